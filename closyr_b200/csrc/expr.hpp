// expr.hpp -- host-side expression trees in Symja's internal (FullForm) shape.
//
// Mirrors what the Clojure side holds as IExpr/IAST (reference: src/closyr/ops/common.clj:135-146 builds
// the canonical expr with `.eval`; src/closyr/ops.clj:127 reads `.leafCount` from it).  The parser accepts
// both IExpr.fullFormString() text and the printed form (str expr) that the reference logs and its tests
// compare against (test/closyr/ops_modify_test.clj:463-751).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace closyr {

enum class Kind : uint8_t {
    Int, Rat, Real,          // exact integer, exact rational, machine real literal
    X, E, Pi,                // the variable and the two real symbolic constants
    Complex,                 // Complex(re, im) literal (I = Complex(0,1))
    DInf,                    // DirectedInfinity(+-1)
    CInf,                    // DirectedInfinity()  (ComplexInfinity)
    Indet,                   // Indeterminate
    Plus, Times, Power,
    Sin, Cos, Tan, Cot, Sec, Csc, ArcSin, ArcCos, Log, Abs,
    Unknown                  // any other head or symbol: stays symbolic in Symja
};

struct Node;
using NodeP = std::unique_ptr<Node>;

struct Node {
    Kind kind = Kind::Unknown;
    // Int: num (exact when !big); Rat: num/den; Real: val; Complex: val + i*val2; DInf: val = +-1
    int64_t num = 0, den = 1;
    bool big = false;        // integer literal does not fit int64: only `val` is meaningful
    double val = 0.0, val2 = 0.0;
    std::string name;        // Unknown heads / symbols, and the decimal text of big integers
    std::vector<NodeP> args;

    bool is_ast() const { return kind >= Kind::Plus && (kind != Kind::Unknown || !args.empty()); }
    bool is_number() const { return kind == Kind::Int || kind == Kind::Rat || kind == Kind::Real; }
};

NodeP make(Kind k);
NodeP make_int(int64_t v);
NodeP make_rat(int64_t n, int64_t d);     // normalises sign and gcd; returns Int when d == 1
NodeP make_real(double v);
NodeP make_unary(Kind k, NodeP a);
NodeP make_power(NodeP b, NodeP e);
NodeP clone(const Node &n);

// numeric value of an Int/Rat/Real/E/Pi/DInf leaf, as the kernels see it (double(num)/double(den))
bool leaf_value(const Node &n, double *out);
bool depends_on_x(const Node &n);
// Symja leafCount: atoms 1, Rational and Complex 3, AST = 1 + sum(args)  
int leaf_count(const Node &n);

// Parses Symja text.  FullForm heads (Plus, Times, Power, Rational, Complex, DirectedInfinity ...) are kept
// exactly as written; infix text is canonicalised the way `->phenotype` would for the shapes the
// reference prints (Subtract/Divide/Sqrt/Exp lowering, flattening, exact rational folding).
// Returns nullptr and fills err on a syntax error.
NodeP parse_expr(const char *text, std::string *err);

std::string to_fullform(const Node &n);

// Canonicalising constructors: what `.eval` in `->phenotype` (common.clj:135-146) does to the shapes the GA
// operators build -- flattening, exact rational folding, x^a*x^b, (u^a)^k, -1*(a+b).  The infix parser uses them.
NodeP canon_plus(std::vector<NodeP> terms);
NodeP canon_times(std::vector<NodeP> factors);
NodeP canon_power(NodeP base, NodeP exponent);
NodeP canon_neg(NodeP a);
// structural equality (same tree, same literals)
bool same_expr(const Node &a, const Node &b);

}  // namespace closyr
