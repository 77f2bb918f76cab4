// expr.cpp -- parser, light canonicaliser and FullForm printer for Symja expression text.
// See expr.hpp for the reference pointers.
#include "expr.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace closyr {

NodeP make(Kind k)
{
    NodeP n(new Node());
    n->kind = k;
    return n;
}

NodeP make_int(int64_t v)
{
    NodeP n = make(Kind::Int);
    n->num = v;
    n->val = (double)v;
    return n;
}

NodeP make_rat(int64_t a, int64_t b)
{
    if (b < 0) { a = -a; b = -b; }
    int64_t g = std::gcd(a < 0 ? -a : a, b);
    if (g > 1) { a /= g; b /= g; }
    if (b == 1) return make_int(a);
    NodeP n = make(Kind::Rat);
    n->num = a;
    n->den = b;
    n->val = (double)a / (double)b;
    return n;
}

NodeP make_real(double v)
{
    NodeP n = make(Kind::Real);
    n->val = v;
    return n;
}

NodeP make_unary(Kind k, NodeP a)
{
    NodeP n = make(k);
    n->args.push_back(std::move(a));
    return n;
}

NodeP make_power(NodeP b, NodeP e)
{
    NodeP n = make(Kind::Power);
    n->args.push_back(std::move(b));
    n->args.push_back(std::move(e));
    return n;
}

NodeP clone(const Node &n)
{
    NodeP c(new Node());
    c->kind = n.kind;
    c->num = n.num;
    c->den = n.den;
    c->big = n.big;
    c->val = n.val;
    c->val2 = n.val2;
    c->name = n.name;
    for (auto &a : n.args) c->args.push_back(clone(*a));
    return c;
}

bool leaf_value(const Node &n, double *out)
{
    switch (n.kind) {
    case Kind::Int: *out = n.val; return true;
    case Kind::Rat: *out = (double)n.num / (double)n.den; return true;
    case Kind::Real: *out = n.val; return true;
    case Kind::E: *out = 2.718281828459045; return true;    // Math.E
    case Kind::Pi: *out = 3.141592653589793; return true;   // Math.PI
    case Kind::DInf: *out = n.val > 0 ? INFINITY : -INFINITY; return true;
    default: return false;
    }
}

bool depends_on_x(const Node &n)
{
    if (n.kind == Kind::X) return true;
    for (auto &a : n.args)
        if (depends_on_x(*a)) return true;
    return false;
}

int leaf_count(const Node &n)
{
    switch (n.kind) {
    case Kind::Rat:
    case Kind::Complex: return 3;
    case Kind::DInf: return n.name == "Infinity" ? 1 : 2;   // symbol vs DirectedInfinity(+-1)
    default: break;
    }
    if (n.args.empty()) return 1;
    int c = 1;
    for (auto &a : n.args) c += leaf_count(*a);
    return c;
}

// ------------------------------------------------------------------------------------------------
// canonicalising constructors (used for infix text only)
// ------------------------------------------------------------------------------------------------
namespace {

struct Coef {                // exact rational, or a double once a Real (or an overflow) is met
    bool exact = true;
    __int128 n = 1, d = 1;
    double r = 1.0;
    void to_real()
    {
        if (exact) { r = (double)n / (double)d; exact = false; }
    }
    void norm()
    {
        if (!exact) return;
        if (d < 0) { n = -n; d = -d; }
        __int128 a = n < 0 ? -n : n, b = d;
        while (b) { __int128 t = a % b; a = b; b = t; }
        if (a > 1) { n /= a; d /= a; }
        const __int128 lim = (__int128)1 << 62;
        if (n > lim || n < -lim || d > lim) to_real();
    }
    void mul(const Node &x)
    {
        if (x.kind == Kind::Real || x.big) { to_real(); r *= x.val; return; }
        if (!exact) { r *= (double)x.num / (double)x.den; return; }
        n *= x.num; d *= x.den; norm();
    }
    void add(const Node &x)
    {
        if (x.kind == Kind::Real || x.big) { to_real(); r += x.val; return; }
        if (!exact) { r += (double)x.num / (double)x.den; return; }
        n = n * x.den + (__int128)x.num * d; d *= x.den; norm();
    }
    NodeP node() const
    {
        if (!exact) return make_real(r);
        return make_rat((int64_t)n, (int64_t)d);
    }
    bool is(int64_t v) const { return exact && d == 1 && n == v; }
};

NodeP c_times(std::vector<NodeP> in);
NodeP c_plus(std::vector<NodeP> in);
NodeP c_power(NodeP b, NodeP e);

NodeP c_neg(NodeP a)
{
    std::vector<NodeP> v;
    v.push_back(make_int(-1));
    v.push_back(std::move(a));
    return c_times(std::move(v));
}

bool is_x_power(const Node &n, int64_t *k)
{
    if (n.kind == Kind::X) { *k = 1; return true; }
    if (n.kind == Kind::Power && n.args[0]->kind == Kind::X && n.args[1]->kind == Kind::Int && !n.args[1]->big) {
        *k = n.args[1]->num;
        return true;
    }
    return false;
}

NodeP c_times(std::vector<NodeP> in)
{
    std::vector<NodeP> flat;
    for (auto &a : in) {
        if (a->kind == Kind::Times)
            for (auto &b : a->args) flat.push_back(std::move(b));
        else
            flat.push_back(std::move(a));
    }
    Coef c;
    std::vector<NodeP> rest;
    int xpos = -1;
    int64_t xk = 0;
    {   // a directed infinity times non-zero numbers is a directed infinity: -1*Infinity -> -Infinity
        int ninf = 0, sign = 1;
        bool only_numbers = true;
        for (auto &a : flat) {
            if (a->kind == Kind::DInf) { ninf++; if (a->val < 0) sign = -sign; }
            else if (a->is_number() && a->val != 0.0) { if (a->val < 0) sign = -sign; }
            else only_numbers = false;
        }
        if (ninf > 0 && only_numbers) {
            NodeP n = make(Kind::DInf);
            n->val = (double)sign;
            return n;
        }
    }
    for (auto &a : flat) {
        int64_t k;
        if (a->is_number()) c.mul(*a);
        else if (is_x_power(*a, &k)) {
            if (xpos < 0) { xpos = (int)rest.size(); rest.push_back(nullptr); }
            xk += k;
        } else rest.push_back(std::move(a));
    }
    if (xpos >= 0) {
        if (xk == 0) rest.erase(rest.begin() + xpos);
        else if (xk == 1) rest[xpos] = make(Kind::X);
        else rest[xpos] = make_power(make(Kind::X), make_int(xk));
    }
    if (c.is(0)) return make_int(0);
    if (rest.empty()) return c.node();
    if (c.is(1)) {
        if (rest.size() == 1) return std::move(rest[0]);
    } else if (c.is(-1) && rest.size() == 1 && rest[0]->kind == Kind::Plus) {
        std::vector<NodeP> terms;                    // Symja distributes a factor of -1 over a sum
        for (auto &t : rest[0]->args) terms.push_back(c_neg(std::move(t)));
        return c_plus(std::move(terms));
    }
    NodeP n = make(Kind::Times);
    if (!c.is(1)) n->args.push_back(c.node());
    for (auto &a : rest) n->args.push_back(std::move(a));
    return n;
}

NodeP c_plus(std::vector<NodeP> in)
{
    std::vector<NodeP> flat;
    for (auto &a : in) {
        if (a->kind == Kind::Plus)
            for (auto &b : a->args) flat.push_back(std::move(b));
        else
            flat.push_back(std::move(a));
    }
    Coef c;
    c.n = 0;
    bool have_num = false;
    std::vector<NodeP> rest;
    {   // a directed infinity absorbs every finite term: x - Infinity -> -Infinity
        int pos = 0, neg = 0;
        bool special = false;
        for (auto &a : flat) {
            if (a->kind == Kind::DInf) (a->val > 0 ? pos : neg)++;
            else if (a->kind == Kind::CInf || a->kind == Kind::Indet) special = true;
        }
        if (!special && (pos > 0) != (neg > 0)) {
            NodeP n = make(Kind::DInf);
            n->val = pos ? 1.0 : -1.0;
            return n;
        }
        if (pos > 0 && neg > 0) return make(Kind::Indet);
    }
    for (auto &a : flat) {
        if (a->is_number()) { c.add(*a); have_num = true; }
        else rest.push_back(std::move(a));
    }
    if (rest.empty()) return c.node();
    if ((!have_num || c.is(0)) && rest.size() == 1) return std::move(rest[0]);
    NodeP n = make(Kind::Plus);
    if (have_num && !c.is(0)) n->args.push_back(c.node());
    for (auto &a : rest) n->args.push_back(std::move(a));
    return n;
}

NodeP c_power(NodeP b, NodeP e)
{
    if (e->kind == Kind::Int && !e->big) {
        int64_t k = e->num;
        if (k == 1) return b;
        if (k == 0) return make_int(1);
        if ((b->kind == Kind::Int || b->kind == Kind::Rat) && !b->big && b->num != 0 && k >= -16 && k <= 16) {
            Coef c;
            Node base;
            base.kind = b->kind;
            base.num = b->num;
            base.den = b->den;
            if (k < 0) { std::swap(base.num, base.den); if (base.den < 0) { base.num = -base.num; base.den = -base.den; } }
            for (int64_t i = 0; i < (k < 0 ? -k : k); i++) c.mul(base);
            return c.node();
        }
        if (b->kind == Kind::Real) return make_real(std::pow(b->val, (double)k));
        if (b->kind == Kind::Power && b->args[1]->is_number()) {     // (u^a)^k -> u^(a*k), k integer
            Coef c;
            c.mul(*b->args[1]);
            c.mul(*e);
            NodeP u = std::move(b->args[0]);
            return c_power(std::move(u), c.node());
        }
        if (b->kind == Kind::Times) {                                 // (f*g)^k -> f^k * g^k
            std::vector<NodeP> fs;
            for (auto &f : b->args) fs.push_back(c_power(std::move(f), make_int(k)));
            return c_times(std::move(fs));
        }
    }
    return make_power(std::move(b), std::move(e));
}

// ------------------------------------------------------------------------------------------------
// recursive-descent parser
// ------------------------------------------------------------------------------------------------
struct Parser {
    const char *s;
    std::string err;
    int depth = 0;

    void ws() { while (*s && isspace((unsigned char)*s)) s++; }
    bool fail(const char *m)
    {
        if (err.empty()) err = std::string(m) + " near '" + std::string(s).substr(0, 16) + "'";
        return false;
    }

    NodeP number()
    {
        const char *b = s;
        bool real = false;
        while (isdigit((unsigned char)*s)) s++;
        if (*s == '.' && (isdigit((unsigned char)s[1]) || s > b)) { real = true; s++; while (isdigit((unsigned char)*s)) s++; }
        if ((*s == 'e' || *s == 'E') && real) {
            const char *r = s + 1;
            if (*r == '-' || *r == '+') r++;
            if (isdigit((unsigned char)*r)) { while (isdigit((unsigned char)*r)) r++; s = r; }
        }
        if (s == b) { fail("number expected"); return nullptr; }
        std::string t(b, s);
        if (real) return make_real(strtod(t.c_str(), nullptr));
        NodeP n = make(Kind::Int);
        n->val = strtod(t.c_str(), nullptr);              // correctly rounded (BigInteger.doubleValue)
        if (t.size() <= 18) n->num = strtoll(t.c_str(), nullptr, 10);
        else { n->big = true; n->name = t; }
        return n;
    }

    NodeP call_or_symbol()
    {
        const char *b = s;
        while (isalnum((unsigned char)*s) || *s == '_' || *s == '$') s++;
        std::string id(b, s);
        ws();
        if (*s != '(') {
            if (id == "x") return make(Kind::X);
            if (id == "E") return make(Kind::E);
            if (id == "Pi") return make(Kind::Pi);
            if (id == "I") { NodeP n = make(Kind::Complex); n->val = 0.0; n->val2 = 1.0; return n; }
            if (id == "Infinity") { NodeP n = make(Kind::DInf); n->val = 1.0; n->name = "Infinity"; return n; }
            if (id == "ComplexInfinity") return make(Kind::CInf);
            if (id == "Indeterminate") return make(Kind::Indet);
            NodeP n = make(Kind::Unknown);
            n->name = id;
            return n;
        }
        s++;
        std::vector<NodeP> args;
        ws();
        if (*s == ')') s++;
        else {
            for (;;) {
                NodeP a = expr();
                if (!a) return nullptr;
                args.push_back(std::move(a));
                ws();
                if (*s == ',') { s++; continue; }
                if (*s == ')') { s++; break; }
                fail("',' or ')' expected");
                return nullptr;
            }
        }
        size_t na = args.size();
        // FullForm heads: kept exactly as written
        if (id == "Plus" && na >= 1) { NodeP n = make(Kind::Plus); n->args = std::move(args); return n; }
        if (id == "Times" && na >= 1) { NodeP n = make(Kind::Times); n->args = std::move(args); return n; }
        if (id == "Power" && na == 2) return make_power(std::move(args[0]), std::move(args[1]));
        if (id == "Rational" && na == 2 && args[0]->kind == Kind::Int && args[1]->kind == Kind::Int &&
            !args[0]->big && !args[1]->big && args[1]->num != 0)
            return make_rat(args[0]->num, args[1]->num);
        if (id == "Complex" && na == 2 && args[0]->is_number() && args[1]->is_number()) {
            NodeP n = make(Kind::Complex);
            leaf_value(*args[0], &n->val);
            leaf_value(*args[1], &n->val2);
            return n;
        }
        if (id == "DirectedInfinity") {
            if (na == 0) return make(Kind::CInf);
            if (na == 1 && args[0]->kind == Kind::Int && (args[0]->num == 1 || args[0]->num == -1)) {
                NodeP n = make(Kind::DInf);
                n->val = (double)args[0]->num;
                return n;
            }
        }
        // convenience heads that `.eval` rewrites
        if (id == "Subtract" && na == 2) {
            std::vector<NodeP> v;
            v.push_back(std::move(args[0]));
            v.push_back(c_neg(std::move(args[1])));
            return c_plus(std::move(v));
        }
        if (id == "Divide" && na == 2) {
            std::vector<NodeP> v;
            v.push_back(std::move(args[0]));
            v.push_back(c_power(std::move(args[1]), make_int(-1)));
            return c_times(std::move(v));
        }
        if (id == "Minus" && na == 1) return c_neg(std::move(args[0]));
        if (id == "Sqrt" && na == 1) return c_power(std::move(args[0]), make_rat(1, 2));
        if (id == "Exp" && na == 1) return c_power(make(Kind::E), std::move(args[0]));
        static const struct { const char *n; Kind k; } U[] = {
            { "Sin", Kind::Sin }, { "Cos", Kind::Cos }, { "Tan", Kind::Tan }, { "Cot", Kind::Cot },
            { "Sec", Kind::Sec }, { "Csc", Kind::Csc }, { "ArcSin", Kind::ArcSin }, { "ArcCos", Kind::ArcCos },
            { "Log", Kind::Log }, { "Abs", Kind::Abs },
        };
        if (na == 1)
            for (auto &u : U)
                if (id == u.n) return make_unary(u.k, std::move(args[0]));
        NodeP n = make(Kind::Unknown);
        n->name = id;
        n->args = std::move(args);
        return n;
    }

    NodeP atom()
    {
        ws();
        if (++depth > 2000) { fail("expression nested too deeply"); return nullptr; }
        NodeP r;
        if (*s == '(') {
            s++;
            r = expr();
            ws();
            if (r && *s != ')') { fail("')' expected"); r = nullptr; }
            else if (r) s++;
        } else if (isdigit((unsigned char)*s) || *s == '.') r = number();
        else if (isalpha((unsigned char)*s) || *s == '$') r = call_or_symbol();
        else fail("unexpected character");
        depth--;
        return r;
    }

    NodeP power()
    {
        NodeP b = atom();
        if (!b) return nullptr;
        ws();
        if (*s == '^') {
            s++;
            NodeP e = unary();               // right associative, allows x^-2
            if (!e) return nullptr;
            return c_power(std::move(b), std::move(e));
        }
        return b;
    }

    NodeP unary()
    {
        ws();
        if (*s == '-') {
            s++;
            NodeP a = unary();
            if (!a) return nullptr;
            return c_neg(std::move(a));
        }
        if (*s == '+') { s++; return unary(); }
        return power();
    }

    NodeP term()
    {
        NodeP a = unary();
        if (!a) return nullptr;
        std::vector<NodeP> fs;
        fs.push_back(std::move(a));
        bool any = false;
        for (;;) {
            ws();
            if (*s == '*') {
                s++;
                NodeP b = unary();
                if (!b) return nullptr;
                fs.push_back(std::move(b));
                any = true;
            } else if (*s == '/') {
                s++;
                NodeP b = unary();
                if (!b) return nullptr;
                fs.push_back(c_power(std::move(b), make_int(-1)));
                any = true;
            } else break;
        }
        if (!any) return std::move(fs[0]);
        return c_times(std::move(fs));
    }

    NodeP expr()
    {
        NodeP a = term();
        if (!a) return nullptr;
        std::vector<NodeP> ts;
        ts.push_back(std::move(a));
        bool any = false;
        for (;;) {
            ws();
            if (*s == '+') {
                s++;
                NodeP b = term();
                if (!b) return nullptr;
                ts.push_back(std::move(b));
                any = true;
            } else if (*s == '-') {
                s++;
                NodeP b = term();
                if (!b) return nullptr;
                ts.push_back(c_neg(std::move(b)));
                any = true;
            } else break;
        }
        if (!any) return std::move(ts[0]);
        return c_plus(std::move(ts));
    }
};

void print_real(std::string &out, double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", v);
    for (int prec = 1; prec < 17; prec++) {               // shortest text that round-trips
        char t[64];
        snprintf(t, sizeof t, "%.*g", prec, v);
        if (strtod(t, nullptr) == v) { memcpy(buf, t, sizeof t); break; }
    }
    std::string t(buf);
    if (t.find('.') == std::string::npos) {
        size_t e = t.find_first_of("eE");
        if (e == std::string::npos) t += ".0";
        else t.insert(e, ".0");
    }
    out += t;
}

void print(std::string &out, const Node &n)
{
    char buf[64];
    switch (n.kind) {
    case Kind::Int:
        if (n.big) out += n.name;
        else { snprintf(buf, sizeof buf, "%lld", (long long)n.num); out += buf; }
        return;
    case Kind::Rat:
        snprintf(buf, sizeof buf, "Rational(%lld,%lld)", (long long)n.num, (long long)n.den);
        out += buf;
        return;
    case Kind::Real:
        if (std::isinf(n.val)) { out += n.val > 0 ? "DirectedInfinity(1)" : "DirectedInfinity(-1)"; return; }
        if (std::isnan(n.val)) { out += "Indeterminate"; return; }
        print_real(out, n.val);
        return;
    case Kind::X: out += "x"; return;
    case Kind::E: out += "E"; return;
    case Kind::Pi: out += "Pi"; return;
    case Kind::Complex:
        out += "Complex(";
        if (std::floor(n.val) == n.val && std::fabs(n.val) < 1e15) { snprintf(buf, sizeof buf, "%lld", (long long)n.val); out += buf; }
        else print_real(out, n.val);
        out += ",";
        if (std::floor(n.val2) == n.val2 && std::fabs(n.val2) < 1e15) { snprintf(buf, sizeof buf, "%lld", (long long)n.val2); out += buf; }
        else print_real(out, n.val2);
        out += ")";
        return;
    case Kind::DInf:
        if (n.name == "Infinity") out += "Infinity";
        else out += n.val > 0 ? "DirectedInfinity(1)" : "DirectedInfinity(-1)";
        return;
    case Kind::CInf: out += "DirectedInfinity()"; return;
    case Kind::Indet: out += "Indeterminate"; return;
    default: break;
    }
    static const char *NAMES[] = { "", "", "", "", "", "", "", "", "", "", "Plus", "Times", "Power", "Sin", "Cos",
                                   "Tan", "Cot", "Sec", "Csc", "ArcSin", "ArcCos", "Log", "Abs" };
    if (n.kind == Kind::Unknown) out += n.name.empty() ? "Unknown" : n.name;
    else out += NAMES[(int)n.kind];
    if (n.kind == Kind::Unknown && n.args.empty()) return;
    out += "(";
    for (size_t i = 0; i < n.args.size(); i++) {
        if (i) out += ",";
        print(out, *n.args[i]);
    }
    out += ")";
}

}  // namespace

NodeP parse_expr(const char *text, std::string *err)
{
    Parser p{ text, {}, 0 };
    NodeP n = p.expr();
    if (n) {
        p.ws();
        if (*p.s) { p.fail("trailing input"); n = nullptr; }
    }
    if (!n && err) *err = p.err.empty() ? "syntax error" : p.err;
    return n;
}

NodeP canon_plus(std::vector<NodeP> terms) { return c_plus(std::move(terms)); }
NodeP canon_times(std::vector<NodeP> factors) { return c_times(std::move(factors)); }
NodeP canon_power(NodeP base, NodeP exponent) { return c_power(std::move(base), std::move(exponent)); }
NodeP canon_neg(NodeP a) { return c_neg(std::move(a)); }

bool same_expr(const Node &a, const Node &b)
{
    if (a.kind != b.kind || a.args.size() != b.args.size()) return false;
    switch (a.kind) {
    case Kind::Int: if (a.big != b.big || (a.big ? a.name != b.name : a.num != b.num)) return false; break;
    case Kind::Rat: if (a.num != b.num || a.den != b.den) return false; break;
    case Kind::Real: if (memcmp(&a.val, &b.val, sizeof(double)) != 0) return false; break;
    case Kind::Complex: if (a.val != b.val || a.val2 != b.val2) return false; break;
    case Kind::DInf: if (a.val != b.val) return false; break;
    case Kind::Unknown: if (a.name != b.name) return false; break;
    default: break;
    }
    for (size_t i = 0; i < a.args.size(); i++)
        if (!same_expr(*a.args[i], *b.args[i])) return false;
    return true;
}

std::string to_fullform(const Node &n)
{
    std::string s;
    print(s, n);
    return s;
}

}  // namespace closyr
