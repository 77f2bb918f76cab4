// ga.cpp -- expression rewrites of the GA (mutation operators, crossover, derivative) and the evolve loop.
// Reference pointers in ga.hpp.  Host only; the scorer is plugged in through the population builder: the
// driver hands out the unscored phenotypes as a compiled population and takes their scores back, exactly the
// role of the `score-fn` closure in ga/initialize (src/closyr/ga.clj:11-17).
#include "ga.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <thread>

#include "../../include/closyr_b200.h"
#include "builder.hpp"
#include "compiler.hpp"
#include "internal.hpp"

namespace closyr {
namespace {

// ------------------------------------------------------------------------------------------------
// constructors that evaluate the way `->phenotype`'s `.eval` would (approximately: see ga.hpp)
// ------------------------------------------------------------------------------------------------
thread_local bool t_div_by_zero = false;      // "Infinite expression 1/0 encountered." (modify.clj:232-234)

NodeP X() { return make(Kind::X); }
NodeP I(int64_t v) { return make_int(v); }
NodeP Q(int64_t n, int64_t d) { return make_rat(n, d); }
NodeP R(double v) { return make_real(v); }

bool is_zero(const Node &n) { return n.is_number() && !n.big && (n.kind == Kind::Real ? n.val == 0.0 : n.num == 0); }
bool is_int(const Node &n, int64_t v) { return n.kind == Kind::Int && !n.big && n.num == v; }
bool is_negative_number(const Node &n) { return n.is_number() && n.val < 0.0; }

NodeP g_plus(std::vector<NodeP> terms);
NodeP g_times(std::vector<NodeP> factors);
NodeP g_power(NodeP b, NodeP e);

NodeP num_add(const Node &a, const Node &b)
{
    std::vector<NodeP> v;
    v.push_back(clone(a));
    v.push_back(clone(b));
    return canon_plus(std::move(v));
}

NodeP plus2(NodeP a, NodeP b)
{
    std::vector<NodeP> v;
    v.push_back(std::move(a));
    v.push_back(std::move(b));
    return g_plus(std::move(v));
}
NodeP times2(NodeP a, NodeP b)
{
    std::vector<NodeP> v;
    v.push_back(std::move(a));
    v.push_back(std::move(b));
    return g_times(std::move(v));
}
NodeP neg(NodeP a) { return times2(I(-1), std::move(a)); }
NodeP sub2(NodeP a, NodeP b) { return plus2(std::move(a), neg(std::move(b))); }          // F/Subtract
NodeP div2(NodeP a, NodeP b) { return times2(std::move(a), g_power(std::move(b), I(-1))); }  // F/Divide

// a + a -> 2*a, 2*a + 3*a -> 5*a
NodeP g_plus(std::vector<NodeP> terms)
{
    for (auto &t : terms)
        if (t->kind == Kind::Indet) return make(Kind::Indet);     // Indeterminate absorbs
    NodeP s = canon_plus(std::move(terms));
    if (s->kind != Kind::Plus) return s;
    struct Term { NodeP coef, body; };
    std::vector<Term> ts;
    for (auto &t : s->args) {
        Term cur;
        if (t->is_number()) { cur.coef = std::move(t); cur.body = nullptr; }
        else if (t->kind == Kind::Times && t->args.size() >= 2 && t->args[0]->is_number()) {
            cur.coef = std::move(t->args[0]);
            if (t->args.size() == 2) cur.body = std::move(t->args[1]);
            else {
                cur.body = make(Kind::Times);
                for (size_t i = 1; i < t->args.size(); i++) cur.body->args.push_back(std::move(t->args[i]));
            }
        } else { cur.coef = I(1); cur.body = std::move(t); }
        bool found = false;
        if (cur.body)
            for (auto &o : ts)
                if (o.body && same_expr(*o.body, *cur.body)) {
                    o.coef = num_add(*o.coef, *cur.coef);
                    found = true;
                    break;
                }
        if (!found) ts.push_back(std::move(cur));
    }
    std::vector<NodeP> out;
    for (auto &t : ts) {
        if (!t.body) { out.push_back(std::move(t.coef)); continue; }
        std::vector<NodeP> f;
        f.push_back(std::move(t.coef));
        f.push_back(std::move(t.body));
        out.push_back(canon_times(std::move(f)));
    }
    return canon_plus(std::move(out));
}

// a*a -> a^2, a^(1/2)*a -> a^(3/2), E^u*E^v -> E^(u+v)
NodeP g_times(std::vector<NodeP> factors)
{
    for (auto &f : factors)
        if (f->kind == Kind::Indet) return make(Kind::Indet);
    NodeP p = canon_times(std::move(factors));
    if (p->kind != Kind::Times) return p;
    struct Fac { NodeP base, exp; };
    std::vector<Fac> fs;
    std::vector<NodeP> nums;
    bool merged = false;
    for (auto &f : p->args) {
        if (f->is_number()) { nums.push_back(std::move(f)); continue; }
        Fac cur;
        if (f->kind == Kind::Power) { cur.base = std::move(f->args[0]); cur.exp = std::move(f->args[1]); }
        else { cur.base = std::move(f); cur.exp = I(1); }
        bool found = false;
        for (auto &o : fs)
            if (same_expr(*o.base, *cur.base)) {
                o.exp = plus2(std::move(o.exp), std::move(cur.exp));
                found = merged = true;
                break;
            }
        if (!found) fs.push_back(std::move(cur));
    }
    std::vector<NodeP> out;
    if (nums.size() == 1 && is_int(*nums[0], -1) && fs.size() >= 2)        // -(a-b)*f -> (b-a)*f
        for (auto &f : fs)
            if (is_int(*f.exp, 1) && f.base->kind == Kind::Plus && !f.base->args.empty() &&
                f.base->args[0]->is_number() && f.base->args[0]->val < 0.0) {
                std::vector<NodeP> terms;
                for (auto &t : f.base->args) terms.push_back(canon_neg(std::move(t)));
                f.base = canon_plus(std::move(terms));
                nums.clear();
                break;
            }
    for (auto &n : nums) out.push_back(std::move(n));
    for (auto &f : fs) {
        if (is_int(*f.exp, 1)) out.push_back(std::move(f.base));
        else if (merged) out.push_back(g_power(std::move(f.base), std::move(f.exp)));
        else out.push_back(make_power(std::move(f.base), std::move(f.exp)));
    }
    return canon_times(std::move(out));
}

Kind reciprocal_fn(Kind k)
{
    switch (k) {
    case Kind::Sin: return Kind::Csc;
    case Kind::Cos: return Kind::Sec;
    case Kind::Tan: return Kind::Cot;
    case Kind::Csc: return Kind::Sin;
    case Kind::Sec: return Kind::Cos;
    case Kind::Cot: return Kind::Tan;
    default: return Kind::Unknown;
    }
}

// Symja writes odd/even functions of a "negative" argument with the sign pulled out: Sin(-1/2+x) is
// -Sin(1/2-x), Cos(-x) is Cos(x), (-1+x)^2 is (1-x)^2 (visible in ops_modify_test.clj:463-751)
bool negative_signed(const Node &n)
{
    if (n.is_number()) return n.val < 0.0;
    if (n.kind == Kind::Times) return !n.args.empty() && n.args[0]->is_number() && n.args[0]->val < 0.0;
    return false;
}
bool negative_leading(const Node &n)
{
    if (n.kind == Kind::Plus) return !n.args.empty() && negative_signed(*n.args[0]);
    return negative_signed(n);
}

bool perfect_root(int64_t v, int64_t *r)
{
    if (v < 0) return false;
    int64_t q = (int64_t)std::llround(std::sqrt((double)v));
    for (int64_t c = std::max<int64_t>(0, q - 1); c <= q + 1; c++)
        if (c * c == v) { *r = c; return true; }
    return false;
}

NodeP g_power(NodeP b, NodeP e)
{
    if (is_zero(*b) && is_negative_number(*e)) {               // 0^-k: Symja throws, the operator fails
        t_div_by_zero = true;
        return I(0);
    }
    if (is_int(*e, -1) && reciprocal_fn(b->kind) != Kind::Unknown) {   // 1/Sin(u) -> Csc(u) (Appendix A1)
        b->kind = reciprocal_fn(b->kind);
        return b;
    }
    if (b->kind == Kind::E && e->kind == Kind::Log) return std::move(e->args[0]);      // E^Log(u) -> u
    if (b->kind == Kind::Real && e->is_number() && b->val > 0.0) {
        double ev;
        leaf_value(*e, &ev);
        const double v = std::pow(b->val, ev);
        if (std::isfinite(v)) return R(v);
    }
    if (b->kind == Kind::E && e->kind == Kind::Real) {
        const double v = std::exp(e->val);
        if (std::isfinite(v)) return R(v);
    }
    const bool e_int = e->kind == Kind::Int && !e->big;
    if (e_int && (e->num % 2) == 0 && e->num != 0 && negative_leading(*b)) b = neg(std::move(b));   // even power
    if (b->kind == Kind::Times && e_int && e->num != 1 && e->num != 0) {          // (f*g)^k -> f^k*g^k
        std::vector<NodeP> fs;
        for (auto &f : b->args) fs.push_back(g_power(std::move(f), I(e->num)));
        return g_times(std::move(fs));
    }
    if (b->kind == Kind::Times && e->kind == Kind::Rat && b->args.size() >= 2 && b->args[0]->is_number() &&
        b->args[0]->val > 0.0) {                                                  // (c*u)^a -> c^a*u^a, c > 0
        NodeP c = std::move(b->args[0]);
        b->args.erase(b->args.begin());
        NodeP rest = b->args.size() == 1 ? std::move(b->args[0]) : std::move(b);
        NodeP ce = g_power(std::move(c), clone(*e));
        return times2(std::move(ce), g_power(std::move(rest), std::move(e)));
    }
    if (b->kind == Kind::Power && (b->args[1]->kind == Kind::Rat || b->args[1]->kind == Kind::Int) && e->is_number() &&
        b->args[1]->val > -1.0 && b->args[1]->val <= 1.0) {                       // (u^a)^b -> u^(a*b), -1 < a <= 1
        NodeP a = std::move(b->args[1]);
        NodeP u = std::move(b->args[0]);
        return g_power(std::move(u), times2(std::move(a), std::move(e)));
    }
    if (b->kind == Kind::Rat && b->num == 1 && e->kind == Kind::Rat && e->num < 0)           // (1/q)^(-a) -> q^a
        return g_power(I(b->den), Q(-e->num, e->den));
    if ((b->kind == Kind::Int || b->kind == Kind::Rat) && !b->big && e->kind == Kind::Rat && e->den == 2 && b->val > 0.0) {
        int64_t rn, rd;                                                           // exact square roots: 100^(-1/2) -> 1/10
        if (perfect_root(b->num, &rn) && perfect_root(b->den, &rd) && rn != 0)
            return canon_power(Q(rn, rd), I(e->num));
    }
    return canon_power(std::move(b), std::move(e));
}

NodeP g_fn(Kind k, NodeP u)
{
    if (u->kind == Kind::Real) {             // machine numbers are evaluated; exact arguments stay symbolic
        const double x = u->val;
        double v = NAN;
        switch (k) {
        case Kind::Sin: v = std::sin(x); break;
        case Kind::Cos: v = std::cos(x); break;
        case Kind::Tan: v = std::tan(x); break;
        case Kind::Cot: v = 1.0 / std::tan(x); break;
        case Kind::Sec: v = 1.0 / std::cos(x); break;
        case Kind::Csc: v = 1.0 / std::sin(x); break;
        case Kind::ArcSin: if (std::fabs(x) <= 1.0) v = std::asin(x); break;
        case Kind::ArcCos: if (std::fabs(x) <= 1.0) v = std::acos(x); break;
        case Kind::Log: if (x > 0.0) v = std::log(x); break;
        case Kind::Abs: v = std::fabs(x); break;
        default: break;
        }
        if (std::isfinite(v)) return R(v);
    }
    if (is_int(*u, 0)) {
        if (k == Kind::Sin || k == Kind::Tan || k == Kind::ArcSin || k == Kind::Abs) return I(0);
        if (k == Kind::Cos || k == Kind::Sec) return I(1);
    }
    if (k == Kind::Log && u->kind == Kind::E) return I(1);
    if (k == Kind::Log && is_int(*u, 1)) return I(0);
    if ((k == Kind::Sin && u->kind == Kind::ArcSin) || (k == Kind::Cos && u->kind == Kind::ArcCos))
        return std::move(u->args[0]);
    if ((k == Kind::Cos && u->kind == Kind::ArcSin) || (k == Kind::Sin && u->kind == Kind::ArcCos))
        return g_power(plus2(I(1), neg(g_power(std::move(u->args[0]), I(2)))), Q(1, 2));      // Sqrt(1-u^2)
    if (negative_leading(*u)) {
        switch (k) {
        case Kind::Sin: case Kind::Tan: case Kind::Cot: case Kind::Csc: case Kind::ArcSin:
            return neg(make_unary(k, neg(std::move(u))));
        case Kind::Cos: case Kind::Sec: case Kind::Abs:
            return make_unary(k, neg(std::move(u)));
        default: break;
        }
    }
    return make_unary(k, std::move(u));
}

NodeP g_exp(NodeP u) { return g_power(make(Kind::E), std::move(u)); }           // F/Exp
NodeP g_sqrt(NodeP u) { return g_power(std::move(u), Q(1, 2)); }                // F/Sqrt

NodeP rebuild(const Node &proto, Kind k, std::vector<NodeP> args)
{
    switch (k) {
    case Kind::Plus: return g_plus(std::move(args));
    case Kind::Times: return g_times(std::move(args));
    case Kind::Power:
        if (args.size() == 2) return g_power(std::move(args[0]), std::move(args[1]));
        break;
    case Kind::Unknown: break;
    default:
        if (args.size() == 1) return g_fn(k, std::move(args[0]));
        break;
    }
    NodeP n = make(k);
    n->name = proto.name;
    n->args = std::move(args);
    return n;
}

// bottom-up rebuild through the evaluating constructors: what `.eval` does to a tree that was put together by hand
NodeP recanon(const Node &n)
{
    if (!n.is_ast()) return clone(n);
    std::vector<NodeP> args;
    for (auto &a : n.args) args.push_back(recanon(*a));
    return rebuild(n, n.kind, std::move(args));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// F/D(expr, x)
// ------------------------------------------------------------------------------------------------
NodeP derivative(const Node &e)
{
    switch (e.kind) {
    case Kind::X: return I(1);
    case Kind::Plus: {
        std::vector<NodeP> ts;
        for (auto &a : e.args) ts.push_back(derivative(*a));
        return g_plus(std::move(ts));
    }
    case Kind::Times: {
        std::vector<NodeP> ts;
        for (size_t i = 0; i < e.args.size(); i++) {
            if (!depends_on_x(*e.args[i])) continue;
            std::vector<NodeP> fs;
            for (size_t j = 0; j < e.args.size(); j++) fs.push_back(j == i ? derivative(*e.args[j]) : clone(*e.args[j]));
            ts.push_back(g_times(std::move(fs)));
        }
        if (ts.empty()) return I(0);
        return g_plus(std::move(ts));
    }
    case Kind::Power: {
        const Node &b = *e.args[0], &x = *e.args[1];
        const bool bx = depends_on_x(b), ex = depends_on_x(x);
        if (!bx && !ex) return I(0);
        if (!ex) {                                              // u^c -> c*u^(c-1)*u'
            std::vector<NodeP> fs;
            fs.push_back(clone(x));
            fs.push_back(g_power(clone(b), plus2(clone(x), I(-1))));
            fs.push_back(derivative(b));
            return g_times(std::move(fs));
        }
        if (!bx) {                                              // c^v -> c^v*Log(c)*v'
            std::vector<NodeP> fs;
            fs.push_back(clone(e));
            if (b.kind != Kind::E) fs.push_back(g_fn(Kind::Log, clone(b)));
            fs.push_back(derivative(x));
            return g_times(std::move(fs));
        }
        // u^v -> u^v*(v'*Log(u) + v*u'/u)
        std::vector<NodeP> q;
        q.push_back(clone(x));
        q.push_back(derivative(b));
        q.push_back(g_power(clone(b), I(-1)));
        return times2(clone(e), plus2(times2(derivative(x), g_fn(Kind::Log, clone(b))), g_times(std::move(q))));
    }
    default: break;
    }
    if (e.args.size() != 1 || e.kind == Kind::Unknown) {
        if (!depends_on_x(e)) return I(0);
        NodeP n = make(Kind::Unknown);                          // stays symbolic: D(f(x), x)
        n->name = "D";
        n->args.push_back(clone(e));
        n->args.push_back(X());
        return n;
    }
    const Node &u = *e.args[0];
    if (!depends_on_x(u)) return I(0);
    NodeP du = derivative(u);
    auto one_minus_sq = [&]() { return plus2(I(1), neg(g_power(clone(u), I(2)))); };
    NodeP outer;
    switch (e.kind) {
    case Kind::Sin: outer = g_fn(Kind::Cos, clone(u)); break;
    case Kind::Cos: outer = neg(g_fn(Kind::Sin, clone(u))); break;
    case Kind::Tan: outer = g_power(g_fn(Kind::Sec, clone(u)), I(2)); break;
    case Kind::Cot: outer = neg(g_power(g_fn(Kind::Csc, clone(u)), I(2))); break;
    case Kind::Sec: outer = times2(g_fn(Kind::Sec, clone(u)), g_fn(Kind::Tan, clone(u))); break;
    case Kind::Csc: outer = neg(times2(g_fn(Kind::Cot, clone(u)), g_fn(Kind::Csc, clone(u)))); break;
    case Kind::ArcSin: outer = g_power(one_minus_sq(), Q(-1, 2)); break;
    case Kind::ArcCos: outer = neg(g_power(one_minus_sq(), Q(-1, 2))); break;
    case Kind::Log: outer = g_power(clone(u), I(-1)); break;
    case Kind::Abs: outer = times2(clone(u), g_power(g_fn(Kind::Abs, clone(u)), I(-1))); break;
    default: outer = I(0); break;
    }
    return times2(std::move(outer), std::move(du));
}

// ------------------------------------------------------------------------------------------------
// the 96 mutation operators (initialize.clj:35-665, same order)
// ------------------------------------------------------------------------------------------------
namespace {

enum Class { FN, LEAF_SYM, LEAF_NUM, SUBST, HEAD, BRANCH };
using Rewrite = std::function<NodeP(NodeP)>;
struct Mutation {
    Class cls;
    const char *label;
    Rewrite f;              // FN: whole expression; LEAF_*: the leaf; BRANCH: the rebuilt branch
    Kind from, to;          // SUBST / HEAD
};

const char *class_name(Class c)
{
    switch (c) {
    case FN: return "modify-fn";
    case LEAF_SYM:
    case LEAF_NUM: return "modify-leafs";
    case SUBST: return "modify-substitute";
    case HEAD: return "modify-ast-head";
    default: return "modify-branches";
    }
}

std::vector<Mutation> build_table()
{
    std::vector<Mutation> t;
    auto add = [&](Class c, const char *label, Rewrite f) { t.push_back({ c, label, std::move(f), Kind::Unknown, Kind::Unknown }); };
    auto heads = [&](Class c, const char *label, Kind from, Kind to) { t.push_back({ c, label, nullptr, from, to }); };
    auto fx = [](Kind k) { return g_fn(k, X()); };

    add(FN, "Derivative", [](NodeP e) { return derivative(*e); });
    add(FN, "+1/2", [](NodeP e) { return plus2(std::move(e), Q(1, 2)); });
    add(FN, "-1/2", [](NodeP e) { return sub2(std::move(e), Q(1, 2)); });
    add(FN, "+1/10", [](NodeP e) { return plus2(std::move(e), Q(1, 10)); });
    add(FN, "-1/10", [](NodeP e) { return sub2(std::move(e), Q(1, 10)); });
    add(FN, "+1/100", [](NodeP e) { return plus2(std::move(e), Q(1, 100)); });
    add(FN, "-1/100", [](NodeP e) { return sub2(std::move(e), Q(1, 100)); });
    add(FN, "+Sin", [fx](NodeP e) { return plus2(std::move(e), fx(Kind::Sin)); });
    add(FN, "-Sin", [fx](NodeP e) { return sub2(std::move(e), fx(Kind::Sin)); });
    add(FN, "+Log", [fx](NodeP e) { return plus2(std::move(e), fx(Kind::Log)); });
    add(FN, "-Log", [fx](NodeP e) { return sub2(std::move(e), fx(Kind::Log)); });
    add(FN, "+Exp", [](NodeP e) { return plus2(std::move(e), g_exp(X())); });
    add(FN, "-Exp", [](NodeP e) { return sub2(std::move(e), g_exp(X())); });
    add(FN, "+Cos", [fx](NodeP e) { return plus2(std::move(e), fx(Kind::Cos)); });
    add(FN, "-Cos", [fx](NodeP e) { return sub2(std::move(e), fx(Kind::Cos)); });
    add(FN, "*Sin", [fx](NodeP e) { return times2(std::move(e), fx(Kind::Sin)); });
    add(FN, "/Sin", [fx](NodeP e) { return times2(std::move(e), div2(I(1), fx(Kind::Sin))); });
    add(FN, "*Cos", [fx](NodeP e) { return times2(std::move(e), fx(Kind::Cos)); });
    add(FN, "/Cos", [fx](NodeP e) { return times2(std::move(e), div2(I(1), fx(Kind::Cos))); });
    add(FN, "+x", [](NodeP e) { return plus2(std::move(e), X()); });
    add(FN, "-x", [](NodeP e) { return sub2(std::move(e), X()); });
    add(FN, "+x^2", [](NodeP e) { return plus2(std::move(e), g_power(X(), I(2))); });
    add(FN, "-x^2", [](NodeP e) { return sub2(std::move(e), g_power(X(), I(2))); });
    add(FN, "+x^1/2", [](NodeP e) { return plus2(std::move(e), g_sqrt(X())); });
    add(FN, "-x^1/2", [](NodeP e) { return sub2(std::move(e), g_sqrt(X())); });
    add(FN, "*x", [](NodeP e) { return times2(std::move(e), X()); });
    add(FN, "/x", [](NodeP e) { return div2(std::move(e), X()); });
    add(FN, "1/f", [](NodeP e) { return div2(I(1), std::move(e)); });
    add(FN, "*-1", [](NodeP e) { return times2(std::move(e), I(-1)); });
    add(FN, "/2", [](NodeP e) { return times2(std::move(e), Q(1, 2)); });
    add(FN, "*2", [](NodeP e) { return times2(std::move(e), I(2)); });
    add(FN, "/10", [](NodeP e) { return div2(std::move(e), I(10)); });
    add(FN, "*10", [](NodeP e) { return times2(std::move(e), I(10)); });
    add(FN, "/100", [](NodeP e) { return times2(std::move(e), Q(1, 100)); });
    add(FN, "*100", [](NodeP e) { return times2(std::move(e), I(100)); });
    add(FN, "*1.1", [](NodeP e) { return times2(std::move(e), R(1.1)); });
    add(FN, "*0.9", [](NodeP e) { return times2(std::move(e), R(0.9)); });

    add(LEAF_SYM, "x+1/2", [](NodeP l) { return plus2(std::move(l), Q(1, 2)); });
    add(LEAF_SYM, "x-1/2", [](NodeP l) { return sub2(std::move(l), Q(1, 2)); });
    add(LEAF_SYM, "x/10", [](NodeP l) { return div2(std::move(l), I(10)); });
    add(LEAF_SYM, "10*x", [](NodeP l) { return times2(std::move(l), I(10)); });
    add(LEAF_SYM, "1/x", [](NodeP l) { return div2(I(1), std::move(l)); });
    add(LEAF_SYM, "x/100", [](NodeP l) { return div2(std::move(l), I(100)); });
    add(LEAF_SYM, "100*x", [](NodeP l) { return times2(std::move(l), I(100)); });
    add(LEAF_SYM, "-1*x", [](NodeP l) { return times2(std::move(l), I(-1)); });
    add(LEAF_SYM, "1.1*x", [](NodeP l) { return times2(std::move(l), R(1.1)); });
    add(LEAF_SYM, "0.9*x", [](NodeP l) { return times2(std::move(l), R(0.9)); });
    add(LEAF_SYM, "sin(x)", [fx](NodeP) { return fx(Kind::Sin); });         // the leaf is replaced by f(x-sym)
    add(LEAF_SYM, "cos(x)", [fx](NodeP) { return fx(Kind::Cos); });
    add(LEAF_SYM, "asin(x)", [fx](NodeP) { return fx(Kind::ArcSin); });
    add(LEAF_SYM, "acos(x)", [fx](NodeP) { return fx(Kind::ArcCos); });
    add(LEAF_SYM, "log(x)", [fx](NodeP) { return fx(Kind::Log); });
    add(LEAF_SYM, "exp(x)", [](NodeP) { return g_exp(X()); });
    add(LEAF_SYM, "x^1/2", [](NodeP l) { return g_power(std::move(l), Q(1, 2)); });
    add(LEAF_SYM, "x^2", [](NodeP l) { return g_power(std::move(l), I(2)); });
    add(LEAF_SYM, "x+1/10", [](NodeP l) { return plus2(std::move(l), Q(1, 10)); });
    add(LEAF_SYM, "x-1/10", [](NodeP l) { return sub2(std::move(l), Q(1, 10)); });
    add(LEAF_SYM, "x+1/100", [](NodeP l) { return plus2(std::move(l), Q(1, 100)); });
    add(LEAF_SYM, "x-1/100", [](NodeP l) { return sub2(std::move(l), Q(1, 100)); });

    add(LEAF_NUM, "c/2", [](NodeP l) { return times2(std::move(l), Q(1, 2)); });
    add(LEAF_NUM, "c*2", [](NodeP l) { return times2(std::move(l), I(2)); });
    add(LEAF_NUM, "c*-1", [](NodeP l) { return times2(std::move(l), I(-1)); });
    add(LEAF_NUM, "c/10", [](NodeP l) { return times2(std::move(l), Q(1, 10)); });
    add(LEAF_NUM, "c*10", [](NodeP l) { return times2(std::move(l), I(10)); });
    add(LEAF_NUM, "c+1/10", [](NodeP l) { return plus2(std::move(l), Q(1, 10)); });
    add(LEAF_NUM, "c-1/10", [](NodeP l) { return sub2(std::move(l), Q(1, 10)); });
    add(LEAF_NUM, "1/c", [](NodeP l) { return div2(I(1), std::move(l)); });
    add(LEAF_NUM, "c+1/100", [](NodeP l) { return plus2(std::move(l), Q(1, 100)); });
    add(LEAF_NUM, "c-1/100", [](NodeP l) { return sub2(std::move(l), Q(1, 100)); });
    add(LEAF_NUM, "c+1/2", [](NodeP l) { return plus2(std::move(l), Q(1, 2)); });
    add(LEAF_NUM, "c-1/2", [](NodeP l) { return sub2(std::move(l), Q(1, 2)); });

    heads(SUBST, "Sin->Cos", Kind::Sin, Kind::Cos);
    heads(SUBST, "Cos->Sin", Kind::Cos, Kind::Sin);
    heads(HEAD, "sin->cos", Kind::Sin, Kind::Cos);
    heads(HEAD, "cos->sin", Kind::Cos, Kind::Sin);
    heads(HEAD, "sin->asin", Kind::Sin, Kind::ArcSin);
    heads(HEAD, "cos->acos", Kind::Cos, Kind::ArcCos);
    heads(HEAD, "+->*", Kind::Plus, Kind::Times);
    heads(HEAD, "*->+", Kind::Times, Kind::Plus);
    heads(HEAD, "^->*", Kind::Power, Kind::Times);

    add(BRANCH, "b derivative", [](NodeP b) { return derivative(*b); });
    add(BRANCH, "b sin", [](NodeP b) { return g_fn(Kind::Sin, std::move(b)); });
    add(BRANCH, "b cos", [](NodeP b) { return g_fn(Kind::Cos, std::move(b)); });
    add(BRANCH, "b asin", [](NodeP b) { return g_fn(Kind::ArcSin, std::move(b)); });
    add(BRANCH, "b acos", [](NodeP b) { return g_fn(Kind::ArcCos, std::move(b)); });
    add(BRANCH, "b exp", [](NodeP b) { return g_exp(std::move(b)); });
    add(BRANCH, "b log", [](NodeP b) { return g_fn(Kind::Log, std::move(b)); });
    add(BRANCH, "b*b", [](NodeP b) { NodeP c = clone(*b); return times2(std::move(b), std::move(c)); });
    add(BRANCH, "b^1/2", [](NodeP b) { return g_power(std::move(b), Q(1, 2)); });
    add(BRANCH, "b^-2", [](NodeP b) { return g_power(std::move(b), I(-2)); });
    add(BRANCH, "b^-1", [](NodeP b) { return g_power(std::move(b), I(-1)); });
    add(BRANCH, "b*-1", [](NodeP b) { return times2(std::move(b), I(-1)); });
    add(BRANCH, "b*1.1", [](NodeP b) { return times2(std::move(b), R(1.1)); });
    add(BRANCH, "b*0.9", [](NodeP b) { return times2(std::move(b), R(0.9)); });
    add(BRANCH, "b+0.1", [](NodeP b) { return plus2(std::move(b), R(0.1)); });
    add(BRANCH, "b-0.1", [](NodeP b) { return sub2(std::move(b), R(0.1)); });
    return t;
}

const std::vector<Mutation> &table()
{
    static const std::vector<Mutation> t = build_table();
    return t;
}

struct Walk {
    const Mutation &m;
    JavaRandom *rng;
    double leaf_scalar;           // min(1, max(0.005, 1/leafCount))  (common.clj:149-156)

    bool should(double scalar)
    {
        const double r = rng ? rng->rand() : 0.0;
        return r < scalar * leaf_scalar;
    }

    NodeP leafs(const Node &n)                                   // modify.clj:27-34
    {
        if (n.is_ast()) {
            std::vector<NodeP> args;
            for (auto &a : n.args) args.push_back(leafs(*a));
            return rebuild(n, n.kind, std::move(args));
        }
        const bool sym = n.kind == Kind::X || n.kind == Kind::E || n.kind == Kind::Pi ||
                         (n.kind == Kind::Unknown && n.args.empty());
        const bool match = m.cls == LEAF_SYM ? sym : n.is_number();
        if (match && should(1.5)) return m.f(clone(n));          // the type test short-circuits the draw
        return clone(n);
    }

    NodeP branches(const Node &n)                                // modify.clj:37-46: children first, then the branch
    {
        if (!n.is_ast()) return clone(n);
        std::vector<NodeP> args;
        for (auto &a : n.args) args.push_back(branches(*a));
        NodeP b = rebuild(n, n.kind, std::move(args));
        if (should(0.5)) return m.f(std::move(b));
        return b;
    }

    NodeP ast_heads(const Node &n, bool always)                  // modify.clj:49-56 / .subs for SUBST
    {
        if (!n.is_ast()) return clone(n);
        std::vector<NodeP> args;
        for (auto &a : n.args) args.push_back(ast_heads(*a, always));
        Kind k = n.kind;
        const bool yes = always || should(0.25);                 // the draw comes before the head test
        if (yes && k == m.from) k = m.to;
        return rebuild(n, k, std::move(args));
    }
};

// common.clj:127-132: NIL or a builtin symbol that is not a real number becomes x
NodeP valid_or_default(NodeP e)
{
    if (!e) return X();
    switch (e->kind) {
    case Kind::DInf:
    case Kind::CInf:
    case Kind::Indet: return X();
    case Kind::Unknown: if (e->args.empty()) return X(); break;
    default: break;
    }
    return e;
}

// modify.clj:259-288
int mods_count_for(int idx)
{
    static const int W[15] = { 40, 30, 20, 15, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1 };
    for (int k = 0; k < 15; k++) {
        if (idx < W[k]) return k + 1;
        idx -= W[k];
    }
    return 15;
}
const int MODS_SAMPLER_SIZE = 172;

}  // namespace

NodeP canonical_form(const Node &n) { return recanon(n); }

int mutation_count() { return (int)table().size(); }

MutationInfo mutation_info(int i)
{
    const Mutation &m = table()[(size_t)i];
    return { class_name(m.cls), m.label };
}

NodeP apply_mutation(int i, const Node &expr, JavaRandom *rng)
{
    if (i < 0 || i >= mutation_count()) return nullptr;
    const Mutation &m = table()[(size_t)i];
    const int lc = leaf_count(expr);
    Walk w{ m, rng, std::min(1.0, std::max(0.005, 1.0 / (double)lc)) };
    t_div_by_zero = false;
    NodeP out;
    switch (m.cls) {
    case FN: out = m.f(clone(expr)); break;
    case LEAF_SYM:
    case LEAF_NUM: out = w.leafs(expr); break;
    case SUBST: out = w.ast_heads(expr, true); break;
    case HEAD: out = w.ast_heads(expr, false); break;
    default: out = w.branches(expr); break;
    }
    if (t_div_by_zero) return nullptr;
    return valid_or_default(std::move(out));
}

NodeP crossover(int max_leafs, const Node &e1, const Node &e2, JavaRandom &rng, bool *changed)
{
    auto part = [&](const Node &e) -> NodeP {
        if (e.is_ast() && !e.args.empty()) return clone(*e.args[(size_t)rng.rand_int((int64_t)e.args.size())]);
        return clone(e);
    };
    NodeP p1 = recanon(*part(e1));
    NodeP p2 = recanon(*part(e2));
    t_div_by_zero = false;
    NodeP n;
    switch (rng.rand_int(6)) {               // crossover-sampler [:plus :times :divide12 :divide21 :minus12 :minus21]
    case 0: n = plus2(std::move(p1), std::move(p2)); break;
    case 1: n = times2(std::move(p1), std::move(p2)); break;
    case 2: n = div2(std::move(p1), std::move(p2)); break;
    case 3: n = div2(std::move(p2), std::move(p1)); break;
    case 4: n = sub2(std::move(p1), std::move(p2)); break;
    default: n = sub2(std::move(p2), std::move(p1)); break;
    }
    if (changed) *changed = false;
    // "Infinite expression 1/0 encountered.": ops-modify/crossover catches the exception and returns nil, and
    // ops/crossover-fn falls back to a mutation of e1 (modify.clj:187-189, ops.clj:175-188)
    if (t_div_by_zero) return nullptr;
    // too big, or nothing changed: the last operation is kept, e1 comes back as it was (modify.clj:176-180)
    if (leaf_count(*n) > max_leafs || same_expr(*n, e1)) return clone(e1);
    if (changed) *changed = true;
    return valid_or_default(std::move(n));
}

NodeP mutate(int max_leafs, const Node &winner, JavaRandom &rng, int *mods_applied)
{
    const int mods_count = mods_count_for(rng.rand_int(MODS_SAMPLER_SIZE));
    int iters = 0, left = mods_count;
    NodeP pheno = clone(winner);
    while (left > 0) {
        const int which = rng.rand_int(mutation_count());
        NodeP next = apply_mutation(which, *pheno, &rng);
        if (!next) next = clone(*pheno);                          // the operator threw: the phenotype is kept
        const bool invalid = leaf_count(*next) > max_leafs;
        const bool discount = invalid || same_expr(*next, *pheno);
        left = invalid ? 0 : left - 1;
        if (!discount) { iters++; pheno = std::move(next); }
    }
    if (mods_applied) *mods_applied = iters;
    return pheno;
}

}  // namespace closyr

// ------------------------------------------------------------------------------------------------
// C ABI: the GA state
// ------------------------------------------------------------------------------------------------
using namespace closyr;

struct closyr_ga {
    struct Pheno { NodeP expr; double score = 0.0; bool scored = false; };
    std::vector<Pheno> pop;
    int max_leafs = 40;
    int threads = 1;
    JavaRandom rng;
    closyr_builder builder;
    std::vector<int64_t> pending;
    bool pending_valid = false;
    closyr_ga() : builder(true) {}

    void refresh_pending()
    {
        pending.clear();
        for (size_t i = 0; i < pop.size(); i++)
            if (!pop[i].scored) pending.push_back((int64_t)i);
    }
};

namespace {

template <typename F> void parallel_for(int64_t n, int threads, F body)
{
    if (threads <= 1 || n < 64) { for (int64_t i = 0; i < n; i++) body(i); return; }
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            const int64_t i0 = next.fetch_add(32);
            if (i0 >= n) break;
            for (int64_t i = i0; i < std::min(n, i0 + 32); i++) body(i);
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<int64_t>(threads, (n + 31) / 32);
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

char *dup_string(const std::string &s)
{
    char *t = (char *)malloc(s.size() + 1);
    if (t) memcpy(t, s.c_str(), s.size() + 1);
    return t;
}

}  // namespace

extern "C" {

int closyr_jrandom_rand_ints(int64_t seed, int64_t n, int32_t count, int32_t *out)
{
    if (!out || count < 0) return CLOSYR_E_ARG;
    JavaRandom r(seed);
    for (int32_t i = 0; i < count; i++) out[i] = r.rand_int(n);
    return CLOSYR_OK;
}

int closyr_jrandom_shuffle(int64_t seed, int32_t n, int32_t *perm)
{
    if (!perm || n < 0) return CLOSYR_E_ARG;
    JavaRandom r(seed);
    std::vector<int32_t> v((size_t)n);
    for (int32_t i = 0; i < n; i++) v[(size_t)i] = i;
    r.shuffle(v);
    memcpy(perm, v.data(), sizeof(int32_t) * (size_t)n);
    return CLOSYR_OK;
}

int32_t closyr_mutation_count(void) { return mutation_count(); }

int closyr_mutation_info(int32_t index, const char **op, const char **label)
{
    if (index < 0 || index >= mutation_count()) return CLOSYR_E_ARG;
    const MutationInfo mi = mutation_info(index);
    if (op) *op = mi.op;
    if (label) *label = mi.label;
    return CLOSYR_OK;
}

int closyr_mutate_expr(const char *text, int32_t index, int32_t use_rng, int64_t seed, char **out)
{
    if (!text || !out || index < 0 || index >= mutation_count()) return CLOSYR_E_ARG;
    std::string err;
    NodeP n = parse_expr(text, &err);
    if (!n) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    JavaRandom r(seed);
    n = canonical_form(*n);
    NodeP m = apply_mutation(index, *n, use_rng ? &r : nullptr);
    if (!m) m = std::move(n);
    *out = dup_string(to_fullform(*m));
    return *out ? CLOSYR_OK : CLOSYR_E_NOMEM;
}

int closyr_crossover_expr(const char *text1, const char *text2, int32_t max_leafs, int64_t seed, char **out)
{
    if (!text1 || !text2 || !out) return CLOSYR_E_ARG;
    std::string err;
    NodeP a = parse_expr(text1, &err), b = parse_expr(text2, &err);
    if (!a || !b) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    JavaRandom r(seed);
    NodeP c = crossover(max_leafs, *a, *b, r, nullptr);
    if (!c) c = clone(*a);                     // ops-modify/crossover returns nil (the caller mutates instead): e1 as it was
    *out = dup_string(to_fullform(*c));
    return *out ? CLOSYR_OK : CLOSYR_E_NOMEM;
}

int closyr_ga_create(int64_t pop_count, int32_t max_leafs, int64_t seed, int32_t threads, closyr_ga **out)
{
    if (!out || pop_count < 1 || max_leafs < 1) return CLOSYR_E_ARG;
    closyr_ga *g = new closyr_ga();
    g->max_leafs = max_leafs;
    g->threads = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    g->rng.set_seed(seed);
    g->pop.resize((size_t)pop_count);
    for (auto &p : g->pop) p.expr = make(Kind::X);            // initialize.clj:17-33: x*1, which evaluates to x
    *out = g;
    return CLOSYR_OK;
}

void closyr_ga_destroy(closyr_ga *g) { delete g; }

int closyr_ga_set_expr(closyr_ga *g, int64_t index, const char *text)
{
    if (!g || !text || index < 0 || index >= (int64_t)g->pop.size()) return CLOSYR_E_ARG;
    std::string err;
    NodeP n = parse_expr(text, &err);
    if (!n) { set_error("parse error: " + err); return CLOSYR_E_PARSE; }
    g->pop[(size_t)index].expr = canonical_form(*n);
    g->pop[(size_t)index].scored = false;
    g->pending_valid = false;
    return CLOSYR_OK;
}

int closyr_ga_pending(closyr_ga *g, closyr_population *view, int64_t *n_pending)
{
    if (!g || !view || !n_pending) return CLOSYR_E_ARG;
    g->refresh_pending();
    const int64_t n = (int64_t)g->pending.size();
    std::vector<Program> progs((size_t)n);
    parallel_for(n, g->threads, [&](int64_t i) { compile(*g->pop[(size_t)g->pending[(size_t)i]].expr, &progs[(size_t)i]); });
    g->builder.clear();
    for (auto &p : progs)
        if (!g->builder.append(p)) { set_error("out of memory"); return CLOSYR_E_NOMEM; }
    g->pending_valid = true;
    *n_pending = n;
    return closyr_builder_view(&g->builder, view);
}

int closyr_ga_pending_indices(closyr_ga *g, int64_t *out, int64_t cap)
{
    if (!g || !g->pending_valid || (!out && cap > 0) || cap < (int64_t)g->pending.size()) return CLOSYR_E_ARG;
    if (!g->pending.empty()) memcpy(out, g->pending.data(), sizeof(int64_t) * g->pending.size());
    return CLOSYR_OK;
}

int closyr_ga_commit_scores(closyr_ga *g, const double *scores, int64_t n)
{
    if (!g || !g->pending_valid || n != (int64_t)g->pending.size() || (!scores && n > 0)) {
        set_error("commit_scores: call closyr_ga_pending first and pass one score per pending phenotype");
        return CLOSYR_E_ARG;
    }
    for (int64_t i = 0; i < n; i++) {
        auto &p = g->pop[(size_t)g->pending[(size_t)i]];
        p.score = scores[i];
        p.scored = true;
    }
    g->pending.clear();
    g->pending_valid = false;
    return CLOSYR_OK;
}

int closyr_ga_evolve(closyr_ga *g, closyr_ga_stats *st)
{
    if (!g) return CLOSYR_E_ARG;
    for (auto &p : g->pop)
        if (!p.scored) { set_error("evolve: unscored phenotypes (score closyr_ga_pending first)"); return CLOSYR_E_ARG; }
    const double t0 = now_ms();
    const int64_t n = (int64_t)g->pop.size();
    std::vector<int64_t> order((size_t)n);
    for (int64_t i = 0; i < n; i++) order[(size_t)i] = i;
    g->rng.shuffle(order);                                       // ga.clj:62-66
    const int64_t pairs = (n + 1) / 2;
    // one generator per pair when the pairs are processed concurrently (the reference's single generator is
    // shared by its pmap workers, so its draw order is not defined either); threads == 1 keeps one stream
    std::vector<int64_t> seeds;
    const bool split = g->threads > 1;
    if (split) {
        seeds.resize((size_t)pairs);
        for (auto &s : seeds) s = g->rng.next_long();
    }
    std::vector<closyr_ga::Pheno> next((size_t)n);
    std::vector<double> pair_sum((size_t)pairs, 0.0);
    std::vector<uint8_t> kind((size_t)pairs, 0);                 // 0 none, 1 mutation child, 2 crossover child
    auto compete = [&](int64_t k, JavaRandom &rng) {            // ga.clj:32-46
        auto &e1 = g->pop[(size_t)order[(size_t)(2 * k)]];
        if (2 * k + 1 >= n) {
            pair_sum[(size_t)k] = e1.score;
            next[(size_t)(2 * k)] = std::move(e1);
            return;
        }
        auto &e2 = g->pop[(size_t)order[(size_t)(2 * k + 1)]];
        const bool use_mutation = rng.rand_int(5) < 4;           // [true true true true false]
        pair_sum[(size_t)k] = e1.score + e2.score;
        if (e1.score >= e2.score) {
            closyr_ga::Pheno child;
            bool changed = false;
            bool crossed = false;
            if (!use_mutation) {
                child.expr = crossover(g->max_leafs, *e1.expr, *e2.expr, rng, &changed);     // ops.clj:175-188
                crossed = child.expr != nullptr;
            }
            if (!child.expr) child.expr = mutate(g->max_leafs, *e1.expr, rng, nullptr);      // also the fallback of a crossover that threw
            kind[(size_t)k] = crossed ? 2 : 1;
            next[(size_t)(2 * k + 1)] = std::move(child);
        } else next[(size_t)(2 * k + 1)] = std::move(e2);
        next[(size_t)(2 * k)] = std::move(e1);
    };
    if (split) parallel_for(pairs, g->threads, [&](int64_t k) { JavaRandom r(seeds[(size_t)k]); compete(k, r); });
    else for (int64_t k = 0; k < pairs; k++) compete(k, g->rng);
    g->pop = std::move(next);
    g->pending_valid = false;
    if (st) {
        memset(st, 0, sizeof(*st));
        st->n_pop = n;
        st->max_pair_sum = -INFINITY;
        for (int64_t k = 0; k < pairs; k++) {
            if (kind[(size_t)k] == 1) st->n_mutation_children++;
            if (kind[(size_t)k] == 2) st->n_crossover_children++;
            st->max_pair_sum = std::max(st->max_pair_sum, pair_sum[(size_t)k]);
            // symbolic_regression.clj:409-414: any pair sum above -1e-3 ends the run
            if (pair_sum[(size_t)k] > -1e-3) st->early_stop = 1;
        }
        st->n_children = st->n_mutation_children + st->n_crossover_children;
        st->ms_modify = now_ms() - t0;
    }
    return CLOSYR_OK;
}

int closyr_ga_best(closyr_ga *g, int64_t *index, double *score, int32_t *leafs, char **expr_fullform)
{
    if (!g) return CLOSYR_E_ARG;
    // ops.clj:191-201: remove unscored, sort-by :score (stable, ascending), reverse, first
    int64_t best = -1;
    for (int64_t i = 0; i < (int64_t)g->pop.size(); i++) {
        const auto &p = g->pop[(size_t)i];
        if (!p.scored) continue;
        if (best < 0 || p.score >= g->pop[(size_t)best].score) best = i;
    }
    if (best < 0) { set_error("best: no scored phenotype"); return CLOSYR_E_ARG; }
    if (index) *index = best;
    if (score) *score = g->pop[(size_t)best].score;
    if (leafs) *leafs = leaf_count(*g->pop[(size_t)best].expr);
    if (expr_fullform) {
        *expr_fullform = dup_string(to_fullform(*g->pop[(size_t)best].expr));
        if (!*expr_fullform) return CLOSYR_E_NOMEM;
    }
    return CLOSYR_OK;
}

int closyr_ga_expr(closyr_ga *g, int64_t index, char **expr_fullform, double *score, int32_t *scored)
{
    if (!g || index < 0 || index >= (int64_t)g->pop.size()) return CLOSYR_E_ARG;
    const auto &p = g->pop[(size_t)index];
    if (expr_fullform) {
        *expr_fullform = dup_string(to_fullform(*p.expr));
        if (!*expr_fullform) return CLOSYR_E_NOMEM;
    }
    if (score) *score = p.score;
    if (scored) *scored = p.scored ? 1 : 0;
    return CLOSYR_OK;
}

int64_t closyr_ga_population_count(closyr_ga *g) { return g ? (int64_t)g->pop.size() : 0; }

}  // extern "C"
