// compiler.cpp -- expression tree -> postfix bytecode (ISA v1.1, include/closyr_b200.h).
//
// Lowering rules (frozen; the CPU oracle's tree walk applies the same per-node semantics):
//   Plus/Times      left fold in argument order.  A leaf operand (x or a real literal) is
//                   fused into ADDX/ADDC/MULX/MULC.  A non-leaf operand is evaluated after spilling the
//                   running value to a STATIC slot (ST s ... ADDS s).  Only the first two operands may be
//                   swapped (IEEE + and * are commutative bit-for-bit): a leading leaf is applied after the
//                   sub-tree that follows it, and of two leading sub-trees the one needing more slots goes first.
//   Power(b, n)     Integer literal, 1 <= |n| <= 32: SQR (2), RECIP (-1), else POWI n (square-and-multiply).
//   Power(b, 1/2)   SQRT;  Power(b, -1/2) RSQRT;  Power(E, e) EXP;  Power(b, literal) POWC;  otherwise POWS.
//   Cot/Sec/Csc     TAN/COS/SIN followed by RECIP (Symja numericises them as reciprocals).
//   Integer/Rational literals become doubles: correctly rounded integer, double(num)/double(den) (A8).
//   No transcendental constant folding: the encoding must be bit-identical in every host language.
//   Peephole (ISA v1.1): `ST0+k ; LDX|LDC` becomes the load with CL_M_PUSH, `LDX ; f` becomes f with CL_M_SRCX.
#include "compiler.hpp"

#include <cmath>
#include <cstring>

#include "../../include/closyr_b200.h"

namespace closyr {
namespace {

struct Leaf { bool is_leaf = false, is_x = false; double v = 0.0; };

Leaf leaf_of(const Node &n)
{
    Leaf l;
    if (n.kind == Kind::X) { l.is_leaf = l.is_x = true; return l; }
    if (leaf_value(n, &l.v)) l.is_leaf = true;
    return l;
}

enum class PowForm { Int, Sqrt, Rsqrt, Exp, Const, General };

PowForm pow_form(const Node &n, int *k)
{
    const Node &b = *n.args[0], &e = *n.args[1];
    if (e.kind == Kind::Int && !e.big && e.num != 0 && e.num >= -32 && e.num <= 32) { *k = (int)e.num; return PowForm::Int; }
    if (e.kind == Kind::Rat && e.den == 2 && e.num == 1) return PowForm::Sqrt;
    if (e.kind == Kind::Rat && e.den == 2 && e.num == -1) return PowForm::Rsqrt;
    if (b.kind == Kind::E) return PowForm::Exp;
    if (leaf_of(e).is_leaf && e.kind != Kind::X) return PowForm::Const;
    return PowForm::General;
}

int need(const Node &n);

int need_nary(const Node &n)
{
    size_t k = n.args.size();
    size_t idx = 1;
    int cur;
    bool l0 = leaf_of(*n.args[0]).is_leaf;
    bool l1 = k > 1 && leaf_of(*n.args[1]).is_leaf;
    if (l0) {
        if (k > 1 && !l1) { cur = need(*n.args[1]); idx = 2; }
        else cur = 0;
    } else if (k > 1 && !l1) {
        int n0 = need(*n.args[0]), n1 = need(*n.args[1]);
        cur = n1 > n0 ? std::max(n1, 1 + n0) : std::max(n0, 1 + n1);
        idx = 2;
    } else cur = need(*n.args[0]);
    for (size_t i = idx; i < k; i++)
        if (!leaf_of(*n.args[i]).is_leaf) cur = std::max(cur, 1 + need(*n.args[i]));
    return cur;
}

int need(const Node &n)
{
    switch (n.kind) {
    case Kind::Plus:
    case Kind::Times: return need_nary(n);
    case Kind::Power: {
        int k;
        switch (pow_form(n, &k)) {
        case PowForm::Exp: return need(*n.args[1]);
        case PowForm::General: return std::max(need(*n.args[0]), 1 + need(*n.args[1]));
        default: return need(*n.args[0]);
        }
    }
    case Kind::Sin: case Kind::Cos: case Kind::Tan: case Kind::Cot: case Kind::Sec: case Kind::Csc:
    case Kind::ArcSin: case Kind::ArcCos: case Kind::Log: case Kind::Abs:
        return need(*n.args[0]);
    default: return 0;
    }
}

struct Emitter {
    Program *p;
    bool overflow = false;
    int max_slot = -1;

    void emit(int op, int a = 0, int b = 0)
    {
        if (p->code.size() + 8 > CLOSYR_MAX_CODE_WORDS) { overflow = true; return; }
        if (!p->code.empty()) {                      // ISA v1.1 fusions (the code is straight-line)
            const uint32_t last = p->code.back();
            const int lop = (int)(last & 0xffu), la = (int)((last >> 8) & 0xffu);
            if ((op == CL_LDX || op == CL_LDC) && lop >= CL_ST0 && lop < CL_ADDS0) {      // ST0+k ; LDX  ->  LDX push k
                p->code.back() = CLOSYR_INSN(op, CL_M_PUSH | ((lop - CL_ST0) << 2), b);
                return;
            }
            if (op >= CL_NEG && op <= CL_ABS && lop == CL_LDX) {                           // LDX ; f  ->  f(x)
                p->code.back() = CLOSYR_INSN(op, la | CL_M_SRCX, b);
                return;
            }
        }
        p->code.push_back(CLOSYR_INSN(op, a, b));
    }
    int konst(double v)
    {
        uint64_t bits;
        memcpy(&bits, &v, 8);
        for (size_t i = 0; i < p->consts.size(); i++) {
            uint64_t o;
            memcpy(&o, &p->consts[i], 8);
            if (o == bits) return (int)i;
        }
        if (p->consts.size() + 4 > CLOSYR_MAX_CONSTS) { overflow = true; return 0; }
        p->consts.push_back(v);
        return (int)p->consts.size() - 1;
    }
    void slot_op(int generic, int fast0, int s)
    {
        if (s >= CLOSYR_MAX_SLOTS) { overflow = true; return; }
        if (s > max_slot) max_slot = s;
        if (s < CLOSYR_REG_SLOTS && fast0 >= 0) emit(fast0 + s, s);
        else emit(generic, s);
    }
    void store(int s) { slot_op(CL_ST, CL_ST0, s); }

    void load_leaf(const Leaf &l)
    {
        if (l.is_x) emit(CL_LDX);
        else emit(CL_LDC, 0, konst(l.v));
    }
    void apply_leaf(bool plus, const Leaf &l)
    {
        if (l.is_x) emit(plus ? CL_ADDX : CL_MULX);
        else emit(plus ? CL_ADDC : CL_MULC, 0, konst(l.v));
    }
    void apply_slot(bool plus, int s)
    {
        if (plus) slot_op(CL_ADDS, CL_ADDS0, s);
        else slot_op(CL_MULS, CL_MULS0, s);
    }

    void nary(const Node &n, int s)
    {
        bool plus = n.kind == Kind::Plus;
        size_t k = n.args.size();
        size_t idx = 1;
        Leaf l0 = leaf_of(*n.args[0]);
        Leaf l1 = k > 1 ? leaf_of(*n.args[1]) : Leaf();
        if (l0.is_leaf) {
            if (k > 1 && !l1.is_leaf) {
                gen(*n.args[1], s);
                apply_leaf(plus, l0);
                idx = 2;
            } else load_leaf(l0);
        } else if (k > 1 && !l1.is_leaf) {
            int first = need(*n.args[1]) > need(*n.args[0]) ? 1 : 0;
            gen(*n.args[first], s);
            store(s);
            gen(*n.args[1 - first], s + 1);
            apply_slot(plus, s);
            idx = 2;
        } else gen(*n.args[0], s);
        for (size_t i = idx; i < k; i++) {
            Leaf l = leaf_of(*n.args[i]);
            if (l.is_leaf) apply_leaf(plus, l);
            else {
                store(s);
                gen(*n.args[i], s + 1);
                apply_slot(plus, s);
            }
        }
    }

    void gen(const Node &n, int s)
    {
        if (overflow) return;
        Leaf l = leaf_of(n);
        if (l.is_leaf) { load_leaf(l); return; }
        switch (n.kind) {
        case Kind::Complex: {
            int i = konst(n.val);
            // the imaginary part must sit right behind the real part
            if (i + 1 != (int)p->consts.size()) { p->consts.push_back(n.val); i = (int)p->consts.size() - 1; }
            p->consts.push_back(n.val2);
            emit(CL_LDZ, 0, i);
            return;
        }
        case Kind::CInf: emit(CL_LDSPEC, 0); return;
        case Kind::Indet: emit(CL_LDSPEC, 1); return;
        case Kind::Plus:
        case Kind::Times: nary(n, s); return;
        case Kind::Power: {
            int k = 0;
            switch (pow_form(n, &k)) {
            case PowForm::Int:
                gen(*n.args[0], s);
                if (k == 2) emit(CL_SQR);
                else if (k == -1) emit(CL_RECIP);
                else if (k != 1) emit(CL_POWI, 0, k);
                return;
            case PowForm::Sqrt: gen(*n.args[0], s); emit(CL_SQRT); return;
            case PowForm::Rsqrt: gen(*n.args[0], s); emit(CL_RSQRT); return;
            case PowForm::Exp: gen(*n.args[1], s); emit(CL_EXP); return;
            case PowForm::Const: gen(*n.args[0], s); emit(CL_POWC, 0, konst(leaf_of(*n.args[1]).v)); return;
            case PowForm::General:
                gen(*n.args[0], s);
                store(s);
                gen(*n.args[1], s + 1);
                slot_op(CL_POWS, -1, s);
                return;
            }
            return;
        }
        case Kind::Sin: gen(*n.args[0], s); emit(CL_SIN); return;
        case Kind::Cos: gen(*n.args[0], s); emit(CL_COS); return;
        case Kind::Tan: gen(*n.args[0], s); emit(CL_TAN); return;
        case Kind::Cot: gen(*n.args[0], s); emit(CL_TAN); emit(CL_RECIP); return;
        case Kind::Sec: gen(*n.args[0], s); emit(CL_COS); emit(CL_RECIP); return;
        case Kind::Csc: gen(*n.args[0], s); emit(CL_SIN); emit(CL_RECIP); return;
        case Kind::ArcSin: gen(*n.args[0], s); emit(CL_ASIN); return;
        case Kind::ArcCos: gen(*n.args[0], s); emit(CL_ACOS); return;
        case Kind::Log: gen(*n.args[0], s); emit(CL_LOG); return;
        case Kind::Abs: gen(*n.args[0], s); emit(CL_ABS); return;
        default: emit(CL_LDSPEC, 1); return;     // unknown head: stays symbolic => never a real number
        }
    }
};

}  // namespace

void compile(const Node &root, Program *out)
{
    out->code.clear();
    out->consts.clear();
    out->aux = 0;
    Emitter em{ out };
    em.gen(root, 0);
    out->leafs = leaf_count(root);
    uint32_t flags = 0;
    bool is_const = !depends_on_x(root);
    if (is_const) flags |= CL_F_CONST;
    if (root.kind == Kind::Indet) flags |= CL_F_NIL;
    if (is_const && !root.args.empty() && !em.overflow) {
        // eval.clj:109 quirk table: each argument's literal value, or +Inf when it is not a real literal
        size_t off = out->consts.size(), argc = root.args.size();
        if (argc < 65536 && off + argc + 4 <= CLOSYR_MAX_CONSTS) {
            for (auto &a : root.args) out->consts.push_back(a->is_number() ? (a->kind == Kind::Rat ? (double)a->num / (double)a->den : a->val) : INFINITY);
            out->aux = (uint32_t)(argc << 16) | (uint32_t)off;
        }
    }
    if (em.overflow) {
        out->code.clear();
        out->consts.clear();
        out->aux = 0;
        flags |= CL_F_UNSUPPORTED;
    }
    out->n_insn = (int)out->code.size();
    out->nslots = em.overflow ? 0 : em.max_slot + 1;
    out->code.push_back(CLOSYR_INSN(CL_END, 0, 0));
    while (out->code.size() % 4) out->code.push_back(CLOSYR_INSN(CL_END, 0, 0));
    while (out->consts.size() % 2) out->consts.push_back(0.0);
    out->meta = CLOSYR_META(out->leafs, flags, out->nslots);
}

}  // namespace closyr
