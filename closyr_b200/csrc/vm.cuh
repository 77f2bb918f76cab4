// vm.cuh -- device-side bytecode interpreters for sm_100a.
//
// Two evaluators share one ISA (include/closyr_b200.h):
//   * run_real<P>:  the fast path.  IEEE-754 fp64 on P points per lane, accumulator and spill slots 0..3 in
//                   registers, every lane of the warp runs the same program so the dispatch branch is uniform.
//                   It is exact whenever the evaluation stays on the real line; it reports (per point) when
//                   it cannot know that: the result is NaN, or an instruction met a case where Symja's value
//                   algebra differs observably from IEEE (E^(-Infinity) vs E^ComplexInfinity, literals that are
//                   not real, general x^y).
//   * full_eval:    the complete value algebra of the scoring path (reals, complex numbers, ComplexInfinity,
//                   Indeterminate; SURVEY.md Appendix A, DESIGN.md "Semantics"), one point per lane.  Only the
//                   128-point chunks in which some point raised the report above are re-evaluated with it.
// What is being replaced: Symja's ExprEvaluator.eval of Function({x},expr)[List(xs)]
// (reference src/closyr/ops/eval.clj:40-56) and the per-element decode (eval.clj:70-80).
#pragma once
#include <cstdint>

#include "../../include/closyr_b200.h"
#include "vmath.cuh"

namespace closyr {
namespace dev {

// ------------------------------------------------------------------------------------------------
// real line helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nonfinite(double x) { return (__double2hiint(x) & 0x7ff00000) == 0x7ff00000; }
__device__ __forceinline__ bool is_inf(double x)
{
    return ((unsigned)__double2hiint(x) << 1) == 0xffe00000u && __double2loint(x) == 0;
}
__device__ __forceinline__ bool is_integer_value(double y) { return !nonfinite(y) && floor(y) == y; }

// ------------------------------------------------------------------------------------------------
// single-copy math.  Every warp runs a different program, so the interpreter jumps between routines all the
// time and the hot part of the kernel image has to fit the instruction cache (first version: every libdevice
// routine inlined per point and per interpreter = 300 KB of SASS, 46 of 47 warp-cycles stalled on
// instruction fetch).  Hence ONE entry point for every real elementary function, called per point; each
// libdevice routine exists exactly once in the image, and both interpreters and the complex algebra use it.
// ------------------------------------------------------------------------------------------------
// integer power by square-and-multiply, MSB first (the same chain the oracle walks); n != 0
__device__ __forceinline__ double powi_real(double a, int n)
{
    unsigned m = (unsigned)(n < 0 ? -n : n);
    int top = 31 - __clz((int)m);
    double r = a;
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        r = r * r;
        if ((m >> b) & 1u) r = r * a;
    }
    return n < 0 ? 1.0 / r : r;
}


// scalar entry points of the lean routines (vmath.cuh): a few hundred bytes each, shared by both evaluators
__device__ __noinline__ double m_exp(double x) { double t[1] = { x }; v_exp<1>(t); return t[0]; }
__device__ __noinline__ double m_log(double x) { double t[1] = { x }; v_log<1>(t); return t[0]; }
__device__ __noinline__ double m_sincos(double x, int shift) { double t[1] = { x }; v_sincos<1>(t, shift); return t[0]; }
struct SC { double s, c; };
__device__ __noinline__ SC m_sincos_both(double x) { SC r; v_sincos_both(x, r.s, r.c); return r; }
__device__ __noinline__ double m_recip(double x) { return 1.0 / x; }
__device__ __noinline__ double m_sqrt(double x) { return sqrt(x); }

// the library routines that have no lean version: one copy each, called directly by both evaluators
__device__ __noinline__ double m_tan(double x) { return tan(x); }
__device__ __noinline__ double m_asin(double x) { return asin(x); }
__device__ __noinline__ double m_acos(double x) { return acos(x); }
__device__ __noinline__ double m_pow(double x, double y) { return pow(x, y); }

// real elementary function `op` of a (c = the literal operand of POWI / POWC): the generic entry, used where the
// operation is not known at compile time
__device__ __noinline__ double u_real(uint32_t op, double a, double c)
{
    switch (op) {
    case CL_RECIP: return m_recip(a);
    case CL_SQRT: return m_sqrt(a);
    case CL_RSQRT: return m_recip(m_sqrt(a));
    case CL_POWI: return powi_real(a, (int)c);
    case CL_POWC: return m_pow(a, c);
    case CL_EXP: return m_exp(a);
    case CL_LOG: return m_log(a);
    case CL_SIN: return m_sincos(a, 0);
    case CL_COS: return m_sincos(a, 1);
    case CL_TAN: return m_tan(a);
    case CL_ASIN: return m_asin(a);
    case CL_ACOS: return m_acos(a);
    default: return fabs(a);            // CL_ABS
    }
}
// sinh and cosh of the same argument (the complex sin / cos / tan need both): one exponential where that is
// accurate (0.5 <= |y| < 700: e - 1/e cancels at most one bit there), the library routines elsewhere
__device__ __noinline__ SC m_sinhcosh(double y)
{
    SC r;
    const double ay = fabs(y);
    if (ay >= 0.5 && ay < 700.0) {
        const double e = m_exp(ay), ei = m_recip(e);
        r.c = 0.5 * e + 0.5 * ei;
        r.s = copysign(0.5 * e - 0.5 * ei, y);
    } else {
        r.s = sinh(y);
        r.c = cosh(y);
    }
    return r;
}
#define k_sin(x) m_sincos((x), 0)
#define k_cos(x) m_sincos((x), 1)
#define k_tan(x) m_tan(x)
#define k_exp(x) m_exp(x)
#define k_log(x) m_log(x)
#define k_asin(x) m_asin(x)
#define k_acos(x) m_acos(x)
#define k_pow(x, y) m_pow((x), (y))
#define k_recip(x) m_recip(x)
#define k_sqrt(x) m_sqrt(x)
__device__ __noinline__ double k_hypot(double x, double y) { return hypot(x, y); }
__device__ __noinline__ double k_atan2(double y, double x) { return atan2(y, x); }

// ------------------------------------------------------------------------------------------------
// full value algebra
// ------------------------------------------------------------------------------------------------
struct Z { double re, im; };
enum { T_REAL = 0, T_CPLX = 1, T_CINF = 2, T_INDET = 3 };

__device__ __forceinline__ int ztag(Z v)
{
    if (isnan(v.re) || isnan(v.im)) return T_INDET;
    if (v.im == 0.0) return T_REAL;
    if (isinf(v.re) || isinf(v.im)) return T_CINF;
    return T_CPLX;
}
__device__ __forceinline__ Z zv(double re, double im) { Z z; z.re = re; z.im = im; return z; }
__device__ __forceinline__ Z zindet() { return zv(__longlong_as_double(0x7ff8000000000000ll), 0.0); }
__device__ __forceinline__ Z zcinf() { return zv(__longlong_as_double(0x7ff0000000000000ll), __longlong_as_double(0x7ff0000000000000ll)); }
// rounding noise must not decide whether a value is real, nor the side of a branch cut (DESIGN.md "Semantics")
#define CLOSYR_CHOP_REL 1e-10
__device__ __forceinline__ Z zmk(double re, double im)
{
    if (im != 0.0 && fabs(im) <= CLOSYR_CHOP_REL * fabs(re)) im = 0.0;
    else if (re != 0.0 && fabs(re) <= CLOSYR_CHOP_REL * fabs(im)) re = 0.0;
    Z v = zv(re, im);
    int t = ztag(v);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zcinf();
    if (t == T_REAL) return zv(re, 0.0);
    return v;
}

__device__ __noinline__ Z zadd(Z a, Z b)
{
    int ta = ztag(a), tb = ztag(b);
    if (ta == T_INDET || tb == T_INDET) return zindet();
    if (ta == T_CINF || tb == T_CINF) {
        Z o = (ta == T_CINF) ? b : a;
        int to = (ta == T_CINF) ? tb : ta;
        if (to == T_CINF) return zindet();
        if (to == T_REAL && isinf(o.re)) return zindet();
        return zcinf();
    }
    return zmk(a.re + b.re, a.im + b.im);
}

__device__ __noinline__ Z zmul(Z a, Z b)
{
    int ta = ztag(a), tb = ztag(b);
    if (ta == T_INDET || tb == T_INDET) return zindet();
    if (ta == T_CINF || tb == T_CINF) {
        Z o = (ta == T_CINF) ? b : a;
        int to = (ta == T_CINF) ? tb : ta;
        if (to == T_REAL && o.re == 0.0) return zindet();
        return zcinf();
    }
    if (ta == T_REAL && tb == T_REAL) return zmk(a.re * b.re, 0.0);
    if (ta == T_REAL) return zmk(a.re * b.re, a.re * b.im);
    if (tb == T_REAL) return zmk(a.re * b.re, a.im * b.re);
    return zmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}

// add / multiply of finite operands with finite results: the arithmetic of zadd / zmul without the tag plumbing
// (the complete-algebra interpreter spends most of its instructions on + and * of ordinary complex numbers)
__device__ __forceinline__ bool zfinite(Z a) { return !(nonfinite(a.re) || nonfinite(a.im)); }
__device__ __forceinline__ Z zchop(Z r)
{
    if (r.im != 0.0 && fabs(r.im) <= CLOSYR_CHOP_REL * fabs(r.re)) r.im = 0.0;
    else if (r.re != 0.0 && fabs(r.re) <= CLOSYR_CHOP_REL * fabs(r.im)) r.re = 0.0;
    return r;
}
__device__ __forceinline__ Z lean_add(Z a, Z b)
{
    const Z r = zv(a.re + b.re, a.im + b.im);
    if (zfinite(a) && zfinite(b) && zfinite(r)) return zchop(r);
    return zadd(a, b);
}
__device__ __forceinline__ Z lean_mul(Z a, Z b)         // at least one operand is off the real line
{
    Z r;
    if (a.im == 0.0) r = zv(a.re * b.re, a.re * b.im);
    else if (b.im == 0.0) r = zv(a.re * b.re, a.im * b.re);
    else r = zv(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
    if (zfinite(a) && zfinite(b) && zfinite(r)) return zchop(r);
    return zmul(a, b);
}

__device__ __forceinline__ Z zneg(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zcinf();
    if (t == T_REAL) return zv(-a.re, 0.0);
    return zv(-a.re, -a.im);
}

__device__ __noinline__ Z zrecip(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zv(0.0, 0.0);
    if (t == T_REAL) {
        if (a.re == 0.0) return zcinf();
        return zv(k_recip(a.re), 0.0);
    }
    double c = a.re, d = a.im;           // Smith's algorithm (Hipparchus Complex.reciprocal)
    if (fabs(c) < fabs(d)) {
        double q = c / d;
        double s = 1.0 / (c * q + d);
        return zmk(s * q, -s);
    }
    double q = d / c;
    double s = 1.0 / (d * q + c);
    return zmk(s, -s * q);
}

__device__ __noinline__ Z zsqrt(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zcinf();
    if (t == T_REAL) {
        if (a.re >= 0.0) return zv(k_sqrt(a.re), 0.0);
        if (isinf(a.re)) return zcinf();
        return zmk(0.0, k_sqrt(-a.re));
    }
    double x = a.re, y = a.im;
    double tt = sqrt((fabs(x) + k_hypot(x, y)) / 2.0);
    if (x >= 0.0) return zmk(tt, y / (2.0 * tt));
    return zmk(fabs(y) / (2.0 * tt), copysign(tt, y));
}

__device__ __noinline__ Z zexp(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_exp(a.re), 0.0);
    double e = k_exp(a.re);
    const SC sc = m_sincos_both(a.im);
    return zmk(e * sc.c, e * sc.s);
}

__device__ __noinline__ Z zlog(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zv(__longlong_as_double(0x7ff0000000000000ll), 0.0);
    if (t == T_REAL) {
        if (a.re > 0.0) return zv(k_log(a.re), 0.0);
        if (a.re == 0.0) return zv(__longlong_as_double(0xfff0000000000000ll), 0.0);
        if (isinf(a.re)) return zcinf();
        return zmk(k_log(-a.re), 3.141592653589793);
    }
    return zmk(k_log(k_hypot(a.re, a.im)), k_atan2(a.im, a.re));
}

__device__ __forceinline__ Z zsin(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_sin(a.re), 0.0);
    const SC sc = m_sincos_both(a.re), h = m_sinhcosh(a.im);
    return zmk(sc.s * h.c, sc.c * h.s);
}

__device__ __forceinline__ Z zcos(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_cos(a.re), 0.0);
    const SC sc = m_sincos_both(a.re), h = m_sinhcosh(a.im);
    return zmk(sc.c * h.c, -(sc.s * h.s));
}

__device__ __noinline__ Z ztan(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_tan(a.re), 0.0);
    if (a.im > 20.0) return zv(0.0, 1.0);
    if (a.im < -20.0) return zv(0.0, -1.0);
    double a2 = 2.0 * a.re, b2 = 2.0 * a.im;
    const SC sc = m_sincos_both(a2), h = m_sinhcosh(b2);
    double d = sc.c + h.c;
    return zmk(sc.s / d, h.s / d);
}

__device__ __forceinline__ Z zmul_i(Z a) { return zmk(-a.im, a.re); }
__device__ __forceinline__ Z zmul_minus_i(Z a) { return zmk(a.im, -a.re); }
__device__ __forceinline__ Z zsqrt1z(Z z) { return zsqrt(lean_add(zv(1.0, 0.0), zneg(lean_mul(z, z)))); }

__device__ __noinline__ Z zasin(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zcinf();
    if (t == T_REAL) {
        if (fabs(a.re) <= 1.0) return zv(k_asin(a.re), 0.0);
        if (isinf(a.re)) return zcinf();
        // real |x| > 1: same principal value, without the cancellation of the general formula
        const double ax = fabs(a.re), l = k_log(ax + sqrt((ax - 1.0) * (ax + 1.0)));
        return a.re > 0.0 ? zmk(3.141592653589793 / 2.0, -l) : zmk(-(3.141592653589793 / 2.0), l);
    }
    Z w = zlog(lean_add(zsqrt1z(a), zmul_i(a)));
    int tw = ztag(w);
    if (tw == T_INDET) return zindet();
    if (tw == T_CINF || (tw == T_REAL && isinf(w.re))) return zcinf();
    return zmul_minus_i(w);
}

__device__ __noinline__ Z zacos(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zcinf();
    if (t == T_REAL) {
        if (fabs(a.re) <= 1.0) return zv(k_acos(a.re), 0.0);
        if (isinf(a.re)) return zcinf();
        const double ax = fabs(a.re), l = k_log(ax + sqrt((ax - 1.0) * (ax + 1.0)));
        return a.re > 0.0 ? zmk(0.0, l) : zmk(3.141592653589793, -l);
    }
    Z s = zsqrt1z(a);
    int ts = ztag(s);
    if (ts == T_INDET) return zindet();
    if (ts == T_CINF || (ts == T_REAL && isinf(s.re))) return zcinf();
    Z w = zlog(lean_add(a, zmul_i(s)));
    int tw = ztag(w);
    if (tw == T_INDET) return zindet();
    if (tw == T_CINF || (tw == T_REAL && isinf(w.re))) return zcinf();
    return zmul_minus_i(w);
}

__device__ __forceinline__ Z zabs(Z a)
{
    int t = ztag(a);
    if (t == T_INDET) return zindet();
    if (t == T_CINF) return zv(__longlong_as_double(0x7ff0000000000000ll), 0.0);
    if (t == T_REAL) return zv(fabs(a.re), 0.0);
    return zmk(k_hypot(a.re, a.im), 0.0);
}

__device__ __noinline__ Z zpowi(Z a, int n)
{
    unsigned m = (unsigned)(n < 0 ? -n : n);
    int top = 31 - __clz((int)m);
    Z r = a;
    for (int b = top - 1; b >= 0; b--) {
        r = lean_mul(r, r);
        if ((m >> b) & 1u) r = lean_mul(r, a);
    }
    if (n < 0) r = zrecip(r);
    return r;
}

__device__ __noinline__ Z zpow(Z z, Z w)
{
    const double PINF = __longlong_as_double(0x7ff0000000000000ll);
    int tz = ztag(z), tw = ztag(w);
    if (tz == T_INDET || tw == T_INDET) return zindet();
    if (tw == T_CINF) return zindet();
    double yr = w.re;
    if (tz == T_CINF) {
        if (tw == T_REAL && isinf(yr)) return zindet();
        if (yr > 0.0) return zcinf();
        if (yr < 0.0) return zv(0.0, 0.0);
        return zindet();
    }
    if (tz == T_REAL && tw == T_REAL) {
        double x = z.re, y = yr;
        if (y == 0.0) return (!nonfinite(x) && x != 0.0) ? zv(1.0, 0.0) : zindet();
        if (isinf(y)) {
            double ax = fabs(x);
            if (ax == 1.0) return zindet();
            bool grows = (ax > 1.0) == (y > 0.0);
            if (!grows) return zv(0.0, 0.0);
            return (x > 0.0) ? zv(PINF, 0.0) : zcinf();
        }
        if (isinf(x)) {
            if (y < 0.0) return zv(0.0, 0.0);
            if (x > 0.0) return zv(PINF, 0.0);
            return is_integer_value(y) ? zmk(k_pow(x, y), 0.0) : zcinf();
        }
        if (x == 0.0) return (y > 0.0) ? zv(0.0, 0.0) : zcinf();
        if (x > 0.0 || is_integer_value(y)) return zmk(k_pow(x, y), 0.0);
        double m = k_pow(-x, y);
        double th = y * 3.141592653589793;
        return zmk(m * k_cos(th), m * k_sin(th));
    }
    if (tz == T_REAL && z.re == 0.0) {
        if (yr > 0.0) return zv(0.0, 0.0);
        if (yr < 0.0) return zcinf();
        return zindet();
    }
    if (tz == T_REAL && isinf(z.re)) {
        if (yr > 0.0) return zcinf();
        if (yr < 0.0) return zv(0.0, 0.0);
        return zindet();
    }
    if (tw == T_REAL && isinf(yr)) {
        double az = k_hypot(z.re, z.im);
        if (az == 1.0) return zindet();
        return ((az > 1.0) == (yr > 0.0)) ? zcinf() : zv(0.0, 0.0);
    }
    return zexp(lean_mul(w, zlog(z)));
}

// One point through the program with the complete algebra: only the points the fast path reported come
// here.  Flat on purpose -- one switch, the real-line case of every instruction inline (the same IEEE sequence
// as the fast path, through the same m_* routines), one call into z* only when a value is off the real line.
// A first version went through one generic per-value routine and then u_real -> m_*: 125 instructions per point-instruction and an
// instruction-fetch stall at almost every step (ncu profile r01_f).
#define CLOSYR_ZSET(expr) { acc = (expr); break; }
#define CLOSYR_REAL1(cond, fast, zf)                        \
    {                                                       \
        const double x_ = acc.re;                           \
        if (acc.im == 0.0 && (cond)) { acc.re = (fast); break; } \
        CLOSYR_ZSET(zf(acc))                                \
    }
// `tab` (or nullptr) points at this point's column of the f(x_i) tables (kernels.cuh): an elementary function of
// the raw variable takes its value from there whenever it would have called the real routine.
__device__ __noinline__ Z run_full(const uint32_t *__restrict__ scode, const double *__restrict__ sconst, double x,
                                   const double *__restrict__ tab, int64_t tab_stride)
{
    Z slot[CLOSYR_MAX_SLOTS];
    Z acc = zv(0.0, 0.0);
#pragma unroll 1
    for (int pc = 0;; pc++) {
        const uint32_t w = scode[pc];
        const uint32_t op = w & 0xffu, a = (w >> 8) & 0xffu;
        const int b = (int)(w >> 16);
        if (acc.re != acc.re) return acc;                   // Indeterminate absorbs every later instruction
        // the frequent instructions first (slot forms, then Sin/Cos): the switch below compiles to a compare tree
        if (op >= CL_ST0) {
            if (op < CL_ADDS0) { slot[a] = acc; continue; }
            const Z sl = slot[a];
            if (sl.im == 0.0 && acc.im == 0.0) acc.re = op < CL_MULS0 ? sl.re + acc.re : sl.re * acc.re;
            else acc = op < CL_MULS0 ? lean_add(sl, acc) : lean_mul(sl, acc);
            continue;
        }
        const double *t = nullptr;
        if ((a & (CL_M_PUSH | CL_M_SRCX)) && op < CL_ST && op != CL_LDSPEC) {   // ISA v1.1 modifiers
            if (a & CL_M_PUSH) slot[CL_M_PUSH_SLOT(a)] = acc;
            if (a & CL_M_SRCX) {
                acc = zv(x, 0.0);
                if (tab != nullptr && op >= CL_RECIP && op <= CL_ACOS && ((0x1FCDu >> (op - CL_RECIP)) & 1u))
                    t = tab + (op - CL_RECIP) * tab_stride;
            }
        }
        if (op == CL_SIN || op == CL_COS) {
            if (acc.im == 0.0) acc.re = t ? *t : m_sincos(acc.re, (int)(op - CL_SIN));
            else acc = op == CL_SIN ? zsin(acc) : zcos(acc);
            continue;
        }
        switch (op) {
        case CL_END: return acc;
        case CL_LDX: acc = zv(x, 0.0); break;
        case CL_LDC: acc = zv(sconst[b], 0.0); break;
        case CL_LDZ: acc = zmk(sconst[b], sconst[b + 1]); break;
        case CL_LDSPEC: acc = a == 0 ? zcinf() : zindet(); break;
        case CL_ADDX: if (acc.im == 0.0) { acc.re += x; break; } CLOSYR_ZSET(lean_add(acc, zv(x, 0.0)))
        case CL_ADDC: if (acc.im == 0.0) { acc.re += sconst[b]; break; } CLOSYR_ZSET(lean_add(acc, zv(sconst[b], 0.0)))
        case CL_MULX: if (acc.im == 0.0) { acc.re *= x; break; } CLOSYR_ZSET(lean_mul(acc, zv(x, 0.0)))
        case CL_MULC: if (acc.im == 0.0) { acc.re *= sconst[b]; break; } CLOSYR_ZSET(lean_mul(acc, zv(sconst[b], 0.0)))
        case CL_NEG: CLOSYR_REAL1(true, -x_, zneg)
        case CL_SQR: if (acc.im == 0.0) { acc.re = acc.re * acc.re; break; } CLOSYR_ZSET(lean_mul(acc, acc))
        case CL_RECIP: CLOSYR_REAL1(x_ != 0.0, t ? *t : m_recip(x_), zrecip)
        case CL_SQRT: CLOSYR_REAL1(x_ >= 0.0, t ? *t : m_sqrt(x_), zsqrt)
        case CL_RSQRT: if (acc.im == 0.0 && acc.re > 0.0) { acc.re = t ? *t : m_recip(m_sqrt(acc.re)); break; } CLOSYR_ZSET(zrecip(zsqrt(acc)))
        case CL_POWI: {
            const int n = (int)(short)b;
            if (acc.im == 0.0) {
                const double r = powi_real(acc.re, n < 0 ? -n : n);
                if (n > 0) acc.re = r;
                else if (r != 0.0) acc.re = m_recip(r);
                else acc = zcinf();
                break;
            }
            CLOSYR_ZSET(zpowi(acc, n))
        }
        case CL_POWC: {
            const double c = sconst[b];
            if (acc.im == 0.0 && acc.re > 0.0 && !is_inf(acc.re) && !nonfinite(c) && c != 0.0) { acc.re = m_pow(acc.re, c); break; }
            CLOSYR_ZSET(zpow(acc, zmk(c, 0.0)))
        }
        case CL_EXP: CLOSYR_REAL1(true, t ? *t : m_exp(x_), zexp)
        case CL_LOG: CLOSYR_REAL1(x_ > 0.0, t ? *t : m_log(x_), zlog)
        case CL_TAN: CLOSYR_REAL1(true, t ? *t : m_tan(x_), ztan)
        case CL_ASIN: CLOSYR_REAL1(fabs(x_) <= 1.0, t ? *t : m_asin(x_), zasin)
        case CL_ACOS: CLOSYR_REAL1(fabs(x_) <= 1.0, t ? *t : m_acos(x_), zacos)
        case CL_ABS: CLOSYR_REAL1(true, fabs(x_), zabs)
        default: {
            // slot instructions: `a` is the slot for both the generic and the register forms
            const uint32_t kind = op;                       // generic forms (the register forms were handled above)
            if (kind == CL_ST) { slot[a] = acc; break; }
            const Z sl = slot[a];
            if (kind == CL_POWS) acc = zpow(sl, acc);
            else if (sl.im == 0.0 && acc.im == 0.0) acc.re = kind == CL_ADDS ? sl.re + acc.re : sl.re * acc.re;
            else acc = kind == CL_ADDS ? lean_add(sl, acc) : lean_mul(sl, acc);
            break;
        }
        }
    }
}
#undef CLOSYR_ZSET
#undef CLOSYR_REAL1

// ------------------------------------------------------------------------------------------------
// fast path: P points per lane, IEEE real arithmetic, register-resident accumulator and slots 0..3.
//
// Invariant kept per point (v = the IEEE value here, s = the value of the full algebra):
//   v finite  <=>  s is a finite real and v == s;   v = +-Inf  =>  s is ComplexInfinity or a directed infinity
//   (IEEE cannot tell which);   v NaN, no event  =>  s is Indeterminate (absorbing: the point is not real).
// The returned mask has bit j set when point j must be re-evaluated with the complete algebra because the
// invariant cannot be kept: an elementary function (anything but + * neg square) produced NaN from a non-NaN
// input (it left its real domain: the algebra continues over the complex numbers) or was handed an infinity
// (1/Inf = 0, E^-Inf = 0 ... are only right for some infinities); literals that are not real; general x^y.
// ------------------------------------------------------------------------------------------------
#define CLOSYR_FORJ _Pragma("unroll") for (int j = 0; j < P; j++)

// spill slots are addressed with a 32-bit shared-window address computed once per evaluation: through a generic
// pointer the compiler rebuilds the window base (S2UR CgaCtaId, UMOV, UIADD3, ULEA) at every access
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// All elementary-function instructions of the fast path, 4 points per call (the struct travels in registers):
// four calls of the scalar routine per instruction, one copy of every routine in the kernel image.  Returns the values and the mask of points that need the complete algebra (see below).
struct D4 { double v[4]; };
struct R4 { D4 d; uint32_t ev; };

__device__ __forceinline__ R4 k1_unary(uint32_t op, D4 in, double c)
{
    R4 out;
    double a[4] = { in.v[0], in.v[1], in.v[2], in.v[3] };
#define CLOSYR_K4(expr) _Pragma("unroll") for (int j = 0; j < 4; j++) a[j] = (expr);
    switch (op) {
    case CL_EXP: CLOSYR_K4(m_exp(a[j])) break;
    case CL_LOG: CLOSYR_K4(m_log(a[j])) break;
    case CL_SIN: CLOSYR_K4(m_sincos(a[j], 0)) break;
    case CL_COS: CLOSYR_K4(m_sincos(a[j], 1)) break;
    case CL_RECIP: CLOSYR_K4(m_recip(a[j])) break;
    case CL_SQRT: CLOSYR_K4(m_sqrt(a[j])) break;
    case CL_TAN: CLOSYR_K4(m_tan(a[j])) break;
    case CL_ASIN: CLOSYR_K4(m_asin(a[j])) break;
    case CL_ACOS: CLOSYR_K4(m_acos(a[j])) break;
    case CL_POWC: CLOSYR_K4(m_pow(a[j], c)) break;
    default: CLOSYR_K4(u_real(op, a[j], c)) break;        // RSQRT POWI ABS
    }
#undef CLOSYR_K4
    uint32_t ev = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double i_ = in.v[j];
        if (is_inf(i_) || (a[j] != a[j] && i_ == i_)) ev |= 1u << j;
        out.d.v[j] = a[j];
    }
    out.ev = ev;
    return out;
}

// Loop-carried state is only the accumulator (P doubles); x is loaded once per chunk.  The spill slots live
// in shared memory ([slot][point][lane], conflict-free): with slots in registers the compiler copied the whole
// 40-register state at every dispatch (36 % of all issued instructions were IMAD.MOV, ncu profile r01_c).
template <int P>
__device__ __forceinline__ uint32_t run_real(const uint32_t *__restrict__ scode, const double *__restrict__ sconst,
                                             double *__restrict__ sslot, const double (&x)[P], double (&acc)[P],
                                             const double *__restrict__ tab, int64_t tab_stride)
{
    double ls[CLOSYR_MAX_SLOTS - CLOSYR_REG_SLOTS][P];    // deep spill slots: local memory, rarely touched
    uint32_t ev = 0;
    int pc = 0;
    uint32_t w = scode[0];
    const uint32_t slot_base = (uint32_t)__cvta_generic_to_shared(sslot);
    CLOSYR_FORJ acc[j] = 0.0;
#pragma unroll 1
    for (;;) {
        const uint32_t op = w & 0xffu, a = (w >> 8) & 0xffu;
        const int b = (int)(w >> 16);
        w = scode[++pc];                                     // prefetch (the buffer has one word of slack)
        if (op < CL_ST && (a & (CL_M_PUSH | CL_M_SRCX)) && op != CL_LDSPEC) {
            // ISA v1.1 modifiers of the leaf loads and the elementary functions: push acc, read x
            if (a & CL_M_PUSH) {
                const uint32_t sl = slot_base + CL_M_PUSH_SLOT(a) * (P * 32 * 8);
                CLOSYR_FORJ sts_f64(sl + j * 256, acc[j]);
            }
            if (a & CL_M_SRCX) { CLOSYR_FORJ acc[j] = x[j]; }
        }
        if (op >= CL_ST0) {
            // register-file-free slot forms: CL_ST0+a, CL_ADDS0+a, CL_MULS0+a with a in 0..3
            const uint32_t sl = slot_base + a * (P * 32 * 8);
            if (op < CL_ADDS0) { CLOSYR_FORJ sts_f64(sl + j * 256, acc[j]); }
            else if (op < CL_MULS0) { CLOSYR_FORJ acc[j] = lds_f64(sl + j * 256) + acc[j]; }
            else { CLOSYR_FORJ acc[j] = lds_f64(sl + j * 256) * acc[j]; }
            continue;
        }
        switch (op) {
        case CL_END: return ev;
        case CL_LDX: CLOSYR_FORJ acc[j] = x[j]; break;
        case CL_LDC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = c; break; }
        case CL_ADDX: CLOSYR_FORJ acc[j] = acc[j] + x[j]; break;
        case CL_ADDC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = acc[j] + c; break; }
        case CL_MULX: CLOSYR_FORJ acc[j] = acc[j] * x[j]; break;
        case CL_MULC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = acc[j] * c; break; }
        case CL_NEG: CLOSYR_FORJ acc[j] = -acc[j]; break;
        case CL_SQR: CLOSYR_FORJ acc[j] = acc[j] * acc[j]; break;
        case CL_ST:
            if (a >= CLOSYR_REG_SLOTS) { CLOSYR_FORJ ls[a - CLOSYR_REG_SLOTS][j] = acc[j]; }
            else ev = (1u << P) - 1u;
            break;
        case CL_ADDS:
            if (a >= CLOSYR_REG_SLOTS) { CLOSYR_FORJ acc[j] = ls[a - CLOSYR_REG_SLOTS][j] + acc[j]; }
            else ev = (1u << P) - 1u;
            break;
        case CL_MULS:
            if (a >= CLOSYR_REG_SLOTS) { CLOSYR_FORJ acc[j] = ls[a - CLOSYR_REG_SLOTS][j] * acc[j]; }
            else ev = (1u << P) - 1u;
            break;
        case CL_LDZ:
        case CL_LDSPEC:
        case CL_POWS: return (1u << P) - 1u;                 // program-uniform: every point needs the complete algebra
        default: {
            // every other instruction is a real elementary function of the accumulator
            double c = (double)(short)b;
            if (op == CL_POWC) {
                c = sconst[b];
                if (nonfinite(c) || c == 0.0) { ev = (1u << P) - 1u; break; }
            }
            static_assert(P == 4, "k1_unary is written for 4 points per lane");
            if ((a & CL_M_SRCX) && tab != nullptr && op >= CL_RECIP && op <= CL_ACOS && ((0x1FCDu >> (op - CL_RECIP)) & 1u)) {
                // f(x) of the raw variable: tabulated per point by leaf_table_kernel with the same routines
                // (`tab` points at this lane's first point); the report rule is k1_unary's
                const double2 t0 = *reinterpret_cast<const double2 *>(tab + (op - CL_RECIP) * tab_stride);
                const double2 t1 = *reinterpret_cast<const double2 *>(tab + (op - CL_RECIP) * tab_stride + 2);
                acc[0] = t0.x; acc[1] = t0.y; acc[2] = t1.x; acc[3] = t1.y;
                uint32_t e4 = 0;
                CLOSYR_FORJ if (is_inf(x[j]) || (acc[j] != acc[j] && x[j] == x[j])) e4 |= 1u << j;
                ev |= e4;
            } else {
                D4 in;
                CLOSYR_FORJ in.v[j] = acc[j];
                const R4 r = k1_unary(op, in, c);
                CLOSYR_FORJ acc[j] = r.d.v[j];
                ev |= r.ev;
            }
            // points are sorted by x, so a domain exit usually takes whole chunks: once every point of the warp is
            // reported the rest of the fast pass is wasted work (the complete algebra starts over anyway)
            if (__all_sync(0xffffffffu, ev == (1u << P) - 1u)) return ev;
            break;
        }
        }
    }
}

}  // namespace dev
}  // namespace closyr
