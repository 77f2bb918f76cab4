// vm.cuh -- device-side bytecode interpreters for sm_100a.
//
// The ISA of include/closyr_b200.h is what the caller writes; a CTA translates the program it staged into
// THREADED CODE once (decode_program: 16 bytes per instruction = handler index, operand word, the fp64 literal
// itself) so that the interpreters dispatch with one LDS.128 and one jump-table branch, without modifier tests,
// slot tests or constant-pool indexing.  Two evaluators run that code:
//   * run_real<P>:  the fast path.  IEEE-754 fp64 on P points per lane, accumulator in registers, spill slots in
//                   shared memory, every lane of the warp runs the same program so the dispatch branch is uniform.
//                   It is exact whenever the evaluation stays on the real line; it reports (per point) when it
//                   cannot know that: the result is NaN (NaN is absorbing on this path, so a domain exit anywhere
//                   shows in the result), or an instruction that maps an infinity to a finite number was handed one
//                   (IEEE cannot tell ComplexInfinity from a directed infinity, nor from an Indeterminate the
//                   algebra reached earlier).
//   * run_cplx<PC>: the complete value algebra of the scoring path (reals, complex numbers, ComplexInfinity,
//                   Indeterminate; SURVEY.md Appendix A, DESIGN.md "Semantics"), PC points per lane over (re, im)
//                   register pairs.  The warp is in REAL MODE while every value of the accumulator is on the real
//                   line (same handlers and same arithmetic as run_real) and switches to COMPLEX MODE at the first
//                   instruction that leaves it; a leaf load switches back.  Special values (infinities,
//                   Indeterminate) take a per-point scalar path through the z* functions below.
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


// scalar entry points of the lean routines (vmath.cuh): a few hundred bytes each, used by the complex algebra
__device__ __noinline__ double m_exp(double x) { double t[1] = { x }; v_exp<1>(t); return t[0]; }
__device__ __noinline__ double m_log(double x) { double t[1] = { x }; v_log<1>(t); return t[0]; }
__device__ __noinline__ double m_sin(double x) { double t[1] = { x }; v_sin<1>(t); return t[0]; }
__device__ __noinline__ double m_cos(double x) { double t[1] = { x }; v_cos<1>(t); return t[0]; }
struct SC { double s, c; };
__device__ __noinline__ SC m_sincos_both(double x)
{
    double s[1] = { x }, c[1] = { x };
    v_sin<1>(s);
    v_cos<1>(c);
    SC r;
    r.s = s[0];
    r.c = c[0];
    return r;
}
__device__ __forceinline__ double m_recip(double x) { return 1.0 / x; }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }

// the library routines that have no lean version: one copy each, called by both evaluators
__device__ __noinline__ double m_tan(double x) { return vm_special(x) ? vm_nan_of(x) : tan(x); }       // tan(Inf), tan(NaN): no argument reduction
__device__ __noinline__ double m_asin(double x) { return asin(x); }
__device__ __noinline__ double m_acos(double x) { return acos(x); }
__device__ __noinline__ double m_pow(double x, double y) { return (x != x && y == y && y != 0.0) ? x : pow(x, y); }   // NaN^y

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
#define k_sin(x) m_sin(x)
#define k_cos(x) m_cos(x)
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

__device__ __noinline__ Z zneg(Z a)
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

__device__ __noinline__ Z zsin(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_sin(a.re), 0.0);
    const SC sc = m_sincos_both(a.re), h = m_sinhcosh(a.im);
    return zmk(sc.s * h.c, sc.c * h.s);
}

__device__ __noinline__ Z zcos(Z a)
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

__device__ __noinline__ Z zabs(Z a)
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

// ------------------------------------------------------------------------------------------------
// threaded code: what a CTA interprets.  One entry per ISA word (same index), 16 bytes:
//   h    handler index (TH_*)
//   aux  leaf loads, TH_TAB, TH_MOD: bit 0 SRCX, bit 1 PUSH, bits 10..11 push slot (= its byte offset, 1 KiB per slot),
//        bits 2..9 the function's own handler (TH_TAB / TH_MOD); bits 16..31 the ISA `b` field (POWI exponent, LDZ
//        constant index) or LDSPEC's `a`;  ST / ADDS / MULS / POWS: slot number << 10
//   c    the fp64 literal of ADDC / MULC / LDC / POWC, or (TH_TAB) the byte offset of the function's table row
// ------------------------------------------------------------------------------------------------
struct __align__(16) TInsn { uint32_t h, aux; double c; };

enum {
    // the "cheap" handlers, 0 .. TH_CHEAP - 1: both interpreters run them inside one PTX block (brx.idx dispatch).
    // The ...P forms store acc to a slot first (ISA v1.1 PUSH); TH_TAB = a tabulated function of the raw variable
    // whose table entry is the algebra's value everywhere (sin cos exp tan); TH_TABR = 1/x (same for the fast path;
    // the complete algebra must turn the pole into ComplexInfinity)
    TH_END = 0, TH_LDX, TH_LDXP, TH_LDC, TH_LDCP, TH_TAB, TH_TABP, TH_TABR, TH_TABRP, TH_ADDX, TH_ADDC, TH_MULX, TH_MULC,
    TH_NEG, TH_SQR, TH_ABS, TH_ST, TH_ADDS, TH_MULS,
    TH_CHEAP,
    // TH_TABC = a tabulated function with a domain (sqrt, x^-1/2, log, asin, acos); TH_MOD = PUSH / SRCX modifiers of
    // a function that has no table, followed by the function itself
    TH_TABC = TH_CHEAP, TH_MOD,
    TH_RECIP, TH_SQRT, TH_RSQRT, TH_POWI, TH_POWC, TH_EXP, TH_LOG, TH_SIN, TH_COS, TH_TAN, TH_ASIN, TH_ACOS,
    TH_LDZ, TH_LDSPEC, TH_POWS,
    TH_COUNT
};
static_assert(TH_CHEAP == 19 && TH_COUNT == 36, "the branch-target tables of the interpreters list the handlers by position");
// program properties found while decoding
constexpr uint32_t PF_NO_FAST = 1u;   // an instruction the IEEE fast path cannot express: every point takes run_cplx
constexpr uint32_t PF_DEEP = 2u;      // uses spill slots >= CLOSYR_REG_SLOTS (kept in local memory by run_cplx)
constexpr uint32_t PF_DOMAIN = 4u;    // calls a function with a real domain (sqrt log asin acos x^c ...): 3 of 4 such programs of the GA mix leave the real line somewhere

// handler of the elementary function CL_NEG + i
__constant__ uint8_t closyr_th_of_function[CL_ABS - CL_NEG + 1] = {
    TH_NEG, TH_RECIP, TH_SQR, TH_SQRT, TH_RSQRT, TH_POWI, TH_POWC, TH_EXP, TH_LOG, TH_SIN, TH_COS, TH_TAN, TH_ASIN, TH_ACOS, TH_ABS,
};

// the functions of the raw variable that closyr_set_points tabulates (bit = op - CL_RECIP):
// RECIP SQRT RSQRT EXP LOG SIN COS TAN ASIN ACOS; SQR, POWI and POWC are not tabulated
constexpr uint32_t VM_TAB_MASK = 0x1FCDu;

// one ISA word -> one threaded-code entry; returns the PF_* bits the instruction implies
// (sconst = the program's constant pool of nk doubles; an index beyond it reads as NaN)
__device__ __forceinline__ uint32_t decode_insn(uint32_t w, const double *__restrict__ sconst, uint32_t nk, bool have_tab,
                                                int64_t tab_stride, TInsn &t)
{
    const uint32_t op = w & 0xffu, a = (w >> 8) & 0xffu, b = w >> 16;
    const double kb = b < nk ? sconst[b] : __longlong_as_double(0x7ff8000000000000ll);
    const uint32_t mods = (a & 3u) | (((a >> 2) & 3u) << 10);          // SRCX, PUSH, push slot << 10
    uint32_t pf = 0;
    t.h = TH_END;
    t.aux = b << 16;
    t.c = 0.0;
    if (op >= CL_ST && op < CL_NUM_OPCODES) {                          // slot forms, generic and register
        t.h = op >= CL_ST0 ? (op < CL_ADDS0 ? TH_ST : (op < CL_MULS0 ? TH_ADDS : TH_MULS))
                           : (op == CL_ST ? TH_ST : (op == CL_ADDS ? TH_ADDS : (op == CL_MULS ? TH_MULS : TH_POWS)));
        const uint32_t k = a;                                          // the register forms repeat the slot in `a`
        t.aux = k << 10;
        if (k >= CLOSYR_REG_SLOTS) pf = PF_DEEP | PF_NO_FAST;
        if (op == CL_POWS) pf |= PF_NO_FAST;
        return pf;
    }
    if (op >= CL_NEG && op <= CL_ABS) {
        const uint32_t f = closyr_th_of_function[op - CL_NEG];
        if (op == CL_POWC) {
            t.c = kb;
            if (nonfinite(t.c) || t.c == 0.0) pf = PF_NO_FAST;
        }
        if ((a & CL_M_SRCX) && have_tab && op >= CL_RECIP && op <= CL_ACOS && ((VM_TAB_MASK >> (op - CL_RECIP)) & 1u)) {
            // real on the whole line: sin cos exp tan 1/x; with a domain: sqrt x^-1/2 log asin acos
            const bool total = op == CL_SIN || op == CL_COS || op == CL_EXP || op == CL_TAN;
            const uint32_t push = (a & CL_M_PUSH) ? 1u : 0u;
            t.h = total ? TH_TAB + push : (op == CL_RECIP ? TH_TABR + push : TH_TABC);
            t.aux = (mods & ~1u) | (f << 2);
            t.c = __longlong_as_double((long long)(op - CL_RECIP) * tab_stride * 8);
        } else if (a & (CL_M_SRCX | CL_M_PUSH)) {
            t.h = TH_MOD;
            t.aux = mods | (f << 2) | (b << 16);
        } else t.h = f;
        return pf;
    }
    switch (op) {
    case CL_END: break;
    case CL_LDX: t.h = (a & CL_M_PUSH) ? TH_LDXP : TH_LDX; t.aux = mods & ~1u; break;
    case CL_LDC: t.h = (a & CL_M_PUSH) ? TH_LDCP : TH_LDC; t.aux = mods & ~1u; t.c = kb; break;
    case CL_LDZ: t.h = b + 1u < nk ? TH_LDZ : TH_LDSPEC; if (t.h == TH_LDSPEC) t.aux = 1u << 16; pf = PF_NO_FAST; break;
    case CL_LDSPEC: t.h = TH_LDSPEC; t.aux = a << 16; pf = PF_NO_FAST; break;
    case CL_ADDX: t.h = TH_ADDX; break;
    case CL_ADDC: t.h = TH_ADDC; t.c = kb; break;
    case CL_MULX: t.h = TH_MULX; break;
    case CL_MULC: t.h = TH_MULC; t.c = kb; break;
    default:                                        // unknown opcode: stays symbolic in Symja, never a real number
        t.h = TH_LDSPEC;
        t.aux = 1u << 16;
        pf = PF_NO_FAST;
        break;
    }
    return pf;
}

// ------------------------------------------------------------------------------------------------
// shared-memory access by 32-bit shared-window address (through a generic pointer the compiler rebuilds the
// window base at every access)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// x of this lane's points, loaded where an instruction needs it (L1-resident): with x in registers for the whole
// program the 64-register budget spilled it to local memory.  volatile: the compiler would otherwise hoist the
// loads of several handlers to the top of the dispatch loop.
__device__ __forceinline__ void ldg_x(const double *p, double (&x)[4])
{
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "l"(p));
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2+16];" : "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ double ldg_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// identity the optimiser cannot see through: keeps per-handler address arithmetic inside the handler
__device__ __forceinline__ uint32_t opaque(uint32_t v)
{
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ double opaque(double v)
{
    asm volatile("" : "+d"(v));
    return v;
}
__device__ __forceinline__ uint4 lds_insn(uint32_t addr)          // the threaded code is read-only while it runs
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------------------------------------
// P-wide real elementary functions shared by run_real, the real mode of run_cplx and leaf_table_kernel, so a
// value is bit-identical wherever it is computed
// ------------------------------------------------------------------------------------------------
#define CLOSYR_FORJ _Pragma("unroll") for (int j = 0; j < P; j++)

template <int P> __device__ __forceinline__ void f_recip(double (&a)[P]) { CLOSYR_FORJ a[j] = 1.0 / a[j]; }
template <int P> __device__ __forceinline__ void f_sqrt(double (&a)[P]) { CLOSYR_FORJ a[j] = sqrt(a[j]); }
template <int P> __device__ __forceinline__ void f_rsqrt(double (&a)[P]) { CLOSYR_FORJ a[j] = 1.0 / sqrt(a[j]); }
template <int P> __device__ __forceinline__ void f_tan(double (&a)[P]) { CLOSYR_FORJ a[j] = m_tan(a[j]); }
template <int P> __device__ __forceinline__ void f_asin(double (&a)[P]) { CLOSYR_FORJ a[j] = m_asin(a[j]); }
template <int P> __device__ __forceinline__ void f_acos(double (&a)[P]) { CLOSYR_FORJ a[j] = m_acos(a[j]); }
template <int P> __device__ __forceinline__ void f_powc(double (&a)[P], double c) { CLOSYR_FORJ a[j] = m_pow(a[j], c); }
// a^m by square-and-multiply, MSB first (the same chain the oracle walks); m >= 1; the loop is warp-uniform
template <int P> __device__ __forceinline__ void f_powu(double (&a)[P], unsigned m)
{
    double r[P];
    CLOSYR_FORJ r[j] = a[j];
#pragma unroll 1
    for (int b = 30 - __clz((int)m); b >= 0; b--) {
        CLOSYR_FORJ r[j] = r[j] * r[j];
        if ((m >> b) & 1u) { CLOSYR_FORJ r[j] = r[j] * a[j]; }
    }
    CLOSYR_FORJ a[j] = r[j];
}
template <int P> __device__ __forceinline__ uint32_t mask_nonfinite(const double (&a)[P])
{
    uint32_t m = 0;
    CLOSYR_FORJ if (nonfinite(a[j])) m |= 1u << j;
    return m;
}
// NaN as the arithmetic and the math routines produce it (quiet); infinities
template <int P> __device__ __forceinline__ uint32_t mask_nan(const double (&a)[P])
{
    uint32_t m = 0;
    CLOSYR_FORJ if (abs_hi(a[j]) >= 0x7ff80000u) m |= 1u << j;
    return m;
}
template <int P> __device__ __forceinline__ uint32_t mask_inf(const double (&a)[P])
{
    uint32_t m = 0;
    CLOSYR_FORJ if (is_inf(a[j])) m |= 1u << j;
    return m;
}
template <int P> __device__ __forceinline__ bool every_point_nan(const double (&a)[P])
{
    bool all = true;
    CLOSYR_FORJ all = all && (abs_hi(a[j]) >= 0x7ff80000u);
    return all;
}

// the tabulated functions of the raw variable, evaluated where no table exists (leaf_table_kernel fills the
// tables through this very function)
template <int P> __device__ __forceinline__ void f_tabulated(uint32_t op, double (&a)[P])
{
    switch (op) {
    case CL_RECIP: f_recip<P>(a); break;
    case CL_SQRT: f_sqrt<P>(a); break;
    case CL_RSQRT: f_rsqrt<P>(a); break;
    case CL_EXP: v_exp<P>(a); break;
    case CL_LOG: v_log<P>(a); break;
    case CL_SIN: v_sin<P>(a); break;
    case CL_COS: v_cos<P>(a); break;
    case CL_TAN: f_tan<P>(a); break;
    case CL_ASIN: f_asin<P>(a); break;
    default: f_acos<P>(a); break;
    }
}

// ------------------------------------------------------------------------------------------------
// fast path: P points per lane, IEEE real arithmetic, accumulator in registers, slots in shared memory
// ([slot][point][lane] doubles, conflict-free; `slot_addr` is this lane's column).
//
// Invariant kept per point (v = the IEEE value here, s = the value of the full algebra):
//   v finite, no report  <=>  s is a finite real and v == s;
//   v = +-Inf            =>   s is ComplexInfinity, a directed infinity or Indeterminate (never a finite number;
//                             the residual is max-resid either way, ops.clj:78-90)
//   v NaN                =>   either the point left a function's real domain somewhere (the algebra continues over the
//                             complex numbers: reported), or the NaN was made by arithmetic on infinities / NaNs and s is
//                             Indeterminate, which absorbs everything that follows (not reported: never a real number).
//                             NaN is absorbing through every instruction of this path, so either shows in the result.
// The returned mask has bit j set when point j must be re-evaluated with the complete algebra: its result is NaN after
// a domain exit, or an instruction that maps an infinity to a finite number (E^-Inf, Inf^-k) or whose result depends on
// the kind of infinity was handed one -- IEEE cannot tell ComplexInfinity from a directed infinity, nor from an
// Indeterminate the algebra reached earlier (ComplexInfinity + ComplexInfinity).
// Programs with PF_NO_FAST never come here.
// ------------------------------------------------------------------------------------------------
// A function of its own (not inlined into the kernel): the dispatch loop gets a register allocation of its own.
//
// The cheap instructions (leaf loads, table loads, + and * with x / a literal / a slot, stores, neg, square, abs:
// two thirds of all dispatches) are interpreted INSIDE one PTX block: LDS.128 of the threaded-code entry, brx.idx
// through a branch-target table, the handler, back to the top -- 5 instructions of dispatch per handler.  Written
// in C++, the same switch became a compare tree (NVVM has no jump tables) with register copies of the whole
// accumulator at the loop head: 25 instructions per dispatch.  The block exits for the elementary functions, which
// C++ handles below, and for END.
// What a warp's fast pass over one 128-point chunk returns: the values, the mask of reported points, and -- computed here
// because the interpreter's registers are free again at that point, while the calling kernel has none to spare -- the
// lane's largest residual; the chunk's residual sum goes straight to shared memory (the leaf of the reduction tree).
template <int P> struct RealResult { double v[P]; double mx; uint32_t ev; };

// residuals (ops.clj:78-90) without FP64 min / max (four instructions each on this chip):
// a = |y - f| is non-negative or NaN, so its bit pattern orders like its value with NaN on top
__device__ __forceinline__ double clamp_resid(double a)
{
    return ((uint32_t)__double2hiint(a) & 0x7fffffffu) >= 0x412e8480u ? CLOSYR_MAX_RESID : a;     // hi word of 1e6; its lo word is 0
}
__device__ __forceinline__ double max_nonneg(double a, double b)
{
    return (unsigned long long)__double_as_longlong(a) >= (unsigned long long)__double_as_longlong(b) ? a : b;
}
// residuals of a lane's 4 points against y, their left-fold sum reduced over the warp with the fixed xor butterfly and
// stored by lane 0 at `sum_addr` (shared); returns the lane's largest residual.  Whole warps only.
__device__ __forceinline__ double chunk_residuals(const double (&v)[4], const double *__restrict__ yp, uint32_t valid,
                                                  uint32_t sum_addr)
{
    double y[4];
    ldg_x(yp, y);                                          // ys is padded to whole chunks like xs
    double r0 = clamp_resid(fabs(y[0] - v[0])), r1 = clamp_resid(fabs(y[1] - v[1]));
    double r2 = clamp_resid(fabs(y[2] - v[2])), r3 = clamp_resid(fabs(y[3] - v[3]));
    if (valid != 0xfu) {                                   // the run's last chunk
        if (!(valid & 1u)) r0 = 0.0;
        if (!(valid & 2u)) r1 = 0.0;
        if (!(valid & 4u)) r2 = 0.0;
        if (!(valid & 8u)) r3 = 0.0;
    }
    double s = ((r0 + r1) + r2) + r3;                      // lane-local left fold (0.0 + r0 == r0)
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, st);
    uint32_t lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    if (lane == 0) sts_f64(sum_addr, s);
    return max_nonneg(max_nonneg(r0, r1), max_nonneg(r2, r3));
}

// The cheap handlers in PTX.  S = this copy's suffix, T = the other copy's: the loop is unrolled twice so that the
// entry of the NEXT instruction is already being loaded (into the other copy's registers) while this one runs.
// Operands: %0..%3 acc, %4 ip, %5..%8 = h aux c.lo c.hi of the instruction the block exits on, %9 this lane's slot
// column, %10..%13 x, %14 table base of this lane's points.
#define CLOSYR_PUSH_PTX(S)                   \
    "and.b32 t_, aux" S "_, 0xc00;\n\t"      \
    "add.u32 t_, t_, %9;\n\t"                \
    "st.shared.f64 [t_], %0;\n\t"            \
    "st.shared.f64 [t_+256], %1;\n\t"        \
    "st.shared.f64 [t_+512], %2;\n\t"        \
    "st.shared.f64 [t_+768], %3;\n\t"
#define CLOSYR_OP4(OP, B0, B1, B2, B3)       \
    OP " %0, %0, " B0 ";\n\t"                \
    OP " %1, %1, " B1 ";\n\t"                \
    OP " %2, %2, " B2 ";\n\t"                \
    OP " %3, %3, " B3 ";\n\t"
#define CLOSYR_CHEAP_HANDLERS(S, T, OFF)                                                          \
    /* leaf loads: the P forms push acc first and fall through */                             \
    "LDXP_" S ":\n\t" CLOSYR_PUSH_PTX(S)                                                       \
    "LDX_" S ":\n\t"                                                                           \
    "mov.f64 %0, %10;\n\t mov.f64 %1, %11;\n\t mov.f64 %2, %12;\n\t mov.f64 %3, %13;\n\t"       \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "LDCP_" S ":\n\t" CLOSYR_PUSH_PTX(S)                                                       \
    "LDC_" S ":\n\t"                                                                           \
    "mov.b64 %0, {lo" S "_, hi" S "_};\n\t mov.f64 %1, %0;\n\t mov.f64 %2, %0;\n\t mov.f64 %3, %0;\n\t" \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "TABP_" S ":\n\t" CLOSYR_PUSH_PTX(S)                                                       \
    "TAB_" S ":\n\t" /* c = byte offset of the function's row */                               \
    "mov.b64 ad_, {lo" S "_, hi" S "_};\n\t"                                                   \
    "add.s64 ad_, ad_, %14;\n\t"                                                               \
    "ld.global.nc.v2.f64 {%0, %1}, [ad_];\n\t"                                                 \
    "ld.global.nc.v2.f64 {%2, %3}, [ad_+16];\n\t"                                              \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    /* arithmetic with a leaf operand (.rn: never contracted into an FMA) */                   \
    "ADDX_" S ":\n\t" CLOSYR_OP4("add.rn.f64", "%10", "%11", "%12", "%13") "bra.uni NEXT_" T ";\n\t" \
    "MULX_" S ":\n\t" CLOSYR_OP4("mul.rn.f64", "%10", "%11", "%12", "%13") "bra.uni NEXT_" T ";\n\t" \
    "ADDC_" S ":\n\t"                                                                          \
    "mov.b64 c_, {lo" S "_, hi" S "_};\n\t"                                                    \
    CLOSYR_OP4("add.rn.f64", "c_", "c_", "c_", "c_") "bra.uni NEXT_" T ";\n\t"                 \
    "MULC_" S ":\n\t"                                                                          \
    "mov.b64 c_, {lo" S "_, hi" S "_};\n\t"                                                    \
    CLOSYR_OP4("mul.rn.f64", "c_", "c_", "c_", "c_") "bra.uni NEXT_" T ";\n\t"                 \
    "NEG_" S ":\n\t"                                                                           \
    "neg.f64 %0, %0;\n\t neg.f64 %1, %1;\n\t neg.f64 %2, %2;\n\t neg.f64 %3, %3;\n\t"          \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "SQR_" S ":\n\t" CLOSYR_OP4("mul.rn.f64", "%0", "%1", "%2", "%3") "bra.uni NEXT_" T ";\n\t" \
    "ABS_" S ":\n\t"                                                                           \
    "abs.f64 %0, %0;\n\t abs.f64 %1, %1;\n\t abs.f64 %2, %2;\n\t abs.f64 %3, %3;\n\t"          \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    /* slots: aux = slot << 10 = its byte offset in this lane's column */                      \
    "ST_" S ":\n\t"                                                                            \
    "add.u32 t_, aux" S "_, %9;\n\t"                                                           \
    "st.shared.f64 [t_], %0;\n\t st.shared.f64 [t_+256], %1;\n\t st.shared.f64 [t_+512], %2;\n\t st.shared.f64 [t_+768], %3;\n\t" \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "ADDS_" S ":\n\t"                                                                          \
    "add.u32 t_, aux" S "_, %9;\n\t"                                                           \
    "ld.shared.f64 s0_, [t_];\n\t ld.shared.f64 s1_, [t_+256];\n\t ld.shared.f64 s2_, [t_+512];\n\t ld.shared.f64 s3_, [t_+768];\n\t" \
    CLOSYR_OP4("add.rn.f64", "s0_", "s1_", "s2_", "s3_") "bra.uni NEXT_" T ";\n\t"             \
    "MULS_" S ":\n\t"                                                                          \
    "add.u32 t_, aux" S "_, %9;\n\t"                                                           \
    "ld.shared.f64 s0_, [t_];\n\t ld.shared.f64 s1_, [t_+256];\n\t ld.shared.f64 s2_, [t_+512];\n\t ld.shared.f64 s3_, [t_+768];\n\t" \
    CLOSYR_OP4("mul.rn.f64", "s0_", "s1_", "s2_", "s3_") "bra.uni NEXT_" T ";\n\t"             \
    /* dispatch of this copy: start loading the following entry into the other copy's registers, then jump */ \
    "NEXT_" S ":\n\t"                                                                          \
    "ld.shared.v4.b32 {h" T "_, aux" T "_, lo" T "_, hi" T "_}, [%4" OFF "];\n\t"                \
    "add.u32 %4, %4, 16;\n\t"                                                                  \
    "tbl" S "_: .branchtargets X_" S ", LDX_" S ", LDXP_" S ", LDC_" S ", LDCP_" S ", TAB_" S ", TABP_" S ", TAB_" S ", TABP_" S \
    ", ADDX_" S ", ADDC_" S ", MULX_" S ", MULC_" S ", NEG_" S ", SQR_" S ", ABS_" S ", ST_" S ", ADDS_" S ", MULS_" S       \
    ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S ", X_" S   \
    ", X_" S ", X_" S ", X_" S ", X_" S ";\n\t"                                                \
    "brx.idx.uni h" S "_, tbl" S "_;\n\t"                                                      \
    /* not a cheap instruction: hand it to the C++ part (ip already points at the following entry) */ \
    "X_" S ":\n\t"                                                                             \
    "mov.b32 %5, h" S "_;\n\t mov.b32 %6, aux" S "_;\n\t mov.b32 %7, lo" S "_;\n\t mov.b32 %8, hi" S "_;\n\t" \
    "bra.uni OUT_;\n\t"

// LEAN = ONE copy of the PTX block, without the prefetch of the next entry: for CTAs of fewer than 32 warps.  Several programs
// share an SM then and the launch is bound by the instruction cache the SMs of a GPC share (profiles/
// r02_design_experiments.md, "instruction supply"): 2.4 KB less code per expression are worth more than the prefetch
// (100K x 256 / 1024 points -2 %, one program per SM +4 %).
template <int P, bool LEAN>
__device__ __noinline__ RealResult<P> run_real(const uint32_t tcode_addr, const uint32_t slot_addr,
                                               const double *__restrict__ xp, const double *__restrict__ tab,
                                               const double *__restrict__ yp, const uint32_t valid, const uint32_t sum_addr)
{
    constexpr uint32_t ALL = (1u << P) - 1u;
    static_assert(P == 4, "slot stride and x loads are written for 4 points per lane");
    RealResult<P> res;
    double (&acc)[P] = res.v;
    uint32_t ev = 0;                 // points handed an infinity where its kind matters
    uint32_t dx = 0;                 // points that left a function's real domain (their NaN is not an Indeterminate)
    uint32_t ip = tcode_addr;
    double x[P];
    ldg_x(xp, x);
    CLOSYR_FORJ acc[j] = 0.0;
#define CLOSYR_PUSH_IF()                                                       \
    if (aux & 2u) {                                                            \
        const uint32_t sl_ = slot_addr + (aux & 0xc00u);                       \
        CLOSYR_FORJ sts_f64(sl_ + j * 256, acc[j]);                            \
    }
    // A function with a real domain: a NaN it makes out of a number is a domain exit (the algebra continues over the complex
    // numbers, only run_cplx knows the value); every other NaN of this path comes from arithmetic on infinities (Inf - Inf,
    // 0 * Inf, Sin(Inf) ...) or from a NaN, and stands for Indeterminate, which absorbs every later operation: such a point
    // is not a real number whatever follows (residual max-resid, ops.clj:78-90) and needs no second evaluation.
    // Points are sorted by x, so a domain exit usually takes whole chunks: once every point of the warp is NaN the rest of
    // the pass cannot change anything.
#define CLOSYR_DOMAIN(body_)                                                                        \
    {                                                                                               \
        const uint32_t pre_ = mask_nan<P>(acc);                                                     \
        body_;                                                                                      \
        dx |= mask_nan<P>(acc) & ~pre_;                                                             \
        if (__all_sync(0xffffffffu, every_point_nan<P>(acc))) goto finish;                          \
    }
#pragma unroll 1
    for (;;) {
        uint32_t h, aux, clo, chi;
        if (LEAN)
            asm volatile(
                "{\n\t"
                ".reg .b32 t_, hA_, auxA_, loA_, hiA_;\n\t"
                ".reg .b64 ad_;\n\t"
                ".reg .f64 c_, s0_, s1_, s2_, s3_;\n\t"
                "bra.uni NEXT_A;\n\t"
                CLOSYR_CHEAP_HANDLERS("A", "A", "")
                "OUT_:\n\t"
                "}\n"
                : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+r"(ip), "=&r"(h), "=&r"(aux), "=&r"(clo), "=&r"(chi)
                : "r"(slot_addr), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "l"(tab)
                : "memory");
        else
            asm volatile(
                "{\n\t"
                ".reg .b32 t_, hA_, auxA_, loA_, hiA_, hB_, auxB_, loB_, hiB_;\n\t"
                ".reg .b64 ad_;\n\t"
                ".reg .f64 c_, s0_, s1_, s2_, s3_;\n\t"
                "ld.shared.v4.b32 {hA_, auxA_, loA_, hiA_}, [%4];\n\t"
                "bra.uni NEXT_A;\n\t"
                CLOSYR_CHEAP_HANDLERS("A", "B", "+16")
                CLOSYR_CHEAP_HANDLERS("B", "A", "+16")
                "OUT_:\n\t"
                "}\n"
                : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+r"(ip), "=&r"(h), "=&r"(aux), "=&r"(clo), "=&r"(chi)
                : "r"(slot_addr), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "l"(tab)
                : "memory");
        const double c = __hiloint2double((int)chi, (int)clo);
    redo:
        switch (h) {
        case TH_END:
        finish: {
            res.ev = (ev | (dx & mask_nan<P>(acc))) & valid;
            res.mx = chunk_residuals(acc, yp, valid, sum_addr);     // overwritten by the slow path when a point was reported
            return res;
        }
        case TH_TABC: {
            // f(x) of the raw variable, tabulated per point by leaf_table_kernel with the same routines, for a function
            // with a domain.  Points are sorted by x, so a domain exit usually takes whole chunks: once every point of
            // the warp is lost the rest of the fast pass is wasted work (the complete algebra starts over anyway)
            CLOSYR_PUSH_IF()
            ldg_x(reinterpret_cast<const double *>(reinterpret_cast<const char *>(tab) + __double_as_longlong(c)), acc);
            dx |= mask_nan<P>(acc);                          // x is a number: a NaN entry is a domain exit
            if (__all_sync(0xffffffffu, every_point_nan<P>(acc))) goto finish;
            break;
        }
        case TH_MOD:                                         // PUSH / SRCX modifiers of a function that has no table
            CLOSYR_PUSH_IF()
            if (aux & 1u) { CLOSYR_FORJ acc[j] = x[j]; }
            h = (aux >> 2) & 0xffu;
            goto redo;
        case TH_NEG: CLOSYR_FORJ acc[j] = -acc[j]; break;      // reached through TH_MOD only
        case TH_SQR: CLOSYR_FORJ acc[j] = acc[j] * acc[j]; break;
        case TH_ABS: CLOSYR_FORJ acc[j] = fabs(acc[j]); break;
        case TH_RECIP: ev |= mask_inf<P>(acc); f_recip<P>(acc); break;
        case TH_SQRT: CLOSYR_DOMAIN(f_sqrt<P>(acc)) break;
        case TH_RSQRT: ev |= mask_inf<P>(acc); CLOSYR_DOMAIN(f_rsqrt<P>(acc)) break;
        case TH_POWI: {
            const int n = (int)(short)(aux >> 16);
            if (n < 0) {
                ev |= mask_inf<P>(acc);
                f_powu<P>(acc, (unsigned)-n);
                f_recip<P>(acc);
            } else f_powu<P>(acc, (unsigned)n);
            break;
        }
        case TH_POWC: ev |= mask_inf<P>(acc); CLOSYR_DOMAIN(f_powc<P>(acc, c)) break;
        case TH_EXP: ev |= mask_inf<P>(acc); v_exp<P>(acc); break;
        case TH_LOG: CLOSYR_DOMAIN(v_log<P>(acc)) break;
        case TH_SIN: v_sin<P>(acc); break;
        case TH_COS: v_cos<P>(acc); break;
        case TH_TAN: f_tan<P>(acc); break;
        case TH_ASIN: CLOSYR_DOMAIN(f_asin<P>(acc)) break;
        case TH_ACOS: CLOSYR_DOMAIN(f_acos<P>(acc)) break;
        default: { res.ev = ALL & valid; res.mx = 0.0; return res; }                 // LDZ LDSPEC POWS: PF_NO_FAST programs do not come here
        }
    }
#undef CLOSYR_DOMAIN
#undef CLOSYR_PUSH_IF
}
#undef CLOSYR_CHEAP_HANDLERS
#undef CLOSYR_OP4
#undef CLOSYR_PUSH_PTX

// ------------------------------------------------------------------------------------------------
// complete algebra, P points per lane.  (re, im) per point with the encoding of Z: Indeterminate = NaN real part,
// ComplexInfinity = (Inf, Inf), directed infinities = (+-Inf, 0).
//
// Real mode (cmo == 0): every value of the accumulator is a real number, a directed infinity or Indeterminate; the
// imaginary parts are not maintained.  IEEE arithmetic IS the algebra there for + and * (oracle v_add / v_mul on two
// T_REAL values), and every elementary function checks the condition under which its real routine is the algebra's
// value; the first point that fails it takes the whole warp to complex mode (cmo == TH_COUNT: the handler index is
// offset into the second half of the jump table) and the instruction is dispatched again there.  Leaf loads return
// to real mode.  `sm` has bit k set when slot k holds values with imaginary parts.
//
// Complex mode: + and * are the finite formulas of lean_add / lean_mul for all points at once; a point whose result
// is not finite (an operand was an infinity / Indeterminate, or the arithmetic overflowed) is redone by zadd / zmul.
// Elementary functions go point by point through the z* functions (they hold the case analysis of the algebra).
//
// Slots: [slot][re | im][point][lane] doubles in shared memory for slots 0..3 (the same bytes run_real uses: its
// values are dead by now), local memory beyond.
// ------------------------------------------------------------------------------------------------
template <int P> struct CAcc { double re[P], im[P]; };

// the chop of zmk on a finite result
__device__ __forceinline__ void c_chop(double &re, double &im)
{
    if (im != 0.0 && fabs(im) <= CLOSYR_CHOP_REL * fabs(re)) im = 0.0;
    else if (re != 0.0 && fabs(re) <= CLOSYR_CHOP_REL * fabs(im)) re = 0.0;
}
// acc = s + acc (s = the other operand; + commutes bit for bit, zadd is symmetric)
template <int P>
__device__ __forceinline__ void c_add(CAcc<P> &a, const double (&sre)[P], const double (&sim)[P])
{
    double rr[P], ri[P];
    uint32_t bad = 0;
    CLOSYR_FORJ {
        rr[j] = sre[j] + a.re[j];
        ri[j] = sim[j] + a.im[j];
        if (nonfinite(rr[j]) || nonfinite(ri[j])) bad |= 1u << j;
        c_chop(rr[j], ri[j]);
    }
    if (bad) {                                   // an operand is an infinity / Indeterminate, or the sum overflowed
        CLOSYR_FORJ if ((bad >> j) & 1u) {
            const Z r = zadd(zv(sre[j], sim[j]), zv(a.re[j], a.im[j]));
            rr[j] = r.re;
            ri[j] = r.im;
        }
    }
    CLOSYR_FORJ { a.re[j] = rr[j]; a.im[j] = ri[j]; }
}
// acc = s * acc with lean_mul's case split on the operands' imaginary parts (s first, as the scalar code has it)
template <int P>
__device__ __forceinline__ void c_mul(CAcc<P> &a, const double (&sre)[P], const double (&sim)[P])
{
    double rr[P], ri[P];
    uint32_t bad = 0;
    CLOSYR_FORJ {
        const double p1 = sre[j] * a.re[j], p2 = sim[j] * a.im[j], p3 = sre[j] * a.im[j], p4 = sim[j] * a.re[j];
        const bool s_real = sim[j] == 0.0, a_real = a.im[j] == 0.0;
        rr[j] = (s_real || a_real) ? p1 : p1 - p2;
        ri[j] = s_real ? p3 : (a_real ? p4 : p3 + p4);
        if (nonfinite(rr[j]) || nonfinite(ri[j])) bad |= 1u << j;
        c_chop(rr[j], ri[j]);
    }
    if (bad) {
        CLOSYR_FORJ if ((bad >> j) & 1u) {
            const Z r = zmul(zv(sre[j], sim[j]), zv(a.re[j], a.im[j]));
            rr[j] = r.re;
            ri[j] = r.im;
        }
    }
    CLOSYR_FORJ { a.re[j] = rr[j]; a.im[j] = ri[j]; }
}

#ifdef CLOSYR_TRACE
__device__ unsigned long long g_ztags[8];
#define CLOSYR_ZSTAT()                                                                                   \
    {                                                                                                    \
        const int t0_ = ztag(zv(a.re[0], a.im[0])), t1_ = ztag(zv(a.re[1], a.im[1]));                     \
        const unsigned both_ = __ballot_sync(0xffffffffu, t0_ == T_CPLX && t1_ == T_CPLX);                \
        const unsigned none_ = __ballot_sync(0xffffffffu, t0_ != T_CPLX && t1_ != T_CPLX);                \
        const unsigned real_ = __ballot_sync(0xffffffffu, t0_ == T_REAL && t1_ == T_REAL);                \
        if ((threadIdx.x & 31) == 0) {                                                                   \
            atomicAdd(&g_ztags[0], 1ull);                                                                \
            if (both_ == 0xffffffffu) atomicAdd(&g_ztags[1], 1ull);                                      \
            else if (none_ == 0xffffffffu) atomicAdd(&g_ztags[2], 1ull);                                 \
            else atomicAdd(&g_ztags[3], 1ull);                                                           \
            atomicAdd(&g_ztags[4], (unsigned long long)__popc(both_));                                   \
            atomicAdd(&g_ztags[5], (unsigned long long)__popc(none_));                                   \
            atomicAdd(&g_ztags[6], (unsigned long long)__popc(real_));                                   \
        }                                                                                                \
    }
#else
#define CLOSYR_ZSTAT()
#endif
#define CLOSYR_ZCALL(f)                                            \
    CLOSYR_ZSTAT()                                                 \
    CLOSYR_FORJ {                                                  \
        const Z r_ = f(zv(a.re[j], a.im[j]));                      \
        a.re[j] = r_.re;                                           \
        a.im[j] = r_.im;                                           \
    }

// real-mode integer power of the scalar interpreter: |n| by the multiplication chain, then 1/r, ComplexInfinity at 0
__device__ __noinline__ Z zpowi_any(Z a, int n)
{
    if (a.im == 0.0) {
        const double r = powi_real(a.re, n < 0 ? -n : n);
        if (n > 0) return zv(r, 0.0);
        if (r != 0.0) return zv(1.0 / r, 0.0);
        return zcinf();
    }
    return zpowi(a, n);
}
__device__ __noinline__ Z zrsqrt(Z a) { return zrecip(zsqrt(a)); }

// Evaluates the program on this lane's P points.  out[j] = the value when it is a real number (possibly +-Inf);
// nonreal = the mask of points whose value is NOT a real number (eval.clj:70-80 turns those into +Inf).
// `tab` = base of the f(x_i) tables (or nullptr), idx[j] = the point's index.  Must be called by whole warps.
template <int P> struct CplxArgs { int64_t idx[P]; };
template <int P> struct CplxResult { double out[P]; uint32_t nonreal, left_real; };   // left_real: the warp left real mode at some instruction

// The cheap handlers of REAL MODE in PTX, two points per lane (the structure of run_real's block; see there), but ONE copy
// without the prefetch of the next entry: this interpreter's code is cold when a CTA arrives (profiles/
// r02_design_experiments.md), the second copy cost more instruction-cache lines than the prefetch saved.
// Operands: %0 %1 the real parts, %2 ip, %3 sm (bit k = slot k holds imaginary parts), %4..%7 = h aux c.lo c.hi of the
// instruction the block exits on, %8 this lane's slot column, %9 %10 x, %11 %12 this lane's two columns of the tables.
// A slot operand with imaginary parts, the 1/x table (pole) and everything that is not cheap leave the block.
#define CLOSYR_PUSH2_PTX(S)                      \
    "bfe.u32 t_, aux" S "_, 10, 2;\n\t"          \
    "bfi.b32 %3, z_, %3, t_, 1;\n\t"             \
    "and.b32 t_, aux" S "_, 0xc00;\n\t"          \
    "add.u32 t_, t_, %8;\n\t"                    \
    "st.shared.f64 [t_], %0;\n\t"                \
    "st.shared.f64 [t_+256], %1;\n\t"
#define CLOSYR_OP2(OP, B0, B1) OP " %0, %0, " B0 ";\n\t" OP " %1, %1, " B1 ";\n\t"
#define CLOSYR_SLOT_REAL_OR_EXIT(S)              \
    "shr.u32 t_, aux" S "_, 10;\n\t"             \
    "shr.u32 t_, %3, t_;\n\t"                    \
    "and.b32 t_, t_, 1;\n\t"                     \
    "setp.ne.u32 p_, t_, 0;\n\t"                 \
    "@p_ bra.uni X_" S ";\n\t"                   \
    "add.u32 t_, aux" S "_, %8;\n\t"             \
    "ld.shared.f64 s0_, [t_];\n\t ld.shared.f64 s1_, [t_+256];\n\t"
#define CLOSYR_TRIG2_GUARD(S)                                                               \
    "mov.b64 {t_, u0_}, %0;\n\t mov.b64 {t_, u1_}, %1;\n\t"                                    \
    "and.b32 u0_, u0_, 0x7fffffff;\n\t and.b32 u1_, u1_, 0x7fffffff;\n\t max.u32 u0_, u0_, u1_;\n\t" \
    "setp.ge.u32 p_, u0_, 0x4197d784;\n\t vote.sync.any.pred p_, p_, 0xffffffff;\n\t @p_ bra.uni X_" S ";\n\t"
#define CLOSYR_CHEAP2_HANDLERS(S, T, OFF)                                                         \
    "LDXP_" S ":\n\t" CLOSYR_PUSH2_PTX(S)                                                      \
    "LDX_" S ":\n\t mov.f64 %0, %9;\n\t mov.f64 %1, %10;\n\t bra.uni NEXT_" T ";\n\t"          \
    "LDCP_" S ":\n\t" CLOSYR_PUSH2_PTX(S)                                                      \
    "LDC_" S ":\n\t mov.b64 %0, {lo" S "_, hi" S "_};\n\t mov.f64 %1, %0;\n\t bra.uni NEXT_" T ";\n\t" \
    "TABP_" S ":\n\t" CLOSYR_PUSH2_PTX(S)                                                      \
    "TAB_" S ":\n\t"                                                                           \
    "mov.b64 ad_, {lo" S "_, hi" S "_};\n\t"                                                   \
    "add.s64 a0_, ad_, %11;\n\t add.s64 a1_, ad_, %12;\n\t"                                    \
    "ld.global.nc.f64 %0, [a0_];\n\t ld.global.nc.f64 %1, [a1_];\n\t"                          \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "ADDX_" S ":\n\t" CLOSYR_OP2("add.rn.f64", "%9", "%10") "bra.uni NEXT_" T ";\n\t"          \
    "MULX_" S ":\n\t" CLOSYR_OP2("mul.rn.f64", "%9", "%10") "bra.uni NEXT_" T ";\n\t"          \
    "ADDC_" S ":\n\t mov.b64 c_, {lo" S "_, hi" S "_};\n\t" CLOSYR_OP2("add.rn.f64", "c_", "c_") "bra.uni NEXT_" T ";\n\t" \
    "MULC_" S ":\n\t mov.b64 c_, {lo" S "_, hi" S "_};\n\t" CLOSYR_OP2("mul.rn.f64", "c_", "c_") "bra.uni NEXT_" T ";\n\t" \
    "NEG_" S ":\n\t neg.f64 %0, %0;\n\t neg.f64 %1, %1;\n\t bra.uni NEXT_" T ";\n\t"           \
    "SQR_" S ":\n\t" CLOSYR_OP2("mul.rn.f64", "%0", "%1") "bra.uni NEXT_" T ";\n\t"            \
    "ABS_" S ":\n\t abs.f64 %0, %0;\n\t abs.f64 %1, %1;\n\t bra.uni NEXT_" T ";\n\t"           \
    "ST_" S ":\n\t"                                                                            \
    "shr.u32 t_, aux" S "_, 10;\n\t"                                                           \
    "bfi.b32 %3, z_, %3, t_, 1;\n\t"                                                           \
    "add.u32 t_, aux" S "_, %8;\n\t"                                                           \
    "st.shared.f64 [t_], %0;\n\t st.shared.f64 [t_+256], %1;\n\t"                              \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "ADDS_" S ":\n\t" CLOSYR_SLOT_REAL_OR_EXIT(S) CLOSYR_OP2("add.rn.f64", "s0_", "s1_") "bra.uni NEXT_" T ";\n\t" \
    "MULS_" S ":\n\t" CLOSYR_SLOT_REAL_OR_EXIT(S) CLOSYR_OP2("mul.rn.f64", "s0_", "s1_") "bra.uni NEXT_" T ";\n\t" \
    /* sin / cos of the accumulator in real mode: the operations of v_sin<2> / v_cos<2> (vmath.cuh), bit for bit; an argument  \
       the lean routine hands to the library (|x| >= 1e8, Inf, NaN: any lane, either point) leaves the block */               \
    /* 1 / acc in real mode (zero is a pole: ComplexInfinity, the C++ handler's business) */       \
    "RECIP_" S ":\n\t"                                                                         \
    "setp.eq.f64 p_, %0, 0d0000000000000000;\n\t setp.eq.or.f64 p_, %1, 0d0000000000000000, p_;\n\t" \
    "vote.sync.any.pred p_, p_, 0xffffffff;\n\t @p_ bra.uni X_" S ";\n\t"                      \
    "ld.const.f64 c_, [closyr_vm_c+280];\n\t div.rn.f64 %0, c_, %0;\n\t div.rn.f64 %1, c_, %1;\n\t" \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    /* exp of the accumulator: v_exp<2> bit for bit; |x| >= 700, Inf, NaN leave the block */    \
    "EXP_" S ":\n\t"                                                                           \
    "mov.b64 {t_, u0_}, %0;\n\t mov.b64 {t_, u1_}, %1;\n\t"                                    \
    "and.b32 u0_, u0_, 0x7fffffff;\n\t and.b32 u1_, u1_, 0x7fffffff;\n\t max.u32 u0_, u0_, u1_;\n\t" \
    "setp.ge.u32 p_, u0_, 0x4085e000;\n\t vote.sync.any.pred p_, p_, 0xffffffff;\n\t @p_ bra.uni X_" S ";\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+0];\n\t ld.const.f64 s0_, [closyr_vm_c+8];\n\t"             \
    "fma.rn.f64 ta_, %0, c_, s0_;\n\t fma.rn.f64 tb_, %1, c_, s0_;\n\t"                        \
    "sub.rn.f64 ka_, ta_, s0_;\n\t sub.rn.f64 kb_, tb_, s0_;\n\t"                              \
    "neg.f64 ka_, ka_;\n\t neg.f64 kb_, kb_;\n\t"                                              \
    "ld.const.f64 c_, [closyr_vm_c+16];\n\t fma.rn.f64 ra_, ka_, c_, %0;\n\t fma.rn.f64 rb_, kb_, c_, %1;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+24];\n\t fma.rn.f64 ra_, ka_, c_, ra_;\n\t fma.rn.f64 rb_, kb_, c_, rb_;\n\t" \
    "ld.const.f64 ka_, [closyr_vm_c+32];\n\t mov.f64 kb_, ka_;\n\t"                            \
    "ld.const.f64 c_, [closyr_vm_c+40];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+48];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+56];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+64];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+72];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+80];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+88];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+96];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+104];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+280];\n\t fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t" \
    "fma.rn.f64 ka_, ka_, ra_, c_;\n\t fma.rn.f64 kb_, kb_, rb_, c_;\n\t"                      \
    "mov.b64 {t_, u0_}, ta_;\n\t shl.b32 t_, t_, 20;\n\t mov.b64 {u1_, u0_}, ka_;\n\t add.s32 u0_, u0_, t_;\n\t mov.b64 %0, {u1_, u0_};\n\t" \
    "mov.b64 {t_, u0_}, tb_;\n\t shl.b32 t_, t_, 20;\n\t mov.b64 {u1_, u0_}, kb_;\n\t add.s32 u0_, u0_, t_;\n\t mov.b64 %1, {u1_, u0_};\n\t" \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "SIN_" S ":\n\t" CLOSYR_TRIG2_GUARD(S)                                                     \
    "ld.const.f64 c_, [closyr_vm_c+112];\n\t ld.const.f64 s0_, [closyr_vm_c+8];\n\t"           \
    "fma.rn.f64 ta_, %0, c_, s0_;\n\t fma.rn.f64 tb_, %1, c_, s0_;\n\t"                        \
    "sub.rn.f64 ka_, ta_, s0_;\n\t sub.rn.f64 kb_, tb_, s0_;\n\t"                              \
    "bra.uni TRIG_" S ";\n\t"                                                                  \
    "COS_" S ":\n\t" CLOSYR_TRIG2_GUARD(S)                                                     \
    "ld.const.f64 c_, [closyr_vm_c+112];\n\t ld.const.f64 s0_, [closyr_vm_c+8];\n\t ld.const.f64 s1_, [closyr_vm_c+256];\n\t" \
    "fma.rn.f64 ta_, %0, c_, s1_;\n\t fma.rn.f64 tb_, %1, c_, s1_;\n\t"                        \
    "add.rn.f64 ta_, ta_, s0_;\n\t add.rn.f64 tb_, tb_, s0_;\n\t"                              \
    "sub.rn.f64 ka_, ta_, s0_;\n\t sub.rn.f64 kb_, tb_, s0_;\n\t"                              \
    "sub.rn.f64 ka_, ka_, s1_;\n\t sub.rn.f64 kb_, kb_, s1_;\n\t"                              \
    "TRIG_" S ":\n\t" /* r = x - k pi (three terms), sin(r) = r + r z Q(z), sign by the parity of k (low word of t) */ \
    "neg.f64 ka_, ka_;\n\t neg.f64 kb_, kb_;\n\t"                                              \
    "ld.const.f64 c_, [closyr_vm_c+120];\n\t fma.rn.f64 ra_, ka_, c_, %0;\n\t fma.rn.f64 rb_, kb_, c_, %1;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+128];\n\t fma.rn.f64 ra_, ka_, c_, ra_;\n\t fma.rn.f64 rb_, kb_, c_, rb_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+184];\n\t fma.rn.f64 ra_, ka_, c_, ra_;\n\t fma.rn.f64 rb_, kb_, c_, rb_;\n\t" \
    "mul.rn.f64 za_, ra_, ra_;\n\t mul.rn.f64 zb_, rb_, rb_;\n\t"                              \
    "ld.const.f64 ka_, [closyr_vm_c+248];\n\t mov.f64 kb_, ka_;\n\t"                           \
    "ld.const.f64 c_, [closyr_vm_c+240];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+232];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+224];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+216];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+208];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+200];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "ld.const.f64 c_, [closyr_vm_c+192];\n\t fma.rn.f64 ka_, ka_, za_, c_;\n\t fma.rn.f64 kb_, kb_, zb_, c_;\n\t" \
    "mul.rn.f64 ka_, ka_, za_;\n\t mul.rn.f64 kb_, kb_, zb_;\n\t"                              \
    "fma.rn.f64 ra_, ka_, ra_, ra_;\n\t fma.rn.f64 rb_, kb_, rb_, rb_;\n\t"                    \
    "mov.b64 {t_, u0_}, ta_;\n\t shl.b32 t_, t_, 31;\n\t mov.b64 {u1_, u0_}, ra_;\n\t xor.b32 u0_, u0_, t_;\n\t mov.b64 %0, {u1_, u0_};\n\t" \
    "mov.b64 {t_, u0_}, tb_;\n\t shl.b32 t_, t_, 31;\n\t mov.b64 {u1_, u0_}, rb_;\n\t xor.b32 u0_, u0_, t_;\n\t mov.b64 %1, {u1_, u0_};\n\t" \
    "bra.uni NEXT_" T ";\n\t"                                                                  \
    "NEXT_" S ":\n\t"                                                                          \
    "ld.shared.v4.b32 {h" T "_, aux" T "_, lo" T "_, hi" T "_}, [%2" OFF "];\n\t"                \
    "add.u32 %2, %2, 16;\n\t"                                                                  \
    "tbl" S "_: .branchtargets X_" S ", LDX_" S ", LDXP_" S ", LDC_" S ", LDCP_" S ", TAB_" S ", TABP_" S ", X_" S ", X_" S \
    ", ADDX_" S ", ADDC_" S ", MULX_" S ", MULC_" S ", NEG_" S ", SQR_" S ", ABS_" S ", ST_" S ", ADDS_" S ", MULS_" S      \
    ", X_" S ", X_" S ", RECIP_" S ", X_" S ", X_" S ", X_" S ", X_" S ", EXP_" S ", X_" S ", SIN_" S ", COS_" S ", X_" S ", X_" S \
    ", X_" S ", X_" S ", X_" S ", X_" S ";\n\t"                                                \
    "brx.idx.uni h" S "_, tbl" S "_;\n\t"                                                      \
    "X_" S ":\n\t"                                                                             \
    "mov.b32 %4, h" S "_;\n\t mov.b32 %5, aux" S "_;\n\t mov.b32 %6, lo" S "_;\n\t mov.b32 %7, hi" S "_;\n\t" \
    "bra.uni OUT_;\n\t"

// DEEP = the program uses spill slots >= CLOSYR_REG_SLOTS (PF_DEEP): those live in local memory and the PTX block is
// not used.  Ordinary programs take the other instantiation, which has neither.
template <int P, bool DEEP>
__device__ __noinline__ CplxResult<P> run_cplx(const uint32_t tcode_addr, const uint32_t slot_addr,
                                               const double *__restrict__ sconst, const double *__restrict__ xs,
                                               const CplxArgs<P> args, const double *__restrict__ tab)
{
    const int64_t (&idx)[P] = args.idx;
    constexpr unsigned FULLW = 0xffffffffu;
    constexpr uint32_t SLOT_SHIFT = 10;                     // 1 KiB per slot, as run_real<4>: [re|im][P points][32 lanes]
    constexpr uint32_t IM_OFF = 512;
    constexpr int NDEEP = DEEP ? CLOSYR_MAX_SLOTS - CLOSYR_REG_SLOTS : 1;
    static_assert(P == 2, "slot layout and the PTX block are written for 2 points per lane");
    CAcc<P> a;
    Z deep[NDEEP][P];                                        // local memory (DEEP only)
    uint32_t cmo = 0;                                        // 0 real mode, TH_COUNT complex mode
    uint32_t sm = 0;                                         // slots holding imaginary parts
    uint32_t left = 0;                                       // complex mode was entered at least once
    uint32_t ip = tcode_addr;
    double x[P];
    CLOSYR_FORJ { x[j] = ldg_f64(xs + idx[j]); a.re[j] = 0.0; a.im[j] = 0.0; }
    const double *const tp0 = tab + idx[0], *const tp1 = tab + idx[1];      // never dereferenced without tables
#define CLOSYR_TO_COMPLEX() { CLOSYR_FORJ a.im[j] = 0.0; cmo = TH_COUNT; left = 1u; }
#define CLOSYR_STORE(k_)                                                                    \
    {                                                                                       \
        const uint32_t k = (k_);                                                            \
        if (!DEEP || k < CLOSYR_REG_SLOTS) {                                                \
            const uint32_t sl_ = slot_addr + (k << SLOT_SHIFT);                             \
            CLOSYR_FORJ sts_f64(sl_ + j * 256, a.re[j]);                                    \
            if (cmo) { CLOSYR_FORJ sts_f64(sl_ + IM_OFF + j * 256, a.im[j]); sm |= 1u << k; } \
            else sm &= ~(1u << k);                                                          \
        } else {                                                                            \
            CLOSYR_FORJ deep[DEEP ? k - CLOSYR_REG_SLOTS : 0][j] = zv(a.re[j], cmo ? a.im[j] : 0.0); \
        }                                                                                   \
    }
#define CLOSYR_PUSH_IF() { const uint32_t ax_ = AUX; if (ax_ & 2u) CLOSYR_STORE((ax_ >> 10) & 3u) }
#define CLOSYR_LOADX(dst) CLOSYR_FORJ dst[j] = x[j];
    // loads slot k into (sre, sim); `scx` = the slot may hold imaginary parts
#define CLOSYR_LOAD(k_)                                                                     \
    double sre[P], sim[P];                                                                  \
    bool scx;                                                                               \
    {                                                                                       \
        const uint32_t k = (k_);                                                            \
        if (!DEEP || k < CLOSYR_REG_SLOTS) {                                                \
            const uint32_t sl_ = slot_addr + (k << SLOT_SHIFT);                             \
            scx = (sm >> k) & 1u;                                                           \
            CLOSYR_FORJ sre[j] = lds_f64(sl_ + j * 256);                                    \
            if (scx) { CLOSYR_FORJ sim[j] = lds_f64(sl_ + IM_OFF + j * 256); }              \
            else { CLOSYR_FORJ sim[j] = 0.0; }                                              \
        } else {                                                                            \
            scx = true;                                                                     \
            CLOSYR_FORJ { sre[j] = deep[DEEP ? k - CLOSYR_REG_SLOTS : 0][j].re; sim[j] = deep[DEEP ? k - CLOSYR_REG_SLOTS : 0][j].im; } \
        }                                                                                   \
    }
    // real-mode elementary function: `cond` (per point, on x_ = a.re[j]) says the real routine is the algebra's value
#define CLOSYR_REALF(cond, body)                                                            \
    {                                                                                       \
        bool ok_ = true;                                                                    \
        CLOSYR_FORJ { const double x_ = a.re[j]; ok_ = ok_ && (cond); }                     \
        if (__all_sync(FULLW, ok_)) { body; break; }                                        \
        CLOSYR_TO_COMPLEX()                                                                 \
        h += TH_COUNT;                                                                      \
        goto redo;                                                                          \
    }
    // AUX / CC: the operand word / literal as the handler sees it.  Through an opaque copy, otherwise the compiler hoists
    // every handler's decoding of them (slot masks, exponents, table addresses) to the head of the dispatch loop:
    // 42 instructions per dispatch in r02b's profile
#define AUX opaque(aux)
#define CC opaque(c)
#pragma unroll 1
    for (;;) {
        uint32_t h, aux, clo, chi;
        if (cmo) {
            // back to real mode as soon as no point of the warp has an imaginary part (a square, an absolute value, a chopped
            // sum): that IS real mode's invariant (reals, directed infinities, Indeterminate = NaN with a zero imaginary part;
            // ComplexInfinity has one), and half of the complex-mode function calls of the GA mix had such operands
            bool real_ = true;
            CLOSYR_FORJ real_ = real_ && a.im[j] == 0.0;
            if (__all_sync(FULLW, real_)) cmo = 0;
        }
        if (!DEEP && cmo == 0) {
            asm volatile(
                "{\n\t"
                ".reg .pred p_;\n\t"
                ".reg .b32 t_, z_, hA_, auxA_, loA_, hiA_, hB_, auxB_, loB_, hiB_;\n\t"
                ".reg .b64 ad_, a0_, a1_;\n\t"
                ".reg .f64 c_, s0_, s1_, ta_, tb_, ka_, kb_, ra_, rb_, za_, zb_;\n\t"
                ".reg .b32 u0_, u1_;\n\t"
                "mov.b32 z_, 0;\n\t"
                "bra.uni NEXT_A;\n\t"
                CLOSYR_CHEAP2_HANDLERS("A", "A", "")
                "OUT_:\n\t"
                "}\n"
                : "+d"(a.re[0]), "+d"(a.re[1]), "+r"(ip), "+r"(sm), "=&r"(h), "=&r"(aux), "=&r"(clo), "=&r"(chi)
                : "r"(slot_addr), "d"(x[0]), "d"(x[1]), "l"(tp0), "l"(tp1)
                : "memory");
        } else {
            const uint4 nx = lds_insn(ip);
            h = nx.x + cmo;
            aux = nx.y;
            clo = nx.z;
            chi = nx.w;
            ip += 16;
        }
        const double c = __hiloint2double((int)chi, (int)clo);
    redo:
        switch (h) {
        case TH_END:
        case TH_END + TH_COUNT: {
            CplxResult<P> res;
            res.nonreal = 0;
            res.left_real = left;
            CLOSYR_FORJ {
                const bool real = a.re[j] == a.re[j] && (cmo == 0 || a.im[j] == 0.0);
                if (!real) res.nonreal |= 1u << j;
                res.out[j] = a.re[j];
            }
            return res;
        }
        // ---- leaf loads: back to real mode
        case TH_LDX:
        case TH_LDXP:
        case TH_LDX + TH_COUNT:
        case TH_LDXP + TH_COUNT: CLOSYR_PUSH_IF() CLOSYR_LOADX(a.re) cmo = 0; break;
        case TH_LDC:
        case TH_LDCP:
        case TH_LDC + TH_COUNT:
        case TH_LDCP + TH_COUNT: CLOSYR_PUSH_IF() CLOSYR_FORJ a.re[j] = c; cmo = 0; break;
        case TH_TAB:
        case TH_TABP:
        case TH_TABR:
        case TH_TABRP:
        case TH_TABC:
        case TH_TAB + TH_COUNT:
        case TH_TABP + TH_COUNT:
        case TH_TABR + TH_COUNT:
        case TH_TABRP + TH_COUNT:
        case TH_TABC + TH_COUNT: {
            CLOSYR_PUSH_IF()
            const double *row = reinterpret_cast<const double *>(reinterpret_cast<const char *>(tab) + __double_as_longlong(opaque(c)));
            double t[P];
            bool fin = true;
            CLOSYR_FORJ { t[j] = row[idx[j]]; fin = fin && !nonfinite(t[j]); }
            cmo = 0;
            if (__all_sync(FULLW, fin)) { CLOSYR_FORJ a.re[j] = t[j]; break; }
            // some point is off the function's real domain (or at a pole): the function itself, complex mode
            CLOSYR_LOADX(a.re)
            CLOSYR_TO_COMPLEX()
            h = ((AUX >> 2) & 0xffu) + TH_COUNT;
            goto redo;
        }
        case TH_MOD:
        case TH_MOD + TH_COUNT:
            CLOSYR_PUSH_IF()
            if (AUX & 1u) { CLOSYR_LOADX(a.re) cmo = 0; }
            h = ((AUX >> 2) & 0xffu) + cmo;
            goto redo;
        case TH_LDZ:
        case TH_LDZ + TH_COUNT: {
            const Z z = zmk(sconst[AUX >> 16], sconst[(AUX >> 16) + 1]);
            CLOSYR_FORJ { a.re[j] = z.re; a.im[j] = z.im; }
            cmo = TH_COUNT;
            left = 1u;
            break;
        }
        case TH_LDSPEC:
        case TH_LDSPEC + TH_COUNT: {
            const Z z = (AUX >> 16) == 0 ? zcinf() : zindet();
            CLOSYR_FORJ { a.re[j] = z.re; a.im[j] = z.im; }
            cmo = TH_COUNT;
            left = 1u;
            break;
        }
        // ---- real mode: IEEE arithmetic is the algebra
        case TH_ADDX: CLOSYR_FORJ a.re[j] = a.re[j] + x[j]; break;
        case TH_ADDC: CLOSYR_FORJ a.re[j] = a.re[j] + c; break;
        case TH_MULX: CLOSYR_FORJ a.re[j] = a.re[j] * x[j]; break;
        case TH_MULC: CLOSYR_FORJ a.re[j] = a.re[j] * c; break;
        case TH_NEG: CLOSYR_FORJ a.re[j] = -a.re[j]; break;
        case TH_SQR: CLOSYR_FORJ a.re[j] = a.re[j] * a.re[j]; break;
        case TH_ABS: CLOSYR_FORJ a.re[j] = fabs(a.re[j]); break;
        case TH_ST:
        case TH_ST + TH_COUNT: CLOSYR_STORE(AUX >> 10) break;
        case TH_ADDS:
        case TH_MULS: {
            const uint32_t k = AUX >> 10;
            const bool scx = (DEEP && k >= CLOSYR_REG_SLOTS) || ((sm >> k) & 1u);
            if (scx) {
                CLOSYR_TO_COMPLEX()
                h += TH_COUNT;
                goto redo;
            }
            const uint32_t sl = slot_addr + (k << SLOT_SHIFT);
            if (h == TH_ADDS) { CLOSYR_FORJ a.re[j] = lds_f64(sl + j * 256) + a.re[j]; }
            else { CLOSYR_FORJ a.re[j] = lds_f64(sl + j * 256) * a.re[j]; }
            break;
        }
        case TH_RECIP: CLOSYR_REALF(x_ != 0.0, f_recip<P>(a.re))
        case TH_SQRT: CLOSYR_REALF(!(x_ < 0.0), f_sqrt<P>(a.re))
        case TH_RSQRT: CLOSYR_REALF(!(x_ <= 0.0), f_rsqrt<P>(a.re))
        case TH_POWI: {
            const int n = (int)(short)(AUX >> 16);
            double r[P];
            CLOSYR_FORJ r[j] = a.re[j];
            f_powu<P>(r, (unsigned)(n < 0 ? -n : n));
            if (n > 0) { CLOSYR_FORJ a.re[j] = r[j]; break; }
            bool ok = true;
            CLOSYR_FORJ ok = ok && (r[j] != 0.0);
            if (__all_sync(FULLW, ok)) { CLOSYR_FORJ a.re[j] = 1.0 / r[j]; break; }
            CLOSYR_TO_COMPLEX()
            h += TH_COUNT;
            goto redo;
        }
        case TH_POWC: CLOSYR_REALF(x_ > 0.0 && !is_inf(x_) && !nonfinite(CC) && CC != 0.0, f_powc<P>(a.re, c))
        case TH_EXP: v_exp<P>(a.re); break;
        case TH_LOG: CLOSYR_REALF(!(x_ <= 0.0), v_log<P>(a.re))
        case TH_SIN: v_sin<P>(a.re); break;
        case TH_COS: v_cos<P>(a.re); break;
        case TH_TAN: f_tan<P>(a.re); break;
        case TH_ASIN: CLOSYR_REALF(!(fabs(x_) > 1.0), f_asin<P>(a.re))
        case TH_ACOS: CLOSYR_REALF(!(fabs(x_) > 1.0), f_acos<P>(a.re))
        case TH_POWS: CLOSYR_TO_COMPLEX() h += TH_COUNT; goto redo;
        // ---- complex mode
        case TH_ADDX + TH_COUNT: { double zr[P]; CLOSYR_FORJ zr[j] = 0.0; c_add<P>(a, x, zr); break; }
        case TH_MULX + TH_COUNT: { double zr[P]; CLOSYR_FORJ zr[j] = 0.0; c_mul<P>(a, x, zr); break; }
        case TH_ADDC + TH_COUNT: { double cr[P], zr[P]; CLOSYR_FORJ { cr[j] = c; zr[j] = 0.0; } c_add<P>(a, cr, zr); break; }
        case TH_MULC + TH_COUNT: { double cr[P], zr[P]; CLOSYR_FORJ { cr[j] = c; zr[j] = 0.0; } c_mul<P>(a, cr, zr); break; }
        case TH_NEG + TH_COUNT: CLOSYR_ZCALL(zneg) break;
        case TH_SQR + TH_COUNT: { double sr[P], si[P]; CLOSYR_FORJ { sr[j] = a.re[j]; si[j] = a.im[j]; } c_mul<P>(a, sr, si); break; }
        case TH_ABS + TH_COUNT: CLOSYR_ZCALL(zabs) break;
        case TH_ADDS + TH_COUNT: { CLOSYR_LOAD(AUX >> 10) (void)scx; c_add<P>(a, sre, sim); break; }
        case TH_MULS + TH_COUNT: { CLOSYR_LOAD(AUX >> 10) (void)scx; c_mul<P>(a, sre, sim); break; }
        case TH_POWS + TH_COUNT: {
            CLOSYR_LOAD(AUX >> 10)
            (void)scx;
            CLOSYR_FORJ {
                const Z r = zpow(zv(sre[j], sim[j]), zv(a.re[j], a.im[j]));
                a.re[j] = r.re;
                a.im[j] = r.im;
            }
            break;
        }
        case TH_RECIP + TH_COUNT: CLOSYR_ZCALL(zrecip) break;
        case TH_SQRT + TH_COUNT: CLOSYR_ZCALL(zsqrt) break;
        case TH_RSQRT + TH_COUNT: CLOSYR_ZCALL(zrsqrt) break;
        case TH_POWI + TH_COUNT: {
            const int n = (int)(short)(AUX >> 16);
            CLOSYR_FORJ {
                const Z r = zpowi_any(zv(a.re[j], a.im[j]), n);
                a.re[j] = r.re;
                a.im[j] = r.im;
            }
            break;
        }
        case TH_POWC + TH_COUNT: {
            const Z w = zmk(c, 0.0);
            CLOSYR_FORJ {
                const Z r = zpow(zv(a.re[j], a.im[j]), w);
                a.re[j] = r.re;
                a.im[j] = r.im;
            }
            break;
        }
        case TH_EXP + TH_COUNT: CLOSYR_ZCALL(zexp) break;
        case TH_LOG + TH_COUNT: CLOSYR_ZCALL(zlog) break;
        case TH_SIN + TH_COUNT: CLOSYR_ZCALL(zsin) break;
        case TH_COS + TH_COUNT: CLOSYR_ZCALL(zcos) break;
        case TH_TAN + TH_COUNT: CLOSYR_ZCALL(ztan) break;
        case TH_ASIN + TH_COUNT: CLOSYR_ZCALL(zasin) break;
        case TH_ACOS + TH_COUNT: CLOSYR_ZCALL(zacos) break;
        default: {                                           // cannot happen: decode_insn emits the handlers above only
            CLOSYR_FORJ { a.re[j] = __longlong_as_double(0x7ff8000000000000ll); a.im[j] = 0.0; }
            cmo = TH_COUNT;
            break;
        }
        }
    }
#undef AUX
#undef CC
#undef CLOSYR_REALF
#undef CLOSYR_LOADX
#undef CLOSYR_LOAD
#undef CLOSYR_PUSH_IF
#undef CLOSYR_STORE
#undef CLOSYR_TO_COMPLEX
}
#undef CLOSYR_CHEAP2_HANDLERS
#undef CLOSYR_TRIG2_GUARD
#undef CLOSYR_SLOT_REAL_OR_EXIT
#undef CLOSYR_OP2
#undef CLOSYR_PUSH2_PTX
#undef CLOSYR_ZCALL

}  // namespace dev
}  // namespace closyr
