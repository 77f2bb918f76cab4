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
// time: the whole kernel image has to stay close to the instruction cache (32 KB L1.5 per SM).  Each libdevice
// routine therefore exists exactly once, behind a call, instead of being inlined at every use
// (inlined per point and per interpreter the kernel was 300 KB of SASS and stalled on instruction fetch).
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ double k_sin(double x) { return sin(x); }
__device__ __noinline__ double k_cos(double x) { return cos(x); }
__device__ __noinline__ double k_tan(double x) { return tan(x); }
__device__ __noinline__ double k_exp(double x) { return exp(x); }
__device__ __noinline__ double k_log(double x) { return log(x); }
__device__ __noinline__ double k_asin(double x) { return asin(x); }
__device__ __noinline__ double k_acos(double x) { return acos(x); }
__device__ __noinline__ double k_pow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double k_sinh(double x) { return sinh(x); }
__device__ __noinline__ double k_cosh(double x) { return cosh(x); }
__device__ __noinline__ double k_hypot(double x, double y) { return hypot(x, y); }
__device__ __noinline__ double k_atan2(double y, double x) { return atan2(y, x); }

// integer power by square-and-multiply, MSB first (the same chain the oracle walks)
__device__ __forceinline__ double powi_real(double a, int n)
{
    unsigned m = (unsigned)(n < 0 ? -n : n);
    int top = 31 - __clz((int)m);
    double r = a;
    for (int b = top - 1; b >= 0; b--) {
        r = r * r;
        if ((m >> b) & 1u) r = r * a;
    }
    return n < 0 ? 1.0 / r : r;
}
__device__ __noinline__ double k_recip(double x) { return 1.0 / x; }
__device__ __noinline__ double k_sqrt(double x) { return sqrt(x); }
__device__ __noinline__ double k_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __noinline__ double k_powi(double x, int n) { return powi_real(x, n); }

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
    return zmk(e * k_cos(a.im), e * k_sin(a.im));
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
    return zmk(k_sin(a.re) * k_cosh(a.im), k_cos(a.re) * k_sinh(a.im));
}

__device__ __noinline__ Z zcos(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_cos(a.re), 0.0);
    return zmk(k_cos(a.re) * k_cosh(a.im), -(k_sin(a.re) * k_sinh(a.im)));
}

__device__ __noinline__ Z ztan(Z a)
{
    int t = ztag(a);
    if (t == T_INDET || t == T_CINF) return zindet();
    if (t == T_REAL) return zmk(k_tan(a.re), 0.0);
    if (a.im > 20.0) return zv(0.0, 1.0);
    if (a.im < -20.0) return zv(0.0, -1.0);
    double a2 = 2.0 * a.re, b2 = 2.0 * a.im;
    double d = k_cos(a2) + k_cosh(b2);
    return zmk(k_sin(a2) / d, k_sinh(b2) / d);
}

__device__ __forceinline__ Z zmul_i(Z a) { return zmk(-a.im, a.re); }
__device__ __forceinline__ Z zmul_minus_i(Z a) { return zmk(a.im, -a.re); }
__device__ __forceinline__ Z zsqrt1z(Z z) { return zsqrt(zadd(zv(1.0, 0.0), zneg(zmul(z, z)))); }

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
    Z w = zlog(zadd(zsqrt1z(a), zmul_i(a)));
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
    Z w = zlog(zadd(a, zmul_i(s)));
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
        r = zmul(r, r);
        if ((m >> b) & 1u) r = zmul(r, a);
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
    return zexp(zmul(w, zlog(z)));
}

// ------------------------------------------------------------------------------------------------
// complete value algebra, P points per lane.  Values are (re, im) register pairs; im == 0 means "on the real
// line" (including directed infinities; Indeterminate is re = NaN), and for such operands every instruction
// takes the same IEEE instruction sequence as run_real, inline.  Anything else (complex numbers,
// ComplexInfinity, poles, domain exits) goes through the scalar z* functions above.  `slow` collects whether
// any point left the inline path, so the caller can drop back to run_real for the next chunk.
// ------------------------------------------------------------------------------------------------
#define CLOSYR_FORJ _Pragma("unroll") for (int j = 0; j < P; j++)
#define CLOSYR_Z(j) zv(re[j], im[j])
#define CLOSYR_SETZ(j, expr) { const Z z_ = (expr); re[j] = z_.re; im[j] = z_.im; slow = 1u; }
// unary instruction: `cond` says the point may take the real instruction sequence `fast`
#define CLOSYR_UNARY(cond, fast, slowf)                                           \
    CLOSYR_FORJ {                                                                 \
        const double a_ = re[j];                                                  \
        if (im[j] == 0.0 && (cond)) re[j] = (fast);                               \
        else CLOSYR_SETZ(j, slowf(CLOSYR_Z(j)))                                   \
    }

template <int P>
__device__ __forceinline__ uint32_t run_full(const uint32_t *__restrict__ scode, const double *__restrict__ sconst,
                                             const double (&x)[P], double (&re)[P], double (&im)[P])
{
    double s0r[P], s0i[P], s1r[P], s1i[P], s2r[P], s2i[P], s3r[P], s3i[P];
    Z ls[CLOSYR_MAX_SLOTS - CLOSYR_REG_SLOTS][P];       // deep spill slots (local memory, rarely touched)
    uint32_t slow = 0;
    int pc = 0;
    uint32_t w = scode[0];
    CLOSYR_FORJ {
        re[j] = 0.0; im[j] = 0.0;
        s0r[j] = s0i[j] = s1r[j] = s1i[j] = s2r[j] = s2i[j] = s3r[j] = s3i[j] = 0.0;
    }
#define CLOSYR_ST(sr, si) CLOSYR_FORJ { sr[j] = re[j]; si[j] = im[j]; }
#define CLOSYR_BIN(sr, si, op, slowf)                                             \
    CLOSYR_FORJ {                                                                 \
        if (im[j] == 0.0 && si[j] == 0.0) re[j] = sr[j] op re[j];                 \
        else CLOSYR_SETZ(j, slowf(zv(sr[j], si[j]), CLOSYR_Z(j)))                 \
    }
    for (;;) {
        const uint32_t op = w & 0xffu, a = (w >> 8) & 0xffu;
        const int b = (int)(w >> 16);
        w = scode[++pc];
        switch (op) {
        case CL_END: return slow;
        case CL_LDX: CLOSYR_FORJ { re[j] = x[j]; im[j] = 0.0; } break;
        case CL_LDC: { const double c = sconst[b]; CLOSYR_FORJ { re[j] = c; im[j] = 0.0; } break; }
        case CL_LDZ: { const Z c = zmk(sconst[b], sconst[b + 1]); CLOSYR_FORJ { re[j] = c.re; im[j] = c.im; } slow = 1u; break; }
        case CL_LDSPEC: { const Z c = a == 0 ? zcinf() : zindet(); CLOSYR_FORJ { re[j] = c.re; im[j] = c.im; } slow = 1u; break; }
        case CL_ADDX:
            CLOSYR_FORJ { if (im[j] == 0.0) re[j] = re[j] + x[j]; else CLOSYR_SETZ(j, zadd(CLOSYR_Z(j), zv(x[j], 0.0))) }
            break;
        case CL_ADDC: {
            const double c = sconst[b];
            CLOSYR_FORJ { if (im[j] == 0.0) re[j] = re[j] + c; else CLOSYR_SETZ(j, zadd(CLOSYR_Z(j), zv(c, 0.0))) }
            break;
        }
        case CL_MULX:
            CLOSYR_FORJ { if (im[j] == 0.0) re[j] = re[j] * x[j]; else CLOSYR_SETZ(j, zmul(CLOSYR_Z(j), zv(x[j], 0.0))) }
            break;
        case CL_MULC: {
            const double c = sconst[b];
            CLOSYR_FORJ { if (im[j] == 0.0) re[j] = re[j] * c; else CLOSYR_SETZ(j, zmul(CLOSYR_Z(j), zv(c, 0.0))) }
            break;
        }
        case CL_NEG: CLOSYR_UNARY(true, -a_, zneg) break;
        case CL_RECIP: CLOSYR_UNARY(a_ != 0.0, k_recip(a_), zrecip) break;
        case CL_SQR: CLOSYR_FORJ { if (im[j] == 0.0) re[j] = re[j] * re[j]; else CLOSYR_SETZ(j, zmul(CLOSYR_Z(j), CLOSYR_Z(j))) } break;
        case CL_SQRT: CLOSYR_UNARY(a_ >= 0.0, k_sqrt(a_), zsqrt) break;
        case CL_RSQRT:
            CLOSYR_FORJ {
                const double a_ = re[j];
                if (im[j] == 0.0 && a_ > 0.0) re[j] = k_rsqrt(a_);
                else CLOSYR_SETZ(j, zrecip(zsqrt(CLOSYR_Z(j))))
            }
            break;
        case CL_POWI: {
            const int n = (int)(short)b;
            CLOSYR_FORJ {
                if (im[j] == 0.0) {
                    const double r = k_powi(re[j], n < 0 ? -n : n);
                    if (n > 0) re[j] = r;
                    else if (r != 0.0) re[j] = k_recip(r);
                    else CLOSYR_SETZ(j, zcinf())
                } else CLOSYR_SETZ(j, zpowi(CLOSYR_Z(j), n))
            }
            break;
        }
        case CL_POWC: {
            const double y = sconst[b];
            const bool plain = !nonfinite(y) && y != 0.0;
            CLOSYR_FORJ {
                const double a_ = re[j];
                if (plain && im[j] == 0.0 && a_ > 0.0 && !is_inf(a_)) re[j] = k_pow(a_, y);
                else CLOSYR_SETZ(j, zpow(CLOSYR_Z(j), zmk(y, 0.0)))
            }
            break;
        }
        case CL_EXP: CLOSYR_UNARY(true, k_exp(a_), zexp) break;
        case CL_LOG: CLOSYR_UNARY(a_ > 0.0, k_log(a_), zlog) break;
        case CL_SIN: CLOSYR_UNARY(true, k_sin(a_), zsin) break;
        case CL_COS: CLOSYR_UNARY(true, k_cos(a_), zcos) break;
        case CL_TAN: CLOSYR_UNARY(true, k_tan(a_), ztan) break;
        case CL_ASIN: CLOSYR_UNARY(fabs(a_) <= 1.0, k_asin(a_), zasin) break;
        case CL_ACOS: CLOSYR_UNARY(fabs(a_) <= 1.0, k_acos(a_), zacos) break;
        case CL_ABS: CLOSYR_UNARY(true, fabs(a_), zabs) break;
        case CL_ST0 + 0: CLOSYR_ST(s0r, s0i) break;
        case CL_ST0 + 1: CLOSYR_ST(s1r, s1i) break;
        case CL_ST0 + 2: CLOSYR_ST(s2r, s2i) break;
        case CL_ST0 + 3: CLOSYR_ST(s3r, s3i) break;
        case CL_ADDS0 + 0: CLOSYR_BIN(s0r, s0i, +, zadd) break;
        case CL_ADDS0 + 1: CLOSYR_BIN(s1r, s1i, +, zadd) break;
        case CL_ADDS0 + 2: CLOSYR_BIN(s2r, s2i, +, zadd) break;
        case CL_ADDS0 + 3: CLOSYR_BIN(s3r, s3i, +, zadd) break;
        case CL_MULS0 + 0: CLOSYR_BIN(s0r, s0i, *, zmul) break;
        case CL_MULS0 + 1: CLOSYR_BIN(s1r, s1i, *, zmul) break;
        case CL_MULS0 + 2: CLOSYR_BIN(s2r, s2i, *, zmul) break;
        case CL_MULS0 + 3: CLOSYR_BIN(s3r, s3i, *, zmul) break;
        default: {
            // generic slot forms (deep slots) and the general power: scalar algebra throughout
            CLOSYR_FORJ {
                Z sl = zv(0.0, 0.0);
                if (op != CL_ST) {
                    if (a >= CLOSYR_REG_SLOTS) sl = ls[a - CLOSYR_REG_SLOTS][j];
                    else sl = a == 0 ? zv(s0r[j], s0i[j]) : a == 1 ? zv(s1r[j], s1i[j]) : a == 2 ? zv(s2r[j], s2i[j]) : zv(s3r[j], s3i[j]);
                }
                if (op == CL_ST) {
                    if (a >= CLOSYR_REG_SLOTS) ls[a - CLOSYR_REG_SLOTS][j] = CLOSYR_Z(j);
                    else if (a == 0) { s0r[j] = re[j]; s0i[j] = im[j]; }
                    else if (a == 1) { s1r[j] = re[j]; s1i[j] = im[j]; }
                    else if (a == 2) { s2r[j] = re[j]; s2i[j] = im[j]; }
                    else { s3r[j] = re[j]; s3i[j] = im[j]; }
                } else if (op == CL_ADDS) CLOSYR_SETZ(j, zadd(sl, CLOSYR_Z(j)))
                else if (op == CL_MULS) CLOSYR_SETZ(j, zmul(sl, CLOSYR_Z(j)))
                else if (op == CL_POWS) CLOSYR_SETZ(j, zpow(sl, CLOSYR_Z(j)))
                else CLOSYR_SETZ(j, zindet())
            }
            break;
        }
        }
    }
#undef CLOSYR_ST
#undef CLOSYR_BIN
}

// ------------------------------------------------------------------------------------------------
// fast path: P points per lane, IEEE real arithmetic, register-resident accumulator and slots 0..3.
//
// Invariant kept per point (v = the IEEE value here, s = the value of the full algebra):
//   v finite  <=>  s is a finite real and v == s;   v = +-Inf  =>  s is ComplexInfinity or a directed infinity
//   (IEEE cannot tell which);   v NaN, no event  =>  s is Indeterminate (absorbing: the point is not real).
// The returned mask has bit j set when point j must be re-evaluated by full_eval because the invariant
// cannot be kept:  (1) a function left its real domain (NaN out of a non-NaN input: the full algebra
// continues over the complex numbers or with ComplexInfinity);  (2) an infinity reaches an instruction that
// can turn it into a finite number (1/Inf = 0, E^-Inf = 0, Inf^-y = 0) -- whether that is right depends on
// which infinity it was;  (3) literals that are not real, and general x^y.
// ------------------------------------------------------------------------------------------------
#define CLOSYR_DOMAIN_OP(expr)                                                 \
    CLOSYR_FORJ {                                                              \
        const double in_ = acc[j];                                             \
        const double r_ = (expr);                                              \
        if (r_ != r_ && in_ == in_) ev |= 1u << j;                             \
        acc[j] = r_;                                                           \
    }

template <int P>
__device__ __forceinline__ uint32_t run_real(const uint32_t *__restrict__ scode, const double *__restrict__ sconst,
                                             const double (&x)[P], double (&acc)[P])
{
    double s0[P], s1[P], s2[P], s3[P];
    double ls[CLOSYR_MAX_SLOTS - CLOSYR_REG_SLOTS][P];    // deep spill slots: local memory, rarely touched
    uint32_t ev = 0;
    int pc = 0;
    uint32_t w = scode[0];
    CLOSYR_FORJ { acc[j] = 0.0; s0[j] = 0.0; s1[j] = 0.0; s2[j] = 0.0; s3[j] = 0.0; }
    for (;;) {
        const uint32_t op = w & 0xffu, a = (w >> 8) & 0xffu;
        const int b = (int)(w >> 16);
        w = scode[++pc];                                     // prefetch (the buffer has one word of slack)
        switch (op) {
        case CL_END: return ev;
        case CL_LDX: CLOSYR_FORJ acc[j] = x[j]; break;
        case CL_LDC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = c; break; }
        case CL_LDZ:
        case CL_LDSPEC: ev = (1u << P) - 1u; break;
        case CL_ADDX: CLOSYR_FORJ acc[j] = acc[j] + x[j]; break;
        case CL_ADDC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = acc[j] + c; break; }
        case CL_MULX: CLOSYR_FORJ acc[j] = acc[j] * x[j]; break;
        case CL_MULC: { const double c = sconst[b]; CLOSYR_FORJ acc[j] = acc[j] * c; break; }
        case CL_NEG: CLOSYR_FORJ acc[j] = -acc[j]; break;
        case CL_RECIP: CLOSYR_FORJ { if (is_inf(acc[j])) ev |= 1u << j; acc[j] = k_recip(acc[j]); } break;
        case CL_SQR: CLOSYR_FORJ acc[j] = acc[j] * acc[j]; break;
        case CL_SQRT: CLOSYR_DOMAIN_OP(k_sqrt(in_)) break;
        case CL_RSQRT: CLOSYR_DOMAIN_OP(is_inf(in_) ? __longlong_as_double(0x7ff8000000000000ll) : k_rsqrt(in_)) break;
        case CL_POWI: {
            const int n = (int)(short)b;
            CLOSYR_FORJ { if (n < 0 && is_inf(acc[j])) ev |= 1u << j; acc[j] = k_powi(acc[j], n); }
            break;
        }
        case CL_POWC: {
            const double y = sconst[b];
            if (nonfinite(y) || y == 0.0) { ev = (1u << P) - 1u; break; }
            CLOSYR_DOMAIN_OP(is_inf(in_) ? __longlong_as_double(0x7ff8000000000000ll) : k_pow(in_, y))
            break;
        }
        case CL_EXP:
            CLOSYR_FORJ { if (is_inf(acc[j])) ev |= 1u << j; acc[j] = k_exp(acc[j]); }
            break;
        case CL_LOG: CLOSYR_DOMAIN_OP(k_log(in_)) break;
        case CL_SIN: CLOSYR_FORJ acc[j] = k_sin(acc[j]); break;
        case CL_COS: CLOSYR_FORJ acc[j] = k_cos(acc[j]); break;
        case CL_TAN: CLOSYR_FORJ acc[j] = k_tan(acc[j]); break;
        case CL_ASIN: CLOSYR_DOMAIN_OP(k_asin(in_)) break;
        case CL_ACOS: CLOSYR_DOMAIN_OP(k_acos(in_)) break;
        case CL_ABS: CLOSYR_FORJ acc[j] = fabs(acc[j]); break;
        case CL_ST0 + 0: CLOSYR_FORJ s0[j] = acc[j]; break;
        case CL_ST0 + 1: CLOSYR_FORJ s1[j] = acc[j]; break;
        case CL_ST0 + 2: CLOSYR_FORJ s2[j] = acc[j]; break;
        case CL_ST0 + 3: CLOSYR_FORJ s3[j] = acc[j]; break;
        case CL_ADDS0 + 0: CLOSYR_FORJ acc[j] = s0[j] + acc[j]; break;
        case CL_ADDS0 + 1: CLOSYR_FORJ acc[j] = s1[j] + acc[j]; break;
        case CL_ADDS0 + 2: CLOSYR_FORJ acc[j] = s2[j] + acc[j]; break;
        case CL_ADDS0 + 3: CLOSYR_FORJ acc[j] = s3[j] + acc[j]; break;
        case CL_MULS0 + 0: CLOSYR_FORJ acc[j] = s0[j] * acc[j]; break;
        case CL_MULS0 + 1: CLOSYR_FORJ acc[j] = s1[j] * acc[j]; break;
        case CL_MULS0 + 2: CLOSYR_FORJ acc[j] = s2[j] * acc[j]; break;
        case CL_MULS0 + 3: CLOSYR_FORJ acc[j] = s3[j] * acc[j]; break;
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
        case CL_POWS:
        default:
            ev = (1u << P) - 1u;       // general x^y (dormant crossover flavours only): full semantics decide
            break;
        }
    }
}

}  // namespace dev
}  // namespace closyr
