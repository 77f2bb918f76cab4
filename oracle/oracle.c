/*
 * oracle.c -- CPU restatement of closyr's scoring hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (closyr_b200/) never links, imports or calls it.
 *
 * What is restated (all paths relative to /root/reference):
 *   score-fn                                   src/closyr/ops.clj:119-136
 *   compute-score-from-actuals-and-expecteds   src/closyr/ops.clj:98-116
 *   compute-residual / not-finite?             src/closyr/ops.clj:72-90
 *   length-deduction, sum, min-score, max-resid src/closyr/ops.clj:24-36, 61-63, 93-95
 *   eval-vec-pheno and its decoders            src/closyr/ops/eval.clj:59-112
 *   Function({x},expr)[List(xs)] evaluation    src/closyr/ops/eval.clj:40-56, common.clj:103-110
 *
 * PARITY STATUS.  The score formula, residual rules, sentinels, constant broadcast, complex=>+Inf
 * and the eval vectors are PINNED against the reference's own tests (tests/test_oracle_golden.py
 * lists every vector with its file:line).  The numeric evaluation itself lives in a third-party
 * dependency that is not vendored under /root/reference and cannot run here (no JVM):
 *   org.matheclipse/matheclipse-core 3.1.0-20240219.223701-64   (doc/depstree.txt:68)
 *   org.hipparchus/hipparchus-core  3.1-20240212.230522-23      (doc/depstree.txt:88)
 * Its published algorithms are restated below: Symja threads Listable heads over the point list and
 * evaluates each element over the complex numbers with Mathematica-style special values; complex
 * elementary functions follow the textbook formulas used by Hipparchus' Complex class (sqrt, log,
 * exp, sin, cos, tan, asin = -i log(sqrt(1-z^2)+iz), acos = -i log(z+i sqrt(1-z^2)), pow = exp(w log z)).
 * Every rule not exercised by a reference test is "PARITY UNPINNED" and is listed in DESIGN.md
 * (section "Semantics"); they are assumptions, implemented identically by the CUDA path.
 *
 * Input format: Symja FullForm strings, e.g.  Plus(Rational(-1,2),Power(x,2))   (IExpr.fullFormString()).
 * This is the exact internal tree, so leafCount (ops.clj:127) is computed without any canonicalisation.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------ sentinels (ops.clj:24-36) */
#define MIN_SCORE (-100000000.0)
#define MAX_RESID (1000000.0)

/* ------------------------------------------------------------------ libm jitter (conditioning probe)
 * Java's Math.*, glibc and CUDA's libdevice are all "within an ulp or two", not bit-identical.  To tell a
 * real mismatch from an expression whose value is not determined by such libraries (e.g. Cos(E^(x^2)) at
 * large x), tests evaluate the oracle again with every transcendental result nudged by one ulp up / down and
 * only hold the 1e-9 bar where the oracle agrees with itself.  0 = off (default). */
static int g_jitter = 0;                      /* 0 off; otherwise the seed of a random +-1 ulp pattern */
static int g_jitter_ulps = 1;                 /* amplitude: each result moves by 1 .. g_jitter_ulps ulp, either way */
static _Thread_local uint64_t t_jcnt = 0;     /* position in the pattern; reset per expression */
void oracle_set_jitter(int seed) { g_jitter = seed; }
/* CUDA's own libm documents up to 2 ulp for sin cos tan asin acos pow (1 ulp for exp log); the explanation probes
 * (oracle.py *_discrepancy_explained) use that envelope, the stability masks keep +-1 ulp */
void oracle_set_jitter_ulps(int ulps) { g_jitter_ulps = ulps < 1 ? 1 : ulps; }
static double J(double r)
{
    if (g_jitter == 0 || !isfinite(r)) return r;
    uint64_t z = (t_jcnt += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double dir = (z & 1) ? INFINITY : -INFINITY;
    int steps = 1 + (int)((z >> 1) % (uint64_t)g_jitter_ulps);
    while (steps-- > 0) r = nextafter(r, dir);
    return r;
}
#define exp(x) J(exp(x))
#define log(x) J(log(x))
#define sin(x) J(sin(x))
#define cos(x) J(cos(x))
#define tan(x) J(tan(x))
#define sinh(x) J(sinh(x))
#define cosh(x) J(cosh(x))
#define asin(x) J(asin(x))
#define acos(x) J(acos(x))
#define atan2(y, x) J(atan2(y, x))
#define hypot(x, y) J(hypot(x, y))
#define pow(x, y) J(pow(x, y))

/* ------------------------------------------------------------------ values */
typedef struct { double re, im; } val_t;
enum { T_REAL = 0, T_CPLX = 1, T_CINF = 2, T_INDET = 3 };

static int tag(val_t v)
{
    if (isnan(v.re) || isnan(v.im)) return T_INDET;
    if (v.im == 0.0) return T_REAL;               /* real line, including directed +-Infinity */
    if (isinf(v.re) || isinf(v.im)) return T_CINF;
    return T_CPLX;
}
static val_t V(double re, double im) { val_t v = { re, im }; return v; }
static val_t INDET(void) { return V(NAN, 0.0); }
static val_t CINF(void) { return V(INFINITY, INFINITY); }
/* normalise: one encoding per class */
/* A complex result whose imaginary (real) part is rounding noise next to the other part is snapped onto
 * the real (imaginary) axis: noise must not decide whether an element "is a real number" (eval.clj:74),
 * nor which side of a branch cut it sits on.  PARITY UNPINNED (Symja's own chop rule is not observable here). */
#define CHOP_REL 1e-10
static val_t mk(double re, double im)
{
    if (im != 0.0 && fabs(im) <= CHOP_REL * fabs(re)) im = 0.0;
    else if (re != 0.0 && fabs(re) <= CHOP_REL * fabs(im)) re = 0.0;
    val_t v = { re, im };
    switch (tag(v)) {
    case T_INDET: return INDET();
    case T_CINF: return CINF();
    case T_REAL: return V(re, 0.0);
    default: return v;
    }
}

static val_t v_add(val_t a, val_t b)
{
    int ta = tag(a), tb = tag(b);
    if (ta == T_INDET || tb == T_INDET) return INDET();
    if (ta == T_CINF || tb == T_CINF) {
        val_t o = (ta == T_CINF) ? b : a;
        int to = (ta == T_CINF) ? tb : ta;
        if (to == T_CINF) return INDET();
        if (to == T_REAL && isinf(o.re)) return INDET();
        return CINF();
    }
    return mk(a.re + b.re, a.im + b.im);
}

static val_t v_mul(val_t a, val_t b)
{
    int ta = tag(a), tb = tag(b);
    if (ta == T_INDET || tb == T_INDET) return INDET();
    if (ta == T_CINF || tb == T_CINF) {
        val_t o = (ta == T_CINF) ? b : a;
        int to = (ta == T_CINF) ? tb : ta;
        if (to == T_REAL && o.re == 0.0) return INDET();
        return CINF();
    }
    if (ta == T_REAL && tb == T_REAL) return mk(a.re * b.re, 0.0);
    if (ta == T_REAL) return mk(a.re * b.re, a.re * b.im);
    if (tb == T_REAL) return mk(a.re * b.re, a.im * b.re);
    return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}

static val_t v_neg(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return CINF();
    if (t == T_REAL) return V(-a.re, 0.0);
    return V(-a.re, -a.im);
}

static val_t v_recip(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return V(0.0, 0.0);
    if (t == T_REAL) {
        if (a.re == 0.0) return CINF();           /* 1/0.0 -> ComplexInfinity (quiet mode) */
        return V(1.0 / a.re, 0.0);
    }
    /* Smith's algorithm, as Hipparchus Complex.reciprocal() */
    double c = a.re, d = a.im;
    if (fabs(c) < fabs(d)) {
        double q = c / d;
        double s = 1.0 / (c * q + d);
        return mk(s * q, -s);
    } else {
        double q = d / c;
        double s = 1.0 / (d * q + c);
        return mk(s, -s * q);
    }
}

static val_t v_sqrt(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return CINF();
    if (t == T_REAL) {
        if (a.re >= 0.0) return V(sqrt(a.re), 0.0);       /* includes +Inf and -0.0 */
        if (isinf(a.re)) return CINF();
        return mk(0.0, sqrt(-a.re));
    }
    double x = a.re, y = a.im;
    double tt = sqrt((fabs(x) + hypot(x, y)) / 2.0);
    if (x >= 0.0) return mk(tt, y / (2.0 * tt));
    return mk(fabs(y) / (2.0 * tt), copysign(tt, y));
}

static val_t v_exp(val_t a)
{
    int t = tag(a);
    if (t == T_INDET || t == T_CINF) return INDET();
    if (t == T_REAL) return mk(exp(a.re), 0.0);
    double e = exp(a.re);
    return mk(e * cos(a.im), e * sin(a.im));
}

static val_t v_log(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return V(INFINITY, 0.0);
    if (t == T_REAL) {
        if (a.re > 0.0) return V(log(a.re), 0.0);
        if (a.re == 0.0) return V(-INFINITY, 0.0);
        if (isinf(a.re)) return CINF();
        return mk(log(-a.re), M_PI);
    }
    return mk(log(hypot(a.re, a.im)), atan2(a.im, a.re));
}

static val_t v_sin(val_t a)
{
    int t = tag(a);
    if (t == T_INDET || t == T_CINF) return INDET();
    if (t == T_REAL) return mk(sin(a.re), 0.0);
    return mk(sin(a.re) * cosh(a.im), cos(a.re) * sinh(a.im));
}

static val_t v_cos(val_t a)
{
    int t = tag(a);
    if (t == T_INDET || t == T_CINF) return INDET();
    if (t == T_REAL) return mk(cos(a.re), 0.0);
    return mk(cos(a.re) * cosh(a.im), -(sin(a.re) * sinh(a.im)));
}

static val_t v_tan(val_t a)
{
    int t = tag(a);
    if (t == T_INDET || t == T_CINF) return INDET();
    if (t == T_REAL) return mk(tan(a.re), 0.0);
    if (a.im > 20.0) return V(0.0, 1.0);
    if (a.im < -20.0) return V(0.0, -1.0);
    double a2 = 2.0 * a.re, b2 = 2.0 * a.im;
    double d = cos(a2) + cosh(b2);
    return mk(sin(a2) / d, sinh(b2) / d);
}

static val_t v_mul_i(val_t a) { return mk(-a.im, a.re); }       /*  i * a, a finite */
static val_t v_mul_minus_i(val_t a) { return mk(a.im, -a.re); } /* -i * a, a finite */

static val_t v_sqrt1z(val_t z) { return v_sqrt(v_add(V(1.0, 0.0), v_neg(v_mul(z, z)))); }

static val_t v_asin(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return CINF();
    if (t == T_REAL) {
        if (fabs(a.re) <= 1.0) return V(asin(a.re), 0.0);
        if (isinf(a.re)) return CINF();
        /* real |x| > 1: the same principal value, written without the cancellation of the general formula */
        double ax = fabs(a.re), l = log(ax + sqrt((ax - 1.0) * (ax + 1.0)));
        return a.re > 0.0 ? mk(M_PI / 2.0, -l) : mk(-(M_PI / 2.0), l);
    }
    val_t w = v_log(v_add(v_sqrt1z(a), v_mul_i(a)));
    int tw = tag(w);
    if (tw == T_INDET) return INDET();
    if (tw == T_CINF || (tw == T_REAL && isinf(w.re))) return CINF();
    return v_mul_minus_i(w);
}

static val_t v_acos(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return CINF();
    if (t == T_REAL) {
        if (fabs(a.re) <= 1.0) return V(acos(a.re), 0.0);
        if (isinf(a.re)) return CINF();
        double ax = fabs(a.re), l = log(ax + sqrt((ax - 1.0) * (ax + 1.0)));
        return a.re > 0.0 ? mk(0.0, l) : mk(M_PI, -l);
    }
    val_t s = v_sqrt1z(a);
    int ts = tag(s);
    if (ts == T_INDET) return INDET();
    if (ts == T_CINF || (ts == T_REAL && isinf(s.re))) return CINF();
    val_t w = v_log(v_add(a, v_mul_i(s)));
    int tw = tag(w);
    if (tw == T_INDET) return INDET();
    if (tw == T_CINF || (tw == T_REAL && isinf(w.re))) return CINF();
    return v_mul_minus_i(w);
}

static val_t v_abs(val_t a)
{
    int t = tag(a);
    if (t == T_INDET) return INDET();
    if (t == T_CINF) return V(INFINITY, 0.0);
    if (t == T_REAL) return V(fabs(a.re), 0.0);
    return mk(hypot(a.re, a.im), 0.0);
}

/* integer power by square-and-multiply, MSB first; |n| <= 32 (frozen spec) */
static val_t v_powi(val_t a, int n)
{
    unsigned m = (unsigned)(n < 0 ? -n : n);
    val_t r = a;
    int top = 31;
    while (top > 0 && !((m >> top) & 1u)) top--;
    for (int b = top - 1; b >= 0; b--) {
        r = v_mul(r, r);
        if ((m >> b) & 1u) r = v_mul(r, a);
    }
    if (n < 0) r = v_recip(r);
    return r;
}

static int is_integer_value(double y) { return isfinite(y) && floor(y) == y; }

/* general power z^w */
static val_t v_pow(val_t z, val_t w)
{
    int tz = tag(z), tw = tag(w);
    if (tz == T_INDET || tw == T_INDET) return INDET();
    if (tw == T_CINF) return INDET();
    double yr = w.re; /* real part of the exponent decides the infinite / zero cases */
    if (tz == T_CINF) {
        if (tw == T_REAL && isinf(yr)) return INDET();
        if (yr > 0.0) return CINF();
        if (yr < 0.0) return V(0.0, 0.0);
        return INDET();
    }
    if (tz == T_REAL && tw == T_REAL) {
        double x = z.re, y = yr;
        if (y == 0.0) return (isfinite(x) && x != 0.0) ? V(1.0, 0.0) : INDET();
        if (isinf(y)) {
            double ax = fabs(x);
            if (ax == 1.0) return INDET();
            int grows = (ax > 1.0) == (y > 0.0);
            if (!grows) return V(0.0, 0.0);
            return (x > 0.0) ? V(INFINITY, 0.0) : CINF();
        }
        if (isinf(x)) {
            if (y < 0.0) return V(0.0, 0.0);
            if (x > 0.0) return V(INFINITY, 0.0);
            return is_integer_value(y) ? mk(pow(x, y), 0.0) : CINF();
        }
        if (x == 0.0) return (y > 0.0) ? V(0.0, 0.0) : CINF();
        if (x > 0.0 || is_integer_value(y)) return mk(pow(x, y), 0.0);
        double m = pow(-x, y);
        double th = y * M_PI;
        return mk(m * cos(th), m * sin(th));
    }
    /* at least one operand is genuinely complex (or a directed infinity met a complex number) */
    if (tz == T_REAL && z.re == 0.0) {
        if (yr > 0.0) return V(0.0, 0.0);
        if (yr < 0.0) return CINF();
        return INDET();
    }
    if (tz == T_REAL && isinf(z.re)) {
        if (yr > 0.0) return CINF();
        if (yr < 0.0) return V(0.0, 0.0);
        return INDET();
    }
    if (tw == T_REAL && isinf(yr)) {
        double az = hypot(z.re, z.im);
        if (az == 1.0) return INDET();
        return ((az > 1.0) == (yr > 0.0)) ? CINF() : V(0.0, 0.0);
    }
    return v_exp(v_mul(w, v_log(z)));
}

/* ------------------------------------------------------------------ expression tree */
enum {
    K_NUM,      /* Integer / Rational / Real literal (real value in v) */
    K_X, K_E, K_PI, K_I, K_COMPLEX, K_DINF, K_CINF, K_INDET,
    K_PLUS, K_TIMES, K_POWER,
    K_SIN, K_COS, K_TAN, K_COT, K_SEC, K_CSC, K_ASIN, K_ACOS, K_LOG, K_ABS, K_EXP, K_SQRT,
    K_UNKNOWN
};

typedef struct node {
    int kind;
    double v, v2;        /* K_NUM: value; K_COMPLEX: re, im; K_DINF: +-1 */
    int leafs;           /* Symja leafCount of this subtree (ops.clj:127) */
    int lit;             /* K_NUM: 1 Integer, 2 Rational, 3 Real */
    long long ival;      /* Integer literal when |value| < 2^62 */
    long long rnum, rden;/* Rational parts when they fit */
    int has_x;
    int nargs;
    struct node **args;
} node;

typedef struct { const char *s; int err; } parser;

static node *new_node(int kind)
{
    node *n = (node *)calloc(1, sizeof(node));
    n->kind = kind;
    n->leafs = 1;
    return n;
}

static void free_node(node *n)
{
    if (!n) return;
    for (int i = 0; i < n->nargs; i++) free_node(n->args[i]);
    free(n->args);
    free(n);
}

static void skip_ws(parser *p) { while (*p->s && isspace((unsigned char)*p->s)) p->s++; }

static node *parse_expr(parser *p);

static const struct { const char *name; int kind; } HEADS[] = {
    { "Plus", K_PLUS }, { "Times", K_TIMES }, { "Power", K_POWER }, { "Sin", K_SIN }, { "Cos", K_COS },
    { "Tan", K_TAN }, { "Cot", K_COT }, { "Sec", K_SEC }, { "Csc", K_CSC }, { "ArcSin", K_ASIN },
    { "ArcCos", K_ACOS }, { "Log", K_LOG }, { "Abs", K_ABS }, { "Exp", K_EXP }, { "Sqrt", K_SQRT },
    { NULL, 0 }
};

static node *parse_number(parser *p)
{
    const char *b = p->s;
    const char *q = b;
    int is_real = 0;
    if (*q == '-' || *q == '+') q++;
    while (isdigit((unsigned char)*q)) q++;
    if (*q == '.') { is_real = 1; q++; while (isdigit((unsigned char)*q)) q++; }
    if ((*q == 'E' || *q == 'e') && is_real) {
        const char *r = q + 1;
        if (*r == '-' || *r == '+') r++;
        if (isdigit((unsigned char)*r)) { while (isdigit((unsigned char)*r)) r++; q = r; }
    }
    if (q == b || (q == b + 1 && (*b == '-' || *b == '+'))) { p->err = 1; return NULL; }
    char buf[512];
    size_t len = (size_t)(q - b);
    if (len >= sizeof buf) { p->err = 1; return NULL; }
    memcpy(buf, b, len);
    buf[len] = 0;
    node *n = new_node(K_NUM);
    n->v = strtod(buf, NULL);          /* correctly rounded, like BigInteger.doubleValue / Double.parseDouble */
    n->lit = is_real ? 3 : 1;
    if (!is_real && len < 19) n->ival = strtoll(buf, NULL, 10);
    else if (!is_real) n->ival = (n->v < 0) ? INT64_MIN : INT64_MAX; /* too big to matter */
    p->s = q;
    return n;
}

static node *parse_expr(parser *p)
{
    skip_ws(p);
    if (p->err) return NULL;
    char c = *p->s;
    if (isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') return parse_number(p);
    if (!isalpha((unsigned char)c) && c != '$') { p->err = 1; return NULL; }
    char id[64];
    int k = 0;
    while ((isalnum((unsigned char)*p->s) || *p->s == '$' || *p->s == '_') && k < 63) id[k++] = *p->s++;
    id[k] = 0;
    skip_ws(p);
    if (*p->s != '(') {
        node *n;
        if (!strcmp(id, "x")) { n = new_node(K_X); n->has_x = 1; }
        else if (!strcmp(id, "E")) n = new_node(K_E);
        else if (!strcmp(id, "Pi")) n = new_node(K_PI);
        else if (!strcmp(id, "I")) { n = new_node(K_I); n->leafs = 3; } /* Complex(0,1): leafCount 3 */
        else if (!strcmp(id, "Infinity")) { n = new_node(K_DINF); n->v = 1.0; }
        else if (!strcmp(id, "ComplexInfinity")) n = new_node(K_CINF);
        else if (!strcmp(id, "Indeterminate")) n = new_node(K_INDET);
        else n = new_node(K_UNKNOWN);
        return n;
    }
    p->s++; /* '(' */
    node **args = NULL;
    int na = 0, cap = 0;
    skip_ws(p);
    if (*p->s == ')') p->s++;
    else {
        for (;;) {
            node *a = parse_expr(p);
            if (p->err || !a) { for (int i = 0; i < na; i++) free_node(args[i]); free(args); p->err = 1; return NULL; }
            if (na == cap) { cap = cap ? cap * 2 : 4; args = (node **)realloc(args, sizeof(node *) * (size_t)cap); }
            args[na++] = a;
            skip_ws(p);
            if (*p->s == ',') { p->s++; continue; }
            if (*p->s == ')') { p->s++; break; }
            for (int i = 0; i < na; i++) free_node(args[i]);
            free(args);
            p->err = 1;
            return NULL;
        }
    }
    node *n = NULL;
    if (!strcmp(id, "Rational") && na == 2 && args[0]->kind == K_NUM && args[1]->kind == K_NUM) {
        n = new_node(K_NUM);
        n->lit = 2;
        n->leafs = 3;
        n->rnum = args[0]->ival;
        n->rden = args[1]->ival;
        n->v = args[0]->v / args[1]->v;   /* double(num)/double(den) (frozen spec) */
        free_node(args[0]); free_node(args[1]); free(args);
        return n;
    }
    if (!strcmp(id, "Complex") && na == 2 && args[0]->kind == K_NUM && args[1]->kind == K_NUM) {
        n = new_node(K_COMPLEX);
        n->leafs = 3;
        n->v = args[0]->v;
        n->v2 = args[1]->v;
        free_node(args[0]); free_node(args[1]); free(args);
        return n;
    }
    if (!strcmp(id, "DirectedInfinity")) {
        if (na == 0) { n = new_node(K_CINF); free(args); return n; }
        if (na == 1 && args[0]->kind == K_NUM && (args[0]->v == 1.0 || args[0]->v == -1.0)) {
            n = new_node(K_DINF);
            n->v = args[0]->v;
            n->leafs = 2;
            free_node(args[0]); free(args);
            return n;
        }
    }
    int kind = K_UNKNOWN;
    for (int i = 0; HEADS[i].name; i++) if (!strcmp(id, HEADS[i].name)) { kind = HEADS[i].kind; break; }
    if (kind != K_UNKNOWN && kind != K_PLUS && kind != K_TIMES && kind != K_POWER && na != 1) kind = K_UNKNOWN;
    if (kind == K_POWER && na != 2) kind = K_UNKNOWN;
    if ((kind == K_PLUS || kind == K_TIMES) && na < 1) kind = K_UNKNOWN;
    n = new_node(kind);
    n->args = args;
    n->nargs = na;
    n->leafs = 1;                       /* the head counts 1 (A12; pinned by -3.0000147 => 7) */
    for (int i = 0; i < na; i++) { n->leafs += args[i]->leafs; n->has_x |= args[i]->has_x; }
    return n;
}

static node *parse(const char *s)
{
    parser p = { s, 0 };
    node *n = parse_expr(&p);
    skip_ws(&p);
    if (p.err || *p.s) { free_node(n); return NULL; }
    return n;
}

/* ------------------------------------------------------------------ evaluation of one element */
static val_t eval(const node *n, double x)
{
    switch (n->kind) {
    case K_NUM: return mk(n->v, 0.0);
    case K_X: return mk(x, 0.0);
    case K_E: return V(M_E, 0.0);
    case K_PI: return V(M_PI, 0.0);
    case K_I: return V(0.0, 1.0);
    case K_COMPLEX: return mk(n->v, n->v2);
    case K_DINF: return V(n->v > 0 ? INFINITY : -INFINITY, 0.0);
    case K_CINF: return CINF();
    case K_INDET: return INDET();
    case K_PLUS: {
        val_t a = eval(n->args[0], x);
        for (int i = 1; i < n->nargs; i++) a = v_add(a, eval(n->args[i], x));   /* left fold (A11) */
        return a;
    }
    case K_TIMES: {
        val_t a = eval(n->args[0], x);
        for (int i = 1; i < n->nargs; i++) a = v_mul(a, eval(n->args[i], x));
        return a;
    }
    case K_POWER: {
        const node *b = n->args[0], *e = n->args[1];
        if (e->kind == K_NUM && e->lit == 1 && e->ival != 0 && e->ival >= -32 && e->ival <= 32)
            return v_powi(eval(b, x), (int)e->ival);
        if (e->kind == K_NUM && e->lit == 2 && e->rden == 2 && e->rnum == 1) return v_sqrt(eval(b, x));
        if (e->kind == K_NUM && e->lit == 2 && e->rden == 2 && e->rnum == -1) return v_recip(v_sqrt(eval(b, x)));
        if (b->kind == K_E) return v_exp(eval(e, x));
        return v_pow(eval(b, x), eval(e, x));
    }
    case K_EXP: return v_exp(eval(n->args[0], x));
    case K_SQRT: return v_sqrt(eval(n->args[0], x));
    case K_SIN: return v_sin(eval(n->args[0], x));
    case K_COS: return v_cos(eval(n->args[0], x));
    case K_TAN: return v_tan(eval(n->args[0], x));
    case K_COT: return v_recip(v_tan(eval(n->args[0], x)));
    case K_SEC: return v_recip(v_cos(eval(n->args[0], x)));
    case K_CSC: return v_recip(v_sin(eval(n->args[0], x)));
    case K_ASIN: return v_asin(eval(n->args[0], x));
    case K_ACOS: return v_acos(eval(n->args[0], x));
    case K_LOG: return v_log(eval(n->args[0], x));
    case K_ABS: return v_abs(eval(n->args[0], x));
    default: return INDET();            /* unknown head stays symbolic in Symja => not a real number */
    }
}

/* eval.clj:70-80: element real ? its double : +Inf */
static double decode_element(val_t v) { return tag(v) == T_REAL ? v.re : INFINITY; }

/*
 * eval-vec-pheno (eval.clj:99-112).  Returns 0 and fills out[0..n) on success, 1 when the reference
 * returns nil (whole result Indeterminate, eval.clj:59-62), 2 when the reference throws
 * (constant expression whose numeric value is not a real number, eval.clj:83-96), 3 on parse error.
 */
static int eval_vec(const node *e, const double *xs, int64_t n, double *out)
{
    if (e->has_x) {
        for (int64_t i = 0; i < n; i++) out[i] = decode_element(eval(e, xs[i]));
        return 0;
    }
    if (e->kind == K_INDET) return 1;
    /* A13 quirk (eval.clj:109): a constant AST whose argument count equals n is decoded per argument */
    if (e->nargs > 0 && (int64_t)e->nargs == n) {
        for (int64_t i = 0; i < n; i++) {
            const node *a = e->args[i];
            out[i] = (a->kind == K_NUM) ? a->v : INFINITY;
        }
        return 0;
    }
    double c;
    if (e->kind == K_NUM) c = e->v;
    else {
        val_t v = eval(e, 0.0);
        if (tag(v) != T_REAL) return 2;
        c = v.re;
    }
    for (int64_t i = 0; i < n; i++) out[i] = c;
    return 0;
}

/* ops.clj:72-90 */
static double compute_residual(double expected, double actual)
{
    double res;
    if (isnan(actual)) return MAX_RESID;
    res = expected - actual;
    if (isnan(res)) return MAX_RESID;
    res = fabs(res);
    return res < MAX_RESID ? res : MAX_RESID;
}

/* the CUDA path's fixed reduction shape (DESIGN.md "Reduction order"): 128-point chunks, lane l owns
 * points 4l..4l+3 (left fold), xor-butterfly over 32 lanes, then an index-aligned binary tree over chunks */
static double chunk_sum(const double *r, int64_t cnt)
{
    double lane[32], tmp[32];
    for (int l = 0; l < 32; l++) {
        double s = 0.0;
        for (int j = 0; j < 4; j++) { int64_t i = 4 * l + j; s = s + (i < cnt ? r[i] : 0.0); }
        lane[l] = s;
    }
    for (int st = 16; st >= 1; st >>= 1) {
        for (int l = 0; l < 32; l++) tmp[l] = lane[l] + lane[l ^ st];
        memcpy(lane, tmp, sizeof lane);
    }
    return lane[0];
}
static double tree_sum(const double *r, int64_t n, int64_t c0, int64_t span)
{
    int64_t nchunks = (n + 127) / 128;
    if (c0 >= nchunks) return -1.0; /* absent */
    if (span == 1) { int64_t cnt = n - c0 * 128; return chunk_sum(r + c0 * 128, cnt > 128 ? 128 : cnt); }
    double a = tree_sum(r, n, c0, span / 2), b = tree_sum(r, n, c0 + span / 2, span / 2);
    return b < 0.0 ? a : a + b;
}

/* ops.clj:98-116.  sum_mode 0 = the reference's left fold from 0.0 (ops.clj:61-63); 1 = CUDA tree shape */
static double score_from_f(const double *f, const double *ys, int64_t n, int leafs, int sum_mode, double *scratch)
{
    if (n <= 0) return MIN_SCORE; /* (reduce max []) throws => min-score */
    double s = 0.0, m = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double r = compute_residual(ys[i], f[i]);
        scratch[i] = r;
        s = s + r;
        if (i == 0 || r > m) m = r;
    }
    if (sum_mode == 1) {
        int64_t nchunks = (n + 127) / 128, span = 1;
        while (span < nchunks) span <<= 1;
        s = tree_sum(scratch, n, 0, span);
    }
    double score = -1.0 * ((2.0 * (s / (double)n)) + m);
    double pen = (0.0000001 * (double)leafs) * (double)leafs;
    if (pen > 0.1) pen = 0.1;
    double ded = fabs(score) * pen;
    return score - ded;
}

/* ------------------------------------------------------------------ exported API (ctypes) */

int oracle_leaf_count(const char *expr)
{
    node *e = parse(expr);
    if (!e) return -1;
    int l = e->leafs;
    free_node(e);
    return l;
}

/* eval-vec-pheno; see eval_vec for the status codes */
int oracle_eval(const char *expr, const double *xs, int64_t n, double *out)
{
    t_jcnt = (uint64_t)g_jitter * 0xD1B54A32D192ED03ull + strlen(expr);
    node *e = parse(expr);
    if (!e) return 3;
    int st = eval_vec(e, xs, n, out);
    free_node(e);
    return st;
}

/* raw element values (re, im, tag) for semantics tests */
int oracle_eval_raw(const char *expr, const double *xs, int64_t n, double *re, double *im, int *tags)
{
    node *e = parse(expr);
    if (!e) return 3;
    for (int64_t i = 0; i < n; i++) {
        val_t v = eval(e, xs[i]);
        re[i] = v.re; im[i] = v.im; tags[i] = tag(v);
    }
    free_node(e);
    return 0;
}

/* (sum, max) of the residuals of chunks [chunk_begin, chunk_end) in the CUDA path's reduction shape; the range
 * must be an index-aligned power-of-two span of 128-point chunks (or its tail): one rank's share under point sharding */
void oracle_tree_partial(const double *resid, int64_t n, int64_t chunk_begin, int64_t chunk_end, double *sum, double *mx)
{
    int64_t span = 1;
    while (span < chunk_end - chunk_begin) span <<= 1;
    double s = chunk_begin < chunk_end ? tree_sum(resid, n, chunk_begin, span) : -1.0;
    double m = 0.0;
    int64_t hi = chunk_end * 128 < n ? chunk_end * 128 : n;
    for (int64_t i = chunk_begin * 128; i < hi; i++) if (resid[i] > m) m = resid[i];
    *sum = s < 0.0 ? 0.0 : s;
    *mx = m;
}

double oracle_compute_residual(double expected, double actual) { return compute_residual(expected, actual); }

double oracle_score_from_f(const double *f, const double *ys, int64_t n, int leafs, int sum_mode)
{
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double s = score_from_f(f, ys, n, leafs, sum_mode, scratch);
    free(scratch);
    return s;
}

static double score_one(const char *expr, const double *xs, const double *ys, int64_t n, int max_leafs,
                        int sum_mode, double *f, double *scratch)
{
    t_jcnt = (uint64_t)g_jitter * 0xD1B54A32D192ED03ull + strlen(expr);
    node *e = parse(expr);
    if (!e) return MIN_SCORE;                       /* any exception => min-score (ops.clj:134-136) */
    double s;
    if (e->leafs > max_leafs) s = MIN_SCORE;        /* ops.clj:128-129 */
    else if (eval_vec(e, xs, n, f) != 0) s = MIN_SCORE; /* nil / throw => min-score (ops.clj:131-136) */
    else s = score_from_f(f, ys, n, e->leafs, sum_mode, scratch);
    free_node(e);
    return s;
}

double oracle_score(const char *expr, const double *xs, const double *ys, int64_t n, int max_leafs, int sum_mode)
{
    double *f = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * 2);
    double s = score_one(expr, xs, ys, n, max_leafs, sum_mode, f, f + (n > 0 ? n : 1));
    free(f);
    return s;
}

/* pmap-style chunk size, ga.clj:49-55 */
static int pop_chunk(int64_t pop) { return pop >= 10000 ? 100 : pop >= 2000 ? 50 : pop >= 500 ? 20 : 10; }

/* score a whole population; exprs = n_expr NUL-terminated strings. Returns the thread count used. */
typedef struct {
    const char *const *exprs; int64_t n_expr; const double *xs, *ys; int64_t n;
    int max_leafs, sum_mode, chunk; double *out; atomic_llong *next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    size_t nn = (size_t)(j->n > 0 ? j->n : 1);
    double *f = (double *)malloc(sizeof(double) * nn * 2);
    int64_t nchunks = (j->n_expr + j->chunk - 1) / j->chunk;
    for (;;) {
        int64_t c = atomic_fetch_add(j->next, 1);
        if (c >= nchunks) break;
        int64_t hi = (c + 1) * j->chunk < j->n_expr ? (c + 1) * j->chunk : j->n_expr;
        for (int64_t i = c * j->chunk; i < hi; i++)
            j->out[i] = score_one(j->exprs[i], j->xs, j->ys, j->n, j->max_leafs, j->sum_mode, f, f + nn);
    }
    free(f);
    return NULL;
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_score_batch(const char *const *exprs, int64_t n_expr, const double *xs, const double *ys, int64_t n,
                       int max_leafs, int sum_mode, int nthreads, double *out_scores)
{
    if (nthreads <= 0) nthreads = oracle_max_threads();
    if (nthreads > 1024) nthreads = 1024;
    atomic_llong next = 0;
    batch_job job = { exprs, n_expr, xs, ys, n, max_leafs, sum_mode, pop_chunk(n_expr), out_scores, &next };
    pthread_t th[1024];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &job);
    batch_worker(&job);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return nthreads;
}
