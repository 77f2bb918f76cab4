// kernels.cu -- sm_100a kernels of the scoring path.
//
// score_kernel: one CTA of up to 32 warps per (expression, tile of points).  The program and its constant
// pool are staged once in shared memory (two TMA bulk copies, cp.async.bulk + mbarrier) and every warp of the CTA interprets it on different 128-point chunks
// (so the warps of a CTA fetch the same instructions at about the same time: 4x fewer distinct instruction
// streams per SM than one program per warp).  Lanes stride over the points, 4 consecutive points per lane per
// pass (two 16-byte loads of xs and of ys per lane, a warp reads 1 KiB contiguous); the residuals are reduced
// with a fixed-shape tree (lane-local left fold of 4, xor-butterfly over the warp, index-aligned binary tree
// over the 128-point chunks) so a score is bit-identical for every tiling, sharding and GPU count.  The epilogue reproduces
// closyr's formula operation by operation (reference src/closyr/ops.clj:78-116); the file is compiled with
// -fmad=false so none of it is contracted.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "vm.cuh"

namespace closyr {
namespace {

using namespace dev;

constexpr unsigned FULL = 0xffffffffu;

// CLOSYR_TRACE builds (scripts/build_trace.sh, never the shipped library): per-CTA phase timestamps of score_kernel
#ifdef CLOSYR_TRACE
constexpr int TRACE_ITEMS = 32768, TRACE_SLOTS = 16;
__device__ unsigned long long g_trace[TRACE_ITEMS * TRACE_SLOTS];
#define TR(k) do { if (threadIdx.x == 0 && blockIdx.x < TRACE_ITEMS) g_trace[blockIdx.x * TRACE_SLOTS + (k)] = clock64(); } while (0)
#else
#define TR(k) do { } while (0)
#endif
constexpr int MAX_LEVELS = 40;
constexpr int MAX_SCORE_WARPS = 32;  // warps per CTA = warps interpreting one staged program (runtime, power of two)

__device__ __forceinline__ double warp_sum(double s)
{
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) s = s + __shfl_xor_sync(FULL, s, st);
    return s;
}
__device__ __forceinline__ double warp_max(double m)
{
#pragma unroll
    for (int st = 16; st >= 1; st >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, st));
    return m;
}

// ops.clj:98-116 with the exact operation order; leafs is Symja's leafCount (ops.clj:127)
__device__ __forceinline__ double score_formula(double sum, double mx, int64_t n, int leafs)
{
    const double mean = sum / (double)n;
    const double score = -1.0 * ((2.0 * mean) + mx);
    double pen = (0.0000001 * (double)leafs) * (double)leafs;
    if (pen > 0.1) pen = 0.1;
    const double ded = fabs(score) * pen;
    return score - ded;
}

struct Staged {
    const double *sconst;      // the program's constant pool in global memory (only LDZ and the quirk table read it)
    uint32_t tcode_addr;       // shared-window address of the threaded code (vm.cuh)
    uint32_t slot_addr;        // shared-window address of this lane's column of the warp's spill slots
    uint32_t pflags;           // PF_* properties of the program
    const double *leaf_tab;    // f(x_i) tables of the resident points (ScoreParams), nullptr when xs are ad hoc
};

constexpr int SLOT_DOUBLES_PER_WARP = CLOSYR_REG_SLOTS * POINTS_PER_LANE * 32;   // 4 KiB
constexpr int LIST_DOUBLES_PER_WARP = CHUNK_POINTS;                              // 1 KiB: the reported points of one chunk

// shared memory of a CTA: [warp slots | warp lists | `nlvl` chunk sums | threaded code (ncode + 1 entries: the dispatch
// loop prefetches one entry past END)]
struct SmemLayout {
    double *list, *slvl;
    TInsn *tcode;
};
__device__ __forceinline__ SmemLayout smem_layout(unsigned char *smem, int nwarps, int nlvl)
{
    SmemLayout L;
    L.list = reinterpret_cast<double *>(smem) + nwarps * SLOT_DOUBLES_PER_WARP;
    L.slvl = L.list + nwarps * LIST_DOUBLES_PER_WARP;
    L.tcode = reinterpret_cast<TInsn *>(L.slvl + ((nlvl + 1) & ~1));      // 16-byte aligned
    return L;
}

// Translates every program of a shard into threaded code once per upload (the layout the interpreters dispatch on:
// vm.cuh decode_insn) and writes one descriptor per expression.  One warp per expression.
__global__ void __launch_bounds__(256) prepare_kernel(const PrepareParams p)
{
    const int lane = threadIdx.x & 31;
    const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= p.pop.n_expr) return;
    const uint32_t c0 = p.pop.code_off[e], c1 = p.pop.code_off[e + 1];
    const uint32_t k0 = p.pop.const_off[e], k1 = p.pop.const_off[e + 1];
    const uint32_t ncode = c1 - c0;
    const uint32_t *gc = p.pop.code + (c0 - p.pop.code_base);
    const double *gk = p.pop.consts + (k0 - p.pop.const_base);
    TInsn *out = reinterpret_cast<TInsn *>(p.tcode) + (c0 - p.pop.code_base);
    uint32_t pf = 0, fm = 0;          // fm: bit (TH_f - TH_RECIP) = the program calls the routine of function f
    for (uint32_t i = lane; i < ncode; i += 32) {
        TInsn t;
        pf |= decode_insn(gc[i], gk, k1 - k0, p.have_tab != 0, p.leaf_stride, t);
        out[i] = t;
        // a function of the raw variable that is real everywhere comes from its table and never runs a routine
        const uint32_t f = (t.h == TH_MOD || t.h == TH_TABC) ? ((t.aux >> 2) & 0xffu) : t.h;
        if (f >= TH_RECIP && f < TH_COUNT) fm |= 1u << (f - TH_RECIP);
    }
    pf = __reduce_or_sync(FULL, pf);
    fm = __reduce_or_sync(FULL, fm);
    {
        constexpr uint32_t DOMAIN = 1u << (TH_SQRT - TH_RECIP) | 1u << (TH_RSQRT - TH_RECIP) | 1u << (TH_POWC - TH_RECIP) | 1u << (TH_LOG - TH_RECIP) |
                                    1u << (TH_ASIN - TH_RECIP) | 1u << (TH_ACOS - TH_RECIP) | 1u << (TH_LDZ - TH_RECIP) | 1u << (TH_LDSPEC - TH_RECIP) |
                                    1u << (TH_POWS - TH_RECIP);
        if (fm & DOMAIN) pf |= PF_DOMAIN;
    }
    if (lane == 0) {
        // grouping key, heaviest routines first: programs with equal keys walk through the same code
        auto has = [fm](uint32_t a, uint32_t b) { return ((fm >> (a - TH_RECIP)) | (fm >> (b - TH_RECIP))) & 1u; };
        const uint32_t len = ncode / 10u;
        p.key[e] = ((pf & (PF_NO_FAST | PF_DEEP)) ? 1u << 11 : 0u) | (has(TH_POWC, TH_POWS) | has(TH_LDZ, TH_LDSPEC)) << 10 |
                   has(TH_LOG, TH_LOG) << 9 | has(TH_SQRT, TH_RSQRT) << 8 | has(TH_ASIN, TH_ASIN) << 7 | has(TH_ACOS, TH_ACOS) << 6 |
                   has(TH_TAN, TH_TAN) << 5 | has(TH_EXP, TH_EXP) << 4 | has(TH_SIN, TH_COS) << 3 | has(TH_RECIP, TH_POWI) << 2 |
                   (len < 3u ? len : 3u);
        ExprDesc d;
        d.tc_off = c0 - p.pop.code_base;
        d.ncode_pf = ncode | (pf << 16);
        d.meta = p.pop.meta[e];
        d.k_off = k0 - p.pop.const_base;
        p.desc[e] = d;
    }
}

// Counting sort of the grouping keys: order[] lists the expressions key by key (the order inside a key is whatever the
// atomics give: a score does not depend on which CTA computes it).  One CTA; 100K expressions take ~20 us.
__global__ void __launch_bounds__(1024) order_kernel(const uint32_t *__restrict__ key, const int64_t n, uint32_t *__restrict__ order)
{
    constexpr int BINS = 1 << ORDER_KEY_BITS, PER = BINS / 1024;
    __shared__ uint32_t bins[BINS];
    __shared__ uint32_t wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = tid; b < BINS; b += 1024) bins[b] = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += 1024) atomicAdd(&bins[key[i] & (BINS - 1)], 1u);
    __syncthreads();
    uint32_t mine[PER], tot = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { mine[k] = tot; tot += bins[tid * PER + k]; }
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t w = wsum[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, wi, d);
        if (lane >= d) wi += t;
    }
    const uint32_t base = __shfl_sync(FULL, wi - w, warp) + incl - tot;
#pragma unroll
    for (int k = 0; k < PER; k++) bins[tid * PER + k] = base + mine[k];
    __syncthreads();
    for (int64_t i = tid; i < n; i += 1024) order[atomicAdd(&bins[key[i] & (BINS - 1)], 1u)] = (uint32_t)i;
}

// The slow half of a chunk, for the warp that owns it: some of its points were reported by the fast pass (or the chunk has
// no fast pass at all).  The reported points are compacted inside the warp and re-evaluated with the complete algebra,
// 64 points (2 per lane) per pass; then the chunk's residuals again from the merged values.  No other warp is involved:
// a CTA's warps take chunks from a shared counter, so a warp that is busy here simply takes fewer chunks.
// ev = this lane's points that need the complete algebra, v = the fast pass' values of the others.
// `mode` (shared, may be nullptr) is the CTA's hint "chunks around here leave the real line almost everywhere": while
// it is set the fast pass is skipped.  Both evaluators give bit-identical values wherever the fast path is valid, so
// the hint cannot change a score.
// A function of its own: the common path of the kernel keeps none of this state in registers.
constexpr int CPLX_POINTS = 2;       // points per lane of the complete-algebra interpreter
struct WarpSlowResult { double v[4]; double mx; uint32_t nr, used; };

__device__ __noinline__ WarpSlowResult warp_slow(const double *__restrict__ sconst, const uint32_t tcode_addr,
                                                 const uint32_t slot_addr, const double *__restrict__ leaf_tab, const bool deep,
                                                 double *__restrict__ list, uint32_t *mode, const bool ran_fast,
                                                 const double *__restrict__ xs, const double *__restrict__ ys,
                                                 const int64_t base, const uint32_t valid, const uint32_t ev_in,
                                                 const uint32_t quirk_off_plus1, const uint32_t sum_addr,
                                                 const double v0, const double v1, const double v2, const double v3)
{
    const int lane = threadIdx.x & 31;
    WarpSlowResult out;
    uint32_t ev = ev_in;
    double v[4] = { v0, v1, v2, v3 };
    if (quirk_off_plus1) {
        // eval.clj:109: constant AST whose argument count equals the point count is decoded per argument
        ev = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = ((valid >> j) & 1u) ? sconst[quirk_off_plus1 - 1 + base + j] : 0.0;
    }
    // warp-wide compaction of the reported points
    const uint32_t mine = __popc(ev);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t pos = incl - mine;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if ((ev >> j) & 1u) list[pos++] = __longlong_as_double(base + j);      // the point's index
    __syncwarp();
    uint32_t left_real = 0;
    constexpr uint32_t TASK = 32 * CPLX_POINTS;
    for (uint32_t t0 = 0; t0 < total; t0 += TASK) {
        CplxArgs<CPLX_POINTS> ca;
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            ca.idx[j] = __double_as_longlong(list[r < total ? r : t0]);          // idle lanes repeat the task's first point
        }
        const CplxResult<CPLX_POINTS> cr = deep ? run_cplx<CPLX_POINTS, true>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab)
                                                : run_cplx<CPLX_POINTS, false>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab);
        left_real |= cr.left_real;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            if (r < total) list[r] = ((cr.nonreal >> j) & 1u) ? __longlong_as_double(0x7ff8000000000000ll) : cr.out[j];   // NaN marks "not a real number"
        }
    }
    __syncwarp();
    pos = incl - mine;
    out.nr = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if ((ev >> j) & 1u) {
            const double r = list[pos++];
            if (r != r) { v[j] = __longlong_as_double(0x7ff0000000000000ll); out.nr |= 1u << j; }
            else v[j] = r;
        }
    }
    __syncwarp();                                   // the list is reused by the warp's next chunk
    out.used = total != 0 ? 1u : 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) out.v[j] = v[j];
    out.mx = chunk_residuals(v, ys + base, valid, sum_addr);
    if (mode && lane == 0) {
        if (ran_fast) { if (total * 4u >= (uint32_t)CHUNK_POINTS * 3u) *(volatile uint32_t *)mode = 1u; }     // >= 75 % of the chunk
        else if (!left_real) *(volatile uint32_t *)mode = 0u;
    }
    return out;
}

// ---- TMA bulk copy (cp.async.bulk + mbarrier): the threaded code of a program is 16-byte aligned and a whole number
// of 16-byte entries
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// largest of a warp's non-negative doubles: their bit patterns order like their values
__device__ __forceinline__ double warp_max_nonneg(double m)
{
    const uint32_t hi = (uint32_t)__double2hiint(m);
    const uint32_t mh = __reduce_max_sync(FULL, hi);
    const uint32_t ml = __reduce_max_sync(FULL, hi == mh ? (uint32_t)__double2loint(m) : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

// One round of a CTA = one 128-point chunk per warp: v[j] = f(x[j]) for the lane's 4 points (+Inf where the value is
// not a real number, as eval.clj:70-80 has it), nr = mask of those points.
//
// Every warp first runs the IEEE-real fast path on its chunk.  The points it reports are compacted across
// the WHOLE CTA into one list and re-evaluated with the complete algebra, 64 points (2 per lane) per warp and
// pass -- so the extra work of a chunk that left the real line is shared by all warps instead of stalling the
// CTA behind one of them.  `mode` is the CTA's adaptive state: when more than half of the points of the previous round needed
// the complete algebra the fast pass is skipped.  Both evaluators give bit-identical values wherever the fast
// path is valid, so the heuristic cannot change a score.
// Must be called by every thread of the CTA (it contains __syncthreads).
struct Round {
    double *list;          // [warps * 128] indices of the reported points, replaced in place by their results
    uint32_t *counts;      // [2][32] reported points per warp, double-buffered by round parity
};

struct CplxPassResult { double v[4]; uint32_t nr, mode, used; };

// The cold half of a round: some point of the CTA was reported.  A function of its own so that the state it needs
// does not live in the registers of the common path.
__device__ __noinline__ CplxPassResult cplx_pass(const double *__restrict__ sconst, const uint32_t tcode_addr,
                                                 const uint32_t slot_addr, const double *__restrict__ leaf_tab, const bool deep,
                                                 double *__restrict__ list, uint32_t *__restrict__ cnt,
                                                 const double *__restrict__ xs, const int64_t base, const uint32_t ev, const uint32_t round_points_first,
                                                 const double v0, const double v1, const double v2, const double v3)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t round_points = round_points_first >> 1;       // 0: the round had a fast pass
    const bool first_round = (round_points_first & 1u) != 0;     // of this CTA's tile
    const bool ran_fast = round_points == 0;
    CplxPassResult res;
    res.v[0] = v0; res.v[1] = v1; res.v[2] = v2; res.v[3] = v3;
    res.nr = 0;
    const uint32_t mine = __popc(ev);
    uint32_t total, pos;
    if (ran_fast) {
        // CTA-wide compaction of the reported points
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) cnt[warp] = incl;
        __syncthreads();
        uint32_t cw = lane < nwarps ? cnt[lane] : 0u, ci = cw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, ci, d);
            if (lane >= d) ci += t;
        }
        total = __shfl_sync(FULL, ci, 31);
        pos = __shfl_sync(FULL, ci - cw, warp) + incl - mine;
    } else {
        // a round without a fast pass: every point of the round is on the list (round_points of them: the points of a run
        // are a prefix of its chunks), each at its own offset -- nothing to count
        total = round_points;
        pos = (uint32_t)warp * CHUNK_POINTS + (uint32_t)lane * POINTS_PER_LANE;
    }
    // the round's first point (rounds without a fast pass: list entry r is point round_base + r, nothing to store or wait for)
    const int64_t round_base = base - (int64_t)((uint32_t)warp * CHUNK_POINTS + (uint32_t)lane * POINTS_PER_LANE);
    if (ran_fast) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((ev >> j) & 1u) list[pos++] = __longlong_as_double(base + j);      // the point's index
        __syncthreads();
    } else {
        pos += mine;
        // the list still holds the previous round's results until every warp has read its own back
        if (!first_round) __syncthreads();
    }
#ifdef CLOSYR_TRACE
    if (threadIdx.x == 0 && blockIdx.x < TRACE_ITEMS && base < 4096) { g_trace[blockIdx.x * TRACE_SLOTS + 11] = total; g_trace[blockIdx.x * TRACE_SLOTS + 13] = clock64(); }
#endif
    constexpr uint32_t TASK = 32 * CPLX_POINTS;
    uint32_t left = 0;                                                             // a task of this warp left real mode
    for (uint32_t tb = 0; tb < total; tb += nwarps * TASK) {
        if (tb) __syncthreads();                                                 // the passes of a round start in step
        const uint32_t t0 = tb + warp * TASK;
        if (t0 >= total) continue;
        CplxArgs<CPLX_POINTS> ca;
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            const uint32_t q = r < total ? r : t0;                               // idle lanes repeat the task's first point
            ca.idx[j] = ran_fast ? __double_as_longlong(list[q]) : round_base + q;
        }
        const CplxResult<CPLX_POINTS> cr = deep ? run_cplx<CPLX_POINTS, true>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab)
                                                : run_cplx<CPLX_POINTS, false>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab);
        left |= cr.left_real;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            if (r < total) list[r] = ((cr.nonreal >> j) & 1u) ? __longlong_as_double(0x7ff8000000000000ll) : cr.out[j];   // NaN marks "not a real number"
        }
    }
#ifdef CLOSYR_TRACE
    if (threadIdx.x == 0 && blockIdx.x < TRACE_ITEMS && base < 4096) g_trace[blockIdx.x * TRACE_SLOTS + 14] = clock64();
#endif
    const uint32_t nleft = (uint32_t)__syncthreads_count(lane == 0 && left != 0);   // warps with a task that left real mode
#ifdef CLOSYR_TRACE
    if (threadIdx.x == 0 && blockIdx.x < TRACE_ITEMS && base < 4096) g_trace[blockIdx.x * TRACE_SLOTS + 15] = clock64();
#endif
    pos -= mine;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if ((ev >> j) & 1u) {
            const double r = list[pos++];
            if (r != r) { res.v[j] = __longlong_as_double(0x7ff0000000000000ll); res.nr |= 1u << j; }
            else res.v[j] = r;
        }
    }
    // the next round of the tile (points are sorted by x, neighbouring rounds behave alike).  After a fast pass: skip it next time
    // when more than half of the round's points were reported -- two passes of the complete algebra are certain then.  After
    // a round without one: try it again when most of the warps' tasks never left real mode.
    {
        const uint32_t busy = (total + TASK - 1) / TASK < (uint32_t)nwarps ? (total + TASK - 1) / TASK : (uint32_t)nwarps;
        res.mode = ran_fast ? ((total * 2u > (uint32_t)blockDim.x * 4u) ? 1u : 0u) : ((nleft * 2u > busy) ? 1u : 0u);
    }
    res.used = nleft != 0 ? 1u : 0u;
    return res;
}

// The slow half of a round, for one warp: called by every warp of the CTA when any point of the round was reported (or
// when the round has no fast pass at all).  Compaction + complete algebra (cplx_pass), then the chunk's residuals again
// from the merged values.  ev = this lane's points that need the complete algebra, v = the fast pass' values of the
// others.  A function of its own: the common path of the kernel keeps none of this state in registers.
struct RoundSlowResult { double v[4]; double mx; uint32_t nr, mode, used; };

__device__ __noinline__ RoundSlowResult round_slow(const double *__restrict__ sconst, const uint32_t tcode_addr,
                                                   const uint32_t slot_addr, const double *__restrict__ leaf_tab, const bool deep,
                                                   double *__restrict__ list, uint32_t *__restrict__ cnt,
                                                   const double *__restrict__ xs, const double *__restrict__ ys,
                                                   const int64_t base, const uint32_t valid, const uint32_t ev_in, const uint32_t round_points,
                                                   const uint32_t quirk_off_plus1, const uint32_t sum_addr,
                                                   const double v0, const double v1, const double v2, const double v3)
{
    RoundSlowResult out;
    uint32_t ev = ev_in;
    double v[4] = { v0, v1, v2, v3 };
    uint32_t rp = round_points;
    if (quirk_off_plus1) {
        // eval.clj:109: constant AST whose argument count equals the point count is decoded per argument
        ev = 0;
        rp = 0;                                       // nothing goes to the complete algebra: an empty compaction
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = ((valid >> j) & 1u) ? sconst[quirk_off_plus1 - 1 + base + j] : 0.0;
    }
    const CplxPassResult cp = cplx_pass(sconst, tcode_addr, slot_addr, leaf_tab, deep, list, cnt, xs, base, ev, rp, v[0], v[1], v[2], v[3]);
    out.nr = cp.nr;
    out.mode = cp.mode;
    out.used = cp.used;                          // the complete algebra was NEEDED: some evaluation left the real line
    out.mx = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) out.v[j] = cp.v[j];
    if (__any_sync(FULL, valid != 0)) out.mx = chunk_residuals(cp.v, ys + base, valid, sum_addr);      // not for an idle warp
    return out;
}


// stages the threaded code of a program with one TMA bulk copy; every thread of the CTA calls it, thread 0 issues the
// copy and everybody waits on the mbarrier (single use: parity 0)
__device__ __forceinline__ void stage_tcode_tma(TInsn *dst, const uint4 *src, uint32_t ncode, unsigned long long *bar, int tid)
{
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, ncode * 16u);
        bulk_g2s(dst, src, ncode * 16u, bar);
    }
    __syncthreads();                 // the initialised barrier is visible to every waiter
    mbar_wait(bar, 0);
}

// what both scoring kernels write for a finished (expression, tile)
__device__ __forceinline__ void emit_tile(const ScoreParams &p, int64_t e, int tile, int leafs, bool const_throws,
                                          double tile_sum, double tile_max, uint32_t tile_nonreal, bool tile_full)
{
    if (p.final_scores) {
        uint8_t rf = tile_full ? CL_R_FULL_SEMANTICS : 0;
        double score;
        if (const_throws) {                                            // eval.clj:83-96 throws => min-score
            score = CLOSYR_MIN_SCORE;
            rf |= CL_R_MIN_SCORE;
        } else score = score_formula(tile_sum, tile_max, p.n, leafs);
        p.scores[e] = score;
        p.rflags[e] = rf;
    } else {
        double *out = p.partials + ((p.partial_tile_offset + tile) * p.pop.n_expr + e) * 3;
        out[0] = tile_sum;
        out[1] = tile_max;
        out[2] = (double)tile_nonreal + (tile_full ? 0.25 : 0.0);     // fraction carries the diagnostic bit
    }
}

// index-aligned pairwise tree over the chunk sums (a missing right child leaves the left one unchanged); one warp
__device__ __forceinline__ void tree_reduce_chunks(double *slvl, int cnt, int lane)
{
    for (int stride = 1; stride < cnt; stride <<= 1) {
        for (int i = lane * 2 * stride; i + stride < cnt; i += 64 * stride) slvl[i] = slvl[i] + slvl[i + stride];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// score_kernel: one CTA of up to 32 warps per (expression, tile of points), ONE program per SM.  The program's threaded
// code (made once per upload by prepare_kernel) arrives with one TMA bulk copy.  The warps move through the tile in
// rounds of one 128-point chunk per warp: fast pass, CTA-wide compaction of the reported points, complete algebra.
// Why one program per SM and rounds in step: the complete-algebra interpreter and the complex functions behind it are
// ~150 KB of code.  Measured in r02 (profiles/r02_design_experiments.md): warps that sit in different places of it --
// several programs per SM, or one program whose warps take chunks on their own -- wait for instructions most of the
// time (47 stalled warp-cycles per issued instruction on `no_instruction`); only the fast interpreter is small enough
// for that, and a fast-only kernel in front of this one did not pay for the second pass over the marked chunks.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAX_SCORE_WARPS * 32, 1) score_kernel(const ScoreParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double s_max[MAX_SCORE_WARPS];
    __shared__ uint32_t s_nonreal[MAX_SCORE_WARPS], s_full[MAX_SCORE_WARPS];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_counts[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int64_t item = blockIdx.x;
    const int64_t slot = item / p.n_tiles;
    const int tile = (int)(item - slot * p.n_tiles);
    const int64_t e = p.pop.order ? (int64_t)__ldg(p.pop.order + slot) : slot;
    TR(1);
#ifdef CLOSYR_TRACE
    if (tid == 0 && blockIdx.x < TRACE_ITEMS) {
        uint32_t smid;
        unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_trace[blockIdx.x * TRACE_SLOTS + 0] = smid;
        g_trace[blockIdx.x * TRACE_SLOTS + 10] = gt;
    }
#endif
    const uint4 dq = __ldg(reinterpret_cast<const uint4 *>(p.pop.desc + e));
    const uint32_t ncode = dq.y & 0xffffu, pflags = dq.y >> 16, meta = dq.z;
    const int leafs = (int)(meta & 0xffffu);
    const uint32_t flags = (meta >> 16) & 0xffu;
#ifdef CLOSYR_TRACE
    if (ncode == 0xfffff) return;
#endif
    TR(2);

    const bool too_big = leafs > p.max_leafs;                               // ops.clj:128-129
    const bool dead = (flags & (CL_F_UNSUPPORTED | CL_F_NIL)) != 0;        // eval nil / unsupported => min-score
    if (too_big || dead || p.n <= 0) {
        if (tid == 0) {
            if (p.final_scores) {
                p.scores[e] = CLOSYR_MIN_SCORE;
                p.rflags[e] = too_big ? CL_R_TOO_BIG : CL_R_MIN_SCORE;
            } else {
                double *out = p.partials + ((p.partial_tile_offset + tile) * p.pop.n_expr + e) * 3;
                out[0] = 0.0; out[1] = 0.0; out[2] = 0.0;
            }
        }
        return;
    }

    // shared memory: [warp slots | compaction list | chunk sums (+ a scratch word) | threaded code]
    double *const list = reinterpret_cast<double *>(smem) + nwarps * SLOT_DOUBLES_PER_WARP;
    double *const slvl = list + nwarps * CHUNK_POINTS;
    TInsn *const tcode = reinterpret_cast<TInsn *>(slvl + ((p.chunks_per_tile + 2) & ~1));
    stage_tcode_tma(tcode, p.pop.tcode + dq.x, ncode, &s_bar, tid);
    TR(3);

    const double *const sconst = p.pop.consts + dq.w;
    const uint32_t tcode_addr = smem_u32(tcode);
    const uint32_t slot_addr = smem_u32(reinterpret_cast<double *>(smem) + warp * SLOT_DOUBLES_PER_WARP + lane);
    const bool is_const = (flags & CL_F_CONST) != 0;
    uint32_t quirk_off_plus1 = 0;
    if (is_const) {
        const uint32_t aux = p.pop.aux[e];
        if ((aux >> 16) != 0 && (int64_t)(aux >> 16) == p.n) quirk_off_plus1 = (aux & 0xffffu) + 1u;
    }
    const int64_t c_lo = p.chunk_begin + (int64_t)tile * p.chunks_per_tile;
    int64_t c_hi = c_lo + p.chunks_per_tile;
    if (c_hi > p.chunk_end) c_hi = p.chunk_end;
    const int cnt = (int)(c_hi - c_lo);
    const uint32_t slvl_addr = smem_u32(slvl);
    const bool fast_ok = !is_const && !(pflags & PF_NO_FAST);
    double warp_mx = 0.0;
    // a program that calls a function with a real domain starts WITHOUT the fast pass: three of four such programs of the GA
    // mix leave the real line somewhere, and then the fast pass is code and time on top of the complete algebra
    // (profiles/r02_design_experiments.md, "the fast pass as a choice per program")
    uint32_t nonreal = 0, mode = (pflags & PF_DOMAIN) ? 1u : 0u, parity = 0, used_full = 0;
    for (int r0 = 0; r0 < cnt; r0 += nwarps, parity ^= 1u) {
        const int ci = r0 + warp;
        const bool have = ci < cnt;
        const int64_t base = (c_lo + ci) * CHUNK_POINTS + lane * POINTS_PER_LANE;
        // xs, ys and every table row are padded to whole chunks (x = 1.0, y = 0.0), so a lane always loads its 4 points
        const int64_t left = have ? p.n - base : 0;
        const uint32_t valid = left >= 4 ? 0xfu : (left <= 0 ? 0u : (1u << (int)left) - 1u);
        const uint32_t sum_addr = slvl_addr + (uint32_t)(have ? ci : p.chunks_per_tile) * 8u;      // idle warps: the scratch word
        const bool fast = fast_ok && mode == 0;
        RealResult<4> rr;
        rr.ev = valid;
        rr.mx = 0.0;
        rr.v[0] = rr.v[1] = rr.v[2] = rr.v[3] = 0.0;
        if (fast && have)
            rr = nwarps < MAX_SCORE_WARPS
                     ? run_real<4, true>(tcode_addr, slot_addr, p.xs + base, p.leaf_tab ? p.leaf_tab + base : nullptr, p.ys + base, valid, sum_addr)
                     : run_real<4, false>(tcode_addr, slot_addr, p.xs + base, p.leaf_tab ? p.leaf_tab + base : nullptr, p.ys + base, valid, sum_addr);
        if (r0 == 0) TR(4);
        // nothing reported anywhere in the CTA (the common round): one barrier, the fast pass has done everything
        if (fast && !__syncthreads_or(rr.ev != 0)) {
            if (r0 == 0) TR(5);
            if (r0 < 4 * nwarps) TR(12 + r0 / nwarps);
            warp_mx = max_nonneg(warp_mx, rr.mx);
            continue;
        }
        if (r0 == 0) TR(5);
        // without a fast pass every point of the round goes to the complete algebra: their number (0 = there was a fast pass)
        uint32_t round_points = 0;
        if (!fast) {
            const int64_t rb = (c_lo + r0) * CHUNK_POINTS, rc = (int64_t)(cnt - r0 < nwarps ? cnt - r0 : nwarps) * CHUNK_POINTS;
            const int64_t nl = p.n - rb;
            round_points = (uint32_t)(nl < rc ? (nl > 0 ? nl : 0) : rc) << 1 | (r0 == 0 ? 1u : 0u);     // bit 0: the tile's first round
        }
        const RoundSlowResult rs = round_slow(sconst, tcode_addr, slot_addr, p.leaf_tab, (pflags & PF_DEEP) != 0, list, s_counts + parity * 32, p.xs, p.ys,
                                              base, valid, rr.ev, round_points, quirk_off_plus1, sum_addr, rr.v[0], rr.v[1], rr.v[2], rr.v[3]);
        warp_mx = max_nonneg(warp_mx, rs.mx);
        nonreal += __popc(rs.nr);
        used_full |= rs.used;
        mode = rs.mode;
        if (r0 == 0) TR(6);
        if (r0 < 4 * nwarps) TR(12 + r0 / nwarps);
    }
    TR(7);
    warp_mx = warp_max_nonneg(warp_mx);
    nonreal = __reduce_add_sync(FULL, nonreal);
    if (lane == 0) { s_max[warp] = warp_mx; s_nonreal[warp] = nonreal; s_full[warp] = used_full; }
    __syncthreads();
    TR(8);
    if (warp != 0) return;
    tree_reduce_chunks(slvl, cnt, lane);
    const double tile_max = warp_max_nonneg(lane < nwarps ? s_max[lane] : 0.0);
    const uint32_t tile_nonreal = __reduce_add_sync(FULL, lane < nwarps ? s_nonreal[lane] : 0u);
    const uint32_t tile_full = __reduce_or_sync(FULL, lane < nwarps ? s_full[lane] : 0u);
    if (lane != 0) return;
    emit_tile(p, e, tile, leafs, is_const && !quirk_off_plus1 && tile_nonreal != 0, cnt > 0 ? slvl[0] : 0.0, tile_max, tile_nonreal, tile_full != 0);
    TR(9);
}

// combines per-tile partials in global tile order with the same aligned binary tree, then the epilogue
__global__ void finalize_kernel(const FinalizeParams p)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_expr) return;
    const uint32_t meta = p.meta[e];
    const int leafs = (int)(meta & 0xffffu);
    const uint32_t flags = (meta >> 16) & 0xffu;
    if (!p.out_partials) {
        if (leafs > p.max_leafs) { p.scores[e] = CLOSYR_MIN_SCORE; p.rflags[e] = CL_R_TOO_BIG; return; }
        if ((flags & (CL_F_UNSUPPORTED | CL_F_NIL)) || p.n <= 0) { p.scores[e] = CLOSYR_MIN_SCORE; p.rflags[e] = CL_R_MIN_SCORE; return; }
    }
    double lvl[32];
    double mx = 0.0, nonreal = 0.0;
    bool used_full = false;
    for (int t = 0; t < p.n_tiles_total; t++) {
        const double *in = p.partials + ((int64_t)t * p.part_stride + e) * 3;
        double s = in[0];
        mx = fmax(mx, in[1]);
        const double nr = floor(in[2]);
        used_full |= in[2] != nr;
        nonreal += nr;
        int k = 0;
        for (uint32_t c = (uint32_t)t; c & 1u; c >>= 1, k++) s = lvl[k] + s;
        lvl[k] = s;
    }
    double sum = 0.0;
    bool have = false;
    for (int k = 0; ((uint32_t)p.n_tiles_total >> k) != 0; k++) {
        if (((uint32_t)p.n_tiles_total >> k) & 1u) {
            sum = have ? lvl[k] + sum : lvl[k];
            have = true;
        }
    }
    if (p.out_partials) {
        double *out = p.out_partials + e * 3;
        out[0] = sum;
        out[1] = mx;
        out[2] = nonreal + (used_full ? 0.25 : 0.0);
        return;
    }
    uint8_t rf = used_full ? CL_R_FULL_SEMANTICS : 0;
    double score;
    if ((flags & CL_F_CONST) && nonreal != 0.0) { score = CLOSYR_MIN_SCORE; rf |= CL_R_MIN_SCORE; }
    else score = score_formula(sum, mx, p.n, leafs);
    p.scores[e] = score;
    p.rflags[e] = rf;
}

// eval-vec-pheno for a batch (eval.clj:99-112): out_f[e][i] = real value or +Inf; status 0 ok, 1 nil, 2 throws
__global__ void __launch_bounds__(32) eval_kernel(const EvalParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const int64_t e = blockIdx.x;
    const uint4 dq = __ldg(reinterpret_cast<const uint4 *>(p.pop.desc + e));
    const uint32_t ncode = dq.y & 0xffffu, pflags = dq.y >> 16, meta = dq.z;
    const uint32_t flags = (meta >> 16) & 0xffu;
    if (flags & (CL_F_UNSUPPORTED | CL_F_NIL)) {
        if (lane == 0) p.status[e] = 1;
        return;
    }
    const SmemLayout L = smem_layout(smem, 1, 2);
    for (uint32_t i = lane; i < ncode; i += 32) reinterpret_cast<uint4 *>(L.tcode)[i] = __ldg(p.pop.tcode + dq.x + i);
    __syncwarp();
    const double *const sconst = p.pop.consts + dq.w;
    const uint32_t tcode_addr = smem_u32(L.tcode);
    const uint32_t slot_addr = smem_u32(reinterpret_cast<double *>(smem) + lane);
    const bool is_const = (flags & CL_F_CONST) != 0;
    uint32_t quirk_off_plus1 = 0;
    if (is_const) {
        const uint32_t aux = p.pop.aux[e];
        if ((aux >> 16) != 0 && (int64_t)(aux >> 16) == p.n) quirk_off_plus1 = (aux & 0xffffu) + 1u;
    }
    const int64_t nchunks = (p.n + CHUNK_POINTS - 1) / CHUNK_POINTS;
    uint32_t nonreal = 0;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t base = c * CHUNK_POINTS + lane * POINTS_PER_LANE;
        const int64_t left = p.n - base;
        const uint32_t valid = left >= 4 ? 0xfu : (left <= 0 ? 0u : (1u << (int)left) - 1u);
        // every point through the complete algebra (the values are wanted, not the residuals: xs doubles as the y vector
        // and the chunk sum goes to a scratch word)
        const WarpSlowResult rs = warp_slow(sconst, tcode_addr, slot_addr, nullptr, (pflags & PF_DEEP) != 0, L.list, nullptr, false, p.xs, p.xs,
                                            base, valid, valid, quirk_off_plus1, smem_u32(L.slvl), 0.0, 0.0, 0.0, 0.0);
        nonreal += __popc(rs.nr);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((valid >> j) & 1u) p.out_f[e * p.n + base + j] = ((rs.nr >> j) & 1u) ? inf : rs.v[j];
    }
    nonreal = __reduce_add_sync(FULL, nonreal);
    if (lane == 0) p.status[e] = (is_const && !quirk_off_plus1 && nonreal != 0) ? 2 : 0;
}

// compute-score-from-actuals-and-expecteds (ops.clj:98-116) on given f values: one warp, same reduction tree
__global__ void __launch_bounds__(32) score_values_kernel(const double *f, const double *ys, int64_t n, int32_t leafs,
                                                          double *out3)
{
    __shared__ double slvl[MAX_LEVELS];
    const int lane = threadIdx.x;
    const int64_t nchunks = (n + CHUNK_POINTS - 1) / CHUNK_POINTS;
    double mx = 0.0;
    uint32_t cnt = 0;
    for (int64_t c = 0; c < nchunks; c++, cnt++) {
        const int64_t base = c * CHUNK_POINTS + lane * POINTS_PER_LANE;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double r = 0.0;
            if (base + j < n) {
                const double actual = f[base + j];
                r = (actual != actual) ? CLOSYR_MAX_RESID : fmin(CLOSYR_MAX_RESID, fabs(ys[base + j] - actual));
            }
            s = s + r;
            mx = fmax(mx, r);
        }
        s = warp_sum(s);
        uint32_t k = 0;
        for (uint32_t t = cnt; t & 1u; t >>= 1, k++) s = slvl[k] + s;
        __syncwarp();
        if (lane == 0) slvl[k] = s;
        __syncwarp();
    }
    double sum = 0.0;
    bool have = false;
    for (uint32_t k = 0; (cnt >> k) != 0; k++)
        if ((cnt >> k) & 1u) { sum = have ? slvl[k] + sum : slvl[k]; have = true; }
    mx = warp_max(mx);
    if (lane == 0) {
        out3[0] = n > 0 ? score_formula(sum, mx, n, leafs) : CLOSYR_MIN_SCORE;
        out3[1] = sum;
        out3[2] = mx;
    }
}

// register-resident DFMA chains: the measured denominator of the FP64 roofline
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double b = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// f(x_i) for the parameter-free elementary functions of the raw variable, through the very routines the
// interpreters use (f_tabulated), so a table entry is bit-identical to computing it in score_kernel.
// Entries past n (the table is padded to whole chunks) take x = 1.0, as the kernel does for absent points.
__global__ void __launch_bounds__(128) leaf_table_kernel(const double *__restrict__ xs, int64_t n, int64_t stride,
                                                           double *__restrict__ tab)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = q * 4;
    if (i0 >= stride) return;
    double in[4];
#pragma unroll
    for (int j = 0; j < 4; j++) in[j] = (i0 + j < n) ? xs[i0 + j] : 1.0;
    for (uint32_t op = CL_RECIP; op <= CL_ACOS; op++) {
        if (!((LEAF_TAB_MASK >> (op - CL_RECIP)) & 1u)) continue;
        double a[4] = { in[0], in[1], in[2], in[3] };
        f_tabulated<4>(op, a);
#pragma unroll
        for (int j = 0; j < 4; j++) tab[(int64_t)(op - CL_RECIP) * stride + i0 + j] = a[j];
    }
}

}  // namespace

#ifdef CLOSYR_TRACE
extern "C" int closyr_debug_ztags(unsigned long long *out)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, dev::g_ztags, sizeof(unsigned long long) * 8);
}
extern "C" int closyr_debug_trace(unsigned long long *out, int n_items)
{
    if (n_items > TRACE_ITEMS) n_items = TRACE_ITEMS;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * TRACE_SLOTS * n_items);
}
#endif

void launch_leaf_table(const double *xs, int64_t n, int64_t stride, double *tab, void *stream)
{
    if (stride <= 0) return;
    const int threads = 128;
    const int64_t quads = stride / 4;
    leaf_table_kernel<<<(unsigned)((quads + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(xs, n, stride, tab);
}

void launch_prepare(const PrepareParams &p, void *stream)
{
    if (p.pop.n_expr <= 0) return;
    const int threads = 256, per_block = threads / 32;
    prepare_kernel<<<(unsigned)((p.pop.n_expr + per_block - 1) / per_block), threads, 0, (cudaStream_t)stream>>>(p);
}

void launch_order(const uint32_t *key, int64_t n_expr, uint32_t *order, void *stream)
{
    if (n_expr <= 0) return;
    order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(key, n_expr, order);
}

int launch_score(const ScoreParams &p, uint32_t max_code_words, void *stream)
{
    const int64_t items = p.pop.n_expr * p.n_tiles;
    if (items <= 0) return 0;
    // [warp slots | compaction list | chunk sums + a scratch word | threaded code + the entry the dispatch loop prefetches past END]
    // CLOSYR_SMEM_PAD (bytes): fewer CTAs per SM, a knob for measurements (profiles/r02_design_experiments.md), not an API
    static const size_t pad = [] { const char *e = getenv("CLOSYR_SMEM_PAD"); return e ? (size_t)atol(e) : 0; }();
    const size_t smem = (size_t)p.warps * (SLOT_DOUBLES_PER_WARP + CHUNK_POINTS) * 8 + (size_t)((p.chunks_per_tile + 2) & ~1) * 8 +
                        ((size_t)max_code_words + 1) * sizeof(dev::TInsn) + 16 + pad;
    // opt in to > 48 KiB of dynamic shared memory; the attribute is per device, so it is set at every launch
    if (smem > 48 * 1024) cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    score_kernel<<<(unsigned)items, p.warps * 32, smem, (cudaStream_t)stream>>>(p);
    return 1;
}

void launch_finalize(const FinalizeParams &p, void *stream)
{
    if (p.n_expr <= 0) return;
    const int threads = 128;
    finalize_kernel<<<(unsigned)((p.n_expr + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(p);
}

void launch_eval(const EvalParams &p, size_t smem_bytes, void *stream)
{
    if (p.pop.n_expr <= 0) return;
    eval_kernel<<<(unsigned)p.pop.n_expr, 32, smem_bytes, (cudaStream_t)stream>>>(p);
}

void launch_score_values(const double *f, const double *ys, int64_t n, int32_t leafs, double *out3, void *stream)
{
    score_values_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, ys, n, leafs, out3);
}

void launch_fp64_probe(double *out, int iters, int blocks, void *stream)
{
    fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters);
}

size_t eval_smem_bytes(uint32_t max_code_words)
{
    // [warp slots | warp list | 2 scratch doubles | threaded code + the entry the dispatch loop prefetches past END]  (smem_layout)
    return (size_t)(SLOT_DOUBLES_PER_WARP + LIST_DOUBLES_PER_WARP) * 8 + 16 + ((size_t)max_code_words + 1) * sizeof(dev::TInsn) + 16;
}
int score_warps(int32_t chunks_per_tile)
{
    // CLOSYR_SCORE_WARPS (a power of two <= 32) caps the CTA size: a tuning knob for measurements, not an API
    static const int cap = [] {
        const char *e = getenv("CLOSYR_SCORE_WARPS");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= MAX_SCORE_WARPS && (v & (v - 1)) == 0) ? v : MAX_SCORE_WARPS;
    }();
    // Tiles of up to 8 chunks get TWO warps per chunk: most programs run such a tile without a fast pass (PF_DOMAIN), and a
    // chunk is two tasks of the complete algebra -- one per warp, one pass instead of two (100K x 512 / 1024 points -23 % /
    // -26 %; 16 chunks x 2 = one 32-warp CTA per SM instead of two 16-warp ones: +10 %, so not there).
    // CLOSYR_WARPS_PER_CHUNK=1 switches it off (measurement knob).
    static const int per_chunk = [] { const char *e = getenv("CLOSYR_WARPS_PER_CHUNK"); const int v = e ? atoi(e) : 2; return v == 1 ? 1 : 2; }();
    const int want = chunks_per_tile <= 8 ? chunks_per_tile * per_chunk : chunks_per_tile;
    int w = 1;
    while (w < cap && w < want) w <<= 1;
    return w;
}

}  // namespace closyr
