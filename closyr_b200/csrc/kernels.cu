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
    const double *sconst;
    double *slvl;              // per-chunk sums of the tile (score_kernel) / scratch
    uint32_t tcode_addr;       // shared-window address of the threaded code (vm.cuh)
    uint32_t slot_addr;        // shared-window address of this lane's column of the warp's spill slots
    uint32_t pflags;           // PF_* properties of the program
    const double *leaf_tab;    // f(x_i) tables of the resident points (ScoreParams), nullptr when xs are ad hoc
};

constexpr int SLOT_DOUBLES_PER_WARP = CLOSYR_REG_SLOTS * POINTS_PER_LANE * 32;

// shared memory of a CTA: [warp slots | constants | `nlvl` doubles of scratch | threaded code (ncode + 1 entries) |
// the program as the caller wrote it (ncode words)]
struct SmemLayout {
    double *sconst, *slvl;
    TInsn *tcode;
    uint32_t *raw;
};
__device__ __forceinline__ SmemLayout smem_layout(unsigned char *smem, int nthreads, uint32_t nk, int nlvl, uint32_t ncode)
{
    SmemLayout L;
    L.sconst = reinterpret_cast<double *>(smem) + (nthreads >> 5) * SLOT_DOUBLES_PER_WARP;
    L.slvl = L.sconst + ((nk + 1) & ~1u);
    L.tcode = reinterpret_cast<TInsn *>(L.slvl + ((nlvl + 1) & ~1));      // 16-byte aligned
    L.raw = reinterpret_cast<uint32_t *>(L.tcode + ncode + 1);
    return L;
}

// translates the staged program into threaded code (one entry per word, one word per thread and pass) and collects
// the program's properties; every thread of the CTA calls it after the program and its constants are visible
__device__ __forceinline__ Staged finish_stage(const SmemLayout &L, uint32_t ncode, unsigned char *smem, int tid, int nthreads,
                                               const double *leaf_tab, int64_t leaf_stride, uint32_t *s_pflags)
{
    uint32_t pf = 0;
    for (uint32_t i = tid; i <= ncode; i += nthreads) {      // entry `ncode` is the slack the prefetch reads
        TInsn t;
        pf |= decode_insn(i < ncode ? L.raw[i] : 0u, L.sconst, leaf_tab != nullptr, leaf_stride, t);
        L.tcode[i] = t;
    }
    if (pf) atomicOr(s_pflags, pf);
    __syncthreads();
    Staged st;
    st.sconst = L.sconst;
    st.slvl = L.slvl;
    st.tcode_addr = (uint32_t)__cvta_generic_to_shared(L.tcode);
    st.slot_addr = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<double *>(smem) + (tid >> 5) * SLOT_DOUBLES_PER_WARP + (tid & 31));
    st.pflags = *s_pflags;
    st.leaf_tab = leaf_tab;
    return st;
}

// copies one program into shared memory with plain loads (eval_kernel: one warp per expression)
__device__ __forceinline__ Staged stage_program(const DevPopulation &pop, int64_t e, unsigned char *smem, int tid,
                                                int nthreads, int nlvl, uint32_t *s_pflags)
{
    const uint32_t c0 = pop.code_off[e], c1 = pop.code_off[e + 1];
    const uint32_t k0 = pop.const_off[e], k1 = pop.const_off[e + 1];
    const uint32_t ncode = c1 - c0, nk = k1 - k0;
    const SmemLayout L = smem_layout(smem, nthreads, nk, nlvl, ncode);
    const double *gk = pop.consts + (k0 - pop.const_base);
    const uint32_t *gc = pop.code + (c0 - pop.code_base);
    for (uint32_t i = tid; i < nk; i += nthreads) L.sconst[i] = gk[i];
    for (uint32_t i = tid; i < ncode; i += nthreads) L.raw[i] = gc[i];
    if (tid == 0) *s_pflags = 0;
    __syncthreads();
    return finish_stage(L, ncode, smem, tid, nthreads, nullptr, 0, s_pflags);
}

// One round of a CTA = one 128-point chunk per warp: v[j] = f(x[j]) for the lane's 4 points (+Inf where the value is
// not a real number, as eval.clj:70-80 has it), nr = mask of those points.
//
// Every warp first runs the IEEE-real fast path on its chunk.  The points it reports are compacted across
// the WHOLE CTA into one list and re-evaluated with the complete algebra, 64 points (2 per lane) per warp and
// pass -- so the extra work of a chunk that left the real line is shared by all warps instead of stalling the
// CTA behind one of them.  `mode` is the CTA's adaptive state: when most points of the previous round needed
// the complete algebra the fast pass is skipped.  Both evaluators give bit-identical values wherever the fast
// path is valid, so the heuristic cannot change a score.
// Must be called by every thread of the CTA (it contains __syncthreads).
struct Round {
    double *list;          // [warps * 128] indices of the reported points, replaced in place by their results
    uint32_t *counts;      // [2][32] reported points per warp, double-buffered by round parity
};
constexpr int CPLX_POINTS = 2;       // points per lane of the complete-algebra interpreter

struct CplxPassResult { double v[4]; uint32_t nr, mode; };

// The cold half of a round: some point of the CTA was reported.  A function of its own so that the state it needs
// does not live in the registers of the common path.
__device__ __noinline__ CplxPassResult cplx_pass(const double *__restrict__ sconst, const uint32_t tcode_addr,
                                                 const uint32_t slot_addr, const double *__restrict__ leaf_tab, const bool deep,
                                                 double *__restrict__ list, uint32_t *__restrict__ cnt,
                                                 const double *__restrict__ xs, const int64_t base, const uint32_t ev,
                                                 const double v0, const double v1, const double v2, const double v3)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    CplxPassResult res;
    res.v[0] = v0; res.v[1] = v1; res.v[2] = v2; res.v[3] = v3;
    res.nr = 0;
    // CTA-wide compaction of the reported points
    const uint32_t mine = __popc(ev);
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
    const uint32_t total = __shfl_sync(FULL, ci, 31);
    uint32_t pos = __shfl_sync(FULL, ci - cw, warp) + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if ((ev >> j) & 1u) list[pos++] = __longlong_as_double(base + j);      // the point's index
    __syncthreads();
    constexpr uint32_t TASK = 32 * CPLX_POINTS;
    for (uint32_t t0 = warp * TASK; t0 < total; t0 += nwarps * TASK) {
        CplxArgs<CPLX_POINTS> ca;
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            ca.idx[j] = __double_as_longlong(list[r < total ? r : t0]);          // idle lanes repeat the task's first point
        }
        const CplxResult<CPLX_POINTS> cr = deep ? run_cplx<CPLX_POINTS, true>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab)
                                                : run_cplx<CPLX_POINTS, false>(tcode_addr, slot_addr, sconst, xs, ca, leaf_tab);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CPLX_POINTS; j++) {
            const uint32_t r = t0 + j * 32 + lane;
            if (r < total) list[r] = ((cr.nonreal >> j) & 1u) ? __longlong_as_double(0x7ff8000000000000ll) : cr.out[j];   // NaN marks "not a real number"
        }
    }
    __syncthreads();
    pos -= mine;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if ((ev >> j) & 1u) {
            const double r = list[pos++];
            if (r != r) { res.v[j] = __longlong_as_double(0x7ff0000000000000ll); res.nr |= 1u << j; }
            else res.v[j] = r;
        }
    }
    res.mode = (total * 4u >= (uint32_t)blockDim.x * 4u * 3u) ? 1u : 0u;      // >= 75 % of the round's points
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
                                                   const int64_t base, const uint32_t valid, const uint32_t ev_in,
                                                   const uint32_t quirk_off_plus1, const uint32_t sum_addr,
                                                   const double v0, const double v1, const double v2, const double v3)
{
    RoundSlowResult out;
    uint32_t ev = ev_in;
    double v[4] = { v0, v1, v2, v3 };
    if (quirk_off_plus1) {
        // eval.clj:109: constant AST whose argument count equals the point count is decoded per argument
        ev = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = ((valid >> j) & 1u) ? sconst[quirk_off_plus1 - 1 + base + j] : 0.0;
    }
    const CplxPassResult cp = cplx_pass(sconst, tcode_addr, slot_addr, leaf_tab, deep, list, cnt, xs, base, ev, v[0], v[1], v[2], v[3]);
    out.nr = cp.nr;
    out.mode = cp.mode;
    out.used = __any_sync(FULL, ev != 0) ? 1u : 0u;
    out.mx = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) out.v[j] = cp.v[j];
    if (__any_sync(FULL, valid != 0)) out.mx = chunk_residuals(cp.v, ys + base, valid, sum_addr);      // not for an idle warp
    return out;
}

// ---- TMA bulk copies (cp.async.bulk + mbarrier): the population buffer is laid out for it (every program and
// every constant pool starts on a 16-byte boundary and is a multiple of 16 bytes long)
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

// stages one program with two TMA bulk copies and translates it; every thread of the CTA calls it, thread 0 issues
// the copies and everybody waits on the mbarrier (single use: parity 0)
__device__ __forceinline__ Staged stage_program_tma(const DevPopulation &pop, int64_t e, unsigned char *smem, int tid,
                                                    int nthreads, int nlvl, unsigned long long *bar, const double *leaf_tab,
                                                    int64_t leaf_stride, uint32_t *s_pflags)
{
    const uint32_t c0 = pop.code_off[e], c1 = pop.code_off[e + 1];
    const uint32_t k0 = pop.const_off[e], k1 = pop.const_off[e + 1];
    const uint32_t ncode = c1 - c0, nk = k1 - k0;
    const SmemLayout L = smem_layout(smem, nthreads, nk, nlvl, ncode);
    if (tid == 0) {
        *s_pflags = 0;
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, ncode * 4u + nk * 8u);
        if (nk) bulk_g2s(L.sconst, pop.consts + (k0 - pop.const_base), nk * 8u, bar);
        bulk_g2s(L.raw, pop.code + (c0 - pop.code_base), ncode * 4u, bar);
    }
    __syncthreads();                 // the initialised barrier is visible to every waiter
    mbar_wait(bar, 0);
    return finish_stage(L, ncode, smem, tid, nthreads, leaf_tab, leaf_stride, s_pflags);
}

__global__ void __launch_bounds__(MAX_SCORE_WARPS * 32, 1) score_kernel(const ScoreParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double s_max[MAX_SCORE_WARPS];
    __shared__ uint32_t s_nonreal[MAX_SCORE_WARPS], s_full[MAX_SCORE_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int64_t item = blockIdx.x;
    const int64_t e = item / p.n_tiles;
    const int tile = (int)(item - e * p.n_tiles);
    const uint32_t meta = p.pop.meta[e];
    const int leafs = (int)(meta & 0xffffu);
    const uint32_t flags = (meta >> 16) & 0xffu;

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

    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_pflags;
    const Staged st = stage_program_tma(p.pop, e, smem, tid, blockDim.x, p.chunks_per_tile + (blockDim.x >> 5) * CHUNK_POINTS,
                                        &s_bar, p.leaf_tab, p.leaf_stride, &s_pflags);
    const bool is_const = (flags & CL_F_CONST) != 0;
    const uint32_t aux = p.pop.aux[e];
    const bool quirk = is_const && (aux >> 16) != 0 && (int64_t)(aux >> 16) == p.n;
    const uint32_t quirk_off = aux & 0xffffu;

    const int64_t c_lo = p.chunk_begin + (int64_t)tile * p.chunks_per_tile;
    int64_t c_hi = c_lo + p.chunks_per_tile;
    if (c_hi > p.chunk_end) c_hi = p.chunk_end;
    const int cnt = (int)(c_hi - c_lo);

    __shared__ uint32_t s_counts[64];
    double *const list = st.slvl + p.chunks_per_tile;
    const uint32_t slvl_addr = smem_u32(st.slvl);
    const bool fast_ok = !is_const && !(st.pflags & PF_NO_FAST);
    double warp_mx = 0.0;
    uint32_t nonreal = 0, mode = 0, parity = 0, used_full = 0;
    for (int64_t c0 = c_lo; c0 < c_hi; c0 += nwarps, parity ^= 1u) {
        const int64_t c = c0 + warp;
        const int64_t base = c * CHUNK_POINTS + lane * POINTS_PER_LANE;
        // xs, ys and every table row are padded to whole chunks (x = 1.0, y = 0.0), so a lane always loads its 4 points
        const int64_t left = c < c_hi ? p.n - base : 0;
        const uint32_t valid = left >= 4 ? 0xfu : (left <= 0 ? 0u : (1u << (int)left) - 1u);
        const uint32_t sum_addr = slvl_addr + (uint32_t)(c - c_lo) * 8u;
        const bool fast = fast_ok && mode == 0;
        RealResult<4> rr;
        rr.ev = valid;
        rr.mx = 0.0;
        rr.v[0] = rr.v[1] = rr.v[2] = rr.v[3] = 0.0;
        if (fast && c < c_hi)
            rr = run_real<4>(st.tcode_addr, st.slot_addr, p.xs + base, p.leaf_tab ? p.leaf_tab + base : nullptr, p.ys + base, valid, sum_addr);
        // nothing reported anywhere in the CTA (the common round): one barrier, the fast pass has done everything
        if (fast && !__syncthreads_or(rr.ev != 0)) {
            warp_mx = max_nonneg(warp_mx, rr.mx);
            continue;
        }
        const RoundSlowResult rs = round_slow(st.sconst, st.tcode_addr, st.slot_addr, p.leaf_tab, (st.pflags & PF_DEEP) != 0, list, s_counts + parity * 32, p.xs, p.ys,
                                              base, valid, rr.ev, quirk ? quirk_off + 1u : 0u, sum_addr, rr.v[0], rr.v[1], rr.v[2], rr.v[3]);
        warp_mx = max_nonneg(warp_mx, rs.mx);
        nonreal += __popc(rs.nr);
        used_full |= rs.used;
        mode = rs.mode;
    }
    warp_mx = warp_max_nonneg(warp_mx);
    nonreal = __reduce_add_sync(FULL, nonreal);
    if (lane == 0) { s_max[warp] = warp_mx; s_nonreal[warp] = nonreal; s_full[warp] = used_full; }
    __syncthreads();
    if (warp != 0) return;

    // index-aligned pairwise tree over the chunk sums (a missing right child leaves the left one unchanged)
    for (int stride = 1; stride < cnt; stride <<= 1) {
        for (int i = lane * 2 * stride; i + stride < cnt; i += 64 * stride) st.slvl[i] = st.slvl[i] + st.slvl[i + stride];
        __syncwarp();
    }
    // max, counts and flags of the warps: one lane per warp, then shuffles (all three are order-independent)
    const double tile_max = warp_max_nonneg(lane < nwarps ? s_max[lane] : 0.0);
    const uint32_t tile_nonreal = __reduce_add_sync(FULL, lane < nwarps ? s_nonreal[lane] : 0u);
    const uint32_t tile_full = __reduce_or_sync(FULL, lane < nwarps ? s_full[lane] : 0u);
    if (lane != 0) return;
    const double tile_sum = cnt > 0 ? st.slvl[0] : 0.0;
    if (p.final_scores) {
        uint8_t rf = tile_full ? CL_R_FULL_SEMANTICS : 0;
        double score;
        if (is_const && !quirk && tile_nonreal != 0) {                 // eval.clj:83-96 throws => min-score
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
        const double *in = p.partials + ((int64_t)t * p.n_expr + e) * 3;
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
    const uint32_t meta = p.pop.meta[e];
    const uint32_t flags = (meta >> 16) & 0xffu;
    if (flags & (CL_F_UNSUPPORTED | CL_F_NIL)) {
        if (lane == 0) p.status[e] = 1;
        return;
    }
    __shared__ uint32_t s_pflags;
    const Staged st = stage_program(p.pop, e, smem, lane, 32, CHUNK_POINTS, &s_pflags);
    const bool is_const = (flags & CL_F_CONST) != 0;
    const uint32_t aux = p.pop.aux[e];
    const bool quirk = is_const && (aux >> 16) != 0 && (int64_t)(aux >> 16) == p.n;
    const uint32_t quirk_off = aux & 0xffffu;
    const int64_t nchunks = (p.n + CHUNK_POINTS - 1) / CHUNK_POINTS;
    __shared__ uint32_t s_counts[64];
    __shared__ double s_sum;                                   // the chunk sums are not used here
    uint32_t nonreal = 0;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t base = c * CHUNK_POINTS + lane * POINTS_PER_LANE;
        const int64_t left = p.n - base;
        const uint32_t valid = left >= 4 ? 0xfu : (left <= 0 ? 0u : (1u << (int)left) - 1u);
        // every point through the complete algebra (one warp per expression: there is no load to balance, and the
        // report path evaluates few points); xs doubles as the y vector of the unused residuals
        const RoundSlowResult rs = round_slow(st.sconst, st.tcode_addr, st.slot_addr, nullptr, (st.pflags & PF_DEEP) != 0, st.slvl, s_counts + (c & 1) * 32, p.xs, p.xs,
                                              base, valid, valid, quirk ? quirk_off + 1u : 0u, smem_u32(&s_sum), 0.0, 0.0, 0.0, 0.0);
        nonreal += __popc(rs.nr);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((valid >> j) & 1u) p.out_f[e * p.n + base + j] = ((rs.nr >> j) & 1u) ? inf : rs.v[j];
    }
    nonreal = __reduce_add_sync(FULL, nonreal);
    if (lane == 0) p.status[e] = (is_const && !quirk && nonreal != 0) ? 2 : 0;
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

int score_warps(int32_t chunks_per_tile);

void launch_leaf_table(const double *xs, int64_t n, int64_t stride, double *tab, void *stream)
{
    if (stride <= 0) return;
    const int threads = 128;
    const int64_t quads = stride / 4;
    leaf_table_kernel<<<(unsigned)((quads + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(xs, n, stride, tab);
}

void launch_score(const ScoreParams &p, size_t smem_bytes, void *stream)
{
    const int64_t items = p.pop.n_expr * p.n_tiles;
    if (items <= 0) return;
    const int warps = score_warps(p.chunks_per_tile);
    // opt in to > 48 KiB of dynamic shared memory; the attribute is per device, so it is set at every launch
    if (smem_bytes > 48 * 1024)
        cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    score_kernel<<<(unsigned)items, warps * 32, smem_bytes, (cudaStream_t)stream>>>(p);
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

size_t score_smem_bytes(uint32_t max_code_words, uint32_t max_consts, int32_t chunks_per_tile, int warps)
{
    // [warp slots | constants | chunk sums + the compaction list | threaded code | the program]  (smem_layout)
    return (size_t)warps * SLOT_DOUBLES_PER_WARP * 8 + (size_t)((max_consts + 1) & ~1u) * 8 +
           (size_t)((chunks_per_tile + warps * CHUNK_POINTS + 1) & ~1) * 8 +
           ((size_t)max_code_words + 1) * sizeof(dev::TInsn) + (size_t)max_code_words * 4 + 16;
}
int score_warps(int32_t chunks_per_tile)
{
    // CLOSYR_SCORE_WARPS (a power of two <= 32) caps the CTA size: a tuning knob for measurements, not an API
    static const int cap = [] {
        const char *e = getenv("CLOSYR_SCORE_WARPS");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= MAX_SCORE_WARPS && (v & (v - 1)) == 0) ? v : MAX_SCORE_WARPS;
    }();
    int w = 1;
    while (w < cap && w < chunks_per_tile) w <<= 1;
    return w;
}

}  // namespace closyr
