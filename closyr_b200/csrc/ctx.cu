// ctx.cu -- the C-ABI shim (north star (2)): owns CUDA streams, device buffers and pinned staging for the
// 1..8 GPUs of one box, keeps xs/ys resident across generations, shards a population across the devices.
// Replaces the per-run marshalling of the reference (src/closyr/symbolic_regression.clj:355-383,
// src/closyr/ops/common.clj:30-76) and the per-phenotype call of ops/score-fn (src/closyr/ops.clj:119-136).
// There is deliberately no CPU path in this file: without a CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/closyr_b200.h"
#include "internal.hpp"
#include "kernels.cuh"

using namespace closyr;

namespace closyr {

size_t rank_scratch_bytes(int64_t n);
int launch_rank(const double *scores, int64_t n, int64_t k, void *scratch, int64_t *out, void *stream);

void *pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pinned_free(void *p) { cudaFreeHost(p); }

namespace {

// NCCL is bound at run time (dlopen): the library loads on a box without it, and inside a process that already carries
// an NCCL (PyTorch) the same copy is used.  Only the entry points of the point-sharding collective are needed.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.handle, "ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.handle, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.handle, "ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

struct DevBuf {                       // grow-only device allocation
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DevPop {                       // one resident population shard
    DevBuf code, code_off, consts, const_off, meta, aux;
    DevBuf tcode, desc;               // what prepare_kernel derives from the arrays above (threaded code, descriptors)
    DevBuf key, order;                // grouping keys and the expression order made from them (order_kernel)
    bool ordered = false;             // `order` holds the order of the current threaded code (made on first use)
    int64_t prepared_stride = -2;     // leaf_stride the threaded code was made for (-1: without tables, -2: not made yet)
    int64_t lo = 0, hi = 0;           // expression range of the caller's population held here
    uint32_t code_base = 0, const_base = 0, max_code = 0, max_consts = 0, n_words = 0;
    DevPopulation view() const
    {
        DevPopulation v;
        v.code = (const uint32_t *)code.p;
        v.code_off = (const uint32_t *)code_off.p;
        v.consts = (const double *)consts.p;
        v.const_off = (const uint32_t *)const_off.p;
        v.meta = (const uint32_t *)meta.p;
        v.aux = (const uint32_t *)aux.p;
        v.code_base = code_base;
        v.const_base = const_base;
        v.n_expr = hi - lo;
        v.tcode = (const uint4 *)tcode.p;
        v.desc = (const ExprDesc *)desc.p;
        v.order = nullptr;            // launch_range decides
        return v;
    }
    void release() { code.release(); code_off.release(); consts.release(); const_off.release(); meta.release(); aux.release(); tcode.release(); desc.release(); key.release(); order.release(); }
};

struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;
    DevBuf xs, ys, leaf_tab, scores, rflags, partials, partials_all, flush, eval_xs, eval_f, eval_status, rank_a, rank_b;
    int64_t leaf_stride = 0;
    DevPop pop, eval_pop;
    int64_t chunk_lo = 0, chunk_hi = 0;   // point-sharded mode: this device's chunk range
    int64_t fin_lo = 0, fin_hi = 0;       // point-sharded mode: the expressions this device finalises and returns
    ncclComm_t comm = nullptr;            // point-sharded mode: this device's communicator
};

}  // namespace
}  // namespace closyr

struct closyr_ctx {
    std::vector<DeviceState> devs;
    uint32_t flags = 0;
    int64_t n_points = -1;
    int64_t n_expr = -1;              // resident population size, -1 = none
    bool scored = false;
    int64_t launches = 0;
    // point sharding across PROCESSES (one device each): this context is rank `rank` of `n_ranks` (closyr_comm_init_rank)
    int n_ranks = 1, rank = 0;
    std::string err;
    std::mutex mu;
};

namespace {

int fail(closyr_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    set_error(msg);
    return code;
}

#define NC(ctx, call)                                                                                     \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(ctx, CLOSYR_E_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r_));       \
    } while (0)

#define CU(ctx, call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            cudaGetLastError();                                                                           \
            return fail(ctx, CLOSYR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
        }                                                                                                 \
    } while (0)

int64_t n_chunks_of(int64_t n) { return (n + CHUNK_POINTS - 1) / CHUNK_POINTS; }
int64_t pow2_ceil(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }

// copies expressions [lo, hi) of `pop` to the device
int upload_shard(closyr_ctx *ctx, DeviceState &d, DevPop &dp, const closyr_population *pop, int64_t lo, int64_t hi)
{
    CU(ctx, cudaSetDevice(d.device));
    dp.lo = lo;
    dp.hi = hi;
    dp.prepared_stride = -2;
    const int64_t ne = hi - lo;
    if (ne <= 0) return CLOSYR_OK;
    dp.code_base = pop->code_off[lo];
    dp.const_base = pop->const_off[lo];
    // the layout the kernel's TMA staging depends on (include/closyr_b200.h): every program starts on a 16-byte
    // boundary and is a whole number of 16-byte units; offsets never decrease.  A violation would be a misaligned
    // bulk copy (sticky context error) or a wrapped size, so it is refused here.
    if (pop->code_off[hi] % 4u != 0 || pop->const_off[hi] % 2u != 0) {
        dp.hi = dp.lo;
        return fail(ctx, CLOSYR_E_ARG, "population layout: code_off must be a multiple of 4 words, const_off of 2 doubles");
    }
    if (pop->code_off[hi] < dp.code_base || pop->const_off[hi] < dp.const_base) {
        dp.hi = dp.lo;
        return fail(ctx, CLOSYR_E_ARG, "population layout: offsets must not decrease");
    }
    const size_t ncode = pop->code_off[hi] - dp.code_base, nk = pop->const_off[hi] - dp.const_base;
    // the copies first: the DMA engine moves the arrays while the host walks the offsets below (0.5 ms per 100K
    // expressions).  Nothing is launched on the shard before the walk has accepted it.
    CU(ctx, dp.code.reserve(ncode * 4 + 16));
    CU(ctx, dp.consts.reserve(nk * 8 + 16));
    CU(ctx, dp.code_off.reserve((ne + 1) * 4));
    CU(ctx, dp.const_off.reserve((ne + 1) * 4));
    CU(ctx, dp.meta.reserve(ne * 4));
    CU(ctx, dp.aux.reserve(ne * 4));
    if (ncode) CU(ctx, cudaMemcpyAsync(dp.code.p, pop->code + dp.code_base, ncode * 4, cudaMemcpyHostToDevice, d.stream));
    if (nk) CU(ctx, cudaMemcpyAsync(dp.consts.p, pop->consts + dp.const_base, nk * 8, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, cudaMemcpyAsync(dp.code_off.p, pop->code_off + lo, (ne + 1) * 4, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, cudaMemcpyAsync(dp.const_off.p, pop->const_off + lo, (ne + 1) * 4, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, cudaMemcpyAsync(dp.meta.p, pop->meta + lo, ne * 4, cudaMemcpyHostToDevice, d.stream));
    CU(ctx, cudaMemcpyAsync(dp.aux.p, pop->aux + lo, ne * 4, cudaMemcpyHostToDevice, d.stream));
    // the walk over the offsets (it streams the whole code array through the host caches for the CL_END test): a few
    // threads for a large shard, each over its own stretch; the first offending expression is reported
    struct Walk { uint32_t mc = 0, mk = 0; const char *bad = nullptr; int64_t bad_e = -1; };
    const uint32_t code_end = pop->code_off[hi], const_end = pop->const_off[hi];
    auto walk = [pop, code_end, const_end](int64_t a, int64_t b, Walk *w) {
        for (int64_t e = a; e < b && !w->bad; e++) {
            const uint32_t c0 = pop->code_off[e], c1 = pop->code_off[e + 1], k0 = pop->const_off[e], k1 = pop->const_off[e + 1];
            if (c0 % 4u != 0 || k0 % 2u != 0) w->bad = "code_off must be a multiple of 4 words, const_off of 2 doubles";
            else if (c1 < c0 || k1 < k0 || c1 > code_end || k1 > const_end) w->bad = "offsets must not decrease";
            else if (c1 == c0) w->bad = "empty program; a program ends with CL_END";
            else if ((pop->code[c1 - 1] & 0xffu) != CL_END) w->bad = "program does not end with CL_END";
            if (w->bad) w->bad_e = e;
            else {
                w->mc = std::max(w->mc, c1 - c0);
                w->mk = std::max(w->mk, k1 - k0);
            }
        }
    };
    const int nt = ne >= 32768 ? (int)std::min<unsigned>(4u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    Walk parts[4];
    if (nt == 1) walk(lo, hi, &parts[0]);
    else {
        std::thread th[3];
        for (int t = 1; t < nt; t++) th[t - 1] = std::thread(walk, lo + ne * t / nt, lo + ne * (t + 1) / nt, &parts[t]);
        walk(lo, lo + ne / nt, &parts[0]);
        for (int t = 1; t < nt; t++) th[t - 1].join();
    }
    uint32_t mc = 0, mk = 0;
    const char *bad = nullptr;
    int64_t bad_e = -1;
    for (int t = 0; t < nt; t++) {
        mc = std::max(mc, parts[t].mc);
        mk = std::max(mk, parts[t].mk);
        if (!bad && parts[t].bad) { bad = parts[t].bad; bad_e = parts[t].bad_e; }
    }
    if (!bad && (mc > CLOSYR_MAX_CODE_WORDS || mk > CLOSYR_MAX_CONSTS)) bad = "a program above CLOSYR_MAX_CODE_WORDS / CLOSYR_MAX_CONSTS";
    if (bad) {
        cudaStreamSynchronize(d.stream);          // the copies read the caller's arrays
        dp.hi = dp.lo;                            // nothing resident
        return fail(ctx, CLOSYR_E_ARG, std::string("population layout: ") + bad + (bad_e >= 0 ? " (expression " + std::to_string(bad_e) + ")" : ""));
    }
    dp.max_code = mc;
    dp.max_consts = mk;
    dp.n_words = (uint32_t)ncode;
    return CLOSYR_OK;
}

// threaded code + descriptors of a resident shard (kernels.cu prepare_kernel), made once per upload and again when the
// leaf tables it refers to change.  leaf_stride < 0: no tables (ad hoc xs).
int prepare_shard(closyr_ctx *ctx, DeviceState &d, DevPop &dp, int64_t leaf_stride)
{
    const int64_t want = leaf_stride > 0 ? leaf_stride : -1;
    if (dp.prepared_stride == want || dp.hi <= dp.lo) return CLOSYR_OK;
    CU(ctx, dp.tcode.reserve((size_t)dp.n_words * 16 + 16));
    CU(ctx, dp.desc.reserve((size_t)(dp.hi - dp.lo) * sizeof(ExprDesc)));
    CU(ctx, dp.key.reserve((size_t)(dp.hi - dp.lo) * 4));
    PrepareParams pp;
    pp.key = (uint32_t *)dp.key.p;
    pp.pop = dp.view();
    pp.tcode = (uint4 *)dp.tcode.p;
    pp.desc = (ExprDesc *)dp.desc.p;
    pp.have_tab = want > 0 ? 1 : 0;
    pp.leaf_stride = want > 0 ? want : 0;
    launch_prepare(pp, d.stream);
    ctx->launches++;
    dp.ordered = false;
    CU(ctx, cudaGetLastError());
    dp.prepared_stride = want;
    return CLOSYR_OK;
}

int check_population(closyr_ctx *ctx, const closyr_population *pop)
{
    if (!pop || pop->n_expr < 0) return fail(ctx, CLOSYR_E_ARG, "bad population");
    if (pop->n_expr > 0 && (!pop->code || !pop->code_off || !pop->const_off || !pop->meta || !pop->aux))
        return fail(ctx, CLOSYR_E_ARG, "population has NULL arrays");
    if (pop->n_expr > 0 && pop->const_off[pop->n_expr] > pop->const_off[0] && !pop->consts)
        return fail(ctx, CLOSYR_E_ARG, "population has NULL constant pool");
    return CLOSYR_OK;
}

// tile size (in chunks, power of two) so that the launch has enough warps to fill the GPU
int32_t pick_chunks_per_tile(int64_t n_expr, int64_t nchunks)
{
    const int64_t target_items = 148 * 6;              // CTAs of up to 32 warps, 1 resident per SM, a few waves
    int64_t cpt = pow2_ceil(std::max<int64_t>(1, nchunks));
    while (cpt > MAX_CHUNKS_PER_TILE) cpt >>= 1;
    while (cpt > 4 && n_expr * ((nchunks + cpt - 1) / cpt) < target_items) cpt >>= 1;
    return (int32_t)cpt;
}

// scores chunks [c_lo, c_hi) of the resident shard; leaves one partial triple per expression in
// d.partials ([n_expr][3]) unless the range covers every point and fits one tile (then scores directly).
int launch_range(closyr_ctx *ctx, DeviceState &d, int32_t max_leafs, int64_t c_lo, int64_t c_hi, bool want_final,
                 double *partials_out)
{
    if (d.pop.hi <= d.pop.lo) return CLOSYR_OK;
    {
        int rc = prepare_shard(ctx, d, d.pop, d.leaf_stride);
        if (rc) return rc;
    }
    const DevPopulation pv = d.pop.view();
    const int64_t ne = pv.n_expr;
    const int64_t nch = c_hi - c_lo;
    const int32_t cpt = pick_chunks_per_tile(ne, nch);
    const int32_t n_tiles = (int32_t)std::max<int64_t>(1, (nch + cpt - 1) / cpt);
    ScoreParams sp;
    sp.pop = pv;
    sp.xs = (const double *)d.xs.p;
    sp.ys = (const double *)d.ys.p;
    sp.leaf_tab = d.leaf_stride > 0 ? (const double *)d.leaf_tab.p : nullptr;
    sp.leaf_stride = d.leaf_stride;
    sp.n = ctx->n_points;
    sp.chunk_begin = c_lo;
    sp.chunk_end = c_hi;
    sp.chunks_per_tile = cpt;
    sp.warps = score_warps(cpt);
    sp.n_tiles = n_tiles;
    {
        // grouped CTA order: pays most when several CTAs share an SM (short tiles: -20 %), still -1 % with one program per SM
        // since the programs that skip the fast pass are grouped together; CLOSYR_ORDER=0 switches it off
        static const int force = [] { const char *e = getenv("CLOSYR_ORDER"); return e ? atoi(e) : -1; }();
        if (force != 0) {
            if (!d.pop.ordered) {
                CU(ctx, d.pop.order.reserve((size_t)ne * 4));
                launch_order((const uint32_t *)d.pop.key.p, ne, (uint32_t *)d.pop.order.p, d.stream);
                ctx->launches++;
                CU(ctx, cudaGetLastError());
                d.pop.ordered = true;
            }
            sp.pop.order = (const uint32_t *)d.pop.order.p;
        }
    }
    sp.max_leafs = max_leafs;
    sp.scores = (double *)d.scores.p;
    sp.rflags = (uint8_t *)d.rflags.p;
    sp.partials = nullptr;
    sp.partial_tile_offset = 0;
    if (ne * n_tiles > 0x7fffffff) return fail(ctx, CLOSYR_E_ARG, "too many (expression, tile) items for one launch");
    const bool direct = want_final && n_tiles == 1;
    sp.final_scores = direct ? 1 : 0;
    if (!direct) {
        if (n_tiles == 1 && partials_out) sp.partials = partials_out;
        else {
            CU(ctx, d.partials.reserve((size_t)n_tiles * ne * 3 * 8));
            sp.partials = (double *)d.partials.p;
        }
    }
    ctx->launches += launch_score(sp, d.pop.max_code, d.stream);
    CU(ctx, cudaGetLastError());
    if (!direct && !(n_tiles == 1 && partials_out)) {
        FinalizeParams fp;
        fp.meta = pv.meta;
        fp.partials = (const double *)d.partials.p;
        fp.part_stride = ne;
        fp.n_expr = ne;
        fp.n_tiles_total = n_tiles;
        fp.max_leafs = max_leafs;
        fp.n = ctx->n_points;
        fp.scores = (double *)d.scores.p;
        fp.rflags = (uint8_t *)d.rflags.p;
        fp.out_partials = want_final ? nullptr : partials_out;
        launch_finalize(fp, d.stream);
        ctx->launches++;
        CU(ctx, cudaGetLastError());
    }
    return CLOSYR_OK;
}

}  // namespace

extern "C" {

const char *closyr_last_error(const closyr_ctx *ctx)
{
    if (ctx && !ctx->err.empty()) return ctx->err.c_str();
    return g_last_error.c_str();
}

int closyr_ctx_create(const int *devices, int n_devices, uint32_t flags, closyr_ctx **out)
{
    if (!out || n_devices < 1 || n_devices > 64) return fail(nullptr, CLOSYR_E_ARG, "bad arguments to closyr_ctx_create");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, CLOSYR_E_NO_GPU,
                    std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
                        "); closyr_b200 has no CPU fallback");
    }
    closyr_ctx *ctx = new closyr_ctx();
    ctx->flags = flags;
    for (int i = 0; i < n_devices; i++) {
        DeviceState d;
        d.device = devices ? devices[i] : i;
        if (d.device < 0 || d.device >= count) {
            delete ctx;
            return fail(nullptr, CLOSYR_E_ARG, "device index out of range");
        }
        ctx->devs.push_back(d);
    }
    for (auto &d : ctx->devs) {
        cudaError_t r = cudaSetDevice(d.device);
        if (r == cudaSuccess) r = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        if (r == cudaSuccess) r = cudaEventCreate(&d.ev0);
        if (r == cudaSuccess) r = cudaEventCreate(&d.ev1);
        if (r == cudaSuccess) r = cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming);
        if (r != cudaSuccess) {
            std::string m = std::string("CUDA initialisation failed: ") + cudaGetErrorString(r);
            closyr_ctx_destroy(ctx);
            return fail(nullptr, CLOSYR_E_CUDA, m);
        }
    }
    if (ctx->devs.size() > 1 && (flags & CLOSYR_SHARD_POINTS)) {
        // the partial residual sums are exchanged with an NCCL all-gather over NVLink: one communicator per device
        if (!nccl().ok) {
            closyr_ctx_destroy(ctx);
            return fail(nullptr, CLOSYR_E_CUDA, "point sharding needs NCCL (libnccl.so.2 not found)");
        }
        std::vector<int> devlist;
        for (auto &d : ctx->devs) devlist.push_back(d.device);
        std::vector<ncclComm_t> comms(devlist.size(), nullptr);
        ncclResult_t r = nccl().CommInitAll(comms.data(), (int)devlist.size(), devlist.data());
        if (r != ncclSuccess) {
            std::string m = std::string("ncclCommInitAll: ") + nccl().GetErrorString(r);
            closyr_ctx_destroy(ctx);
            return fail(nullptr, CLOSYR_E_CUDA, m);
        }
        for (size_t i = 0; i < ctx->devs.size(); i++) ctx->devs[i].comm = comms[i];
    }
    if (ctx->devs.size() > 1) {
        for (size_t i = 1; i < ctx->devs.size(); i++) {       // closyr_rank_resident gathers the scores on device 0
            cudaSetDevice(ctx->devs[0].device);
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->devs[0].device, ctx->devs[i].device);
            if (can) { cudaDeviceEnablePeerAccess(ctx->devs[i].device, 0); cudaGetLastError(); }
        }
    }
    *out = ctx;
    return CLOSYR_OK;
}

void closyr_ctx_destroy(closyr_ctx *ctx)
{
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.comm) nccl().CommDestroy(d.comm);
        for (DevBuf *b : { &d.xs, &d.ys, &d.leaf_tab, &d.scores, &d.rflags, &d.partials, &d.partials_all, &d.flush, &d.eval_xs,
                           &d.eval_f, &d.eval_status, &d.rank_a, &d.rank_b })
            b->release();
        d.pop.release();
        d.eval_pop.release();
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.done) cudaEventDestroy(d.done);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

int closyr_ctx_device_count(const closyr_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int closyr_set_points(closyr_ctx *ctx, const double *xs, const double *ys, int64_t n)
{
    if (!ctx || n < 0 || (n > 0 && (!xs || !ys))) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_set_points");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    const std::vector<double> ones(CHUNK_POINTS, 1.0);      // alive until the streams are synchronised below
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        // padded to whole 128-point chunks (x = 1.0, y = 0.0) so that a lane always loads its 4 points
        const int64_t padded = n_chunks_of(n) * CHUNK_POINTS;
        CU(ctx, d.xs.reserve((size_t)padded * 8 + 64));
        CU(ctx, d.ys.reserve((size_t)padded * 8 + 64));
        if (n) {
            CU(ctx, cudaMemcpyAsync(d.xs.p, xs, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
            CU(ctx, cudaMemcpyAsync(d.ys.p, ys, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
            if (padded > n) {
                CU(ctx, cudaMemcpyAsync((double *)d.xs.p + n, ones.data(), (size_t)(padded - n) * 8, cudaMemcpyHostToDevice, d.stream));
                CU(ctx, cudaMemsetAsync((double *)d.ys.p + n, 0, (size_t)(padded - n) * 8, d.stream));
            }
        }
        // f(x_i) tables of the raw variable, shared by every expression of every generation scored on these points
        d.leaf_stride = padded;
        if (d.leaf_stride > 0) {
            CU(ctx, d.leaf_tab.reserve((size_t)d.leaf_stride * LEAF_TAB_ROWS * 8));
            launch_leaf_table((const double *)d.xs.p, n, d.leaf_stride, (double *)d.leaf_tab.p, d.stream);
            ctx->launches++;
            CU(ctx, cudaGetLastError());
        }
    }
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaStreamSynchronize(d.stream));
    }
    ctx->n_points = n;
    return CLOSYR_OK;
}

int closyr_upload_population(closyr_ctx *ctx, const closyr_population *pop)
{
    if (!ctx) return CLOSYR_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    int rc = check_population(ctx, pop);
    if (rc) return rc;
    const int D = (int)ctx->devs.size();
    const int64_t ne = pop->n_expr;
    std::vector<int64_t> cut((size_t)D + 1, 0);
    cut[(size_t)D] = ne;
    if (ctx->flags & CLOSYR_SHARD_POINTS) {
        for (int i = 1; i < D; i++) cut[(size_t)i] = 0;     // every device holds the whole population
    } else if (ne > 0) {
        // contiguous blocks balanced by program length (a proxy of the work per expression)
        const uint64_t total = (uint64_t)(pop->code_off[ne] - pop->code_off[0]);
        int64_t e = 0;
        for (int i = 1; i < D; i++) {
            const uint64_t want = total * (uint64_t)i / (uint64_t)D;
            while (e < ne && (uint64_t)(pop->code_off[e] - pop->code_off[0]) < want) e++;
            cut[(size_t)i] = e;
        }
    }
    for (int i = 0; i < D; i++) {
        DeviceState &d = ctx->devs[(size_t)i];
        const bool all = (ctx->flags & CLOSYR_SHARD_POINTS) != 0;
        rc = upload_shard(ctx, d, d.pop, pop, all ? 0 : cut[(size_t)i], all ? ne : cut[(size_t)i + 1]);
        if (rc) return rc;
        const int64_t n_here = d.pop.hi - d.pop.lo;
        CU(ctx, d.scores.reserve((size_t)std::max<int64_t>(1, n_here) * 8));
        CU(ctx, d.rflags.reserve((size_t)std::max<int64_t>(1, n_here)));
    }
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaStreamSynchronize(d.stream));   // the caller's buffers may be reused after return
    }
    ctx->n_expr = ne;
    ctx->scored = false;
    return CLOSYR_OK;
}

int closyr_score_resident(closyr_ctx *ctx, int32_t max_leafs, float *device_ms)
{
    if (!ctx) return CLOSYR_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (ctx->n_points < 0) return fail(ctx, CLOSYR_E_STATE, "closyr_set_points has not been called");
    if (ctx->n_expr < 0) return fail(ctx, CLOSYR_E_STATE, "no resident population");
    const int D = (int)ctx->devs.size();
    const int64_t nch = n_chunks_of(ctx->n_points);
    const bool by_points = (ctx->flags & CLOSYR_SHARD_POINTS) && (D > 1 || ctx->n_ranks > 1);
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaEventRecord(d.ev0, d.stream));
    }
    if (!by_points) {
        for (auto &d : ctx->devs) {
            CU(ctx, cudaSetDevice(d.device));
            int rc = launch_range(ctx, d, max_leafs, 0, nch, true, nullptr);
            if (rc) return rc;
        }
    } else {
        // Point sharding: part q of P evaluates all expressions on an aligned power-of-two span of chunks and leaves one
        // (sum, max, non-real) triple per expression in row q of the [P][n_expr][3] table; ONE in-place NCCL all-gather
        // completes the table on every device; every device then combines the rows with the fixed tree for its share
        // of the expressions (in one process: a slice each; across processes: all of them, every rank returns all scores).
        const int P = D > 1 ? D : ctx->n_ranks;
        const int64_t span = pow2_ceil((nch + P - 1) / P);
        const int64_t ne = ctx->n_expr;
        int parts = 0;
        for (int q = 0; q < P; q++)
            if (std::min(nch, (int64_t)q * span) < std::min(nch, (int64_t)q * span + span)) parts = q + 1;
        for (int i = 0; i < D; i++) {
            DeviceState &d = ctx->devs[(size_t)i];
            const int q = D > 1 ? i : ctx->rank;
            const int64_t lo = std::min(nch, (int64_t)q * span), hi = std::min(nch, lo + span);
            d.chunk_lo = lo;
            d.chunk_hi = hi;
            d.fin_lo = D > 1 ? ne * i / D : 0;
            d.fin_hi = D > 1 ? ne * (i + 1) / D : ne;
            CU(ctx, cudaSetDevice(d.device));
            CU(ctx, d.partials_all.reserve((size_t)P * std::max<int64_t>(1, ne) * 3 * 8));
            double *mine = (double *)d.partials_all.p + (size_t)q * ne * 3;
            if (hi > lo) {
                int rc = launch_range(ctx, d, max_leafs, lo, hi, false, mine);
                if (rc) return rc;
            } else if (ne > 0) CU(ctx, cudaMemsetAsync(mine, 0, (size_t)ne * 3 * 8, d.stream));
        }
        if (ne > 0) {
            NC(ctx, nccl().GroupStart());
            for (int i = 0; i < D; i++) {
                DeviceState &d = ctx->devs[(size_t)i];
                const int q = D > 1 ? i : ctx->rank;
                NC(ctx, nccl().AllGather((const double *)d.partials_all.p + (size_t)q * ne * 3, d.partials_all.p, (size_t)ne * 3, ncclDouble, d.comm, d.stream));
            }
            NC(ctx, nccl().GroupEnd());
            for (int i = 0; i < D; i++) {
                DeviceState &d = ctx->devs[(size_t)i];
                if (d.fin_hi <= d.fin_lo) continue;
                CU(ctx, cudaSetDevice(d.device));
                FinalizeParams fp;
                fp.meta = (const uint32_t *)d.pop.meta.p + d.fin_lo;
                fp.partials = (const double *)d.partials_all.p + (size_t)d.fin_lo * 3;
                fp.part_stride = ne;
                fp.n_expr = d.fin_hi - d.fin_lo;
                fp.n_tiles_total = std::max(1, parts);
                fp.max_leafs = max_leafs;
                fp.n = ctx->n_points;
                fp.scores = (double *)d.scores.p + d.fin_lo;
                fp.rflags = (uint8_t *)d.rflags.p + d.fin_lo;
                fp.out_partials = nullptr;
                launch_finalize(fp, d.stream);
                ctx->launches++;
                CU(ctx, cudaGetLastError());
            }
        }
    }
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaEventRecord(d.ev1, d.stream));
    }
    float worst = 0.f;
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        worst = std::max(worst, ms);
    }
    if (device_ms) *device_ms = worst;
    ctx->scored = true;
    return CLOSYR_OK;
}

int closyr_comm_unique_id(uint8_t *out_id)
{
    if (!out_id) return fail(nullptr, CLOSYR_E_ARG, "bad arguments to closyr_comm_unique_id");
    if (!nccl().ok) return fail(nullptr, CLOSYR_E_CUDA, "NCCL (libnccl.so.2) not found");
    static_assert(sizeof(ncclUniqueId) == CLOSYR_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, CLOSYR_E_CUDA, std::string("ncclGetUniqueId: ") + nccl().GetErrorString(r));
    memcpy(out_id, &id, sizeof(id));
    return CLOSYR_OK;
}

int closyr_comm_init_rank(closyr_ctx *ctx, int n_ranks, int rank, const uint8_t *id_bytes)
{
    if (!ctx || !id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_comm_init_rank");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (ctx->devs.size() != 1 || !(ctx->flags & CLOSYR_SHARD_POINTS))
        return fail(ctx, CLOSYR_E_STATE, "closyr_comm_init_rank: a context with ONE device created with CLOSYR_SHARD_POINTS");
    if (!nccl().ok) return fail(ctx, CLOSYR_E_CUDA, "NCCL (libnccl.so.2) not found");
    DeviceState &d = ctx->devs[0];
    if (d.comm) return fail(ctx, CLOSYR_E_STATE, "closyr_comm_init_rank: already initialised");
    CU(ctx, cudaSetDevice(d.device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    NC(ctx, nccl().CommInitRank(&d.comm, n_ranks, id, rank));
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return CLOSYR_OK;
}

int closyr_download_scores(closyr_ctx *ctx, double *out_scores, uint8_t *out_flags)
{
    if (!ctx || !out_scores) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_download_scores");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (!ctx->scored) return fail(ctx, CLOSYR_E_STATE, "no scores: call closyr_score_resident first");
    const bool by_points = (ctx->flags & CLOSYR_SHARD_POINTS) && (ctx->devs.size() > 1 || ctx->n_ranks > 1);
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        DeviceState &d = ctx->devs[i];
        // what this device returns: its block of the population, or (point sharding) the slice it finalised
        const int64_t lo = by_points ? d.fin_lo : d.pop.lo, off = by_points ? d.fin_lo : 0;
        const int64_t ne = by_points ? d.fin_hi - d.fin_lo : d.pop.hi - d.pop.lo;
        if (ne <= 0) continue;
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaMemcpyAsync(out_scores + lo, (const double *)d.scores.p + off, (size_t)ne * 8, cudaMemcpyDeviceToHost, d.stream));
        if (out_flags) CU(ctx, cudaMemcpyAsync(out_flags + lo, (const uint8_t *)d.rflags.p + off, (size_t)ne, cudaMemcpyDeviceToHost, d.stream));
    }
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaStreamSynchronize(d.stream));
    }
    return CLOSYR_OK;
}

int closyr_score_batch(closyr_ctx *ctx, const closyr_population *pop, int32_t max_leafs, double *out_scores,
                       uint8_t *out_flags)
{
    int rc = closyr_upload_population(ctx, pop);
    if (rc) return rc;
    rc = closyr_score_resident(ctx, max_leafs, nullptr);
    if (rc) return rc;
    return closyr_download_scores(ctx, out_scores, out_flags);
}

int closyr_eval_batch(closyr_ctx *ctx, const closyr_population *pop, const double *xs, int64_t n, double *out_f,
                      uint8_t *out_status)
{
    if (!ctx || n < 0 || (n > 0 && (!xs || !out_f)) || !out_status) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_eval_batch");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    int rc = check_population(ctx, pop);
    if (rc) return rc;
    DeviceState &d = ctx->devs[0];
    const int64_t ne = pop->n_expr;
    if (ne == 0) return CLOSYR_OK;
    rc = upload_shard(ctx, d, d.eval_pop, pop, 0, ne);
    if (rc) return rc;
    const int64_t padded = n_chunks_of(n) * CHUNK_POINTS;      // whole chunks, x = 1.0 beyond n (as closyr_set_points)
    const std::vector<double> ones(CHUNK_POINTS, 1.0);
    CU(ctx, d.eval_xs.reserve((size_t)padded * 8 + 64));
    CU(ctx, d.eval_f.reserve((size_t)std::max<int64_t>(1, ne * n) * 8));
    CU(ctx, d.eval_status.reserve((size_t)ne));
    if (n) CU(ctx, cudaMemcpyAsync(d.eval_xs.p, xs, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
    if (padded > n) CU(ctx, cudaMemcpyAsync((double *)d.eval_xs.p + n, ones.data(), (size_t)(padded - n) * 8, cudaMemcpyHostToDevice, d.stream));
    rc = prepare_shard(ctx, d, d.eval_pop, -1);        // ad hoc xs: no leaf tables
    if (rc) return rc;
    EvalParams ep;
    ep.pop = d.eval_pop.view();
    ep.xs = (const double *)d.eval_xs.p;
    ep.n = n;
    ep.out_f = (double *)d.eval_f.p;
    ep.status = (uint8_t *)d.eval_status.p;
    launch_eval(ep, eval_smem_bytes(d.eval_pop.max_code), d.stream);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    if (n) CU(ctx, cudaMemcpyAsync(out_f, d.eval_f.p, (size_t)(ne * n) * 8, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaMemcpyAsync(out_status, d.eval_status.p, (size_t)ne, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return CLOSYR_OK;
}

int closyr_score_partials_resident(closyr_ctx *ctx, int32_t max_leafs, int64_t chunk_begin, int64_t chunk_end,
                                   void *dev_partials)
{
    if (!ctx || !dev_partials || chunk_begin < 0 || chunk_end < chunk_begin)
        return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_score_partials_resident");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (ctx->n_points < 0) return fail(ctx, CLOSYR_E_STATE, "closyr_set_points has not been called");
    if (ctx->n_expr < 0) return fail(ctx, CLOSYR_E_STATE, "no resident population");
    const int64_t nch = n_chunks_of(ctx->n_points);
    chunk_end = std::min(chunk_end, nch);
    const int64_t span = pow2_ceil(std::max<int64_t>(1, chunk_end - chunk_begin));
    if (chunk_begin % span != 0 && chunk_begin < chunk_end)
        return fail(ctx, CLOSYR_E_ARG, "chunk range must be an aligned power-of-two span (fixed reduction tree)");
    DeviceState &d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.device));
    if (chunk_begin >= chunk_end) {
        CU(ctx, cudaMemsetAsync(dev_partials, 0, (size_t)ctx->n_expr * 3 * 8, d.stream));
    } else {
        int rc = launch_range(ctx, d, max_leafs, chunk_begin, chunk_end, false, (double *)dev_partials);
        if (rc) return rc;
    }
    CU(ctx, cudaStreamSynchronize(d.stream));
    return CLOSYR_OK;
}

int closyr_finalize_partials(closyr_ctx *ctx, const void *dev_partials_all, int n_parts, int32_t max_leafs)
{
    if (!ctx || !dev_partials_all || n_parts < 1) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_finalize_partials");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (ctx->n_expr < 0) return fail(ctx, CLOSYR_E_STATE, "no resident population");
    DeviceState &d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.device));
    if (ctx->n_expr > 0) {
        FinalizeParams fp;
        fp.meta = (const uint32_t *)d.pop.meta.p;
        fp.partials = (const double *)dev_partials_all;
        fp.part_stride = ctx->n_expr;
        fp.n_expr = ctx->n_expr;
        fp.n_tiles_total = n_parts;
        fp.max_leafs = max_leafs;
        fp.n = ctx->n_points;
        fp.scores = (double *)d.scores.p;
        fp.rflags = (uint8_t *)d.rflags.p;
        fp.out_partials = nullptr;
        launch_finalize(fp, d.stream);
        ctx->launches++;
        CU(ctx, cudaGetLastError());
    }
    CU(ctx, cudaStreamSynchronize(d.stream));
    ctx->scored = true;
    return CLOSYR_OK;
}

int closyr_score_from_values(closyr_ctx *ctx, const double *f, const double *ys, int64_t n, int32_t leafs, double *out3)
{
    if (!ctx || n < 0 || (n > 0 && (!f || !ys)) || !out3) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_score_from_values");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    DeviceState &d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.device));
    CU(ctx, d.eval_f.reserve((size_t)(2 * n + 4) * 8));
    double *df = (double *)d.eval_f.p, *dy = df + n, *dout = dy + n;
    if (n) {
        CU(ctx, cudaMemcpyAsync(df, f, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
        CU(ctx, cudaMemcpyAsync(dy, ys, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
    }
    launch_score_values(df, dy, n, leafs, dout, d.stream);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out3, dout, 3 * 8, cudaMemcpyDeviceToHost, d.stream));
    CU(ctx, cudaStreamSynchronize(d.stream));
    return CLOSYR_OK;
}

int closyr_rank_resident(closyr_ctx *ctx, int64_t k, int64_t *out_indices)
{
    if (!ctx || k < 0 || (k > 0 && !out_indices)) return fail(ctx, CLOSYR_E_ARG, "bad arguments to closyr_rank_resident");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err.clear();
    if (!ctx->scored) return fail(ctx, CLOSYR_E_STATE, "no scores: call closyr_score_resident first");
    const int64_t ne = ctx->n_expr;
    if (k > ne) k = ne;
    if (k == 0) return CLOSYR_OK;
    DeviceState &d0 = ctx->devs[0];
    CU(ctx, cudaSetDevice(d0.device));
    const bool by_points = (ctx->flags & CLOSYR_SHARD_POINTS) && ctx->devs.size() > 1;
    const double *scores = (const double *)d0.scores.p;
    CU(ctx, d0.rank_b.reserve((size_t)ne * 8 + (size_t)k * 8));
    if (ctx->devs.size() > 1) {
        // gather every device's scores on device 0 (peer copy over NVLink when available)
        for (auto &d : ctx->devs) {
            const int64_t lo = by_points ? d.fin_lo : d.pop.lo, off = by_points ? d.fin_lo : 0;
            const int64_t n_here = by_points ? d.fin_hi - d.fin_lo : d.pop.hi - d.pop.lo;
            if (n_here <= 0) continue;
            CU(ctx, cudaSetDevice(d.device));
            CU(ctx, cudaEventRecord(d.done, d.stream));
            CU(ctx, cudaSetDevice(d0.device));
            CU(ctx, cudaStreamWaitEvent(d0.stream, d.done, 0));
            CU(ctx, cudaMemcpyPeerAsync((double *)d0.rank_b.p + lo, d0.device, (const double *)d.scores.p + off, d.device,
                                        (size_t)n_here * 8, d0.stream));
        }
        scores = (const double *)d0.rank_b.p;
    }
    CU(ctx, d0.rank_a.reserve(rank_scratch_bytes(ne)));
    int64_t *d_out = (int64_t *)((double *)d0.rank_b.p + ne);
    ctx->launches += launch_rank(scores, ne, k, d0.rank_a.p, d_out, d0.stream);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out_indices, d_out, (size_t)k * 8, cudaMemcpyDeviceToHost, d0.stream));
    CU(ctx, cudaStreamSynchronize(d0.stream));
    return CLOSYR_OK;
}

int closyr_l2_flush(closyr_ctx *ctx)
{
    if (!ctx) return CLOSYR_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t bytes = (size_t)256 << 20;      // twice the 126 MB L2
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, d.flush.reserve(bytes));
        CU(ctx, cudaMemsetAsync(d.flush.p, 0x5a, bytes, d.stream));
    }
    for (auto &d : ctx->devs) {
        CU(ctx, cudaSetDevice(d.device));
        CU(ctx, cudaStreamSynchronize(d.stream));
    }
    return CLOSYR_OK;
}

int closyr_fp64_peak_probe(closyr_ctx *ctx, double *dfma_per_second)
{
    if (!ctx || !dfma_per_second) return CLOSYR_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &d = ctx->devs[0];
    CU(ctx, cudaSetDevice(d.device));
    const int blocks = 148 * 8, iters = 20000;
    CU(ctx, d.flush.reserve((size_t)blocks * 256 * 8));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CU(ctx, cudaEventRecord(d.ev0, d.stream));
        launch_fp64_probe((double *)d.flush.p, iters, blocks, d.stream);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaEventRecord(d.ev1, d.stream));
        CU(ctx, cudaStreamSynchronize(d.stream));
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        const double rate = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
        if (rep > 0) best = std::max(best, rate);
    }
    *dfma_per_second = best;
    return CLOSYR_OK;
}

int64_t closyr_kernel_launch_count(const closyr_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *closyr_host_alloc(size_t bytes) { return pinned_alloc(bytes); }
void closyr_host_free(void *p) { if (p) pinned_free(p); }

}  // extern "C"
