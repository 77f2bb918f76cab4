// rank.cu -- device-side ranking of a scored population.
//
// The reference ranks with (reverse (sort-by :score pop)) (src/closyr/ops.clj:195-201): a stable ascending
// sort followed by a reversal, i.e. descending score with ties in DEscending original index.  Here every
// element gets the strict total-order key (~orderable(score), ~index) and is sorted ascending with a bitonic
// network, so the result is bit-identical to the host procedure without needing a stable sort.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace closyr {
namespace {

struct Key { unsigned long long hi; unsigned long long lo; };

__device__ __forceinline__ bool key_less(const Key &a, const Key &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

__global__ void rank_keys_kernel(const double *scores, int64_t n, int64_t n_pad, Key *keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    Key k;
    if (i < n) {
        double s = scores[i];
        if (s == 0.0) s = 0.0;                                  // -0.0 and +0.0 compare equal on the host
        unsigned long long u = (unsigned long long)__double_as_longlong(s);
        u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);   // ascending order of doubles
        k.hi = ~u;                                              // descending score
        k.lo = ~(unsigned long long)i;                          // ties: descending index
    } else {
        k.hi = ~0ull;                                           // padding sorts last
        k.lo = ~0ull;
    }
    keys[i] = k;
}

__global__ void bitonic_step_kernel(Key *keys, int64_t n_pad, int64_t j, int64_t k)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int64_t l = i ^ j;
    if (l <= i) return;
    const Key a = keys[i], b = keys[l];
    const bool up = (i & k) == 0;
    if (key_less(b, a) == up) { keys[i] = b; keys[l] = a; }
}

// The steps with j < RANK_TILE stay inside aligned tiles of RANK_TILE keys: one block keeps its tile in shared memory
// (32 KiB) and runs all of them between two passes over HBM.  `k_first`..`k_last`: the merge sizes to run (k_first ==
// k_last > RANK_TILE: only the tail j = RANK_TILE / 2 .. 1 of that merge; otherwise every step of every k in the range).
constexpr int RANK_TILE = 2048;

__global__ void __launch_bounds__(RANK_TILE / 2) bitonic_tile_kernel(Key *keys, int64_t k_first, int64_t k_last)
{
    __shared__ Key tile[RANK_TILE];
    const int64_t base = (int64_t)blockIdx.x * RANK_TILE;
    const int t = threadIdx.x;
    tile[t] = keys[base + t];
    tile[t + RANK_TILE / 2] = keys[base + t + RANK_TILE / 2];
    __syncthreads();
    for (int64_t k = k_first; k <= k_last; k <<= 1) {
        for (int64_t j = (k > RANK_TILE ? RANK_TILE : k) >> 1; j > 0; j >>= 1) {
            // thread t owns the pair (i, i ^ j) with bit j of i clear
            const int i = (int)(((t & ~((int)j - 1)) << 1) | (t & ((int)j - 1)));
            const int l = i | (int)j;
            const bool up = ((base + i) & k) == 0;
            const Key a = tile[i], b = tile[l];
            if (key_less(b, a) == up) { tile[i] = b; tile[l] = a; }
            __syncthreads();
        }
    }
    keys[base + t] = tile[t];
    keys[base + t + RANK_TILE / 2] = tile[t + RANK_TILE / 2];
}

__global__ void rank_emit_kernel(const Key *keys, int64_t k, int64_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = (int64_t)(~keys[i].lo);
}

}  // namespace

size_t rank_scratch_bytes(int64_t n)
{
    int64_t p = 1;
    while (p < n) p <<= 1;
    return (size_t)p * sizeof(Key);
}

// scores: device pointer [n]; scratch: rank_scratch_bytes(n); out: device pointer [k].  Returns launches made.
int launch_rank(const double *scores, int64_t n, int64_t k, void *scratch, int64_t *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    Key *keys = (Key *)scratch;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n_pad + threads - 1) / threads);
    int launches = 0;
    rank_keys_kernel<<<blocks, threads, 0, stream>>>(scores, n, n_pad, keys);
    launches++;
    if (n_pad < RANK_TILE) {
        for (int64_t kk = 2; kk <= n_pad; kk <<= 1)
            for (int64_t j = kk >> 1; j > 0; j >>= 1) {
                bitonic_step_kernel<<<blocks, threads, 0, stream>>>(keys, n_pad, j, kk);
                launches++;
            }
    } else {
        // 16 bytes per key per pass over HBM: every merge size up to the tile in ONE pass, then per merge size the steps
        // that cross tiles one pass each and the rest in one pass (1M keys: 55 launches instead of 211)
        const unsigned tiles = (unsigned)(n_pad / RANK_TILE);
        bitonic_tile_kernel<<<tiles, RANK_TILE / 2, 0, stream>>>(keys, 2, RANK_TILE);
        launches++;
        for (int64_t kk = 2 * RANK_TILE; kk <= n_pad; kk <<= 1) {
            for (int64_t j = kk >> 1; j >= RANK_TILE; j >>= 1) {
                bitonic_step_kernel<<<blocks, threads, 0, stream>>>(keys, n_pad, j, kk);
                launches++;
            }
            bitonic_tile_kernel<<<tiles, RANK_TILE / 2, 0, stream>>>(keys, kk, kk);
            launches++;
        }
    }
    rank_emit_kernel<<<(unsigned)((k + threads - 1) / threads), threads, 0, stream>>>(keys, k, out);
    launches++;
    return launches;
}

}  // namespace closyr
