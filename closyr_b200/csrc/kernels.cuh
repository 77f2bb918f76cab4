// kernels.cuh -- launch parameters shared by kernels.cu and ctx.cu
#pragma once
#include <cstdint>

#include "../../include/closyr_b200.h"

namespace closyr {

constexpr int CHUNK_POINTS = 128;     // 32 lanes x 4 consecutive points: the leaf of the fixed reduction tree
constexpr int POINTS_PER_LANE = 4;

struct DevPopulation {                // device pointers of one resident shard
    const uint32_t *code;
    const uint32_t *code_off;         // [n_expr + 1], absolute word offsets of the caller's buffer
    const double *consts;
    const uint32_t *const_off;        // [n_expr + 1]
    const uint32_t *meta;
    const uint32_t *aux;
    uint32_t code_base;               // code_off value of the shard's first expression
    uint32_t const_base;
    int64_t n_expr;
    // what prepare_kernel leaves per shard (ctx.cu launches it once per upload / set_points): the threaded code of every
    // program (vm.cuh TInsn, 16 bytes per ISA word, same word index relative to the shard) and one descriptor per expression
    const uint4 *tcode;
    const struct ExprDesc *desc;
    // expression indices of the shard grouped by the handlers their programs use (order_kernel), or nullptr: the CTAs that
    // share an SM at the same time then run through the same routines (profiles/r02_design_experiments.md, "order")
    const uint32_t *order;
};

// one 16-byte load tells a CTA everything about its expression
struct __align__(16) ExprDesc {
    uint32_t tc_off;                  // first threaded-code entry of the program (index into DevPopulation::tcode)
    uint32_t ncode_pf;                // number of entries (low 16 bits) | PF_* program properties << 16
    uint32_t meta;                    // CLOSYR_META word
    uint32_t k_off;                   // first constant of the program (index into DevPopulation::consts)
};

struct PrepareParams {
    DevPopulation pop;                // tcode / desc are the outputs
    uint4 *tcode;
    ExprDesc *desc;
    int32_t have_tab;                 // leaf tables exist: SRCX functions become table loads
    int64_t leaf_stride;
    uint32_t *key;                    // [n_expr] ORDER_KEY_BITS-bit grouping key per expression (output)
};
constexpr int ORDER_KEY_BITS = 12;
void launch_prepare(const PrepareParams &p, void *stream);
// order[] = the expressions of a shard key by key (counting sort of prepare_kernel's keys), for DevPopulation::order
void launch_order(const uint32_t *key, int64_t n_expr, uint32_t *order, void *stream);

struct ScoreParams {
    DevPopulation pop;
    const double *xs;
    const double *ys;
    int64_t n;                        // total number of points of the run
    int64_t chunk_begin, chunk_end;   // 128-point chunks [begin, end) evaluated by this launch
    int32_t chunks_per_tile;          // power of two; one CTA reduces one tile
    int32_t warps;                    // warps per CTA (score_warps)
    int32_t n_tiles;                  // tiles per expression in this launch
    int32_t max_leafs;
    int32_t final_scores;             // 1: one tile covers every point -> write scores; 0: write partials
    double *scores;                   // [n_expr]
    uint8_t *rflags;                  // [n_expr]
    double *partials;                 // [n_tiles][n_expr][3] = sum, max, non-real count
    int64_t partial_tile_offset;      // index of this launch's first tile inside `partials`
    // f(x_i) of the raw variable for the parameter-free elementary functions, tabulated once per closyr_set_points
    // (every expression of every generation shares the points): row (op - CL_RECIP) of LEAF_TAB_ROWS rows of
    // `leaf_stride` doubles; nullptr = compute in the kernel
    const double *leaf_tab;
    int64_t leaf_stride;
};

constexpr int LEAF_TAB_ROWS = CL_ACOS - CL_RECIP + 1;
// RECIP SQRT RSQRT EXP LOG SIN COS TAN ASIN ACOS (bit = op - CL_RECIP); SQR, POWI and POWC are not tabulated
constexpr uint32_t LEAF_TAB_MASK = 0x1FCDu;
void launch_leaf_table(const double *xs, int64_t n, int64_t stride, double *tab, void *stream);

struct FinalizeParams {
    const uint32_t *meta;
    const double *partials;           // [n_tiles_total][part_stride][3], first expression of this launch at column 0
    int64_t part_stride;              // expressions per row of `partials` (>= n_expr: a launch may combine a slice)
    int64_t n_expr;
    int32_t n_tiles_total;
    int32_t max_leafs;
    int64_t n;
    double *scores;
    uint8_t *rflags;
    double *out_partials;             // non-NULL: emit one combined triple per expression instead of a score
};

struct EvalParams {
    DevPopulation pop;
    const double *xs;
    int64_t n;
    double *out_f;                    // [n_expr][n]
    uint8_t *status;                  // [n_expr]
};

int launch_score(const ScoreParams &p, uint32_t max_code_words, void *stream);     // returns the number of kernels launched
void launch_finalize(const FinalizeParams &p, void *stream);
void launch_eval(const EvalParams &p, size_t smem_bytes, void *stream);
void launch_score_values(const double *f, const double *ys, int64_t n, int32_t leafs, double *out3, void *stream);
void launch_fp64_probe(double *out, int iters, int blocks, void *stream);
size_t eval_smem_bytes(uint32_t max_code_words);
int score_warps(int32_t chunks_per_tile);
constexpr int MAX_CHUNKS_PER_TILE = 1024;   // per-chunk sums of a tile live in shared memory (8 KiB)

}  // namespace closyr
