/*
 * closyr_b200.h -- C ABI of the B200-native scoring path for closyr (plain C, no exceptions, no torch types).
 *
 * Drop-in boundary.  The reference has no FFI; the boundary is the Clojure closure
 *     (partial ops/score-fn run-args run-config)        src/closyr/symbolic_regression.clj:474-478
 * stored as :score-fn (src/closyr/ga.clj:11-17) and called per phenotype (src/closyr/ga.clj:25-29), plus
 * ops-eval/eval-vec-pheno used by the report (src/closyr/ops.clj:291-293).  The entry points below are
 * what a Java FFM (Panama) / JNA binding of those two functions binds; INTEGRATION.md shows the binding.
 * All file:line citations are relative to the reference checkout.
 *
 * Every function returns 0 on success or a negative CLOSYR_E_* code; closyr_last_error() gives the text.
 * The library never keeps a caller pointer after a call returns.
 */
#ifndef CLOSYR_B200_H
#define CLOSYR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLOSYR_ABI_VERSION 1

/* ------------------------------------------------------------------------------------------------
 * Sentinels (src/closyr/ops.clj:24-36)
 * ---------------------------------------------------------------------------------------------- */
#define CLOSYR_MIN_SCORE (-100000000.0)
#define CLOSYR_MAX_RESID (1000000.0)
#define CLOSYR_DEFAULT_MAX_LEAFS 40

/* ------------------------------------------------------------------------------------------------
 * Bytecode ISA v1.1 (DESIGN.md "Bytecode").  One instruction = one uint32:
 *     bits 0..7 opcode, bits 8..15 `a` (spill slot, or modifiers), bits 16..31 `b` (constant index or signed exponent).
 * The machine is a postfix evaluator whose top of stack is the accumulator `acc`; the rest of the
 * operand stack is addressed by STATIC slot numbers chosen by the compiler (the depth of a postfix
 * program is known at every instruction), so the kernel keeps slots 0..3 in registers.
 * Leaf operands (x, constants) are fused into the arithmetic instruction that consumes them.
 *
 * ISA v1.1 modifiers (a v1 program, all modifier bits 0, means what it meant).  The `a` field of the
 * leaf loads CL_LDX / CL_LDC and of the elementary functions CL_NEG .. CL_ABS is a modifier byte:
 *     CL_M_SRCX   elementary functions only: the operand is x, not acc         (was: LDX ; f)
 *     CL_M_PUSH   before anything else, slot[(a >> 2) & 3] = acc               (was: ST0+k ; ...)
 * so the very common  ST0 ; LDX ; SIN  is ONE instruction: SIN with a = CL_M_SRCX | CL_M_PUSH | 0 << 2.
 * The compiler applies both fusions wherever they are possible; a third of all dispatches disappear.
 * ---------------------------------------------------------------------------------------------- */
#define CL_M_SRCX 0x01
#define CL_M_PUSH 0x02
#define CL_M_PUSH_SLOT(a) (((a) >> 2) & 3)
enum closyr_opcode {
    CL_END = 0,
    CL_LDX = 1,      /* acc = x                                                        */
    CL_LDC = 2,      /* acc = c[b]                                                     */
    CL_LDZ = 3,      /* acc = c[b] + i*c[b+1]   (Complex literal; forces full semantics) */
    CL_LDSPEC = 4,   /* acc = a==0 ? ComplexInfinity : Indeterminate                   */
    CL_ADDX = 5,     /* acc = acc + x                                                  */
    CL_ADDC = 6,     /* acc = acc + c[b]                                               */
    CL_MULX = 7,     /* acc = acc * x                                                  */
    CL_MULC = 8,     /* acc = acc * c[b]                                               */
    CL_NEG = 9,      /* acc = -acc          (Times(-1, acc))                           */
    CL_RECIP = 10,   /* acc = acc^-1        (1/0.0 -> ComplexInfinity)                 */
    CL_SQR = 11,     /* acc = acc*acc       (Power(acc, 2))                            */
    CL_SQRT = 12,    /* acc = acc^(1/2)                                                */
    CL_RSQRT = 13,   /* acc = 1/(acc^(1/2)) (Power(acc, -1/2))                         */
    CL_POWI = 14,    /* acc = acc^b, b = int16, 3 <= |b| <= 32 or b == -2: square-and-multiply, MSB first */
    CL_POWC = 15,    /* acc = acc^c[b]      (real literal exponent, general pow)       */
    CL_EXP = 16,     /* acc = E^acc                                                    */
    CL_LOG = 17,
    CL_SIN = 18,
    CL_COS = 19,
    CL_TAN = 20,
    CL_ASIN = 21,
    CL_ACOS = 22,
    CL_ABS = 23,
    CL_ST = 24,      /* slot[a] = acc                (generic form, any a < 64)        */
    CL_ADDS = 25,    /* acc = slot[a] + acc                                            */
    CL_MULS = 26,    /* acc = slot[a] * acc                                            */
    CL_POWS = 27,    /* acc = slot[a] ^ acc          (general pow, any a)              */
    CL_ST0 = 28,     /* CL_ST0 + a   : slot[a] = acc       for a in 0..3 (register slots) */
    CL_ADDS0 = 32,   /* CL_ADDS0 + a : acc = slot[a] + acc for a in 0..3               */
    CL_MULS0 = 36,   /* CL_MULS0 + a : acc = slot[a] * acc for a in 0..3               */
    CL_NUM_OPCODES = 40
};
#define CLOSYR_REG_SLOTS 4
#define CLOSYR_MAX_SLOTS 64
#define CLOSYR_MAX_CODE_WORDS 2048   /* per expression, including END and padding */
#define CLOSYR_MAX_CONSTS 1024       /* per expression */
#define CLOSYR_INSN(op, a, b) ((uint32_t)(op) | ((uint32_t)(a) << 8) | ((uint32_t)((uint16_t)(b)) << 16))

/* per-expression metadata word: bits 0..15 leafCount (saturating), 16..23 CL_F_* flags, 24..31 slots used */
#define CL_F_CONST 0x01        /* expression does not depend on x (eval.clj:83-96 constant path)       */
#define CL_F_UNSUPPORTED 0x02  /* exceeds a structural limit above; scored as min-score                */
#define CL_F_NIL 0x04          /* whole expression is Indeterminate: eval returns nil (eval.clj:59-62) */
#define CLOSYR_META(leafs, flags, nslots) \
    ((uint32_t)((leafs) > 65535 ? 65535 : (leafs)) | ((uint32_t)(flags) << 16) | ((uint32_t)(nslots) << 24))

/* result flags written per expression by closyr_score_batch */
#define CL_R_TOO_BIG 0x01        /* leafs > max-leafs => min-score (ops.clj:128-129)                   */
#define CL_R_MIN_SCORE 0x02      /* nil / throwing eval => min-score (ops.clj:131-136)                 */
#define CL_R_FULL_SEMANTICS 0x04 /* diagnostic: some point's evaluation left the real line (complex values, ComplexInfinity) */

/*
 * The population: one flat struct-of-arrays buffer (north star (1)).  Programs are concatenated;
 * each program starts on a 16-byte boundary (code padded with CL_END to a multiple of 4 words,
 * constants padded to a multiple of 2 doubles) so a program can be staged with one TMA bulk copy.
 * For a constant expression whose head has `argc` arguments, aux = argc<<16 | offset (in doubles, inside
 * this expression's pool) of an argc-entry table holding each argument's literal value or +Inf: the
 * reference decodes such a result per argument when argc equals the point count (eval.clj:109).
 */
typedef struct closyr_population {
    int64_t n_expr;
    const uint32_t *code;       /* all instruction words                                   */
    const uint32_t *code_off;   /* [n_expr + 1] word offsets into code, multiples of 4     */
    const double *consts;       /* all constant pools                                      */
    const uint32_t *const_off;  /* [n_expr + 1] offsets into consts, multiples of 2        */
    const uint32_t *meta;       /* [n_expr] CLOSYR_META words                              */
    const uint32_t *aux;        /* [n_expr] see above; 0 when unused                       */
} closyr_population;

/* ------------------------------------------------------------------------------------------------
 * Error codes
 * ---------------------------------------------------------------------------------------------- */
#define CLOSYR_OK 0
#define CLOSYR_E_ARG (-1)      /* bad argument                                 */
#define CLOSYR_E_CUDA (-2)     /* CUDA runtime error, see closyr_last_error    */
#define CLOSYR_E_NO_GPU (-3)   /* no usable CUDA device: there is NO CPU fallback */
#define CLOSYR_E_STATE (-4)    /* call order (e.g. score before set_points)    */
#define CLOSYR_E_PARSE (-5)    /* expression text could not be parsed          */
#define CLOSYR_E_NOMEM (-6)

typedef struct closyr_ctx closyr_ctx;
typedef struct closyr_builder closyr_builder;

const char *closyr_version(void);
/* text of the last error on this thread (ctx may be NULL for creation / builder errors) */
const char *closyr_last_error(const closyr_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * Context: owns the CUDA streams, device buffers and pinned staging of 1..8 GPUs of one box.
 * Replaces the per-run state the reference keeps in `run-args`
 * (src/closyr/symbolic_regression.clj:355-383): xs/ys are marshalled once and stay resident.
 * ---------------------------------------------------------------------------------------------- */
#define CLOSYR_SHARD_POPULATION 0u /* expressions split across the devices, no collective (default)   */
#define CLOSYR_SHARD_POINTS 1u     /* points split across devices; partial (sum,max) combined in a fixed tree */

int closyr_ctx_create(const int *devices, int n_devices, uint32_t flags, closyr_ctx **out);
void closyr_ctx_destroy(closyr_ctx *ctx);
int closyr_ctx_device_count(const closyr_ctx *ctx);

/* replaces doubles->exprs / exprs->exprs-list / :input-ys-vec (common.clj:30-35, 68-76); copies */
int closyr_set_points(closyr_ctx *ctx, const double *xs, const double *ys, int64_t n);

/*
 * Batch form of ops/score-fn (src/closyr/ops.clj:119-136) for a whole generation: host buffers in,
 * host buffers out (H2D of the population, kernels, D2H of scores inside the call).
 * out_scores[n_expr] doubles; out_flags[n_expr] CL_R_* bytes (may be NULL).
 */
int closyr_score_batch(closyr_ctx *ctx, const closyr_population *pop, int32_t max_leafs, double *out_scores,
                       uint8_t *out_flags);

/* the same in three steps, so a caller (and bench.py) can keep a population resident in HBM */
int closyr_upload_population(closyr_ctx *ctx, const closyr_population *pop);
int closyr_score_resident(closyr_ctx *ctx, int32_t max_leafs, float *device_ms /* may be NULL */);
int closyr_download_scores(closyr_ctx *ctx, double *out_scores, uint8_t *out_flags);

/*
 * Batch form of ops-eval/eval-vec-pheno (src/closyr/ops/eval.clj:99-112) on an explicit xs vector
 * (the report path evaluates on xs and on the +-35 % extended domain, eval.clj:123-175).
 * out_f is [n_expr * n] row-major; out_status[i]: 0 vector valid, 1 nil, 2 the reference would throw.
 */
int closyr_eval_batch(closyr_ctx *ctx, const closyr_population *pop, const double *xs, int64_t n, double *out_f,
                      uint8_t *out_status);

/*
 * ops/compute-score-from-actuals-and-expecteds (src/closyr/ops.clj:98-116) on caller-provided f(xs):
 * residuals (ops.clj:78-90), their sum and max, and the score with the length deduction for `leafs`.
 * out3 = { score, resid-sum, max-resid }.  The same device reduction and epilogue as closyr_score_batch.
 */
int closyr_score_from_values(closyr_ctx *ctx, const double *f_of_xs, const double *ys, int64_t n, int32_t leafs,
                             double *out3);

/*
 * Device-side ranking of the scores of the resident population, bit-identical to the host's
 * (reverse (sort-by :score pop)) including tie order (src/closyr/ops.clj:195-201): descending score,
 * ties in DEscending original index.  Writes the first k indices.
 */
int closyr_rank_resident(closyr_ctx *ctx, int64_t k, int64_t *out_indices);

/*
 * Point sharding (CLOSYR_SHARD_POINTS): the points are split into aligned power-of-two spans of 128-point chunks, one
 * per GPU; every GPU evaluates all expressions on its span, the per-expression partial (sum, max, non-real count)
 * triples are exchanged with ONE NCCL all-gather over NVLink issued by this library on its own streams, and combined by
 * the same fixed binary tree, so a score is bit-identical for 1, 2, 4 or 8 GPUs.  With several devices in one context
 * (one process driving the box: the JVM's shape) closyr_ctx_create builds the communicators itself and
 * closyr_score_resident / closyr_score_batch do everything.  With one process per GPU each process creates a context
 * with ONE device and CLOSYR_SHARD_POINTS, rank 0 calls closyr_comm_unique_id, the caller hands the 128 bytes to every
 * rank (any transport), every rank calls closyr_comm_init_rank, and from then on closyr_score_resident is collective:
 * every rank calls it with the same resident population and returns ALL scores.
 * Replaces the pmap over the population for very large point counts (src/closyr/ga.clj:58-88).
 */
#define CLOSYR_COMM_ID_BYTES 128
int closyr_comm_unique_id(uint8_t *out_id /* CLOSYR_COMM_ID_BYTES */);
int closyr_comm_init_rank(closyr_ctx *ctx, int n_ranks, int rank, const uint8_t *id /* CLOSYR_COMM_ID_BYTES */);

/* point-sharded building blocks for callers that bring their own exchange (closyr_b200/sharder.py emulates ranks with them):
 * partials[3*n_expr] = {sum, max, nonreal_count} of this device's point range, combined afterwards. */
int closyr_score_partials_resident(closyr_ctx *ctx, int32_t max_leafs, int64_t chunk_begin, int64_t chunk_end,
                                   void *dev_partials /* device pointer, 3*n_expr doubles */);
int closyr_finalize_partials(closyr_ctx *ctx, const void *dev_partials_all /* [n_parts][3*n_expr] */, int n_parts,
                             int32_t max_leafs);

/* measurement helpers used by bench.py (not part of the scoring semantics) */
int closyr_l2_flush(closyr_ctx *ctx);
int closyr_fp64_peak_probe(closyr_ctx *ctx, double *dfma_per_second); /* register-resident DFMA chain */
int64_t closyr_kernel_launch_count(const closyr_ctx *ctx);

/* self test (no GPU needed): the builder's buffer falls back to pageable memory when page-locked memory runs out
 * and releases every block with the allocator it came from; 0 = ok */
int closyr_selftest_pinned_fallback(void);

/* pinned host memory for population buffers written directly by the caller (FFM MemorySegment) */
void *closyr_host_alloc(size_t bytes);
void closyr_host_free(void *p);

/* ------------------------------------------------------------------------------------------------
 * Expression compiler: lowers expression trees to the bytecode above (north star (1)).  In the Clojure
 * integration the same lowering walks the IExpr directly (clj/closyr/gpu/compile.clj); this C entry
 * point takes Symja text -- IExpr.fullFormString() or the printed form (str expr).
 * ---------------------------------------------------------------------------------------------- */
#define CLOSYR_BUILDER_PINNED 1u

int closyr_builder_create(uint32_t flags, closyr_builder **out);
void closyr_builder_destroy(closyr_builder *b);
void closyr_builder_clear(closyr_builder *b);
/* appends one expression; on CLOSYR_E_PARSE nothing is appended */
int closyr_builder_add(closyr_builder *b, const char *expr_text);
/* appends n expressions, compiling on nthreads host threads (0 = all cores); order is preserved */
int closyr_builder_add_many(closyr_builder *b, const char *const *expr_texts, int64_t n, int nthreads);
int closyr_builder_view(const closyr_builder *b, closyr_population *out);

/* canonical Symja FullForm text of a parsed expression (malloc'd; free with closyr_free) */
int closyr_to_fullform(const char *expr_text, char **out);
int closyr_leaf_count(const char *expr_text, int32_t *out);
void closyr_free(void *p);

/*
 * Synthetic populations for measurement: seeded random expression trees.
 * mix 0 = "GA mix" (max-leafs 40 configs), 1 = "transcendental-heavy" (config 4).
 * Appends n_expr programs to `b`; when out_texts != NULL it receives n_expr malloc'd FullForm strings
 * (free each with closyr_free) so a checker can score the same population.
 */
int closyr_generate_population(closyr_builder *b, uint64_t seed, int64_t n_expr, int32_t max_leafs, int32_t mix,
                               char **out_texts);

/*
 * Stand-alone GA driver: the host loop around the scorer for runs without a
 * JVM.  Replaces, for measurement purposes, ga/evolve (src/closyr/ga.clj:58-88), ops/mutation-fn and
 * ops/crossover-fn (src/closyr/ops.clj:143-188), ops.modify (src/closyr/ops/modify.clj:77-288) with the 96
 * operators of src/closyr/ops/initialize.clj:35-665, and the generator of src/closyr/util/prng.clj:16-59.
 * The scorer stays outside, as the `score-fn` closure does in ga/initialize (ga.clj:11-17):
 *     closyr_ga_pending -> closyr_score_batch (or any scorer) -> closyr_ga_commit_scores -> closyr_ga_evolve -> ...
 * No GPU is needed for anything in this block.
 */
typedef struct closyr_ga closyr_ga;

typedef struct closyr_ga_stats {
    int64_t n_pop;
    int64_t n_children;             /* phenotypes created by this evolve (they are the next pending batch) */
    int64_t n_mutation_children;
    int64_t n_crossover_children;
    double max_pair_sum;            /* max over pairs of e1-score + e2-score (ga.clj:46, :pop-scores) */
    int32_t early_stop;             /* any pair sum > -1e-3 (symbolic_regression.clj:409-414) */
    int32_t reserved;
    double ms_modify;               /* host time spent in shuffle + compete */
} closyr_ga_stats;

/* java.util.Random restated (golden pairs: test/closyr/util_prng_test.clj:22-25): count values of rand-int(n) */
int closyr_jrandom_rand_ints(int64_t seed, int64_t n, int32_t count, int32_t *out);
/* Collections.shuffle of 0..n-1 with a fresh generator (prng.clj:49-55) */
int closyr_jrandom_shuffle(int64_t seed, int32_t n, int32_t *perm);

/* the mutation table, in the order of initialize.clj */
int32_t closyr_mutation_count(void);
int closyr_mutation_info(int32_t index, const char **op, const char **label);
/* apply one operator to an expression; use_rng 0 = every should-modify draw says yes (rand stubbed to 0.0 as in
 * test/closyr/ops_modify_test.clj:754-756).  *out is malloc'd FullForm text (closyr_free). */
int closyr_mutate_expr(const char *expr_text, int32_t index, int32_t use_rng, int64_t seed, char **out);
int closyr_crossover_expr(const char *expr1, const char *expr2, int32_t max_leafs, int64_t seed, char **out);

/* population of pop_count phenotypes `x` (initialize.clj:17-33); threads <= 0: all host cores; threads == 1
 * keeps a single generator stream for the whole run */
int closyr_ga_create(int64_t pop_count, int32_t max_leafs, int64_t seed, int32_t threads, closyr_ga **out);
void closyr_ga_destroy(closyr_ga *ga);
int64_t closyr_ga_population_count(closyr_ga *ga);
int closyr_ga_set_expr(closyr_ga *ga, int64_t index, const char *expr_text);
/* the phenotypes without a score, compiled into a population (valid until the next closyr_ga_* call) */
int closyr_ga_pending(closyr_ga *ga, closyr_population *view, int64_t *n_pending);
int closyr_ga_pending_indices(closyr_ga *ga, int64_t *out, int64_t capacity);
int closyr_ga_commit_scores(closyr_ga *ga, const double *scores, int64_t n);
/* one ga/evolve on a fully scored population: shuffle, compete in pairs, create the children */
int closyr_ga_evolve(closyr_ga *ga, closyr_ga_stats *stats);
/* ops/sort-population's first element (ops.clj:191-201) */
int closyr_ga_best(closyr_ga *ga, int64_t *index, double *score, int32_t *leafs, char **expr_fullform);
int closyr_ga_expr(closyr_ga *ga, int64_t index, char **expr_fullform, double *score, int32_t *scored);

#ifdef __cplusplus
}
#endif
#endif /* CLOSYR_B200_H */
