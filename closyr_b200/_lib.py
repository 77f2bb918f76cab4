"""ctypes binding of include/closyr_b200.h (the same C ABI a Clojure FFM / JNA binding uses).

The shared library is built in-tree by `__graft_entry__.build()` (closyr_b200/csrc/Makefile).  There is no
fallback: if the library is missing, or there is no CUDA device when a context is created, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CLOSYR_B200_LIB", os.path.join(_HERE, "libclosyr_b200.so"))   # override: A/B builds

MIN_SCORE = -100000000.0      # reference src/closyr/ops.clj:24-26
MAX_RESID = 1000000.0         # ops.clj:34-36
DEFAULT_MAX_LEAFS = 40        # ops.clj:29-31

SHARD_POPULATION = 0
SHARD_POINTS = 1
COMM_ID_BYTES = 128
BUILDER_PINNED = 1

R_TOO_BIG = 0x01
R_MIN_SCORE = 0x02
R_FULL_SEMANTICS = 0x04

F_CONST = 0x01
F_UNSUPPORTED = 0x02
F_NIL = 0x04

E_NO_GPU = -3
E_PARSE = -5


class Population(C.Structure):
    """closyr_population: the flat struct-of-arrays buffer of one generation."""
    _fields_ = [
        ("n_expr", C.c_int64),
        ("code", C.POINTER(C.c_uint32)),
        ("code_off", C.POINTER(C.c_uint32)),
        ("consts", C.POINTER(C.c_double)),
        ("const_off", C.POINTER(C.c_uint32)),
        ("meta", C.POINTER(C.c_uint32)),
        ("aux", C.POINTER(C.c_uint32)),
    ]


# every symbol include/closyr_b200.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "closyr_version", "closyr_last_error",
    "closyr_ctx_create", "closyr_ctx_destroy", "closyr_ctx_device_count", "closyr_set_points",
    "closyr_score_batch", "closyr_upload_population", "closyr_score_resident", "closyr_download_scores",
    "closyr_eval_batch", "closyr_score_from_values", "closyr_rank_resident", "closyr_score_partials_resident", "closyr_finalize_partials",
    "closyr_comm_unique_id", "closyr_comm_init_rank",
    "closyr_l2_flush", "closyr_fp64_peak_probe", "closyr_kernel_launch_count", "closyr_host_alloc", "closyr_host_free",
    "closyr_selftest_pinned_fallback",
    "closyr_builder_create", "closyr_builder_destroy", "closyr_builder_clear", "closyr_builder_add",
    "closyr_builder_add_many", "closyr_builder_view", "closyr_to_fullform", "closyr_leaf_count", "closyr_free",
    "closyr_generate_population",
    "closyr_jrandom_rand_ints", "closyr_jrandom_shuffle", "closyr_mutation_count", "closyr_mutation_info",
    "closyr_mutate_expr", "closyr_crossover_expr", "closyr_ga_create", "closyr_ga_destroy",
    "closyr_ga_population_count", "closyr_ga_set_expr", "closyr_ga_pending", "closyr_ga_pending_indices",
    "closyr_ga_commit_scores", "closyr_ga_evolve", "closyr_ga_best", "closyr_ga_expr",
]


class GAStats(C.Structure):
    """closyr_ga_stats"""
    _fields_ = [
        ("n_pop", C.c_int64),
        ("n_children", C.c_int64),
        ("n_mutation_children", C.c_int64),
        ("n_crossover_children", C.c_int64),
        ("max_pair_sum", C.c_double),
        ("early_stop", C.c_int32),
        ("reserved", C.c_int32),
        ("ms_modify", C.c_double),
    ]

_lib = None


class ClosyrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("closyr_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the scoring path)" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    dp, u8p = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    popp = C.POINTER(Population)
    sig = {
        "closyr_version": (C.c_char_p, []),
        "closyr_last_error": (C.c_char_p, [vp]),
        "closyr_ctx_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, u32, C.POINTER(vp)]),
        "closyr_ctx_destroy": (None, [vp]),
        "closyr_ctx_device_count": (C.c_int, [vp]),
        "closyr_set_points": (C.c_int, [vp, dp, dp, i64]),
        "closyr_score_batch": (C.c_int, [vp, popp, i32, dp, u8p]),
        "closyr_upload_population": (C.c_int, [vp, popp]),
        "closyr_score_resident": (C.c_int, [vp, i32, C.POINTER(C.c_float)]),
        "closyr_download_scores": (C.c_int, [vp, dp, u8p]),
        "closyr_eval_batch": (C.c_int, [vp, popp, dp, i64, dp, u8p]),
        "closyr_score_from_values": (C.c_int, [vp, dp, dp, i64, i32, dp]),
        "closyr_rank_resident": (C.c_int, [vp, i64, C.POINTER(i64)]),
        "closyr_score_partials_resident": (C.c_int, [vp, i32, i64, i64, vp]),
        "closyr_finalize_partials": (C.c_int, [vp, vp, C.c_int, i32]),
        "closyr_l2_flush": (C.c_int, [vp]),
        "closyr_fp64_peak_probe": (C.c_int, [vp, dp]),
        "closyr_kernel_launch_count": (i64, [vp]),
        "closyr_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
        "closyr_comm_init_rank": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
        "closyr_host_alloc": (vp, [C.c_size_t]),
        "closyr_host_free": (None, [vp]),
        "closyr_builder_create": (C.c_int, [u32, C.POINTER(vp)]),
        "closyr_builder_destroy": (None, [vp]),
        "closyr_builder_clear": (None, [vp]),
        "closyr_builder_add": (C.c_int, [vp, C.c_char_p]),
        "closyr_builder_add_many": (C.c_int, [vp, C.POINTER(C.c_char_p), i64, C.c_int]),
        "closyr_builder_view": (C.c_int, [vp, popp]),
        "closyr_to_fullform": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
        "closyr_leaf_count": (C.c_int, [C.c_char_p, C.POINTER(i32)]),
        "closyr_free": (None, [vp]),
        "closyr_generate_population": (C.c_int, [vp, u64, i64, i32, i32, C.POINTER(vp)]),
        "closyr_jrandom_rand_ints": (C.c_int, [i64, i64, i32, C.POINTER(i32)]),
        "closyr_jrandom_shuffle": (C.c_int, [i64, i32, C.POINTER(i32)]),
        "closyr_mutation_count": (i32, []),
        "closyr_mutation_info": (C.c_int, [i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
        "closyr_mutate_expr": (C.c_int, [C.c_char_p, i32, i32, i64, C.POINTER(vp)]),
        "closyr_crossover_expr": (C.c_int, [C.c_char_p, C.c_char_p, i32, i64, C.POINTER(vp)]),
        "closyr_ga_create": (C.c_int, [i64, i32, i64, i32, C.POINTER(vp)]),
        "closyr_ga_destroy": (None, [vp]),
        "closyr_ga_population_count": (i64, [vp]),
        "closyr_ga_set_expr": (C.c_int, [vp, i64, C.c_char_p]),
        "closyr_ga_pending": (C.c_int, [vp, popp, C.POINTER(i64)]),
        "closyr_ga_pending_indices": (C.c_int, [vp, C.POINTER(i64), i64]),
        "closyr_ga_commit_scores": (C.c_int, [vp, dp, i64]),
        "closyr_ga_evolve": (C.c_int, [vp, C.POINTER(GAStats)]),
        "closyr_ga_best": (C.c_int, [vp, C.POINTER(i64), dp, C.POINTER(i32), C.POINTER(vp)]),
        "closyr_ga_expr": (C.c_int, [vp, i64, C.POINTER(vp), dp, C.POINTER(i32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(rc, ctx=None):
    if rc != 0:
        msg = lib().closyr_last_error(ctx)
        raise ClosyrError(rc, msg.decode() if msg else "unknown error")
