"""The C-ABI library loads on a machine without a GPU, exports every symbol include/closyr_b200.h declares,
and fails loudly (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from closyr_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "closyr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(closyr_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = C.CDLL(L.SO_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(L.EXPORTS) == names, "closyr_b200/_lib.py EXPORTS and the header disagree"
    assert b"closyr_b200" in L.lib().closyr_version()


def test_builder_buffer_falls_back_when_page_locked_memory_runs_out():
    """a failed page-locked growth must release the old (page-locked) block with the page-locked
    deallocator, keep the data and continue in pageable memory."""
    lib = C.CDLL(L.SO_PATH)
    lib.closyr_selftest_pinned_fallback.restype = C.c_int
    assert lib.closyr_selftest_pinned_fallback() == 0


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "closyr_b200.h")).read()
    for cite in ("src/closyr/ops.clj:119-136", "src/closyr/ops/eval.clj:99-112", "src/closyr/symbolic_regression.clj:474-478",
                 "src/closyr/ga.clj:25-29", "src/closyr/ops.clj:195-201"):
        assert cite in text


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import closyr_b200 as cb
    with pytest.raises(cb.ClosyrError) as ei:
        cb.Context([0])
    assert ei.value.code == L.E_NO_GPU
    assert "no CPU fallback" in str(ei.value)


def test_isa_constants_match_header():
    text = open(os.path.join(ROOT, "include", "closyr_b200.h")).read()
    ops = dict(re.findall(r"\b(CL_[A-Z0-9]+)\s*=\s*(\d+)", text))
    from closyr_b200.population import OPNAMES
    for name, val in ops.items():
        if name == "CL_NUM_OPCODES":
            continue
        assert OPNAMES[int(val)].startswith(name[3:].rstrip("0")[:3]) or OPNAMES[int(val)] == name[3:], (name, val)


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: the header must compile as C11 (no C++-isms), and a C program that uses the host-side
    entry points (builder, text front-end, GA driver; nothing that needs a GPU) must link and run."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "abi_smoke.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "closyr_b200.h"
int main(void) {
    closyr_builder *b = 0;
    closyr_population view;
    int32_t leafs = 0;
    char *ff = 0;
    if (closyr_builder_create(0, &b) != CLOSYR_OK) return 1;
    if (closyr_builder_add(b, "x*x-1/2") != CLOSYR_OK) return 2;
    if (closyr_builder_add(b, "Plus(((") == CLOSYR_OK) return 3;              /* parse errors are reported */
    if (closyr_builder_view(b, &view) != CLOSYR_OK || view.n_expr != 1) return 4;
    if ((view.meta[0] & 0xffffu) != 7u) return 5;                              /* leafCount of -1/2+x^2 */
    if (closyr_leaf_count("x*x-1/2", &leafs) != CLOSYR_OK || leafs != 7) return 6;
    if (closyr_to_fullform("x*x-1/2", &ff) != CLOSYR_OK || strcmp(ff, "Plus(Rational(-1,2),Power(x,2))") != 0) return 7;
    closyr_free(ff);
    closyr_builder_destroy(b);
    int32_t r[2];
    if (closyr_jrandom_rand_ints(1, 50, 2, r) != CLOSYR_OK || r[0] != 36 || r[1] != 5) return 8;   /* util_prng_test.clj */
    if (closyr_mutation_count() != 96) return 9;
    closyr_ctx *ctx = 0;
    int dev = 0;
    int rc = closyr_ctx_create(&dev, 1, 0, &ctx);
    printf("ctx_create rc=%d\n", rc);                                           /* CLOSYR_E_NO_GPU without a device */
    if (rc == CLOSYR_OK) closyr_ctx_destroy(ctx);
    return 0;
}
''')
    exe = tmp_path / "abi_smoke"
    inc = os.path.join(ROOT, "include")
    so_dir = os.path.dirname(L.SO_PATH)
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe),
                    "-L", so_dir, "-lclosyr_b200", "-Wl,-rpath," + so_dir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "ctx_create rc=" in out.stdout


def test_comm_entry_points_without_a_gpu():
    """The NCCL id is plain bytes (no device needed to make one); joining a communicator needs a context, and a context
    needs a GPU: without one the calls fail with a code, they do not crash."""
    import ctypes as C
    lib = L.lib()
    ident = (C.c_uint8 * L.COMM_ID_BYTES)()
    rc = lib.closyr_comm_unique_id(ident)
    assert rc in (0, L.E_CUDA if hasattr(L, "E_CUDA") else -2)
    assert lib.closyr_comm_unique_id(None) != 0
    assert lib.closyr_comm_init_rank(None, 2, 0, ident) != 0


def test_kernels_use_the_hardware_paths_the_design_names():
    """SASS of the in-tree library: the threaded code arrives by TMA bulk copy (UBLKCP) against an mbarrier (SYNCS), the
    interpreters dispatch through branch-target tables (BRX), and nothing uses tensor cores."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(ROOT, "closyr_b200", "libclosyr_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    kernel = sass[sass.index("score_kernel"):]
    kernel = kernel[:kernel.index("Function :", 10)] if "Function :" in kernel[10:] else kernel
    assert "UBLKCP" in kernel and "SYNCS" in kernel and "BRX" in kernel
    assert "HMMA" not in sass and "UTCMMA" not in sass and "DMMA" not in sass
