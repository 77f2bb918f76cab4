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
