"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the driver reads,
and the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--workload", "c2", "--cpu-sample", "100", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "expr-point evals/sec" and d["unit"] == "expr-points/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["population"] == 1000 and d["config"]["points"] == 1024


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT,
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--workload", "c2", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""            # no JSON line, no silent CPU path


def test_reference_arm_loads_nothing_of_the_product():
    """The reference arm's process maps oracle/ only: no closyr_b200 module, no libclosyr_b200.so."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c2', '--steps', '1', '--warmup', '0']\n"
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libclosyr_b200' not in maps, 'the product library is mapped'\n"
            "assert 'liboracle' in maps\n"
            "assert not any(m == 'closyr_b200' or m.startswith('closyr_b200.') for m in sys.modules), 'closyr_b200 imported'\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_both_arms_describe_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.config_of("c3", "population")
    assert a["population"] == 100000 and a["points"] == 4096 and a["max_leafs"] == 40
    assert list(a) == ["workload", "population", "points", "max_leafs", "sharding", "generator_seed", "l2", "step"]
    # the committed cache of the default workload holds the whole population, so the reference arm needs no generator
    texts, source = bench.population_texts("c3", 100000)
    assert len(texts) == 100000 and source.startswith("committed cache")
