"""Row N3: the patch that lifts closyr's schema / CLI bounds and adds the GPU options applies to the reference checkout
(skipped on boxes without /root/reference, e.g. the GPU box), and touches exactly the lines SURVEY.md names."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "clj", "patches", "0001-gpu-scoring-limits-and-options.patch")
REF = "/root/reference"


def test_patch_names_the_three_files():
    text = open(PATCH).read()
    files = [l.split(" b/")[1].strip() for l in text.splitlines() if l.startswith("diff --git ")]
    assert files == ["src/closyr/core.clj", "src/closyr/symbolic_regression.clj", "src/closyr/util/spec.clj"]
    for needle in ("closyr.max-points", "closyr.max-population", "--gpus", "--shard-points", "gpu-ops/->gpu-run-args",
                   "gpu-ga/evolve", ":shard-points?"):
        assert needle in text, needle


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "closyr")) or shutil.which("patch") is None,
                    reason="needs the reference checkout and patch(1)")
def test_patch_applies_to_the_reference():
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(os.path.join(REF, "src"), os.path.join(tmp, "src"))
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", PATCH], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        subprocess.check_call(["patch", "-p1", "-s", "-i", PATCH], cwd=tmp)
        spec = open(os.path.join(tmp, "src", "closyr", "util", "spec.clj")).read()
        # the default bounds are the reference's own
        assert '(limit "closyr.max-points" 10000)' in spec and '(limit "closyr.max-population" 100000)' in spec
        # balanced parentheses in every touched file (a cheap syntax check without a Clojure reader)
        for f in ("core.clj", "symbolic_regression.clj", os.path.join("util", "spec.clj")):
            src = open(os.path.join(tmp, "src", "closyr", f)).read()
            depth = 0
            in_str = esc = False
            in_comment = False
            for ch in src:
                if in_comment:
                    in_comment = ch != "\n"
                    continue
                if in_str:
                    if esc:
                        esc = False
                    elif ch == "\\":
                        esc = True
                    elif ch == '"':
                        in_str = False
                    continue
                if ch == '"':
                    in_str = True
                elif ch == ";":
                    in_comment = True
                elif ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
            assert depth == 0, f
