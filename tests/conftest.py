import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The native pieces are built in-tree and kept out of git: on a checkout that has not been built yet (no
    # closyr_b200/libclosyr_b200.so or no oracle library) build them once, as __graft_entry__.build() does.
    so = os.path.join(ROOT, "closyr_b200", "libclosyr_b200.so")
    oracle_so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(so) and os.path.exists(oracle_so)):
        import __graft_entry__
        __graft_entry__.build()
