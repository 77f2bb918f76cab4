"""closyr_b200 -- B200-native scoring path for closyr.

The package holds only what the hot path needs: csrc/ (CUDA kernels, the C ABI, the expression compiler) and
the host-side mirror of the reference's operator interface (ops.py).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import ClosyrError, MIN_SCORE, MAX_RESID, DEFAULT_MAX_LEAFS  # noqa: F401
from .context import Context  # noqa: F401
from .population import PopulationBuilder, to_fullform, leaf_count, disassemble  # noqa: F401
