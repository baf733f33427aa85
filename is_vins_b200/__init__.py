"""B200-native IS-VINS sliding-window marginalization + sparsification backend.

Product code: the CUDA library `libisv_b200.so` (is_vins_b200/csrc, C ABI in include/isv_capi.h) and
the thin host mirror in this package.  Nothing here imports `oracle/` (test infrastructure).
"""
from . import capi  # noqa: F401
from .batch import WindowBatch, WindowOutputs, pack_events  # noqa: F401
from .backend import DeviceBatch, MargBackend  # noqa: F401
from .evaluate import DeviceProblem, FactorProblem, eval_problem  # noqa: F401
from .sequence import SequenceState  # noqa: F401
from .marginalization import MarginalizationInfo, PriorState, ResidualBlockInfo, add_margin_old_blocks  # noqa: F401

__all__ = ["capi", "WindowBatch", "WindowOutputs", "pack_events", "DeviceBatch", "MargBackend",
           "FactorProblem", "DeviceProblem", "eval_problem", "SequenceState", "MarginalizationInfo", "PriorState", "ResidualBlockInfo",
           "add_margin_old_blocks"]
from .forensic import forensic_batch, literal_forward  # noqa: F401
