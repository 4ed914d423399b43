"""needlestack_b200 — the per-site robust negative-binomial error model of IARCbioinfo/needlestack
(bin/glm_rob_nb.r + the per-allele body of bin/needlestack.r) as hand-written FP64 CUDA for B200."""
from ._abi import (NS_EMIT_ALL, NS_EMIT_CALLS, NS_EMIT_FITTED, NS_F_EXTRA_ROB, NS_F_GATE1, NS_F_GATE2, NS_F_GATED,
                   NS_F_MAXIT, NS_F_NAN_ROWS, NS_F_QUAD_ERROR, NS_F_QUAD_OVERFLOW, NS_F_REF_INV, NS_F_SIG_AT_MIN,
                   NS_F_SKIPPED, NS_GT, NS_STATUS)
from .api import Caller, MultiCaller, NeedlestackError, PinnedArray, glmrob_nb, load_library, make_params

__all__ = ["Caller", "MultiCaller", "NeedlestackError", "PinnedArray", "glmrob_nb", "load_library", "make_params"]
