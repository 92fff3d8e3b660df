"""gen_fvgn_steady_b200 -- B200-native (sm_100a) implementation of the Gen-FVGN training hot path:
GN message passing (Encoder / GnBlock / Decoder) + differentiable finite-volume PDE loss, behind the reference's
src/FVMmodel module API.  Hand-written CUDA kernels through a C-ABI library (include/fvgn_b200.h); no CPU path."""
from .data import Data  # noqa: F401
from .plan import GraphPlan  # noqa: F401

__all__ = ["Data", "GraphPlan"]
