"""admm_optim_b200 -- B200-native backend for the hot path of MultigridShapeOpt/admm_optim:
P1 assembly -> GMG-preconditioned BiCGStab -> ADMM prox/dual, behind the UG4-style object API
the reference's Lua drivers call.  The arithmetic lives in libadmm_b200.so (CUDA, sm_100a);
there is no CPU fallback."""
from ._lib import AdmmB200Error, LIB_PATH  # noqa: F401

__all__ = ["AdmmB200Error", "LIB_PATH"]
