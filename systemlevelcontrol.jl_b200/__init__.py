"""systemlevelcontrol.jl_b200 — B200-native drop-in for the column-wise localized H2 SLS synthesis hot
path of aaltoKEPO/SystemLevelControl.jl (`SLS_𝓗₂`, src/synthesis.jl:11-72; `sparsity_dim_reduction`,
src/reduction.jl:11-27).  The arithmetic lives in the CUDA library built from csrc/ (C ABI:
include/slsb200.h); this package is the host-side mirror of the reference interface.

The directory name contains a dot, so import it through the repo-root module `slc_b200`."""
from .synthesis import JuliaCSC, Plant, SLS_H2, Synthesizer, sparsity_dim_reduction  # noqa: F401
from . import _capi  # noqa: F401

__all__ = ["Plant", "SLS_H2", "sparsity_dim_reduction", "Synthesizer", "JuliaCSC"]
