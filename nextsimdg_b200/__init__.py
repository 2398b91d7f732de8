"""nextsimdg_b200 -- B200-native (sm_100a, FP64) dynamics hot path of neXtSIM_DG.

The product is the C-ABI shared library ``libnsdg_cuda.so`` (``include/nsdg.h``), built from
``nextsimdg_b200/csrc``.  This package is the thin Python host layer above it:

* :mod:`nextsimdg_b200.capi`      -- ctypes binding of every ``nsdg_*`` entry point
* :mod:`nextsimdg_b200.dynamics`  -- ``CUDAMEVPDynamics`` / ``CUDABBMDynamics``: mirrors of the
  reference's ``IDynamics`` modules (core/src/modules/DynamicsModule/MEVPDynamics.cpp,
  BBMDynamics.cpp) used by the parity tests and the benchmark
* :mod:`nextsimdg_b200.synthetic` -- the deterministic synthetic inputs of SURVEY.md 8(d)
* :mod:`nextsimdg_b200.partition` -- 2-D box decomposition + halo plumbing for N GPUs

There is no CPU fallback anywhere in this package: if the CUDA library is missing or no GPU is
usable, the calls raise.
"""
from .capi import NsdgError, load_library, library_path  # noqa: F401
from .dynamics import CUDABBMDynamics, CUDADynamicsBase, CUDAFreeDriftDynamics, CUDAMEVPDynamics  # noqa: F401

__all__ = [
    "NsdgError",
    "load_library",
    "library_path",
    "CUDAMEVPDynamics",
    "CUDABBMDynamics",
    "CUDAFreeDriftDynamics",
    "CUDADynamicsBase",
]
