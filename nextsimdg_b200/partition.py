"""2-D box decomposition of the global grid across ranks (run/partition.cdl semantics) and the halo
plumbing for N GPUs.  Filled in by the multi-GPU milestone; single-domain runs never import this."""
from __future__ import annotations


def make_weak_scaling_box(*args, **kwargs):
    raise NotImplementedError("multi-GPU halo exchange is not built yet")
