"""CPU-side checks of the drop-in boundary: the product library loads, exports every symbol that
include/nsdg.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nsdg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsdg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    fns = declared_functions()
    for must in ("nsdg_create", "nsdg_set_mesh", "nsdg_set_field", "nsdg_step", "nsdg_get_field", "nsdg_update",
                 "nsdg_destroy", "nsdg_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(cuda_lib):
    from nextsimdg_b200 import capi

    fns = declared_functions()
    for fn in fns:
        assert hasattr(cuda_lib, fn), f"{fn} declared in nsdg.h but not exported"
    assert sorted(capi.SIGNATURES) == fns, "capi.SIGNATURES must list exactly the functions of nsdg.h"


def test_config_struct_matches_header(cuda_lib):
    from nextsimdg_b200 import capi

    cfg = capi.Config()
    ctypes.memset(ctypes.byref(cfg), 0xFF, ctypes.sizeof(cfg))
    cuda_lib.nsdg_config_default(ctypes.byref(cfg))
    assert (cfg.rheology, cfg.dgadv, cfg.cgdegree, cfg.nsteps, cfg.device) == (0, 6, 2, 100, -1)
    assert (cfg.alpha, cfg.beta) == (1500.0, 1500.0)
    assert list(cfg.neighbour) == [-1, -1, -1, -1] and cfg.nranks == 1 and cfg.global_nx == 0


def test_no_cpu_fallback_without_gpu(cuda_lib):
    """Without a CUDA device nsdg_create must fail loudly; with one it must succeed."""
    from nextsimdg_b200 import capi

    cfg = capi.Config()
    cuda_lib.nsdg_config_default(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    status = cuda_lib.nsdg_create(ctypes.byref(cfg), ctypes.byref(h))
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        assert status == 0
        cuda_lib.nsdg_destroy(h)
    else:
        assert status != 0
        assert b"CUDA" in cuda_lib.nsdg_last_error()


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (checked textually)."""
    pkg = os.path.join(ROOT, "nextsimdg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt and "nsdg_oracle" not in txt, f
