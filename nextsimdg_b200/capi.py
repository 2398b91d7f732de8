"""ctypes binding of libnsdg_cuda.so (include/nsdg.h).  Plumbing only: no arithmetic here."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_ubyte, c_void_p

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))

# enum nsdg_rheology / nsdg_field (include/nsdg.h)
MEVP, BBM, FREEDRIFT = 0, 1, 2
HICE, CICE, DAMAGE, U, V, UWIND, VWIND, UOCEAN, VOCEAN, SSH, TAUX, TAUY = range(12)
FIELD_IDS = {
    "hice": HICE, "cice": CICE, "damage": DAMAGE, "u": U, "v": V, "uwind": UWIND, "vwind": VWIND,
    "uocean": UOCEAN, "vocean": VOCEAN, "ssh": SSH, "uiostress": TAUX, "viostress": TAUY,
}
IPC_HANDLE_BYTES = 64


class NsdgError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message = nsdg_last_error())."""


class Config(Structure):
    _fields_ = [
        ("rheology", c_int), ("dgadv", c_int), ("cgdegree", c_int), ("nsteps", c_int), ("device", c_int),
        ("use_cuda_graph", c_int), ("force_general", c_int), ("pin_host_buffers", c_int),
        ("alpha", c_double), ("beta", c_double),
        ("global_nx", c_int), ("global_ny", c_int), ("box_x0", c_int), ("box_y0", c_int),
        ("rank", c_int), ("nranks", c_int), ("neighbour", c_int * 4), ("keep_dg_moments", c_int),
    ]


class UpdateIO(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "hice_in", "cice_in", "damage_in", "uwind", "vwind", "uocean", "vocean", "ssh",
        "hice_out", "cice_out", "damage_out", "u_out", "v_out", "taux_out", "tauy_out")]


class Timing(Structure):
    _fields_ = [("advection_ms", c_float), ("prepare_ms", c_float), ("subcycle_ms", c_float),
                ("total_ms", c_float), ("kernel_launches", c_long), ("uniform_path", c_int), ("halo_ms", c_float)]


# name -> (restype, argtypes); must list EVERY function declared in include/nsdg.h
SIGNATURES = {
    "nsdg_config_default": (None, [POINTER(Config)]),
    "nsdg_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "nsdg_destroy": (c_int, [c_void_p]),
    "nsdg_set_mesh": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int]),
    "nsdg_set_field": (c_int, [c_void_p, c_int, c_void_p, c_int]),
    "nsdg_step": (c_int, [c_void_p, c_double]),
    "nsdg_get_field": (c_int, [c_void_p, c_int, c_void_p, c_int]),
    "nsdg_set_boundaries": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "nsdg_advect_field": (c_int, [c_void_p, c_int, c_double, c_int, c_int, c_int, c_double, c_double]),
    "nsdg_set_benchmark_forcing": (c_int, [c_void_p, c_double, c_double, c_double]),
    "nsdg_update": (c_int, [c_void_p, POINTER(UpdateIO), c_double]),
    "nsdg_get_landmask": (c_int, [c_void_p, c_void_p]),
    "nsdg_get_dirichlet": (c_int, [c_void_p, c_int, c_void_p, c_size_t, POINTER(c_size_t)]),
    "nsdg_get_internal": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t, POINTER(c_size_t)]),
    "nsdg_set_internal": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t]),
    "nsdg_heal_damage": (c_int, [c_void_p, ctypes.c_double, ctypes.c_double, c_void_p]),
    "nsdg_get_state": (c_int, [c_void_p, c_void_p, c_size_t, POINTER(c_size_t)]),
    "nsdg_set_state": (c_int, [c_void_p, c_void_p, c_size_t]),
    "nsdg_subcycles": (c_int, [c_void_p, c_int, POINTER(c_float)]),
    "nsdg_time_kernels": (c_int, [c_void_p, c_int, POINTER(c_float), POINTER(c_float)]),
    "nsdg_get_timing": (c_int, [c_void_p, POINTER(Timing)]),
    "nsdg_halo_export": (c_int, [c_void_p, c_void_p]),
    "nsdg_halo_connect": (c_int, [c_void_p, c_int, c_void_p]),
    "nsdg_halo_ready": (c_int, [c_void_p]),
    "nsdg_last_error": (c_char_p, []),
    "nsdg_version": (c_char_p, []),
}

_lib = None


def library_path() -> str:
    return os.environ.get("NSDG_CUDA_LIB", os.path.join(_PKG_DIR, "libnsdg_cuda.so"))


def load_library():
    """Load libnsdg_cuda.so and type its entry points.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise NsdgError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int):
    if status != 0:
        raise NsdgError(load_library().nsdg_last_error().decode())


def as_c(a: np.ndarray):
    """Pointer to a C-contiguous float64 array (no copy; the caller keeps `a` alive)."""
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise NsdgError("arrays crossing the C ABI must be C-contiguous float64")
    return a.ctypes.data_as(c_void_p)
