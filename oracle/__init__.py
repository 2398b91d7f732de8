"""ctypes binding of the CPU oracle (oracle/libnsdg_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (nextsimdg_b200) never does.

PARITY STATUS of the oracle: advection half pinned by the reference's KATs
(dynamics/test/Advection_test.cpp, AdvectionPeriodicBC_test.cpp) and mesh lists by the
25km_NH fixture; the momentum half (mEVP/BBM subcycle) is "parity unpinned" -- the reference
has no golden vector for it -- and is covered by analytic self-consistency tests only.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_int, c_long, c_void_p

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "libnsdg_oracle.so")
_lib = None

RADIANS = float.fromhex("0x1.1df46a2529d39p-6")


def build(force: bool = False):
    """Compile the C++ restatement (g++ -O3 -fopenmp) via oracle/Makefile."""
    args = ["make", "-C", _DIR] + (["-B"] if force else [])
    subprocess.run(args, check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.nso_create.restype = c_void_p
        L.nso_create.argtypes = [c_int, c_int, c_int, c_int]
        L.nso_destroy.argtypes = [c_void_p]
        L.nso_set_mesh.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int]
        L.nso_set_field.argtypes = [c_void_p, c_char_p, c_void_p, c_int]
        L.nso_update.argtypes = [c_void_p, c_double]
        L.nso_subcycles.argtypes = [c_void_p, c_int]
        L.nso_subcycles.restype = c_double
        L.nso_sweep.argtypes = [c_void_p, c_char_p]
        L.nso_set_delta_t.argtypes = [c_void_p, c_double]
        L.nso_last_subcycle_seconds.argtypes = [c_void_p]
        L.nso_last_subcycle_seconds.restype = c_double
        L.nso_get_dg0.argtypes = [c_void_p, c_char_p, c_void_p]
        L.nso_get_dg.argtypes = [c_void_p, c_char_p, c_void_p]
        L.nso_raw_size.argtypes = [c_void_p, c_char_p]
        L.nso_raw_size.restype = c_long
        L.nso_get_raw.argtypes = [c_void_p, c_char_p, c_void_p]
        L.nso_set_raw.argtypes = [c_void_p, c_char_p, c_void_p]
        L.nso_set_param.argtypes = [c_void_p, c_char_p, c_double]
        L.nso_dirichlet_size.argtypes = [c_void_p, c_int]
        L.nso_dirichlet_size.restype = c_long
        L.nso_get_dirichlet.argtypes = [c_void_p, c_int, c_void_p]
        L.nso_get_landmask.argtypes = [c_void_p, c_void_p]
        L.nso_get_vertices.argtypes = [c_void_p, c_void_p]
        L.nso_last_error.restype = c_char_p
        L.nso_set_threads.argtypes = [c_int]
        L.nso_max_threads.restype = c_int
        L.nso_kat_advection.restype = c_double
        L.nso_kat_advection.argtypes = [c_int, c_int, c_double, c_void_p]
        L.nso_kat_periodic.restype = c_double
        L.nso_kat_periodic.argtypes = [c_int, c_int, c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p)


class OracleDynamics:
    """Same surface as nextsimdg_b200.dynamics.CUDADynamicsBase, computed by the CPU restatement."""

    def __init__(self, rheology: str = "mevp", dgadv: int = 6, cgdegree: int = 2, nsteps: int = 100):
        self.L = load()
        self.rheology = rheology
        self.dgadv, self.cgdegree, self.nsteps = dgadv, cgdegree, nsteps
        self.h = self.L.nso_create({"mevp": 0, "bbm": 1, "freedrift": 2}[rheology], dgadv, cgdegree, nsteps)
        if not self.h:
            raise RuntimeError("oracle: unsupported (dgadv, cgdegree)")
        self.shared = {}
        self.nx = self.ny = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.nso_destroy(self.h)
            self.h = None

    def _chk(self, status):
        if status < 0:
            raise RuntimeError(self.L.nso_last_error().decode())

    def _set(self, name, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        self._chk(self.L.nso_set_field(self.h, name.encode(), _p(a), a.size // (self.nx * self.ny)))

    def setData(self, ms: dict):
        spherical = "longitude" in ms and "latitude" in ms
        coords = np.array(ms["coords"], dtype=np.float64, copy=True)
        if spherical:
            coords *= RADIANS
        mask = np.ascontiguousarray(ms["mask"], dtype=np.float64)
        self.ny, self.nx = mask.shape
        coords = np.ascontiguousarray(coords.reshape(-1))
        self._chk(self.L.nso_set_mesh(self.h, self.nx, self.ny, _p(coords), _p(mask), int(spherical)))
        for name in ("hice", "cice", "u", "v"):
            self._set(name, ms[name])
        if self.rheology == "bbm":
            self.damage = (np.array(ms["damage"], dtype=np.float64).reshape(self.ny, self.nx, -1)[..., 0].copy()
                           if "damage" in ms else np.where(mask == 1.0, 1.0, 1.7e38))
            self._set("damage", ms.get("damage", self.damage))

    def update(self, dt: float):
        s = self.shared
        if self.rheology == "bbm" and "damage" in s:
            np.copyto(self.damage, s["damage"])
        self._set("hice", s["hice"])
        self._set("cice", s["cice"])
        if self.rheology == "bbm":
            self._set("damage", self.damage)
        # FreeDriftDynamics::update passes only the ocean velocity (FreeDriftDynamics.hpp:44-45)
        for name in (("uocean", "vocean") if self.rheology == "freedrift" else ("uwind", "vwind", "uocean", "vocean", "ssh")):
            self._set(name, s[name])
        self._chk(self.L.nso_update(self.h, float(dt)))
        np.copyto(s["hice"], self.getDG0Data("hice"))
        np.copyto(s["cice"], self.getDG0Data("cice"))
        if self.rheology == "bbm":
            self.damage = self.getDG0Data("damage")
        self.uice, self.vice = self.getDG0Data("u"), self.getDG0Data("v")
        self.taux, self.tauy = self.getDG0Data("uiostress"), self.getDG0Data("viostress")

    def step(self, dt: float):
        self._chk(self.L.nso_update(self.h, float(dt)))

    def subcycles(self, n: int) -> float:
        return self.L.nso_subcycles(self.h, n)

    def sweep(self, which: str):
        self._chk(self.L.nso_sweep(self.h, which.encode()))

    def set_delta_t(self, deltaT: float):
        self.L.nso_set_delta_t(self.h, float(deltaT))

    def last_subcycle_seconds(self) -> float:
        return self.L.nso_last_subcycle_seconds(self.h)

    def getDG0Data(self, name):
        out = np.empty((self.ny, self.nx))
        self._chk(self.L.nso_get_dg0(self.h, name.encode(), _p(out)))
        return out

    def getDGData(self, name):
        out = np.empty((self.ny, self.nx, self.dgadv))
        self._chk(self.L.nso_get_dg(self.h, name.encode(), _p(out)))
        return out

    def internal(self, name):
        n = self.L.nso_raw_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n)
        self._chk(self.L.nso_get_raw(self.h, name.encode(), _p(out)))
        return out

    def set_internal(self, name, data):
        a = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
        assert a.size == self.L.nso_raw_size(self.h, name.encode())
        self._chk(self.L.nso_set_raw(self.h, name.encode(), _p(a)))

    def set_param(self, name, value):
        self._chk(self.L.nso_set_param(self.h, name.encode(), float(value)))

    def landmask(self):
        out = np.empty(self.nx * self.ny, dtype=np.uint8)
        self.L.nso_get_landmask(self.h, _p(out))
        return out

    def dirichlet(self, edge):
        out = np.empty(self.L.nso_dirichlet_size(self.h, edge), dtype=np.int64)
        self.L.nso_get_dirichlet(self.h, edge, _p(out))
        return out

    def vertices(self):
        out = np.empty(((self.ny + 1) * (self.nx + 1), 2))
        self.L.nso_get_vertices(self.h, _p(out))
        return out
