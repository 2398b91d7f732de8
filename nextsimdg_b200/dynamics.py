"""Host-side mirrors of the reference's IDynamics modules, above the C ABI.

``CUDAMEVPDynamics`` / ``CUDABBMDynamics`` follow core/src/modules/DynamicsModule/MEVPDynamics.cpp
and BBMDynamics.cpp method for method (same names, argument meaning and error behaviour), so the
parity tests read like the reference's own module code:

    dyn = CUDAMEVPDynamics()
    dyn.setData(ms)            # ms: dict keyed by core/src/include/gridNames.hpp names
    dyn.update(tst_seconds)    # reads / writes the shared arrays in dyn.shared
    dyn.getState()

The shared-array store of the reference (ModelArrayRef<Shared::H_ICE> ..., IDynamics.hpp:89-99)
is the dict ``dyn.shared`` of C-contiguous float64 (ny, nx) arrays, index [j, i] = element i + nx*j.
All arithmetic happens in libnsdg_cuda.so; nothing here computes.
"""
from __future__ import annotations

import ctypes
from ctypes import c_size_t, c_void_p

import numpy as np

from . import capi
from .capi import NsdgError, as_c, check

# Degrees to radians as the reference's hex float (MEVPDynamics.cpp:21)
RADIANS = float.fromhex("0x1.1df46a2529d39p-6")


class CUDADynamicsBase:
    """Common part: IDynamics (core/src/modules/include/IDynamics.hpp:17-121)."""

    rheology = capi.MEVP
    _uses_damage = False
    named_fields = ("hice", "cice", "u", "v")  # MEVPDynamics.cpp:28

    def __init__(self, dgadv: int = 6, cgdegree: int = 2, nsteps: int = 100, device: int = -1,
                 use_cuda_graph: bool = True, force_general: bool = False, pin_host_buffers: bool = False,
                 partition=None, keep_dg_moments: bool = False):
        self._lib = capi.load_library()
        cfg = capi.Config()
        self._lib.nsdg_config_default(ctypes.byref(cfg))
        cfg.rheology = self.rheology
        cfg.dgadv = dgadv
        cfg.cgdegree = cgdegree
        cfg.nsteps = nsteps
        cfg.device = device
        cfg.use_cuda_graph = int(use_cuda_graph)
        cfg.force_general = int(force_general)
        cfg.pin_host_buffers = int(pin_host_buffers)
        cfg.keep_dg_moments = int(keep_dg_moments)  # opt-in extension (include/nsdg.h); the reference zeroes the moments (Q4)
        if partition is not None:
            partition.fill_config(cfg)
        self.cfg = cfg
        self.dgadv = dgadv
        self.cgdegree = cgdegree
        self.nsteps = nsteps
        self._h = c_void_p()
        check(self._lib.nsdg_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.nx = self.ny = 0
        self.shared = {}
        # members of IDynamics (IDynamics.hpp:79-88)
        self.uice = self.vice = self.damage = self.taux = self.tauy = None

    # -- lifetime: Module<IDynamics> owns one instance, destroyed by the Finalizer (Module.hpp:171-175)
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nsdg_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getName(self):
        return "IDynamics"

    def usesDamage(self):
        return self._uses_damage

    @staticmethod
    def checkSpherical(ms) -> bool:
        """IDynamics::checkSpherical, IDynamics.hpp:107-120."""
        if "longitude" in ms and "latitude" in ms:
            return True
        if "x" in ms and "y" in ms:
            return False
        raise RuntimeError("Input data must contain either Cartesian (x, y) or spherical (longitude, latitude) coordinates.")

    # -- low-level pass-throughs to the kernel object (DynamicsKernel::setData / getDG0Data / getDGData)
    def _set(self, name: str, data: np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.float64)
        n = self.nx * self.ny
        if a.size % n:
            raise NsdgError(f"field {name}: size {a.size} is not a multiple of nx*ny={n}")
        check(self._lib.nsdg_set_field(self._h, capi.FIELD_IDS[name], as_c(a), a.size // n))

    def getDG0Data(self, name: str) -> np.ndarray:
        out = np.empty((self.ny, self.nx))
        check(self._lib.nsdg_get_field(self._h, capi.FIELD_IDS[name], as_c(out), 1))
        return out

    def getDGData(self, name: str) -> np.ndarray:
        out = np.empty((self.ny, self.nx, self.dgadv))
        check(self._lib.nsdg_get_field(self._h, capi.FIELD_IDS[name], as_c(out), self.dgadv))
        return out

    # -- IDynamics::setData + MEVPDynamics::setData (MEVPDynamics.cpp:37-57)
    def setData(self, ms: dict):
        spherical = self.checkSpherical(ms)
        coords = np.array(ms["coords"], dtype=np.float64, copy=True)  # KeyError like ms.at()
        mask = np.ascontiguousarray(ms["mask"], dtype=np.float64)
        if spherical:
            coords *= RADIANS
        ny, nx = mask.shape
        if coords.size != (nx + 1) * (ny + 1) * 2:
            raise NsdgError("coords must be a VERTEX array of (ny+1, nx+1, 2)")
        self.nx, self.ny = nx, ny
        coords = np.ascontiguousarray(coords.reshape(-1))
        check(self._lib.nsdg_set_mesh(self._h, nx, ny, as_c(coords), as_c(mask), int(spherical)))
        self.uice = np.array(ms["u"], dtype=np.float64).reshape(ny, nx, -1)[..., 0].copy()
        self.vice = np.array(ms["v"], dtype=np.float64).reshape(ny, nx, -1)[..., 0].copy()
        self.damage = np.zeros((ny, nx))  # IDynamics.hpp:62-70: damage = 0 unless the module uses it
        self.taux = np.zeros((ny, nx))
        self.tauy = np.zeros((ny, nx))
        for name in self.named_fields:
            self._set(name, ms[name])
        self._set_defaults(ms, mask)

    def _set_defaults(self, ms, mask):
        pass

    def update(self, tst_seconds: float):
        raise NotImplementedError

    # -- IDynamics::getState (IDynamics.hpp:46-53): u, v masked
    def getState(self):
        return {"u": self.uice, "v": self.vice}

    # -- diagnostics / test access
    def internal(self, name: str) -> np.ndarray:
        n = c_size_t()
        check(self._lib.nsdg_get_internal(self._h, name.encode(), None, 0, ctypes.byref(n)))
        out = np.empty(n.value)
        check(self._lib.nsdg_get_internal(self._h, name.encode(), as_c(out), out.size, ctypes.byref(n)))
        return out

    def set_internal(self, name: str, data: np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
        check(self._lib.nsdg_set_internal(self._h, name.encode(), as_c(a), a.size))

    def heal_damage(self, dt: float, td_seconds: float = 15 * 86400.0, delta_cice: np.ndarray | None = None):
        """Nextsim::ConstantHealing on the device-resident DG0 damage (ConstantHealing.cpp:53-72); BBM handles only."""
        ptr = None
        if delta_cice is not None:
            self._dci = np.ascontiguousarray(delta_cice, dtype=np.float64)
            ptr = self._dci.ctypes.data_as(c_void_p)
        check(self._lib.nsdg_heal_damage(self._h, float(dt), float(td_seconds), ptr))

    # -- restart state: everything carried from one timestep to the next, one flat float64 buffer
    def get_state(self) -> np.ndarray:
        n = ctypes.c_size_t()
        check(self._lib.nsdg_get_state(self._h, None, 0, ctypes.byref(n)))
        out = np.empty(n.value)
        check(self._lib.nsdg_get_state(self._h, out.ctypes.data_as(c_void_p), out.size, ctypes.byref(n)))
        return out

    def set_state(self, state: np.ndarray):
        a = np.ascontiguousarray(state, dtype=np.float64)
        check(self._lib.nsdg_set_state(self._h, a.ctypes.data_as(c_void_p), a.size))

    def landmask(self) -> np.ndarray:
        out = np.empty(self.nx * self.ny, dtype=np.uint8)
        check(self._lib.nsdg_get_landmask(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def dirichlet(self, edge: int) -> np.ndarray:
        n = c_size_t()
        check(self._lib.nsdg_get_dirichlet(self._h, edge, None, 0, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.int64)
        check(self._lib.nsdg_get_dirichlet(self._h, edge, out.ctypes.data_as(c_void_p), out.size, ctypes.byref(n)))
        return out

    def step(self, dt: float):
        check(self._lib.nsdg_step(self._h, float(dt)))

    def set_boundaries(self, dirichlet=None, periodic=None):
        """ParametricMesh::dirichlet / ::periodic assigned directly (ParametricMesh.hpp:76-79), as the reference's advection
        tests do.  dirichlet: None or a list of 4 element-index lists (None entries keep the mask-derived list);
        periodic: None or a list of segments, each a list of (type, c1, c2, edge)."""
        keep = []
        dptr = nptr = None
        if dirichlet is not None:
            # one spare entry so that an EMPTY list still has a non-null pointer (NULL means "keep the mask-derived list")
            arrs = [None if d is None else np.ascontiguousarray(np.append(np.asarray(d, dtype=np.int64), 0)) for d in dirichlet]
            keep += arrs
            dp = (c_void_p * 4)(*[None if a is None else a.ctypes.data_as(c_void_p) for a in arrs])
            nd = (c_size_t * 4)(*[0 if a is None else a.size - 1 for a in arrs])
            keep += [dp, nd]
            dptr, nptr = ctypes.cast(dp, c_void_p), ctypes.cast(nd, c_void_p)
        segs = [np.ascontiguousarray(sg, dtype=np.int64).reshape(-1, 4) for sg in (periodic or [])]
        flat = np.ascontiguousarray(np.concatenate(segs, axis=0)) if segs else np.zeros((0, 4), dtype=np.int64)
        sizes = (c_size_t * max(len(segs), 1))(*[sg.shape[0] for sg in segs])
        check(self._lib.nsdg_set_boundaries(self._h, dptr, nptr, flat.ctypes.data_as(c_void_p) if flat.size else None,
                                            ctypes.cast(sizes, c_void_p), len(segs)))
        del keep

    def advect_field(self, name: str, dt: float, rk_order: int = 2, nsteps: int = 1, limit_max=None, limit_min=None):
        """nsteps x {DGTransport::reinitnormalvelocity; DGTransport::step (rk1 / rk2 / rk3); LimitMax; LimitMin} on one DG
        field, with the DG velocity currently in the transport object (set_internal("velx" / "vely"))."""
        mode = (1 if limit_max is not None else 0) | (2 if limit_min is not None else 0)
        check(self._lib.nsdg_advect_field(self._h, capi.FIELD_IDS[name], float(dt), int(rk_order), int(nsteps), mode,
                                          float(limit_max or 0.0), float(limit_min or 0.0)))

    def set_benchmark_forcing(self, elapsed_seconds: float, domain_x: float = 512e3, domain_y: float = 512e3):
        """Benchmark{Atmosphere,Ocean} evaluated on the device (no forcing crosses PCIe)."""
        check(self._lib.nsdg_set_benchmark_forcing(self._h, float(elapsed_seconds), float(domain_x), float(domain_y)))

    def subcycles(self, n: int) -> float:
        ms = ctypes.c_float()
        check(self._lib.nsdg_subcycles(self._h, int(n), ctypes.byref(ms)))
        return ms.value

    def timing(self) -> capi.Timing:
        t = capi.Timing()
        check(self._lib.nsdg_get_timing(self._h, ctypes.byref(t)))
        return t

    def _update_io(self, use_damage: bool):
        """One fused nsdg_update call on the shared arrays (the module's whole update())."""
        s = self.shared
        io = capi.UpdateIO()
        keep = []

        def ptr(a):
            keep.append(a)
            return as_c(a)

        io.hice_in = ptr(s["hice"])
        io.cice_in = ptr(s["cice"])
        io.uwind, io.vwind = ptr(s["uwind"]), ptr(s["vwind"])
        io.uocean, io.vocean = ptr(s["uocean"]), ptr(s["vocean"])
        io.ssh = ptr(s["ssh"])
        io.hice_out, io.cice_out = ptr(s["hice"]), ptr(s["cice"])  # written back in place (MEVPDynamics.cpp:79-80)
        io.u_out, io.v_out = ptr(self.uice), ptr(self.vice)
        io.taux_out, io.tauy_out = ptr(self.taux), ptr(self.tauy)
        if use_damage:
            io.damage_in = ptr(self.damage)
            io.damage_out = ptr(self.damage)
        return io, keep


class CUDAMEVPDynamics(CUDADynamicsBase):
    """Drop-in for Nextsim::MEVPDynamics (MEVPDynamics.cpp:29-101)."""

    rheology = capi.MEVP

    def getName(self):
        return "CUDAMEVPDynamics"

    def update(self, tst_seconds: float):
        io, keep = self._update_io(False)
        check(self._lib.nsdg_update(self._h, ctypes.byref(io), float(tst_seconds)))
        del keep


class CUDABBMDynamics(CUDADynamicsBase):
    """Drop-in for Nextsim::BBMDynamics (BBMDynamics.cpp:19-132)."""

    rheology = capi.BBM
    _uses_damage = True

    def getName(self):
        return "CUDABBMDynamics"

    def _set_defaults(self, ms, mask):
        # BBMDynamics.cpp:49-63: damage defaults to 1.0, masked: ModelComponent::mask puts
        # MissingData::value (1.7e38, MissingData.hpp:17) on land (ModelComponent.cpp:96-110)
        if "damage" in ms:
            self._set("damage", ms["damage"])
            self.damage = np.array(ms["damage"], dtype=np.float64).reshape(self.ny, self.nx, -1)[..., 0].copy()
        else:
            self.damage = np.where(mask == 1.0, 1.0, 1.7e38)
            self._set("damage", self.damage)

    def update(self, tst_seconds: float):
        # BBMDynamics.cpp:71: damage = damage0 (the Protected::DAMAGE array of the store)
        if "damage" in self.shared:
            np.copyto(self.damage, self.shared["damage"])
        io, keep = self._update_io(True)
        check(self._lib.nsdg_update(self._h, ctypes.byref(io), float(tst_seconds)))
        del keep

    def getState(self):
        # BBMDynamics.cpp:104-117: full-DG hice, cice, damage
        st = super().getState()
        st.update({"hice": self.getDGData("hice"), "cice": self.getDGData("cice"), "damage": self.getDGData("damage")})
        return st


class CUDAFreeDriftDynamics(CUDADynamicsBase):
    """Drop-in for Nextsim::FreeDriftDynamics (core/src/modules/DynamicsModule/include/FreeDriftDynamics.hpp:26-83)."""

    rheology = capi.FREEDRIFT

    def getName(self):
        return "CUDAFreeDriftDynamics"

    def setData(self, ms: dict):
        # quirk: the reference passes isSpherical = false to kernel.initialise whatever the coordinates are
        # (FreeDriftDynamics.hpp:69), after scaling lon/lat to radians
        ms2 = dict(ms)
        if self.checkSpherical(ms):
            ms2["coords"] = np.asarray(ms["coords"], dtype=np.float64) * RADIANS
            ms2.pop("longitude")
            ms2.pop("latitude")
            ms2["x"] = ms2["y"] = np.zeros_like(np.asarray(ms["mask"], dtype=np.float64))
        super().setData(ms2)

    def update(self, tst_seconds: float):
        # FreeDriftDynamics::update sets hice, cice, uocean, vocean only (FreeDriftDynamics.hpp:39-55)
        s = self.shared
        for name in ("hice", "cice", "uocean", "vocean"):
            self._set(name, s[name])
        self.step(tst_seconds)
        np.copyto(s["hice"], self.getDG0Data("hice"))
        np.copyto(s["cice"], self.getDG0Data("cice"))
        self.uice = self.getDG0Data("u")
        self.vice = self.getDG0Data("v")
