"""2-D box decomposition of the global grid across ranks and the halo plumbing for N GPUs.

Semantics follow the reference's partition metadata (run/partition.cdl; bounding boxes read at
core/src/ModelMetadata.cpp:42-62; per-side neighbour ids as in core/test/partition_metadata_3.cdl:28-66):
rank r owns the box [x0, x0+nx) x [y0, y0+ny) of the global element grid.  The reference's dynamics has
no halo exchange at all, so everything beyond the box geometry is this repo's design:

* every box is extended by a one-element overlap ring towards each neighbour; the LOCAL mesh handed to
  ``nsdg_set_mesh`` includes that ring;
* ``libnsdg_cuda`` exchanges the non-owned CG node lines of (u, v) every subcycle and the ring column/row of
  each advected DG field every RK stage, by peer stores over NVLink into the neighbour's IPC-mapped arena;
* this module builds the boxes, routes the IPC handles between ranks (any transport with all-gather
  semantics: torch.distributed here) and crops results back to the owned box.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import check

BOTTOM, RIGHT, TOP, LEFT = 0, 1, 2, 3
OPPOSITE = {BOTTOM: TOP, RIGHT: LEFT, TOP: BOTTOM, LEFT: RIGHT}


def grid_shape(nranks: int) -> tuple[int, int]:
    """(px, py) of the box grid: 1x1, 2x1, 2x2, 4x2 ... (x gets the extra factor of two)."""
    px = py = 1
    n = nranks
    while n > 1:
        if n % 2:
            raise ValueError("number of ranks must be a power of two")
        if px <= py:
            px *= 2
        else:
            py *= 2
        n //= 2
    return px, py


@dataclass
class Partition:
    """One rank's box of a px x py decomposition of a global_nx x global_ny element grid."""

    rank: int
    nranks: int
    global_nx: int
    global_ny: int
    px: int
    py: int

    def __post_init__(self):
        if self.px * self.py != self.nranks:
            raise ValueError("px * py must equal the number of ranks")
        if self.global_nx % self.px or self.global_ny % self.py:
            raise ValueError("the box grid must divide the element grid")
        self.ix, self.iy = self.rank % self.px, self.rank // self.px
        self.nx, self.ny = self.global_nx // self.px, self.global_ny // self.py  # owned extent
        self.x0, self.y0 = self.ix * self.nx, self.iy * self.ny
        nb = [-1, -1, -1, -1]
        if self.iy > 0:
            nb[BOTTOM] = self.rank - self.px
        if self.ix < self.px - 1:
            nb[RIGHT] = self.rank + 1
        if self.iy < self.py - 1:
            nb[TOP] = self.rank + self.px
        if self.ix > 0:
            nb[LEFT] = self.rank - 1
        self.neighbour = nb
        self.ring = [1 if r >= 0 else 0 for r in nb]  # ring width towards each side
        # local (ring-extended) window in global element indices
        self.lx0, self.ly0 = self.x0 - self.ring[LEFT], self.y0 - self.ring[BOTTOM]
        self.lnx = self.nx + self.ring[LEFT] + self.ring[RIGHT]
        self.lny = self.ny + self.ring[BOTTOM] + self.ring[TOP]

    @classmethod
    def weak(cls, rank: int, nranks: int, n_per_rank: int) -> "Partition":
        px, py = grid_shape(nranks)
        return cls(rank, nranks, px * n_per_rank, py * n_per_rank, px, py)

    @classmethod
    def strong(cls, rank: int, nranks: int, global_nx: int, global_ny: int) -> "Partition":
        px, py = grid_shape(nranks)
        return cls(rank, nranks, global_nx, global_ny, px, py)

    def fill_config(self, cfg: capi.Config):
        cfg.global_nx, cfg.global_ny = self.global_nx, self.global_ny
        cfg.box_x0, cfg.box_y0 = self.x0, self.y0
        cfg.rank, cfg.nranks = self.rank, self.nranks
        for s in range(4):
            cfg.neighbour[s] = self.neighbour[s]

    # -- windows -------------------------------------------------------------------------------
    def local_window(self):
        """slices of the GLOBAL element arrays covered by the local (ring-extended) mesh"""
        return slice(self.ly0, self.ly0 + self.lny), slice(self.lx0, self.lx0 + self.lnx)

    def local_vertex_window(self):
        return slice(self.ly0, self.ly0 + self.lny + 1), slice(self.lx0, self.lx0 + self.lnx + 1)

    def owned_in_local(self):
        """slices of a LOCAL element array that hold the owned box"""
        return (slice(self.ring[BOTTOM], self.ring[BOTTOM] + self.ny), slice(self.ring[LEFT], self.ring[LEFT] + self.nx))

    def owned_window(self):
        return slice(self.y0, self.y0 + self.ny), slice(self.x0, self.x0 + self.nx)

    def crop_state(self, ms_global: dict) -> dict:
        """local (ring-extended) copy of a global model state dict (coords are VERTEX arrays)"""
        ew, vw = self.local_window(), self.local_vertex_window()
        out = {}
        for k, v in ms_global.items():
            a = np.asarray(v)
            out[k] = np.ascontiguousarray(a[vw] if k == "coords" else a[ew])
        return out


def connect_halos(dyn, part: Partition, all_gather):
    """Exchange the arena IPC handles and connect every neighbour side.

    ``all_gather(obj) -> list`` returns every rank's object in rank order (e.g. torch.distributed
    ``all_gather_object``); ``dyn`` is a CUDADynamicsBase whose mesh is already set.
    """
    lib = dyn._lib
    mine = (ctypes.c_ubyte * capi.IPC_HANDLE_BYTES)()
    check(lib.nsdg_halo_export(dyn._h, ctypes.cast(mine, ctypes.c_void_p)))
    handles = all_gather(bytes(mine))
    for side in range(4):
        peer = part.neighbour[side]
        if peer < 0:
            continue
        buf = (ctypes.c_ubyte * capi.IPC_HANDLE_BYTES).from_buffer_copy(handles[peer])
        check(lib.nsdg_halo_connect(dyn._h, side, ctypes.cast(buf, ctypes.c_void_p)))
    check(lib.nsdg_halo_ready(dyn._h))
    all_gather(b"ready")  # nobody starts exchanging before every box has mapped its neighbours


def torch_all_gather(dist):
    def gather(obj):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out

    return gather


# -- synthetic inputs generated directly for a window of a large global grid ---------------------------
def benchmark_window(part: Partition, L: float, t: float = 0.0):
    """The cyclone-box state and forcing (synthetic.benchmark_box / benchmark_forcing) for the local window
    of `part`, computed from global indices without materialising the global arrays."""
    gx, gy = part.global_nx, part.global_ny
    dx = L / gx
    iy = np.arange(part.ly0, part.ly0 + part.lny, dtype=np.float64)
    ix = np.arange(part.lx0, part.lx0 + part.lnx, dtype=np.float64)
    vx = np.arange(part.lx0, part.lx0 + part.lnx + 1, dtype=np.float64) * L / gx
    vy = np.arange(part.ly0, part.ly0 + part.lny + 1, dtype=np.float64) * L / gx  # square cells
    X, Y = np.meshgrid(vx, vy)
    coords = np.ascontiguousarray(np.stack([X, Y], axis=-1))
    mask = np.ones((part.lny, part.lnx))
    mask[(iy == 0) | (iy == gy - 1), :] = 0.0
    mask[:, (ix == 0) | (ix == gx - 1)] = 0.0
    a = iy[:, None] * dx
    b = ix[None, :] * dx
    hice = (0.3 + 0.005 * (np.sin(60e-6 * a) + np.sin(30e-6 * b))) * mask
    cice = np.ones_like(mask) * mask
    z = np.zeros_like(mask)
    ms = {"coords": coords, "mask": mask, "x": z, "y": z, "hice": np.ascontiguousarray(hice),
          "cice": np.ascontiguousarray(cice), "damage": np.ascontiguousarray(mask.copy()), "u": z.copy(), "v": z.copy()}
    x = ix[None, :] * dx + 0 * iy[:, None]
    y = iy[:, None] * dx + 0 * ix[None, :]
    x0 = y0 = (L / 2) * (1 + t / (5 * 86400.0))
    xp, yp = x - x0, y - y0
    s = 1e-5 * np.exp(-1e-5 * np.hypot(xp, yp))
    al = np.deg2rad(72.0)
    c = np.ascontiguousarray
    forcing = {"uwind": c(-s * 30.0 * (np.cos(al) * xp + np.sin(al) * yp)),
               "vwind": c(-s * 30.0 * (-np.sin(al) * xp + np.cos(al) * yp)),
               "uocean": c(0.01 * (2 * y / L - 1)), "vocean": c(0.01 * (1 - 2 * x / L)), "ssh": z.copy()}
    return ms, forcing


def make_weak_scaling_box(cls, n, rheo, rank, world, local_rank, dist, nsteps=100, make_inputs=None, cell=4000.0, strong=False):
    """bench.py helper: this rank's box of the weak-scaling run (n x n owned elements per GPU) or, with strong=True,
    of the strong-scaling run (n x n elements in total, split into px x py boxes)."""
    part = Partition.strong(rank, world, n, n) if strong else Partition.weak(rank, world, n)
    L = cell * part.global_nx
    ms, forcing = benchmark_window(part, L)
    dyn = cls(nsteps=nsteps, device=local_rank, pin_host_buffers=True, partition=part)
    dyn.setData(ms)
    connect_halos(dyn, part, torch_all_gather(dist))
    dyn.partition = part
    dyn.owned_elements = lambda: part.nx * part.ny
    return dyn, ms, forcing


# -- self-consistency probe: partitioned boxes against the same problem on one GPU ------------------------------------
PROBE_CASES = (("mevp", "uniform"), ("mevp", "distorted"), ("mevp", "spherical"), ("bbm", "uniform"), ("bbm", "distorted"))


def probe_inputs(kind: str):
    """Seeded small problems for the partition probe: a 96 x 64 para24x30-shaped box (uniform / distorted, irregular land
    mask) or the 64 x 64 TOPAZ-like spherical grid (run/init_topaz128x128.py:106-121 at half resolution)."""
    from . import synthetic

    if kind == "spherical":
        ms = synthetic.topaz_like_spherical(64)
        return ms, synthetic.smooth_forcing(64, 64)
    ms = synthetic.para_state(96, 64, dxy=8000.0, distort=0.04 if kind == "distorted" else 0.0, irregular_mask=True)
    return ms, synthetic.smooth_forcing(96, 64)


def _dg0(a):
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a[..., 0] if a.ndim == 3 else a)


def partition_probe(rheo: str, kind: str, rank: int, world: int, device: int, all_gather, nsteps: int = 40, nts: int = 2):
    """Run one probe case as `world` boxes with halo exchange AND (on rank 0) as a single domain on the same GPU.

    Returns on rank 0 {field: norm-wise relative error over ice elements of the gathered owned boxes against the
    single-domain result} plus "single" = the single-domain exports (so that a caller can check THEM against an
    independent checker); None on the other ranks.  The reference has no distributed dynamics: equality
    with the single-domain result is what multi-GPU parity means here."""
    from .dynamics import CUDABBMDynamics, CUDAMEVPDynamics

    # BBM is an explicit elastic scheme: keep c_elastic * dt_sub / dx well below 1, otherwise rounding noise is amplified
    # and no two summation orders agree ('conditioning'); on the sphere its damage time scale blows up within
    # a few subcycles in the reference itself (BBMStressUpdateStep.hpp:158), hence mEVP only there
    dt = 120.0 if rheo == "bbm" else 600.0
    ms, forc = probe_inputs(kind)
    gny, gnx = np.asarray(ms["mask"]).shape
    cls = CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics

    def drive(dyn, state, window):
        dyn.shared = {"hice": np.ascontiguousarray(_dg0(state["hice"])[window]), "cice": np.ascontiguousarray(_dg0(state["cice"])[window]),
                      **{k: np.ascontiguousarray(v[window]) for k, v in forc.items()}}
        for _ in range(nts):
            dyn.update(dt)
        out = {"u": dyn.uice, "v": dyn.vice, "hice": dyn.shared["hice"], "cice": dyn.shared["cice"], "taux": dyn.taux, "tauy": dyn.tauy}
        if rheo == "bbm":
            out["damage"] = dyn.damage
        return out

    part = Partition.strong(rank, world, gnx, gny)
    dyn = cls(nsteps=nsteps, device=device, partition=part)
    dyn.setData(part.crop_state(ms))
    connect_halos(dyn, part, all_gather)
    ow = part.owned_in_local()
    mine = {k: np.ascontiguousarray(v[ow]) for k, v in drive(dyn, ms, part.local_window()).items()}
    everyone = all_gather((part.owned_window(), mine))
    dyn.close()
    if rank != 0:
        return None
    ref = cls(nsteps=nsteps, device=device)
    ref.setData(ms)
    full = {k: v.copy() for k, v in drive(ref, ms, (slice(None), slice(None))).items()}
    ref.close()
    ice = np.asarray(ms["mask"]).astype(bool)
    errs = {}
    for name, whole in full.items():
        got = np.full_like(whole, np.nan)
        for win, fields in everyone:
            got[win] = fields[name]
        errs[name] = float(np.abs(got - whole)[ice].max() / max(np.abs(whole[ice]).max(), 1e-300))  # NaN if a box is missing
    errs["single"] = full
    return errs
