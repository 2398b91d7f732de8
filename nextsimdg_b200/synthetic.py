"""Deterministic synthetic inputs for the dynamics path; no files, no network.

Arrays follow the ModelArray conventions: element fields are (ny, nx) with [j, i] = element
i + nx*j, vertex coordinates are (ny+1, nx+1, 2).  Each generator states the reference file
whose *specification* it restates (the reference builds netCDF files; we build arrays).
"""
from __future__ import annotations

import numpy as np


def box_coords(nx: int, ny: int, Lx: float, Ly: float) -> np.ndarray:
    """Vertex coordinates x = i*Lx/nx, y = j*Ly/ny (run/make_init_base.py:133-142)."""
    x = np.arange(nx + 1, dtype=np.float64) * Lx / nx
    y = np.arange(ny + 1, dtype=np.float64) * Ly / ny
    X, Y = np.meshgrid(x, y)  # shape (ny+1, nx+1)
    return np.ascontiguousarray(np.stack([X, Y], axis=-1))


def distort_coords(coords: np.ndarray, distort: float = 0.05) -> np.ndarray:
    """The distorted test mesh of dynamics/test/Advection_test.cpp:212-218."""
    ny1, nx1, _ = coords.shape
    Nx, Ny = nx1 - 1, ny1 - 1
    Lx, Ly = coords[0, -1, 0] - coords[0, 0, 0], coords[-1, 0, 1] - coords[0, 0, 1]
    ix = np.arange(nx1)[None, :]
    iy = np.arange(ny1)[:, None]
    out = coords.copy()
    out[..., 0] += Lx * distort * np.sin(np.pi * ix / Nx * 3.0) * np.sin(np.pi * iy / Ny)
    out[..., 1] += Ly * distort * np.sin(np.pi * iy / Ny * 2.0) * np.sin(np.pi * ix / Nx * 2.0)
    return np.ascontiguousarray(out)


def benchmark_box(n: int, L: float = 512000.0, ring_mask: bool = True) -> dict:
    """The Mehlmann et al. (2021) cyclone box (run/make_init_benchmark.py:6-40, config_benchmark.cfg).

    mask = 1 inside, 0 on the outermost ring; cice = 1; hice = 0.3 + 0.005(sin(60e-6 a D)+sin(30e-6 b D))
    with a the slow and b the fast array index and D = L/n; damage = 1; u = v = 0.
    """
    coords = box_coords(n, n, L, L)
    mask = np.ones((n, n))
    if ring_mask:
        mask[0, :] = mask[-1, :] = 0.0
        mask[:, 0] = mask[:, -1] = 0.0
    D = L / n
    a = np.arange(n, dtype=np.float64)[:, None] * D
    b = np.arange(n, dtype=np.float64)[None, :] * D
    hice = (0.3 + 0.005 * (np.sin(60e-6 * a) + np.sin(30e-6 * b))) * mask
    cice = np.ones((n, n)) * mask
    zeros = np.zeros((n, n))
    return {
        "coords": coords, "mask": mask, "x": b + 0 * a, "y": a + 0 * b,
        "hice": np.ascontiguousarray(hice), "cice": np.ascontiguousarray(cice),
        "damage": np.ascontiguousarray(mask.copy()), "u": zeros.copy(), "v": zeros.copy(),
    }


def benchmark_forcing(n: int, t: float, L: float = 512000.0) -> dict:
    """BenchmarkAtmosphere.cpp:38-74 and BenchmarkOcean.cpp:27-36 at time t [s] since the first update.

    Coordinates are the cell lower-left corners x = i*dx, y = j*dy (BenchmarkCoordinates.cpp:20-42).
    """
    dx = L / n
    x = (np.arange(n, dtype=np.float64) * dx)[None, :] + np.zeros((n, 1))
    y = (np.arange(n, dtype=np.float64) * dx)[:, None] + np.zeros((1, n))
    x0 = y0 = (L / 2) * (1 + t / (5 * 86400.0))
    xp, yp = x - x0, y - y0
    s = 1e-5 * np.exp(-1e-5 * np.hypot(xp, yp))
    al = np.deg2rad(72.0)
    uw = -s * 30.0 * (np.cos(al) * xp + np.sin(al) * yp)
    vw = -s * 30.0 * (-np.sin(al) * xp + np.cos(al) * yp)
    uo = 0.01 * (2 * y / L - 1)
    vo = 0.01 * (1 - 2 * x / L)
    c = np.ascontiguousarray
    return {"uwind": c(uw), "vwind": c(vw), "uocean": c(uo), "vocean": c(vo), "ssh": np.zeros((n, n))}


def smooth_forcing(nx: int, ny: int, seed: int = 20241017, ssh_amp: float = 0.05) -> dict:
    """Smooth O(1) test forcing with a non-trivial sea-surface height (exercises the SSH gradient)."""
    x = (np.arange(nx) + 0.5)[None, :] / nx
    y = (np.arange(ny) + 0.5)[:, None] / ny
    c = np.ascontiguousarray
    return {
        "uwind": c(8.0 * np.sin(2 * np.pi * y) * np.cos(np.pi * x) + 2.0),
        "vwind": c(-6.0 * np.cos(2 * np.pi * x) * np.sin(np.pi * y) + 1.0),
        "uocean": c(0.05 * (2 * y - 1) + 0 * x),
        "vocean": c(0.05 * (1 - 2 * x) + 0 * y),
        "ssh": c(ssh_amp * np.sin(2 * np.pi * x) * np.sin(3 * np.pi * y)),
    }


def para_state(nx: int = 30, ny: int = 24, dxy: float = 25000.0, distort: float = 0.0, irregular_mask: bool = False,
               seed: int = 20241017) -> dict:
    """para24x30-shaped case (run/make_init_para24x30.py:10-11,37-38,83-90): 30 x 24 elements of 25 km,
    cice = 0.96875, hice = (1.5, 2.0, 3.0) in the first three DG components, u = v = 0; optionally the
    Advection_test distortion and an irregular land mask in the spirit of run/make_init_rect20x30.py:26-56."""
    coords = box_coords(nx, ny, nx * dxy, ny * dxy)
    if distort:
        coords = distort_coords(coords, distort)
    mask = np.ones((ny, nx))
    if irregular_mask:
        jj, ii = np.mgrid[0:ny, 0:nx]
        mask[(ii - 0.3 * nx) ** 2 + (jj - 0.6 * ny) ** 2 < (0.12 * min(nx, ny)) ** 2] = 0.0  # island
        mask[:2, : nx // 3] = 0.0  # coast strip touching the domain edge
        mask[ny // 2, nx - 3:] = 0.0  # one-element-wide peninsula
    hice = np.zeros((ny, nx, 6))
    hice[..., 0], hice[..., 1], hice[..., 2] = 1.5, 2.0, 3.0
    # keep the linear part small enough that h stays positive inside the element
    hice[..., 1] *= 0.1
    hice[..., 2] *= 0.1
    rng = np.random.default_rng(seed)
    hice[..., 0] += 0.2 * rng.uniform(-1, 1, size=(ny, nx))
    cice = np.zeros((ny, nx, 6))
    cice[..., 0] = 0.96875
    hice *= mask[..., None]
    cice *= mask[..., None]
    z = np.zeros((ny, nx))
    return {"coords": coords, "mask": mask, "x": z, "y": z, "hice": hice, "cice": cice, "u": z.copy(), "v": z.copy(),
            "damage": np.ascontiguousarray(mask.copy())}


def topaz_like_spherical(n: int = 128, land_lat: float = 72.0) -> dict:
    """TOPAZ-like polar azimuthal-equidistant grid (run/init_topaz128x128.py:106-121,152-154):
    X, Y = linspace(-20, 20, n+1) degrees, lat = 90 - sqrt(X^2+Y^2), lon = atan2(Y, X) [degrees];
    land where the element-centre latitude is south of `land_lat`."""
    g = np.linspace(-20.0, 20.0, n + 1)
    X, Y = np.meshgrid(g, g)
    lat = 90.0 - np.sqrt(X ** 2 + Y ** 2)
    lon = np.degrees(np.arctan2(Y, X))
    coords = np.ascontiguousarray(np.stack([lon, lat], axis=-1))
    gc = 0.5 * (g[1:] + g[:-1])
    Xc, Yc = np.meshgrid(gc, gc)
    latc = 90.0 - np.sqrt(Xc ** 2 + Yc ** 2)
    mask = (latc > land_lat).astype(np.float64)
    hice = (1.0 + 0.5 * np.cos(np.radians(Xc * 9)) * np.sin(np.radians(Yc * 9))) * mask
    cice = (0.9 + 0.05 * np.sin(np.radians(Xc * 18))) * mask
    z = np.zeros((n, n))
    return {"coords": coords, "mask": mask, "longitude": z, "latitude": z, "hice": np.ascontiguousarray(hice),
            "cice": np.ascontiguousarray(cice), "u": z.copy(), "v": z.copy(), "damage": np.ascontiguousarray(mask.copy())}
