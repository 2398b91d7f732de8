"""Pins the oracle's mesh-derived integer state against the reference's 25km_NH fixture
(dynamics/test/ParametricMesh_test.cpp:39-131): land mask and sorted Dirichlet lists, bit-exact."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mesh_25km_NH.npz")


def golden_mesh():
    g = np.load(GOLD)
    nx, ny = int(g["nx"]), int(g["ny"])
    mask = np.unpackbits(g["landmask_bits"])[: nx * ny].astype(np.float64).reshape(ny, nx)
    x = np.arange(nx + 1) * float(g["dx"])
    y = np.arange(ny + 1) * float(g["dy"])
    X, Y = np.meshgrid(x, y)
    coords = np.ascontiguousarray(np.stack([X, Y], -1))
    lists = [np.asarray(g[f"dirichlet{e}"], dtype=np.int64) for e in range(4)]
    return nx, ny, coords, mask, lists


def test_reference_fixture_facts():
    """The literal REQUIREs of ParametricMesh_test.cpp:44-76."""
    nx, ny, coords, mask, lists = golden_mesh()
    assert (nx, ny) == (154, 121)
    lm = mask.ravel()
    assert lm[0] == 0 and lm[4] == 0 and lm[5] == 1 and lm[nx - 1] == 1 and lm[(ny - 1) * nx] == 0
    assert lists[0][0] == 5 and lists[3][0] == 5 and lists[1][0] == 7 and lists[2][0] == 48
    assert [len(l) for l in lists] == [369, 384, 369, 384]


def test_oracle_lists_match_smesh_file():
    """'Compare readmesh and landmask reading' (ParametricMesh_test.cpp:89-131) on the oracle."""
    nx, ny, coords, mask, lists = golden_mesh()
    z = np.zeros((ny, nx))
    o = oracle.OracleDynamics("mevp", 6, 2, 1)
    o.setData({"coords": coords, "mask": mask, "x": z, "y": z, "hice": z, "cice": z, "u": z, "v": z})
    assert np.array_equal(o.landmask(), mask.ravel().astype(np.uint8))
    for e in range(4):
        assert np.array_equal(o.dirichlet(e), lists[e]), f"edge {e}"
    assert np.array_equal(o.vertices(), coords.reshape(-1, 2))
