"""Generates tests/golden/mesh_25km_NH.npz from the reference's own mesh fixture.

Run in the build container (needs /root/reference); the .npz is committed so that nothing reads
/root/reference at test time.  Source: dynamics/test/25km_NH.smesh (format: ParametricMesh.cpp:8-190
`readmesh`), the fixture behind dynamics/test/ParametricMesh_test.cpp:39-131.

Stored: nx, ny, dx, dy, the land mask (1 = ocean/ice) as packed bits, and the four Dirichlet element
lists exactly as the file gives them (the reference test sorts nothing on the readmesh side; the
file is already sorted per edge).
"""
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dynamics/test/25km_NH.smesh"
tok = open(src).read().split()
assert tok[0] == "ParametricMesh" and tok[1] == "2.0"
nx, ny = int(tok[2]), int(tok[3])
p = 4
nn = (nx + 1) * (ny + 1)
verts = np.array(tok[p:p + 2 * nn], dtype=np.float64).reshape(nn, 2)
p += 2 * nn
assert tok[p] == "landmask"
ne = int(tok[p + 1])
p += 2
assert ne == nx * ny
mask = np.array(tok[p:p + ne], dtype=np.int64).astype(np.uint8)
p += ne
assert tok[p] == "dirichlet"
nd = int(tok[p + 1])
p += 2
d = np.array(tok[p:p + 2 * nd], dtype=np.int64).reshape(nd, 2)
p += 2 * nd
assert tok[p] == "periodic" and int(tok[p + 1]) == 0
dx = verts[1, 0] - verts[0, 0]
dy = verts[nx + 1, 1] - verts[0, 1]
# the vertices are a regular lattice: store only dx, dy after checking
X, Y = np.meshgrid(np.arange(nx + 1) * dx, np.arange(ny + 1) * dy)
assert np.array_equal(verts[:, 0], X.ravel()) and np.array_equal(verts[:, 1], Y.ravel())
out = {"nx": nx, "ny": ny, "dx": dx, "dy": dy, "landmask_bits": np.packbits(mask)}
for e in range(4):
    out[f"dirichlet{e}"] = d[d[:, 1] == e, 0]  # file order
np.savez_compressed(__file__.replace("make_golden.py", "mesh_25km_NH.npz"), **out)
print(nx, ny, dx, dy, mask.sum(), [len(out[f"dirichlet{e}"]) for e in range(4)])
