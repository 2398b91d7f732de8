"""Stress version of tests/test_gpu_parity.py::test_restart_state_resumes_bitwise (experiment tooling): loops inside one
process, churning device memory between iterations, and reports which field differs when a resume is not bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, capi, synthetic

rheo = os.environ.get("QB_RHEO", "mevp"); iters = int(os.environ.get("ITERS", "40"))
cls = CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics
nx, ny, dt = 40, 33, 900.0
ms = synthetic.para_state(nx, ny, distort=0.04, irregular_mask=True)
f = [synthetic.smooth_forcing(nx, ny, seed=s) for s in (1, 2)]

def fresh():
    d = cls(nsteps=60); d.setData(ms)
    d.shared = {"hice": np.array(ms["hice"][..., 0], order="C", copy=True), "cice": np.array(ms["cice"][..., 0], order="C", copy=True)}
    return d

def churn(i):  # leave differently-sized garbage behind in the device heap
    n = 24 + 7 * (i % 5)
    d = (CUDABBMDynamics if i % 2 else CUDAMEVPDynamics)(nsteps=3)
    m = synthetic.benchmark_box(n); d.setData(m)
    d.shared = {"hice": m["hice"].copy(), "cice": m["cice"].copy(), **{a: b.copy() for a, b in synthetic.benchmark_forcing(n, 0.0).items()}}
    d.update(120.0); d.close()

bad = 0
for it in range(iters):
    churn(it)
    a = fresh()
    for k in range(2):
        a.shared.update({n: v.copy() for n, v in f[k].items()}); a.update(dt)
    b = fresh(); b.shared.update({n: v.copy() for n, v in f[0].items()}); b.update(dt)
    state, shared = b.get_state(), {k: v.copy() for k, v in b.shared.items()}
    dmg = None if b.damage is None else b.damage.copy(); b.close()
    churn(it + 1)
    c = fresh(); c.set_state(state); c.shared.update(shared)
    if rheo == "bbm": c.damage = dmg
    c.shared.update({n: v.copy() for n, v in f[1].items()}); c.update(dt)
    diffs = [n for n in ("uice", "vice", "taux", "tauy") if not np.array_equal(getattr(a, n), getattr(c, n))]
    for n in ("s11", "s12", "s22", "cg_u", "cg_v", "hice", "cice"):
        x, y = a.internal(n), c.internal(n)
        if not np.array_equal(x, y):
            w = np.flatnonzero(x != y)
            diffs.append(f"{n}[{w.size} of {x.size}, first {w[:4].tolist()}, max {np.nanmax(np.abs(x - y)):.2e}]")
    if diffs:
        bad += 1; print("iter", it, "DIFF", diffs, flush=True)
    a.close(); c.close()
print("STRESS", os.path.basename(capi.library_path()), rheo, f"{bad} of {iters} iterations not bitwise")
