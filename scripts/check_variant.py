"""Quick parity check of one library build (NSDG_CUDA_LIB) against the oracle: two updates on a 64x64 benchmark box.
Experiment tooling for kernel variants; the tests proper are tests/test_gpu_*.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, capi, synthetic

n, dt, nsteps = 64, 120.0, 20
worst = 0.0
for cls, rheo in ((CUDAMEVPDynamics, "mevp"), (CUDABBMDynamics, "bbm")):
    ms = synthetic.benchmark_box(n)
    if os.environ.get("QB_DISTORT"):  # parametric kernels
        ms["coords"] = synthetic.distort_coords(ms["coords"], 0.02)
    gpu = cls(nsteps=nsteps)
    ref = oracle.OracleDynamics(rheo, nsteps=nsteps, impl="port")
    gpu.setData(ms); ref.setData(ms)
    for k in range(2):
        f = synthetic.benchmark_forcing(n, k * dt)
        for d in (gpu, ref):
            d.shared = {"hice": ms["hice"].copy(), "cice": ms["cice"].copy(), **{a: b.copy() for a, b in f.items()}}
            d.update(dt)
    names = [("u", gpu.uice, ref.uice), ("v", gpu.vice, ref.vice)]
    if rheo == "bbm":
        names.append(("damage", gpu.damage, ref.damage))
    for name, a, b in names:
        err = float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))
        worst = max(worst, err)
        print(f"{os.path.basename(capi.library_path())} {rheo} {name}: {err:.3e}")
    gpu.close()
print("CHECK", os.path.basename(capi.library_path()), "OK" if worst < 1e-9 else "FAIL", f"{worst:.3e}")
