import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nextsimdg_b200 import CUDAMEVPDynamics, CUDABBMDynamics, capi, synthetic
n = int(os.environ.get("QB_N", "2048")); rheo = os.environ.get("QB_RHEO", "mevp")
L = 4000.0 * n
ms = synthetic.benchmark_box(n, L=L); f = synthetic.benchmark_forcing(n, 0.0, L=L)
if os.environ.get("QB_DISTORT"):
    ms["coords"] = synthetic.distort_coords(ms["coords"], 0.02)
if os.environ.get("QB_SPH"):  # TOPAZ-like spherical grid of the same size (polar azimuthal-equidistant, all ocean)
    ms = synthetic.topaz_like_spherical(n, land_lat=0.0)
    f = synthetic.smooth_forcing(n, n)
    ms["hice"] = np.ascontiguousarray(np.asarray(ms["hice"]).reshape(n, n, -1)[..., 0])
    ms["cice"] = np.ascontiguousarray(np.asarray(ms["cice"]).reshape(n, n, -1)[..., 0])
kw = {"dgadv": 3, "cgdegree": 1} if os.environ.get("QB_CG1") else {}  # the reference's other compile-time build (DG1 / CG1)
d = (CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics)(nsteps=100, **kw)
d.setData(ms)
d.shared = {"hice": ms["hice"].copy(), "cice": ms["cice"].copy(), **{k: v.copy() for k, v in f.items()}}
d.update(120.0)
d.step(120.0)
s = ctypes.c_float(); l = ctypes.c_float()
capi.check(d._lib.nsdg_time_kernels(d._h, 30, ctypes.byref(s), ctypes.byref(l)))
d.step(120.0); t = d.timing()
print(json.dumps({"lib": os.path.basename(capi.library_path()), "R": os.environ.get("NSDG_STRIP_ROWS"), "rheo": rheo, "n": n,
                  "strip_ms": round(s.value, 4), "lines_ms": round(l.value, 4), "subcycle_ms_per": round(t.subcycle_ms / 100, 4),
                  "adv_ms": round(t.advection_ms, 3), "prep_ms": round(t.prepare_ms, 3), "uniform": t.uniform_path, "mesh": "spherical" if os.environ.get("QB_SPH") else ("distorted" if os.environ.get("QB_DISTORT") else "rect")}))
