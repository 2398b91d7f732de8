import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, synthetic
nx, ny, dt = 36, 28, 900.0
ms = synthetic.para_state(nx, ny, distort=0.03, irregular_mask=True)
f = synthetic.smooth_forcing(nx, ny)
names = ("hice","cice","cgH","cgA","uAtmos","vAtmos","uOcean","vOcean","uGradSSH","s11","s12","cg_u","cg_v")
for rheo, cls in (("mevp", CUDAMEVPDynamics), ("bbm", CUDABBMDynamics)):
    res = {}
    for second in ("update", "step", "update2"):
        d = cls(nsteps=40, keep_dg_moments=True)
        d.setData(ms)
        d.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0]), "cice": np.ascontiguousarray(ms["cice"][..., 0]), **{k: v.copy() for k, v in f.items()}}
        d.update(dt)
        if second.startswith("update"):
            d.update(dt)
        else:
            d.step(dt)
        res[second] = {n: d.internal(n) for n in names}
        d.close()
    for n in names:
        a, b, c = res["update"][n], res["step"][n], res["update2"][n]
        print(rheo, n, "update-vs-step %.2e" % (np.abs(a-b).max()/max(np.abs(b).max(),1e-300)), "update-vs-update %.2e" % (np.abs(a-c).max()/max(np.abs(c).max(),1e-300)))
