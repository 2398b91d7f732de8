import os, sys, subprocess, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from nextsimdg_b200 import CUDABBMDynamics, synthetic
    ms = synthetic.para_state(96, 64, dxy=8000.0, distort=0.04, irregular_mask=True)
    forc = synthetic.smooth_forcing(96, 64)
    d = CUDABBMDynamics(nsteps=40); d.setData(ms)
    d.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0]), "cice": np.ascontiguousarray(ms["cice"][..., 0]), **{k: v.copy() for k, v in forc.items()}}
    for _ in range(2): d.update(600.0)
    pickle.dump((d.uice, d.vice, d.damage), open(sys.argv[1], "wb"))
else:
    outs = []
    for R in ("16", "7"):
        env = dict(os.environ, NSDG_STRIP_ROWS=R)
        subprocess.run([sys.executable, __file__, f"/tmp/o{R}.pkl"], env=env, check=True)
        outs.append(pickle.load(open(f"/tmp/o{R}.pkl", "rb")))
    for i, n in enumerate(("u", "v", "damage")):
        a, b = outs[0][i], outs[1][i]
        print(n, np.abs(a - b).max() / np.abs(a).max())
