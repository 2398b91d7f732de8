"""Generate tests/golden/ref_outputs.npz: outputs of the REFERENCE's own dynamics kernels
(oracle/_ref/libnsdg_ref_cg{1,2}.so = /root/reference/dynamics/src compiled by `make -C oracle ref`) on the seeded
cases of tests/refcases.py.  Run in the build container (needs /root/reference): python tests/golden/make_golden_ref.py

The reference is run single-threaded: its CG1->CG2 sea-surface-height interpolation increments a loop counter that is
shared between OpenMP threads (CGDynamicsKernel.cpp:213-221), so multi-threaded runs with ssh != 0 are racy.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import refcases  # noqa: E402


def main():
    assert oracle.build_ref(), "reference tree not available: cannot regenerate the fixtures"
    out = {}
    for name, (ms, forcings, dt, (dg, cg), rheos) in refcases.cases().items():
        oracle.load_ref(cg).nso_set_threads(1)
        out[f"{name}/digest"] = np.frombuffer(bytes.fromhex(refcases.inputs_digest(ms, forcings)), dtype=np.uint8)
        for rheo, nsteps in rheos.items():
            d = oracle.OracleDynamics(rheo, dg, cg, nsteps, impl="reference")
            res = refcases.run_case(d, ms, forcings, dt, keep=refcases.KEEP.get(name, "all"))
            for k, v in res.items():
                out[f"{name}/{rheo}/{k}"] = v
            print(name, rheo, {k: float(np.abs(v).max()) for k, v in res.items() if k in ("uice", "hice")})
    np.savez_compressed(refcases.GOLDEN, **out)
    print("wrote", refcases.GOLDEN, os.path.getsize(refcases.GOLDEN), "bytes")


if __name__ == "__main__":
    main()
