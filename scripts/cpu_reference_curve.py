"""Throughput of the reference's own CPU dynamics (oracle/_ref) on the box's host cores at 512^2, 1024^2 and 2048^2:
is the per-element-subcycle cost size independent, i.e. is bench.py's bounded --impl reference sample representative of the
2048^2 workload?  One warm-up update and one timed update per size.  Writes gpurun_out/cpu_reference_curve.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

out = {"cores": os.cpu_count(), "points": []}
for rheo in ("mevp", "bbm"):
    for n in (512, 1024, 2048):
        if rheo == "bbm" and n == 2048:
            continue  # keeps the run bounded; the mEVP curve answers the question
        r = bench.cpu_oracle_run(n, rheo, 100, 1, 1)
        out["points"].append({"rheology": rheo, "n": n, "value": r["value"], "seconds_per_update": r["seconds"], "kind": r["kind"]})
        print(out["points"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_reference_curve.json"), "w"), indent=1)
