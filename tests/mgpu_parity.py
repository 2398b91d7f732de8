"""Multi-GPU parity driver (run under torch.distributed.run with 2, 4 or 8 ranks, one GPU each).

Every rank runs its box of a partitioned domain with NVLink halo exchange; rank 0 also runs the same
problem as a single domain on its GPU and compares the gathered owned boxes against it
(nextsimdg_b200.partition.partition_probe: mEVP and BBM on uniform and distorted Cartesian meshes with an
irregular land mask, mEVP on the TOPAZ-like spherical mesh).  The reference has no distributed dynamics, so
"parity" is equality (to rounding) with the single-domain result; rank 0 additionally checks the single-domain
result against the reference's own kernels (oracle/_ref) when that library travelled with the snapshot, which
ties the boxes to the reference.  With NSDG_MGPU_LOG set, the report is also written to that file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL = 1e-10


def reference_check(rheo, kind, single, nsteps, nts):
    """single-domain GPU exports vs the reference's kernels run on the host (None if oracle/_ref is absent)"""
    import oracle
    from nextsimdg_b200.partition import probe_inputs

    if not oracle.have_ref(2):
        return None
    oracle.load_ref(2).nso_set_threads(1)  # ssh != 0: the reference's CG1 -> CG2 interpolation is racy with threads (Q15)
    ms, forc = probe_inputs(kind)
    ref = oracle.OracleDynamics(rheo, 6, 2, nsteps, impl="reference")
    ref.setData(ms)
    a = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64)[..., 0] if np.asarray(x).ndim == 3 else x)  # noqa: E731
    ref.shared = {"hice": a(ms["hice"]).copy(), "cice": a(ms["cice"]).copy(), **{k: v.copy() for k, v in forc.items()}}
    for _ in range(nts):
        ref.update(120.0 if rheo == "bbm" else 600.0)
    ice = np.asarray(ms["mask"]).astype(bool)
    worst = 0.0
    for name, want in (("u", ref.uice), ("v", ref.vice), ("hice", ref.shared["hice"]), ("cice", ref.shared["cice"]), ("taux", ref.taux)):
        worst = max(worst, float(np.abs(single[name] - want)[ice].max() / max(np.abs(want[ice]).max(), 1e-300)))
    return worst


def main():
    import torch
    import torch.distributed as dist

    from nextsimdg_b200.partition import PROBE_CASES, partition_probe, torch_all_gather

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gather = torch_all_gather(dist)
    failures, report = [], []

    def say(line):
        print(line, flush=True)
        report.append(line)

    nsteps, nts = 40, 2
    # + BASELINE.json configs[3]: the TOPAZ-like 128 x 128 spherical grid, 120 subcycles ("100+"), 2-D partitioned
    for rheo, kind in list(PROBE_CASES) + [("mevp", "topaz128")]:
        nsteps, nts = (120, 1) if kind == "topaz128" else (40, 2)
        errs = partition_probe(rheo, kind, rank, world, local, gather, nsteps=nsteps, nts=nts)
        if rank == 0:
            single = errs.pop("single")
            for name, err in errs.items():
                ok = err < TOL  # False for NaN
                say(f"mgpu[{world}] {rheo:4s} {kind:9s} {name:6s} rel err vs single domain = {err:.3e} {'ok' if ok else 'FAIL'}")
                if not ok:
                    failures.append((rheo, kind, name, err))
            r = reference_check(rheo, kind, single, nsteps, nts)
            if r is not None:
                ok = r < 1e-9
                say(f"mgpu[{world}] {rheo:4s} {kind:9s} single domain vs the reference's kernels (oracle/_ref) = {r:.3e} {'ok' if ok else 'FAIL'}")
                if not ok:
                    failures.append((rheo, kind, "reference", r))
        dist.barrier()
    # a re-mesh on live, connected boxes: nsdg_set_mesh disconnects the halos, every box exports / connects / readies again
    # (epochs restart at zero on all of them) and the rerun from the same initial data reproduces the first run bit for bit
    from nextsimdg_b200 import CUDAMEVPDynamics
    from nextsimdg_b200.partition import Partition, connect_halos, probe_inputs

    ms, forc = probe_inputs("uniform")
    gny, gnx = ms["mask"].shape
    part = Partition.strong(rank, world, gnx, gny)
    dyn = CUDAMEVPDynamics(nsteps=40, device=local, partition=part)
    runs = []
    for attempt in range(2):
        dyn.setData(part.crop_state(ms))
        connect_halos(dyn, part, gather)
        lw = part.local_window()
        dyn.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0][lw]), "cice": np.ascontiguousarray(ms["cice"][..., 0][lw]),
                      **{k: np.ascontiguousarray(v[lw]) for k, v in forc.items()}}
        dyn.update(600.0)
        runs.append((dyn.uice.copy(), dyn.vice.copy(), dyn.shared["hice"].copy()))
    same = all(np.array_equal(a, b) for a, b in zip(*runs)) and float(np.abs(runs[0][0]).max()) > 0
    flags = gather(bool(same))
    dyn.close()
    if rank == 0:
        say(f"mgpu[{world}] re-mesh + reconnect on live boxes reproduces the first run bitwise: {'ok' if all(flags) else 'FAIL'}")
        if not all(flags):
            failures.append(("remesh", flags))
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    if rank == 0:
        say("MGPU PARITY " + ("OK" if not failures else f"FAILED {failures}"))
        log = os.environ.get("NSDG_MGPU_LOG")
        if log:
            os.makedirs(os.path.dirname(os.path.abspath(log)), exist_ok=True)
            with open(log, "w") as f:
                f.write("\n".join(report) + "\n")
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
