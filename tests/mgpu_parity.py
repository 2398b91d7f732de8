"""Multi-GPU parity driver (run under torch.distributed.run with 2, 4 or 8 ranks, one GPU each).

Every rank runs its box of a partitioned domain with NVLink halo exchange; rank 0 also runs the same
problem as a single domain on its GPU and compares the gathered owned boxes against it.  The reference
has no distributed dynamics, so "parity" is equality (to rounding) with the single-domain result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, synthetic
    from nextsimdg_b200.partition import Partition, connect_halos, torch_all_gather

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gather = torch_all_gather(dist)
    failures = []
    cases = [("mevp", "uniform"), ("mevp", "distorted"), ("bbm", "uniform"), ("bbm", "distorted")]
    for rheo, kind in cases:
        gnx, gny, nsteps, nts = 96, 64, 40, 2
        # BBM is an explicit elastic scheme: keep the sub-step c_elastic * deltaT / dx well below 1 (here 0.3), otherwise
        # rounding noise is amplified and no two summation orders agree ('conditioning')
        dt = 120.0 if rheo == "bbm" else 600.0
        ms = synthetic.para_state(gnx, gny, dxy=8000.0, distort=0.04 if kind == "distorted" else 0.0, irregular_mask=True)
        forc = synthetic.smooth_forcing(gnx, gny)
        cls = CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics
        part = Partition.strong(rank, world, gnx, gny)
        dyn = cls(nsteps=nsteps, device=local, partition=part)
        dyn.setData(part.crop_state(ms))
        connect_halos(dyn, part, gather)
        lw = part.local_window()
        dyn.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0][lw]), "cice": np.ascontiguousarray(ms["cice"][..., 0][lw]),
                      **{k: np.ascontiguousarray(v[lw]) for k, v in forc.items()}}
        for _ in range(nts):
            dyn.update(dt)
        ow = part.owned_in_local()
        mine = {"u": dyn.uice[ow], "v": dyn.vice[ow], "hice": dyn.shared["hice"][ow], "cice": dyn.shared["cice"][ow],
                "taux": dyn.taux[ow]}
        if rheo == "bbm":
            mine["damage"] = dyn.damage[ow]
        everyone = gather((part.owned_window(), mine))
        dyn.close()
        if rank == 0:
            ref = cls(nsteps=nsteps, device=local)
            ref.setData(ms)
            ref.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0]), "cice": np.ascontiguousarray(ms["cice"][..., 0]),
                          **{k: v.copy() for k, v in forc.items()}}
            for _ in range(nts):
                ref.update(dt)
            full = {"u": ref.uice, "v": ref.vice, "hice": ref.shared["hice"], "cice": ref.shared["cice"], "taux": ref.taux}
            if rheo == "bbm":
                full["damage"] = ref.damage
            ice = ms["mask"].astype(bool)
            for name, whole in full.items():
                got = np.full_like(whole, np.nan)
                for win, fields in everyone:
                    got[win] = fields[name]
                err = np.abs(got - whole)[ice].max() / max(np.abs(whole[ice]).max(), 1e-300)
                status = "ok" if err < 1e-10 else "FAIL"
                print(f"mgpu[{world}] {rheo:4s} {kind:9s} {name:6s} rel err vs single domain = {err:.3e} {status}", flush=True)
                if not err < 1e-10:
                    failures.append((rheo, kind, name, err))
            ref.close()
        dist.barrier()
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU PARITY", "OK" if not failures else f"FAILED {failures}", flush=True)
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
