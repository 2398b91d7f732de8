"""Shared case table for the parity tests: the same seeded inputs are fed to the real reference kernels
(oracle/_ref, `impl="reference"`), to the CPU restatement (`impl="port"`) and to the CUDA library, and the
reference's outputs are also committed as tests/golden/ref_outputs.npz (made by tests/golden/make_golden_ref.py).

A case = (model state for setData, list of per-update forcings, dt, build (dgadv, cgdegree), nSteps per rheology).
"""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz")

# what the module exports after update() (MEVPDynamics.cpp:79-86, BBMDynamics.cpp:92-100) + the prognostic
# internals that carry over to the next step
EXPORTS = ("uice", "vice", "taux", "tauy", "hice", "cice")
INTERNALS = ("cg_u", "cg_v", "s11", "s12", "s22")


def _trim(ms, dg):
    ms = dict(ms)
    for k in ("hice", "cice"):
        a = np.asarray(ms[k])
        if a.ndim == 3 and a.shape[-1] > dg:
            ms[k] = np.ascontiguousarray(a[..., :dg]) if dg > 1 else np.ascontiguousarray(a[..., 0])
    return ms


def cases():
    from nextsimdg_b200 import synthetic as S

    c = {
        # name: (ms, forcings, dt, (dgadv, cg), {rheology: nSteps})
        "box32": (S.benchmark_box(32), [S.benchmark_forcing(32, 0.0), S.benchmark_forcing(32, 120.0)], 120.0, (6, 2),
                  {"mevp": 100, "bbm": 100, "freedrift": 1}),
        "para_distorted_land": (S.para_state(30, 24, distort=0.05, irregular_mask=True), [S.smooth_forcing(30, 24)], 900.0, (6, 2),
                                {"mevp": 100, "bbm": 100, "freedrift": 1}),
        "para_uniform_land": (S.para_state(37, 21, irregular_mask=True), [S.smooth_forcing(37, 21)], 900.0, (6, 2),
                              {"mevp": 100, "bbm": 100}),
        # BBM on a spherical mesh: the reference's damage time scale uses smesh.h(i) in radians
        # (BBMStressUpdateStep.hpp:158) and blows up within a few subcycles; compared while still finite
        "spherical": (S.topaz_like_spherical(32), [S.smooth_forcing(32, 32)], 600.0, (6, 2), {"mevp": 100, "bbm": 2}),
        "dg1cg1_distorted_land": (_trim(S.para_state(30, 24, distort=0.05, irregular_mask=True), 3), [S.smooth_forcing(30, 24)], 900.0,
                                  (3, 1), {"mevp": 100, "bbm": 100}),
        "dg1cg2_distorted_land": (_trim(S.para_state(22, 18, distort=0.04, irregular_mask=True), 3), [S.smooth_forcing(22, 18)], 900.0,
                                  (3, 2), {"mevp": 30, "bbm": 30}),
        "dg0cg2_uniform": (_trim(S.para_state(20, 16), 1), [S.smooth_forcing(20, 16)], 900.0, (1, 2), {"mevp": 30}),
    }
    return c


def inputs_digest(ms, forcings):
    h = hashlib.sha256()
    for k in sorted(ms):
        h.update(k.encode())
        h.update(np.ascontiguousarray(ms[k], dtype=np.float64).tobytes())
    for f in forcings:
        for k in sorted(f):
            h.update(np.ascontiguousarray(f[k], dtype=np.float64).tobytes())
    return h.hexdigest()


def run_case(d, ms, forcings, dt):
    """Drive a dynamics object (CUDA module mirror or oracle.OracleDynamics) the way the model does."""
    ny, nx = np.asarray(ms["mask"]).shape
    d.setData(ms)
    d.shared = {"hice": np.array(np.asarray(ms["hice"]).reshape(ny, nx, -1)[..., 0], dtype=np.float64, order="C", copy=True),
                "cice": np.array(np.asarray(ms["cice"]).reshape(ny, nx, -1)[..., 0], dtype=np.float64, order="C", copy=True)}
    for f in forcings:
        d.shared.update({k: v.copy() for k, v in f.items()})
        d.update(dt)
    out = {"uice": d.uice, "vice": d.vice, "taux": d.taux, "tauy": d.tauy, "hice": d.shared["hice"], "cice": d.shared["cice"]}
    if getattr(d, "rheology", "") == "bbm" or getattr(d, "_uses_damage", False):
        out["damage"] = d.damage
    if getattr(d, "rheology", "") in ("freedrift", 2):
        # FreeDriftDynamics::update never exports the ice-ocean stress (FreeDriftDynamics.hpp:39-58)
        out.pop("taux"), out.pop("tauy")
    for n in INTERNALS:
        out[n] = d.internal(n)
    return {k: np.array(v, dtype=np.float64, copy=True) for k, v in out.items() if v is not None}


def compare(got, want, mask, tol, tol_stress=None):
    """Norm-wise relative error per field over ICE elements (land values are unspecified in the reference)."""
    ice = np.asarray(mask).astype(bool)
    worst = {}
    for k, w in want.items():
        if k not in got:
            continue
        g = np.asarray(got[k], dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
        if k in ("cg_u", "cg_v"):
            pass
        elif g.size % ice.size == 0:
            g = g.reshape(ice.size, -1)[ice.ravel()]
            w = w.reshape(ice.size, -1)[ice.ravel()]
        assert np.isfinite(g).all(), f"{k}: non-finite values"
        worst[k] = np.abs(g - w).max() / max(np.abs(w).max(), 1e-300)
    bad = {k: v for k, v in worst.items() if v > (tol_stress if (tol_stress and k in ("s11", "s12", "s22")) else tol)}
    return worst, bad
