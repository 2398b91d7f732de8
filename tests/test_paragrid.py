"""ParaGrid restart layout (core/src/ParaGridIO.cpp:272-348): CDL writer / reader on the CPU, bit-reproducible resume through it on the GPU."""
import numpy as np
import pytest

from nextsimdg_b200 import paragrid


def _fake_state(nx, ny, dg=6, cg=2, bbm=True, seed=3):
    rng = np.random.default_rng(seed)
    d = paragrid.dimensions(nx, ny, dg, cg)
    names = [k for k in paragrid.DIMS if bbm or k not in ("damage", "dynamics_avg_u", "dynamics_avg_v")]
    return {k: rng.standard_normal([d[x] for x in paragrid.DIMS[k]]) * 10.0 ** rng.integers(-12, 6) for k in names}


@pytest.mark.parametrize("dg,cg", [(6, 2), (3, 1)])
def test_cdl_round_trip_is_exact_and_has_the_reference_layout(dg, cg):
    nx, ny = 7, 5
    st = _fake_state(nx, ny, dg, cg)
    text = paragrid.to_cdl(st, nx, ny, dg, cg, time_unix=946684800, time_formatted="2000-01-01T00:00:00Z")
    # the layout of ParaGridIO::dumpModelState / run/make_init_base.py:112-186
    for needle in ("group: structure {", ':type = "parametric_rectangular"', "group: metadata {", "group: time {",
                   "int64 time ;", "group: configuration {", "group: data {", f"xdim = {nx} ;", f"ydim = {ny} ;",
                   f"x_cg = {cg * nx + 1} ;", f"dg_comp = {dg} ;", f"dgstress_comp = {8 if cg == 2 else 3} ;", "ncoords = 2 ;",
                   "double hice(ydim, xdim, dg_comp) ;", "double u(ydim, xdim) ;", "double coords(yvertex, xvertex, ncoords) ;",
                   "hice:missing_value = 1.7e+38 ;"):
        assert needle in text, needle
    back, dims = paragrid.from_cdl(text)
    assert dims == paragrid.dimensions(nx, ny, dg, cg)
    assert set(back) == set(st)
    for k in st:
        assert np.array_equal(back[k], st[k]), k  # every double survives the text form bit for bit
    with pytest.raises(ValueError):
        bad = dict(st)
        bad["hice"] = bad["hice"][..., :2]
        paragrid.to_cdl(bad, nx, ny, dg, cg)


@pytest.mark.gpu
@pytest.mark.parametrize("rheo", ["mevp", "bbm"])
def test_resume_through_the_paragrid_file_is_bitwise(rheo, cuda_lib):
    from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, synthetic

    cls = CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics
    nx, ny, dt = 24, 18, 900.0
    ms = synthetic.para_state(nx, ny, distort=0.04, irregular_mask=True)
    f = [synthetic.smooth_forcing(nx, ny, seed=s) for s in (1, 2)]

    def fresh():
        d = cls(nsteps=40)
        d.setData(ms)
        d.shared = {"hice": np.array(ms["hice"][..., 0], order="C", copy=True), "cice": np.array(ms["cice"][..., 0], order="C", copy=True)}
        return d

    a = fresh()
    for k in range(2):
        a.shared.update({n: v.copy() for n, v in f[k].items()})
        a.update(dt)
    b = fresh()
    b.shared.update({n: v.copy() for n, v in f[0].items()})
    b.update(dt)
    text = paragrid.to_cdl(paragrid.restart_state(b, ms), nx, ny)
    shared = {k: v.copy() for k, v in b.shared.items()}
    dmg = None if b.damage is None else b.damage.copy()
    b.close()
    state, dims = paragrid.from_cdl(text)
    assert dims["xdim"] == nx and state["hice"].shape == (ny, nx, 6) and state["dynamics_cg_u"].shape == (2 * ny + 1, 2 * nx + 1)
    c = fresh()
    paragrid.restore(c, state)
    c.shared.update(shared)
    if rheo == "bbm":
        c.damage = dmg
    c.shared.update({n: v.copy() for n, v in f[1].items()})
    c.update(dt)
    for name in ("uice", "vice", "taux", "tauy"):
        assert np.array_equal(getattr(a, name), getattr(c, name)), name
    for name in ("s11", "cg_u", "hice", "cice"):
        assert np.array_equal(a.internal(name), c.internal(name)), name
