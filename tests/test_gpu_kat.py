"""The reference's own advection known-answer tests, met BY THE CUDA PATH through the C ABI.

dynamics/test/Advection_test.cpp:243-308 (rotating bump on a straight and on a distorted box mesh, rk3) and
dynamics/test/AdvectionPeriodicBC_test.cpp:279-298 (ring mesh, periodic left/right edges, LimitMax/LimitMin after every
step) store L2 errors "taken" from upstream runs and check them to rel 1e-7.  Here the oracle only stages the problem
(mesh, Function2DG projections of the initial field and of the velocity) and measures the final error
(L2ErrorFunctionDG); every time step -- reinitnormalvelocity, the three Runge-Kutta stages with cell, edge, periodic and
Dirichlet terms, the limiters -- runs in libnsdg_cuda.so (nsdg_set_boundaries, nsdg_advect_field).
"""
import numpy as np
import pytest

from test_oracle_kat import ADVECTION, DISTORTED, PERIODIC, TOL

pytestmark = pytest.mark.gpu


def _handle(dg, st, kind):
    from nextsimdg_b200 import CUDAMEVPDynamics

    nx, ny = st["nx"], st["ny"]
    d = CUDAMEVPDynamics(dgadv=dg, cgdegree=2 if dg == 6 else 1, nsteps=1)
    z = np.zeros((ny, nx))
    d.setData({"coords": st["coords"], "mask": np.ones((ny, nx)), "x": z, "y": z, "hice": st["phi0"], "cice": z.copy(),
               "u": z.copy(), "v": z.copy()})
    if kind == 1:  # AdvectionPeriodicBC_test.cpp:228-249: Dirichlet bottom / top, periodic left / right
        d.set_boundaries(dirichlet=[np.arange(nx), [], nx * (ny - 1) + np.arange(nx), []],
                         periodic=[[(1, (i + 1) * nx - 1, i * nx, i * (nx + 1) + i) for i in range(ny)]])
    d.set_internal("velx", st["velx"])
    d.set_internal("vely", st["vely"])
    return d


@pytest.mark.parametrize("dg", [3, 6])
@pytest.mark.parametrize("it", [0, 1])
@pytest.mark.parametrize("distort,table", [(0.0, ADVECTION), (0.05, DISTORTED)])
def test_cuda_meets_the_advection_kat(cuda_lib, oracle_lib, dg, it, distort, table):
    import oracle

    st = oracle.kat_stage(0, dg, it, distort)
    d = _handle(dg, st, 0)
    d.advect_field("hice", st["dt"], rk_order=3, nsteps=st["nt"])  # Advection_test.cpp:131-135: rk3 for DG >= 3
    err = oracle.kat_error(0, dg, it, distort, d.getDGData("hice"))
    d.close()
    assert err == pytest.approx(table[dg][it], rel=TOL)


@pytest.mark.parametrize("dg", [3, 6])
@pytest.mark.parametrize("it", [0, 1])
def test_cuda_meets_the_periodic_ring_kat(cuda_lib, oracle_lib, dg, it):
    import oracle

    st = oracle.kat_stage(1, dg, it)
    d = _handle(dg, st, 1)
    d.advect_field("hice", st["dt"], rk_order=3, nsteps=st["nt"], limit_max=1.0, limit_min=0.0)
    err = oracle.kat_error(1, dg, it, 0.0, d.getDGData("hice"))
    d.close()
    assert err == pytest.approx(PERIODIC[dg][it], rel=TOL)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_rk_orders_match_the_oracle_step_by_step(cuda_lib, oracle_lib, order):
    """step_rk1 / rk2 / rk3 (DGTransport.cpp:514-566): ten steps on the distorted 12 x 13 box, field by field against the
    oracle's transport object at the per-sweep tolerance."""
    import ctypes

    import oracle

    st = oracle.kat_stage(0, 6, 0, 0.05)
    d = _handle(6, st, 0)
    d.advect_field("hice", st["dt"], rk_order=order, nsteps=10)
    got = d.getDGData("hice")
    d.close()
    L = oracle.load()
    L.nso_kat_steps.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    want = np.empty_like(got)
    assert L.nso_kat_steps(0, 6, 0, 0.05, order, 10, want.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-12


@pytest.mark.parametrize("rheo,seam,distort", [("mevp", "x", 0.03), ("bbm", "x", 0.03), ("mevp", "y", 0.0), ("mevp", "xy", 0.03)])
def test_periodic_seam_in_the_module_path(cuda_lib, rheo, seam, distort):
    """Periodic edges set on a module handle: the advection inside update() crosses the seam (DGTransport.cpp:466-481),
    prepareIteration averages cgH, cgA across it (CGDynamicsKernel.cpp:264-266) and every subcycle averages the stress
    divergence across it before the momentum update (CGDynamicsKernel.cpp:395-397; VectorManipulations.hpp:26-65) --
    compared with the reference's own kernels given the same lists.  seam: "x" left / right, "y" bottom / top (a uniform
    mesh: the handle leaves its fast path), "xy" both (corner nodes are averaged twice, segment after segment)."""
    import oracle
    from nextsimdg_b200 import CUDABBMDynamics, CUDAMEVPDynamics, synthetic

    nx, ny = 18, 11
    ms = synthetic.para_state(nx, ny, distort=distort)
    f = synthetic.smooth_forcing(nx, ny)
    per = []
    if "x" in seam:  # {type 1, right element, left element, edge id}
        per.append([(1, (i + 1) * nx - 1, i * nx, i * (nx + 1)) for i in range(ny)])
    if "y" in seam:  # {type 0, top element, bottom element, edge id}
        per.append([(0, (ny - 1) * nx + i, i, i) for i in range(nx)])
    impl = "reference" if oracle.have_ref(2) else "port"
    if impl == "reference":
        oracle.load_ref(2).nso_set_threads(1)
    if rheo == "bbm":
        ms["damage"] = 1.0 + 0.0 * np.asarray(ms["mask"])
    gpu = (CUDABBMDynamics if rheo == "bbm" else CUDAMEVPDynamics)(nsteps=3)
    ref = oracle.OracleDynamics(rheo, 6, 2, 3, impl=impl)
    for d in (gpu, ref):
        d.setData(ms)
        d.set_boundaries(dirichlet=[[] if "y" in seam else None, [] if "x" in seam else None, [] if "y" in seam else None,
                                    [] if "x" in seam else None], periodic=per)
        d.shared = {"hice": np.ascontiguousarray(ms["hice"][..., 0]), "cice": np.ascontiguousarray(ms["cice"][..., 0]),
                    **{k: v.copy() for k, v in f.items()}}
        if rheo == "bbm":
            d.shared["damage"] = ms["damage"].copy()
    # a velocity that crosses the seam, then one update: the advection uses it, prepareIteration and three subcycles follow
    # (not a uniform velocity: its strain would be rounding noise, which BBM's elastic predictor multiplies by 6e8; and a
    # short step for BBM, whose explicit scheme is unstable with three subcycles of 300 s)
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    u0 = 0.3 + 0.05 * np.cos(2 * np.pi * ii / nx) * np.cos(2 * np.pi * jj / ny) + 0.0 * ms["mask"]
    for d in (gpu, ref):
        d._set("u", u0)
        d._set("v", 0.1 * u0)
        d.update(30.0 if rheo == "bbm" else 900.0)
    for name in ("hice", "cice", "cgH", "cgA"):
        a, b = gpu.internal(name), ref.internal(name)
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-12, name
    # the momentum solve: three subcycles with the seam average of the stress divergence
    for name in ("cg_u", "cg_v"):
        a, b = gpu.internal(name), ref.internal(name)
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-11, name
    if "x" in seam and rheo == "mevp":  # the seam is really periodic: mass left through the right edge arrives at the left edge
        # (BBM advects with the mean velocity of the previous step's subcycles, which is zero in the first update)
        h = gpu.internal("hice").reshape(ny, nx, 6)[..., 0]
        assert np.abs(h[:, 0] - ms["hice"][:, 0, 0]).max() > 1e-6
    gpu.close()
