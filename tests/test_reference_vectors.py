"""Known answers the REFERENCE's own unit tests hold for the two helpers next to the dynamics path:

* physics/test/DamageHealing_test.cpp:33-157 -- Nextsim::ConstantHealing with td = 20 days, dt = 1 day: seven (cice, deltaCi,
  damage) -> damage vectors, checked upstream to 1e-8;
* physics/test/BenchmarkBoundaries_test.cpp:20-84 -- BenchmarkOcean at 256 x 256: four corner values, exact; BenchmarkAtmosphere:
  the cyclone weakens at point (50, 40) over the first hour.

They pin the checkers (oracle/healing.py, nextsimdg_b200.synthetic.benchmark_forcing) on the CPU and the device
implementations (nsdg_heal_damage, nsdg_set_benchmark_forcing) on the GPU.
"""
import numpy as np
import pytest

TD = 20 * 86400.0  # DamageHealing_test.cpp:34-35: "td = 20" days
DT = 86400.0  # Duration("P0-1T00:00:00")
# (cice, deltaCi, damage before, damage after): DamageHealing_test.cpp:75-81 and :129-157
HEALING = [(0.5, 0.0, 0.5, 0.55), (0.5, 0.0, 0.99, 1.0), (0.6, 0.3, 0.0, 0.55), (0.6, 0.3, 0.5, 0.80), (0.5, 0.1, 0.5, 0.65),
           (1.0, 0.1, 1.0, 1.0), (0.5, -0.5, 0.5, 0.55)]
PREC = 1e-8  # DamageHealing_test.cpp:73


def test_healing_checker_meets_the_reference_vectors():
    from oracle.healing import constant_healing

    for cice, dci, d0, want in HEALING:
        got = constant_healing(np.array([d0]), np.array([cice]), np.array([dci]), DT, TD)[0]
        assert got == pytest.approx(want, rel=PREC) and got <= 1.0


def test_benchmark_forcing_checker_meets_the_reference_vectors():
    from nextsimdg_b200 import synthetic

    n, vmax = 256, 0.01  # BenchmarkBoundaries_test.cpp:23-24,41
    f0 = synthetic.benchmark_forcing(n, 0.0)
    uo, vo = f0["uocean"], f0["vocean"]  # arrays are [j, i]
    assert uo[0, 0] == -vmax and vo[0, 0] == vmax  # :42-43
    assert uo[n - 1, n - 1] == pytest.approx((n - 2.0) / n * vmax, rel=1e-15)  # :44
    assert vo[n - 1, n - 1] == pytest.approx(-(n - 2.0) / n * vmax, rel=1e-15)  # :45
    assert uo[40, 50] != 0.0 and vo[40, 50] != 0.0  # :38-39
    # the cyclone moves away from (i, j) = (50, 40): the wind there weakens over the first hour (:70-79)
    f1 = synthetic.benchmark_forcing(n, 3600.0)
    assert f0["uwind"][40, 50] != 0.0 and f0["vwind"][40, 50] != 0.0
    assert abs(f1["uwind"][40, 50]) < abs(f0["uwind"][40, 50]) and abs(f1["vwind"][40, 50]) < abs(f0["vwind"][40, 50])


@pytest.mark.gpu
def test_device_healing_meets_the_reference_vectors(cuda_lib):
    from nextsimdg_b200 import CUDABBMDynamics, synthetic

    n = 8
    ms = synthetic.benchmark_box(n, ring_mask=False)
    for cice, dci, d0, want in HEALING:
        d = CUDABBMDynamics(nsteps=1)
        ms["cice"] = np.full((n, n), cice)
        ms["damage"] = np.full((n, n), d0)
        d.setData(ms)
        d.heal_damage(DT, TD, np.full((n, n), dci))
        got = d.getDG0Data("damage")
        d.close()
        assert np.all(got <= 1.0)
        assert got == pytest.approx(np.full((n, n), want), rel=PREC)


@pytest.mark.gpu
def test_device_benchmark_forcing_meets_the_reference_vectors(cuda_lib):
    from nextsimdg_b200 import CUDAMEVPDynamics, synthetic

    n, vmax = 256, 0.01
    d = CUDAMEVPDynamics(nsteps=1)
    d.setData(synthetic.benchmark_box(n, ring_mask=False))
    d.set_benchmark_forcing(0.0)
    m = 2 * n + 1  # CG2 nodes per side; a corner node of the CG field carries the corner element's DG0 value (DG2CG, x2 x2 x 1/4)
    uo, vo = d.internal("uOcean").reshape(m, m), d.internal("vOcean").reshape(m, m)
    assert uo[0, 0] == pytest.approx(-vmax, rel=1e-14) and vo[0, 0] == pytest.approx(vmax, rel=1e-14)
    assert uo[-1, -1] == pytest.approx((n - 2.0) / n * vmax, rel=1e-13) and vo[-1, -1] == pytest.approx(-(n - 2.0) / n * vmax, rel=1e-13)
    # element-centre nodes carry the element's DG0 value unchanged: the wind at element (50, 40) weakens over the first hour
    ua0 = d.internal("uAtmos").reshape(m, m)[2 * 40 + 1, 2 * 50 + 1]
    d.set_benchmark_forcing(3600.0)
    ua1 = d.internal("uAtmos").reshape(m, m)[2 * 40 + 1, 2 * 50 + 1]
    d.close()
    assert ua0 != 0.0 and abs(ua1) < abs(ua0)
