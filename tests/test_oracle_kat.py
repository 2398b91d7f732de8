"""Pins the CPU oracle's advection half against the reference's own known-answer tests.

Values and tolerance are the ones stored in dynamics/test/Advection_test.cpp:26,272-301 and
dynamics/test/AdvectionPeriodicBC_test.cpp:50,281-291 (rel 1e-7; the oracle actually lands within 1e-13).
"""
import pytest

TOL = 1.0e-7  # Advection_test.cpp:26

ADVECTION = {  # Advection_test.cpp:274-283  [DG][it]
    1: (5.1256149074257538e-02, 4.8288256703303903e-02),
    3: (2.9450967798560313e-02, 1.3427281824939470e-02),
    6: (9.9340386651904228e-03, 4.0274889287136816e-03),
}
DISTORTED = {  # Advection_test.cpp:292-301
    1: (5.0958748236875594e-02, 4.8461087243594887e-02),
    3: (3.2033423984965226e-02, 1.5639052766007147e-02),
    6: (1.1830071142946669e-02, 4.9207680503709564e-03),
}
PERIODIC = {  # AdvectionPeriodicBC_test.cpp:281-291
    1: (1.0338503986019776e+00, 1.1451366598186583e+00),
    3: (1.0949680727313791e+00, 8.3851748134278226e-01),
    6: (7.1704413076190610e-01, 4.6135258391583606e-01),
}


@pytest.fixture(autouse=True)
def single_thread(oracle_lib):
    oracle_lib.nso_set_threads(1)  # tiny meshes: OpenMP overhead dominates
    yield
    oracle_lib.nso_set_threads(8)


@pytest.mark.parametrize("dg", [1, 3, 6])
@pytest.mark.parametrize("it", [0, 1])
def test_advection_straight(oracle_lib, dg, it):
    err = oracle_lib.nso_kat_advection(dg, it, 0.0, None)
    assert err == pytest.approx(ADVECTION[dg][it], rel=TOL)


@pytest.mark.parametrize("dg", [1, 3, 6])
@pytest.mark.parametrize("it", [0, 1])
def test_advection_distorted(oracle_lib, dg, it):
    err = oracle_lib.nso_kat_advection(dg, it, 0.05, None)
    assert err == pytest.approx(DISTORTED[dg][it], rel=TOL)


@pytest.mark.parametrize("dg", [1, 3, 6])
@pytest.mark.parametrize("it", [0, 1])
def test_advection_periodic_ring_with_limiter(oracle_lib, dg, it):
    err = oracle_lib.nso_kat_periodic(dg, it, None)
    assert err == pytest.approx(PERIODIC[dg][it], rel=TOL)
