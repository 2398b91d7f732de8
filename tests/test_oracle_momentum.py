"""Analytic self-consistency checks of the oracle's momentum half.

The reference holds NO golden vector for DG2CG, CG2DG, strain, mEVP/BBM stress, divergence,
momentum or IDynamics::update: this half is "parity unpinned".  These checks
pin what can be pinned analytically (items (i)-(vi) of SURVEY.md 8(c)).
"""
import numpy as np
import pytest

import oracle
from nextsimdg_b200 import synthetic

RHO_ICE, F_ATM, F_OCEAN = 900.0, 1.2e-3 * 1.3, 5.5e-3 * 1026.0


def make(rheo="mevp", nx=12, ny=10, distort=0.0, dgadv=6, cg=2, Lx=120e3, Ly=100e3):
    coords = synthetic.box_coords(nx, ny, Lx, Ly)
    if distort:
        coords = synthetic.distort_coords(coords, distort)
    z = np.zeros((ny, nx))
    ms = {"coords": coords, "mask": np.ones((ny, nx)), "x": z, "y": z, "hice": z + 1.0, "cice": z + 1.0, "u": z, "v": z,
          "damage": z + 1.0}
    o = oracle.OracleDynamics(rheo, dgadv, cg, 1)
    o.setData(ms)
    return o, coords


def cg_nodes(coords, cg):
    """physical coordinates of the CG nodes (bilinear map of the element lattice)"""
    ny1, nx1, _ = coords.shape
    nx, ny = nx1 - 1, ny1 - 1
    X = np.zeros((cg * ny + 1, cg * nx + 1, 2))
    for j in range(cg * ny + 1):
        for i in range(cg * nx + 1):
            ex, ey = min(i // cg, nx - 1), min(j // cg, ny - 1)
            s, t = (i - cg * ex) / cg, (j - cg * ey) / cg
            c = coords
            X[j, i] = ((1 - s) * (1 - t) * c[ey, ex] + s * (1 - t) * c[ey, ex + 1] + (1 - s) * t * c[ey + 1, ex]
                       + s * t * c[ey + 1, ex + 1])
    return X


@pytest.mark.parametrize("cg,dgadv", [(2, 6), (1, 3)])
def test_lumped_mass_sums_to_area(cg, dgadv):
    o, coords = make(distort=0.05, dgadv=dgadv, cg=cg)
    assert o.internal("lumpedcgmass").sum() == pytest.approx(120e3 * 100e3, rel=1e-12)
    assert o.internal("lumpedcg1mass").sum() == pytest.approx(120e3 * 100e3, rel=1e-12)


@pytest.mark.parametrize("cg,dgadv", [(2, 6), (1, 3)])
def test_dg2cg_cg2dg_reproduce_linear(cg, dgadv):
    """(i) a globally linear field survives DG -> CG -> DG on a uniform mesh"""
    nx, ny = 12, 10
    o, coords = make(nx=nx, ny=ny, dgadv=dgadv, cg=cg)
    dx, dy = 120e3 / nx, 100e3 / ny
    f = lambda x, y: 3.0 + 2e-5 * x - 1e-5 * y
    xc = (np.arange(nx) + 0.5) * dx
    yc = (np.arange(ny) + 0.5) * dy
    dg = np.zeros((ny, nx, dgadv))
    dg[..., 0] = f(xc[None, :], yc[:, None])
    dg[..., 1] = 2e-5 * dx
    dg[..., 2] = -1e-5 * dy
    o._set("u", dg)  # ma2dg + DG2CG
    X = cg_nodes(coords, cg)
    assert np.allclose(o.internal("cg_u").reshape(X.shape[:2]), f(X[..., 0], X[..., 1]), rtol=1e-13, atol=0)
    back = o.getDG0Data("u")  # CG2DG, component 0
    assert np.allclose(back, dg[..., 0], rtol=1e-13, atol=0)


@pytest.mark.parametrize("distort", [0.0, 0.05])
def test_strain_of_linear_velocity_is_exact(distort):
    """(ii) u = a x + b y, v = c x + d y  =>  e11 = a, e22 = d, e12 = (b+c)/2 in component 0, zero above"""
    o, coords = make(distort=distort)
    X = cg_nodes(coords, 2)
    a, b, c, d = 1e-6, -2e-6, 3e-6, 0.5e-6
    o.set_internal("cg_u", a * X[..., 0] + b * X[..., 1])
    o.set_internal("cg_v", c * X[..., 0] + d * X[..., 1])
    o.sweep("strain")
    for name, val in (("e11", a), ("e22", d), ("e12", 0.5 * (b + c))):
        e = o.internal(name).reshape(-1, 8)
        assert np.allclose(e[:, 0], val, rtol=1e-11, atol=0)
        assert np.abs(e[:, 1:]).max() < 1e-11 * abs(val)


@pytest.mark.parametrize("distort", [0.0, 0.05])
def test_divergence_of_constant_stress_vanishes_inside(distort):
    """(iii) (S, grad phi_i) = 0 for interior basis functions when S is constant; the whole sum
    over nodes is zero before the Dirichlet zeroing only through boundary terms"""
    o, coords = make(distort=distort)
    ny, nx = 10, 12
    s = np.zeros((ny * nx, 8))
    s[:, 0] = 1000.0
    o.set_internal("s11", s)
    o.set_internal("s12", 0.5 * s)
    o.set_internal("s22", -2.0 * s)
    o.sweep("divergence")
    for name in ("dStressX", "dStressY"):
        d = o.internal(name).reshape(2 * ny + 1, 2 * nx + 1)
        scale = 1000.0 * 1e4  # stress * element edge length
        assert np.abs(d[1:-1, 1:-1]).max() < 1e-10 * scale
        assert np.abs(d[0, :]).max() == 0.0 and np.abs(d[:, -1]).max() == 0.0  # dirichletZero on the closed domain edge


def test_one_mevp_subcycle_uniform_state():
    """(v) u = v = 0, uniform h, a, constant wind, no ocean: strain 0 => Delta = DeltaMin, the stress
    becomes the isotropic pressure -P/(2 alpha) and interior nodes follow the hand formula"""
    o, coords = make()
    nx, ny = 12, 10
    h, a, dt, alpha, beta = 1.0, 1.0, 120.0, 1500.0, 1500.0
    one = np.ones((ny, nx))
    o.shared = {"hice": h * one, "cice": a * one, "uwind": 10 * one, "vwind": -5 * one, "uocean": 0 * one, "vocean": 0 * one,
                "ssh": 0 * one}
    for k, v in o.shared.items():
        o._set(k, v)
    o.sweep("prepare")
    o.set_delta_t(dt)
    o.subcycles(1)
    P = 27500.0 * h * np.exp(-20.0 * (1 - a))
    s11 = o.internal("s11").reshape(-1, 8)
    assert np.allclose(s11[:, 0], -0.5 * P / alpha, rtol=1e-12)
    assert np.abs(s11[:, 1:]).max() < 1e-12 * P
    assert np.abs(o.internal("s12")).max() < 1e-12 * P
    absatm = np.hypot(10.0, 5.0)
    denom = RHO_ICE * h / dt * (1 + beta)
    u_expect = a * F_ATM * absatm * 10.0 / denom
    v_expect = a * F_ATM * absatm * (-5.0) / denom
    u = o.internal("cg_u").reshape(2 * ny + 1, 2 * nx + 1)
    v = o.internal("cg_v").reshape(2 * ny + 1, 2 * nx + 1)
    assert np.allclose(u[1:-1, 1:-1], u_expect, rtol=1e-9)
    assert np.allclose(v[1:-1, 1:-1], v_expect, rtol=1e-9)
    assert np.abs(u[0]).max() == 0 and np.abs(v[:, 0]).max() == 0  # closed boundary


def test_bbm_undamaged_elastic_step():
    """(vi) d = 1, s = 0, a = 1, small uniform strain: one stress update gives
    s = lambda/(lambda+dt) * dt E/(1-nu^2) K:e and leaves the damage untouched"""
    o, coords = make("bbm")
    X = cg_nodes(coords, 2)
    e11, e22, e12 = 1e-9, -0.5e-9, 0.25e-9
    o.set_internal("cg_u", e11 * X[..., 0] + e12 * X[..., 1])
    o.set_internal("cg_v", e12 * X[..., 0] + e22 * X[..., 1])
    dt = 1.2
    o.set_delta_t(dt)
    o.sweep("strain")
    o.sweep("stress")
    nu, Y, lam = 1.0 / 3.0, 5.96e8, 1e7
    E = 1.0 * Y * 1.0 * 1.0
    D = dt * E / (1 - nu * nu)
    m = lam / (lam + dt)
    s11 = o.internal("s11").reshape(-1, 8)
    s22 = o.internal("s22").reshape(-1, 8)
    s12 = o.internal("s12").reshape(-1, 8)
    assert np.allclose(s11[:, 0], m * D * (e11 + nu * e22), rtol=1e-9)
    assert np.allclose(s22[:, 0], m * D * (nu * e11 + e22), rtol=1e-9)
    assert np.allclose(s12[:, 0], m * D * e12 * (1 - nu), rtol=1e-9)
    d = o.internal("damage").reshape(-1, 6)
    assert np.allclose(d[:, 0], 1.0, rtol=1e-13) and np.abs(d[:, 1:]).max() < 1e-12
