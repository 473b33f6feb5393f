"""Pins the CPU oracle (oracle/msed_oracle.c) against everything the reference offers for this path:
the one solver test program (src/test/test_Solver.F90) and closed forms that follow from the
reference source alone.  No GPU needed."""
import numpy as np
import pytest

from tests.cases import C1_BDYS, C1_FLUXES, config_case, make_case, scaled_err


# ---- 1. grid closed form, fabm_sediment_driver.F90:147-168 ----------------------------------------
@pytest.mark.parametrize("K,dzmin,self_fac,dz_last,zi_last", [
    (15, 0.004, 4.625, 0.0185, 0.16875),
    (30, 0.002, 4.806451612903226, 0.0096129032258064524, 0.17419354838709677),
    (40, 0.0015, 4.8536585365853657, None, 0.17560975609756097),
])
def test_grid_closed_form(oracle, K, dzmin, self_fac, dz_last, zi_last):
    s = oracle.OracleSediment(2, 3, K, dzmin)
    dz, zi, zc, dzc = (s.field3d(n) for n in ("dz", "zi", "zc", "dzc"))
    assert 0.18 / ((K + 1) / 2.0 * dzmin) - 1.0 == pytest.approx(self_fac, rel=1e-15)
    assert dz[0, 0, 0] == dzmin
    if dz_last is not None:
        assert dz[1, 2, K - 1] == pytest.approx(dz_last, rel=1e-14)
    assert dz[0, 0, K - 1] / dz[0, 0, 0] == pytest.approx(self_fac, rel=1e-13)
    assert zi[0, 0, K] == pytest.approx(zi_last, rel=1e-14)
    assert zi[0, 0, K] == pytest.approx(0.18 * K / (K + 1), rel=1e-13)   # sum of the linear ramp
    assert np.allclose(zc[0, 0], 0.5 * (zi[0, 0, :-1] + zi[0, 0, 1:]), rtol=1e-14)
    assert np.array_equal(dzc[0, 0], zc[0, 0, 1:] - zc[0, 0, :-1])
    assert np.all(dz == dz[:1, :1, :])                                     # horizontally uniform


# ---- 2. porosity / flux_cap / bioturbation_factor profiles, :278-304,:434-435 -------------------
def test_static_profiles(oracle):
    s = oracle.OracleSediment(1, 1, 15, 0.004)
    por, intf, zc, zi, dz = (s.field3d(n)[0, 0] for n in ("porosity", "intf_porosity", "zc", "zi", "dz"))
    assert por[0] == pytest.approx(0.69874, rel=1e-12)
    assert por[14] == pytest.approx(0.599515, rel=1e-12)
    assert np.allclose(por, 0.7 * (1 - 0.9 * zc), rtol=1e-15)
    assert intf[0] == por[0] and np.array_equal(intf[1:], 0.5 * (por[:-1] + por[1:]))
    cap = s.field3d("flux_cap")[0, 0]
    raw = 2.0e4 / 86400.0 * (1.0 - por) * dz
    assert cap[0] == raw[0] and cap[1] == raw[1]                          # k=2 is not clamped (:285)
    assert np.all(np.diff(cap[1:]) <= 0)
    bf = s.field3d("bioturbation_factor")[0, 0]                            # profile 1 (:296-298)
    want = np.maximum(0.2 / 0.9, np.maximum(5.0 - 100.0 * zi[:-1], 0.0) / 5.0)
    assert np.array_equal(bf, want)
    s2 = oracle.OracleSediment(1, 1, 15, 0.004, nml=oracle.sed_nml(bioturbation_profile=2))
    assert np.array_equal(s2.field3d("bioturbation_factor")[0, 0], np.exp(-100.0 * zi[:-1] / 5.0))
    s0 = oracle.OracleSediment(1, 1, 15, 0.004, nml=oracle.sed_nml(bioturbation_profile=0))
    assert np.all(s0.field3d("bioturbation_factor") == 1.0)


def test_update_porosity_from_surface(oracle):
    s = oracle.OracleSediment(2, 2, 10, 0.005)
    surf = np.array([[0.5, 0.6], [0.7, 0.8]])
    s.update_porosity(surf)
    por, zc = s.field3d("porosity"), s.field3d("zc")
    for k in range(1, 10):
        assert np.array_equal(por[:, :, k], surf * (1.0 - 0.9 * (zc[:, :, k] - zc[:, :, 0])))


def test_init_concentrations_divides_all_by_porosity(oracle):
    s = oracle.OracleSediment(1, 1, 15, 0.004)
    s.init_concentrations()
    por = s.field3d("porosity")[0, 0]
    init = [4e3, 4e3, 40., 10., 20., 40., 100., 100.]
    for n in range(8):                                                     # incl. particulates (:456)
        assert np.array_equal(s.conc[0, 0, :, n], init[n] / por)


# ---- 3. diff3d analytic cases, :739-825 -----------------------------------------------------------
def _transport_only(oracle, K=12, bcup=2, **nml):
    s = oracle.OracleSediment(1, 1, K, 0.004, nml=oracle.sed_nml(**nml) if nml else None,
                              model=oracle.MODEL_NONE, bcup_dissolved_variables=bcup)
    s.init_concentrations()
    return s


def test_diff3d_uniform_zero_gradient(oracle):
    s = _transport_only(oracle, bcup=3)
    s.conc[...] = 7.0
    s.field3d("porosity")[...] = 0.6           # C = conc*por uniform too
    s.update_porosity(from_surface=False)
    s.set_boundary(np.zeros((1, 1, 9)), np.zeros((1, 1, 8)))
    rhs = s.get_rhs()
    assert np.all(rhs == 0.0)                  # BcUp=3/BcDown=3 (dissolved), zero flux (particulate)


def test_diff3d_flux_boundary_and_mass_conservation(oracle):
    s = _transport_only(oracle, bcup=2)
    rng = np.random.default_rng(0)
    s.conc[...] = s.conc * (1 + 0.3 * rng.uniform(-1, 1, s.conc.shape))
    b = C1_BDYS.reshape(1, 1, 9).copy()
    f = C1_FLUXES.reshape(1, 1, 8).copy()
    s.set_boundary(b, f)
    rhs = s.get_rhs()[0, 0]                    # transport only
    por, dz = s.field3d("porosity")[0, 0], s.field3d("dz")[0, 0]
    flux_top = s.fluxes[0, 0]
    for n in range(8):
        # inventory change = Flux(1) - Flux(K+1) = Flux(1)   (BcDown=3)
        inv = np.sum(rhs[:, n] * por * dz)
        assert inv == pytest.approx(flux_top[n], rel=1e-11, abs=1e-22)
    # Dirichlet top flux (:786): -D(1)*(C(1)-Cup)/dz(1), D from :652,:682
    temp = 5.0
    f_T = np.exp(-4500.0 * (1.0 / (temp + 273.0) - 1.0 / 288.0))
    D1 = 0.9 * f_T / 86400.0 / 10000.0 * (1 - por[0]) * 1.0 + (0.9 + temp * 0.035) * por[0] / 86400.0 / 10000.0
    for n in range(3, 8):
        assert flux_top[n] == pytest.approx(-D1 * (s.conc[0, 0, 0, n] - b[0, 0, n + 1]) / dz[0], rel=1e-13)
    # flux BC for particulates (:783): dC(1) gets F/(VF*dz) then the (1-por)/por rescale (:677)
    s.conc[...] = 5.0
    s.field3d("porosity")[...] = 0.5
    s.update_porosity(from_surface=False)
    rhs = s.get_rhs()[0, 0]
    assert rhs[0, 0] == pytest.approx(f[0, 0, 0] / (0.5 * dz[0]), rel=1e-14)
    assert np.all(rhs[1:, 0] == 0.0)


def test_diff3d_distributed_pom_cascade(oracle):
    """BcUp=4 (:791-803): flux above the cap is handed down interface by interface."""
    K = 6
    s = _transport_only(oracle, K=K, distributed_pom_flux=1, pom_flux_max=8.64)  # cap rate 1e-4*(1-por)*dz
    s.conc[...] = 3.0
    s.field3d("porosity")[...] = 0.5
    s.update_porosity(from_surface=True, porosity_surface=np.full((1, 1), 0.5))
    por, dz = s.field3d("porosity")[0, 0], s.field3d("dz")[0, 0]
    cap = s.field3d("flux_cap")[0, 0].copy()
    F = 2.5 * cap[0]
    f = np.zeros((1, 1, 8)); f[0, 0, 0] = F
    s.set_boundary(np.zeros((1, 1, 9)), f)
    s.conc[...] = 3.0 / por[None, None, :, None]      # uniform bulk concentration: no diffusive flux
    rhs = s.get_rhs()[0, 0, :, 0]
    flux = np.zeros(K + 1); flux[0] = F
    rest, k = F - cap[0], 1
    while rest > 0 and k < K:
        flux[k] += rest; rest -= cap[k]; k += 1
    if k >= K:
        flux[K - 1] += rest
    want = (flux[:-1] - flux[1:]) / ((1 - por) * dz) * (1 - por) / por
    assert np.allclose(rhs, want, rtol=1e-9, atol=1e-18)
    assert np.sum(rhs * por * dz) == pytest.approx(F, rel=1e-9)


# ---- 4. ode_solver KAT from src/test/test_Solver.F90 -------------------------------------------------
def test_solver_kat_euler_closed_form(oracle):
    inum, jnum, knum, nvar, n = 100, 1, 24, 8, 1000
    conc = np.zeros((inum, jnum, knum, nvar), order="F")
    for k in range(1, knum + 1):      # solv%conc(:,:,k,:)=1.0 + k*0.1 in DEFAULT REAL (:76)
        conc[:, :, k - 1, :] = np.float64(np.float32(1.0) + np.float32(k) * np.float32(0.1))
    start = conc.copy()
    out = oracle.test_solver_kat(inum, jnum, knum, nvar, conc, 1.0, 0, n)
    i = np.arange(1, inum + 1)[:, None, None, None]
    k = np.arange(1, knum + 1)[None, None, :, None]
    closed = start + n * 1.0 * (i + 1 + k) * 1.0e-8
    assert np.max(np.abs(out - closed)) < 5e-13      # only accumulated rounding of n additions
    # and exactly the sequential floating-point sum
    seq = start.copy()
    for _ in range(n):
        seq = seq + 1.0 * ((i + 1 + k) * 1.0e-8)
    assert np.array_equal(out, seq)


# ---- 5. adaptive Euler control flow, solver_library.F90:104-140 -----------------------------------------
def test_adaptive_euler_control_flow(oracle):
    case = make_case("sub", 3, 2, 15, 0.004, seed=2)
    par = oracle.omexdia_params(rnit=2.0e3, rODUox=2.0e3)
    a = oracle.OracleSediment(3, 2, 15, 0.004, params=par, dt_min=1.0)
    a.init_concentrations(); a.set_boundary(case.bdys, case.fluxes)
    c0 = a.conc.copy()
    rhs0 = a.get_rhs()
    assert np.any(c0 + 360.0 * rhs0 - 0.1 * c0 < 0)          # the full step violates (:121)
    a.ode_solver(360.0, 2)
    d = a.solver_diag()
    assert d["subcycles"] >= 1
    m = d["subcycles"]
    # dt_red = dt/4^m, then 4^m accepted sub-steps reproduce explicit Euler with that dt
    b = oracle.OracleSediment(3, 2, 15, 0.004, params=par, dt_min=1.0)
    b.init_concentrations(); b.set_boundary(case.bdys, case.fluxes)
    for _ in range(4 ** m):
        b.ode_solver(360.0 / 4 ** m, 0)
    assert np.array_equal(a.conc, b.conc)
    # dt_min >= dt accepts the violating step as is (:126)
    c = oracle.OracleSediment(3, 2, 15, 0.004, params=par, dt_min=1000.0)
    c.init_concentrations(); c.set_boundary(case.bdys, case.fluxes)
    c.ode_solver(360.0, 2)
    assert c.solver_diag()["subcycles"] == 0
    assert np.array_equal(c.conc, c0 + 360.0 * rhs0)


def test_adaptive_diagnostics_minloc(oracle):
    case = config_case("C1")
    s = oracle.OracleSediment(1, 1, 30, 0.002, dt_min=1.0, adaptive_solver_diagnostics=True)
    s.init_concentrations(); s.set_boundary(case.bdys, case.fluxes)
    c0 = s.conc.copy()
    s.ode_solver(360.0, 2)
    d = s.solver_diag()
    assert d["last_min_dt"] == 360.0
    rel = (s.conc - c0) / c0
    i, j, k, n = np.unravel_index(np.argmin(rel.ravel(order="F")), rel.shape, order="F")
    assert d["last_min_dt_grid_cell"] == [i + 1, j + 1, k + 1, n + 1]


def test_adaptive_decision_sensitivity(oracle):
    """Evidence for the parity protocol: with whole-domain accept/reject decisions the 10-day C1
    state of the ORACLE ITSELF moves by more than 1e-8 when the initial state is perturbed by 1e-15,
    because the sub-cycling history changes.  Parity over long adaptive runs is therefore asserted
    while both sides take identical decisions (tests/test_gpu_parity.py::_lockstep)."""
    case = config_case("C1")
    out, hist = [], []
    for eps in (0.0, 1e-15):
        s = oracle.OracleSediment(1, 1, 30, 0.002, dt_min=1.0)
        s.init_concentrations(); s.set_boundary(case.bdys, case.fluxes)
        s.conc[...] = s.conc * (1 + eps)
        h = []
        for _ in range(240):
            s.step(360.0, 2, 10)
            h.append(s.solver_diag()["subcycles"])
        out.append(s.conc.copy()); hist.append(h)
    first_diff = next((i for i, (a, b) in enumerate(zip(*hist)) if a != b), None)
    assert first_diff is not None and first_diff > 100      # same decisions for > 1000 steps
    assert scaled_err(out[1], out[0]) > 1e-9                 # ... and then visibly different states


# ---- 6. RK4 and RK4-3/8 on linear decay, :142-185 ------------------------------------------------------
def test_rk4_variants_linear_decay(oracle):
    """With transport off (uniform state, zero fluxes) ldetC obeys dc/dt = -f_T*rLabile*c."""
    T = 10.0
    for method in (1, 3):
        s = oracle.OracleSediment(1, 1, 8, 0.01, dt_min=1.0, bcup_dissolved_variables=3)
        s.init_concentrations()
        por = s.field3d("porosity")
        s.conc[0, 0, :, :] = (4000.0 / por[0, 0])[:, None]     # uniform bulk -> no particulate flux
        b = np.zeros((1, 1, 9)); b[0, 0, 0] = T
        s.set_boundary(b, np.zeros((1, 1, 8)))
        c0 = s.conc[0, 0, :, 0].copy()
        dt = 86400.0 * 5
        s.ode_solver(dt, method)
        E_a = 0.1 * np.log(1.5) * 288.15 * 298.15
        lam = np.exp(-E_a * (1 / (T + 273.15) - 1 / 288.15)) * 0.043 / 86400.0
        x = lam * dt
        want = c0 * (1 - x + x ** 2 / 2 - x ** 3 / 6 + x ** 4 / 24)
        assert np.allclose(s.conc[0, 0, :, 0], want, rtol=1e-13)


# ---- 7. masked columns, :464,:703-709 ---------------------------------------------------------------------
def test_masked_column(oracle):
    mask = np.zeros((3, 2), dtype=np.int32); mask[1, 0] = 1
    case = make_case("m", 3, 2, 10, 0.005, seed=3)
    s = oracle.OracleSediment(3, 2, 10, 0.005, mask2d=mask, dt_min=1.0)
    assert s.check_domain() == 0
    s.init_concentrations(); s.set_boundary(case.bdys, case.fluxes)
    assert np.all(s.conc[1, 0] == 1e20) and np.all(s.field3d("porosity")[1, 0] == 1.0)
    rhs = s.get_rhs()
    assert np.all(rhs[1, 0] == 0.0)
    for method in (0, 1, 2, 3):
        assert s.step(360.0, method, 2) == 0
        assert np.all(s.conc[1, 0] == 1e20)
    assert s.solver_diag()["subcycles"] == 0
    assert np.all(s.field3d("temp3d")[1, 0] == -999.0)


# ---- reaction term: independent numpy transcription of SURVEY.md Appendix B ---------------------------------
def _omexdia_numpy(c, T):
    ldetC, sdetC, detP, po4, no3, nh3, oxy, odu = c
    rL, rS, rnit, rodu = 0.043 / 86400, 0.001 / 86400, 200. / 86400, 20. / 86400
    E_a = 0.1 * np.log(1.5) * 288.15 * 298.15
    fT = np.exp(-E_a * (1 / (T + 273.15) - 1 / 288.15))
    ox = oxy / (oxy + 3. + 0.04 * (nh3 + odu))
    de = (1 - oxy / (oxy + 70.)) * no3 / (no3 + 1.)
    an = (1 - oxy / (oxy + 1.)) * (1 - no3 / (no3 + 1.))
    resc = 1 / (ox + de + an)
    cl, cs = rL * ldetC, rS * sdetC
    cp = min(cl + cs, 9600. / 86400)
    npr = cl * 0.22 + cs * 0.005
    rads = 0.01 * rS * po4 * max(odu, 70.)
    pp = rL * (1 - ox) * detP
    nit = fT * rnit * nh3 * oxy / (oxy + 20. + 0.04 * (ldetC + odu))
    oo = fT * rodu * odu * oxy / (oxy + 1. + 0.04 * (nh3 + ldetC))
    return np.array([-fT * cl, -fT * cs, fT * (rads - pp), fT * (pp - rads), -0.8 * cp * de * resc + nit,
                     (npr - nit) / 1.0, -cp * ox * resc - 2 * nit - oo, cp * an * resc - oo]), 0.8 * cp * de * resc


def test_omexdia_reaction_spec(oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        c = np.array([6000., 6000., 60., 15., 30., 60., 150., 150.]) * rng.uniform(0.0, 2.0, 8)
        T = rng.uniform(0, 25)
        r, d = oracle.omexdia_cell(c, T)
        wr, wd = _omexdia_numpy(c, T)
        assert np.allclose(r, wr, rtol=1e-12, atol=1e-25)
        assert d == pytest.approx(wd, rel=1e-12)
    # conservation built into the formulation: P is only exchanged between detP and po4
    assert r[2] + r[3] == pytest.approx(0.0, abs=1e-20)


def test_boundary_conditions_formulas(oracle):
    """get_boundary_conditions, component :1930-2020."""
    s = oracle.OracleSediment(2, 2, 10, 0.005, bcup_dissolved_variables=1)
    s.init_concentrations()
    rng = np.random.default_rng(1)
    temp = 10 * rng.random((2, 2))
    cs = [rng.random((2, 2)) for _ in range(8)]
    wz = [-rng.random((2, 2)) for _ in range(3)] + [None] * 5
    s.get_boundary_conditions(temp, cs, wz)
    assert np.array_equal(s.bdys[:, :, 0], temp)
    por, dz = s.field3d("porosity"), s.field3d("dz")
    for n in range(3):
        assert np.array_equal(s.fluxes[:, :, n], -cs[n] * wz[n])
    for n in range(3, 8):
        assert np.array_equal(s.bdys[:, :, n + 1], cs[n])
        want = -(s.conc[:, :, 0, n] - cs[n]) / dz[:, :, 0] * (0.9 + 0.9 + temp * 0.035) * por[:, :, 0] / 86400. / 10000.
        assert np.allclose(s.fluxes[:, :, n], want, rtol=1e-15)


def test_soil_pelagic_connector_formulas(oracle):
    """soil_pelagic_connector Run against the formulas read off the reference
    (src/mediators/soil_pelagic_connector.F90:333-359,:409-411,:467-472,:529-533,:660-720,:842-874)."""
    rng = np.random.default_rng(12)
    up = np.asfortranarray(rng.normal(size=(6, 4, 8)) * 1e-5)
    year = float(np.float32(86400.0) * np.float32(365.0))
    assert year == 31536000.0
    ldetC, sdetC, po4, no3, nh3, oxy, odu = (up[:, :, n] for n in (0, 1, 3, 4, 5, 6, 7))
    r = oracle.soil_pelagic_connector(up, dinflux_const=0.3, convertN=1.5, convertP=0.75)
    assert np.array_equal(r["nitrate"], no3)                                  # no convertN on nitrate
    assert np.array_equal(r["ammonium"], 1.5 * nh3)
    assert np.array_equal(r["DIN"], (nh3 + no3 + 0.3 / year) * 1.5)
    assert np.array_equal(r["DIP"], 0.75 * (po4 + (0.3 / 16.0) / year))       # dipflux_const < 0: Redfield (:156)
    assert np.array_equal(r["oxygen"], oxy) and np.array_equal(r["odu"], odu)
    assert np.array_equal(r["detC"], ldetC + sdetC)
    assert not r["detN"].any() and not r["detP"].any()                        # no matching import fields
    r = oracle.soil_pelagic_connector(up, want=("odu", "DIP"), dipflux_const=0.02)
    assert np.array_equal(r["odu"], odu - oxy) and set(r) == {"odu", "DIP"}
    assert np.array_equal(r["DIP"], po4 + 0.02 / year)
    r = oracle.soil_pelagic_connector(up, want=("oxygen",))
    assert np.array_equal(r["oxygen"], oxy - odu)
