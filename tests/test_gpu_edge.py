"""Edge cases of the path: minimum/maximum layer counts, ragged tile shapes (column counts that are
not multiples of the warp / CTA / 128-byte plane padding), all-land and single-column tiles, non-default
model parameters (active CprodMax cap, NH3 adsorption, non-zero minima), bcup_dissolved=0/3."""
import numpy as np
import pytest

from tests.cases import make_case, rel_err, scaled_err

pytestmark = pytest.mark.gpu
DT = 360.0


def _both(oracle, case, method=2, nsteps=2, **cfgkw):
    from mossco_code_b200 import SedimentDriver, default_config
    kw = dict(inum=case.inum, jnum=case.jnum, knum=case.knum, dzmin=case.dzmin, dt_min=1.0)
    kw.update(cfgkw)
    cfg = default_config(**kw)
    ref = oracle.OracleSediment.from_config(cfg, mask2d=case.mask)
    ref.init_concentrations(); ref.set_boundary(case.bdys, case.fluxes)
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        rhs_g = sed.get_rhs()
        rc = sed.step(DT, method, nsteps)
        got, flux, info_sub = sed.conc, sed.fluxes, sed.info.subcycle_warnings
    rhs_r = ref.get_rhs()
    assert ref.step(DT, method, nsteps) == rc == 0
    wet = case.mask == 0
    assert scaled_err(rhs_g[wet], rhs_r[wet]) < 1e-12 if wet.any() else True
    assert rel_err(got[wet], ref.conc[wet]) <= 1e-11
    assert np.all(got[~wet] == 1e20)
    assert info_sub == ref.solver_diag()["subcycles"]
    if wet.any():
        assert scaled_err(flux[wet], ref.fluxes[wet]) <= 1e-11
    return got


@pytest.mark.parametrize("knum,dzmin", [(2, 0.05), (3, 0.03), (5, 0.02), (63, 0.0005), (64, 0.0005)])
@pytest.mark.parametrize("method", [2, 1])
def test_layer_count_limits(gpu, oracle, knum, dzmin, method):
    case = make_case("k", 5, 3, knum, dzmin, seed=knum)
    _both(oracle, case, method=method, nsteps=2)


def test_knum_above_limit_is_rejected(gpu):
    from mossco_code_b200 import MsedError, SedimentDriver, default_config
    with pytest.raises(MsedError) as e:
        SedimentDriver(default_config(inum=2, jnum=2, knum=65, dzmin=0.001))
    assert e.value.code == -1


@pytest.mark.parametrize("inum,jnum", [(1, 1), (1, 37), (37, 1), (31, 1), (33, 4), (127, 1), (129, 1),
                                        (17, 15), (255, 3), (128, 2)])
def test_ragged_tile_shapes(gpu, oracle, inum, jnum):
    case = make_case("r", inum, jnum, 12, 0.004, seed=inum * 100 + jnum, land_fraction=0.3 if inum * jnum > 8 else 0.0)
    _both(oracle, case, nsteps=2)
    _both(oracle, case, method=3, nsteps=1)


def test_all_land_tile(gpu, oracle):
    case = make_case("land", 9, 4, 10, 0.005, seed=1)
    case.mask[...] = 1
    from mossco_code_b200 import SedimentDriver, default_config
    cfg = default_config(inum=9, jnum=4, knum=10, dzmin=0.005, dt_min=1.0)
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        assert sed.check_domain() == 0
        for method in (0, 1, 2, 3):
            assert sed.step(DT, method, 2) == 0
        assert np.all(sed.conc == 1e20)
        assert np.all(sed.get_rhs() == 0.0)
        assert sed.info.subcycle_warnings == 0


@pytest.mark.parametrize("kw", [
    dict(CprodMax=20.0),                      # the cap on Cprod is active (rLabile*ldetC ~ 245/d > 20/d)
    dict(NH3Ads=1.3),                         # ammonium adsorption divides the nh3 rate
    dict(rLabile=0.1, rSemilabile=0.01, NCrLdet=0.15, NCrSdet=0.13, PAds=0.3, PAdsODU=10.0, rnit=20.0,
         ksO2nitri=1.0, rODUox=5.0, ksO2oduox=2.0, ksO2oxic=1.0, ksNO3denit=30.0, kinO2denit=10.0,
         kinNO3anox=5.0, kinO2anox=5.0),      # "reference values" column of fabm_sed.nml
    dict(diffusivity=2.5, bioturbation=3.0, bioturbation_depth=11.0, bioturbation_min=1.0, porosity_max=0.9,
         porosity_fac=1.5, bioturbation_profile=2),
    dict(initial_value=[1e3, 2e3, 10., 1., 5., 80., 300., 0.], minimum=[1., 2., 3., 0.5, 30., 1., 2., 150.]),
])
def test_non_default_parameters(gpu, oracle, kw):
    case = make_case("par", 7, 5, 15, 0.004, seed=31)
    got = _both(oracle, case, nsteps=3, **kw)
    if "minimum" in kw:                        # clip to state_variables(n)%minimum, component :1726-1732
        for n, m in enumerate(kw["minimum"]):
            assert got[..., n].min() >= m
        assert got[..., 7].min() == 150.0      # odu starts at 0/porosity: lifted to its minimum


@pytest.mark.parametrize("bcup", [0, 3])
def test_no_flux_upper_boundary_for_dissolved(gpu, oracle, bcup):
    """bcup_dissolved 3 = zero flux (:789).  With 0, diff3d never assigns Flux(1) (:782-803): it keeps
    what the previous variable's call left in get_rhs's intFlux array, i.e. the detP input flux --
    reproduced by oracle and kernel alike."""
    case = make_case("bc", 6, 4, 12, 0.004, seed=9)
    from mossco_code_b200 import SedimentDriver, default_config
    cfg = default_config(inum=6, jnum=4, knum=12, dzmin=0.004, dt_min=1.0, bcup_dissolved_variables=bcup)
    with SedimentDriver(cfg) as sed:
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        rhs = sed.get_rhs()
        if bcup == 3:
            assert np.all(sed.fluxes[:, :, 3:] == 0.0)
        else:
            assert np.array_equal(sed.fluxes[:, :, 3:], np.repeat(case.fluxes[:, :, 2:3], 5, axis=2))
    ref = oracle.OracleSediment.from_config(cfg)
    ref.init_concentrations(); ref.set_boundary(case.bdys, case.fluxes)
    assert scaled_err(rhs, ref.get_rhs()) < 1e-12


def test_large_time_step_rk_negative_concentrations(gpu, oracle):
    """RK stages may drive concentrations negative; rates must stay finite and match the oracle."""
    case = make_case("neg", 5, 4, 12, 0.004, seed=13)
    from mossco_code_b200 import SedimentDriver, default_config
    cfg = default_config(inum=5, jnum=4, knum=12, dzmin=0.004, dt_min=1.0)
    ref = oracle.OracleSediment.from_config(cfg)
    ref.init_concentrations(); ref.set_boundary(case.bdys, case.fluxes)
    with SedimentDriver(cfg) as sed:
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        sed.ode_solver(1800.0, 1)
        got = sed.conc
    ref.ode_solver(1800.0, 1)
    assert np.isfinite(ref.conc).all()
    assert scaled_err(got, ref.conc) <= 1e-10


def test_rk_stages_per_launch_argument_is_checked(gpu):
    """msed_set_rk_stages_per_launch takes 2 (stage pairs) or 4 (the whole call in one launch), nothing else."""
    from mossco_code_b200 import MsedError, SedimentDriver, default_config
    with SedimentDriver(default_config(inum=4, jnum=3, knum=8, dzmin=0.01)) as sed:
        sed.set_rk_stages_per_launch(2)
        sed.set_rk_stages_per_launch(4)
        for bad in (0, 1, 3, 8):
            with pytest.raises(MsedError):
                sed.set_rk_stages_per_launch(bad)
