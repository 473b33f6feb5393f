"""Batched 1-D pre-simulation (msed_spinup_batch; fabm_sediment_component.F90:557-632): every member must be
the single-column spin-up (msed_spinup_column) bit for bit -- state, sub-cycle and attempt counts, last_min_dt and
its grid cell -- with its own accept decisions, and match the restated reference within the 10-day bar."""
import numpy as np
import pytest

from tests.cases import C1_BDYS, C1_FLUXES, C1B_BDYS, C1B_FLUXES, scaled_err

pytestmark = pytest.mark.gpu


def _members(P, seed=5):
    rng = np.random.default_rng(seed)
    bd = np.empty((P, 9), order="F")
    fl = np.empty((P, 8), order="F")
    mem = []
    for m in range(P):
        base_b, base_f = (C1_BDYS, C1_FLUXES) if m % 2 == 0 else (C1B_BDYS, C1B_FLUXES)
        bd[m] = base_b * (1 + 0.2 * rng.uniform(-1, 1, 9))
        fl[m] = base_f * (1 + 0.2 * rng.uniform(-1, 1, 8))
        over = {}
        if m % 3 == 1:      # a stiff member: sub-cycles while its neighbours do not
            over.update(rnit=2.0e3 * (1 + rng.random()), rODUox=2.0e3)
        if m % 4 == 2:
            over.update(rLabile=0.043 * (0.5 + rng.random()), initial_value=[3e3, 5e3, 30., 8., 25., 30., 150., 50.])
        mem.append(over)
    return bd, fl, mem


@pytest.mark.parametrize("method", [2, 0, 1, 3])
@pytest.mark.parametrize("knum,dzmin", [(15, 0.004), (30, 0.002), (32, 0.002), (33, 0.002), (40, 0.0015), (64, 0.0005)])
def test_spinup_batch_equals_single_columns(gpu, method, knum, dzmin):
    from mossco_code_b200 import default_config, spinup_batch, spinup_column
    cfg = default_config(knum=knum, dzmin=dzmin, dt_min=1.0, bioturbation_profile=1)
    P, nsteps = 13, 24 * 6 if method == 2 else 24 * 2
    bd, fl, mem = _members(P)
    got, infos = spinup_batch(cfg, bd, fl, nsteps, method, members=mem)
    assert got.shape == (P, 1, knum, 8)
    subs = []
    for m in range(P):
        cm = default_config(knum=knum, dzmin=dzmin, dt_min=1.0, bioturbation_profile=1, **mem[m])
        want, wi = spinup_column(cm, bd[m], fl[m], nsteps, method, launch_per_attempt=True)
        assert np.array_equal(got[m], want[0], equal_nan=True), m
        assert infos[m].steps_done == nsteps
        assert infos[m].rhs_evaluations == wi.rhs_evaluations, m
        assert infos[m].subcycle_warnings == wi.subcycle_warnings, m
        if method == 2:
            assert infos[m].last_min_dt == wi.last_min_dt, m
            assert list(infos[m].last_min_dt_grid_cell) == list(wi.last_min_dt_grid_cell), m
        subs.append(infos[m].subcycle_warnings)
    if method == 2:
        assert min(subs) > 0 and max(subs) > 1.5 * min(subs)     # members really decide for themselves


def test_spinup_batch_matches_oracle_and_shared_parameters(gpu, oracle):
    from mossco_code_b200 import default_config, spinup_batch
    cfg = default_config(knum=15, dzmin=0.004, dt_min=1.0, bioturbation_profile=2)
    P, nsteps = 6, 24 * 20
    bd, fl, _ = _members(P, seed=9)
    got, infos = spinup_batch(cfg, bd, fl, nsteps, 2)        # no per-member parameters: cfg for all
    nml, par = oracle.from_config(cfg)
    for m in range(P):
        want = oracle.spinup_column(nml, par, 15, 0.004, 1.0, -0.9, bd[m], fl[m], nsteps)
        assert scaled_err(got[m], want[0]) <= 1e-8, m


def test_spinup_batch_outside_the_kernels_scope_loops_over_columns(gpu):
    """A distributed POM flux: member by member through msed_spinup_column, same interface."""
    from mossco_code_b200 import default_config, spinup_batch, spinup_column
    for kw in (dict(knum=12, dzmin=0.004, distributed_pom_flux=1),):
        cfg = default_config(dt_min=1.0, **kw)
        bd, fl, mem = _members(3)
        got, infos = spinup_batch(cfg, bd, fl, 30, 2, members=mem)
        for m in range(3):
            cm = default_config(dt_min=1.0, **kw, **mem[m])
            want, wi = spinup_column(cm, bd[m], fl[m], 30, 2)
            assert np.array_equal(got[m], want[0]) and infos[m].subcycle_warnings == wi.subcycle_warnings
