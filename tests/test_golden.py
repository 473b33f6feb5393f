"""Committed fixtures under tests/golden/ (self-goldens written by the oracle, see make_golden.py):
the oracle must still reproduce them on this host, and the CUDA path must hit the same targets."""
import os

import numpy as np
import pytest

from tests.cases import rel_err, scaled_err
from tests.golden.make_golden import CASES, DT, run_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, f"{name}.npz")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    gold = _load(name)
    now = run_case(name)
    for key, want in gold.items():
        if key == "oracle_git_rev":
            continue
        got = now[key]
        if key.startswith("subcycles"):
            assert int(got) == int(want), key
        else:
            wet = np.abs(want) < 1e19
            # same source, same flags; libm's exp may differ in the last bit between hosts
            assert scaled_err(got[wet][..., None], want[wet][..., None]) < 1e-13, key


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden(gpu, name):
    from mossco_code_b200 import SedimentDriver, default_config
    factory, kw, runs = CASES[name]
    case = factory()
    gold = _load(name)
    for method, nsteps in runs:
        cfg = default_config(inum=case.inum, jnum=case.jnum, knum=case.knum, dzmin=case.dzmin, dt_min=1.0,
                             **kw.get("nml", {}))
        with SedimentDriver(cfg) as sed:
            sed.set_mask(case.mask)
            sed.init_concentrations()
            sed.set_boundary(case.bdys, case.fluxes)
            sed.set_par_surface(case.par_surface)
            wet = case.mask == 0
            if nsteps == 1:
                assert scaled_err(sed.get_rhs()[wet], gold[f"rhs_m{method}"][wet]) < 1e-12
            assert sed.step(DT, method, nsteps) == 0
            want = gold[f"conc_m{method}_n{nsteps}"]
            tol = 1e-12 if nsteps == 1 else 1e-10
            assert rel_err(sed.conc[wet], want[wet]) <= tol, (name, method, nsteps)
            assert np.all(sed.conc[~wet] == 1e20)
            assert sed.info.subcycle_warnings == int(gold[f"subcycles_m{method}_n{nsteps}"])
            assert scaled_err(sed.fluxes[wet], gold[f"fluxes_m{method}_n{nsteps}"][wet]) <= tol
