#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU ORACLE (oracle/msed_oracle.c).

These are SELF-GOLDENS: the reference (Fortran + ESMF + FABM) cannot be built or run in this image and
holds no golden vectors of its own for this path, so the files pin the oracle's behaviour against
regressions and give the GPU tests fixed targets.  Inputs are regenerated from seeds by tests/cases.py.

    python tests/golden/make_golden.py        # rewrites the fixtures, records the git revision
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import msed_oracle as orc  # noqa: E402
from tests.cases import config_case, make_case  # noqa: E402

DT = 360.0
CASES = {
    # name: (case factory, oracle kwargs, [(method, nsteps), ...])
    "c1": (lambda: config_case("C1"), {}, [(2, 1), (2, 100), (1, 1), (1, 100), (3, 1), (0, 1)]),
    "c1b": (lambda: config_case("C1b"), {}, [(2, 1), (2, 100)]),
    "c2_8x8": (lambda: make_case("C2s", 8, 8, 30, 0.002, seed=1234), {}, [(2, 1), (2, 50), (1, 1)]),
    "c3_12x10": (lambda: make_case("C3s", 12, 10, 30, 0.002, seed=2024, land_fraction=0.45,
                                   smooth_temperature=True, par_max=50.0), {}, [(2, 1), (2, 50), (3, 1)]),
    "c4_6x5x40": (lambda: make_case("C4s", 6, 5, 40, 0.0015, seed=4096), {}, [(2, 1), (2, 50)]),
    "pom_flux": (lambda: make_case("pom", 5, 4, 12, 0.004, seed=5),
                 dict(nml=dict(distributed_pom_flux=1, pom_flux_max=0.3)), [(2, 1), (2, 20)]),
    "profile3": (lambda: make_case("p3", 4, 3, 15, 0.004, seed=41),
                 dict(nml=dict(bioturbation_profile=3)), [(2, 1), (1, 1)]),
}


def run_case(name):
    factory, kw, runs = CASES[name]
    case = factory()
    out = {}
    for method, nsteps in runs:
        nml = orc.sed_nml(**kw.get("nml", {}))
        o = orc.OracleSediment(case.inum, case.jnum, case.knum, case.dzmin, nml=nml, mask2d=case.mask,
                               dt_min=1.0)
        o.init_concentrations()
        o.set_boundary(case.bdys, case.fluxes)
        o.par_surface[...] = case.par_surface
        if nsteps == 1:
            out[f"rhs_m{method}"] = o.get_rhs()
        rc = o.step(DT, method, nsteps)
        assert rc == 0
        out[f"conc_m{method}_n{nsteps}"] = o.conc.copy()
        out[f"fluxes_m{method}_n{nsteps}"] = o.fluxes.copy()
        out[f"subcycles_m{method}_n{nsteps}"] = np.array(o.solver_diag()["subcycles"])
        o.finalize()
    return out


def main():
    orc.build()
    rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                         text=True).stdout.strip() or "unknown"
    for name in CASES:
        data = run_case(name)
        data["oracle_git_rev"] = np.array(rev)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **data)
        print(name, {k: getattr(v, "shape", None) for k, v in data.items() if k.startswith("conc")})


if __name__ == "__main__":
    main()
