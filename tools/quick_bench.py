#!/usr/bin/env python
"""Quick device-side timing of the step kernels (development helper; bench.py is the contract)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mossco_code_b200 import SedimentDriver, default_config  # noqa: E402
from tests.cases import make_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--inum", type=int, default=4096)
ap.add_argument("--jnum", type=int, default=512)
ap.add_argument("--knum", type=int, default=40)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--method", type=int, default=2)
ap.add_argument("--spin", type=int, default=10)
ap.add_argument("--perturb", type=float, default=0.1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fusion", choices=["on", "off", "pairs", "chains"], default="on")
ap.add_argument("--land", type=float, default=0.0)
ap.add_argument("--rk-stages", type=int, default=4, help="Runge-Kutta, thread per column: stages per launch (4 or 2)")
a = ap.parse_args()
dzmin = 0.0015 if a.knum == 40 else 0.002
t0 = time.time()
case = make_case("qb", a.inum, a.jnum, a.knum, dzmin, seed=4096, perturb=a.perturb, land_fraction=a.land)
cfg = default_config(inum=a.inum, jnum=a.jnum, knum=a.knum, dzmin=dzmin, dt_min=1.0)
sed = SedimentDriver(cfg)
if a.land > 0:
    sed.set_mask(case.mask)
sed.init_concentrations()
sed.set_step_fusion({"on": "auto"}.get(a.fusion, a.fusion))
sed.set_rk_stages_per_launch(a.rk_stages)
sed.set_boundary(case.bdys, case.fluxes)
print(f"setup {time.time()-t0:.1f}s")
sed.step(360.0, a.method, a.spin)
print("spin: subcycles", sed.info.subcycle_warnings, "rhs evals", sed.info.rhs_evaluations)
for rep in range(a.reps):
    sed.step(360.0, a.method, a.steps)
    i = sed.info
    cells = a.inum * a.jnum * a.knum * (float((case.mask == 0).mean()) if a.land > 0 else 1.0)
    ms = i.kernel_ms / a.steps
    balg = 136.0 + 216.0 / a.knum
    print(f"rep{rep}: {ms:.3f} ms/step  {cells/ms/1e6:.2f} Gcell-updates/s  alg {cells*balg/ms/1e6:.0f} GB/s "
          f"({cells*balg/ms/1e6/6547.2*100:.1f}% of measured HBM) subcycles={i.subcycle_warnings} "
          f"rhs={i.rhs_evaluations} launches={i.kernel_launches} ms/rhs={i.kernel_ms/max(i.rhs_evaluations,1):.3f}")
