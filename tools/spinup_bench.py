#!/usr/bin/env python
"""The 1-D pre-simulation of the component (fabm_sediment_component.F90:557-632: presimulation_years x 8760 steps of
3600 s) for a batch of members on the device, beside the same pre-simulation of ONE member through msed_spinup_column
(VERDICT r1 #7).  usage: python tools/spinup_bench.py [--members 1024] [--years 2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mossco_code_b200 import default_config, spinup_batch, spinup_column  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=1024)
ap.add_argument("--years", type=float, default=2.0)
ap.add_argument("--knum", type=int, default=30)
a = ap.parse_args()
nsteps = int(round(a.years * 8760))
cfg = default_config(knum=a.knum, dzmin=0.002, dt_min=1.0, bioturbation_profile=1)
rng = np.random.default_rng(5)
P = a.members
from tests.cases import C1_BDYS, C1_FLUXES  # noqa: E402  (the forcing of BASELINE config 1)
bd = np.asfortranarray(C1_BDYS[None, :] * (1 + 0.2 * rng.uniform(-1, 1, (P, 9))))      # forcing classes
fl = np.asfortranarray(C1_FLUXES[None, :] * (1 + 0.2 * rng.uniform(-1, 1, (P, 8))))
mem = [dict(rLabile=0.043 * (0.5 + rng.random())) for _ in range(P)]     # a parameter ensemble
spinup_batch(cfg, bd[:8], fl[:8], 48, 2, members=mem[:8])                # warm-up (context, module load)
t0 = time.perf_counter()
conc, infos = spinup_batch(cfg, bd, fl, nsteps, 2, members=mem)
wall = time.perf_counter() - t0
rhs = sum(i.rhs_evaluations for i in infos); sub = sum(i.subcycle_warnings for i in infos)
t0 = time.perf_counter()
one, info1 = spinup_column(default_config(knum=a.knum, dzmin=0.002, dt_min=1.0, bioturbation_profile=1, **mem[0]), bd[0], fl[0], nsteps, 2)
wall1 = time.perf_counter() - t0
out = {"members": P, "years": a.years, "steps_per_member": nsteps, "knum": a.knum, "method": "adaptive Euler",
       "batch": {"kernel_ms": infos[0].kernel_ms, "wall_s": round(wall, 4), "rhs_evaluations": int(rhs), "subcycles": int(sub),
                 "member_steps_per_s": P * nsteps / (infos[0].kernel_ms * 1e-3),
                 "cell_updates_per_s": rhs * a.knum / (infos[0].kernel_ms * 1e-3)},
       "one_member_on_device": {"kernel_ms": info1.kernel_ms, "wall_s": round(wall1, 4)},
       "bit_identical_member_0": bool(np.array_equal(conc[0], one[0], equal_nan=True))}
print(json.dumps(out, indent=1))
