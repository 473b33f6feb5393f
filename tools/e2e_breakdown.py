#!/usr/bin/env python
"""Where does a host-buffer Run() spend its time?  (development helper)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, pinned_fortran, slab_forcing  # noqa: E402
from mossco_code_b200 import SedimentDriver, default_config  # noqa: E402

wlname = sys.argv[1] if len(sys.argv) > 1 else "c4"
inum, jnum, knum, dzmin, seed, land, desc = WORKLOADS[wlname]
bdys, fluxes, mask, par = slab_forcing(WORKLOADS[wlname], 0, jnum)
cfg = default_config(inum=inum, jnum=jnum, knum=knum, dzmin=dzmin, dt_min=1.0)
sed = SedimentDriver(cfg)
sed.init_concentrations()
keep = []
t_, temp = pinned_fortran((inum, jnum)); temp[...] = bdys[:, :, 0]; keep.append(t_)
cs, wz = [], []
for n in range(8):
    t_, a = pinned_fortran((inum, jnum)); keep.append(t_)
    a[...] = -fluxes[:, :, n] if n < 3 else bdys[:, :, n + 1]
    cs.append(a)
    if n < 3:
        t_, w = pinned_fortran((inum, jnum)); w[...] = 1.0; keep.append(t_); wz.append(w)
    else:
        wz.append(None)
t_, up = pinned_fortran((inum, jnum, 8)); keep.append(t_)
for rep in range(3):
    t0 = time.perf_counter()
    sed.get_boundary_conditions(temp, cs, wz)
    t1 = time.perf_counter()
    sed.run(360.0, 2, 3600.0)
    t2 = time.perf_counter()
    sed.upward_fluxes(up)
    t3 = time.perf_counter()
    h2d = 12 * inum * jnum * 8 / 1e9
    d2h = 8 * inum * jnum * 8 / 1e9
    print(f"rep{rep}: bc {1e3*(t1-t0):.1f} ms ({h2d/(t1-t0):.1f} GB/s H2D)  run {1e3*(t2-t1):.1f} ms "
          f"(kernels {sed.info.kernel_ms:.1f})  export {1e3*(t3-t2):.1f} ms ({d2h/(t3-t2):.1f} GB/s D2H)")
