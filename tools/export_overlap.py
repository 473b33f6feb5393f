#!/usr/bin/env python
"""Does the asynchronous state export hide under the Runs that follow it?  (development helper)
usage: python tools/export_overlap.py [workload]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, pinned_fortran, slab_forcing  # noqa: E402
from mossco_code_b200 import SedimentDriver, default_config  # noqa: E402
from mossco_code_b200.sediment import PARTICULATE  # noqa: E402

wlname = sys.argv[1] if len(sys.argv) > 1 else "c4slab"
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
    a[...] = -fluxes[:, :, n] if PARTICULATE[n] else bdys[:, :, n + 1]
    cs.append(a)
    if PARTICULATE[n]:
        t_, w = pinned_fortran((inum, jnum)); w[...] = 1.0; keep.append(t_); wz.append(w)
    else:
        wz.append(None)
t_, up = pinned_fortran((inum, jnum, 8)); keep.append(t_)
t_, state = pinned_fortran((inum, jnum, knum, 8)); keep.append(t_)


def runs(n):
    for _ in range(n):
        sed.run_exchange(360.0, 2, 3600.0, temp, cs, wz, out=up)


runs(2)
sed.export_state_begin(state); sed.export_state_wait()          # first call allocates the snapshot
for rep in range(2):
    sed.init_concentrations()
    torch.cuda.synchronize(); t0 = time.perf_counter(); runs(12); torch.cuda.synchronize(); t_runs = time.perf_counter() - t0
    t0 = time.perf_counter(); sed.export_state_begin(state); t_begin = time.perf_counter() - t0
    sed.export_state_wait(); t_exp = time.perf_counter() - t0
    sed.init_concentrations()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    runs(1); sed.export_state_begin(state); t1 = time.perf_counter(); runs(11); t2 = time.perf_counter(); sed.export_state_wait()
    torch.cuda.synchronize(); t_both = time.perf_counter() - t0
    # the same with device-resident steps instead of Runs (no PCIe traffic of their own)
    sed.init_concentrations()
    torch.cuda.synchronize(); t0 = time.perf_counter(); sed.step(360.0, 2, 110); torch.cuda.synchronize(); t_steps = time.perf_counter() - t0
    sed.init_concentrations(); sed.step(360.0, 2, 10)
    torch.cuda.synchronize(); sed.export_state_begin(state); t0 = time.perf_counter(); sed.step(360.0, 2, 110); torch.cuda.synchronize()
    t_steps_exp = time.perf_counter() - t0; sed.export_state_wait()
    print(f"      110 device-resident steps {1e3*t_steps:.1f} ms, during an export {1e3*t_steps_exp:.1f} ms")
    # per-call wall times of 11 calls of 10 device-resident steps, without and with a pending export
    for pend in (False, True):
        sed.init_concentrations(); sed.step(360.0, 2, 10); torch.cuda.synchronize()
        if pend:
            sed.export_state_begin(state)
        ts = []
        for _ in range(11):
            t0 = time.perf_counter(); sed.step(360.0, 2, 10); ts.append(1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter(); sed.export_state_wait(); tw = 1e3 * (time.perf_counter() - t0)
        print(f"      step(10) x 11, export pending={pend}: " + " ".join(f"{t:.1f}" for t in ts) + f" ms; wait {tw:.1f} ms")
    for pend in (False, True):
        sed.init_concentrations(); runs(1); torch.cuda.synchronize()
        if pend:
            sed.export_state_begin(state)
        ts = []
        for _ in range(11):
            t0 = time.perf_counter(); runs(1); ts.append((1e3 * (time.perf_counter() - t0), [round(x, 1) for x in sed.exchange_timing()]))
        t0 = time.perf_counter(); sed.export_state_wait(); tw = 1e3 * (time.perf_counter() - t0)
        print(f"      Run x 11, export pending={pend}: " + " ".join(f"{t:.1f}{ph}" for t, ph in ts) + f" ms; wait {tw:.1f} ms")
    gb = state.nbytes / 1e9
    print(f"rep{rep}: 12 Runs {1e3*t_runs:.1f} ms | export alone {1e3*t_exp:.1f} ms ({gb/t_exp:.1f} GB/s, begin() returns after "
          f"{1e3*t_begin:.1f} ms) | 12 Runs with the export started after the first {1e3*t_both:.1f} ms "
          f"(11 Runs after begin: {1e3*(t2-t1):.1f} ms)")
