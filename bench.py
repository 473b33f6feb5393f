#!/usr/bin/env python
"""bench.py -- sediment cell-updates/s of the fused fabm_sediment RHS + adaptive-Euler integrator.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c4slab|c3|c2|c5] [--impl reference]
                  [--regime fresh|subcycling] [--tts-days D]

One "step" is one ``ode_solver(sed, dt=360 s, ode_method=2)`` call + check_NaN + clip over the whole
grid (one iteration of src/components/fabm_sediment_component.F90:1700-1769).  The default workload
is BASELINE.json's metric config: the 4096x4096x40 grid (C4), j-slab sharded over N ranks (strong
scaling: the total grid is fixed).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT = 360.0                 # examples/esmf/sediment/run_sed.nml
COUPLING_SECONDS = 3600.0  # examples/esmf/sediment/toplevel_component.F90:71 (1 h coupling)
METHOD = 2                 # ADAPTIVE_EULER, component default (:62)
JSON_OUT = sys.stdout
RAMP_SECONDS = 0.4         # untimed stepping before a short timed region so that it runs at load clocks
SEGMENT = 40               # steps between restarts from the initial state inside a long timed region
SUBCYCLING_START = 120     # --regime subcycling: untimed steps from the initial state before the timed region (the
                           # synthetic forcing enters its sub-cycling episode after ~115 steps)
FP64_INSTR_FALLBACK = {"pair_kernel": 145.0, "chain_kernel": 155.0, "column_kernel": 170.0}  # r01 ncu source pages
NVAR = 8
ROW_BLOCK = 512            # forcing is seeded per block of 512 rows so the field is independent of N

WORKLOADS = {
    # name: (inum, jnum, knum, dzmin, seed, land_fraction, description)
    "c4": (4096, 4096, 40, 0.0015, 4096, 0.0, "C4 4096x4096x40, no mask, C2-style forcing seed 4096"),
    "c4slab": (4096, 512, 40, 0.0015, 4096, 0.0, "one 4096x512x40 slab of C4 (its 8-GPU tile)"),
    "c3": (1000, 1000, 30, 0.002, 2024, 0.45, "C3 1000x1000x30, 45% land mask, smooth T, PAR"),
    "c2": (100, 100, 30, 0.002, 1234, 0.0, "C2 100x100x30 (L2-resident, launch-latency regime)"),
    "c5": (2048, 2048, 30, 0.002, 2048, 0.0, "C5 2048x2048x30 + one pelagic box per column, bed-flux "
           "exchange every 3600 s on device (msed_coupled_run)"),
}


def b_alg(knum: int) -> float:
    """Algorithmic bytes per cell-update, SURVEY.md 8(d) / BASELINE.md 2: 136 + 216/K."""
    return 2 * NVAR * 8 + 8 + 8.0 * (3 * NVAR + 3) / knum


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_figure(kernel, key):
    """A per-cell-update figure of a stepping kernel from the committed ncu captures (profiles/traffic.json:
    DRAM bytes from one ``--set full`` capture, fp64-pipe instructions from its source page), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(p))[kernel][key])
    except Exception:
        return None


def traffic_per_cell(kernel="column_kernel"):
    return profile_figure(kernel, "dram_bytes_per_cell_update")


def probe_reference_toolchain():
    """the reference's own CPU build needs gfortran, ESMF (ESMFMKFILE), a FABM
    install and MPI on the box.  Returns what was found; the gfortran arm is only possible if all are there."""
    import shutil
    found = {
        "gfortran": shutil.which("gfortran") or shutil.which("f95") or shutil.which("nvfortran") or shutil.which("flang"),
        "mpirun": shutil.which("mpirun") or shutil.which("mpiexec"),
        "ESMFMKFILE": os.environ.get("ESMFMKFILE") if os.path.exists(os.environ.get("ESMFMKFILE", "/nonexistent")) else None,
        "FABMDIR": next((d for d in (os.environ.get("FABMDIR"), os.environ.get("FABM_PREFIX"))
                         if d and os.path.isdir(d)), None),
    }
    found["buildable"] = all(found[k] for k in ("gfortran", "mpirun", "ESMFMKFILE", "FABMDIR"))
    return found


def slab_forcing(wl, j0, j1):
    """Forcing rows [j0,j1) of the workload; seeded per ROW_BLOCK so every N sees the same field."""
    from tests.cases import make_case
    inum, jnum, knum, dzmin, seed, land, _ = wl
    parts = []
    for b0 in range((j0 // ROW_BLOCK) * ROW_BLOCK, j1, ROW_BLOCK):
        b1 = min(b0 + ROW_BLOCK, jnum)
        c = make_case("blk", inum, b1 - b0, knum, dzmin, seed=seed + b0 // ROW_BLOCK, land_fraction=land,
                      smooth_temperature=land > 0, par_max=50.0 if land > 0 else 0.0)
        lo, hi = max(j0, b0) - b0, min(j1, b1) - b0
        parts.append((c.bdys[:, lo:hi], c.fluxes[:, lo:hi], c.mask[:, lo:hi], c.par_surface[:, lo:hi]))
    cat = lambda i: np.asfortranarray(np.concatenate([p[i] for p in parts], axis=1))  # noqa: E731
    return cat(0), cat(1), cat(2), cat(3)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def pin_rank_to_gpu_numa(local_rank):
    """Keep this rank -- and with it the pinned host buffers it is about to touch first -- on the NUMA node its GPU
    hangs off, when the box has more than one: eight ranks that all allocate on node 0 push every import and
    export field of a Run through one socket's memory.  Returns what was done."""
    import torch
    try:
        nodes = [n for n in os.listdir("/sys/devices/system/node") if n.startswith("node") and n[4:].isdigit()]
        p = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(path + "/numa_node").read().strip())
        if len(nodes) < 2 or node < 0:
            return {"numa_nodes": len(nodes), "gpu_numa_node": node, "pinned": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_nodes": len(nodes), "gpu_numa_node": node, "pinned": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_nodes": len(nodes), "gpu_numa_node": node, "pinned": True, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001  (sysfs layout differs between hosts: measuring must not depend on it)
        return {"pinned": False, "error": repr(e)[:80]}


def pinned_fortran(shape):
    """Pinned host buffer viewed as a Fortran-ordered numpy array of ``shape``."""
    import torch
    t = torch.empty(tuple(reversed(shape)), dtype=torch.float64).pin_memory()
    return t, t.numpy().T


# ------------------------------------------------------------------------------------------------
def cpu_reference(wl, steps, warmup, target_seconds=None, as_arm=False, fused=False, regime="fresh"):
    """The restated reference CPU path (oracle, OpenMP j-slab tiles on all host cores) timed on a
    bounded sample of the same workload.  ``fused``: the fused-loop CPU variant of the same step.  Nothing of
    the product is loaded here: the defaults come from the oracle.
    Returns (cell_updates_per_s, cores, sample_text, ms/step, sample_grid)."""
    from oracle import msed_oracle as orc
    inum, jnum, knum, dzmin, seed, land, _ = wl
    cores = os.cpu_count() or 1
    cfg = orc.default_config(inum=inum, knum=knum, dzmin=dzmin, dt_min=1.0)

    def run(rows, nsteps, conc=None):
        rows = max(cores, (rows // cores) * cores)
        cfg.jnum = rows
        bd, fl, mask, _par = slab_forcing((inum, max(rows, 1), knum, dzmin, seed, land, ""), 0, rows)
        if conc is None:
            o = orc.OracleSediment.from_config(cfg, mask2d=mask)
            o.init_concentrations()
            conc = o.conc.copy(order="F")
            o.finalize()
        secs, sub, conc = orc.bench_tiled(cfg, mask, conc, bd, fl, DT, METHOD, nsteps, cores, native=True, fused=fused)
        if secs <= 0:
            raise RuntimeError("oracle bench failed")
        return secs, rows, sub, conc

    # probe, then size the sample: ~1 s per step, at most 16 M cell-layers (the un-fused reference
    # structure needs ~400 B of temporaries per cell-layer), 3..100 steps for ~15 s in total
    secs, rows, _, _ = run(2 * cores, 3)
    rate = 3 * inum * rows * knum / secs
    rows_cap = max(cores, int(16e6 / (inum * knum)))
    want_rows = int(min(max(int(rate * 1.0 / (inum * knum)), cores), rows_cap, jnum))
    conc = None
    if as_arm:
        nsteps = steps
        if warmup:
            run(want_rows, warmup)
    else:
        step_s = inum * want_rows * knum / rate
        nsteps = int(min(max((target_seconds or 15.0) / step_s, 3), 100))
    if regime == "subcycling":   # advance (untimed) into the sub-cycling episode the GPU arm times
        _, _, _, conc = run(want_rows, SUBCYCLING_START)
    # same policy as the GPU arm: restart from the initial state every SEGMENT steps (no sub-cycling regime)
    secs, sub, left = 0.0, 0, nsteps
    while left > 0:
        m = left if regime == "subcycling" else min(SEGMENT, left)
        s_, rows, sub_, conc = run(want_rows, m, conc if regime == "subcycling" else None)
        secs, sub, left = secs + s_, sub + sub_, left - m
    n = nsteps
    cells = inum * rows * knum
    sample = (f"{inum}x{rows}x{knum} slab of the workload, {n} steps, {cores} OpenMP threads "
              f"(one j-slab tile per thread), oracle -O3 -march=native, "
              f"{'fused loop (one pass over the state per attempt)' if fused else 'un-fused whole-array passes as in the Fortran'}, "
              f"subcycles={sub}")
    return cells * n / secs, cores, sample, secs / n * 1e3, [inum, rows, knum]


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inum, jnum, knum = wl[0], wl[1], wl[2]
    value, cores, sample, ms, sample_grid = cpu_reference(wl, args.steps, args.warmup, as_arm=True, regime=args.regime)
    line = {
        "impl": "reference", "metric": "sediment cell-updates/sec", "value": value,
        "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        # the same workload keys as the B200 arm; the CPU times a bounded slab of that grid (sample_grid) and
        # the metric is a rate, so the two are compared per cell-update
        "config": {"workload": WORKLOADS[args.workload][6], "grid": [inum, jnum, knum],
                   "rows_per_gpu": jnum // max(args.gpus, 1), "dt_s": DT, "ode_method": METHOD,
                   "regime": args.regime, "sample_grid": sample_grid,
                   "note": "reference = C restatement of the Fortran CPU path (kind: port); the gfortran/ESMF/FABM "
                           "build is probed for and absent, see cpu_baseline.toolchain_probe"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                         "sample": sample, "toolchain_probe": probe_reference_toolchain()},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--regime", default="fresh", choices=["fresh", "subcycling"],
                    help="fresh (default, SURVEY 8d): the regime without rejected attempts, long timed regions restart "
                         "from the initial state every 40 steps.  subcycling: 120 untimed steps from the initial state "
                         "first, then the timed steps without restarts -- the stiff episode in which every step is "
                         "rejected at dt and runs as four quarter steps")
    ap.add_argument("--tts-days", type=float, default=0.0,
                    help="also report the time to solution of this many simulated days through the component's Run "
                         "(3600 s coupling intervals, host buffers), no restarts: config.time_to_solution")
    ap.add_argument("--fusion", default="on", choices=["on", "off", "pairs", "chains"],
                    help="speculative fused launches (msed_set_step_fusion); results are identical.  on = auto: "
                         "chains (warp per column, up to 16 steps per launch) where knum <= 32, else pairs "
                         "(thread per column, two steps per launch)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: native libraries that printf to fd 1 (NCCL's "NCCL version ..."
    # banner) are sent to stderr, the JSON line goes to the original stdout
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.component import FabmSedimentComponent
    from mossco_code_b200.sediment import PARTICULATE, VARIABLE_NAMES
    from mossco_code_b200.sharding import init_flag_collective, slab_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    numa_info = pin_rank_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    inum, jnum, knum, dzmin, seed, land, desc = wl
    j0, j1 = slab_bounds(jnum, world, rank)
    rows = j1 - j0
    bdys, fluxes, mask, par = slab_forcing(wl, j0, j1)
    cfg = default_config(inum=inum, jnum=rows, knum=knum, dzmin=dzmin, dt_min=1.0, device=local_rank,
                         j_offset=j0)
    sed = SedimentDriver(cfg)
    stream = torch.cuda.Stream()
    sed.set_stream(stream.cuda_stream)
    if land > 0:
        sed.set_mask(mask)
        sed.set_par_surface(par)
    sed.init_concentrations()
    sed.set_boundary(bdys, fluxes)
    if os.environ.get("MSED_BENCH_CHUNKS"):                 # diagnosis only: PCIe chunking of the e2e Run
        sed.set_exchange_chunks(int(os.environ["MSED_BENCH_CHUNKS"]))
    if os.environ.get("MSED_BENCH_LOCAL_ACCEPT") != "1":   # diagnosis only: per-tile accept decision
        init_flag_collective(sed)
    sed.set_step_fusion({"on": "auto"}.get(args.fusion, args.fusion))
    cells_local = inum * rows * knum * (1.0 if land == 0 else float((mask == 0).mean()))
    cells_total = torch.tensor([cells_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cells_total)
    cells_total = float(cells_total.item())

    # ---- device-resident throughput ("value") ---------------------------------------------------
    coupled = args.workload == "c5"
    steps_per_coupling = int(round(COUPLING_SECONDS / DT))
    if coupled:   # pelagic boxes: bottom-water concentrations = the forcing, 10 m boxes, w = -1e-6 m/s
        pel = np.asfortranarray(np.concatenate([np.abs(fluxes[:, :, :3]) * 1e6, bdys[:, :, 4:]], axis=2))
        wzp = np.zeros((inum, rows, NVAR), order="F"); wzp[:, :, :3] = -1.0e-6
        sed.pelagic_init(pel, wzp, np.full((inum, rows), 10.0, order="F"), np.asfortranarray(bdys[:, :, 0]))
        if args.steps % steps_per_coupling:
            raise SystemExit(f"--workload c5 needs --steps to be a multiple of {steps_per_coupling}")

    # SURVEY 8(d) quotes the metric in the regime where no attempt is rejected.  From the namelist initial
    # state the synthetic forcing enters a stiff episode after ~115 steps in which every step is rejected at dt
    # and runs as four quarter steps, so in the default regime a long timed region restarts from the initial
    # state every SEGMENT steps (the re-initialisation kernel runs inside the timed region and is counted in
    # gpu_launches).  --regime subcycling times that episode instead (no restarts).
    subcyc = args.regime == "subcycling"
    totals = dict(kernel_ms=0.0, fused_ms=0.0, kernel_launches=0, subcycle_warnings=0, rhs_evaluations=0,
                  steps_done=0, fused_pairs=0, fused_steps=0, reinits=0)

    def timed_steps(n):
        done = 0
        while done < n:
            if done and not subcyc:
                sed.init_concentrations()
                totals["reinits"] += 1
            m = min(SEGMENT, n - done)
            if coupled:
                rc = sed.coupled_run(DT, METHOD, COUPLING_SECONDS, m // steps_per_coupling)
            else:
                rc = sed.step(DT, METHOD, m)
            i = sed.info
            for k in totals:
                if k != "reinits":
                    totals[k] += getattr(i, k)
            if rc:
                return rc
            done += m
        return 0

    def to_subcycling_episode():
        """From the initial state to the start of the stiff episode, in Run-sized calls (untimed)."""
        sed.init_concentrations()
        for _ in range(SUBCYCLING_START // 10):
            if coupled:
                sed.coupled_run(DT, METHOD, COUPLING_SECONDS, 1)
            else:
                sed.step(DT, METHOD, 10)

    # W warm-up steps, then -- for workloads whose whole timed region lasts only milliseconds -- further
    # untimed steps for about RAMP_SECONDS: a B200 that sat idle while the host prepared the forcing is at its
    # idle clocks and needs tens of milliseconds under load to reach the clocks a production run sees (a C3
    # step measured 0.30 ms warm and 1.1 ms straight after start-up).  Same count on every rank (the adaptive
    # step holds a collective); the state is re-initialised before the timed region either way.
    sed.step(DT, METHOD, args.warmup)
    tw0 = time.perf_counter()            # two more untimed steps, past the first call's lazy initialisation
    sed.step(DT, METHOD, 2)
    per_step = torch.tensor([(time.perf_counter() - tw0) / 2.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(per_step, op=dist.ReduceOp.MAX)
    ramp_steps = 0
    if float(per_step.item()) * args.steps < RAMP_SECONDS and not subcyc:
        ramp_steps = int(min(4000, max(10, RAMP_SECONDS / max(float(per_step.item()), 1e-6)))) // 10 * 10
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()          # samples the ramp (same kernels, same load) and the timed region
    for r0 in range(0, ramp_steps, 10):
        if r0 % SEGMENT == 0:
            sed.init_concentrations()
        if coupled:
            sed.coupled_run(DT, METHOD, COUPLING_SECONDS, 1)
        else:
            sed.step(DT, METHOD, 10)
    if subcyc:
        to_subcycling_episode()  # 120 untimed steps: also the clock ramp of this regime
    else:
        sed.init_concentrations()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        rc = timed_steps(args.steps)
        e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    if rc != 0:
        raise SystemExit(f"msed_step returned {rc}")
    # A tiling-independent checksum of the state the timed steps left behind: the multi-GPU runs decide every
    # accept globally, so every N must print the same pair as N = 1 (msed_state_checksum)
    cs_sum, cs_xor = sed.state_checksum(global_ncol=inum * jnum)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (cs_sum, cs_xor))
        cs_sum, cs_xor = 0, 0
        for a_, b_ in parts:
            cs_sum = (cs_sum + a_) % 2 ** 64
            cs_xor ^= b_
    ms = torch.tensor([e0.elapsed_time(e1), totals["kernel_ms"], totals["fused_ms"]], dtype=torch.float64,
                      device="cuda")
    counts = torch.tensor([totals["kernel_launches"] + totals["reinits"], totals["subcycle_warnings"],
                           totals["rhs_evaluations"], totals["steps_done"], totals["fused_pairs"],
                           totals["fused_steps"]],
                          dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms, fused_ms = float(ms[0]), float(ms[1]), float(ms[2])
    launches, subcycles, rhs_evals, steps_done, fused_pairs, fused_steps = (int(x) for x in counts.tolist())
    assert steps_done == args.steps
    value = cells_total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the component's Run with HOST buffers ("e2e") -------------------------
    # Run(3600 s): H2D of the import fields (temperature + 8 surface concentrations + 3 sinking
    # velocities), get_boundary_conditions, 10 ode_solver steps, D2H of the 8 upward bed fluxes.
    comp = FabmSedimentComponent()
    comp.sed, comp.cfg = sed, cfg
    comp.mask = np.asfortranarray(mask > 0)
    comp.run_nml.update(dt=DT, ode_method=METHOD, numlayers=knum, dzmin=dzmin, dt_min=1.0)
    comp.export_3d_every_run = False   # <name>_in_soil 3-D fields: at an output cadence, see e2e_with_export
    keep, imp, exp = [], {}, {}
    t_, a_ = pinned_fortran((inum, rows)); a_[...] = bdys[:, :, 0]; keep.append(t_)
    imp["temperature_at_soil_surface"] = a_
    h2d = a_.nbytes
    for n, v in enumerate(VARIABLE_NAMES):
        t_, a_ = pinned_fortran((inum, rows)); keep.append(t_)
        if PARTICULATE[n]:
            a_[...] = -fluxes[:, :, n]                      # C_surface with w_z = 1: flux = -C*w (:1986)
            tw, aw = pinned_fortran((inum, rows)); aw[...] = 1.0; keep.append(tw)
            imp[f"{v}_z_velocity_at_soil_surface"] = aw
            h2d += aw.nbytes
        else:
            a_[...] = bdys[:, :, n + 1]
        imp[f"{v}_at_soil_surface"] = a_
        h2d += a_.nbytes
    tf_, comp.flux_buffer = pinned_fortran((inum, rows, NVAR)); keep.append(tf_)
    steps_per_run = int(round(COUPLING_SECONDS / DT))
    nruns = max(1, args.steps // steps_per_run)

    def timed_runs(n, on_run=None):
        """n Runs through the component; returns max-over-ranks wall seconds."""
        barrier()
        t0 = time.perf_counter()
        for r in range(n):
            if r and r % (SEGMENT // steps_per_run) == 0 and not subcyc:   # stay in the regime without rejections
                sed.init_concentrations()
            comp.run(imp, exp, run_seconds=COUPLING_SECONDS)
            if on_run:
                on_run(r)
        torch.cuda.synchronize()
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        return float(secs.item())

    if subcyc:
        to_subcycling_episode()
    else:
        sed.init_concentrations()
    # The three sinking velocities are constants of the synthetic pelagic side, as they are in most pelagic
    # models: the coupler declares them static (component.static_import_suffixes -> msed_set_import_generations)
    # and they cross PCIe once, in the warm-up Run; the other nine fields are uploaded every Run.
    # e2e_all_fields below is the same measurement with every field uploaded every Run.
    static_suffix = "_z_velocity_at_soil_surface"
    h2d_static = sum(a.nbytes for k, a in imp.items() if k.endswith(static_suffix))
    # (a one-chunk tile with pinned fields is read by the boundary kernel straight through PCIe -- nothing is staged,
    #  so there is nothing to keep on the device and every field crosses every Run)
    small_tile = inum * rows < (1 << 17)
    if small_tile:
        h2d_static = 0
    else:
        comp.static_import_suffixes = (static_suffix,)
    comp.run(imp, exp, run_seconds=COUPLING_SECONDS)        # warm-up Run
    e2e_s = timed_runs(nruns)
    d2h = sum(exp[f"{v}_upward_flux_at_soil_surface"].nbytes for v in VARIABLE_NAMES)
    e2e_value = cells_total * nruns * steps_per_run / e2e_s
    # where the last timed Run spent its time on each rank's device (msed_get_exchange_timing): with transfers fully
    # hidden the span is the kernels' time; the excess is exposed PCIe traffic (profiles/r02_n8_pcie_contention.json)
    ph = [round(x, 3) for x in sed.exchange_timing()]
    phases = [ph]
    if world > 1:
        phases = [None] * world
        dist.all_gather_object(phases, ph)
    e2e_phases = {"wall_ms_per_run": round(1e3 * e2e_s / nruns, 3),
                  "per_rank_ms": {"last_h2d_landed": [p_[0] for p_ in phases], "last_kernel_done": [p_[1] for p_ in phases],
                                  "d2h_tail_after_kernels": [p_[2] for p_ in phases], "device_span": [p_[3] for p_ in phases]},
                  "of": "the last timed Run, ms since its first import copy was issued"}
    comp.static_import_suffixes = ()
    sed.set_import_generations(None)
    if subcyc:
        to_subcycling_episode()
    else:
        sed.init_concentrations()
    comp.run(imp, exp, run_seconds=COUPLING_SECONDS)
    e2e_all_value = cells_total * nruns * steps_per_run / timed_runs(nruns)

    # ---- the same with the <name>_in_soil write-back at an output cadence ---------------------------------
    # The reference copies all nvar 3-D states to their fields every Run (:1773-1822); consumers read them at the
    # output cadence.  Here: EXPORT_CADENCE Runs, one asynchronous export of the whole state started after the
    # first of them (snapshot on the device, PCIe copy under the following Runs), completed inside the timed
    # region.  Needs a pinned host buffer of the state's size: measured when that is <= 12 GB per rank (C4 on 8
    # GPUs, the slab, C2, C3, C5) or MSED_BENCH_EXPORT=1.
    EXPORT_CADENCE = 24
    state_bytes = inum * rows * knum * NVAR * 8
    e2e_export = None
    if not coupled and (state_bytes <= 12e9 or os.environ.get("MSED_BENCH_EXPORT") == "1"):
        tb_, comp.export_buffer = pinned_fortran((inum, rows, knum, NVAR)); keep.append(tb_)
        sed.export_state_begin(comp.export_buffer); sed.export_state_wait()   # untimed: allocates the snapshot buffer, creates the stream
        if not subcyc:
            sed.init_concentrations()
        comp.export_cadence, comp._runs = EXPORT_CADENCE, EXPORT_CADENCE - 1    # the first timed Run is a cadence Run
        secs = timed_runs(EXPORT_CADENCE, on_run=lambda r: comp.export_ready(exp) if r == EXPORT_CADENCE - 1 else None)
        comp.export_cadence = 0
        e2e_export = {"value": cells_total * EXPORT_CADENCE * steps_per_run / secs, "unit": "cell-updates/s",
                      "cadence_runs": EXPORT_CADENCE, "export_bytes_per_cadence": state_bytes,
                      "d2h_bytes_per_step": (d2h * EXPORT_CADENCE + state_bytes) / (EXPORT_CADENCE * steps_per_run)}

    # ---- time to solution of a simulated period through the component (no restarts) ---------------------
    tts = None
    if args.tts_days > 0 and not coupled:
        sed.init_concentrations()
        n_tts = int(round(args.tts_days * 86400.0 / COUPLING_SECONDS))
        agg = dict(sub=0, rhs=0, fused=0, launches=0)

        def acc(r):
            i = comp.last_info
            agg["sub"] += i.subcycle_warnings; agg["rhs"] += i.rhs_evaluations
            agg["fused"] += i.fused_steps; agg["launches"] += i.kernel_launches
        saved, subcyc = subcyc, True          # (no restarts inside timed_runs)
        secs = timed_runs(n_tts, on_run=acc)
        subcyc = saved
        tts = {"simulated_days": args.tts_days, "runs": n_tts, "steps": n_tts * steps_per_run, "seconds": secs,
               "subcycles": agg["sub"], "rhs_evaluations": agg["rhs"], "fused_steps": agg["fused"],
               "gpu_launches": agg["launches"],
               "call": f"{n_tts} x FabmSedimentComponent.run({int(COUPLING_SECONDS)} s), host buffers"}

    if rank == 0:
        from mossco_code_b200 import measure_fp64_peak
        peak, peak_src = measured_peak()
        fp64_peak = measure_fp64_peak(local_rank)           # TFLOP/s (2 x FMA/s), measured now on this GPU
        balg = b_alg(knum)
        cells_per_gpu = cells_total / world
        accepts_per_step = rhs_evals / max(args.steps, 1)   # attempts per ode_solver call (1 without sub-cycling)
        if fused_pairs > 0 and fused_steps >= args.steps // 2:
            # accepted sub-steps (= RHS evaluations that advance the state) per launch: 2 for a pair, up to 16
            # for a chain
            subs_total = fused_steps * (rhs_evals - subcycles) / max(steps_done, 1)
            m = subs_total / fused_pairs
            kname = "chain_kernel" if m > 2.01 else "pair_kernel"
            kernel_name = (f"msed::{kname}<OMEXDIA_P, adaptive> ({m:.1f} accepted sub-steps per launch on average)")
            cell_updates_per_launch = m * cells_per_gpu     # RHS evaluation + update of one cell-layer = one unit
            avg_launch_s = fused_ms * 1e-3 / fused_pairs
            tpc = traffic_per_cell(kname)
            fpi = profile_figure(kname, "fp64_instr_per_cell_update") or FP64_INSTR_FALLBACK[kname]
            achieved_tf = 2.0 * fpi * cell_updates_per_launch / avg_launch_s / 1e12
            fused_bytes = (2 * NVAR * 8 + 8.0 * (3 * NVAR + 3) / knum) / m   # state once per launch, closed-form porosity
            hbm_fused = fused_bytes * cell_updates_per_launch / avg_launch_s / 1e9
            per_step_alg = balg * cell_updates_per_launch / avg_launch_s / 1e9
            roof = {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                    "traffic": None if tpc is None else tpc * cell_updates_per_launch,
                    "kernel": kernel_name, "cell_updates_per_launch": cell_updates_per_launch,
                    "avg_launch_ms": avg_launch_s * 1e3,
                    "fp64_instr_per_cell_update": fpi,
                    "peak_source": "msed_measure_fp64_peak in this process (DFMA micro-kernel, 2 flop per FMA); "
                                   "achieved counts every fp64-pipe instruction of the launch as one FMA slot "
                                   "(ncu source page of the kernel, profiles/traffic.json)",
                    "hbm": {"achieved": hbm_fused, "peak": peak, "unit": "GB/s", "frac": hbm_fused / peak,
                            "algorithmic_bytes_per_cell_update": fused_bytes, "peak_source": peak_src,
                            "note": "algorithmic bytes of the FUSED launch: the state is read and written once per "
                                    "launch, (128 + 216/K)/m per cell-update"},
                    "frac_per_step_bytes": per_step_alg / peak,
                    "per_step_algorithmic_bytes": balg,
                    "achieved_dram": None if tpc is None else tpc * cell_updates_per_launch / avg_launch_s / 1e9,
                    "fused_launches": fused_pairs, "fused_steps": fused_steps,
                    "note": "fused launches advance several sub-steps per HBM round trip and are bound by the fp64 "
                            "pipe, not by HBM; frac_per_step_bytes is SURVEY 8(d)'s per-step figure (136 + 216/K B "
                            "per cell-update) over the measured copy bandwidth and exceeds 1 for that reason; "
                            "--fusion off times the single-step HBM-bound kernel"}
        else:
            kernel_name = "msed::column_kernel<OMEXDIA_P, OP_ADAPTIVE>"
            cell_updates_per_launch = cells_per_gpu
            avg_launch_s = kernel_ms * 1e-3 / max(rhs_evals, 1)
            tpc = traffic_per_cell("column_kernel")
            achieved = balg * cell_updates_per_launch / avg_launch_s / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None if tpc is None else tpc * cell_updates_per_launch,
                    "achieved_dram": None if tpc is None else tpc * cell_updates_per_launch / avg_launch_s / 1e9,
                    "kernel": kernel_name, "cell_updates_per_launch": cell_updates_per_launch,
                    "algorithmic_bytes_per_cell_update": balg, "peak_source": peak_src,
                    "avg_launch_ms": avg_launch_s * 1e3, "fp64_peak_tflops": fp64_peak,
                    "note": "single-step kernel: state read once and written once per attempt"}
        line = {
            "metric": "sediment cell-updates/sec", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "grid": [inum, jnum, knum], "rows_per_gpu": rows, "dt_s": DT,
                       "ode_method": METHOD, "regime": args.regime, "sharding": f"j-slabs x{world}, no halo",
                       "step_fusion": args.fusion,
                       "l2": "per-GPU state >= 5.4 GB >> 126 MB L2 (inputs larger than L2)"
                       if cells_per_gpu * 128 > 1e9 else "state fits L2 (launch-latency regime)",
                       "subcycles_in_timed_region": subcycles, "rhs_evaluations": rhs_evals,
                       "attempts_per_step": accepts_per_step,
                       "state_restarts_in_timed_region": totals["reinits"],
                       "clock_ramp_steps_after_warmup": SUBCYCLING_START if subcyc else ramp_steps,
                       "restart_every_steps": None if subcyc else SEGMENT,
                       "state_checksum": {"sum_mod_2_64": f"{cs_sum:016x}", "xor": f"{cs_xor:016x}",
                                          "of": "state after the timed steps, msed_state_checksum over all tiles: "
                                                "identical for every N"},
                       "host_numa": numa_info,
                       "e2e_call": f"FabmSedimentComponent.run({int(COUPLING_SECONDS)} s) -> msed_run_exchange: H2D of 9 "
                                   f"pinned import fields (12 in e2e_all_fields) + get_boundary_conditions + {steps_per_run} ode_solver "
                                   f"steps + D2H of 8 upward-flux fields (chunk-major: every chunk runs the whole "
                                   f"interval as soon as its fields have landed); {nruns} timed Run(s); WITHOUT the "
                                   f"3-D <name>_in_soil write-back, which e2e_with_export adds at its cadence"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "cell-updates/s",
                    "h2d_bytes_per_step": (h2d - h2d_static) / steps_per_run, "d2h_bytes_per_step": d2h / steps_per_run,
                    "h2d_bytes_per_run": h2d - h2d_static, "d2h_bytes_per_run": d2h, "steps_per_run": steps_per_run,
                    "includes_3d_export": False,
                    "static_import_fields": ("none: one-chunk tile, pinned fields read and written in place through PCIe "
                                             "by the boundary / export kernels (zero-copy)") if small_tile else
                                            ("the 3 *_z_velocity_at_soil_surface fields (constant sinking speeds) are "
                                             "declared static and uploaded once, outside the timed Runs")},
            "e2e_phases": e2e_phases,
            "e2e_all_fields": {"value": e2e_all_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / steps_per_run,
                               "d2h_bytes_per_step": d2h / steps_per_run,
                               "note": "every one of the 12 import fields uploaded every Run"},
            "e2e_with_export": e2e_export,
            "gpu_launches": launches,
            "roofline": roof,
        }
        if tts is not None:
            line["config"]["time_to_solution"] = tts
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _, sgrid = cpu_reference(wl, 0, 0, target_seconds=12.0, regime=args.regime)
            vf, _, samplef, _, _ = cpu_reference(wl, 0, 0, target_seconds=6.0, fused=True, regime=args.regime)
            line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                    "sample": sample, "sample_grid": sgrid,
                                    "fused_loop_variant": {"value": vf, "unit": "cell-updates/s", "cores": cores,
                                                           "kind": "port", "sample": samplef},
                                    "toolchain_probe": probe_reference_toolchain()}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    sed.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
