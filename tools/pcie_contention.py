#!/usr/bin/env python
"""PCIe / host-memory contention between the ranks of one box (development helper).

Launched like bench.py (python -m torch.distributed.run --nproc-per-node N tools/pcie_contention.py): every rank
copies pinned host buffers of a C4 tile's Run (12 import fields in, 8 flux fields out) to and from its GPU
  solo  : one rank at a time, the others idle
  all   : every rank at once (what a Run of the sharded bench does)
in each direction and in both directions together, and rank 0 prints per-rank and aggregate GB/s, the NUMA node
and CPU list of every GPU, and which CPUs each rank may run on."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gpu_numa(idx):
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = torch.cuda.get_device_properties(idx).pci_domain_id
        dev = torch.cuda.get_device_properties(idx).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = open(path + "/numa_node").read().strip()
        cpus = open(path + "/local_cpulist").read().strip()
        return {"pci": os.path.basename(path), "numa_node": node, "local_cpulist": cpus}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


cols = 4096 * (4096 // max(world, 1))
nin, nout = 12, 8
hin = torch.empty(nin * cols, dtype=torch.float64).pin_memory(); hin.fill_(1.0)
hout = torch.empty(nout * cols, dtype=torch.float64).pin_memory(); hout.fill_(0.0)
din = torch.empty_like(hin, device="cuda"); dout = torch.ones(nout * cols, dtype=torch.float64, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def one(kind, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s_in):
                din.copy_(hin, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s_out):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    nbytes = (hin.numel() * 8 if kind in ("h2d", "both") else 0) + (hout.numel() * 8 if kind in ("d2h", "both") else 0)
    return nbytes / best / 1e9


res = {"rank": rank, "gpu": gpu_numa(local), "cpus_allowed": sorted(os.sched_getaffinity(0))[:4] + ["..."] +
       [len(os.sched_getaffinity(0))], "mb_in": hin.numel() * 8 / 1e6, "mb_out": hout.numel() * 8 / 1e6}
one("both", 1)  # warm up
for kind in ("h2d", "d2h", "both"):
    for r in range(world):          # solo: rank r alone
        barrier()
        if r == rank:
            res[f"solo_{kind}_gbs"] = round(one(kind), 1)
        barrier()
    barrier()
    res[f"all_{kind}_gbs"] = round(one(kind), 1)
    barrier()
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    agg = {k: round(sum(o[k] for o in out), 1) for k in res if k.endswith("_gbs")}
    nodes = {}
    try:
        for n in sorted(os.listdir("/sys/devices/system/node")):
            if n.startswith("node"):
                nodes[n] = open(f"/sys/devices/system/node/{n}/cpulist").read().strip()
    except Exception as e:  # noqa: BLE001
        nodes = {"error": repr(e)}
    print(json.dumps({"world": world, "numa_nodes": nodes, "host_cpus": os.cpu_count(), "aggregate": agg, "ranks": out}, indent=1))
if world > 1:
    dist.destroy_process_group()
