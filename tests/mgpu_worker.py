"""Worker for tests/test_gpu_multi.py, launched by torch.distributed.run with one rank per GPU.
Each rank owns a j-slab; the adaptive accept flag is MAX-reduced over NCCL inside libmsed_b200.
Rank 0 also integrates the whole tile on its own GPU and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mossco_code_b200 import SedimentDriver, default_config  # noqa: E402
from mossco_code_b200.sharding import gather_slabs, init_flag_collective, local_slab, slab_bounds  # noqa: E402
from tests.cases import make_case  # noqa: E402

DT, NSTEPS = 360.0, 4


def prepared(case, j0, j1, device, bump_row):
    cfg = default_config(inum=case.inum, jnum=j1 - j0, knum=case.knum, dzmin=case.dzmin, dt_min=1.0,
                         device=device, j_offset=j0)
    sed = SedimentDriver(cfg)
    sed.init_concentrations()
    sed.set_boundary(np.asfortranarray(case.bdys[:, j0:j1]), np.asfortranarray(case.fluxes[:, j0:j1]))
    if j0 <= bump_row < j1:     # one column that violates relative_change_min on a full step
        c = sed.conc
        c[2, bump_row - j0, :, 5] *= 50.0
        c[2, bump_row - j0, :, 6] *= 0.02
        sed.conc = c
    return sed


def scenario(case, world, rank, local, bump_row, nsteps):
    """One sharded run against the whole tile on rank 0; returns (ok, text)."""
    j0, j1 = slab_bounds(case.jnum, world, rank)
    sed = prepared(case, j0, j1, local, bump_row)
    init_flag_collective(sed)
    sed.set_step_fusion("chains")   # under a collective auto mode keeps to pairs: the choice must not depend on the tile
    rc = sed.step(DT, 2, nsteps)
    mine = torch.from_numpy(np.ascontiguousarray(sed.conc)).cuda()
    sub = torch.tensor([sed.info.subcycle_warnings, sed.info.rhs_evaluations, rc, sed.info.fused_steps], device="cuda")
    subs = [torch.zeros_like(sub) for _ in range(world)]
    dist.all_gather(subs, sub)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)                  # equal slab sizes by construction
    ok, text = True, ""
    if rank == 0:
        whole = prepared(case, 0, case.jnum, local, bump_row)
        whole.set_step_fusion("chains")
        rc0 = whole.step(DT, 2, nsteps)
        got = gather_slabs([np.asfortranarray(p.cpu().numpy()) for p in parts])
        info = whole.info
        same_sub = all(int(s[0]) == info.subcycle_warnings and int(s[1]) == info.rhs_evaluations and
                       int(s[2]) == 0 and int(s[3]) == info.fused_steps for s in subs)
        ok = rc0 == 0 and same_sub and np.array_equal(got, whole.conc)
        text = (f"subcycles={info.subcycle_warnings} rhs={info.rhs_evaluations} fused_steps={info.fused_steps} "
                f"per-rank={[s.tolist() for s in subs]} bit_exact={np.array_equal(got, whole.conc)}")
        ok = ok and ((info.subcycle_warnings > 0) if bump_row >= 0 else (info.fused_steps == nsteps))
        whole.finalize()
    sed.finalize()
    return ok, text


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = make_case("mgpu", 40, 8 * world, 20, 0.003, seed=77)
    # 1. a violating column in the LAST rank's slab: every rank must reject the same attempts (the fused
    #    launch is dropped everywhere, single steps sub-cycle in lock-step)
    ok1, t1 = scenario(case, world, rank, local, case.jnum - 2, NSTEPS)
    # 2. no violation: the fused launches (chains: knum = 20) are committed on every rank
    ok2, t2 = scenario(case, world, rank, local, -1, 12)
    if rank == 0:
        print(f"MGPU world={world} rejected: {t1}")
        print(f"MGPU world={world} fused: {t2}")
        print(f"MGPU bit_exact={ok1 and ok2}")
    flag = torch.tensor([0 if (ok1 and ok2) else 1], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
