"""N>1 host logic on CPU: j-slab partitioning and the one collective on the path -- the MAX reduction of
the adaptive-step accept flag (solver_library.F90:121) -- exercised with world_size=2 over gloo.

The compute inside each rank is the ORACLE (this is a test; the product never computes on the CPU): two
ranks step their slabs with the global decision taken through the product's
``sharding.reduce_flags_max`` and must reproduce the single-domain oracle run bit for bit, which
per-tile decisions (what the reference does under MPI) do not."""
import os
import tempfile

import numpy as np
import pytest

from tests.cases import make_case

DT = 360.0
NSTEPS = 4


def test_slab_bounds_partition():
    from mossco_code_b200.sharding import gather_slabs, local_slab, slab_bounds
    for jnum in (2, 7, 512, 4096, 1000):
        for world in (1, 2, 3, 4, 8):
            if jnum < world:
                with pytest.raises(ValueError):
                    slab_bounds(jnum, world, 0)
                continue
            b = [slab_bounds(jnum, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == jnum
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [j1 - j0 for j0, j1 in b]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    a = np.asfortranarray(np.arange(3 * 7 * 5, dtype=float).reshape(3, 7, 5))
    parts = [local_slab(a, 3, r) for r in range(3)]
    assert all(p.flags.f_contiguous for p in parts)
    assert np.array_equal(gather_slabs(parts), a)
    with pytest.raises(ValueError):
        slab_bounds(8, 2, 2)


def _case():
    """6x8 tile; one column in the upper half (rank 1's slab) is loaded with ammonium and oxygen so that
    only it violates relative_change_min on a full 360 s step."""
    case = make_case("shard", 6, 8, 12, 0.004, seed=12)
    return case


def _prepare(orc, case, j0, j1):
    from mossco_code_b200.sharding import slab_bounds  # noqa: F401
    sub_b = np.asfortranarray(case.bdys[:, j0:j1])
    sub_f = np.asfortranarray(case.fluxes[:, j0:j1])
    par = orc.omexdia_params()
    o = orc.OracleSediment(case.inum, j1 - j0, case.knum, case.dzmin, params=par, dt_min=1.0)
    o.init_concentrations()
    o.set_boundary(sub_b, sub_f)
    c = o.conc
    if j0 <= 6 < j1:            # global row 6 lives in the last slab
        c[2, 6 - j0, :, 5] *= 50.0     # nh3
        c[2, 6 - j0, :, 6] *= 0.02     # oxy: nitrification eats >90 % of it in one step
    return o


def _adaptive_step(o, dt, dt_min, fac, reduce):
    """ode_solver ADAPTIVE_EULER (solver_library.F90:104-140) with the any() taken through `reduce`."""
    import torch
    dt_int, dt_red, sub = 0.0, dt, 0
    while dt_int < dt:
        rhs = o.get_rhs()
        c1 = o.conc + dt_red * rhs
        flags = torch.tensor([int(np.any(c1 - fac * o.conc < 0.0)), 0], dtype=torch.int32)
        reduce(flags)
        if int(flags[0]) and dt_red > dt_min:
            dt_red *= 0.25
            sub += 1
        else:
            o.conc[...] = c1
            dt_int += dt_red
    return sub


def _worker(rank, world, tmp, use_global):
    import torch.distributed as dist

    from mossco_code_b200.sharding import reduce_flags_max, slab_bounds
    from oracle import msed_oracle as orc
    dist.init_process_group("gloo", init_method=f"file://{tmp}/rdzv_{int(use_global)}", rank=rank,
                            world_size=world)
    case = _case()
    j0, j1 = slab_bounds(case.jnum, world, rank)
    o = _prepare(orc, case, j0, j1)
    reduce = reduce_flags_max if use_global else (lambda f: f)
    sub = 0
    for _ in range(NSTEPS):
        sub += _adaptive_step(o, DT, 1.0, 1.0 + (-0.9), reduce)
        np.maximum(o.conc, 0.0, out=o.conc)           # clip to minimum, component :1728
    np.save(os.path.join(tmp, f"conc_{int(use_global)}_{rank}.npy"), o.conc)
    np.save(os.path.join(tmp, f"sub_{int(use_global)}_{rank}.npy"), np.array(sub))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_global_accept_flag_matches_single_domain(oracle):
    import torch.multiprocessing as mp

    from mossco_code_b200.sharding import gather_slabs
    case = _case()
    ref = _prepare(oracle, case, 0, case.jnum)
    assert ref.step(DT, 2, NSTEPS) == 0
    ref_sub = ref.solver_diag()["subcycles"]
    assert ref_sub > 0, "the crafted column must force sub-cycling"
    with tempfile.TemporaryDirectory() as tmp:
        for use_global in (True, False):
            mp.spawn(_worker, args=(2, tmp, use_global), nprocs=2, join=True)
        glob = gather_slabs([np.load(os.path.join(tmp, f"conc_1_{r}.npy")) for r in range(2)])
        loc = gather_slabs([np.load(os.path.join(tmp, f"conc_0_{r}.npy")) for r in range(2)])
        sub_g = [int(np.load(os.path.join(tmp, f"sub_1_{r}.npy"))) for r in range(2)]
        sub_l = [int(np.load(os.path.join(tmp, f"sub_0_{r}.npy"))) for r in range(2)]
    # one flag all-reduce per attempt reproduces the single-domain run exactly
    assert sub_g == [ref_sub, ref_sub]
    assert np.array_equal(glob, ref.conc)
    # per-tile decisions (no collective): rank 0 never sub-cycles, so its slab differs
    assert sub_l[0] == 0 and sub_l[1] == ref_sub
    assert not np.array_equal(loc[:, :4], ref.conc[:, :4])
    assert np.array_equal(loc[:, 4:], ref.conc[:, 4:])


@pytest.mark.timeout(120)
def test_broadcast_bytes_gloo():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_bcast_worker, args=(2, tmp), nprocs=2, join=True)
        got = [open(os.path.join(tmp, f"id_{r}"), "rb").read() for r in range(2)]
    assert got[0] == got[1] == bytes(range(128))


def _bcast_worker(rank, world, tmp):
    import torch.distributed as dist

    from mossco_code_b200.sharding import broadcast_bytes
    dist.init_process_group("gloo", init_method=f"file://{tmp}/rdzv", rank=rank, world_size=world)
    payload = bytes(range(128)) if rank == 0 else b""
    out = broadcast_bytes(payload, 0)
    open(os.path.join(tmp, f"id_{rank}"), "wb").write(out)
    dist.barrier()
    dist.destroy_process_group()
