"""Multi-GPU path: j-slab tiles + NCCL MAX-reduction of the accept flag inside the library reproduce the
single-GPU run of the whole tile bit for bit (needs >= 2 GPUs on the box; the CPU-side protocol is
covered by tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_run_is_bit_exact(gpu):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "bit_exact=True" in res.stdout


@pytest.mark.parametrize("mode,bump", [("pairs", True), ("pairs", False), ("chains", True), ("off", True)])
def test_two_tiles_on_one_gpu_share_the_accept_decision(gpu, mode, bump):
    """The global accept logic (solver_library.F90:121 is a whole-domain any()) on ONE device: two handles own the
    two j-slabs of a tile, each driven by its own host thread, and msed_set_allreduce_hook MAX-reduces their
    flags through the host at every decision point -- what NCCL does between ranks.  The tiles together must
    reproduce the single-handle run of the whole tile bit for bit, including the run in which only ONE tile holds
    the column that forces sub-cycling, fused groups that plan rejected attempts, and the sub-cycle counters."""
    import ctypes as C
    import threading

    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sharding import gather_slabs
    from tests.cases import make_case
    cudart = None
    for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            cudart = C.CDLL(name)
            break
        except OSError:
            continue
    if cudart is None:
        import glob
        import torch
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        cudart = C.CDLL(cands[0])
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    cudart.cudaStreamSynchronize.argtypes = [C.c_void_p]
    H2D, D2H = 1, 2
    case = make_case("twotiles", 24, 12, 20, 0.003, seed=78)
    nsteps, calls = 6, 3

    def prepared(j0, j1):
        cfg = default_config(inum=24, jnum=j1 - j0, knum=20, dzmin=0.003, dt_min=1.0, j_offset=j0)
        sed = SedimentDriver(cfg)
        sed.init_concentrations()
        sed.set_boundary(np.asfortranarray(case.bdys[:, j0:j1]), np.asfortranarray(case.fluxes[:, j0:j1]))
        if bump and j0 <= 9 < j1:       # one column of the SECOND tile violates relative_change_min on a full step
            c = sed.conc
            c[2, 9 - j0, :, 5] *= 50.0
            c[2, 9 - j0, :, 6] *= 0.02
            sed.conc = c
        sed.set_step_fusion(mode)
        return sed

    whole = prepared(0, 12)
    tot = dict(sub=0, rhs=0, fused=0)
    for _ in range(calls):
        assert whole.step(360.0, 2, nsteps) == 0
        tot["sub"] += whole.info.subcycle_warnings
        tot["rhs"] += whole.info.rhs_evaluations
        tot["fused"] += whole.info.fused_steps
    assert (tot["sub"] > 0) == bump

    tiles = [prepared(0, 6), prepared(6, 12)]
    barrier = threading.Barrier(2)
    box = [None, None]
    nred = [0, 0]

    def make_hook(r):
        def hook(ptr, count, stream):
            cudart.cudaStreamSynchronize(C.c_void_p(stream))
            mine = (C.c_int32 * count)()
            cudart.cudaMemcpy(mine, C.c_void_p(ptr), 4 * count, D2H)
            box[r] = list(mine)
            barrier.wait(timeout=60)
            other = box[1 - r]
            assert len(other) == count          # both tiles reach the same reduction
            red = (C.c_int32 * count)(*[max(a, b) for a, b in zip(mine, other)])
            barrier.wait(timeout=60)            # both have read before anyone overwrites
            cudart.cudaMemcpy(C.c_void_p(ptr), red, 4 * count, H2D)
            nred[r] += 1
            return 0
        return hook

    got = [dict(sub=0, rhs=0, fused=0, rc=0) for _ in tiles]
    errs = []

    def drive(r):
        try:
            for _ in range(calls):
                rc = tiles[r].step(360.0, 2, nsteps)
                got[r]["rc"] |= rc
                got[r]["sub"] += tiles[r].info.subcycle_warnings
                got[r]["rhs"] += tiles[r].info.rhs_evaluations
                got[r]["fused"] += tiles[r].info.fused_steps
        except Exception as exc:   # a dead thread must not leave the other one at the barrier
            errs.append(exc)
            barrier.abort()

    for r, t in enumerate(tiles):
        t.set_allreduce_hook(make_hook(r))
    th = [threading.Thread(target=drive, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    assert nred[0] == nred[1] > 0
    for r in range(2):
        assert got[r]["rc"] == 0 and got[r]["sub"] == tot["sub"] and got[r]["rhs"] == tot["rhs"], (r, got[r], tot)
    assert np.array_equal(gather_slabs([t.conc for t in tiles]), whole.conc)
    assert np.array_equal(gather_slabs([t.fluxes for t in tiles]), whole.fluxes)
    # the tiling-independent checksum bench.py prints for every N
    w = whole.state_checksum(global_ncol=24 * 12, col_offset=0)
    parts = [t.state_checksum(global_ncol=24 * 12) for t in tiles]
    assert ((parts[0][0] + parts[1][0]) % 2 ** 64, parts[0][1] ^ parts[1][1]) == w
    for t in tiles + [whole]:
        t.finalize()
