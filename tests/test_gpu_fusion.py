"""Step fusion: the speculative fused launches -- pairs (msed_pair.cuh: thread per column, two steps) and
chains (msed_chain.cuh: warp per column, state in registers, up to 16 steps) -- must be invisible in the
results: bit-identical state, bed fluxes, diagnostics, sub-cycle counts and NaN behaviour with fusion on
or off, and must actually be used (fewer launches) in the regime they are meant for."""
import numpy as np
import pytest

from tests.cases import make_case, rel_err

pytestmark = pytest.mark.gpu
DT = 360.0


def _run(case, fusion, method, nsteps, calls=1, mutate=None, **cfgkw):
    # Runge-Kutta with a thread per column: "quad" = four stages per launch (rk_quad_kernel), "pairs" = two
    rk_stages = 4 if fusion == "quad" else 2
    if fusion == "quad":
        fusion = "pairs"
    from mossco_code_b200 import SedimentDriver, default_config
    kw = dict(inum=case.inum, jnum=case.jnum, knum=case.knum, dzmin=case.dzmin, dt_min=1.0)
    kw.update(cfgkw)
    cfg = default_config(**kw)
    with SedimentDriver(cfg) as sed:
        sed.set_step_fusion(fusion)
        sed.set_rk_stages_per_launch(rk_stages)
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        if mutate:
            mutate(sed)
        launches = sub = rhs = done = fused = 0
        rc = 0
        for _ in range(calls):
            rc = sed.step(DT, method, nsteps)
            launches += sed.info.kernel_launches
            fused += sed.info.fused_steps
            sub += sed.info.subcycle_warnings
            rhs += sed.info.rhs_evaluations
            done += sed.info.steps_done
            if rc:
                break
        return dict(conc=sed.conc, fluxes=sed.fluxes, denit=sed.field("denit"), launches=launches, sub=sub,
                    rhs=rhs, done=done, rc=rc, fused=fused)


def _same(a, b):
    assert a["rc"] == b["rc"] and a["done"] == b["done"] and a["sub"] == b["sub"] and a["rhs"] == b["rhs"]
    assert np.array_equal(a["conc"], b["conc"], equal_nan=True)
    assert np.array_equal(a["fluxes"], b["fluxes"], equal_nan=True)
    assert np.array_equal(a["denit"], b["denit"], equal_nan=True)


MODES = ["pairs", "chains"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("method", [2, 0])
@pytest.mark.parametrize("nsteps", [2, 3, 4, 10, 11])
def test_fusion_is_bit_identical(gpu, oracle, method, nsteps, mode):
    case = make_case("fuse", 37, 21, 20, 0.003, seed=101, land_fraction=0.2, smooth_temperature=True)
    on = _run(case, mode, method, nsteps, calls=2)
    off = _run(case, False, method, nsteps, calls=2)
    _same(on, off)
    assert on["launches"] < off["launches"]                 # pairs were really used
    ref = oracle.OracleSediment(37, 21, 20, 0.003, mask2d=case.mask, dt_min=1.0)
    ref.init_concentrations(); ref.set_boundary(case.bdys, case.fluxes)
    assert ref.step(DT, method, 2 * nsteps) == 0
    wet = case.mask == 0
    assert rel_err(on["conc"][wet], ref.conc[wet]) <= 1e-11


@pytest.mark.parametrize("kw", [dict(bcup_dissolved_variables=1), dict(bioturbation_profile=2),
                                dict(bcup_dissolved_variables=3, bioturbation_profile=0),
                                dict(minimum=[1., 2., 3., 0.5, 30., 1., 2., 150.]), dict(model=1)])
@pytest.mark.parametrize("mode", MODES)
def test_fusion_variants(gpu, kw, mode):
    case = make_case("fusev", 19, 9, 15, 0.004, seed=7)

    def mutate(sed):
        if kw.get("bcup_dissolved_variables") == 1:
            fl = case.fluxes.copy(); fl[:, :, 3:] = 1e-6 * (1 + np.arange(5))
            sed.set_boundary(None, fl)
        sed.update_porosity(0.5 + 0.3 * np.random.default_rng(2).random((19, 9)))   # porosity mode 2

    _same(_run(case, mode, 2, 9, mutate=mutate, **kw), _run(case, False, 2, 9, mutate=mutate, **kw))


@pytest.mark.parametrize("method", [2, 0])
@pytest.mark.parametrize("knum,dzmin,nsteps", [(2, 0.05, 7), (3, 0.03, 5), (31, 0.002, 16), (32, 0.002, 17), (30, 0.002, 33),
                                               (33, 0.002, 5), (40, 0.0015, 17), (63, 0.0005, 4), (64, 0.0005, 10)])
def test_chain_layer_counts_and_lengths(gpu, knum, dzmin, nsteps, method):
    """Chains at the edges of their range: the thinnest columns the driver accepts (knum = 2: every layer
    touches a boundary), a full warp (knum = 32), idle lanes (knum < 32), two layers per lane (knum = 33: the
    second layer of lane 16 is a shadow; 40; 63; 64: every lane full), and step counts that split into one, two and
    three launches."""
    case = make_case("chk", 21, 13, knum, dzmin, seed=40 + knum, land_fraction=0.15)
    on = _run(case, "chains", method, nsteps, calls=2)
    off = _run(case, False, method, nsteps, calls=2)
    _same(on, off)
    if off["sub"] == 0:      # (the thinnest layers sub-cycle irregularly at this dt: then only the bits are compared)
        assert on["launches"] < off["launches"]
    else:
        assert knum >= 63


def test_chain_is_the_default_on_small_tiles(gpu):
    """auto mode: chains on tiles too small for a thread per column (fewer launches than pairs for the same call),
    one layer per lane up to 32 layers, two above."""
    small = make_case("chd", 16, 8, 30, 0.002, seed=3)
    deep = make_case("chd", 16, 8, 40, 0.0015, seed=3)
    for case in (small, deep):
        assert _run(case, "auto", 2, 10)["launches"] == _run(case, "chains", 2, 10)["launches"] \
            < _run(case, "pairs", 2, 10)["launches"]
    _same(_run(deep, "auto", 2, 10), _run(deep, False, 2, 10))
    _same(_run(deep, "pairs", 2, 10), _run(deep, False, 2, 10))


@pytest.mark.parametrize("mode", MODES)
def test_fusion_falls_back_when_a_step_is_rejected(gpu, mode):
    """rnit/rODUox boosted: the first steps sub-cycle.  A pair containing a rejected step is not committed;
    the single-step path redoes it, and fusion stays off for a while afterwards."""
    case = make_case("fuser", 12, 8, 15, 0.004, seed=2)
    kw = dict(rnit=2.0e3, rODUox=2.0e3)
    on = _run(case, mode, 2, 7, calls=3, **kw)
    off = _run(case, False, 2, 7, calls=3, **kw)
    assert off["sub"] > 0
    _same(on, off)


# per-step rejection counts of these regimes on this tile (oracle, 60 steps):
#   rnit=rODUox=600     110111111111...           steady sub-cycling at dt/4
#   rnit=rODUox=4000    2222222222222222100100200120...   steady dt/16, then a mix
#   rnit=rODUox=2000    2111111111111111111110000000010101000100...   an episode that ends
#   rLabile=0.6         0000000111212221212121212...   period-2 pattern: every prediction fails
#   rnit=rODUox=2e4     3333003000300030...        deeper than the fused kernels plan (dt/64)
SUBCYCLING = [dict(rnit=600., rODUox=600.), dict(rnit=4000., rODUox=4000.), dict(rnit=2000., rODUox=2000.),
              dict(rLabile=0.6), dict(rnit=2.0e4, rODUox=2.0e4)]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("nsteps,calls", [(10, 5), (3, 9), (1, 12)])
@pytest.mark.parametrize("kw", SUBCYCLING)
def test_subcycling_regime_is_fused_and_bit_identical(gpu, kw, nsteps, calls, mode):
    """Sub-cycled steps (solver_library.F90:126-128) go through the fused kernels too: a call is planned the way
    the last step went -- rejected attempts at dt, dt/4 with the RHS of the first accepted sub-step, then 4 or 16
    sub-steps without a clip in between -- and committed only if the reference would have taken exactly those
    decisions.  State, bed fluxes, diagnostics and the attempt / sub-cycle counters must be those of the
    single-attempt path in every regime, including the ones where the prediction keeps failing."""
    case = make_case("fuser", 12, 8, 15, 0.004, seed=2)
    on = _run(case, mode, 2, nsteps, calls=calls, **kw)
    off = _run(case, False, 2, nsteps, calls=calls, **kw)
    assert off["sub"] > 0
    _same(on, off)
    if kw.get("rnit") == 600. and nsteps * calls >= 27:
        # the steady regime: once a call has been planned from a sub-cycled step everything is fused
        assert on["fused"] >= nsteps * (calls - 2)
        assert on["launches"] < off["launches"]
    if kw.get("rnit") == 4000. and nsteps == 3:
        assert on["fused"] >= 9                      # steps 3..14 run at dt/16, planned from step 2


@pytest.mark.parametrize("mode", MODES)
def test_subcycling_fused_on_a_k40_masked_tile(gpu, mode):
    """knum = 40 on a tile with land in the steady dt/4 regime: the pair path over the wet-column list, and chains with
    two layers per lane."""
    case = make_case("fusek", 23, 11, 40, 0.0015, seed=12, land_fraction=0.3)
    kw = dict(rnit=600., rODUox=600.)
    on = _run(case, mode, 2, 10, calls=4, **kw)
    off = _run(case, False, 2, 10, calls=4, **kw)
    assert off["sub"] >= 30
    _same(on, off)
    assert on["fused"] >= 20 and on["launches"] < off["launches"]


@pytest.mark.parametrize("mode", MODES)
def test_fusion_accepts_violations_below_dt_min(gpu, mode):
    """dt <= dt_min: a violating step is accepted (solver_library.F90:126), so fused launches commit too."""
    case = make_case("fusem", 12, 8, 15, 0.004, seed=2)
    kw = dict(rnit=2.0e3, rODUox=2.0e3, dt_min=400.0)
    on = _run(case, mode, 2, 6, **kw)
    off = _run(case, False, 2, 6, **kw)
    assert off["sub"] == 0
    _same(on, off)
    assert on["launches"] < off["launches"]


@pytest.mark.parametrize("method", [0, 2])
def test_pair_lazy_clip_redoes_columns_that_need_the_clip(gpu, method):
    """pair_kernel walks a column without the clip and again with it if some new value had its sign bit set (zero
    minima): with reaction rates that overshoot, some columns do need the clip -- cells end exactly at the
    minimum -- and the result must still be the bits of the single steps, which clip every value."""
    case = make_case("lazy", 40, 24, 20, 0.003, seed=77, land_fraction=0.15)
    kw = dict(rnit=2.0e3, rODUox=2.0e3, dt_min=400.0)      # dt <= dt_min: every attempt is accepted as it is
    on = _run(case, "pairs", method, 6, **kw)
    off = _run(case, False, method, 6, **kw)
    _same(on, off)
    wet = case.mask == 0
    clipped = (on["conc"][wet] == 0.0).sum()
    assert 0 < clipped < on["conc"][wet].size // 2          # the clip fired somewhere, not everywhere
    assert on["launches"] < off["launches"]


@pytest.mark.parametrize("mode", MODES)
def test_fusion_nan_stops_at_the_same_step(gpu, mode):
    case = make_case("fusen", 6, 5, 12, 0.004, seed=4)

    def poison(sed):
        c = sed.conc
        c[3, 2, 5, 6] = np.nan
        sed.conc = c

    on = _run(case, mode, 2, 9, mutate=poison)
    off = _run(case, False, 2, 9, mutate=poison)
    assert on["rc"] == off["rc"] == 1 and on["done"] == off["done"] == 1


def test_fusion_stream_porosity_uses_single_steps(gpu):
    """An arbitrary 3-D porosity field (restart) is outside the fused kernel's scope: same results, no pairs."""
    case = make_case("fusep", 9, 6, 12, 0.004, seed=8)

    def mutate(sed):
        por = sed.field("porosity") * (1 + 0.05 * np.random.default_rng(1).random((9, 6, 12)))
        sed.set_porosity(np.minimum(por, 0.95))

    on = _run(case, True, 2, 8, mutate=mutate)
    off = _run(case, False, 2, 8, mutate=mutate)
    _same(on, off)
    assert on["launches"] == off["launches"]


# ---- Runge-Kutta stages fused (msed_rkpair.cuh, msed_rkquad.cuh, msed_chain.cuh) ------------------------
RK_MODES = ["pairs", "quad", "chains"]


@pytest.mark.parametrize("method", [1, 3])
@pytest.mark.parametrize("knum", [2, 3, 4, 5, 6, 7, 20, 40])
@pytest.mark.parametrize("mode", RK_MODES)
def test_rk_stage_fusion_is_bit_identical(gpu, oracle, method, knum, mode):
    """Runge-Kutta calls fused two stages per launch (mode pairs: rk_pair_kernel, thread per column), four stages per
    launch (mode quad: rk_quad_kernel, thread per column, one pass over the state per call; knum < 5 falls back to
    pairs) or four stages and several calls per launch with the column in registers (mode chains: rk_chain_kernel,
    warp per column)."""
    case = make_case("rkf", 37, 21, knum, 0.003, seed=55 + knum, land_fraction=0.2, smooth_temperature=True)
    on = _run(case, mode, method, 5, calls=2)
    off = _run(case, False, method, 5, calls=2)
    _same(on, off)
    assert on["launches"] < off["launches"]                 # two launches per step, or one per four steps, instead of four
    if mode == "chains" or (mode == "quad" and knum >= 5):
        assert on["launches"] < _run(case, "pairs", method, 5, calls=2)["launches"]
    ref = oracle.OracleSediment(37, 21, knum, 0.003, mask2d=case.mask, dt_min=1.0)
    ref.init_concentrations(); ref.set_boundary(case.bdys, case.fluxes)
    assert ref.step(DT, method, 10) == 0
    wet = case.mask == 0
    assert rel_err(on["conc"][wet], ref.conc[wet]) <= 1e-11


@pytest.mark.parametrize("method", [1, 3])
@pytest.mark.parametrize("kw", [dict(bcup_dissolved_variables=1), dict(bioturbation_profile=2),
                                dict(bcup_dissolved_variables=0),
                                dict(minimum=[1., 2., 3., 0.5, 30., 1., 2., 150.]), dict(model=1)])
@pytest.mark.parametrize("mode", RK_MODES)
def test_rk_stage_fusion_variants(gpu, method, kw, mode):
    case = make_case("rkfv", 19, 9, 15, 0.004, seed=9)

    def mutate(sed):
        if kw.get("bcup_dissolved_variables") == 1:
            fl = case.fluxes.copy(); fl[:, :, 3:] = 1e-6 * (1 + np.arange(5))
            sed.set_boundary(None, fl)
        sed.update_porosity(0.5 + 0.3 * np.random.default_rng(2).random((19, 9)))   # porosity mode 2

    _same(_run(case, mode, method, 6, mutate=mutate, **kw), _run(case, False, method, 6, mutate=mutate, **kw))


@pytest.mark.parametrize("method", [1, 3])
@pytest.mark.parametrize("mode", RK_MODES)
def test_rk_stage_fusion_nan_stops_at_the_same_step(gpu, method, mode):
    case = make_case("rkfn", 12, 8, 15, 0.004, seed=3)
    kw = dict(rnit=5.0e5, rODUox=5.0e5)                     # explicit RK blows up within a few steps
    on = _run(case, mode, method, 40, **kw)
    off = _run(case, False, method, 40, **kw)
    assert on["rc"] == off["rc"] and on["done"] == off["done"]
    assert np.array_equal(on["conc"], off["conc"], equal_nan=True)


@pytest.mark.parametrize("method", [1, 3])
def test_rk_unfusable_configs_take_the_staged_path(gpu, method):
    case = make_case("rkfs", 12, 8, 15, 0.004, seed=4)
    for kw in (dict(bioturbation_profile=3), dict(distributed_pom_flux=1)):
        on = _run(case, True, method, 4, **kw)
        off = _run(case, False, method, 4, **kw)
        _same(on, off)
        assert on["launches"] == off["launches"]
