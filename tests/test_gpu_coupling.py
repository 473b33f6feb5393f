"""BASELINE config 5 (benthic-pelagic coupling) at test size: one pelagic box per column exchanges
surface concentrations / sinking fluxes and bed fluxes with the sediment each coupling step, entirely on
the device (msed_coupled_run), against the same sequence on the oracle + numpy."""
import numpy as np
import pytest

from tests.cases import make_case, rel_err, scaled_err

pytestmark = pytest.mark.gpu
DT, COUPLING = 360.0, 3600.0


def _pelagic(case, rng):
    shape2 = case.mask.shape
    conc = np.empty(shape2 + (8,), order="F")
    base = np.array([30.0, 300.0, 1.0, 0.6, 14.0, 4.0, 250.0, 0.5])   # mmol m-3 in the bottom water
    for n in range(8):
        conc[:, :, n] = base[n] * (1.0 + 0.1 * rng.uniform(-1, 1, shape2))
    wz = np.zeros(shape2 + (8,), order="F")
    wz[:, :, :3] = -(1.0 + rng.random(shape2 + (3,))) * 1e-6          # sinking (downward = negative)
    height = 5.0 + 10.0 * rng.random(shape2)
    height[0, 0] = 0.0                                                 # where(layer_height > 0) branch
    temp = 4.0 + 8.0 * rng.random(shape2)
    return conc, wz, np.asfortranarray(height), np.asfortranarray(temp)


@pytest.mark.parametrize("bcup", [2, 1])
def test_coupled_run_matches_oracle_sequence(gpu, oracle, bcup):
    from mossco_code_b200 import SedimentDriver, default_config
    case = make_case("c5", 10, 6, 20, 0.003, seed=55, land_fraction=0.2)
    rng = np.random.default_rng(9)
    conc, wz, height, temp = _pelagic(case, rng)
    cfg = default_config(inum=10, jnum=6, knum=20, dzmin=0.003, dt_min=1.0, bcup_dissolved_variables=bcup)
    ncoup = 4
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.pelagic_init(conc, wz, height, temp)
        assert sed.coupled_run(DT, 2, COUPLING, ncoup) == 0
        assert sed.info.steps_done == 10 * ncoup
        got_pel, got_sed, got_flux = sed.pelagic_conc, sed.conc, sed.fluxes
    ref = oracle.OracleSediment.from_config(cfg, mask2d=case.mask)
    ref.init_concentrations()
    pel = conc.copy(order="F")
    wet = case.mask == 0
    for _ in range(ncoup):
        cs = [pel[:, :, n] for n in range(8)]
        ws = [wz[:, :, n] if n < 3 else None for n in range(8)]
        ref.get_boundary_conditions(temp, cs, ws)
        assert ref.step(DT, 2, 10) == 0
        up = -ref.fluxes
        ok = wet & (height > 0)
        for n in range(8):   # conc + bfl*dt/layer_height, fabm_pelagic_component.F90:2100-2105
            pel[:, :, n][ok] = pel[:, :, n][ok] + up[:, :, n][ok] * COUPLING / height[ok]
    assert rel_err(got_sed[wet], ref.conc[wet]) <= 1e-10
    assert scaled_err(got_pel[wet], pel[wet]) <= 1e-12
    assert np.array_equal(got_pel[0, 0], conc[0, 0])          # zero-height box untouched
    assert np.array_equal(got_pel[~wet], conc[~wet])          # masked columns untouched
    assert scaled_err(got_flux[wet], ref.fluxes[wet]) <= 1e-10
    # exchange actually happened: oxygen is drawn down, particulates are lost to the bed
    assert np.all(got_pel[wet & (height > 0)][:, 6] < conc[wet & (height > 0)][:, 6])
    assert np.all(got_pel[wet & (height > 0)][:, 0] < conc[wet & (height > 0)][:, 0])


@pytest.mark.parametrize("full", [True, False])
def test_pelagic_soil_couplers(gpu, oracle, full):
    """pelagic_benthic_coupler fused with get_boundary_conditions, and benthic_pelagic_coupler
    (src/mediators/pelagic_benthic_coupler.F90:281-492, benthic_pelagic_coupler.F90:188-287)."""
    from mossco_code_b200 import SedimentDriver, default_config
    case = make_case("cpl", 9, 7, 15, 0.004, seed=66)
    rng = np.random.default_rng(3)
    sh = (9, 7)
    f = dict(temperature=4 + 8 * rng.random(sh), oxygen=250 * (rng.random(sh) - 0.2),
             detN=2 + rng.random(sh), detN_z_velocity=-1e-5 * (1 + rng.random(sh)), DIN=10 + 5 * rng.random(sh))
    if full:
        f.update(detC=(6 + 2 * rng.random(sh)) * f["detN"], detP=0.1 + 0.1 * rng.random(sh),
                 detP_z_velocity=-2e-5 * (1 + rng.random(sh)), nitrate=8 + rng.random(sh),
                 ammonium=3 + rng.random(sh), DIP=0.5 + rng.random(sh))
    cfg = default_config(inum=9, jnum=7, knum=15, dzmin=0.004, dt_min=1.0)
    with SedimentDriver(cfg) as sed:
        sed.init_concentrations()
        sed.pelagic_benthic_coupler(**f)
        ref = oracle.OracleSediment.from_config(cfg)
        ref.init_concentrations()
        cs, wz = oracle.pelagic_benthic_coupler(sh, **{k: v for k, v in f.items() if k != "temperature"})
        ref.get_boundary_conditions(f["temperature"], cs, wz)
        assert np.array_equal(sed.bdys, ref.bdys)            # same IEEE operation sequence
        assert np.array_equal(sed.fluxes, ref.fluxes)
        assert np.all(sed.bdys[:, :, 7] >= 0) and np.all(sed.bdys[:, :, 8] >= 0)
        assert np.any(sed.bdys[:, :, 8] > 0)                  # negative oxygen became reduced substances
        assert sed.run(360.0, 2, 3600.0) == 0 and ref.step(360.0, 2, 10) == 0
        got = sed.benthic_pelagic_coupler(want=oracle.B2P_FIELDS, dinflux_const=0.3, convertN=1.0 if full else 2.0)
        want = oracle.benthic_pelagic_coupler(-sed.fluxes, dinflux_const=0.3, convertN=1.0 if full else 2.0)
        for k in oracle.B2P_FIELDS:
            assert np.array_equal(got[k], want[k]), k
        assert scaled_err(sed.fluxes, ref.fluxes) <= 1e-10


@pytest.mark.parametrize("want", [("nitrate", "ammonium", "DIN", "DIP", "oxygen", "odu", "detN", "detC", "detP"),
                                  ("DIN", "DIP", "odu", "detC"), ("nitrate", "ammonium", "DIP", "oxygen", "detP")])
def test_soil_pelagic_connector(gpu, oracle, want):
    """soil_pelagic_connector Run (src/mediators/soil_pelagic_connector.F90:179-981) on the device, incl.
    the three oxygen/odu branches (:660-720) and the namelist factors (:140)."""
    from mossco_code_b200 import SedimentDriver, default_config
    case = make_case("s2p", 11, 5, 15, 0.004, seed=67, land_fraction=0.2)
    cfg = default_config(inum=11, jnum=5, knum=15, dzmin=0.004, dt_min=1.0)
    kw = dict(dinflux_const=0.3, dipflux_const=-1.0 if "odu" in want else 0.01, convertN=1.5, convertP=0.75)
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        assert sed.step(360.0, 2, 3) == 0
        got = sed.soil_pelagic_connector(want=want, **kw)
        ref = oracle.soil_pelagic_connector(-sed.fluxes, want=want, **kw)
        assert set(got) == set(want) == set(ref)
        for k in want:
            assert np.array_equal(got[k], ref[k]), k      # same IEEE operation sequence
        up = -sed.fluxes
        if "odu" in want and "oxygen" not in want:
            assert np.array_equal(got["odu"], up[:, :, 7] - up[:, :, 6])
        if "detC" in want:
            assert np.array_equal(got["detC"], up[:, :, 0] + up[:, :, 1])
        for k in ("detN", "detP"):
            if k in want:
                assert not got[k].any()


def test_mediator_mirrors_in_a_coupled_sequence(gpu, oracle):
    """pelagic model -> PelagicBenthicCoupler -> sediment Run -> SoilPelagicConnector / BenthicPelagicCoupler ->
    pelagic model, with the field names of an ECOSMO-like and a MAECS-like pelagic model
    (mossco_code_b200/mediators.py), against the oracle's restatement of the same mediators."""
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.mediators import BenthicPelagicCoupler, PelagicBenthicCoupler, SoilPelagicConnector
    from mossco_code_b200.sediment import VARIABLE_NAMES
    sh = (9, 7)
    rng = np.random.default_rng(8)
    f3 = lambda lo, hi: np.asfortranarray(lo + (hi - lo) * rng.random(sh + (4,)))      # (i, j, layer): bottom = layer 0
    pel = {"temperature_in_water": f3(4, 12), "oxygen_in_water": f3(150, 300),
           "Detritus_Nitrogen_detN_in_water": f3(2, 3), "Detritus_Nitrogen_detN_z_velocity_in_water": -1e-5 * f3(1, 2),
           "Detritus_Carbon_detC_in_water": f3(14, 20), "Dissolved_Inorganic_Nitrogen_DIN_nutN_in_water": f3(10, 15)}
    cfg = default_config(inum=9, jnum=7, knum=15, dzmin=0.004, dt_min=1.0)
    with SedimentDriver(cfg) as sed:
        sed.init_concentrations()
        exp = {}
        assert PelagicBenthicCoupler(sed).run(pel, exp, fill_export=True) == 0
        ref = oracle.OracleSediment.from_config(cfg)
        ref.init_concentrations()
        b = lambda k: pel[k][:, :, 0]
        cs, wz = oracle.pelagic_benthic_coupler(sh, oxygen=b("oxygen_in_water"), detN=b("Detritus_Nitrogen_detN_in_water"),
                                                detN_z_velocity=b("Detritus_Nitrogen_detN_z_velocity_in_water"),
                                                detC=b("Detritus_Carbon_detC_in_water"),
                                                DIN=b("Dissolved_Inorganic_Nitrogen_DIN_nutN_in_water"))
        ref.get_boundary_conditions(b("temperature_in_water"), cs, wz)
        assert np.array_equal(sed.bdys, ref.bdys) and np.array_equal(sed.fluxes, ref.fluxes)
        assert np.array_equal(exp["dissolved_oxygen_at_soil_surface"], ref.bdys[:, :, 7])
        assert sed.run(360.0, 2, 3600.0) == 0
        soil = {f"{v}_upward_flux_at_soil_surface": None for v in VARIABLE_NAMES}     # presence is what is checked
        up = -sed.fluxes
        # ECOSMO-like names: separate nitrate/ammonium, phosphate, oxygen only (-> oxygen minus reduced substances)
        eco = {n: np.zeros(sh, order="F") for n in (
            "hzg_ecosmo_no3_upward_flux_at_soil_surface", "hzg_ecosmo_nh4_upward_flux_at_soil_surface",
            "hzg_ecosmo_pho_upward_flux_at_soil_surface", "hzg_ecosmo_oxy_upward_flux_at_soil_surface")}
        kw = dict(dinflux_const=0.3, convertN=1.5, convertP=0.75)
        assert SoilPelagicConnector(sed, **kw).run(soil, eco) == 0
        want = oracle.soil_pelagic_connector(up, want=("nitrate", "ammonium", "DIP", "oxygen"), **kw)
        assert np.array_equal(eco["hzg_ecosmo_no3_upward_flux_at_soil_surface"], want["nitrate"])
        assert np.array_equal(eco["hzg_ecosmo_nh4_upward_flux_at_soil_surface"], want["ammonium"])
        assert np.array_equal(eco["hzg_ecosmo_pho_upward_flux_at_soil_surface"], want["DIP"])
        assert np.array_equal(eco["hzg_ecosmo_oxy_upward_flux_at_soil_surface"], up[:, :, 6] - up[:, :, 7])
        # MAECS-like names through the older coupler: DIN branch, detritus N from the carbon fluxes
        maecs = {n: np.zeros(sh, order="F") for n in (
            "Dissolved_Inorganic_Nitrogen_DIN_nutN_upward_flux_at_soil_surface",
            "Dissolved_Inorganic_Phosphorus_DIP_nutP_upward_flux_at_soil_surface",
            "Detritus_Nitrogen_detN_upward_flux_at_soil_surface", "Detritus_Carbon_detC_upward_flux_at_soil_surface",
            "oxygen_upward_flux_at_soil_surface")}
        assert BenthicPelagicCoupler(sed, dinflux_const=0.3).run(soil, maecs) == 0
        want = oracle.benthic_pelagic_coupler(up, dinflux_const=0.3)
        for name, key in (("Dissolved_Inorganic_Nitrogen_DIN_nutN", "DIN"), ("Dissolved_Inorganic_Phosphorus_DIP_nutP", "DIP"),
                          ("Detritus_Nitrogen_detN", "detN"), ("Detritus_Carbon_detC", "detC"), ("oxygen", "oxygen")):
            assert np.array_equal(maecs[f"{name}_upward_flux_at_soil_surface"], want[key]), key


@pytest.mark.parametrize("nchunks,seconds", [(4, 3600.0), (3, 1000.0), (5, 360.0), (1, 3600.0), (4, 720.0), (0, 3600.0), (0, 1000.0), (0, 360.0)])
@pytest.mark.parametrize("mode", ["auto", "pairs"])   # auto: chains on this tile; pairs: the wet-column list, chunked
def test_run_exchange_equals_separate_calls(gpu, nchunks, seconds, mode):
    """msed_run_exchange (chunk-pipelined PCIe/compute overlap) must be bit-identical to
    get_boundary_conditions + run + upward_fluxes, incl. the shortened last step and a single step."""
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    case = make_case("xchg", 70, 37, 20, 0.003, seed=91, land_fraction=0.25)
    rng = np.random.default_rng(5)
    sh = (70, 37)
    temp = 4 + 8 * rng.random(sh)
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if PARTICULATE[n] else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.ones(sh, order="F") if PARTICULATE[n] else None for n in range(8)]
    cs[5] = None                                              # ammonium absent from the import state
    res = []
    for fused in (False, True):
        cfg = default_config(inum=70, jnum=37, knum=20, dzmin=0.003, dt_min=1.0)
        with SedimentDriver(cfg) as sed:
            sed.set_mask(case.mask)
            sed.init_concentrations()
            sed.set_boundary(case.bdys, case.fluxes)
            sed.set_step_fusion(mode if fused else "off")
            for _ in range(2):
                if fused:
                    sed.set_exchange_chunks(nchunks)
                    rc, up = sed.run_exchange(360.0, 2, seconds, temp, cs, wz)
                else:
                    sed.get_boundary_conditions(temp, cs, wz)
                    rc = sed.run(360.0, 2, seconds)
                    up = sed.upward_fluxes()
                assert rc == 0
                steps = sed.info.steps_done
            res.append((sed.conc, up.copy(), sed.bdys, steps, sed.field("denit")))
    assert res[0][3] == res[1][3]
    assert np.array_equal(res[0][4], res[1][4])           # diagnostic of the last get_rhs state
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize("nchunks,seconds,boost", [(4, 3600.0, False), (3, 2880.0, False), (5, 1440.0, False),
                                                   (4, 3600.0, True), (2, 2880.0, True)])
def test_run_exchange_chunk_major(gpu, nchunks, seconds, boost):
    """Chunk-major order of a coupling interval (msed_set_exchange_order): every chunk runs all pairs of the
    interval through explicitly named buffers and one controller commits them -- odd and even numbers of pairs
    (the even case ends in the staging buffer: pointer rotation), a tile with land (wet-column list), several
    Runs in a row, a Run with rejected attempts (nothing committed, redone step by step), and ordinary calls
    afterwards must all give the bits of the separate unfused calls."""
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    case = make_case("xcm", 70, 37, 20, 0.003, seed=93, land_fraction=0.25)
    rng = np.random.default_rng(6)
    sh = (70, 37)
    temp = 4 + 8 * rng.random(sh)
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if PARTICULATE[n] else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.ones(sh, order="F") if PARTICULATE[n] else None for n in range(8)]
    kw = dict(rnit=2.0e3, rODUox=2.0e3) if boost else {}
    res = []
    for fused in (False, True):
        cfg = default_config(inum=70, jnum=37, knum=20, dzmin=0.003, dt_min=1.0, **kw)
        with SedimentDriver(cfg) as sed:
            sed.set_mask(case.mask)
            sed.init_concentrations()
            sed.set_boundary(case.bdys, case.fluxes)
            sed.set_step_fusion("pairs" if fused else "off")
            ups, sub, fusedsteps, concs, fields = [], 0, 0, [], []
            for _ in range(3):
                if fused:
                    sed.set_exchange_chunks(nchunks)
                    sed.set_exchange_order(True)
                    rc, up = sed.run_exchange(360.0, 2, seconds, temp, cs, wz)
                else:
                    sed.get_boundary_conditions(temp, cs, wz)
                    rc = sed.run(360.0, 2, seconds)
                    up = sed.upward_fluxes()
                assert rc == 0
                ups.append(up.copy())
                sub += sed.info.subcycle_warnings
                fusedsteps += sed.info.fused_steps
                # the state after EVERY Run (an even number of pairs leaves it in the rotated-in staging buffer,
                # whose land columns no stepping kernel writes): land stays missing_value, driver :464
                concs.append(sed.conc)
                assert np.all(concs[-1][case.mask > 0] == 1e20)
                # calls that use the staging buffer between Runs: derived fields and one RHS evaluation
                fields.append((sed.field("photosynthetically_active_radiation"), sed.get_rhs()))
            denit = sed.field("denit")
            rhs = sed.get_rhs()                       # uses the staging buffer: must still be a valid one
            assert sed.step(360.0, 2, 3) == 0         # and ordinary calls carry on from the rotated buffers
            assert np.all(sed.conc[case.mask > 0] == 1e20)
            res.append((sed.conc, ups, sed.bdys, denit, rhs, sub, fusedsteps, concs, fields))
    a, b = res
    for x, y in zip(a[7], b[7]):
        assert np.array_equal(x, y)
    for (fx, rx), (fy, ry) in zip(a[8], b[8]):
        assert np.array_equal(fx, fy) and np.array_equal(rx, ry)
    assert a[5] == b[5] and (a[5] > 0) == boost
    if not boost and b[5] == 0:
        assert b[6] == 3 * int(round(seconds / 360.0))          # every step of every Run went through the sequence
    assert np.array_equal(a[0], b[0])
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


@pytest.mark.parametrize("nchunks", [0, 4])
def test_run_exchange_static_import_fields(gpu, nchunks):
    """msed_set_import_generations: a field whose generation counter did not move is not uploaded again -- the
    Run uses the copy the device kept (so host edits without a bump are invisible), a bumped counter or a new
    host array uploads it, and with unchanged data the results are bit-identical to uploading everything."""
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    case = make_case("xgen", 70, 37, 20, 0.003, seed=94, land_fraction=0.2)
    rng = np.random.default_rng(8)
    sh = (70, 37)
    temp = np.asfortranarray(4 + 8 * rng.random(sh))
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if PARTICULATE[n] else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.asfortranarray(1.0 + rng.random(sh)) if PARTICULATE[n] else None for n in range(8)]
    cfg = default_config(inum=70, jnum=37, knum=20, dzmin=0.003, dt_min=1.0)

    def runs(static, edit=None):
        out = []
        with SedimentDriver(cfg) as sed:
            sed.set_mask(case.mask)
            sed.init_concentrations()
            sed.set_step_fusion("pairs")
            sed.set_exchange_chunks(nchunks)
            w = [None if a is None else a.copy(order="F") for a in wz]
            gen = [1] * 17
            for r in range(3):
                if static:
                    for k in range(17):            # everything but the z-velocities changes every Run
                        if not (k >= 2 and k % 2 == 0):
                            gen[k] += 1
                    if edit == "bump" and r == 2:
                        gen[2] += 1
                    sed.set_import_generations(gen)
                if edit and r == 2:
                    w[0][...] *= 3.0               # the coupler rewrites detritus' sinking velocity in place
                rc, up = sed.run_exchange(360.0, 2, 3600.0, temp, cs, w)
                assert rc == 0
                out.append((up.copy(), sed.conc))
        return out

    plain, static = runs(False), runs(True)
    for (u0, c0), (u1, c1) in zip(plain, static):
        assert np.array_equal(u0, u1) and np.array_equal(c0, c1)
    silent, bumped, plain_edit = runs(True, "silent"), runs(True, "bump"), runs(False, "plain")
    assert np.array_equal(silent[2][1], static[2][1])            # not announced: the device copy is used
    assert np.array_equal(bumped[2][1], plain_edit[2][1])        # announced: uploaded
    assert not np.array_equal(bumped[2][1], static[2][1])


def test_run_exchange_zero_copy_with_pinned_fields(gpu):
    """A one-chunk tile whose import fields and flux buffer are pinned host memory takes the zero-copy path of
    msed_run_exchange (the boundary kernel reads the fields, the export kernel writes the fluxes through PCIe):
    same bits as the staged copies pageable arrays get, Run after Run, also when the coupler rewrites a field in
    place between Runs."""
    import torch
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    case = make_case("xzc", 50, 31, 20, 0.003, seed=95, land_fraction=0.2)
    rng = np.random.default_rng(10)
    sh = (50, 31)
    keep = []

    def pinned(a):
        t = torch.empty(tuple(reversed(a.shape)), dtype=torch.float64).pin_memory()
        keep.append(t)
        v = t.numpy().T
        v[...] = a
        return v

    temp = np.asfortranarray(4 + 8 * rng.random(sh))
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if PARTICULATE[n] else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.asfortranarray(1.0 + rng.random(sh)) if PARTICULATE[n] else None for n in range(8)]
    cfg = default_config(inum=50, jnum=31, knum=20, dzmin=0.003, dt_min=1.0)

    def runs(pin):
        conv = pinned if pin else (lambda a: a.copy(order="F"))
        t, c, w = conv(temp), [conv(a) for a in cs], [None if a is None else conv(a) for a in wz]
        out = conv(np.zeros(sh + (8,), order="F"))
        res = []
        with SedimentDriver(cfg) as sed:
            sed.set_mask(case.mask)
            sed.init_concentrations()
            for r in range(3):
                if r == 2:
                    c[6][...] *= 0.5            # oxygen above the bed halves
                rc, up = sed.run_exchange(360.0, 2, 3600.0, t, c, w, out=out)
                assert rc == 0 and up is out
                res.append((up.copy(), sed.conc))
        return res

    a, b = runs(False), runs(True)
    for (u0, c0), (u1, c1) in zip(a, b):
        assert np.array_equal(u0, u1) and np.array_equal(c0, c1)
    assert not np.array_equal(a[1][0], a[2][0])


def test_run_exchange_with_rejected_attempt(gpu):
    """If an attempt is rejected the chunk-wise export is stale and must be redone from the final state."""
    from mossco_code_b200 import SedimentDriver, default_config
    case = make_case("xrej", 40, 16, 15, 0.004, seed=2)
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if n < 3 else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.ones((40, 16), order="F") if n < 3 else None for n in range(8)]
    temp = np.asfortranarray(case.bdys[:, :, 0])
    out = []
    for fused in (False, True):
        cfg = default_config(inum=40, jnum=16, knum=15, dzmin=0.004, dt_min=1.0, rnit=2.0e3, rODUox=2.0e3)
        with SedimentDriver(cfg) as sed:
            sed.init_concentrations()
            if fused:
                sed.set_exchange_chunks(4)
                rc, up = sed.run_exchange(360.0, 2, 1800.0, temp, cs, wz)
            else:
                sed.get_boundary_conditions(temp, cs, wz)
                rc = sed.run(360.0, 2, 1800.0)
                up = sed.upward_fluxes()
            assert rc == 0 and sed.info.subcycle_warnings > 0
            out.append((sed.conc, up.copy(), sed.info.subcycle_warnings))
    assert out[0][2] == out[1][2]
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_check_domain_flags_bad_porosity(gpu):
    """fabm_sed_check_domain 'stop' conditions (driver :503-511) come back as MSED_BAD_DOMAIN."""
    from mossco_code_b200 import MsedError, SedimentDriver, default_config
    cfg = default_config(inum=6, jnum=5, knum=12, dzmin=0.004)
    mask = np.zeros((6, 5), dtype=np.int32); mask[1, 1] = 1
    for bad, text in ((1.2, "> 1"), (0.0, "<=0")):
        with SedimentDriver(cfg) as sed:
            sed.set_mask(mask)
            assert sed.check_domain() == 0
            por = sed.field("porosity")
            por[1, 1, 3] = bad                       # masked column: ignored (and reset to 1)
            sed.set_porosity(por)
            assert sed.check_domain() == 0
            por[4, 2, 7] = bad
            sed.set_porosity(por)
            with pytest.raises(MsedError) as e:
                sed.check_domain()
            assert e.value.code == 2 and text in str(e.value)


P2S_CASES = [
    # which pelagic fields exist -> which of the connector's branches run
    dict(detC=True, detP=True, detP_z_velocity=True, nitrate=True, ammonium=True, DIN=True, DIP=True, oxygen=True,
         odu=True, water_depth=True, tke=True),                                  # every field
    dict(DIN=True, oxygen=True),                                                 # NPZD-like: DIN only, oxygen only
    dict(nitrate=True, DIN=True, odu=True, detC=True, water_depth=True),         # ammonium = DIN - nitrate; odu only
    dict(ammonium=True, detP=True, tke=True),                                    # nitrate = ammonium; DIN = 2*ammonium
    dict(nitrate=True, ammonium=True, DIP=True),                                 # neither oxygen nor odu: rows untouched
]


@pytest.mark.parametrize("head", [False, True])
@pytest.mark.parametrize("present", P2S_CASES)
def test_pelagic_soil_connector(gpu, oracle, present, head):
    """pelagic_soil_connector (src/mediators/pelagic_soil_connector.F90:176-2122) fused with
    get_boundary_conditions against the restatement, for the default (intended) algebra and for what the HEAD
    revision computes (msed_set_compat), with non-default namelist values."""
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    case = make_case("p2s", 9, 7, 15, 0.004, seed=67)
    rng = np.random.default_rng(4)
    sh = (9, 7)
    allf = dict(temperature=4 + 8 * rng.random(sh), par=30 * rng.random(sh), oxygen=250 * (rng.random(sh) - 0.2),
                odu=40 * (rng.random(sh) - 0.5), detN=2 + rng.random(sh), detN_z_velocity=-1e-5 * (1 + rng.random(sh)),
                detC=(2 + 60 * rng.random(sh)), detP=0.1 + 0.1 * rng.random(sh),
                detP_z_velocity=-2e-5 * (1 + rng.random(sh)), nitrate=8 + rng.random(sh), ammonium=3 + rng.random(sh),
                DIN=12 + 5 * rng.random(sh), DIP=0.5 + rng.random(sh), water_depth=0.05 + 30 * rng.random(sh) ** 3,
                tke=2.0e3 * rng.random(sh))
    f = {k: v for k, v in allf.items() if k in ("temperature", "par", "detN", "detN_z_velocity") or present.get(k)}
    params = dict(sinking_factor=0.25, NC_ldet=0.21, convertN=1.5, convertP=0.8, critical_detritus=45.0)
    cfg = default_config(inum=9, jnum=7, knum=15, dzmin=0.004, dt_min=1.0)
    with SedimentDriver(cfg) as sed:
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        ref = oracle.OracleSediment.from_config(cfg)
        ref.init_concentrations()
        ref.set_boundary(case.bdys, case.fluxes)
        sed.set_compat(p2s_head=head)
        sed.pelagic_soil_connector(params=params, **f)
        cs, wz = oracle.pelagic_soil_connector(sh, params=params, head_compat=head,
                                               **{k: v for k, v in f.items() if k not in ("temperature", "par")})
        if not (present.get("oxygen") or present.get("odu")):
            cs[6] = cs[7] = None                      # not in the transfer: the sediment keeps its boundary values
        ref.get_boundary_conditions(f["temperature"], cs, wz)
        assert np.array_equal(sed.bdys, ref.bdys)            # same IEEE operation sequence
        assert np.array_equal(sed.fluxes, ref.fluxes)
        assert np.array_equal(sed.field("photosynthetically_active_radiation")[:, :, 0], f["par"])
        if not head:   # the split conserves detritus carbon: fac_ldet + fac_sdet = C:N
            cn = f["detC"] / (np.float64(np.float32(1e-5)) + f["detN"]) if "detC" in f else 106.0 / 16.0
            total = (cs[0] + cs[1]) / (params["convertN"] * f["detN"])
            assert np.allclose(total, cn, rtol=1e-14)
        else:
            assert np.all(sed.fluxes[:, :, :2] == 0.0)       # HEAD never writes the carbon velocities
        assert sed.step(360.0, 2, 2) == 0 and ref.step(360.0, 2, 2) == 0
        assert scaled_err(sed.conc, ref.conc) <= 1e-11
    assert sum(PARTICULATE) == 3


def test_pelagic_benthic_coupler_oxygen_quirk_switch(gpu, oracle):
    """MSED_COMPAT_P2B_OXYGEN_LAST_CELL: the whole-array assignment of pelagic_benthic_coupler.F90:344-349 as
    written (every column gets the last cell's oxygen split) against the per-column default."""
    from mossco_code_b200 import SedimentDriver, default_config
    rng = np.random.default_rng(8)
    sh = (6, 5)
    f = dict(temperature=4 + 8 * rng.random(sh), oxygen=250 * (rng.random(sh) - 0.3), detN=2 + rng.random(sh),
             detN_z_velocity=-1e-5 * (1 + rng.random(sh)), DIN=10 + 5 * rng.random(sh))
    cfg = default_config(inum=6, jnum=5, knum=12, dzmin=0.004, dt_min=1.0)
    for quirk in (False, True):
        with SedimentDriver(cfg) as sed:
            sed.init_concentrations()
            sed.set_compat(p2b_oxygen_last_cell=quirk)
            sed.pelagic_benthic_coupler(**f)
            cs, wz = oracle.pelagic_benthic_coupler(sh, oxy_last_cell=quirk, **{k: v for k, v in f.items() if k != "temperature"})
            assert np.array_equal(sed.bdys[:, :, 7], cs[6]) and np.array_equal(sed.bdys[:, :, 8], cs[7])
            if quirk:
                last = f["oxygen"][-1, -1]
                assert np.all(sed.bdys[:, :, 7] == max(0.0, last)) and np.all(sed.bdys[:, :, 8] == max(0.0, -last))
            else:
                assert len(np.unique(sed.bdys[:, :, 7])) > 2


def test_diagnostics_and_checksum(gpu, oracle):
    """msed_diagnostics (bed-flux sums, inventories over the wet columns) against numpy on the downloaded
    state, and msed_state_checksum: two half tiles reproduce the whole tile's pair."""
    from mossco_code_b200 import SedimentDriver, default_config
    case = make_case("diag", 12, 10, 15, 0.004, seed=71, land_fraction=0.25)

    def tile(j0, j1):
        cfg = default_config(inum=12, jnum=j1 - j0, knum=15, dzmin=0.004, dt_min=1.0, j_offset=j0)
        sed = SedimentDriver(cfg)
        sed.set_mask(np.asfortranarray(case.mask[:, j0:j1]))
        sed.init_concentrations()
        sed.set_boundary(np.asfortranarray(case.bdys[:, j0:j1]), np.asfortranarray(case.fluxes[:, j0:j1]))
        assert sed.step(360.0, 0, 4) == 0          # Euler: no whole-domain decision, tiles are independent
        return sed

    whole = tile(0, 10)
    b, inv = whole.diagnostics()
    wet = case.mask == 0
    conc, por, fl = whole.conc, whole.field("porosity"), whole.fluxes
    _, _, dz, _ = whole.grid()
    for n in range(8):
        assert np.isclose(b[n], fl[:, :, n][wet].sum(), rtol=1e-12, atol=1e-300)
        assert np.isclose(inv[n], (conc[:, :, :, n] * por * dz)[wet].sum(), rtol=1e-12)
    want = whole.state_checksum(global_ncol=120, col_offset=0)
    parts = [tile(0, 4), tile(4, 10)]
    got = [p.state_checksum(global_ncol=120) for p in parts]
    assert ((got[0][0] + got[1][0]) % 2 ** 64, got[0][1] ^ got[1][1]) == want
    i, j = np.argwhere(wet)[7]
    c = conc.copy(); c[i, j, 5, 1] = np.nextafter(c[i, j, 5, 1], 1e30)
    whole.conc = c
    assert whole.state_checksum(global_ncol=120, col_offset=0) != want     # one ulp in one cell shows
    for s in [whole] + parts:
        s.finalize()
