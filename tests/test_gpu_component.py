"""The ESMF-component mirror (mossco_code_b200/component.py) against the same sequence driven on the
oracle: phases, import/export field names and one coupling interval of Run
(src/components/fabm_sediment_component.F90:99-131, :943-1192, :1493-1829)."""
import numpy as np
import pytest

from tests.cases import make_case, rel_err, scaled_err

pytestmark = pytest.mark.gpu


def _import_state(case, rng):
    from mossco_code_b200.sediment import PARTICULATE, VARIABLE_NAMES
    s = {"temperature_at_soil_surface": np.asfortranarray(case.bdys[:, :, 0].copy())}
    for n, v in enumerate(VARIABLE_NAMES):
        if PARTICULATE[n]:   # default.dat style: negative "concentration", w = 1 => downward flux
            s[f"{v}_at_soil_surface"] = np.asfortranarray(-case.fluxes[:, :, n])
            s[f"{v}_z_velocity_at_soil_surface"] = np.ones(case.mask.shape, order="F")
        else:
            s[f"{v}_at_soil_surface"] = np.asfortranarray(case.bdys[:, :, n + 1].copy())
    s["photosynthetically_active_radiation_at_soil_surface"] = np.asfortranarray(case.par_surface)
    return s


def test_component_run_matches_oracle_sequence(gpu, oracle):
    from mossco_code_b200.component import FabmSedimentComponent
    from mossco_code_b200.sediment import PARTICULATE, VARIABLE_NAMES
    case = make_case("comp", 9, 7, 15, 0.004, seed=17, land_fraction=0.2, par_max=40.0)
    comp = FabmSedimentComponent()
    imp, exp = {}, {}
    assert set(k for k in comp.set_services()) >= {("initialize", 0), ("initialize", 1), ("initialize", 2),
                                                    ("readrestart", 1), ("run", 1), ("finalize", 1)}
    comp.initialize_p0(imp, exp)
    assert comp.phase_map == ["IPDv00p1=1", "IPDv00p2=2"]
    gridmask = 1 - case.mask                      # ESMF_GRIDITEM_MASK: <=0 is masked (:499)
    comp.initialize_p1(imp, exp, grid_shape=(9, 7), grid_mask=gridmask,
                       run_nml=dict(numlayers=15, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2))
    comp.initialize_p2(imp, exp)
    # import / export catalogue
    assert "temperature_at_soil_surface" in imp and "porosity_at_soil_surface" in imp
    for n, v in enumerate(VARIABLE_NAMES):
        assert f"{v}_at_soil_surface" in imp
        assert (f"{v}_z_velocity_at_soil_surface" in imp) == PARTICULATE[n]
        assert exp[f"{v}_in_soil"].shape == (9, 7, 15)
        assert exp[f"{v}_upward_flux_at_soil_surface"].shape == (9, 7)
    for s in ("porosity", "layer_height", "layer_center_depth", "temperature",
              "photosynthetically_active_radiation"):
        assert exp[f"{s}_in_soil"].shape == (9, 7, 15)
    assert set(comp.export_field_names()) <= set(exp)

    # the same on the oracle
    ref = oracle.OracleSediment.from_config(comp.cfg, mask2d=case.mask)
    ref.init_concentrations()
    wet = case.mask == 0
    assert np.array_equal(exp["dissolved_oxygen_in_soil"][wet], ref.conc[:, :, :, 6][wet])

    rng = np.random.default_rng(0)
    state = _import_state(case, rng)
    imp.update(state)
    for it in range(3):                           # three coupling intervals of 1 h
        comp.run(imp, exp, clock={"currTime": 3600.0 * it, "stopTime": 3600.0 * (it + 1)})
        ref.par_surface[...] = case.par_surface
        cs = [state[f"{v}_at_soil_surface"] for v in VARIABLE_NAMES]
        wz = [state.get(f"{v}_z_velocity_at_soil_surface") for v in VARIABLE_NAMES]
        ref.get_boundary_conditions(state["temperature_at_soil_surface"], cs, wz)
        assert ref.step(360.0, 2, 10) == 0
        assert comp.last_info.steps_done == 10
    for n, v in enumerate(VARIABLE_NAMES):
        assert rel_err(exp[f"{v}_in_soil"][wet], ref.conc[:, :, :, n][wet]) <= 1e-10, v
        assert scaled_err(exp[f"{v}_upward_flux_at_soil_surface"][wet][..., None],
                          -ref.fluxes[:, :, n][wet][..., None]) <= 1e-10, v
    assert np.array_equal(exp["temperature_in_soil"][wet], ref.field3d("temp3d")[wet])
    assert scaled_err(exp["photosynthetically_active_radiation_in_soil"][wet][..., None],
                      ref.field3d("par")[wet][..., None]) < 1e-14
    assert np.array_equal(exp["porosity_in_soil"], ref.field3d("porosity"))
    comp.finalize()


def test_component_shortened_last_step(gpu, oracle):
    """Run stops exactly at stopTime: 1000 s = 2 x 360 s + 280 s (component :1705-1708)."""
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("short", 4, 3, 12, 0.004, seed=5)
    comp = FabmSedimentComponent()
    imp, exp = {}, {}
    comp.initialize_p1(imp, exp, grid_shape=(4, 3),
                       run_nml=dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2))
    imp.update(_import_state(case, None))
    comp.run(imp, exp, run_seconds=1000.0)
    assert comp.last_info.steps_done == 3
    ref = oracle.OracleSediment.from_config(comp.cfg)
    ref.init_concentrations()
    from mossco_code_b200.sediment import VARIABLE_NAMES
    cs = [imp[f"{v}_at_soil_surface"] for v in VARIABLE_NAMES]
    wz = [imp.get(f"{v}_z_velocity_at_soil_surface") for v in VARIABLE_NAMES]
    ref.get_boundary_conditions(imp["temperature_at_soil_surface"], cs, wz)
    ref.step(360.0, 2, 2)
    ref.step(280.0, 2, 1)
    assert rel_err(comp.sed.conc, ref.conc) <= 1e-11
    comp.finalize()


def test_component_restart_roundtrip(gpu):
    """ReadRestart (:1328-1489): <name>_in_soil fields of a finished run restart a fresh component."""
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("rst", 5, 4, 12, 0.004, seed=8, land_fraction=0.2)
    nml = dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2)
    a, b = FabmSedimentComponent(), FabmSedimentComponent()
    ia, ea, ib, eb = {}, {}, {}, {}
    a.initialize_p1(ia, ea, grid_shape=(5, 4), grid_mask=1 - case.mask, run_nml=nml)
    ia.update(_import_state(case, None))
    a.run(ia, ea, run_seconds=7200.0)
    b.initialize_p1(ib, eb, grid_shape=(5, 4), grid_mask=1 - case.mask, run_nml=nml)
    b.read_restart({k: v for k, v in ea.items() if k.endswith("_in_soil")}, eb)
    assert np.array_equal(b.sed.conc, a.sed.conc)
    ib.update(_import_state(case, None))
    a.run(ia, ea, run_seconds=3600.0)
    b.run(ib, eb, run_seconds=3600.0)
    assert np.array_equal(b.sed.conc, a.sed.conc)       # deterministic: bit-identical continuation
    a.finalize(); b.finalize()


def test_component_restart_through_netcdf_file(gpu, tmp_path):
    """The same through a file in mossco_netcdf.F90's layout (soil_netcdf.py): doubles survive bit for bit."""
    from scipy.io import netcdf_file
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("rstf", 6, 5, 12, 0.004, seed=9, land_fraction=0.2)
    nml = dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2)
    a, b = FabmSedimentComponent(), FabmSedimentComponent()
    ia, ea, ib, eb = {}, {}, {}, {}
    a.initialize_p1(ia, ea, grid_shape=(6, 5), grid_mask=1 - case.mask, run_nml=nml)
    ia.update(_import_state(case, None))
    a.run(ia, ea, run_seconds=3600.0)
    path = str(tmp_path / "restart_soil.nc")
    a.write_restart_file(path, 3600.0)
    a.run(ia, ea, run_seconds=3600.0)
    a.write_restart_file(path, 7200.0, append=True)
    nc = netcdf_file(path, "r", mmap=False)
    v = nc.variables["dissolved_oxygen_in_soil"]
    assert v.dimensions == ("time", "ungridded00012", "sedimentFluxes_2_O", "sedimentFluxes_1_O")
    assert v.shape == (2, 12, 5, 6) and v.missing_value == -1.0e30
    nc.close()
    b.initialize_p1(ib, eb, grid_shape=(6, 5), grid_mask=1 - case.mask, run_nml=nml)
    b.read_restart_file(path)                            # last record
    assert np.array_equal(b.sed.conc, a.sed.conc)
    ib.update(_import_state(case, None))
    a.run(ia, ea, run_seconds=3600.0)
    b.run(ib, eb, run_seconds=3600.0)
    assert np.array_equal(b.sed.conc, a.sed.conc)
    a.finalize(); b.finalize()


def test_component_on_a_mesh(gpu, tmp_path):
    """Mesh geometry (:391-446, :693-770): columns = owned elements, rank-1 surface fields, rank-2
    <name>_in_soil fields; the numbers are those of the same columns on an (n,1) grid."""
    from mossco_code_b200.component import FabmSedimentComponent
    from mossco_code_b200.sediment import VARIABLE_NAMES
    case = make_case("mesh", 23, 1, 12, 0.004, seed=12, land_fraction=0.2)
    nml = dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2)
    g, m = FabmSedimentComponent(), FabmSedimentComponent()
    ig, eg, im, em = {}, {}, {}, {}
    g.initialize_p1(ig, eg, grid_shape=(23, 1), grid_mask=1 - case.mask, run_nml=nml)
    m.initialize_p1(im, em, grid_shape=(23,), grid_mask=(1 - case.mask)[:, 0], run_nml=nml)
    st = _import_state(case, None)
    ig.update(st)
    im.update({k: v[:, 0].copy() for k, v in st.items()})           # rank-1 fields
    for _ in range(2):
        g.run(ig, eg, run_seconds=3600.0)
        m.run(im, em, run_seconds=3600.0)
    for v in VARIABLE_NAMES:
        assert em[f"{v}_in_soil"].shape == (23, 12) and em[f"{v}_upward_flux_at_soil_surface"].shape == (23,)
        assert np.array_equal(em[f"{v}_in_soil"], eg[f"{v}_in_soil"][:, 0, :])
        assert np.array_equal(em[f"{v}_upward_flux_at_soil_surface"], eg[f"{v}_upward_flux_at_soil_surface"][:, 0])
    assert em["temperature_in_soil"].shape == (23, 12) and em["denit_in_soil"].shape == (23, 12)
    # restart hand-over with rank-2 <name>_in_soil fields, and through a file
    r = FabmSedimentComponent()
    ir, er = {}, {}
    r.initialize_p1(ir, er, grid_shape=(23,), grid_mask=(1 - case.mask)[:, 0], run_nml=nml)
    r.read_restart({k: v for k, v in em.items() if k.endswith("_in_soil")}, er)
    assert np.array_equal(r.sed.conc, m.sed.conc)
    path = str(tmp_path / "mesh_restart.nc")
    m.write_restart_file(path)
    r2 = FabmSedimentComponent()
    r2.initialize_p1({}, {}, grid_shape=(23,), grid_mask=(1 - case.mask)[:, 0], run_nml=nml)
    r2.read_restart_file(path)
    assert np.array_equal(r2.sed.conc, m.sed.conc)
    for c in (g, m, r, r2):
        c.finalize()


def test_component_presimulation(gpu, oracle):
    """presimulation_years > 0: 1-D spin-up broadcast to every wet column (:557-632)."""
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("pre", 4, 3, 12, 0.004, seed=3, land_fraction=0.3)
    comp = FabmSedimentComponent()
    imp, exp = {}, {}
    years = 10.0 / 365.0
    comp.initialize_p1(imp, exp, grid_shape=(4, 3), grid_mask=1 - case.mask,
                       run_nml=dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2,
                                    presimulation_years=years, pel_Temp=5.0, pel_NO3=14.0, pel_NH4=4.0,
                                    pel_PO4=0.6, pel_O2=250.0, pflux_lDetC=2.0, pflux_sDetC=24.0,
                                    pflux_lDetP=0.08))
    from tests.cases import C1_BDYS, C1_FLUXES
    nml, par = oracle.from_config(comp.cfg)
    want = oracle.spinup_column(nml, par, 12, 0.004, 1.0, -0.9, C1_BDYS, C1_FLUXES, 240, method=2)
    conc = comp.sed.conc
    wet = case.mask == 0
    for i, j in zip(*np.nonzero(wet)):
        assert scaled_err(conc[i, j], want[0, 0]) <= 1e-8
    assert np.all(conc[~wet] == 1e20)
    comp.finalize()


def test_component_output_dat(gpu, tmp_path):
    """output.dat (component :677-685, :1734-1759): header + one block per output step, parseable the way
    examples/standalone/omexdia_p/plotsed1d.py parses it."""
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("out", 3, 2, 12, 0.004, seed=4)
    comp = FabmSedimentComponent()
    imp, exp = {}, {}
    path = tmp_path / "output.dat"
    comp.initialize_p1(imp, exp, grid_shape=(3, 2), output_path=str(path),
                       run_nml=dict(numlayers=12, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2, output=5))
    imp.update(_import_state(case, None))
    comp.run(imp, exp, run_seconds=3600.0)          # steps 0..9 -> blocks after steps 0 and 5
    comp.run(imp, exp, run_seconds=3600.0)          # steps 10..19 -> blocks after steps 10 and 15
    conc = comp.sed.conc
    comp.finalize()
    lines = path.read_text().splitlines()
    head = lines[0].split()
    assert head[:4] == ["time(s)", "depth(m)", "layer-height(m)", "porosity()"]
    assert "hzg_omexdia_p_denit" in head and head[4] == "hzg_omexdia_p_ldetC"
    blocks = [i for i, l in enumerate(lines) if len(l.split()) > 1 and l.split()[1] == "fluxes"]
    assert len(blocks) == 4
    assert [float(lines[i].split()[0]) for i in blocks] == [0.0, 1800.0, 3600.0, 5400.0]
    first = lines[blocks[0] + 1].split()
    assert len(first) == 4 + 8 + 1 and len(lines[blocks[0] + 1]) == 13 * 15 + 12
    assert first[0] == "0.000E+00" and first[1].endswith("E-002")
    rows = np.array([[float(x) for x in l.split()] for l in lines[blocks[3] + 1: blocks[3] + 13]])
    assert rows.shape == (12, 13) and np.all(rows[:, 0] == 5400.0)
    assert np.all(np.diff(rows[:, 1]) > 0)                       # depth increases
    assert FabmSedimentComponent._fortran_e(-1234.5678, 15, 4, 3).strip() == "-0.1235E+004"
    assert FabmSedimentComponent._fortran_e(0.99996, 15, 4, 3).strip() == "0.1000E+001"


@pytest.mark.parametrize("seconds", [1440.0, 3600.0])     # 2 pairs (state ends in the rotated-in staging buffer), 5 pairs
def test_component_masked_tile_pairs_chunk_major(gpu, seconds):
    """A masked K=40 tile driven through the component with fused pairs in several chunks (chunk-major Run):
    every Run uses the staging buffer (porosity import, 3-D exports) AND may rotate it in as the state buffer,
    so land cells of every <name>_in_soil export must stay at missing_value and the wet cells must carry the
    bits of the unfused, unchunked component."""
    from mossco_code_b200.component import FabmSedimentComponent
    from mossco_code_b200.sediment import VARIABLE_NAMES
    case = make_case("cm", 70, 23, 40, 0.0015, seed=29, land_fraction=0.3, par_max=30.0)
    land = case.mask > 0
    surf = np.asfortranarray(0.55 + 0.2 * np.random.default_rng(3).random((70, 23)))
    out = []
    for fused in (False, True):
        comp = FabmSedimentComponent()
        imp, exp = {}, {}
        comp.initialize_p1(imp, exp, grid_shape=(70, 23), grid_mask=1 - case.mask,
                           run_nml=dict(numlayers=40, dzmin=0.0015, dt=360.0, dt_min=1.0, ode_method=2))
        comp.sed.set_step_fusion("pairs" if fused else "off")
        comp.sed.set_exchange_chunks(4 if fused else 1)
        comp.sed.set_exchange_order(True)
        imp.update(_import_state(case, None))
        imp["porosity_at_soil_surface"] = surf
        snaps = []
        for it in range(4):
            comp.run(imp, exp, run_seconds=seconds)
            for v in VARIABLE_NAMES:
                assert np.all(exp[f"{v}_in_soil"][land] == 1e20), (it, v)
            snaps.append({k: np.array(a, copy=True) for k, a in exp.items()})
        if fused:
            assert comp.last_info.fused_steps == int(round(seconds / 360.0))
        out.append(snaps)
        comp.finalize()
    for a, b in zip(*out):
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("shape,pin", [((37, 11), False), ((64, 8), True), ((64, 8), False), ((9, 1), True)])
def test_asynchronous_state_export_is_a_snapshot(gpu, shape, pin):
    """msed_export_state_begin / _wait (the <name>_in_soil write-back of fabm_sediment_component.F90:1773-1822 at an
    output cadence): the copy is handed to the engine in slices under the stepping calls that follow, and what
    arrives is the state at the moment of begin -- padded device rows (37x11) and unpadded ones (64x8), pinned and
    pageable destinations, Runs with host buffers and plain steps in between, two exports in a row."""
    import torch
    from mossco_code_b200 import SedimentDriver, default_config
    from mossco_code_b200.sediment import PARTICULATE
    inum, jnum = shape
    case = make_case("xexp", inum, jnum, 20, 0.003, seed=31, land_fraction=0.2 if inum > 9 else 0.0)
    cfg = default_config(inum=inum, jnum=jnum, knum=20, dzmin=0.003, dt_min=1.0)
    keep = []

    def buf():
        if not pin:
            return np.full((inum, jnum, 20, 8), -7.0, order="F")
        t = torch.full((8, 20, jnum, inum), -7.0, dtype=torch.float64).pin_memory()
        keep.append(t)
        return t.numpy().T

    temp = np.asfortranarray(case.bdys[:, :, 0])
    cs = [np.asfortranarray((-case.fluxes[:, :, n]) if PARTICULATE[n] else case.bdys[:, :, n + 1]) for n in range(8)]
    wz = [np.ones(shape, order="F") if PARTICULATE[n] else None for n in range(8)]
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        assert sed.step(360.0, 2, 3) == 0
        want1 = sed.conc
        out1 = buf()
        sed.export_state_begin(out1)
        for _ in range(2):                                   # Runs with host buffers while the export is pending
            rc, _ = sed.run_exchange(360.0, 2, 3600.0, temp, cs, wz)
            assert rc == 0
        assert sed.step(360.0, 2, 5) == 0                    # ... and plain steps
        want2 = sed.conc
        out2 = buf()
        sed.export_state_begin(out2)                          # completes the first export, starts the second
        assert np.array_equal(out1, want1)
        assert sed.step(360.0, 0, 2) == 0
        sed.export_state_wait()
        assert np.array_equal(out2, want2)
        assert not np.array_equal(want1, want2) and not np.array_equal(sed.conc, want2)
        sed.export_state_wait()                               # nothing pending: a no-op


def test_component_export_cadence(gpu):
    """export_cadence: every n-th Run starts the asynchronous export; export_ready() completes it and points the
    <var>_in_soil fields at the buffer, which holds the state of that Run while later Runs have moved on."""
    from mossco_code_b200.component import FabmSedimentComponent
    case = make_case("xcad", 12, 7, 15, 0.004, seed=8)
    comp = FabmSedimentComponent()
    imp, exp = {}, {}
    comp.initialize_p0(imp, exp)
    comp.initialize_p1(imp, exp, grid_shape=(12, 7), grid_mask=1 - case.mask,
                       run_nml=dict(numlayers=15, dzmin=0.004, dt=360.0, dt_min=1.0, ode_method=2))
    comp.initialize_p2(imp, exp)
    imp.update(_import_state(case, np.random.default_rng(0)))
    comp.export_3d_every_run = False
    comp.export_cadence = 2
    states = []
    for r in range(4):
        assert comp.run(imp, exp, run_seconds=1800.0) == 0
        states.append(comp.sed.conc)
        if r == 2:
            assert comp.export_ready(exp)                     # the export started after Run 2 (index 1)
            assert np.array_equal(comp.export_buffer, states[1])
            assert np.array_equal(exp["dissolved_oxygen_in_soil"], states[1][:, :, :, 6])
    assert comp.export_ready(exp)                             # the one started after Run 4
    assert np.array_equal(comp.export_buffer, states[3])
    assert not comp.export_ready(exp)
    comp.finalize()
