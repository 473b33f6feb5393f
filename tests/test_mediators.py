"""Name handling and branch selection of the mediator mirrors (mossco_code_b200/mediators.py) against the
lookups of the Fortran mediators -- CPU only: the sediment driver is replaced by a recorder."""
import numpy as np
import pytest

from mossco_code_b200.component import ComponentError
from mossco_code_b200.mediators import (ESMF_RC_ARG_BAD, ESMF_RC_NOT_FOUND, BenthicPelagicCoupler,
                                        PelagicBenthicCoupler, PelagicSoilConnector, SoilPelagicConnector)
from mossco_code_b200.sediment import VARIABLE_NAMES

SH = (4, 3)


class Recorder:
    """Stands in for SedimentDriver: records what the mediators ask the device for."""
    def __init__(self):
        self.calls = []

    def _answer(self, kind, want, kw):
        self.calls.append((kind, tuple(want), kw))
        return {k: np.full(SH, float(i + 1), order="F") for i, k in enumerate(want)}

    def soil_pelagic_connector(self, want, **kw):
        return self._answer("s2p", want, kw)

    def benthic_pelagic_coupler(self, want, **kw):
        return self._answer("b2p", want, kw)

    def pelagic_benthic_coupler(self, **fields):
        self.calls.append(("p2b", tuple(sorted(fields)), fields))

    def pelagic_soil_connector(self, params=None, **fields):
        self.calls.append(("p2s", tuple(sorted(fields)), dict(fields, params=params)))

    def set_compat(self, **kw):
        self.calls.append(("compat", (), kw))


def soil_export():
    return {f"{v}_upward_flux_at_soil_surface": np.zeros(SH, order="F") for v in VARIABLE_NAMES}


def test_soil_pelagic_connector_finds_ecosmo_names_and_fills_in_place():
    rec = Recorder()
    exp = {n: np.zeros(SH, order="F") for n in (
        "hzg_ecosmo_no3_upward_flux_at_soil_surface", "hzg_ecosmo_nh4_upward_flux_at_soil_surface",
        "hzg_ecosmo_pho_upward_flux_at_soil_surface", "hzg_ecosmo_oxy_upward_flux_at_soil_surface",
        "some_other_field")}
    keep = exp["hzg_ecosmo_no3_upward_flux_at_soil_surface"]
    med = SoilPelagicConnector(rec, dinflux_const=0.3, convertN=2.0)
    assert med.run(soil_export(), exp) == 0
    kind, want, kw = rec.calls[0]
    assert kind == "s2p" and set(want) == {"nitrate", "ammonium", "DIP", "oxygen"}      # only oxygen: oxy - odu branch
    assert kw == dict(dinflux_const=0.3, convertN=2.0)
    assert exp["hzg_ecosmo_no3_upward_flux_at_soil_surface"] is keep and keep[0, 0] == 1.0 + want.index("nitrate")
    assert not exp["some_other_field"].any()


def test_soil_pelagic_connector_maecs_names_and_missing_soil_field():
    rec = Recorder()
    exp = {"Dissolved_Inorganic_Nitrogen_DIN_nutN_upward_flux_at_soil_surface": None,
           "nutrients_upward_flux_at_soil_surface": np.zeros(SH, order="F"),
           "Dissolved_Inorganic_Phosphorus_DIP_nutP_upward_flux_at_soil_surface": np.zeros(SH, order="F"),
           "Detritus_Carbon_detC_upward_flux_at_soil_surface": np.zeros(SH, order="F"),
           "dissolved_reduced_substances_upward_flux_at_soil_surface": np.zeros(SH, order="F")}
    SoilPelagicConnector(rec).run(soil_export(), exp)
    assert set(rec.calls[0][1]) == {"DIN", "DIP", "detC", "odu"}
    assert exp["nutrients_upward_flux_at_soil_surface"].any()           # the first name the state really holds
    imp = soil_export()
    del imp["dissolved_oxygen_upward_flux_at_soil_surface"]
    with pytest.raises(ComponentError) as e:
        SoilPelagicConnector(rec).run(imp, exp)
    assert e.value.rc == ESMF_RC_ARG_BAD                                 # soil_pelagic_connector.F90:548-552
    assert SoilPelagicConnector(rec).run(soil_export(), {}) == 0 and len(rec.calls) == 1   # nothing wanted


def test_benthic_pelagic_coupler_din_branch_follows_the_ammonium_lookup():
    rec = Recorder()
    din = {"DIN_upward_flux_at_soil_surface": np.zeros(SH, order="F"),
           "nitrate_upward_flux_at_soil_surface": np.zeros(SH, order="F")}
    BenthicPelagicCoupler(rec).run(soil_export(), din)                  # no ammonium field: DIN branch (:224-236)
    assert set(rec.calls[-1][1]) == {"DIN", "nitrate"}
    both = dict(din, **{"ammonium_upward_flux_at_soil_surface": np.zeros(SH, order="F")})
    BenthicPelagicCoupler(rec).run(soil_export(), both)
    assert set(rec.calls[-1][1]) == {"nitrate", "ammonium"}
    with pytest.raises(ComponentError) as e:                             # neither: the reference finalizes (:231)
        BenthicPelagicCoupler(rec).run(soil_export(), {"oxygen_upward_flux_at_soil_surface": np.zeros(SH)})
    assert e.value.rc == ESMF_RC_NOT_FOUND


def test_pelagic_benthic_coupler_takes_the_bottom_layer_and_checks_required_fields():
    rec = Recorder()
    rng = np.random.default_rng(1)
    f3 = lambda: rng.random(SH + (5,))
    imp = {"temperature_in_water": f3(), "oxygen_in_water": f3(), "detN_in_water": f3(),
           "detN_z_velocity_in_water": f3(), "nutrients_in_water": rng.random(SH),
           "Detritus_Carbon_detC_in_water": f3()}
    assert PelagicBenthicCoupler(rec).run(imp) == 0
    kind, keys, fields = rec.calls[0]
    assert set(keys) == {"temperature", "oxygen", "detN", "detN_z_velocity", "DIN", "detC"}
    assert np.array_equal(fields["oxygen"], imp["oxygen_in_water"][:, :, 0]) and fields["oxygen"].flags.f_contiguous
    assert np.array_equal(fields["DIN"], imp["nutrients_in_water"])
    del imp["nutrients_in_water"]
    with pytest.raises(ComponentError):
        PelagicBenthicCoupler(rec).run(imp)
    imp["nutrients_in_water"] = rng.random(SH)
    del imp["detN_z_velocity_in_water"]
    with pytest.raises(ComponentError) as e:
        PelagicBenthicCoupler(rec).run(imp)
    assert e.value.rc == ESMF_RC_NOT_FOUND


def test_pelagic_soil_connector_names_velocity_splice_and_export_side_fields():
    """pelagic_soil_connector.F90: detritus under its eight alternative names (:923-930), the velocity field named
    after the detritus field that was found (:961-972), water depth and TKE looked up in the EXPORT state
    (:1126, :1164), nothing transferred without detritus (:919), ESMF_RC_NOT_FOUND without any nitrogen (:1846)."""
    rec = Recorder()
    rng = np.random.default_rng(2)
    f3 = lambda: rng.random(SH + (5,))
    imp = {"temperature_in_water": f3(), "hzg_ecosmo_oxy_in_water": f3(), "hzg_ecosmo_det_in_water": f3(),
           "hzg_ecosmo_det_z_velocity_in_water": f3(), "dissolved_ammonium_nh3_in_water": f3(),
           "nitrate_in_water": f3(), "downwelling_photosynthetic_radiative_flux_in_water": f3(),
           "detN_z_velocity_in_water": f3()}                     # a velocity of ANOTHER name must not be picked
    exp = {"water_depth_at_soil_surface": rng.random(SH), "turbulent_diffusivity_of_momentum_at_soil_surface": rng.random(SH)}
    med = PelagicSoilConnector(rec, sinking_factor=0.2, head_compat=True)
    assert med.run(imp, exp) == 0
    assert rec.calls[0] == ("compat", (), dict(p2s_head=True))
    kind, keys, fields = rec.calls[1]
    assert kind == "p2s" and set(keys) == {"temperature", "par", "oxygen", "detN", "detN_z_velocity", "ammonium",
                                           "nitrate", "water_depth", "tke"}
    assert fields["params"] == dict(sinking_factor=0.2)
    assert np.array_equal(fields["detN_z_velocity"], imp["hzg_ecosmo_det_z_velocity_in_water"][:, :, 0])
    assert np.array_equal(fields["tke"], exp["turbulent_diffusivity_of_momentum_at_soil_surface"])
    del imp["hzg_ecosmo_det_z_velocity_in_water"]
    with pytest.raises(ComponentError) as e:
        med.run(imp, exp)
    assert e.value.rc == ESMF_RC_NOT_FOUND
    del imp["hzg_ecosmo_det_in_water"]
    n = len(rec.calls)
    assert med.run(imp, exp) == 0 and len(rec.calls) == n        # no detritus: the routine returns early
    imp2 = {"temperature_in_water": f3(), "detritus_in_water": f3(), "detritus_z_velocity_in_water": f3()}
    with pytest.raises(ComponentError) as e:
        PelagicSoilConnector(rec).run(imp2, {})
    assert e.value.rc == ESMF_RC_NOT_FOUND


def test_pelagic_soil_connector_closed_forms(oracle):
    """The restated algebra on hand-computed numbers (pelagic_soil_connector.F90:1063-1232, :1816-2110)."""
    one = lambda v: np.full((1, 1), float(v))
    eps = float(np.float32(1e-5))
    f = dict(detN=one(2.0), detN_z_velocity=one(-1e-4), detC=one(13.0), DIN=one(8.0), nitrate=one(5.0),
             oxygen=one(-12.0), water_depth=one(0.3), tke=one(500.0))
    par = dict(oracle.P2S_DEFAULTS)
    cs, wz = oracle.pelagic_soil_connector((1, 1), **f)
    cn = 13.0 / (eps + 2.0)
    fl = (1 - par["NC_sdet"] * cn) / (par["NC_ldet"] - par["NC_sdet"])
    assert 0 < fl < cn
    assert cs[0][0, 0] == fl * 1.0 * 2.0 and cs[1][0, 0] == (cn - fl) * 1.0 * 2.0
    hsd = par["half_sedimentation_depth"]
    env = 1.0 * (0.3 * 0.3) / (0.3 * 0.3 + hsd * hsd)
    env = env * 1000.0 / (500.0 + 1000.0)
    env = env + par["sinking_factor_min"] / 0.3
    x = 13.0 / 60.0
    env = env * 1.0 / (1.0 + (x * x) * (x * x))
    assert wz[0][0, 0] == 0.3 * env * -1e-4 == wz[1][0, 0] == wz[2][0, 0]
    assert cs[2][0, 0] == 1.0 / 16.0 * 1.0 * 2.0                          # detP from detN (:1521)
    assert cs[5][0, 0] == 1.0 * (8.0 - 5.0) and cs[4][0, 0] == 5.0        # ammonium = DIN - nitrate (:1821)
    assert cs[3][0, 0] == 1.0 * (1.0 / 16.0 * 1.0 * 8.0)                  # phosphate from DIN (:2099)
    assert cs[6][0, 0] == 0.0 and cs[7][0, 0] == 12.0                     # negative oxygen is odu
    # C:N beyond the labile end member: everything labile (:1084-1087); below the semilabile one: nothing (:1088)
    cs, _ = oracle.pelagic_soil_connector((1, 1), **dict(f, detC=one(2.0)))
    cn = 2.0 / (eps + 2.0)
    assert cs[0][0, 0] == cn * 2.0 and cs[1][0, 0] == 0.0
    cs, _ = oracle.pelagic_soil_connector((1, 1), **dict(f, detC=one(400.0)))
    assert cs[0][0, 0] == 0.0 and cs[1][0, 0] == (400.0 / (eps + 2.0)) * 2.0
    # HEAD: the concentration fields receive the velocity expression with detN in it (:1293-1295)
    cs, wz = oracle.pelagic_soil_connector((1, 1), head_compat=True, **f)
    assert cs[0][0, 0] == 0.3 * env * 2.0 == cs[1][0, 0] and wz[0][0, 0] == 0.0 and wz[2][0, 0] == 0.3 * env * 2.0
