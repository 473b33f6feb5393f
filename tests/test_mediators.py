"""Name handling and branch selection of the mediator mirrors (mossco_code_b200/mediators.py) against the
lookups of the Fortran mediators -- CPU only: the sediment driver is replaced by a recorder."""
import numpy as np
import pytest

from mossco_code_b200.component import ComponentError
from mossco_code_b200.mediators import (ESMF_RC_ARG_BAD, ESMF_RC_NOT_FOUND, BenthicPelagicCoupler,
                                        PelagicBenthicCoupler, SoilPelagicConnector)
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
