"""Host-side mirrors of the mediators either side of ``fabm_sediment_component``.

The reference's couplers are ESMF coupler components whose ``Run(importState, exportState)`` looks fields up
by name -- each quantity under a list of alternative names, because every pelagic model spells its variables
differently -- and does 2-D field algebra on what it finds.  With the sediment resident on the GPU that
algebra runs there (``msed_pelagic_benthic_coupler``, ``msed_benthic_pelagic_coupler``,
``msed_soil_pelagic_connector``); these classes keep the name handling, the "which fields are present"
branches and the error behaviour of the Fortran, so that the tests read like a coupled MOSSCO configuration.
States are plain dicts ``name -> numpy array`` as in ``component.py``; pelagic ``*_in_water`` fields may be
rank 3 (i, j, layer), in which case the bottom layer ``[:, :, 0]`` is taken (``lbnd(3)`` in the Fortran).

    pelagic model --PelagicBenthicCoupler / PelagicSoilConnector--> sediment
    sediment --BenthicPelagicCoupler / SoilPelagicConnector--> pelagic model
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .component import ESMF_SUCCESS, ComponentError
from .sediment import VARIABLE_NAMES, SedimentDriver

ESMF_RC_ARG_BAD = 505       # ESMF_RC_ARG_BAD
ESMF_RC_NOT_FOUND = 541     # ESMF_RC_NOT_FOUND
State = Dict[str, np.ndarray]


def _first(state: State, names: Sequence[str]) -> Optional[str]:
    """``mossco_state_get(state, (/names/), ...)``: the first of the alternative names the state holds."""
    for n in names:
        if n in state and state[n] is not None:
            return n
    return None


def _bottom(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    return np.asfortranarray(a[:, :, 0] if a.ndim == 3 else a)


class PelagicBenthicCoupler:
    """``pelagic_benthic_coupler`` (src/mediators/pelagic_benthic_coupler.F90:281-492): bottom-water
    concentrations and sinking velocities of the pelagic model become the sediment's
    ``*_at_soil_surface`` / ``*_z_velocity_at_soil_surface`` import fields.  Here the result goes straight
    into the sediment's boundary arrays on the device (``get_boundary_conditions`` fused in); the
    ``export_state`` receives the assembled fields only when ``fill_export`` is set."""

    IMPORT_NAMES = {
        "temperature": ("temperature_in_water",),                                                   # :323
        "oxygen": ("concentration_of_dissolved_oxygen_in_water", "oxygen_in_water",
                   "dissolved_oxygen_in_water"),                                                    # :332-335
        "detN": ("detritus_in_water", "detN_in_water", "Detritus_Nitrogen_detN_in_water"),          # :353-356
        "detN_z_velocity": ("detritus_z_velocity_in_water", "detN_z_velocity_in_water",
                            "Detritus_Nitrogen_detN_z_velocity_in_water"),                          # :366-369
        "detC": ("Detritus_Carbon_detC_in_water",),                                                 # :378
        "detP": ("detP_in_water", "Detritus_Phosphorus_detP_in_water"),                             # :413-415
        "detP_z_velocity": ("detP_z_velocity_in_water", "Detritus_Phosphorus_detP_z_velocity_in_water"),  # :423-425
        "nitrate": ("nitrate_in_water",),                                                           # :438
        "DIN": ("nutrients_in_water", "DIN_in_water", "Dissolved_Inorganic_Nitrogen_DIN_nutN_in_water"),  # :440-443
        "ammonium": ("ammonium_in_water",),                                                         # :445
        "DIP": ("DIP_in_water", "phosphate_in_water", "Dissolved_Inorganic_Phosphorus_DIP_nutP_in_water"),  # :471-474
    }
    REQUIRED = ("temperature", "oxygen", "detN", "detN_z_velocity")

    def __init__(self, sed: SedimentDriver):
        self.sed = sed

    def run(self, import_state: State, export_state: Optional[State] = None, fill_export: bool = False) -> int:
        fields = {}
        for key, names in self.IMPORT_NAMES.items():
            n = _first(import_state, names)
            if n is not None:
                fields[key] = _bottom(import_state[n])
        for key in self.REQUIRED:       # the Fortran aborts through ESMF_LogFoundError on these (:324,:336,:357,:370)
            if key not in fields:
                raise ComponentError(ESMF_RC_NOT_FOUND, f"pelagic_benthic_coupler: no {self.IMPORT_NAMES[key][-1]}")
        # nitrate / ammonium / DIP fall back on DIN (:446-480): without it there is nothing to fall back on
        if "DIN" not in fields and not all(k in fields for k in ("nitrate", "ammonium", "DIP")):
            raise ComponentError(ESMF_RC_NOT_FOUND, "pelagic_benthic_coupler: neither DIN nor nitrate+ammonium+DIP")
        self.sed.pelagic_benthic_coupler(**fields)
        if fill_export and export_state is not None:
            bdys, fluxes = self.sed.bdys, self.sed.fluxes
            export_state["temperature_at_soil_surface"] = bdys[:, :, 0].copy()
            for n, v in enumerate(VARIABLE_NAMES[3:], start=3):
                export_state[f"{v}_at_soil_surface"] = bdys[:, :, n + 1].copy()
            for n, v in enumerate(VARIABLE_NAMES[:3]):
                export_state[f"{v}_sinking_flux_at_soil_surface"] = fluxes[:, :, n].copy()
        return ESMF_SUCCESS


class PelagicSoilConnector:
    """``pelagic_soil_connector`` (src/mediators/pelagic_soil_connector.F90:176-2122), the generic successor of
    ``pelagic_benthic_coupler``: the same direction, with a C:N-dependent split of detritus into the labile and
    semilabile pools (:1082-1095), an environmental sinking factor from water depth, turbulent kinetic energy
    and the detritus concentration (:1150-1232), and ammonium / nitrate / phosphate assembled from whichever of
    nitrate, ammonium, DIN, DIP the pelagic model exports (:1775-2110).  Namelist ``/pelagic_soil_connector/``
    (:146-148): sinking_factor, sinking_factor_min, NC_ldet, NC_sdet, half_sedimentation_depth,
    critical_detritus, half_sedimentation_tke, convertN, convertP.  ``head_compat`` reproduces the HEAD
    revision's detritus / phosphate branches (``msed_set_compat``, include/msed.h)."""

    IMPORT_NAMES = {
        "temperature": ("temperature_in_water",),                                                   # :351
        "par": ("photosynthetically_active_radiation_in_water",
                "downwelling_photosynthetic_radiative_flux_in_water"),                              # :331-332
        "oxygen": ("concentration_of_dissolved_oxygen_in_water", "oxygen_in_water", "dissolved_oxygen_oxy_in_water",
                   "hzg_ecosmo_oxy_in_water", "dissolved_oxygen_in_water"),                         # :652-656
        "odu": ("dissolved_reduced_substances_odu_in_water", "dissolved_reduced_substances_in_water"),   # :697-698
        "detN": ("detritus_at_soil_surface", "detritus_in_water", "detN_at_soil_surface", "detN_in_water",
                 "Detritus_Nitrogen_detN_at_soil_surface", "Detritus_Nitrogen_detN_in_water",
                 "hzg_ecosmo_det_at_soil_surface", "hzg_ecosmo_det_in_water"),                      # :923-930
        "detC": ("Detritus_Carbon_detC_at_soil_surface", "Detritus_Carbon_detC_in_water"),          # :1027-1028
        "detP": ("detP_in_water", "Detritus_Phosphorus_detP_in_water"),                             # :1530-1531
        "detP_z_velocity": ("detP_z_velocity_in_water", "Detritus_Phosphorus_detP_z_velocity_in_water"),  # :1604-1605
        "nitrate": ("nitrate_in_water",),                                                           # :1655
        "DIN": ("nutrients_in_water", "DIN_in_water", "Dissolved_Inorganic_Nitrogen_DIN_nutN_in_water"),  # :1685-1687
        "ammonium": ("ammonium_in_water", "dissolved_ammonium_nh3_in_water"),                       # :1725-1726
        "DIP": ("DIP_in_water", "phosphate_in_water", "Dissolved_Inorganic_Phosphorus_DIP_nutP_in_water"),  # :1988-1990
    }
    # looked up in the EXPORT state: the sediment side of the coupling holds them (:1126, :1164-1165)
    EXPORT_SIDE_NAMES = {
        "water_depth": ("water_depth_at_soil_surface",),
        "tke": ("turbulent_kinetic_energy_at_soil_surface", "turbulent_diffusivity_of_momentum_at_soil_surface"),
    }

    def __init__(self, sed: SedimentDriver, head_compat: bool = False, **namelist):
        self.sed = sed
        self.namelist = namelist
        self.head_compat = head_compat

    def run(self, import_state: State, export_state: Optional[State] = None, fill_export: bool = False) -> int:
        fields = {}
        for key, names in self.IMPORT_NAMES.items():
            n = _first(import_state, names)
            if n is not None:
                fields[key] = _bottom(import_state[n])
        n = _first(import_state, self.IMPORT_NAMES["detN"])
        if n is None:           # "skip the rest of this routine" (:919-921): nothing is transferred
            return ESMF_SUCCESS
        # the velocity field carries the detritus field's name with z_velocity spliced in (:961-972)
        for suffix in ("_in_water", "_at_soil_surface"):
            if n.endswith(suffix):
                vname = n[:-len(suffix)] + "_z_velocity" + suffix
                break
        if import_state.get(vname) is None:
            raise ComponentError(ESMF_RC_NOT_FOUND, f"pelagic_soil_connector: no {vname}")
        fields["detN_z_velocity"] = _bottom(import_state[vname])
        if "temperature" not in fields:
            raise ComponentError(ESMF_RC_NOT_FOUND, "pelagic_soil_connector: no temperature_in_water")
        if not any(k in fields for k in ("nitrate", "ammonium", "DIN")):        # rc = ESMF_RC_NOT_FOUND, :1846
            raise ComponentError(ESMF_RC_NOT_FOUND, "pelagic_soil_connector: none of nitrate, ammonium, DIN")
        for key, names in self.EXPORT_SIDE_NAMES.items():
            e = _first(export_state or {}, names)
            if e is not None:
                fields[key] = _bottom(export_state[e])
        self.sed.set_compat(p2s_head=self.head_compat)
        self.sed.pelagic_soil_connector(params=self.namelist, **fields)
        if fill_export and export_state is not None:
            bdys, fluxes = self.sed.bdys, self.sed.fluxes
            export_state["temperature_at_soil_surface"] = bdys[:, :, 0].copy()
            for n, v in enumerate(VARIABLE_NAMES[3:], start=3):
                export_state[f"{v}_at_soil_surface"] = bdys[:, :, n + 1].copy()
            for n, v in enumerate(VARIABLE_NAMES[:3]):
                export_state[f"{v}_sinking_flux_at_soil_surface"] = fluxes[:, :, n].copy()
        return ESMF_SUCCESS


class _UpwardFluxMediator:
    """Common part of the two soil -> pelagic mediators: find the pelagic flux fields in the export state
    under their alternative names, have the device compute them, copy them in."""
    EXPORT_NAMES: Dict[str, Sequence[str]] = {}
    SOIL_REQUIRED: Sequence[str] = ()

    def __init__(self, sed: SedimentDriver, **namelist):
        self.sed = sed
        self.namelist = namelist

    def _wanted(self, export_state: State) -> Dict[str, str]:
        return {key: n for key, names in self.EXPORT_NAMES.items()
                if (n := _first(export_state, names)) is not None}

    def _check_soil(self, import_state: State):
        for v in self.SOIL_REQUIRED:
            if f"{v}_upward_flux_at_soil_surface" not in import_state:
                raise ComponentError(ESMF_RC_ARG_BAD, f"expected exactly one field for {v}_upward_flux_at_soil_surface")

    def _store(self, export_state: State, found: Dict[str, str], values: Dict[str, np.ndarray]):
        for key, name in found.items():
            dst = export_state[name]
            if isinstance(dst, np.ndarray) and dst.shape == values[key].shape:
                dst[...] = values[key]                      # the pelagic model's field keeps its memory
            else:
                export_state[name] = values[key]


class SoilPelagicConnector(_UpwardFluxMediator):
    """``soil_pelagic_connector`` (src/mediators/soil_pelagic_connector.F90:179-981).  ``import_state`` is the
    sediment's export state (its ``<var>_upward_flux_at_soil_surface`` fields are what the device holds);
    ``export_state`` holds the pelagic model's flux fields, filled in place.  Namelist
    ``/soil_pelagic_connector/`` (:140): dinflux_const, dipflux_const, convertN, convertP."""

    EXPORT_NAMES = {
        "nitrate": ("nitrate_upward_flux_at_soil_surface", "hzg_ecosmo_no3_upward_flux_at_soil_surface"),      # :320-321
        "ammonium": ("ammonium_upward_flux_at_soil_surface", "dissolved_ammonium_nh3_upward_flux_at_soil_surface",
                     "hzg_ecosmo_nh4_upward_flux_at_soil_surface"),                                            # :375-377
        "DIN": ("nutrients_upward_flux_at_soil_surface", "DIN_upward_flux_at_soil_surface",
                "Dissolved_Inorganic_Nitrogen_DIN_nutN_upward_flux_at_soil_surface"),                          # :423-425
        "DIP": ("DIP_upward_flux_at_soil_surface", "phosphate_upward_flux_at_soil_surface",
                "Dissolved_Inorganic_Phosphorus_DIP_nutP_upward_flux_at_soil_surface",
                "hzg_ecosmo_pho_upward_flux_at_soil_surface"),                                                 # :491-494
        "oxygen": ("oxygen_upward_flux_at_soil_surface", "dissolved_oxygen_oxy_upward_flux_at_soil_surface",
                   "hzg_ecosmo_oxy_upward_flux_at_soil_surface"),                                              # :596-598
        "odu": ("dissolved_reduced_substances_upward_flux_at_soil_surface",),                                  # :625
        "detN": ("detritus_upward_flux_at_soil_surface", "detN_upward_flux_at_soil_surface",
                 "Detritus_Nitrogen_detN_upward_flux_at_soil_surface"),                                        # :736-738
        "detC": ("Detritus_Carbon_detC_upward_flux_at_soil_surface",),                                         # :818-822
        "detP": ("detP_upward_flux_at_soil_surface", "Detritus_Phosphorus_detP_upward_flux_at_soil_surface"),  # :884-885
    }
    # fieldCount /= 1 -> ESMF_RC_ARG_BAD (:242-248, :269-274, :294-299, :548-552, :573-577)
    SOIL_REQUIRED = ("mole_concentration_of_phosphate", "mole_concentration_of_nitrate", "mole_concentration_of_ammonium",
                     "dissolved_oxygen", "dissolved_reduced_substances")

    def run(self, import_state: State, export_state: State) -> int:
        self._check_soil(import_state)
        found = self._wanted(export_state)
        if found:
            self._store(export_state, found, self.sed.soil_pelagic_connector(want=tuple(found), **self.namelist))
        return ESMF_SUCCESS


class BenthicPelagicCoupler(_UpwardFluxMediator):
    """``benthic_pelagic_coupler`` (src/mediators/benthic_pelagic_coupler.F90:188-287), the predecessor of
    ``soil_pelagic_connector``: same direction, detritus nitrogen from the carbon fluxes through the N:C
    ratios (:258), oxygen as oxygen minus reduced substances (:281).  Namelist (:39-43, :205-206):
    dinflux_const, dipflux_const, convertN, NC_fdet, NC_sdet."""

    EXPORT_NAMES = {
        "nitrate": ("nitrate_upward_flux_at_soil_surface",),                                                   # :220
        "ammonium": ("ammonium_upward_flux_at_soil_surface",),                                                 # :222
        "DIN": ("nutrients_upward_flux_at_soil_surface", "DIN_upward_flux_at_soil_surface",
                "Dissolved_Inorganic_Nitrogen_DIN_nutN_upward_flux_at_soil_surface"),                          # :227-230
        "DIP": ("DIP_upward_flux_at_soil_surface", "phosphate_upward_flux_at_soil_surface",
                "Dissolved_Inorganic_Phosphorus_DIP_nutP_upward_flux_at_soil_surface"),                        # :238-241
        "detN": ("detritus_upward_flux_at_soil_surface", "detN_upward_flux_at_soil_surface",
                 "Detritus_Nitrogen_detN_upward_flux_at_soil_surface"),                                        # :253-256
        "detC": ("Detritus_Carbon_detC_upward_flux_at_soil_surface",),                                         # :261-262
        "detP": ("detP_upward_flux_at_soil_surface", "Detritus_Phosphorus_detP_upward_flux_at_soil_surface"),  # :269-271
        "oxygen": ("oxygen_upward_flux_at_soil_surface",),                                                     # :277
    }

    def run(self, import_state: State, export_state: State) -> int:
        found = self._wanted(export_state)
        # "weak check" of the reference (:224-226): the DIN branch is taken exactly when the export state has
        # no ammonium flux field (nitrc is the status of that last lookup), and then a DIN field must exist
        if "ammonium" in found:
            found.pop("DIN", None)
        elif "DIN" not in found:
            raise ComponentError(ESMF_RC_NOT_FOUND, "benthic_pelagic_coupler: no ammonium and no DIN flux field "
                                                    "in the export state (ESMF_Finalize at :231)")
        if found:
            self._store(export_state, found, self.sed.benthic_pelagic_coupler(want=tuple(found), **self.namelist))
        return ESMF_SUCCESS
