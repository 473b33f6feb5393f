"""ctypes binding of the C ABI in ``include/msed.h`` (libmsed_b200.so).

This is the only way Python reaches the CUDA path; there is no fallback.  If the shared library
has not been built (``python -c 'import __graft_entry__ as g; g.build()'``) importing the
symbols raises ``MsedLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os

NVAR = 8
MAX_LAYERS = 64
ABI_VERSION = 1

EULER, RUNGE_KUTTA_4, ADAPTIVE_EULER, RUNGE_KUTTA_4_38 = 0, 1, 2, 3
MODEL_OMEXDIA_P, MODEL_NONE, MODEL_TEST_SOLVER = 0, 1, 2

OK, NAN_DETECTED, BAD_DOMAIN = 0, 1, 2
ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_ALLOC, ERR_STATE = -1, -2, -3, -4, -5

FIELDS = {
    "porosity": 0, "layer_height": 1, "layer_center_depth": 2, "temperature": 3,
    "photosynthetically_active_radiation": 4, "biomass": 5, "bioturbation": 6,
    "weighted_toc": 7, "denit": 8, "intf_porosity": 9, "flux_cap": 10,
}


class MsedLibraryError(RuntimeError):
    pass


class MsedError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"msed error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """``msed_config`` -- sed_nml + run_nml + hzg_omexdia_p namelists in one struct."""
    _fields_ = [
        ("abi_version", C.c_int32), ("inum", C.c_int32), ("jnum", C.c_int32), ("knum", C.c_int32),
        ("device", C.c_int32), ("model", C.c_int32),
        ("dzmin", C.c_double),
        ("diffusivity", C.c_double), ("bioturbation", C.c_double), ("porosity_max", C.c_double),
        ("porosity_fac", C.c_double), ("k_par", C.c_double), ("pom_flux_max", C.c_double),
        ("bioturbation_depth", C.c_double), ("bioturbation_min", C.c_double),
        ("bioturb_k_l", C.c_double), ("bioturb_L1", C.c_double), ("bioturb_L2", C.c_double),
        ("bioturb_beta", C.c_double), ("bioturb_b", C.c_double), ("bioturb_dry_density", C.c_double),
        ("bioturbation_profile", C.c_int32), ("distributed_pom_flux", C.c_int32),
        ("dt_min", C.c_double), ("relative_change_min", C.c_double),
        ("bcup_dissolved_variables", C.c_int32), ("adaptive_solver_diagnostics", C.c_int32),
        ("rLabile", C.c_double), ("rSemilabile", C.c_double), ("NCrLdet", C.c_double),
        ("NCrSdet", C.c_double), ("PAds", C.c_double), ("PAdsODU", C.c_double),
        ("NH3Ads", C.c_double), ("CprodMax", C.c_double), ("rnit", C.c_double),
        ("ksO2nitri", C.c_double), ("rODUox", C.c_double), ("ksO2oduox", C.c_double),
        ("ksO2oxic", C.c_double), ("ksNO3denit", C.c_double), ("kinO2denit", C.c_double),
        ("kinNO3anox", C.c_double), ("kinO2anox", C.c_double),
        ("initial_value", C.c_double * NVAR), ("minimum", C.c_double * NVAR),
        ("i_offset", C.c_int32), ("j_offset", C.c_int32),
    ]


class StepInfo(C.Structure):
    """``msed_step_info``"""
    _fields_ = [
        ("steps_done", C.c_int64), ("rhs_evaluations", C.c_int64), ("subcycle_warnings", C.c_int64),
        ("last_min_dt", C.c_double), ("last_min_dt_grid_cell", C.c_int32 * 4),
        ("nan_detected", C.c_int32), ("kernel_ms", C.c_double), ("kernel_launches", C.c_int64),
        ("fused_pairs", C.c_int64), ("fused_ms", C.c_double), ("fused_steps", C.c_int64),
    ]


class PelagicState(C.Structure):
    """``msed_pelagic_state`` -- inputs of pelagic_benthic_coupler (NULL = field absent)."""
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "temperature", "oxygen", "detN", "detN_z_velocity", "detC", "detP", "detP_z_velocity",
        "nitrate", "ammonium", "DIN", "DIP")]


class BenthicPelagicParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dinflux_const", "dipflux_const", "convertN", "NC_fdet", "NC_sdet")]


class PelagicFluxes(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "nitrate", "ammonium", "DIN", "DIP", "detN", "detC", "detP", "oxygen")]


class SoilPelagicParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dinflux_const", "dipflux_const", "convertN", "convertP")]


class SoilPelagicFluxes(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "nitrate", "ammonium", "DIN", "DIP", "oxygen", "odu", "detN", "detC", "detP")]


class PelagicSoilState(C.Structure):
    """``msed_pelagic_soil_state`` -- inputs of pelagic_soil_connector (NULL = field absent)."""
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "temperature", "par", "oxygen", "odu", "detN", "detN_z_velocity", "detC", "detP", "detP_z_velocity",
        "nitrate", "ammonium", "DIN", "DIP", "water_depth", "tke")]


class PelagicSoilParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "sinking_factor", "sinking_factor_min", "NC_ldet", "NC_sdet", "half_sedimentation_depth",
        "half_sedimentation_tke", "critical_detritus", "convertN", "convertP")]


class SpinupMember(C.Structure):
    """``msed_spinup_member`` -- reaction parameters and initial values of one member of a spin-up batch."""
    _fields_ = [(n, C.c_double) for n in (
        "rLabile", "rSemilabile", "NCrLdet", "NCrSdet", "PAds", "PAdsODU", "NH3Ads", "CprodMax", "rnit",
        "ksO2nitri", "rODUox", "ksO2oduox", "ksO2oxic", "ksNO3denit", "kinO2denit", "kinNO3anox",
        "kinO2anox")] + [("initial_value", C.c_double * NVAR)]


COMPAT_P2B_OXYGEN_LAST_CELL, COMPAT_P2S_HEAD = 1, 2

ALLREDUCE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

_dp = C.POINTER(C.c_double)
_h = C.c_void_p

# every symbol include/msed.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "msed_config_defaults": (C.c_int, [C.POINTER(Config)]),
    "msed_create": (C.c_int, [C.POINTER(Config), C.POINTER(_h)]),
    "msed_destroy": (C.c_int, [_h]),
    "msed_last_error": (C.c_char_p, [_h]),
    "msed_version": (C.c_char_p, []),
    "msed_sizeof": (C.c_size_t, [C.c_int]),
    "msed_get_grid": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "msed_set_mask": (C.c_int, [_h, C.POINTER(C.c_int32)]),
    "msed_set_porosity": (C.c_int, [_h, _dp]),
    "msed_update_porosity_from_surface": (C.c_int, [_h, _dp]),
    "msed_set_par_surface": (C.c_int, [_h, _dp]),
    "msed_check_domain": (C.c_int, [_h]),
    "msed_init_concentrations": (C.c_int, [_h]),
    "msed_set_state": (C.c_int, [_h, _dp]),
    "msed_get_state": (C.c_int, [_h, _dp]),
    "msed_set_state_from_column": (C.c_int, [_h, _dp]),
    "msed_set_boundary": (C.c_int, [_h, _dp, _dp]),
    "msed_get_boundary_conditions": (C.c_int, [_h, _dp, C.POINTER(_dp), C.POINTER(_dp)]),
    "msed_get_boundary": (C.c_int, [_h, _dp, _dp]),
    "msed_get_fluxes": (C.c_int, [_h, _dp]),
    "msed_get_upward_fluxes": (C.c_int, [_h, _dp]),
    "msed_get_field": (C.c_int, [_h, C.c_int, _dp]),
    "msed_export_state_begin": (C.c_int, [_h, _dp]),
    "msed_export_state_wait": (C.c_int, [_h]),
    "msed_get_rhs": (C.c_int, [_h, _dp]),
    "msed_ode_solver": (C.c_int, [_h, C.c_double, C.c_int, C.POINTER(StepInfo)]),
    "msed_step": (C.c_int, [_h, C.c_double, C.c_int, C.c_int64, C.POINTER(StepInfo)]),
    "msed_run": (C.c_int, [_h, C.c_double, C.c_int, C.c_double, C.POINTER(StepInfo)]),
    "msed_set_import_generations": (C.c_int, [_h, C.POINTER(C.c_uint64)]),
    "msed_get_exchange_timing": (C.c_int, [_h, C.POINTER(C.c_double)]),
    "msed_run_exchange": (C.c_int, [_h, C.c_double, C.c_int, C.c_double, _dp, C.POINTER(_dp), C.POINTER(_dp),
                                    _dp, C.POINTER(StepInfo)]),
    "msed_set_exchange_chunks": (C.c_int, [_h, C.c_int]),
    "msed_set_exchange_order": (C.c_int, [_h, C.c_int]),
    "msed_set_rk_stages_per_launch": (C.c_int, [_h, C.c_int]),
    "msed_set_step_fusion": (C.c_int, [_h, C.c_int]),
    "msed_spinup_column": (C.c_int, [C.POINTER(Config), _dp, _dp, C.c_int64, C.c_int, _dp,
                                     C.POINTER(StepInfo)]),
    "msed_pelagic_init": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "msed_pelagic_get": (C.c_int, [_h, _dp]),
    "msed_coupled_run": (C.c_int, [_h, C.c_double, C.c_int, C.c_double, C.c_int64, C.POINTER(StepInfo)]),
    "msed_pelagic_benthic_coupler": (C.c_int, [_h, C.POINTER(PelagicState)]),
    "msed_benthic_pelagic_coupler": (C.c_int, [_h, C.POINTER(BenthicPelagicParams), C.POINTER(PelagicFluxes)]),
    "msed_soil_pelagic_connector": (C.c_int, [_h, C.POINTER(SoilPelagicParams), C.POINTER(SoilPelagicFluxes)]),
    "msed_spinup_batch": (C.c_int, [C.POINTER(Config), C.c_int32, C.POINTER(SpinupMember), _dp, _dp, C.c_int64,
                                    C.c_int, _dp, C.POINTER(StepInfo)]),
    "msed_pelagic_soil_params_defaults": (C.c_int, [C.POINTER(PelagicSoilParams)]),
    "msed_pelagic_soil_connector": (C.c_int, [_h, C.POINTER(PelagicSoilState), C.POINTER(PelagicSoilParams)]),
    "msed_set_compat": (C.c_int, [_h, C.c_int]),
    "msed_diagnostics": (C.c_int, [_h, _dp, _dp, C.c_int]),
    "msed_state_checksum": (C.c_int, [_h, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]),
    "msed_measure_fp64_peak": (C.c_int, [C.c_int, _dp]),
    "msed_set_stream": (C.c_int, [_h, C.c_void_p]),
    "msed_synchronize": (C.c_int, [_h]),
    "msed_device_state": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "msed_set_boundary_device": (C.c_int, [_h, C.c_void_p, C.c_void_p]),
    "msed_get_fluxes_device": (C.c_int, [_h, C.c_void_p]),
    "msed_nccl_unique_id": (C.c_int, [C.c_char * 128]),
    "msed_comm_init": (C.c_int, [_h, C.c_char * 128, C.c_int, C.c_int]),
    "msed_comm_destroy": (C.c_int, [_h]),
    "msed_set_allreduce_hook": (C.c_int, [_h, ALLREDUCE_HOOK, C.c_void_p]),
}

LIB_PATH = os.environ.get("MSED_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                    "libmsed_b200.so")
_lib = None


def load(path: str | None = None):
    """Load libmsed_b200.so and bind every declared symbol.  No fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise MsedLibraryError(
            f"{p} not found: build the CUDA extension first (__graft_entry__.build()); "
            "mossco_code_b200 has no CPU path")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:  # pragma: no cover - build error
            raise MsedLibraryError(f"{p} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    if lib.msed_sizeof(0) != C.sizeof(Config) or lib.msed_sizeof(1) != C.sizeof(StepInfo):
        raise MsedLibraryError(
            f"struct layout mismatch: msed_config {lib.msed_sizeof(0)} vs {C.sizeof(Config)}, "
            f"msed_step_info {lib.msed_sizeof(1)} vs {C.sizeof(StepInfo)}")
    if path is None:
        _lib = lib
    return lib
