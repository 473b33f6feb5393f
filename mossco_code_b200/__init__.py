"""mossco_code_b200 -- B200-native fabm_sediment column solver (MOSSCO hot path only).

The package holds the CUDA kernels + C ABI (``csrc/`` -> ``libmsed_b200.so``), the host-side
mirror of the reference's ``type_sed`` / ``ode_solver`` contract (``sediment``), the ESMF-style
component wrapper (``component``) and the j-slab sharding helpers (``sharding``).
"""
from .sediment import (ADAPTIVE_EULER, EULER, MODEL_NONE, MODEL_OMEXDIA_P, MODEL_TEST_SOLVER,  # noqa: F401
                       PARTICULATE, RUNGE_KUTTA_4, RUNGE_KUTTA_4_38, STATE_NAMES, VARIABLE_NAMES,
                       SedimentDriver, default_config, measure_fp64_peak, nccl_unique_id, ode_solver,
                       spinup_batch, spinup_column)
from ._abi import MsedError, MsedLibraryError, StepInfo  # noqa: F401

__version__ = "0.1.0"
