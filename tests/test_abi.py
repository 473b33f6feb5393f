"""The C-ABI library loads and exports every symbol include/msed.h declares; struct layouts match the
ctypes mirror; and without a GPU every compute entry fails loudly (there is no CPU path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "msed.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(msed_[a-z0-9_]+)\s*\(", text))
    names.discard("msed_allreduce_hook")
    return names


def test_header_symbols_are_exported(msed_lib):
    from mossco_code_b200 import _abi
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/msed.h"
    assert declared == set(_abi.SYMBOLS), (declared ^ set(_abi.SYMBOLS))
    for name in declared:
        assert hasattr(msed_lib, name), f"{name} not exported by libmsed_b200.so"


def test_struct_layouts(msed_lib):
    from mossco_code_b200 import _abi
    assert msed_lib.msed_sizeof(0) == C.sizeof(_abi.Config)
    assert msed_lib.msed_sizeof(1) == C.sizeof(_abi.StepInfo)
    assert msed_lib.msed_version().decode().startswith("msed_b200 abi1 sm_100a")


def test_defaults_match_reference_namelists(msed_lib):
    from mossco_code_b200 import default_config
    c = default_config()
    # fabm_sediment_driver.F90:217-231
    assert (c.diffusivity, c.bioturbation, c.bioturbation_depth, c.bioturbation_min) == (0.9, 0.9, 5.0, 0.2)
    assert (c.porosity_max, c.porosity_fac, c.k_par, c.pom_flux_max) == (0.7, 0.9, 2.0e-3, 2.0e4)
    assert c.bioturbation_profile == 1 and c.distributed_pom_flux == 0
    # fabm_sediment_component.F90:59-67
    assert (c.dt_min, c.relative_change_min, c.bcup_dissolved_variables) == (1.0e-8, -0.9, 2)
    # examples/standalone/omexdia_p/fabm_sed.nml:51-77
    assert (c.rLabile, c.rSemilabile, c.NCrLdet, c.NCrSdet) == (0.043, 0.001, 0.22, 0.005)
    assert (c.rnit, c.ksO2nitri, c.rODUox, c.kinO2denit, c.CprodMax) == (200., 20., 20., 70., 9600.)
    assert list(c.initial_value) == [4e3, 4e3, 40., 10., 20., 40., 100., 100.]


def test_bad_arguments_are_rejected(msed_lib):
    from mossco_code_b200 import _abi, default_config
    h = C.c_void_p()
    for kw in (dict(knum=1), dict(knum=_abi.MAX_LAYERS + 1), dict(inum=0), dict(model=7), dict(abi_version=99)):
        cfg = default_config(**kw)
        assert msed_lib.msed_create(C.byref(cfg), C.byref(h)) == _abi.ERR_ARG
        assert not h.value
    # the minimum clip compares bit patterns (clip_min): a negative floor or a NaN is refused up front
    for bad in (-1.0e-3, float("nan")):
        cfg = default_config(minimum=[0.0, 0.0, bad, 0.0, 0.0, 0.0, 0.0, 0.0])
        assert msed_lib.msed_create(C.byref(cfg), C.byref(h)) == _abi.ERR_ARG
        assert b"minimum" in msed_lib.msed_last_error(None)
    assert msed_lib.msed_destroy(None) == 0


def test_no_cpu_fallback(msed_lib):
    """On a box without a CUDA device msed_create must fail with MSED_ERR_CUDA, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the no-GPU behaviour is checked on the CPU box")
    from mossco_code_b200 import MsedError, SedimentDriver, _abi, default_config
    with pytest.raises(MsedError) as e:
        SedimentDriver(default_config(inum=2, jnum=2, knum=10))
    assert e.value.code == _abi.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_missing_library_fails_loudly(tmp_path):
    from mossco_code_b200 import _abi
    with pytest.raises(_abi.MsedLibraryError):
        _abi.load(str(tmp_path / "libmsed_b200.so"))


def test_product_does_not_touch_oracle():
    """Nothing under mossco_code_b200/ or include/ may reference oracle/ (parity claims depend on it)."""
    bad = []
    for base in ("mossco_code_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".F90", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"msed_oracle|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
