"""Host-side logic of the component and driver mirrors that needs no GPU: the generation counters the component
hands to msed_set_import_generations (static import fields), and the pointer-table cache of run_exchange -- the
native library is replaced by a recorder."""
import ctypes as C

import numpy as np

from mossco_code_b200.component import FabmSedimentComponent
from mossco_code_b200.sediment import NVAR, PARTICULATE, VARIABLE_NAMES, SedimentDriver

SH = (5, 4)


class FakeSed:
    """Stands in for SedimentDriver under FabmSedimentComponent.run."""
    shape2d = SH

    def __init__(self):
        self.gens, self.runs, self.info = [], 0, None

    def set_import_generations(self, gen):
        self.gens.append(None if gen is None else list(gen))

    def run_exchange(self, dt, method, seconds, temp, cs, wz, out=None):
        self.runs += 1
        return 0, np.zeros(SH + (NVAR,), order="F")


def _imports():
    imp = {"temperature_at_soil_surface": np.full(SH, 8.0, order="F")}
    for n, v in enumerate(VARIABLE_NAMES):
        imp[f"{v}_at_soil_surface"] = np.ones(SH, order="F")
        if PARTICULATE[n]:
            imp[f"{v}_z_velocity_at_soil_surface"] = np.full(SH, -1e-5, order="F")
    return imp


def test_static_import_fields_keep_their_generation():
    comp = FabmSedimentComponent()
    comp.sed = FakeSed()
    comp.export_3d_every_run = False
    imp, exp = _imports(), {}
    comp.run(imp, exp, run_seconds=3600.0)
    assert comp.sed.gens == []                                   # the reference's behaviour: nothing declared
    comp.static_import_suffixes = ("_z_velocity_at_soil_surface",)
    for _ in range(3):
        comp.run(imp, exp, run_seconds=3600.0)
    g = comp.sed.gens
    assert len(g) == 3 and all(len(x) == 1 + 2 * NVAR for x in g)
    wz_keys = [2 + 2 * n for n in range(NVAR)]
    other = [k for k in range(1 + 2 * NVAR) if k not in wz_keys]
    assert all(g[0][k] == g[1][k] == g[2][k] for k in wz_keys)  # static: same counter every Run
    assert all(g[0][k] < g[1][k] < g[2][k] for k in other)      # everything else: a new one every Run
    comp.import_changed("_z_velocity_at_soil_surface")           # the coupler rewrote them
    comp.run(imp, exp, run_seconds=3600.0)
    assert all(comp.sed.gens[3][k] > g[2][k] for k in wz_keys)


class FakeLib:
    def __init__(self):
        self.calls = []

    def msed_run_exchange(self, h, dt, method, seconds, tp, cs, ws, outp, info):
        self.calls.append((C.cast(tp, C.c_void_p).value, [C.cast(cs[n], C.c_void_p).value for n in range(NVAR)],
                           [C.cast(ws[n], C.c_void_p).value for n in range(NVAR)], C.cast(outp, C.c_void_p).value))
        return 0


def _driver():
    sed = SedimentDriver.__new__(SedimentDriver)        # no library, no handle: only run_exchange's host side
    sed._lib, sed._h, sed.info = FakeLib(), None, C.c_int()
    sed.inum, sed.jnum, sed.knum, sed.nvar = SH[0], SH[1], 3, NVAR
    sed._check = lambda rc, allow=(): rc
    return sed


def test_run_exchange_pointer_tables_are_cached_per_set_of_arrays():
    sed = _driver()
    assert sed.shape2d == SH
    imp = _imports()
    temp = imp["temperature_at_soil_surface"]
    cs = [imp[f"{v}_at_soil_surface"] for v in VARIABLE_NAMES]
    wz = [imp.get(f"{v}_z_velocity_at_soil_surface") for v in VARIABLE_NAMES]
    out = np.zeros(SH + (NVAR,), order="F")
    for _ in range(3):                                   # fresh list objects, same arrays: what a coupler does
        rc, up = sed.run_exchange(360.0, 2, 3600.0, temp, list(cs), list(wz), out=out)
        assert rc == 0 and up is out
    a, b, c = sed._lib.calls
    assert a == b == c
    assert a[0] == temp.ctypes.data and a[3] == out.ctypes.data
    assert a[1] == [x.ctypes.data for x in cs]
    assert a[2] == [None if x is None else x.ctypes.data for x in wz]
    # another array for one field: new tables
    cs2 = list(cs); cs2[6] = np.full(SH, 2.0, order="F")
    sed.run_exchange(360.0, 2, 3600.0, temp, cs2, wz, out=out)
    assert sed._lib.calls[-1][1][6] == cs2[6].ctypes.data != a[1][6]
    # a field that has to be converted (C order, float32) is copied on every call and never cached
    cs3 = list(cs); cs3[4] = np.ones(SH, dtype=np.float32)
    sed.run_exchange(360.0, 2, 3600.0, temp, cs3, wz, out=out)
    assert sed._rx_cache is None
    # without a caller-owned output buffer a fresh one is returned each time
    _, u1 = sed.run_exchange(360.0, 2, 3600.0, temp, cs, wz)
    _, u2 = sed.run_exchange(360.0, 2, 3600.0, temp, cs, wz)
    assert u1 is not u2
