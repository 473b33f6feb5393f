"""ctypes harness around oracle/libmsed_oracle.so -- TEST INFRASTRUCTURE ONLY.

Mirrors the method names of ``mossco_code_b200.sediment.SedimentDriver`` so parity tests can drive
the CPU restatement and the CUDA product with the same lines.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NVAR = 8
EULER, RUNGE_KUTTA_4, ADAPTIVE_EULER, RUNGE_KUTTA_4_38 = 0, 1, 2, 3
MODEL_OMEXDIA_P, MODEL_NONE = 0, 1

_PTR = dict(conc=0, bdys=1, fluxes=2, porosity=3, intf_porosity=4, bioturbation_factor=5, par=6,
            par_surface=7, temp3d=8, flux_cap=9, biomass=10, weighted_toc=11, denit=12, zi=13, zc=14,
            dz=15, dzc=16, transport=17)


class SedNml(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "diffusivity", "bioturbation", "porosity_max", "porosity_fac", "k_par", "pom_flux_max",
        "bioturbation_depth", "bioturbation_min", "bioturb_k_l", "bioturb_L1", "bioturb_L2",
        "bioturb_beta", "bioturb_b", "bioturb_dry_density")] + [
        ("bioturbation_profile", C.c_int), ("distributed_pom_flux", C.c_int)]


class OmexdiaParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "rLabile", "rSemilabile", "NCrLdet", "NCrSdet", "PAds", "PAdsODU", "NH3Ads", "CprodMax",
        "rnit", "ksO2nitri", "rODUox", "ksO2oduox", "ksO2oxic", "ksNO3denit", "kinO2denit",
        "kinNO3anox", "kinO2anox")] + [("init", C.c_double * NVAR), ("minimum", C.c_double * NVAR)]


def build(fast: bool = False, native: bool = False) -> str:
    """Compile the oracle with the committed Makefile; returns the library path."""
    target = "libmsed_oracle_fast.so" if fast else "libmsed_oracle.so"
    if native:
        subprocess.run(["make", "-C", HERE, "fast-native"], check=True, capture_output=True)
    else:
        subprocess.run(["make", "-C", HERE, target], check=True, capture_output=True)
    return os.path.join(HERE, target)


_libs = {}


def load(fast: bool = False):
    key = "fast" if fast else "parity"
    if key in _libs:
        return _libs[key]
    path = os.path.join(HERE, "libmsed_oracle_fast.so" if fast else "libmsed_oracle.so")
    if not os.path.exists(path):
        build(fast=fast)
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    lib.osed_sed_nml_defaults.argtypes = [C.POINTER(SedNml)]
    lib.osed_omexdia_defaults.argtypes = [C.POINTER(OmexdiaParams)]
    lib.osedpy_create.restype = C.c_void_p
    lib.osedpy_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(SedNml), C.c_int,
                                  C.POINTER(OmexdiaParams), C.POINTER(C.c_int)]
    lib.osedpy_destroy.argtypes = [C.c_void_p]
    lib.osedpy_ptr.restype = dp
    lib.osedpy_ptr.argtypes = [C.c_void_p, C.c_int]
    lib.osedpy_sed.restype = C.c_void_p
    lib.osedpy_sed.argtypes = [C.c_void_p]
    lib.osedpy_set_solver.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.osedpy_get_solver_diag.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int), C.POINTER(C.c_long), dp]
    lib.osedpy_test_solver.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_double, C.c_int, C.c_long]
    lib.osed_update_porosity.argtypes = [C.c_void_p, C.c_int]
    lib.osed_init_concentrations.argtypes = [C.c_void_p]
    lib.osed_check_domain.argtypes = [C.c_void_p]
    lib.osed_check_domain.restype = C.c_int
    lib.osed_get_rhs.argtypes = [C.c_void_p, dp]
    lib.osed_ode_solver.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.osed_component_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.osed_component_step.restype = C.c_int
    lib.osed_get_boundary_conditions.argtypes = [C.c_void_p, dp, C.POINTER(dp), C.POINTER(dp)]
    lib.osed_spinup_column.argtypes = [C.POINTER(SedNml), C.POINTER(OmexdiaParams), C.c_int, C.c_double,
                                       C.c_double, C.c_double, dp, dp, C.c_long, C.c_int, dp]
    lib.osed_omexdia_p_cell.argtypes = [C.POINTER(OmexdiaParams), dp, C.c_double, dp, dp]
    lib.osed_diff3d.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp]
    lib.osed_bench_tiled.restype = C.c_double
    lib.osed_bench_tiled.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(SedNml),
                                     C.POINTER(OmexdiaParams), C.POINTER(C.c_int), dp, dp, dp,
                                     C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.POINTER(C.c_long)]
    lib.osed_bench_fused.restype = C.c_double
    lib.osed_bench_fused.argtypes = lib.osed_bench_tiled.argtypes
    lib.osed_pelagic_benthic_coupler.argtypes = [C.c_size_t, C.POINTER(dp), dp, dp]
    lib.osed_pelagic_benthic_coupler_ex.argtypes = [C.c_size_t, C.POINTER(dp), dp, dp, C.c_int]
    lib.osed_pelagic_soil_connector.argtypes = [C.c_size_t, C.POINTER(dp), dp, C.c_int, dp, dp]
    lib.osed_pelagic_soil_connector.restype = None
    lib.osed_soil_pelagic_connector.argtypes = [C.c_size_t, dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_int, C.c_int, dp]
    lib.osed_soil_pelagic_connector.restype = None
    lib.osed_benthic_pelagic_coupler.argtypes = [C.c_size_t, dp, C.c_double, C.c_double, C.c_double,
                                                 C.c_double, C.c_double, dp]
    _libs[key] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def sed_nml(**kw) -> SedNml:
    nml = SedNml()
    load().osed_sed_nml_defaults(C.byref(nml))
    for k, v in kw.items():
        if not hasattr(nml, k):
            raise AttributeError(k)
        setattr(nml, k, v)
    return nml


def omexdia_params(**kw) -> OmexdiaParams:
    p = OmexdiaParams()
    load().osed_omexdia_defaults(C.byref(p))
    for k, v in kw.items():
        if k in ("init", "minimum"):
            for n in range(NVAR):
                getattr(p, k)[n] = float(v[n])
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


def from_config(cfg):
    """Translate a product ``msed_config`` (ctypes struct) into the oracle's namelist structs."""
    nml = sed_nml(**{k: getattr(cfg, k) for k in (
        "diffusivity", "bioturbation", "porosity_max", "porosity_fac", "k_par", "pom_flux_max",
        "bioturbation_depth", "bioturbation_min", "bioturb_k_l", "bioturb_L1", "bioturb_L2",
        "bioturb_beta", "bioturb_b", "bioturb_dry_density", "bioturbation_profile",
        "distributed_pom_flux")})
    par = omexdia_params(**{k: getattr(cfg, k) for k in (
        "rLabile", "rSemilabile", "NCrLdet", "NCrSdet", "PAds", "PAdsODU", "NH3Ads", "CprodMax",
        "rnit", "ksO2nitri", "rODUox", "ksO2oduox", "ksO2oxic", "ksNO3denit", "kinO2denit",
        "kinNO3anox", "kinO2anox")},
        init=[cfg.initial_value[n] for n in range(NVAR)],
        minimum=[cfg.minimum[n] for n in range(NVAR)])
    return nml, par


def default_config(**kw):
    """A plain-Python stand-in for the product's ``msed_config`` holding the reference defaults
    (fabm_sediment_driver.F90:217-231, component :59-67, fabm_sed.nml:51-77), taken from THIS library --
    for callers that must not load the product (bench.py --impl reference)."""
    from types import SimpleNamespace
    nml, par = sed_nml(), omexdia_params()
    d = {n: getattr(nml, n) for n, _ in SedNml._fields_}
    d.update({n: getattr(par, n) for n, _ in OmexdiaParams._fields_ if n not in ("init", "minimum")})
    d.update(inum=1, jnum=1, knum=10, dzmin=0.005, model=MODEL_OMEXDIA_P, dt_min=1.0e-8,
             relative_change_min=-0.9, bcup_dissolved_variables=2, adaptive_solver_diagnostics=0,
             initial_value=[par.init[n] for n in range(NVAR)], minimum=[par.minimum[n] for n in range(NVAR)])
    for k, v in kw.items():
        if k not in d:
            raise AttributeError(k)
        d[k] = v
    return SimpleNamespace(**d)


class OracleSediment:
    """``type_sed`` on the CPU (restated reference).  Arrays are live numpy views, Fortran order."""

    def __init__(self, inum, jnum, knum, dzmin, nml=None, params=None, mask2d=None,
                 model=MODEL_OMEXDIA_P, dt_min=1.0e-8, relative_change_min=-0.9,
                 bcup_dissolved_variables=2, adaptive_solver_diagnostics=False, verbose=False,
                 fast=False):
        self._lib = load(fast)
        self.inum, self.jnum, self.knum, self.nvar = inum, jnum, knum, NVAR
        self._nml = nml or sed_nml()
        self._par = params or omexdia_params()
        m = None
        if mask2d is not None:
            m = np.asfortranarray(np.asarray(mask2d, dtype=np.int32))
            assert m.shape == (inum, jnum)
        self._h = self._lib.osedpy_create(inum, jnum, knum, dzmin, C.byref(self._nml), model,
                                          C.byref(self._par),
                                          None if m is None else m.ctypes.data_as(C.POINTER(C.c_int)))
        if not self._h:
            raise RuntimeError("osedpy_create failed")
        self._sed = self._lib.osedpy_sed(self._h)
        self._lib.osedpy_set_solver(self._h, dt_min, relative_change_min, bcup_dissolved_variables,
                                    int(adaptive_solver_diagnostics), int(verbose))

    @classmethod
    def from_config(cls, cfg, mask2d=None, fast=False, verbose=False):
        nml, par = from_config(cfg)
        return cls(cfg.inum, cfg.jnum, cfg.knum, cfg.dzmin, nml, par, mask2d, model=cfg.model,
                   dt_min=cfg.dt_min, relative_change_min=cfg.relative_change_min,
                   bcup_dissolved_variables=cfg.bcup_dissolved_variables,
                   adaptive_solver_diagnostics=bool(cfg.adaptive_solver_diagnostics), fast=fast,
                   verbose=verbose)

    def finalize(self):
        if self._h:
            self._lib.osedpy_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.finalize()
        except Exception:
            pass

    def _view(self, name, shape):
        ptr = self._lib.osedpy_ptr(self._h, _PTR[name])
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape, order="F")

    # live views ---------------------------------------------------------------------------
    @property
    def conc(self):
        return self._view("conc", (self.inum, self.jnum, self.knum, self.nvar))

    @conc.setter
    def conc(self, v):
        self.conc[...] = v

    @property
    def bdys(self):
        return self._view("bdys", (self.inum, self.jnum, self.nvar + 1))

    @property
    def fluxes(self):
        return self._view("fluxes", (self.inum, self.jnum, self.nvar))

    def field3d(self, name):
        k = {"zi": self.knum + 1, "dzc": self.knum - 1}.get(name, self.knum)
        return self._view(name, (self.inum, self.jnum, k))

    @property
    def par_surface(self):
        return self._view("par_surface", (self.inum, self.jnum))

    def set_boundary(self, bdys=None, fluxes=None):
        if bdys is not None:
            self.bdys[...] = bdys
        if fluxes is not None:
            self.fluxes[...] = fluxes

    # driver procedures ----------------------------------------------------------------------
    def update_porosity(self, porosity_surface=None, from_surface=True):
        if from_surface:
            self.field3d("porosity")[:, :, 0] = porosity_surface
        self._lib.osed_update_porosity(self._sed, int(from_surface))

    def init_concentrations(self):
        self._lib.osed_init_concentrations(self._sed)

    def check_domain(self):
        return self._lib.osed_check_domain(self._sed)

    def get_rhs(self):
        out = np.zeros((self.inum, self.jnum, self.knum, self.nvar), order="F")
        self._lib.osed_get_rhs(self._sed, _p(out))
        return out

    def ode_solver(self, dt, method=ADAPTIVE_EULER):
        self._lib.osed_ode_solver(self._sed, float(dt), int(method))

    def step(self, dt, method=ADAPTIVE_EULER, nsteps=1):
        for _ in range(int(nsteps)):
            if self._lib.osed_component_step(self._sed, float(dt), int(method)):
                return 1
        return 0

    def get_boundary_conditions(self, temperature=None, csurf=None, wz=None):
        dp = C.POINTER(C.c_double)
        keep = []
        cs, ws = (dp * NVAR)(), (dp * NVAR)()
        t = None
        if temperature is not None:
            t = np.asfortranarray(np.asarray(temperature, dtype=np.float64))
        for n in range(NVAR):
            for arr, src in ((cs, csurf), (ws, wz)):
                if src is not None and src[n] is not None:
                    a = np.asfortranarray(np.asarray(src[n], dtype=np.float64))
                    keep.append(a)
                    arr[n] = _p(a)
        self._lib.osed_get_boundary_conditions(self._sed, _p(t), cs, ws)

    def solver_diag(self):
        lm, cell, sub, bio = C.c_double(), (C.c_int * 4)(), C.c_long(), C.c_double()
        self._lib.osedpy_get_solver_diag(self._h, C.byref(lm), cell, C.byref(sub), C.byref(bio))
        return dict(last_min_dt=lm.value, last_min_dt_grid_cell=list(cell), subcycles=sub.value,
                    bioturbation=bio.value)


def test_solver_kat(inum, jnum, knum, nvar, conc, dt, method, nsteps):
    """ode_solver driven by the src/test/test_Solver.F90:40 right-hand side; conc updated in place."""
    a = np.asfortranarray(conc, dtype=np.float64)
    load().osedpy_test_solver(inum, jnum, knum, nvar, _p(a), float(dt), int(method), int(nsteps))
    return a


test_solver_kat.__test__ = False  # not a pytest test


def omexdia_cell(c8, temp, params=None):
    par = params or omexdia_params()
    c = np.ascontiguousarray(c8, dtype=np.float64)
    r = np.zeros(8)
    d = C.c_double()
    load().osed_omexdia_p_cell(C.byref(par), _p(c), float(temp), _p(r), C.byref(d))
    return r, d.value


def spinup_column(nml, par, knum, dzmin, dt_min, rcm, bdys1d, fluxes1d, nsteps, method=ADAPTIVE_EULER):
    b = np.ascontiguousarray(bdys1d, dtype=np.float64).reshape(-1)
    f = np.ascontiguousarray(fluxes1d, dtype=np.float64).reshape(-1)
    out = np.zeros((1, 1, knum, NVAR), order="F")
    load().osed_spinup_column(C.byref(nml), C.byref(par), knum, dzmin, dt_min, rcm, _p(b), _p(f),
                              int(nsteps), int(method), _p(out))
    return out


_native_built = False


def bench_tiled(cfg, mask2d, conc, bdys, fluxes, dt, method, nsteps, nthreads, native=True, fused=False):
    """Times the restated reference CPU path (j-slab tiles, one OpenMP thread per tile).  ``fused``: the
    fused-loop variant of the same step (one pass over the state per attempt)."""
    global _native_built
    if native and not _native_built:
        try:
            build(fast=True, native=True)
        except Exception:
            build(fast=True)
        _native_built = True
        _libs.pop("fast", None)
    lib = load(fast=True)
    nml, par = from_config(cfg)
    m = None if mask2d is None else np.asfortranarray(np.asarray(mask2d, dtype=np.int32))
    sub = C.c_long()
    conc = np.asfortranarray(conc, dtype=np.float64)
    b = np.asfortranarray(bdys, dtype=np.float64)
    f = np.asfortranarray(fluxes, dtype=np.float64)
    fn = lib.osed_bench_fused if fused else lib.osed_bench_tiled
    secs = fn(cfg.inum, cfg.jnum, cfg.knum, cfg.dzmin, C.byref(nml), C.byref(par),
                                None if m is None else m.ctypes.data_as(C.POINTER(C.c_int)),
                                _p(conc), _p(b), _p(f), float(dt), int(method), int(nsteps),
                                cfg.dt_min, cfg.relative_change_min, cfg.bcup_dissolved_variables,
                                int(nthreads), C.byref(sub))
    return secs, sub.value, conc


P2B_FIELDS = ("oxygen", "detN", "detN_z_velocity", "detC", "detP", "detP_z_velocity", "nitrate", "ammonium",
              "DIN", "DIP")
B2P_FIELDS = ("nitrate", "ammonium", "DIN", "DIP", "detN", "detC", "detP", "oxygen")


P2S_FIELDS = ("oxygen", "odu", "detN", "detN_z_velocity", "detC", "detP", "detP_z_velocity", "nitrate", "ammonium",
              "DIN", "DIP", "water_depth", "tke")
P2S_PARAMS = ("sinking_factor", "sinking_factor_min", "NC_ldet", "NC_sdet", "half_sedimentation_depth",
              "half_sedimentation_tke", "critical_detritus", "convertN", "convertP")
# module defaults, pelagic_soil_connector.F90:38-46 (three of them are default-real literals)
P2S_DEFAULTS = dict(sinking_factor=0.3, sinking_factor_min=float(np.float32(0.02)), NC_ldet=0.23, NC_sdet=0.01,
                    half_sedimentation_depth=float(np.float32(0.1)), half_sedimentation_tke=1.0e3,
                    critical_detritus=60.0, convertN=1.0, convertP=1.0)


def pelagic_soil_connector(shape2d, params=None, head_compat=False, csurf0=None, wz0=None, **fields):
    """Restated pelagic_soil_connector Run: returns (csurf list of 8, wz list of 3 + 5 None).  ``csurf0`` /
    ``wz0``: previous contents of the export fields (rows the connector leaves alone keep them)."""
    dp = C.POINTER(C.c_double)
    n2 = int(np.prod(shape2d))
    arr, keep = (dp * 13)(), []
    for i, name in enumerate(P2S_FIELDS):
        a = fields.get(name)
        if a is not None:
            a = np.asfortranarray(np.asarray(a, dtype=np.float64))
            keep.append(a)
            arr[i] = _p(a)
    par = dict(P2S_DEFAULTS)
    par.update(params or {})
    pv = np.array([par[k] for k in P2S_PARAMS], dtype=np.float64)
    cs = np.zeros(tuple(shape2d) + (8,), order="F") if csurf0 is None else np.asfortranarray(csurf0, dtype=np.float64).copy(order="F")
    wz = np.zeros(tuple(shape2d) + (3,), order="F") if wz0 is None else np.asfortranarray(wz0, dtype=np.float64).copy(order="F")
    load().osed_pelagic_soil_connector(n2, arr, _p(pv), int(head_compat), _p(cs), _p(wz))
    return [cs[..., n] for n in range(8)], [wz[..., n] for n in range(3)] + [None] * 5


def pelagic_benthic_coupler(shape2d, oxy_last_cell=False, **fields):
    """Restated pelagic_benthic_coupler Run: returns (csurf list of 8, wz list of 3 + 5 None)."""
    dp = C.POINTER(C.c_double)
    n2 = int(np.prod(shape2d))
    arr, keep = (dp * 10)(), []
    for i, name in enumerate(P2B_FIELDS):
        a = fields.get(name)
        if a is not None:
            a = np.asfortranarray(np.asarray(a, dtype=np.float64))
            keep.append(a)
            arr[i] = _p(a)
    cs = np.zeros(tuple(shape2d) + (8,), order="F")
    wz = np.zeros(tuple(shape2d) + (3,), order="F")
    load().osed_pelagic_benthic_coupler_ex(n2, arr, _p(cs), _p(wz), int(oxy_last_cell))
    return [cs[..., n] for n in range(8)], [wz[..., n] for n in range(3)] + [None] * 5


def benthic_pelagic_coupler(up, dinflux_const=0.0, dipflux_const=-1.0, convertN=1.0, NC_fdet=0.20,
                            NC_sdet=0.04):
    up = np.asfortranarray(np.asarray(up, dtype=np.float64))
    n2 = int(np.prod(up.shape[:-1]))
    out = np.zeros(up.shape[:-1] + (8,), order="F")
    load().osed_benthic_pelagic_coupler(n2, _p(up), dinflux_const, dipflux_const, convertN, NC_fdet, NC_sdet,
                                        _p(out))
    return {name: out[..., i] for i, name in enumerate(B2P_FIELDS)}


S2P_FIELDS = ("nitrate", "ammonium", "DIN", "DIP", "oxygen", "odu", "detN", "detC", "detP")


def soil_pelagic_connector(up, want=S2P_FIELDS, dinflux_const=0.0, dipflux_const=-1.0, convertN=1.0,
                           convertP=1.0):
    """Restated soil_pelagic_connector Run; ``want`` = the fields the export state holds."""
    up = np.asfortranarray(np.asarray(up, dtype=np.float64))
    n2 = int(np.prod(up.shape[:-1]))
    out = np.zeros(up.shape[:-1] + (9,), order="F")
    load().osed_soil_pelagic_connector(n2, _p(up), dinflux_const, dipflux_const, convertN, convertP,
                                       int("oxygen" in want), int("odu" in want), _p(out))
    return {name: out[..., i] for i, name in enumerate(S2P_FIELDS) if name in want}
