"""Deterministic synthetic forcing for the five BASELINE.json configurations.

Shared by the tests, ``__graft_entry__.smoke()`` and ``bench.py``; pure numpy, no reference or
oracle dependency.  State order: ldetC sdetC detP po4 no3 nh3 oxy odu.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

NVAR = 8
# examples/esmf/sediment/run_sed.nml: pel_Temp, pel_PO4, pel_NO3, pel_NH4, pel_O2 and the
# spin-up quirk that feeds pel_O2 into dissolved_reduced_substances (component :595)
C1_BDYS = np.array([5.0, 0.0, 0.0, 0.0, 0.6, 14.0, 4.0, 250.0, 250.0])
# pflux_lDetC, pflux_sDetC, pflux_lDetP in mmol m-2 d-1 -> s-1 (component :596-600)
C1_FLUXES = np.array([2.0 / 86400.0, 24.0 / 86400.0, 0.08 / 86400.0, 0.0, 0.0, 0.0, 0.0, 0.0])
# examples/esmf/sediment/default.dat (the ESMF variant of the same example)
C1B_BDYS = np.array([10.0, 0.0, 0.0, 0.0, 1.0, 5.0, 5.0, 250.0, 0.0])
C1B_FLUXES = np.array([5.787e-5, 5.787e-5, 9.26e-7, 0.0, 0.0, 0.0, 0.0, 0.0])


def _smooth_field(rng, inum, jnum, coarse=9):
    """Seeded low-pass noise in [0,1]: bilinear interpolation of a coarse random lattice."""
    g = rng.random((coarse, coarse))
    xi = np.linspace(0.0, coarse - 1.0, inum)
    xj = np.linspace(0.0, coarse - 1.0, jnum)
    i0 = np.minimum(xi.astype(int), coarse - 2)
    j0 = np.minimum(xj.astype(int), coarse - 2)
    fi = (xi - i0)[:, None]
    fj = (xj - j0)[None, :]
    a = g[i0][:, j0]
    b = g[i0 + 1][:, j0]
    c = g[i0][:, j0 + 1]
    d = g[i0 + 1][:, j0 + 1]
    return (a * (1 - fi) + b * fi) * (1 - fj) + (c * (1 - fi) + d * fi) * fj


def make_case(name, inum, jnum, knum, dzmin, seed=1234, perturb=0.1, land_fraction=0.0,
              smooth_temperature=False, par_max=0.0, base_bdys=C1_BDYS, base_fluxes=C1_FLUXES):
    """Forcing = base values x (1 + perturb*U(-1,1)) per column and field (seeded)."""
    rng = np.random.default_rng(seed)
    shape2 = (inum, jnum)
    bdys = np.empty(shape2 + (NVAR + 1,), order="F")
    fluxes = np.empty(shape2 + (NVAR,), order="F")
    for n in range(NVAR + 1):
        u = rng.uniform(-1.0, 1.0, size=shape2) if perturb else 0.0
        bdys[:, :, n] = base_bdys[n] * (1.0 + perturb * u)
    for n in range(NVAR):
        u = rng.uniform(-1.0, 1.0, size=shape2) if perturb else 0.0
        fluxes[:, :, n] = base_fluxes[n] * (1.0 + perturb * u)
    if smooth_temperature:  # 2..18 degC smooth field (C3)
        bdys[:, :, 0] = 2.0 + 16.0 * _smooth_field(rng, inum, jnum)
    par = np.zeros(shape2, order="F")
    if par_max > 0.0:
        par[...] = par_max * _smooth_field(rng, inum, jnum)
    mask = np.zeros(shape2, dtype=np.int32, order="F")
    if land_fraction > 0.0:
        f = _smooth_field(np.random.default_rng(seed + 790), inum, jnum)
        thr = np.quantile(f, land_fraction)
        mask[...] = (f < thr).astype(np.int32)
        if mask.all():
            mask[0, 0] = 0
    return SimpleNamespace(name=name, inum=inum, jnum=jnum, knum=knum, dzmin=dzmin, mask=mask,
                           bdys=bdys, fluxes=fluxes, par_surface=par)


def config_case(which: str, scale: float = 1.0):
    """The BASELINE.json configs, optionally shrunk (scale<1) for CPU-sized parity runs."""
    s = lambda n: max(2, int(round(n * scale)))  # noqa: E731
    if which == "C1":
        return make_case("C1", 1, 1, 30, 0.002, perturb=0.0)
    if which == "C1b":
        return make_case("C1b", 1, 1, 30, 0.002, perturb=0.0, base_bdys=C1B_BDYS, base_fluxes=C1B_FLUXES)
    if which == "C2":
        return make_case("C2", s(100), s(100), 30, 0.002, seed=1234)
    if which == "C3":
        return make_case("C3", s(1000), s(1000), 30, 0.002, seed=2024, land_fraction=0.45,
                         smooth_temperature=True, par_max=50.0)
    if which == "C4":
        return make_case("C4", s(4096), s(4096), 40, 0.0015, seed=4096)
    if which == "C5":
        return make_case("C5", s(2048), s(2048), 30, 0.002, seed=2048)
    raise KeyError(which)


def rel_err(got, ref, wet=None):
    """max |got-ref| / max(|ref|, tiny) over unmasked entries."""
    g, r = np.asarray(got), np.asarray(ref)
    if wet is not None:
        g, r = g[wet], r[wet]
    return float(np.max(np.abs(g - r) / np.maximum(np.abs(r), 1e-300))) if g.size else 0.0


def scaled_err(got, ref, wet=None):
    """max |got-ref| / max|ref| per variable (last axis): robust where single entries cross zero."""
    g, r = np.asarray(got), np.asarray(ref)
    if wet is not None:
        g, r = g[wet], r[wet]
    if g.size == 0:
        return 0.0
    ax = tuple(range(g.ndim - 1))
    scale = np.maximum(np.max(np.abs(r), axis=ax), 1e-300)
    return float(np.max(np.max(np.abs(g - r), axis=ax) / scale))
