"""Parity at BASELINE.json's full sizes through size-independent properties: columns are independent,
so a full-size grid whose forcing repeats a small tile must reproduce that tile's (oracle-checked)
result in every column, bit for bit; and transport alone conserves mass column by column."""
import numpy as np
import pytest

from tests.cases import make_case, rel_err

pytestmark = pytest.mark.gpu
DT = 360.0


def _tiled(tile, reps_i, reps_j):
    t = lambda a: np.asfortranarray(np.tile(a, (reps_i, reps_j) + (1,) * (a.ndim - 2)))  # noqa: E731
    return t(tile.bdys), t(tile.fluxes), t(tile.mask), t(tile.par_surface)


def _run(cfgkw, bdys, fluxes, mask, par, method, nsteps):
    from mossco_code_b200 import SedimentDriver, default_config
    cfg = default_config(dt_min=1.0, **cfgkw)
    with SedimentDriver(cfg) as sed:
        sed.set_mask(mask)
        sed.set_par_surface(par)
        sed.init_concentrations()
        sed.set_boundary(bdys, fluxes)
        assert sed.step(DT, method, nsteps) == 0
        return sed.conc, sed.fluxes, sed.info.subcycle_warnings


@pytest.mark.parametrize("name,ti,tj,ri,rj,K,dzmin,land", [
    ("C3 1000x1000x30 with land mask", 20, 25, 50, 40, 30, 0.002, 0.45),
    ("C4 slab 4096x512x40", 64, 8, 64, 64, 40, 0.0015, 0.0),
])
def test_full_size_equals_replicated_tile(gpu, oracle, name, ti, tj, ri, rj, K, dzmin, land):
    tile = make_case("tile", ti, tj, K, dzmin, seed=99, land_fraction=land, smooth_temperature=land > 0,
                     par_max=50.0 if land > 0 else 0.0)
    small, small_f, sub_s = _run(dict(inum=ti, jnum=tj, knum=K, dzmin=dzmin), tile.bdys, tile.fluxes,
                                 tile.mask, tile.par_surface, 2, 3)
    # the tile itself is oracle-checked
    ref = oracle.OracleSediment(ti, tj, K, dzmin, mask2d=tile.mask, dt_min=1.0)
    ref.init_concentrations(); ref.set_boundary(tile.bdys, tile.fluxes)
    assert ref.step(DT, 2, 3) == 0
    wet = tile.mask == 0
    assert rel_err(small[wet], ref.conc[wet]) <= 1e-12
    # the full-size grid
    b, f, m, p = _tiled(tile, ri, rj)
    big, big_f, sub_b = _run(dict(inum=ti * ri, jnum=tj * rj, knum=K, dzmin=dzmin), b, f, m, p, 2, 3)
    assert sub_b == sub_s
    assert big.shape == (ti * ri, tj * rj, K, 8)
    # np.tile: i = rep*ti + ii with ii fastest -> Fortran-split the i axis as (ti, ri), j as (tj, rj)
    view = big.reshape((ti, ri, tj, rj, K, 8), order="F")
    assert np.array_equal(view, np.broadcast_to(small[:, None, :, None], view.shape))
    fview = big_f.reshape((ti, ri, tj, rj, 8), order="F")
    assert np.array_equal(fview, np.broadcast_to(small_f[:, None, :, None], fview.shape))


def test_transport_conserves_mass_full_c3(gpu):
    """MODEL_NONE on the full C3 grid: d/dt of every column inventory equals the bed flux
    (BcDown=3 closes the bottom, diff3d :813)."""
    from mossco_code_b200 import MODEL_NONE, SedimentDriver, default_config
    case = make_case("C3", 1000, 1000, 30, 0.002, seed=2024, land_fraction=0.45, smooth_temperature=True)
    cfg = default_config(inum=1000, jnum=1000, knum=30, dzmin=0.002, model=MODEL_NONE)
    with SedimentDriver(cfg) as sed:
        sed.set_mask(case.mask)
        sed.init_concentrations()
        rng = np.random.default_rng(1)
        surf = 0.5 + 0.3 * rng.random((1000, 1000))
        sed.update_porosity(surf)
        sed.init_concentrations()
        sed.set_boundary(case.bdys, case.fluxes)
        rhs = sed.get_rhs()
        por = sed.field("porosity")
        top = sed.fluxes
        _, _, dz, _ = sed.grid()
    wet = case.mask == 0
    inv = np.einsum("ijkn,ijk,k->ijn", rhs, por, dz)
    scale = np.max(np.abs(top[wet]), axis=0) + 1e-300
    assert np.max(np.abs(inv[wet] - top[wet]) / scale) < 1e-10
    assert np.all(rhs[~wet] == 0.0)
