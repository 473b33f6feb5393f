"""soil_netcdf.py: the on-disk layout of ``*_in_soil`` / ``*_at_soil_surface`` fields follows
src/utilities/mossco_netcdf.F90 (dimension names :3004/:4157, attributes :1340-1365, time axis :2413-2470)."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from mossco_code_b200 import soil_netcdf as sn


def test_layout_and_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    f = {"dissolved_oxygen_in_soil": np.asfortranarray(rng.random((7, 5, 30))),
         "dissolved_oxygen_upward_flux_at_soil_surface": np.asfortranarray(rng.standard_normal((7, 5)))}
    f["dissolved_oxygen_in_soil"][2, 3, :] = 1.0e20        # a masked column as the driver leaves it
    p = str(tmp_path / "soil.nc")
    sn.write_fields(p, f, 3600.0, units={"dissolved_oxygen_in_soil": "mmol m-3"})
    f2 = {k: v * 1.5 for k, v in f.items()}
    sn.write_fields(p, f2, 7200.0, append=True)

    nc = netcdf_file(p, "r", mmap=False)
    assert nc.dimensions["time"] is None                   # unlimited
    assert nc.dimensions["sedimentFluxes_1_O"] == 7 and nc.dimensions["sedimentFluxes_2_O"] == 5
    assert nc.dimensions["ungridded00030"] == 30
    v3 = nc.variables["dissolved_oxygen_in_soil"]
    assert v3.dimensions == ("time", "ungridded00030", "sedimentFluxes_2_O", "sedimentFluxes_1_O")
    assert v3.typecode() == "d" and v3.units == b"mmol m-3"
    assert v3.missing_value == -1.0e30 and v3._FillValue == -1.0e30
    assert v3.coordinates == b"sedimentFluxes_lon sedimentFluxes_lat"
    v2 = nc.variables["dissolved_oxygen_upward_flux_at_soil_surface"]
    assert v2.dimensions == ("time", "sedimentFluxes_2_O", "sedimentFluxes_1_O")
    assert nc.variables["time"].units.startswith(b"seconds since")
    assert np.array_equal(nc.variables["time"][:], [3600.0, 7200.0])
    # Fortran element (i,j,k) sits at C index [t,k,j,i]
    assert v3[0, 4, 2, 6] == f["dissolved_oxygen_in_soil"][6, 2, 4]
    nc.close()

    g, t = sn.read_fields(p)
    assert t == 7200.0 and all(np.array_equal(g[k], f2[k]) for k in f)
    g, t = sn.read_fields(p, 0, names={"dissolved_oxygen_in_soil"})
    assert t == 3600.0 and list(g) == ["dissolved_oxygen_in_soil"]
    assert np.array_equal(g["dissolved_oxygen_in_soil"], f["dissolved_oxygen_in_soil"])
    assert g["dissolved_oxygen_in_soil"].flags.f_contiguous


def test_mask_writes_fill_value_and_bad_shape_is_refused(tmp_path):
    a = np.ones((4, 3, 2))
    mask = np.zeros((4, 3), dtype=np.int32); mask[1, 1] = 1
    p = str(tmp_path / "m.nc")
    sn.write_fields(p, {"x_in_soil": a}, 0.0, mask=mask)
    g, _ = sn.read_fields(p)
    assert np.all(g["x_in_soil"][1, 1, :] == sn.MISSING_R8) and g["x_in_soil"][0, 0, 0] == 1.0
    with pytest.raises(ValueError):
        sn.write_fields(str(tmp_path / "b.nc"), {"x": a, "y": np.ones((3, 4))}, 0.0)
    assert sn.ungridded_dim_name(40) == "ungridded00040"
