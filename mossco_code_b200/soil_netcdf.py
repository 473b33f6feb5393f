"""``<name>_in_soil`` / ``*_at_soil_surface`` fields on disk, in the layout MOSSCO's netcdf_component writes
and netcdf_input_component reads back for ReadRestart.

What is reproduced from src/utilities/mossco_netcdf.F90 (the reference's netCDF writer):
- one unlimited ``time`` dimension, variable ``time`` (double, ``units = 'seconds since ...'``) (:2413-2470);
- horizontal dimensions ``<geomName>_<i>_<stagger>`` with the stagger suffix ``O`` for cell centres
  (:3004, staggerLocSuffix :3992-4020); the sediment's flux grid is called ``sedimentFluxes``
  (src/components/fabm_sediment_component.F90:903);
- the vertical (ungridded) dimension ``ungridded%05d`` named after its length (:4157);
- variables in Fortran dimension order (x, y, ungridded, time), i.e. C order (time, ungridded, y, x);
- per-variable attributes ``long_name``, ``units``, ``coordinates``, ``missing_value`` and ``_FillValue``
  = -1e30 for R8 fields (:180, :1340-1365); fields are stored as NF90_DOUBLE (``precision='NF90_DOUBLE'``,
  :1315 -- a restart must not lose bits).

Written with scipy's pure-python NetCDF-3 (64-bit offset) writer: any netCDF library opens it.  This is
file-format glue next to the hot path; it never touches the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
from scipy.io import netcdf_file

MISSING_R8 = -1.0e30
STAGGER_CENTER = "O"


def ungridded_dim_name(length: int) -> str:
    return "ungridded%05d" % length            # '(A9,I5.5)', mossco_netcdf.F90:4157


def write_fields(path: str, fields: Dict[str, np.ndarray], time_seconds: float, *,
                 geom_name: str = "sedimentFluxes", units: Optional[Dict[str, str]] = None,
                 time_units: str = "seconds since 2000-01-01 00:00:00", mask: Optional[np.ndarray] = None,
                 append: bool = False) -> None:
    """Write (or append one time record of) rank-2 ``(inum,jnum)`` / rank-3 ``(inum,jnum,knum)`` fields."""
    units = units or {}
    first = next(iter(fields.values()))
    inum, jnum = first.shape[0], first.shape[1]
    xdim, ydim = f"{geom_name}_1_{STAGGER_CENTER}", f"{geom_name}_2_{STAGGER_CENTER}"
    if append:
        nc = netcdf_file(path, "a")
        rec = nc.variables["time"].shape[0]
    else:
        nc = netcdf_file(path, "w", version=2)
        nc.createDimension("time", None)
        nc.createDimension(xdim, inum)
        nc.createDimension(ydim, jnum)
        t = nc.createVariable("time", "d", ("time",))
        t.units = time_units
        t.standard_name = "time"
        rec = 0
        nc.mossco_layout = "mossco_netcdf.F90 dimension/attribute scheme, written by mossco_code_b200"
    try:
        nc.variables["time"][rec] = float(time_seconds)
        for name, a in fields.items():
            a = np.asarray(a, dtype=np.float64)
            if a.shape[:2] != (inum, jnum) or a.ndim not in (2, 3):
                raise ValueError(f"{name}: shape {a.shape} does not fit the ({inum},{jnum}) grid")
            if name not in nc.variables:
                dims = ["time"]
                if a.ndim == 3:
                    zdim = ungridded_dim_name(a.shape[2])
                    if zdim not in nc.dimensions:
                        nc.createDimension(zdim, a.shape[2])
                    dims.append(zdim)
                dims += [ydim, xdim]
                v = nc.createVariable(name, "d", tuple(dims))
                v.long_name = name
                v.units = units.get(name, "")
                v.coordinates = f"{geom_name}_lon {geom_name}_lat"
                v.missing_value = MISSING_R8
                v._FillValue = MISSING_R8
            out = a
            if mask is not None:                               # masked cells carry the fill value
                out = a.copy()
                out[np.asarray(mask) != 0] = MISSING_R8
            nc.variables[name][rec] = np.ascontiguousarray(out.T)   # (i,j[,k]) -> ([k,]j,i)
    finally:
        nc.close()


def read_fields(path: str, record: int = -1, names=None):
    """Return ``(fields, time_seconds)`` of one time record; arrays come back in ``(inum,jnum[,knum])``
    Fortran-order layout, fill values as stored (ReadRestart leaves masked cells to check_domain)."""
    nc = netcdf_file(path, "r", mmap=False)
    try:
        t = float(nc.variables["time"][record])
        out = {}
        for name, v in nc.variables.items():
            if name == "time" or (names is not None and name not in names):
                continue
            if not v.dimensions or v.dimensions[0] != "time":
                continue
            out[name] = np.asfortranarray(np.array(v[record], dtype=np.float64).T)
        return out, t
    finally:
        nc.close()
