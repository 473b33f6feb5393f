"""Host-side mirror of the ``fabm_sediment_component`` ESMF interface
(src/components/fabm_sediment_component.F90): the same phases (InitializeP0/P1/P2, ReadRestart, Run,
Finalize, :99-131), the same import/export field names (:943-1192) and the same per-Run sequence
(PAR/porosity import -> get_boundary_conditions -> step loop -> export write-back, :1493-1829),
with the body replaced by C-ABI calls into libmsed_b200.so.

ESMF itself is not available here, so an ``ESMF_State`` is represented by a plain ``dict`` mapping
field names to numpy fp64 arrays in Fortran order (rank 2 ``(inum,jnum)`` for surface fields, rank 3
``(inum,jnum,knum)`` for ``*_in_soil``).  The Fortran shim in ``fortran/msed_b200.F90`` is the
production equivalent of this file.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import _abi
from .sediment import (ADAPTIVE_EULER, NVAR, PARTICULATE, STATE_NAMES, VARIABLE_NAMES, SedimentDriver,
                       default_config, spinup_column)

State = Dict[str, np.ndarray]

# legacy aliases still accepted on import (component :602-605; examples/esmf/sediment/default.dat)
LEGACY_ALIASES = {
    "detritus_labile_carbon": ("fast_detritus_C",),
    "detritus_semilabile_carbon": ("slow_detritus_C",),
    "detritus_labile_phosphorus": ("detritus-P", "detritus_phosphorus"),
}

# export_states catalogue, fabm_sediment_driver.F90:877-924
STATIC_EXPORTS = ("porosity", "layer_height", "layer_center_depth", "temperature",
                  "photosynthetically_active_radiation")
PROFILE3_EXPORTS = ("biomass", "bioturbation", "weighted_toc")

# run_nml defaults, component :59-67
RUN_NML_DEFAULTS = dict(
    numyears=1, dt=360.0, output=-1, numlayers=10, dzmin=0.005, ode_method=ADAPTIVE_EULER,
    dt_min=1.0e-8, relative_change_min=-0.9, bcup_dissolved_variables=2, presimulation_years=-1,
    pel_Temp=5.0, pel_NO3=5.0, pel_NH4=5.0, pel_PO4=0.5, pel_O2=250.0,
    pflux_lDetC=10.0, pflux_sDetC=10.0, pflux_lDetN=1.5, pflux_sDetN=1.5, pflux_lDetP=0.2,
)

ESMF_SUCCESS = 0
ESMF_RC_VAL_OUTOFRANGE = 547  # value ESMF assigns; only identity matters here


class ComponentError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(msg)
        self.rc = rc


class FabmSedimentComponent:
    """One instance per process, like the module-level ``sed`` of the reference (:80)."""

    def __init__(self, name: str = "fabm_sediment"):
        self.name = name
        self.sed: Optional[SedimentDriver] = None
        self.run_nml = dict(RUN_NML_DEFAULTS)
        self.phase_map = None
        self.clock_seconds = 0.0
        self.last_info = None
        self.export_3d_every_run = True
        # <name>_in_soil write-back at an output cadence instead of every Run: every export_cadence-th Run starts
        # an asynchronous export of the state into export_buffer (caller-owned, ideally pinned, shape
        # (inum,jnum,knum,nvar)) that overlaps the following Runs; export_ready() completes it and refreshes the
        # <var>_in_soil entries of the export state (views of export_buffer).  0 = off.
        self.export_cadence = 0
        self.export_buffer = None
        self._export_started = False
        self._runs = 0
        self.flux_buffer = None   # optional caller-owned (pinned) (inum,jnum,nvar) export buffer
        # Import fields the coupler promises not to change between Runs unless it says so with
        # import_changed(name): name suffixes, e.g. ("_z_velocity_at_soil_surface",) for the constant sinking
        # velocities of a pelagic component.  They cross PCIe once (msed_set_import_generations); every other
        # field is uploaded every Run.  Empty = the reference's behaviour (:1865-2030 reads every field every Run).
        self.static_import_suffixes = ()
        self._import_gen = {}     # per field key of msed_set_import_generations
        self._out = None          # output.dat handle (run_nml output > 0, component :266-269)
        self.advance_count = 0    # ESMF clock advanceCount
        self.on_mesh = False      # geometry is a mesh: rank-1 surface fields (see initialize_p1)

    # ---- SetServices -----------------------------------------------------------------------
    def set_services(self):
        """Entry points as registered at :108-128."""
        return {
            ("initialize", 0): self.initialize_p0, ("initialize", 1): self.initialize_p1,
            ("initialize", 2): self.initialize_p2, ("readrestart", 1): self.read_restart,
            ("run", 1): self.run, ("finalize", 1): self.finalize,
        }

    # ---- InitializeP0 (:135-173) -------------------------------------------------------------
    def initialize_p0(self, import_state: State, export_state: State, clock=None):
        self.phase_map = ["IPDv00p1=1", "IPDv00p2=2"]  # NUOPC InitializePhaseMap, :156-164
        return ESMF_SUCCESS

    # ---- InitializeP1 (:177-1201) -------------------------------------------------------------
    def initialize_p1(self, import_state: State, export_state: State, clock=None, *, grid_shape,
                      run_nml: Optional[dict] = None, sed_nml: Optional[dict] = None,
                      fabm_nml: Optional[dict] = None, grid_mask: Optional[np.ndarray] = None,
                      device: int = -1, j_offset: int = 0, output_path: str = "output.dat"):
        """``grid_shape`` = (inum, jnum) of the foreign grid tile (:314-448), or ``(numElements,)`` for a
        mesh (:391-446): the columns then are the owned mesh elements (``_INUM_`` = numElements,
        ``_JNUM_`` = 1), surface fields have rank 1 and ``*_in_soil`` fields rank 2 (element, layer;
        :693-770).  ``grid_mask`` is the ESMF_GRIDITEM_MASK item: columns with grid_mask <= 0 are masked
        (:497-501)."""
        self.on_mesh = len(tuple(grid_shape)) == 1
        if self.on_mesh:
            grid_shape = (int(grid_shape[0]), 1)
            if grid_mask is not None:
                grid_mask = np.asarray(grid_mask).reshape(grid_shape)
        if run_nml:
            unknown = set(run_nml) - set(RUN_NML_DEFAULTS)
            if unknown:
                raise ComponentError(1, f"unknown run_nml entries {sorted(unknown)}")
            self.run_nml.update(run_nml)
        r = self.run_nml
        kw = dict(inum=int(grid_shape[0]), jnum=int(grid_shape[1]), knum=int(r["numlayers"]),
                  dzmin=float(r["dzmin"]), dt_min=float(r["dt_min"]),
                  relative_change_min=float(r["relative_change_min"]),
                  bcup_dissolved_variables=int(r["bcup_dissolved_variables"]), device=device,
                  j_offset=int(j_offset))
        kw.update(sed_nml or {})
        kw.update(fabm_nml or {})
        cfg = default_config(**kw)
        self.cfg = cfg
        self.sed = SedimentDriver(cfg)                      # init_grid (:512) + initialize (:517)
        sed = self.sed
        if grid_mask is not None:
            sed.set_mask((np.asarray(grid_mask) <= 0).astype(np.int32))
            self.mask = np.asfortranarray((np.asarray(grid_mask) <= 0))
        else:
            self.mask = np.zeros(sed.shape2d, dtype=bool, order="F")
        rc = sed.check_domain()                             # :529
        if rc:
            raise ComponentError(rc, "check_domain failed")
        sed.init_concentrations()                           # :536

        # boundary defaults for the pre-simulation (:588-606)
        bd = np.zeros(NVAR + 1)
        fl = np.zeros(NVAR)
        bd[0] = r["pel_Temp"]
        bd[1 + STATE_NAMES.index("no3")] = r["pel_NO3"]
        bd[1 + STATE_NAMES.index("nh3")] = r["pel_NH4"]
        bd[1 + STATE_NAMES.index("po4")] = r["pel_PO4"]
        bd[1 + STATE_NAMES.index("oxy")] = r["pel_O2"]
        bd[1 + STATE_NAMES.index("odu")] = r["pel_O2"]      # reference quirk, :595
        fl[STATE_NAMES.index("ldetC")] = r["pflux_lDetC"] / 86400.0
        fl[STATE_NAMES.index("sdetC")] = r["pflux_sDetC"] / 86400.0
        fl[STATE_NAMES.index("detP")] = r["pflux_lDetP"] / 86400.0
        bdys = np.empty(sed.shape2d + (NVAR + 1,), order="F")
        fluxes = np.empty(sed.shape2d + (NVAR,), order="F")
        bdys[...] = bd
        fluxes[...] = fl
        sed.set_boundary(bdys, fluxes)

        if r["presimulation_years"] > 0:                    # :614-632
            nsteps = int(r["presimulation_years"] * 365 * 24)
            col, info = spinup_column(cfg, bd, fl, nsteps, int(r["ode_method"]))
            sed.set_state_from_column(col)
            self.spinup_info = info
        sed.get_rhs()                                       # fills diagnostics, :638-645

        # export fields (:943-1040) and import fields (:1045-1192)
        self._fill_exports(export_state, with_3d=True)
        for name in self.import_field_names():
            import_state.setdefault(name, None)
        self.clock_seconds = 0.0
        self.advance_count = 0
        if int(r["output"]) > 0:                            # sed%do_output = output .gt. 0, :266
            self._open_output(output_path)
        return ESMF_SUCCESS

    def initialize_p2(self, import_state: State, export_state: State, clock=None):
        return ESMF_SUCCESS                                 # :1206-1324: field completion only

    # ---- names ---------------------------------------------------------------------------------
    def import_field_names(self):
        names = ["temperature_at_soil_surface", "porosity_at_soil_surface",
                 "photosynthetically_active_radiation_at_soil_surface"]
        for n, v in enumerate(VARIABLE_NAMES):
            names.append(f"{v}_at_soil_surface")
            if PARTICULATE[n]:
                names.append(f"{v}_z_velocity_at_soil_surface")
        return names

    def export_field_names(self):
        names = [f"{s}_in_soil" for s in STATIC_EXPORTS]
        names += [f"{v}_in_soil" for v in VARIABLE_NAMES]
        names += [f"{v}_upward_flux_at_soil_surface" for v in VARIABLE_NAMES]
        if self.cfg.bioturbation_profile == 3:
            names += [f"{s}_in_soil" for s in PROFILE3_EXPORTS]
        names.append("denit_in_soil")                       # FABM diagnostic, :1012-1040
        return names

    def _surface(self, f):
        """Import field -> the driver's (inum, jnum) layout (rank-1 mesh fields get a unit j axis)."""
        if f is None or not self.on_mesh:
            return f
        return np.asarray(f, dtype=np.float64).reshape(self.sed.shape2d, order="F")

    def _export(self, a):
        """Driver array (inum, 1[, k]) -> the rank the geometry has (mesh: drop the unit j axis)."""
        return a[:, 0] if self.on_mesh else a

    @staticmethod
    def _lookup(state: State, base: str, suffix: str):
        key = base + suffix
        if state.get(key) is not None:
            return state[key]
        for alias in LEGACY_ALIASES.get(base, ()):
            if state.get(alias + suffix) is not None:
                return state[alias + suffix]
        return None

    # ---- ReadRestart (:1328-1489) -----------------------------------------------------------------
    def read_restart(self, import_state: State, export_state: State, clock=None):
        sed = self.sed
        conc = None
        for n, v in enumerate(VARIABLE_NAMES):
            f = import_state.get(f"{v}_in_soil")
            if f is None:
                continue
            if self.on_mesh and f.ndim == 2:
                f = np.asarray(f)[:, None, :]
            if tuple(f.shape) != sed.shape3d:               # bounds must match (:1440-1460)
                continue
            if conc is None:
                conc = sed.conc
            conc[:, :, :, n] = f
        if conc is not None:
            sed.conc = conc
        por = import_state.get("porosity_in_soil")
        if por is not None and self.on_mesh and por.ndim == 2:
            por = np.asfortranarray(np.asarray(por)[:, None, :])
        if por is not None and tuple(por.shape) == sed.shape3d:
            sed.set_porosity(por)
        rc = sed.check_domain()                              # :1485
        if rc:
            raise ComponentError(rc, "check_domain failed after restart")
        return ESMF_SUCCESS

    # ---- restart files: what netcdf_component writes from the export state and netcdf_input_component
    # ---- feeds to ReadRestart, in mossco_netcdf.F90's layout (see soil_netcdf.py) -----------------
    def write_restart_file(self, path: str, time_seconds: float = None, *, append: bool = False,
                           with_diagnostics: bool = False):
        from . import soil_netcdf
        export: State = {}
        self._fill_exports(export, with_3d=True)
        keep = {f"{v}_in_soil" for v in VARIABLE_NAMES} | {"porosity_in_soil"}
        if with_diagnostics:
            keep = set(export)
        fields = {k: v for k, v in export.items() if k in keep}
        if self.on_mesh:                                     # the file layout is (x, y[, layer]): unit y axis
            fields = {k: np.asarray(v)[:, None] for k, v in fields.items()}
        units ={f"{v}_in_soil": "mmol m-3" for v in VARIABLE_NAMES}
        units.update({f"{v}_upward_flux_at_soil_surface": "mmol m-2 s-1" for v in VARIABLE_NAMES})
        t = self.clock_seconds if time_seconds is None else time_seconds
        soil_netcdf.write_fields(path, fields, t, units=units, append=append)
        return ESMF_SUCCESS

    # ---- import fields that stay on the device (static_import_suffixes) -----------------------------------
    def _generations(self):
        """Counters for msed_set_import_generations: a static field keeps its counter, every other field gets a
        new one each Run."""
        keys = [(0, "temperature_at_soil_surface")]
        for n, v in enumerate(VARIABLE_NAMES):
            keys.append((1 + 2 * n, v + "_at_soil_surface"))
            keys.append((2 + 2 * n, v + "_z_velocity_at_soil_surface"))
        gen = [0] * (1 + 2 * len(VARIABLE_NAMES))
        for k, name in keys:
            static = any(name.endswith(s) for s in self.static_import_suffixes)
            g = self._import_gen.get(k, 1)
            if not static:
                g += 1
            self._import_gen[k] = g
            gen[k] = g
        return gen

    def import_changed(self, name_suffix: str):
        """The coupler has written new data into the static import fields whose names end in ``name_suffix``."""
        for k in list(self._import_gen):
            self._import_gen[k] += 1 << 32

    def read_restart_file(self, path: str, record: int = -1):
        from . import soil_netcdf
        fields, t = soil_netcdf.read_fields(path, record)
        rc = self.read_restart(fields, {})
        self.clock_seconds = t
        return rc

    # ---- Run (:1493-1829) -------------------------------------------------------------------------
    def run(self, import_state: State, export_state: State, clock=None, *, run_seconds: float = None):
        """One coupling interval.  ``clock`` may be a dict with 'currTime'/'stopTime' in seconds."""
        sed = self.sed
        r = self.run_nml
        if run_seconds is None:
            run_seconds = float(clock["stopTime"] - clock["currTime"])
        par = self._surface(import_state.get("photosynthetically_active_radiation_at_soil_surface"))
        if par is not None:                                  # :1568-1596
            sed.set_par_surface(par)
        por = self._surface(import_state.get("porosity_at_soil_surface"))
        if por is not None:                                  # :1619-1643
            sed.update_porosity(por, from_surface=True)
        temp = self._surface(import_state.get("temperature_at_soil_surface"))
        cs = [self._surface(self._lookup(import_state, v, "_at_soil_surface")) for v in VARIABLE_NAMES]
        wz = [self._surface(self._lookup(import_state, v, "_z_velocity_at_soil_surface")) if PARTICULATE[n]
              else None for n, v in enumerate(VARIABLE_NAMES)]
        if self.static_import_suffixes:
            sed.set_import_generations(self._generations())
        if self._out is not None:
            rc, up = self._run_with_output(temp, cs, wz, float(run_seconds))
        else:
            # get_boundary_conditions (:1665) + the step loop (:1700-1769) + the flux export (:1819) in
            # one C-ABI call, so the library can overlap the PCIe transfers with the first/last attempt
            rc, up = sed.run_exchange(float(r["dt"]), int(r["ode_method"]), float(run_seconds), temp, cs,
                                      wz, out=self.flux_buffer)
        self.last_info = sed.info
        self.clock_seconds += float(run_seconds)
        if rc == _abi.NAN_DETECTED:                          # :1718-1723
            raise ComponentError(ESMF_RC_VAL_OUTOFRANGE, "NaN detected applying ode_solver")
        self._fill_exports(export_state, with_3d=self.export_3d_every_run, up=up)   # :1773-1822
        self._runs += 1
        if self.export_cadence > 0 and self._runs % self.export_cadence == 0:
            if self.export_buffer is None:
                self.export_buffer = np.zeros(sed.shape4d, order="F")
            sed.export_state_begin(self.export_buffer)
            self._export_started = True
        return ESMF_SUCCESS

    def export_ready(self, export_state: State) -> bool:
        """Complete the export started by the last cadence Run (if any): waits for the copy and points the
        ``<var>_in_soil`` entries of ``export_state`` at it.  Returns whether an export was pending."""
        if not self._export_started:
            return False
        self.sed.export_state_wait()
        self._export_started = False
        for n, v in enumerate(VARIABLE_NAMES):
            export_state[f"{v}_in_soil"] = self._export(self.export_buffer[:, :, :, n])
        return True

    # ---- output.dat (:677-685 header, :1734-1759 rows) ---------------------------------------------
    @staticmethod
    def _fortran_e(x: float, width: int, digits: int, expw: int = 2) -> str:
        """Fortran Ew.d / Ew.dEe edit descriptor (0.dddE+ee normalisation)."""
        if x != x:
            return "NaN".rjust(width)
        if x == 0.0:
            mant, ex = 0.0, 0
        else:
            ex = int(np.floor(np.log10(abs(x)))) + 1
            mant = abs(x) / 10.0 ** ex
            if round(mant, digits) >= 1.0:
                mant, ex = mant / 10.0, ex + 1
        s = f"{mant:.{digits}f}E{'+' if ex >= 0 else '-'}{abs(ex):0{expw}d}"
        return (("-" if x < 0 else "") + s).rjust(width)

    def _open_output(self, path: str):
        self._out = open(path, "w")
        names = [f"hzg_omexdia_p_{s}" for s in STATE_NAMES] + ["hzg_omexdia_p_denit"]
        self._out.write("time(s) depth(m) layer-height(m) porosity() " + "".join(" " + n for n in names) + "\n")

    def _write_output(self, time_s: float):
        sed, e = self.sed, self._fortran_e
        i0, j0 = np.argwhere(~self.mask)[0] if (~self.mask).any() else (0, 0)   # column (lbnd1,lbnd2)
        conc, por, denit = sed.conc, sed.field("porosity"), sed.field("denit")
        _, zc, dz, _ = sed.grid()
        fl = sed.fluxes
        self._out.write(f" {time_s!r} fluxes " + " ".join(repr(float(v)) for v in fl[0, 0, :]) + "\n")
        for k in range(sed.knum):
            row = e(time_s, 15, 3) + " " + e(zc[k], 15, 4, 3) + " " + e(dz[k], 15, 4, 3) + " " + \
                e(por[i0, j0, k], 15, 4, 3)
            row += "".join(" " + e(conc[i0, j0, k, n], 15, 4, 3) for n in range(NVAR))
            row += " " + e(denit[i0, j0, k], 15, 4, 3)
            self._out.write(row + "\n")
        self._out.flush()

    def _run_with_output(self, temp, cs, wz, run_seconds: float):
        """Run loop split at the output steps: row block after every step with
        mod(advanceCount, output) == 0 (:1737), time label advanceCount*dt."""
        sed, r = self.sed, self.run_nml
        dt, method, every = float(r["dt"]), int(r["ode_method"]), int(r["output"])
        sed.get_boundary_conditions(temp, cs, wz)
        nfull = int(np.floor(run_seconds / dt * (1.0 + 1e-14)))
        rem = run_seconds - nfull * dt
        rc = 0
        for _ in range(nfull):
            rc = sed.step(dt, method, 1)
            if rc:
                break
            if self.advance_count % every == 0:
                self._write_output(self.advance_count * dt)
            self.advance_count += 1
        if rc == 0 and rem > 1e-9 * dt:
            rc = sed.step(rem, method, 1)
            if rc == 0 and self.advance_count % every == 0:
                self._write_output(self.advance_count * rem)   # advanceCount*dt with the shortened dt (:1739)
            self.advance_count += 1
        return rc, sed.upward_fluxes(self.flux_buffer)

    def _fill_exports(self, export_state: State, with_3d: bool, up=None):
        sed = self.sed
        if up is None:
            up = sed.upward_fluxes(self.flux_buffer)
        for n, v in enumerate(VARIABLE_NAMES):
            export_state[f"{v}_upward_flux_at_soil_surface"] = self._export(up[:, :, n])
        if not with_3d:
            return
        conc = sed.conc
        for n, v in enumerate(VARIABLE_NAMES):
            export_state[f"{v}_in_soil"] = self._export(conc[:, :, :, n])
        for s in STATIC_EXPORTS:
            export_state[f"{s}_in_soil"] = self._export(sed.field(s))
        if self.cfg.bioturbation_profile == 3:
            for s in PROFILE3_EXPORTS:
                export_state[f"{s}_in_soil"] = self._export(sed.field(s))
        export_state["denit_in_soil"] = self._export(sed.field("denit"))

    # ---- Finalize (:1833-1861) ------------------------------------------------------------------------
    def finalize(self, import_state: State = None, export_state: State = None, clock=None):
        if self._out is not None:                            # close(funit), :1852
            self._out.close()
            self._out = None
        if self.sed is not None:
            self.sed.finalize()
            self.sed = None
        return ESMF_SUCCESS
