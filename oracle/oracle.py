"""ctypes front-end of the C oracle (oracle/mflbm_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg - never by the product.
Arrays are exposed as numpy views of the oracle's own memory, shaped [z, y, x] (x fastest) with the
reference's ghost widths (pdf: [2, 19, nz+2, ny+2, nx+2]).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def build(force: bool = False) -> None:
    src = HERE / "mflbm_oracle.c"
    libs = [HERE / "liboracle_f64.so", HERE / "liboracle_f32.so"]
    if force or any((not l.exists()) or l.stat().st_mtime < src.stat().st_mtime for l in libs):
        subprocess.run(["make", "-C", str(HERE), "all"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def _params_struct(real):
    class Params(C.Structure):
        _fields_ = [
            ("nx", C.c_longlong), ("ny", C.c_longlong), ("nz", C.c_longlong),
            ("iper", C.c_int), ("jper", C.c_int), ("kper", C.c_int),
            ("wall_x_min", C.c_int), ("wall_x_max", C.c_int), ("wall_y_min", C.c_int), ("wall_y_max", C.c_int),
            ("wall_z_min", C.c_int), ("wall_z_max", C.c_int),
            ("inlet_BC", C.c_int), ("outlet_BC", C.c_int),
            ("porous_plate_cmd", C.c_int), ("Z_porous_plate", C.c_int),
            ("n_exclude_inlet", C.c_int), ("n_exclude_outlet", C.c_int),
            ("mrt", C.c_int), ("rho_out_BC", C.c_int),
            ("la_nu1", real), ("la_nu2", real), ("lbm_gamma", real), ("theta_deg", real), ("lbm_beta", real),
            ("sa_inject", real), ("ca_0", real), ("force_z0", real),
        ]
    return Params


class Oracle:
    """One oracle context = one lattice in the reference's global layouts."""

    def __init__(self, control: dict, prec: str = "f64", mrt: int = 2):
        build()
        self.prec = prec
        self.rt = np.float64 if prec == "f64" else np.float32
        real = C.c_double if prec == "f64" else C.c_float
        self.lib = C.CDLL(str(HERE / f"liboracle_{prec}.so"))
        self.Params = _params_struct(real)
        L = self.lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(self.Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_array.restype = C.c_void_p
        L.orc_array.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_scalar.restype = C.c_double
        L.orc_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_set_walls.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int]
        L.orc_geometry_preprocess.argtypes = [C.c_void_p]
        L.orc_init_new.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_init_new.restype = C.c_int
        L.orc_color_gradient.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_monitor.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cal_saturation.argtypes = [C.c_void_p]
        L.orc_cal_saturation.restype = C.c_double
        for name in ("orc_collide",):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong]
        for name in ("orc_inlet_velocity", "orc_inlet_pressure", "orc_outlet_convective", "orc_outlet_pressure", "orc_porous_plate",
                     "orc_periodic_pdf_edges"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong]
        L.orc_periodic_pdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong]
        L.orc_periodic_phi.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong]
        for name in ("orc_extrapolate_phi_to_solid", "orc_normal_directions", "orc_alter_color_gradient",
                     "orc_extrapolate_normal_to_solid", "orc_csf_curvature"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]

        c = control
        p = self.Params()
        p.nx, p.ny, p.nz = c["nxGlobal"], c["nyGlobal"], c["nzGlobal"]
        p.iper, p.jper, p.kper = c["iper"], c["jper"], c["kper"]
        p.wall_x_min, p.wall_x_max = c["domain_wall_status_x_min"], c["domain_wall_status_x_max"]
        p.wall_y_min, p.wall_y_max = c["domain_wall_status_y_min"], c["domain_wall_status_y_max"]
        p.wall_z_min, p.wall_z_max = c["domain_wall_status_z_min"], c["domain_wall_status_z_max"]
        p.inlet_BC, p.outlet_BC = c["inlet_BC"], c["outlet_BC"]
        p.porous_plate_cmd, p.Z_porous_plate = c["porous_plate_cmd"], c["Z_porous_plate"]
        p.n_exclude_inlet, p.n_exclude_outlet = c["n_exclude_inlet"], c["n_exclude_outlet"]
        p.mrt = mrt
        p.rho_out_BC = c["rho_out_BC"]
        p.la_nu1, p.la_nu2 = c["fluid1_viscosity"], c["fluid2_viscosity"]
        p.lbm_gamma, p.theta_deg, p.lbm_beta = c["surface_tension"], c["theta"], c["RK_beta"]
        p.sa_inject, p.ca_0, p.force_z0 = c["saturation_injection"], c["capillary_number"], c["body_force_0"]
        self.params = p
        self.control = dict(c)
        self.nx, self.ny, self.nz = int(p.nx), int(p.ny), int(p.nz)
        self.h = L.orc_create(C.byref(p))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- array views -----------------------------------------------------------------------------------
    def _shape(self, g):
        return (self.nz + 2 * g, self.ny + 2 * g, self.nx + 2 * g)

    def arr(self, name: str) -> np.ndarray:
        ptr = self.lib.orc_array(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        shapes = {
            "walls_global": (np.int32, self._shape(0)), "walls": (np.int32, self._shape(2)), "walls_type": (np.int32, self._shape(4)),
            "pore_profile_z": (np.int32, (self.nz,)),
            "s_nx": (self.rt, self._shape(4)), "s_ny": (self.rt, self._shape(4)), "s_nz": (self.rt, self._shape(4)),
            "phi": (self.rt, self._shape(4)),
            "cn_x": (self.rt, self._shape(2)), "cn_y": (self.rt, self._shape(2)), "cn_z": (self.rt, self._shape(2)),
            "c_norm": (self.rt, self._shape(2)), "curv": (self.rt, self._shape(1)),
            "pdf": (self.rt, (2, 19) + self._shape(1)),
            "W_in": (self.rt, (self.ny + 2, self.nx + 2)),
            "f_convec": (self.rt, (19, self.ny + 2, self.nx + 2)), "g_convec": (self.rt, (19, self.ny + 2, self.nx + 2)),
            "phi_convec": (self.rt, (self.ny + 2, self.nx + 2)),
            "u": (self.rt, self._shape(1)), "v": (self.rt, self._shape(1)), "w": (self.rt, self._shape(1)), "rho": (self.rt, self._shape(1)),
        }
        dt, shp = shapes[name]
        n = int(np.prod(shp))
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shp)

    def scalar(self, name: str) -> float:
        return float(self.lib.orc_scalar(self.h, name.encode()))

    def set_scalar(self, name: str, v: float) -> None:
        self.lib.orc_set_scalar(self.h, name.encode(), float(v))

    # -- set-up ----------------------------------------------------------------------------------------
    def set_walls(self, solid: np.ndarray | None, compat_stride: bool = True) -> None:
        mg = int(self.control.get("modify_geometry_cmd", 0))
        if solid is None:
            self.lib.orc_set_walls(self.h, None, 0, 0, 0, int(compat_stride), mg)
        else:
            s = np.ascontiguousarray(solid, dtype=np.int8)
            nz, ny, nx = s.shape
            self.lib.orc_set_walls(self.h, s.ctypes.data, nx, ny, nz, int(compat_stride), mg)

    def geometry_preprocess(self) -> None:
        self.lib.orc_geometry_preprocess(self.h)

    def init_new(self) -> None:
        rc = self.lib.orc_init_new(self.h, int(self.control["initial_fluid_distribution_option"]),
                                   float(self.rt(self.control["initial_interface_position"])))
        if rc:
            raise ValueError("unsupported initial_fluid_distribution_option")

    def color_gradient(self) -> None:
        self.lib.orc_color_gradient(self.h)

    def setup(self, solid: np.ndarray | None, compat_stride: bool = True) -> "Oracle":
        """everything src/main.cpp:71-92 does for a new simulation"""
        self.set_walls(solid, compat_stride)
        self.geometry_preprocess()
        self.init_new()
        self.color_gradient()
        return self

    # -- stepping --------------------------------------------------------------------------------------
    def step(self, ntime: int) -> None:
        self.lib.orc_step(self.h, int(ntime))

    def run(self, ntime_first: int, nsteps: int) -> None:
        self.lib.orc_run(self.h, int(ntime_first), int(nsteps))

    def monitor(self):
        out = np.zeros(10, dtype=np.float64)
        prof = np.zeros(7 * self.nz, dtype=np.float64)
        self.lib.orc_monitor(self.h, out.ctypes.data, prof.ctypes.data)
        names = ["saturation", "saturation_full_domain", "vol1_sum", "vol2_sum", "mass1_sum", "mass2_sum", "ca", "umax_global",
                 "kinetic_energy1", "kinetic_energy2"]
        return dict(zip(names, out.tolist())), prof.reshape(7, self.nz)

    def cal_saturation(self) -> float:
        return float(self.lib.orc_cal_saturation(self.h))

    def state(self) -> dict:
        names = ["pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"]
        if self.control["outlet_BC"] == 1:
            names += ["f_convec", "g_convec", "phi_convec"]
        return {n: self.arr(n).copy() for n in names}

    def load_state(self, st: dict) -> None:
        for k, v in st.items():
            key = {"f_convec_bc": "f_convec", "g_convec_bc": "g_convec", "phi_convec_bc": "phi_convec"}.get(k, k)
            self.arr(key)[...] = v
