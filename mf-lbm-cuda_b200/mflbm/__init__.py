"""Thin ctypes binding of libmflbm.so (include/mflbm.h) for tests and bench.py.

This is plumbing, not the product: the product is the CUDA library and the C++ host driver.  There is no
CPU fallback anywhere in this module - if the CUDA library is missing or fails to load, importing
:func:`load_library` raises.

Arrays cross the boundary as numpy arrays in the reference's layouts, shaped [z, y, x] with x fastest
(/root/reference/includes/Idx_gpu.cuh:52-70); pdf is [2, 19, nz+2, ny+2, nx+2].
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_ROOT / "lib" / "libmflbm.so"
CSRC = PKG_ROOT / "csrc"

_lib = None


def build_library(force: bool = False) -> Path:
    """nvcc-compile csrc/ for sm_100a into lib/libmflbm.so (in-tree, so it travels to the GPU box)."""
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_ROOT.parent / "include" / "mflbm.h"]
    newest = max(s.stat().st_mtime for s in srcs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        r = subprocess.run(["make", "-C", str(CSRC), "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libmflbm.so failed:\n" + r.stdout[-4000:])
    return LIB_PATH


class MonitorOut(C.Structure):
    _fields_ = [
        ("saturation", C.c_double), ("saturation_full_domain", C.c_double),
        ("vol1_sum", C.c_double), ("vol2_sum", C.c_double), ("mass1_sum", C.c_double), ("mass2_sum", C.c_double),
        ("vol1_full", C.c_double), ("vol2_full", C.c_double), ("mass1_full", C.c_double), ("mass2_full", C.c_double),
        ("fl1_avg", C.c_double), ("fl2_avg", C.c_double), ("fl1_avg_whole", C.c_double), ("fl2_avg_whole", C.c_double),
        ("ca", C.c_double), ("umax", C.c_double), ("kinetic_energy", C.c_double * 2),
        ("nan_detected", C.c_int32), ("reserved", C.c_int32),
        ("fl1", C.c_void_p), ("fl2", C.c_void_p), ("pre", C.c_void_p), ("mass1", C.c_void_p), ("mass2", C.c_void_p),
        ("vol1", C.c_void_p), ("vol2", C.c_void_p),
        ("pre_w_sum", C.c_double), ("pre_nw_sum", C.c_double), ("n_w", C.c_int64), ("n_nw", C.c_int64), ("outlet_phase1_count", C.c_int64),
    ]


class Slab(C.Structure):
    _fields_ = [("x0", C.c_int64), ("nx_local", C.c_int64), ("has_left", C.c_int32), ("has_right", C.c_int32)]


def _params_struct(real):
    class Params(C.Structure):
        _fields_ = [
            ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64),
            ("iper", C.c_int32), ("jper", C.c_int32), ("kper", C.c_int32),
            ("wall_z_min", C.c_int32), ("wall_z_max", C.c_int32),
            ("inlet_BC", C.c_int32), ("outlet_BC", C.c_int32),
            ("porous_plate_cmd", C.c_int32), ("Z_porous_plate", C.c_int32),
            ("n_exclude_inlet", C.c_int32), ("n_exclude_outlet", C.c_int32),
            ("mrt", C.c_int32),
            ("lbm_gamma", real), ("lbm_beta", real), ("la_nu1", real), ("la_nui1", real), ("la_nui2", real), ("cos_theta", real),
            ("force_z", real), ("rho_in", real), ("rho_out", real), ("phi_inlet", real), ("sa_inject", real), ("uin_avg", real),
            ("relaxation", real), ("A_xy", real),
        ]
    return Params


PARAMS = {"f32": _params_struct(C.c_float), "f64": _params_struct(C.c_double)}
REAL = {"f32": np.float32, "f64": np.float64}
CREAL = {"f32": C.c_float, "f64": C.c_double}


def load_library() -> C.CDLL:
    """Load libmflbm.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.mflbm_last_error.restype = C.c_char_p
    lib.mflbm_version.restype = C.c_int
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    for p in ("f32", "f64"):
        r = CREAL[p]
        sig = {
            "create": [C.POINTER(PARAMS[p]), C.POINTER(Slab), i32, vp, C.POINTER(vp)],
            "destroy": [vp],
            "set_params": [vp, C.POINTER(PARAMS[p])],
            "upload_geometry": [vp, vp, vp, vp, vp, vp],
            "preprocess_geometry": [vp, vp],
            "download_geometry": [vp, vp, vp, vp, vp, vp, vp],
            "upload_state": [vp] + [vp] * 11,
            "init_state": [vp, i32, r, vp],
            "init_state_from_phi": [vp, vp, vp],
            "download_state": [vp] + [vp] * 10,
            "step": [vp, i32],
            "run": [vp, i32, i32],
            "color_gradient": [vp],
            "monitor": [vp, C.POINTER(MonitorOut)],
            "sync": [vp],
            "phi_change": [vp, i32, C.POINTER(C.c_double)],
            "download_macro": [vp, vp, vp, vp, vp],
            "halo_buffers": [vp, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)],
            "halo_pack": [vp, i32],
            "halo_unpack": [vp, i32],
            "halo_p2p_local": [vp, i32, i32, C.POINTER(vp), C.POINTER(vp)],
            "halo_p2p_region": [vp, C.POINTER(vp), C.POINTER(i64)],
            "halo_p2p_connect": [vp, i32, i32, vp, vp],
            "halo_push": [vp, i32],
            "halo_unpack_wait": [vp, i32],
            "step_phase": [vp, i32, i32],
        }
        for name, argtypes in sig.items():
            fn = getattr(lib, f"mflbm_{p}_{name}")
            fn.argtypes = argtypes
            fn.restype = C.c_int
        getattr(lib, f"mflbm_{p}_num_fluid_nodes").argtypes = [vp]
        getattr(lib, f"mflbm_{p}_num_fluid_nodes").restype = i64
        getattr(lib, f"mflbm_{p}_kernel_launches").argtypes = [vp]
        getattr(lib, f"mflbm_{p}_kernel_launches").restype = i64
        getattr(lib, f"mflbm_{p}_stream").argtypes = [vp]
        getattr(lib, f"mflbm_{p}_stream").restype = vp
        getattr(lib, f"mflbm_{p}_device_ptr").argtypes = [vp, C.c_char_p]
        getattr(lib, f"mflbm_{p}_device_ptr").restype = vp
    for name, argtypes in (("mflbm_ipc_export", [vp, vp]), ("mflbm_ipc_import", [vp, C.POINTER(vp)]), ("mflbm_ipc_release", [vp])):
        getattr(lib, name).argtypes = argtypes
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


EXPORTED = ["create", "destroy", "set_params", "upload_geometry", "preprocess_geometry", "download_geometry", "upload_state",
            "init_state", "init_state_from_phi", "download_state", "step", "run", "color_gradient", "monitor", "sync", "phi_change", "download_macro", "halo_buffers", "halo_pack",
            "halo_unpack", "halo_p2p_local", "halo_p2p_region", "halo_p2p_connect", "halo_push", "halo_unpack_wait", "step_phase",
            "num_fluid_nodes", "kernel_launches", "stream", "device_ptr"]


class MflbmError(RuntimeError):
    pass


def derive_params(control: dict, prec: str, mrt: int = 2):
    """Host-side derivation of the GPU layer's scalars from control-file keys, in the solver precision, following
    /root/reference/src/IO_multiphase.cpp:200-206 and src/Init_multiphase.cpp:128-214,258-266 step by step.
    (The C++ host driver does the same in mf-lbm-cuda_b200/host/case.hpp; tests cross-check both against the oracle.)"""
    R = REAL[prec]
    c = control
    P = PARAMS[prec]()
    P.nx, P.ny, P.nz = c["nxGlobal"], c["nyGlobal"], c["nzGlobal"]
    P.iper, P.jper, P.kper = c["iper"], c["jper"], c["kper"]
    P.wall_z_min, P.wall_z_max = c["domain_wall_status_z_min"], c["domain_wall_status_z_max"]
    P.inlet_BC, P.outlet_BC = c["inlet_BC"], c["outlet_BC"]
    P.porous_plate_cmd, P.Z_porous_plate = c["porous_plate_cmd"], c["Z_porous_plate"]
    P.n_exclude_inlet, P.n_exclude_outlet = c["n_exclude_inlet"], c["n_exclude_outlet"]
    P.mrt = mrt
    la_nu1, la_nu2 = R(c["fluid1_viscosity"]), R(c["fluid2_viscosity"])
    gamma, beta = R(c["surface_tension"]), R(c["RK_beta"])
    sa, ca0, f0 = R(c["saturation_injection"]), R(c["capillary_number"]), R(c["body_force_0"])
    Pi = R(3.14159265358979323846)
    theta = R(180.) - R(c["theta"])
    theta = R(theta * Pi) / R(180.)
    cos_theta = R(math.cos(float(theta))) if prec == "f64" else np.cos(theta, dtype=np.float32)
    la_y = R(R(R(c["nyGlobal"] - 1) - R(0.5)) - R(0.5))
    la_x = R(R(R(c["nxGlobal"] - 1) - R(0.5)) - R(0.5))
    A_xy = R(la_x * la_y)
    force_z, rho_out, rho_in, uin_avg = f0, R(1.), R(0.), R(0.)
    if c["kper"] == 0 and c["domain_wall_status_z_min"] == 0 and c["domain_wall_status_z_max"] == 0:
        if c["inlet_BC"] == 1:
            force_z = R(0.)
            uin_avg = R(R(ca0 * gamma) / la_nu1)
        elif c["inlet_BC"] == 2:
            force_z = R(0.)
            p_gradient = R(-f0 / R(3.))
            if c["rho_out_BC"]:
                rho_out = R(R(1.) - R(p_gradient * R(c["nzGlobal"])))
            else:
                rho_in = R(rho_out - R(p_gradient * R(c["nzGlobal"])))
    P.lbm_gamma, P.lbm_beta, P.la_nu1 = gamma, beta, la_nu1
    P.la_nui1, P.la_nui2 = R(R(1.) / la_nu1), R(R(1.) / la_nu2)
    P.cos_theta = cos_theta
    P.force_z, P.rho_in, P.rho_out = force_z, rho_in, rho_out
    P.phi_inlet = R(R(2.) * sa - R(1.))
    P.sa_inject, P.uin_avg, P.relaxation, P.A_xy = sa, uin_avg, R(1.), A_xy
    return P


DEFAULT_CHAIN = "csr"   # what csrc/mflbm.cu does when MFLBM_CHAIN is unset


class Solver:
    """One solver handle = one lattice (or x-slab) resident on one GPU."""

    def __init__(self, params, prec: str = "f64", slab: Slab | None = None, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        self.prec = prec
        self.rt = REAL[prec]
        self.params = params
        self.h = C.c_void_p()
        self._slab = slab
        self.nx = int(slab.nx_local) if slab is not None else int(params.nx)
        self.ny, self.nz = int(params.ny), int(params.nz)
        # the library reads MFLBM_CHAIN when a solver is created ("list": the four list kernels of kernels_step.cuh instead of
        # the brick chain of kernels_chain.cuh; "csr": the brick chain with per-brick site lists; results are identical);
        # recorded here so that callers can report it
        self.chain = os.environ.get("MFLBM_CHAIN", "")
        if self.chain not in ("list", "brick", "csr"):
            self.chain = DEFAULT_CHAIN
        rc = self._fn("create")(C.byref(params), C.byref(slab) if slab is not None else None, device, stream, C.byref(self.h))
        self._check(rc)

    # -- plumbing ------------------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, f"mflbm_{self.prec}_{name}")

    def _check(self, rc):
        if rc != 0:
            raise MflbmError(self.lib.mflbm_last_error().decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self._fn("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, g):
        return (self.nz + 2 * g, self.ny + 2 * g, self.nx + 2 * g)

    def _in(self, a, dtype, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.size != int(np.prod(shape)):
            raise ValueError(f"array has {a.size} elements, expected shape {shape}")
        self._keep.append(a)
        return a.ctypes.data

    # -- geometry ------------------------------------------------------------------------------------
    def upload_geometry(self, walls, walls_type, s_nx, s_ny, s_nz):
        self._keep = []
        rc = self._fn("upload_geometry")(self.h, self._in(walls, np.int32, self._shape(2)), self._in(walls_type, np.int32, self._shape(4)),
                                         self._in(s_nx, self.rt, self._shape(4)), self._in(s_ny, self.rt, self._shape(4)),
                                         self._in(s_nz, self.rt, self._shape(4)))
        self._check(rc)

    def preprocess_geometry(self, walls_interior_global):
        """walls_interior_global: int8 [nzGlobal, nyGlobal, nxGlobal], 1 = solid, after set_walls"""
        a = np.ascontiguousarray(walls_interior_global, dtype=np.int8)
        if a.shape != (int(self.params.nz), int(self.params.ny), int(self.params.nx)):
            raise ValueError("walls_interior_global must have the GLOBAL lattice shape [nz, ny, nx]")
        self._check(self._fn("preprocess_geometry")(self.h, a.ctypes.data))

    def download_geometry(self):
        out = dict(walls=np.empty(self._shape(2), np.int32), walls_type=np.empty(self._shape(4), np.int32),
                   s_nx=np.empty(self._shape(4), self.rt), s_ny=np.empty(self._shape(4), self.rt), s_nz=np.empty(self._shape(4), self.rt))
        counts = np.zeros(4, np.int64)
        rc = self._fn("download_geometry")(self.h, out["walls"].ctypes.data, out["walls_type"].ctypes.data, out["s_nx"].ctypes.data,
                                           out["s_ny"].ctypes.data, out["s_nz"].ctypes.data, counts.ctypes.data)
        self._check(rc)
        out["counts"] = counts
        return out

    # -- state ---------------------------------------------------------------------------------------
    def upload_state(self, pdf=None, phi=None, cn_x=None, cn_y=None, cn_z=None, c_norm=None, curv=None, W_in=None,
                     f_convec=None, g_convec=None, phi_convec=None):
        self._keep = []
        s1, s2, s4 = self._shape(1), self._shape(2), self._shape(4)
        pl = (self.ny + 2, self.nx + 2)
        rc = self._fn("upload_state")(self.h, self._in(pdf, self.rt, (38,) + s1), self._in(phi, self.rt, s4), self._in(cn_x, self.rt, s2),
                                      self._in(cn_y, self.rt, s2), self._in(cn_z, self.rt, s2), self._in(c_norm, self.rt, s2),
                                      self._in(curv, self.rt, s1), self._in(W_in, self.rt, pl), self._in(f_convec, self.rt, (19,) + pl),
                                      self._in(g_convec, self.rt, (19,) + pl), self._in(phi_convec, self.rt, pl))
        self._check(rc)

    def init_state(self, option: int, interface_z0: float, W_in=None):
        self._keep = []
        rc = self._fn("init_state")(self.h, int(option), CREAL[self.prec](float(self.rt(interface_z0))),
                                    self._in(W_in, self.rt, (self.ny + 2, self.nx + 2)))
        self._check(rc)

    def init_state_from_phi(self, phi, W_in=None):
        """equilibrium PDFs at rest + colour gradient from a caller-supplied phase field [nz+8, ny+8, nx+8]"""
        self._keep = []
        rc = self._fn("init_state_from_phi")(self.h, self._in(phi, self.rt, self._shape(4)), self._in(W_in, self.rt, (self.ny + 2, self.nx + 2)))
        self._check(rc)

    def state_shapes(self, convective: bool | None = None) -> dict:
        """shapes of the arrays download_state returns (reference layouts), e.g. to allocate pinned host buffers"""
        if convective is None:
            convective = self.params.outlet_BC == 1
        s1, s2, s4 = self._shape(1), self._shape(2), self._shape(4)
        pl = (self.ny + 2, self.nx + 2)
        out = dict(pdf=(2, 19) + s1, phi=s4, cn_x=s2, cn_y=s2, cn_z=s2, c_norm=s2, curv=s1)
        if convective:
            out.update(f_convec=(19,) + pl, g_convec=(19,) + pl, phi_convec=pl)
        return out

    def download_state(self, convective: bool | None = None, fields=None) -> dict:
        """fields: optional subset of ("pdf","phi","cn_x","cn_y","cn_z","c_norm","curv") to fetch (default all)"""
        if convective is None:
            convective = self.params.outlet_BC == 1 and fields is None
        s1, s2, s4 = self._shape(1), self._shape(2), self._shape(4)
        pl = (self.ny + 2, self.nx + 2)
        shapes = dict(pdf=(2, 19) + s1, phi=s4, cn_x=s2, cn_y=s2, cn_z=s2, c_norm=s2, curv=s1)
        out = {k: np.empty(shp, self.rt) for k, shp in shapes.items() if fields is None or k in fields}
        if convective:
            out.update(f_convec=np.empty((19,) + pl, self.rt), g_convec=np.empty((19,) + pl, self.rt), phi_convec=np.empty(pl, self.rt))
        self.download_state_into(out)
        return out

    def download_state_into(self, out: dict) -> None:
        """fill caller-owned (e.g. pinned) numpy arrays; keys as returned by download_state, missing keys are skipped"""
        ptr = lambda k: out[k].ctypes.data if k in out else None
        rc = self._fn("download_state")(self.h, ptr("pdf"), ptr("phi"), ptr("cn_x"), ptr("cn_y"), ptr("cn_z"), ptr("c_norm"), ptr("curv"),
                                        ptr("f_convec"), ptr("g_convec"), ptr("phi_convec"))
        self._check(rc)

    # -- stepping ------------------------------------------------------------------------------------
    def step(self, ntime: int):
        self._check(self._fn("step")(self.h, int(ntime)))

    def run(self, ntime_first: int, nsteps: int):
        self._check(self._fn("run")(self.h, int(ntime_first), int(nsteps)))

    def step_phase(self, ntime: int, phase: int):
        self._check(self._fn("step_phase")(self.h, int(ntime), int(phase)))

    def color_gradient(self):
        self._check(self._fn("color_gradient")(self.h))

    def sync(self):
        self._check(self._fn("sync")(self.h))

    def set_params(self, params):
        self.params = params
        self._check(self._fn("set_params")(self.h, C.byref(params)))

    def monitor(self, profiles: bool = False):
        m = MonitorOut()
        prof = None
        if profiles:
            prof = {k: np.zeros(self.nz, np.float64) for k in ("fl1", "fl2", "pre", "mass1", "mass2", "vol1", "vol2")}
            for k, a in prof.items():
                setattr(m, k, a.ctypes.data)
        self._check(self._fn("monitor")(self.h, C.byref(m)))
        out = {k: getattr(m, k) for k in ("saturation", "saturation_full_domain", "vol1_sum", "vol2_sum", "mass1_sum", "mass2_sum",
                                          "vol1_full", "vol2_full", "mass1_full", "mass2_full", "fl1_avg", "fl2_avg", "fl1_avg_whole",
                                          "fl2_avg_whole", "ca", "umax", "nan_detected", "pre_w_sum", "pre_nw_sum", "n_w", "n_nw",
                                          "outlet_phase1_count")}
        out["kinetic_energy"] = [m.kinetic_energy[0], m.kinetic_energy[1]]
        if prof is not None:
            out["profiles"] = prof
        return out

    def phi_change(self, seed: bool = False) -> float:
        """max |phi - phi_old| over fluid nodes, then phi_old <- phi (src/Monitor.cpp:279-312)"""
        d = C.c_double(0.0)
        self._check(self._fn("phi_change")(self.h, int(seed), C.byref(d)))
        return d.value

    def download_macro(self) -> dict:
        """rho, u, v, w of compute_macro_vars (src/Misc.cpp:222-274), [nz+2, ny+2, nx+2]"""
        out = {k: np.zeros(self._shape(1), self.rt) for k in ("rho", "u", "v", "w")}
        self._check(self._fn("download_macro")(self.h, *[out[k].ctypes.data for k in ("rho", "u", "v", "w")]))
        return out

    # -- halo exchange -------------------------------------------------------------------------------
    def halo_buffers(self, kind: int, side: int):
        send, recv, cnt = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._fn("halo_buffers")(self.h, kind, side, C.byref(send), C.byref(recv), C.byref(cnt)))
        return send.value, recv.value, cnt.value

    def halo_pack(self, kind: int):
        self._check(self._fn("halo_pack")(self.h, kind))

    def halo_unpack(self, kind: int):
        self._check(self._fn("halo_unpack")(self.h, kind))

    # -- halo messages through peer memory -------------------------------------------------------------
    def halo_p2p_local(self, kind: int, side: int):
        """-> (device address of my receive buffer, device address of my arrival flag) for message `kind` from `side`"""
        recv, flag = C.c_void_p(), C.c_void_p()
        self._check(self._fn("halo_p2p_local")(self.h, kind, side, C.byref(recv), C.byref(flag)))
        return recv.value, flag.value

    def halo_p2p_region(self):
        """-> (base address, bytes) of the single allocation that holds every pointer of halo_p2p_local"""
        base, n = C.c_void_p(), C.c_int64()
        self._check(self._fn("halo_p2p_region")(self.h, C.byref(base), C.byref(n)))
        return base.value, n.value

    def halo_p2p_connect(self, kind: int, side: int, peer_recv: int, peer_flag: int):
        self._check(self._fn("halo_p2p_connect")(self.h, kind, side, C.c_void_p(peer_recv), C.c_void_p(peer_flag)))

    def halo_push(self, kind: int):
        self._check(self._fn("halo_push")(self.h, kind))

    def halo_unpack_wait(self, kind: int):
        self._check(self._fn("halo_unpack_wait")(self.h, kind))

    # -- bookkeeping ---------------------------------------------------------------------------------
    @property
    def num_fluid_nodes(self) -> int:
        return int(self._fn("num_fluid_nodes")(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(self._fn("kernel_launches")(self.h))

    @property
    def stream(self) -> int:
        return int(self._fn("stream")(self.h) or 0)

    def chain_bricks(self) -> tuple[int, int]:
        """(bricks the last gradient chain processed, bricks of the lattice) - brick chain only (kernels_chain.cuh)"""
        import torch
        ptr = self.device_ptr("chain_bricks_processed")
        nbx, nby, nbz = (16 * -(-(self.nx + 8) // 16)) // 8, -(-(self.ny + 8) // 4), -(-(self.nz + 8) // 4)
        if not ptr:
            return 0, nbx * nby * nbz
        class _A:
            __cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        self.sync()
        return int(torch.as_tensor(_A(), device="cuda")[0]), nbx * nby * nbz

    def device_ptr(self, name: str) -> int:
        return int(self._fn("device_ptr")(self.h, name.encode()) or 0)


def ipc_export(device_ptr: int) -> bytes:
    """64-byte CUDA IPC handle of the allocation that starts at device_ptr"""
    lib = load_library()
    buf = C.create_string_buffer(64)
    if lib.mflbm_ipc_export(C.c_void_p(device_ptr), buf) != 0:
        raise MflbmError(lib.mflbm_last_error().decode())
    return buf.raw


def ipc_import(handle: bytes) -> int:
    lib = load_library()
    out = C.c_void_p()
    if lib.mflbm_ipc_import(C.create_string_buffer(handle, 64), C.byref(out)) != 0:
        raise MflbmError(lib.mflbm_last_error().decode())
    return out.value


def ipc_release(device_ptr: int) -> None:
    """close a mapping opened by ipc_import"""
    lib = load_library()
    if lib.mflbm_ipc_release(C.c_void_p(device_ptr)) != 0:
        raise MflbmError(lib.mflbm_last_error().decode())
