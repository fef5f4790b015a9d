"""Shared test cases and helpers (test infrastructure)."""
from __future__ import annotations

import numpy as np

import refcase as rc

TOL = {"f64": 1e-12, "f32": 1e-5}   # north_star: field-normalised relative tolerance on pdf / phi


def relerr(a: np.ndarray, b: np.ndarray, scale_of: np.ndarray | None = None) -> float:
    """field-normalised relative error max|a-b| / max|b| (hard part ii).
    scale_of: take the normalisation from this array instead of b (the convective-outlet buffers hold five PDF
    populations of ONE component, which are ~0 while the other phase occupies the outlet: they are compared on the
    scale of the PDF field they are copies of)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = float(np.max(np.abs(b if scale_of is None else scale_of)))
    d = float(np.max(np.abs(a - b)))
    return d / scale if scale > 0 else d


def derived_close(a: np.ndarray, b: np.ndarray, tol: float, weight: np.ndarray | None = None) -> tuple[bool, str]:
    """Comparison for cn_*, c_norm, curv.  These pass through hard thresholds (c_norm < 1e-6 zeroes the normal,
    the secant solver's err > eps branches: /root/reference/src/main_iteration_GPU.cu:795,839,850), so a rounding
    difference can flip a handful of nodes by O(1).  Criterion: all but 0.1 % of the entries within 1e3*tol
    (absolute, the fields are O(1)), and nothing worse than the fields' own magnitude.

    weight: the reference c_norm on the same grid.  The unit normal cn = grad(phi)/|grad(phi)| and the curvature built
    from it are pure rounding noise where |grad(phi)| is barely above the 1e-6 cut-off (bulk of either phase), and the
    time step only ever consumes them multiplied by c_norm (CSF force 0.5*gamma*curv*c_norm*cn, :147-150) or by
    rho1*rho2 (recolouring, :305), both ~0 there.  With a weight the comparison is made on weight*field, i.e. on the
    quantities that enter the update."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    scale = max(1.0, float(np.abs(b).max()))
    if weight is not None:
        w = np.minimum(np.abs(np.asarray(weight, dtype=np.float64)), 1.0)
        d = d * w
    frac = float(np.mean(d > 1e3 * tol * scale))
    worst = float(d.max())
    ok = frac <= 1e-3 and worst <= 2.5 * scale
    return ok, f"outlier fraction {frac:.2e}, worst {worst:.3e}"


def threshold_flips(c_norm_a: np.ndarray, c_norm_b: np.ndarray) -> tuple[int, int]:
    """(nodes where exactly one of the two c_norm arrays is zero, nodes where the reference's is non-zero): how many nodes sit on
    different sides of the |grad phi| < 1e-6 cut-off (/root/reference/src/main_iteration_GPU.cu:795) in the two runs"""
    za, zb = np.asarray(c_norm_a) == 0, np.asarray(c_norm_b) == 0
    return int(np.count_nonzero(za != zb)), int(np.count_nonzero(~zb))


def saturation_of(phi4: np.ndarray, walls2: np.ndarray, ctl: dict) -> tuple[float, float]:
    """saturation and saturation_full_domain of /root/reference/src/Monitor.cpp:34-120 from a phase field (4-ghost layout) and
    the wall flags (2-ghost layout), summed in double: vol1 = sum 0.5 (1 + phi) over fluid nodes, vol2 likewise with (1 - phi);
    `saturation` over the slices n_exclude_inlet < k <= nz - n_exclude_outlet"""
    phi = np.asarray(phi4, dtype=np.float64)[4:-4, 4:-4, 4:-4]
    fluid = np.asarray(walls2)[2:-2, 2:-2, 2:-2] == 0
    v1 = (0.5 * (1.0 + phi) * fluid).sum(axis=(1, 2))
    v2 = (0.5 * (1.0 - phi) * fluid).sum(axis=(1, 2))
    nz = phi.shape[0]
    lo, hi = int(ctl["n_exclude_inlet"]), nz - int(ctl["n_exclude_outlet"])
    return float(v1[lo:hi].sum() / (v1[lo:hi].sum() + v2[lo:hi].sum())), float(v1.sum() / (v1.sum() + v2.sum()))


def derived_weight(c_norm_ref: np.ndarray, key: str) -> np.ndarray | None:
    """weight array for derived_close: c_norm (2 ghosts) cut to the grid of `key` (curv carries 1 ghost)"""
    if key == "c_norm":
        return None
    return c_norm_ref[1:-1, 1:-1, 1:-1] if key == "curv" else c_norm_ref


def case_tube_pressure():
    """tube + sphere, drainage, Zou-He pressure inlet/outlet, porous plate blocking fluid 1 near the outlet"""
    solid = rc.tube_sphere(20, 20, 28, buffer=4)
    ctl = dict(initial_fluid_distribution_option=1, saturation_injection=1.0, inlet_BC=2, outlet_BC=2, theta=65,
               initial_interface_position=6.0, porous_plate_cmd=1, Z_porous_plate=25, n_exclude_inlet=3, n_exclude_outlet=3,
               body_force_0=1e-4)
    return ctl, solid


def case_pack_velocity():
    """sphere pack, drainage, velocity inlet + convective outlet, contact angle 45"""
    solid = rc.sphere_pack(24, 24, 32, radius=4.0, porosity=0.55, buffer=5, seed=7)
    ctl = dict(initial_fluid_distribution_option=1, saturation_injection=1.0, inlet_BC=1, outlet_BC=1, theta=45,
               initial_interface_position=4.0, capillary_number=1e-3, n_exclude_inlet=5, n_exclude_outlet=5, body_force_0=0.0)
    return ctl, solid


def case_imbibition_plate2():
    """shipped-control flavour: imbibition, pressure/pressure, porous plate blocking fluid 2 near the outlet"""
    solid = rc.sphere_pack(20, 20, 30, radius=3.5, porosity=0.6, buffer=5, seed=11)
    ctl = dict(initial_fluid_distribution_option=2, saturation_injection=0.0, inlet_BC=2, outlet_BC=2, theta=65,
               initial_interface_position=4.0, porous_plate_cmd=2, Z_porous_plate=26, n_exclude_inlet=5, n_exclude_outlet=5)
    return ctl, solid


def case_periodic_drop():
    """y/z periodic, body force, a drop of fluid 1 around an obstacle: exercises the 9 periodic kernels"""
    nx, ny, nz = 18, 16, 20
    solid = np.zeros((nz, ny, nx), dtype=np.int8)
    k, j, i = np.meshgrid(np.arange(1, nz + 1), np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
    solid[(i - 9.5) ** 2 + (j - 3.0) ** 2 + (k - 4.0) ** 2 < 9.0] = 1     # obstacle cut by the periodic faces
    ctl = dict(initial_fluid_distribution_option=5, saturation_injection=1.0, inlet_BC=2, outlet_BC=2, theta=30,
               initial_interface_position=5.0, jper=1, kper=1, domain_wall_status_y_min=0, domain_wall_status_y_max=0,
               body_force_0=2e-5, n_exclude_inlet=0, n_exclude_outlet=0)
    return ctl, solid


def case_duct_no_geometry():
    """no external geometry: open duct with the hard-coded modify_geometry obstacle (src/Misc.cpp:105)"""
    ctl = dict(nxGlobal=16, nyGlobal=16, nzGlobal=36, modify_geometry_cmd=1, initial_fluid_distribution_option=1,
               saturation_injection=1.0, inlet_BC=1, outlet_BC=2, theta=60, initial_interface_position=5.0, capillary_number=5e-4)
    return ctl, None


def case_rect_quirk():
    """nx != ny: the reference reads the geometry with swapped strides (src/Misc.cpp:171)"""
    solid = rc.sphere_pack(22, 18, 24, radius=3.5, porosity=0.6, buffer=4, seed=3)
    ctl = dict(initial_fluid_distribution_option=1, saturation_injection=1.0, inlet_BC=2, outlet_BC=2, theta=50,
               initial_interface_position=4.0)
    return ctl, solid


CASES = {
    "tube_pressure": case_tube_pressure,
    "pack_velocity": case_pack_velocity,
    "imbibition_plate2": case_imbibition_plate2,
    "periodic_drop": case_periodic_drop,
    "duct_no_geometry": case_duct_no_geometry,
    "rect_quirk": case_rect_quirk,
}


def full_control(name: str) -> tuple[dict, np.ndarray | None]:
    ctl, solid = CASES[name]()
    full = dict(rc.DEFAULT_CONTROL)
    full.update(ctl)
    if solid is not None:
        nz, ny, nx = solid.shape
        full.update(nxGlobal=nx, nyGlobal=ny, nzGlobal=nz, external_geometry_read_cmd=1)
    else:
        full["external_geometry_read_cmd"] = 0
    return full, solid


def make_oracle(name: str, prec: str):
    from oracle import Oracle
    ctl, solid = full_control(name)
    o = Oracle(ctl, prec)
    o.setup(solid)
    return o, ctl, solid


def solver_from_oracle(o, ctl, prec: str, **kw):
    """A CUDA solver whose geometry and state are uploaded from an oracle context (both in reference layouts)."""
    import mflbm
    P = mflbm.derive_params(ctl, prec)
    s = mflbm.Solver(P, prec, **kw)
    s.upload_geometry(o.arr("walls"), o.arr("walls_type"), o.arr("s_nx"), o.arr("s_ny"), o.arr("s_nz"))
    s.upload_state(pdf=o.arr("pdf"), phi=o.arr("phi"), cn_x=o.arr("cn_x"), cn_y=o.arr("cn_y"), cn_z=o.arr("cn_z"),
                   c_norm=o.arr("c_norm"), curv=o.arr("curv"), W_in=o.arr("W_in"), f_convec=o.arr("f_convec"),
                   g_convec=o.arr("g_convec"), phi_convec=o.arr("phi_convec"))
    return s


def fluid_mask(o, g: int) -> np.ndarray:
    """boolean [nz+2g, ny+2g, nx+2g] mask of real fluid nodes"""
    w = o.arr("walls")   # 2 ghosts
    nz, ny, nx = o.nz, o.ny, o.nx
    m = np.zeros((nz + 2 * g, ny + 2 * g, nx + 2 * g), dtype=bool)
    m[g:g + nz, g:g + ny, g:g + nx] = w[2:2 + nz, 2:2 + ny, 2:2 + nx] == 0
    return m
