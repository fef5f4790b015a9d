#!/usr/bin/env python
"""bench.py — MLUPS of the colour-gradient two-phase LBM time step on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--prec f64|f32] [--size S] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic SxSxS random sphere pack (S=256, porosity ~0.4, radius 12), drainage,
velocity inlet + convective outlet, contact angle 45 deg, FP64.  For N > 1 the lattice is (S*N) x S x S cut into N
x-slabs (weak scaling; BASELINE configs[4] is --size 512), PDF/phi halos pushed into the neighbour's memory over NVLink
(--halo nccl: NCCL send/recv).  --global-nx NX fixes the lattice at NX x S x S for every N (strong scaling; BASELINE
configs[3] is --global-nx 1024 --size 512).

One JSON line on stdout (rank 0).  `value` = MLUPS with the reference's definition nx*ny*nz*steps/1e6/s
(/root/reference/src/main.cpp:270), state resident in HBM; `e2e` = the same through the C ABI starting from pinned
host arrays (H2D of the full state, K steps, monitor + D2H of the full state - the post-condition the reference's
main_iteration_kernel_GPU gives its host on timer steps, src/main_iteration_GPU.cu:2059-2076).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
for p in (REPO / "mf-lbm-cuda_b200", REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------
def workload_control(nx: int, ny: int, nz: int, case: str = "drainage") -> dict:
    """drainage (BASELINE configs[1-3]): the lattice starts full of fluid 2 and fluid 1 is injected (option 1, saturation_injection 1);
    imbibition (configs[4]): the roles are swapped (option 2, saturation_injection 0)"""
    import refcase as rc
    ctl = dict(rc.DEFAULT_CONTROL)
    opt, inj = (1, 1.0) if case == "drainage" else (2, 0.0)
    ctl.update(nxGlobal=nx, nyGlobal=ny, nzGlobal=nz, initial_fluid_distribution_option=opt, saturation_injection=inj, theta=45,
               initial_interface_position=8.0, inlet_BC=1, outlet_BC=1, capillary_number=1e-4, n_exclude_inlet=10, n_exclude_outlet=10,
               fluid1_viscosity=0.04, fluid2_viscosity=0.4, surface_tension=0.03, RK_beta=0.95, body_force_0=0.0)
    return ctl


def workload_geometry(nx: int, ny: int, nz: int, seed: int = 20240229, kind: str = "pack") -> np.ndarray:
    """interior walls after set_walls (src/Misc.cpp:59-87): sphere pack + solid x/y faces, int8 [nz, ny, nx]"""
    import refcase as rc
    if kind == "open":   # diagnostic only: an empty duct (porosity ~1), upper bound for the kernels' bandwidth efficiency
        solid = np.zeros((nz, ny, nx), dtype=np.int8)
        solid[:, :, 0] = 1; solid[:, :, -1] = 1; solid[:, 0, :] = 1; solid[:, -1, :] = 1
        return solid
    radius = 12.0 * min(nx, ny, nz) / 256.0 if min(nx, ny, nz) < 256 else 12.0
    solid = rc.sphere_pack(nx, ny, nz, radius=max(radius, 3.0), porosity=0.4, buffer=max(2, round(10 * nz / 256)), seed=seed)
    solid[:, :, 0] = 1; solid[:, :, -1] = 1; solid[:, 0, :] = 1; solid[:, -1, :] = 1
    return solid


def workload_geometry_window(nx: int, ny: int, nz: int, xlo: int, xhi: int, seed: int = 20240229, kind: str = "pack") -> np.ndarray:
    """the same geometry as workload_geometry(nx, ny, nz), rasterised only for global columns xlo..xhi (1-based, clipped):
    what one x-slab's preprocessing reads.  The array has the global shape (untouched pages are never committed)."""
    import refcase as rc
    xlo, xhi = max(1, xlo), min(nx, xhi)
    solid = np.zeros((nz, ny, nx), dtype=np.int8)
    if kind == "pack":
        radius = 12.0 * min(nx, ny, nz) / 256.0 if min(nx, ny, nz) < 256 else 12.0
        radius = max(radius, 3.0)
        buffer = max(2, round(10 * nz / 256))
        rng = np.random.Generator(np.random.MT19937(seed))
        n = int(round(-np.log(0.4) * float(nx) * ny * nz / (4.0 / 3.0 * np.pi * radius ** 3)))
        cx, cy, cz = rng.uniform(0.5, nx + 0.5, n), rng.uniform(0.5, ny + 0.5, n), rng.uniform(0.5, nz + 0.5, n)   # same draws as rc.sphere_pack
        r = int(np.ceil(radius))
        sel = (cx + r + 1 >= xlo) & (cx - r - 1 <= xhi)
        for x0, y0, z0 in zip(cx[sel], cy[sel], cz[sel]):
            i0, i1 = max(xlo, int(np.floor(x0)) - r), min(xhi, int(np.ceil(x0)) + r)
            j0, j1 = max(1, int(np.floor(y0)) - r), min(ny, int(np.ceil(y0)) + r)
            k0, k1 = max(1, int(np.floor(z0)) - r), min(nz, int(np.ceil(z0)) + r)
            if i0 > i1 or j0 > j1 or k0 > k1:
                continue
            kk, jj, ii = np.meshgrid(np.arange(k0, k1 + 1), np.arange(j0, j1 + 1), np.arange(i0, i1 + 1), indexing="ij")
            sub = solid[k0 - 1:k1, j0 - 1:j1, i0 - 1:i1]
            sub[(ii - x0) ** 2 + (jj - y0) ** 2 + (kk - z0) ** 2 <= radius ** 2] = 1
        solid[:buffer, :, xlo - 1:xhi] = 0
        solid[nz - buffer:, :, xlo - 1:xhi] = 0
    if xlo == 1:
        solid[:, :, 0] = 1
    if xhi == nx:
        solid[:, :, -1] = 1
    solid[:, 0, xlo - 1:xhi] = 1
    solid[:, -1, xlo - 1:xhi] = 1
    return solid


def inlet_profile(ctl: dict, prec: str) -> np.ndarray:
    """W_in of the velocity inlet (src/Init_multiphase.cpp:258-284, src/Misc.cpp:387-419), evaluated in numpy."""
    R = np.float32 if prec == "f32" else np.float64
    nx, ny = ctl["nxGlobal"], ctl["nyGlobal"]
    la_x, la_y = R(nx - 2), R(ny - 2)
    uin = R(R(ctl["capillary_number"]) * R(ctl["surface_tension"]) / R(ctl["fluid1_viscosity"]))
    a, b = R(0.5) * la_x, R(0.5) * la_y
    n = np.arange(1, 1000, 2).astype(R)
    pi = R(np.pi)
    t1 = np.sum(np.tanh(R(0.5) * n * pi * b / a) / n ** 5, dtype=R)
    t2 = R(1.) - R(192.) / pi ** 5 * (a / b) * t1
    t2 = R(-3.) * uin / (t2 * a ** 2)
    W = np.zeros((ny + 2, nx + 2), dtype=R)
    xx = (np.arange(2, nx).astype(R) - R(1.5) - a)[None, :, None]
    yy = (np.arange(2, ny).astype(R) - R(1.5) - b)[:, None, None]
    nn = n[None, None, :]
    sign = np.where(((n - 1) / 2) % 2 == 0, R(1.), R(-1.))[None, None, :]
    with np.errstate(over="ignore", under="ignore"):
        term = sign * np.cos(R(0.5) * nn * pi * xx / a) / nn ** 3 * (
            R(1.) - (np.exp(R(0.5) * nn * pi * (yy - b) / a) + np.exp(R(0.5) * nn * pi * (-yy - b) / a)) / (R(1.) + np.exp(R(0.5) * nn * pi * (-b - b) / a)))
    W[2:ny, 2:nx] = term.sum(axis=2, dtype=R) * (R(-16.) * t2 * a ** 2 * pi ** -3)
    return W


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def hbm_peak():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def cpu_baseline(prec: str) -> dict:
    """the C oracle (a port: the reference has no CPU solver) on a bounded sample of the same workload"""
    sys.path.insert(0, str(REPO / "oracle"))
    from oracle import Oracle
    n = 64
    ctl = workload_control(n, n, n)
    solid = workload_geometry(n, n, n)
    o = Oracle(ctl, prec)
    o.set_walls(solid)
    o.geometry_preprocess(); o.init_new(); o.color_gradient()
    o.run(1, 2)
    steps, dt, nt = 0, 0.0, 3
    while dt < 12.0:   # bounded sample: ~12 s of CPU work
        t = time.perf_counter(); o.run(nt, 10); dt += time.perf_counter() - t
        steps += 10; nt += 10
    return {"value": n ** 3 * steps / 1e6 / dt, "unit": "MLUPS", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
            "kind": "port", "sample": f"{n}^3 sphere pack, {steps} steps, C oracle with OpenMP (the reference has no CPU solver path)"}


def run_ours(args) -> dict:
    import torch
    import mflbm
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from mflbm import slab as slabmod
        return slabmod.bench_slabs(args, rank, world, local)
    S = args.size
    NX = args.global_nx or S
    prec = args.prec
    ctl = workload_control(NX, S, S, args.case)
    solid = workload_geometry(NX, S, S, seed=args.seed, kind=args.geometry)
    stream = torch.cuda.Stream()
    solver = mflbm.Solver(mflbm.derive_params(ctl, prec), prec, device=local, stream=stream.cuda_stream)
    t0 = time.perf_counter()
    solver.preprocess_geometry(solid)
    t_geo = time.perf_counter() - t0
    W = inlet_profile(ctl, prec)
    solver.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=W)
    n_fluid, n_site = solver.num_fluid_nodes, NX * S * S
    dims = f"{S}^3" if NX == S else f"{NX}x{S}x{S}"
    # ---- device-resident timing -------------------------------------------------------------------
    solver.run(1, args.warmup)
    solver.sync()
    l0 = solver.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nt = 1 + args.warmup
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        e0.record(stream)
        solver.run(nt, args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = solver.kernel_launches - l0
    nt += args.steps
    # ---- per-kernel timing of the dominant launches (collide odd / even), CUDA events on the launching stream ----------
    kpairs = max(2, min(args.steps // 2, 25))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * kpairs)]
    for a, b in ev:
        a.record(stream)
        solver.step_phase(nt, 0)
        b.record(stream)
        solver.step_phase(nt, 1)
        solver.step_phase(nt, 2)
        nt += 1
    torch.cuda.synchronize()
    first_odd = (nt - 2 * kpairs) % 2 == 1
    t_all = [a.elapsed_time(b) for a, b in ev]
    t_odd = float(np.mean(t_all[0::2] if first_odd else t_all[1::2]))
    t_even = float(np.mean(t_all[1::2] if first_odd else t_all[0::2]))
    mon = solver.monitor()
    assert mon["nan_detected"] == 0, "simulation produced non-finite values"
    ms_step = ms / args.steps
    mlups = n_site * args.steps / 1e6 / (ms * 1e-3)
    s_bytes = 8 if prec == "f64" else 4
    bytes_step = 78 * s_bytes * n_fluid + n_site
    peak, peak_src = hbm_peak()
    achieved_step = bytes_step / (ms_step * 1e-3) / 1e9
    # dominant kernel = the collide launch (one per step, alternating odd / even): 38 PDF reads + 38 PDF writes + the phi
    # write per fluid node; the remaining s + 1 B/site of the step model belong to the gradient chain
    bytes_collide = 77 * s_bytes * n_fluid
    t_collide = 0.5 * (t_odd + t_even)
    achieved = bytes_collide / (t_collide * 1e-3) / 1e9
    traffic = None
    tf = REPO / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(f"{prec}_{S}", None) if NX == S else None
        except Exception:
            traffic = None
    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    if args.no_e2e:
        h2d = d2h = 0; e2e_mlups = None; mon2 = mon
    else:
        st = solver.download_state()
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in st.items()}
        host = {k: v.numpy() for k, v in pinned.items()}
        h2d = sum(v.nbytes for v in host.values())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.upload_state(pdf=host["pdf"], phi=host["phi"], cn_x=host["cn_x"], cn_y=host["cn_y"], cn_z=host["cn_z"], c_norm=host["c_norm"],
                            curv=host["curv"], f_convec=host.get("f_convec"), g_convec=host.get("g_convec"), phi_convec=host.get("phi_convec"))
        solver.run(nt, args.steps)
        mon2 = solver.monitor()
        solver.download_state_into(host)
        t_e2e = time.perf_counter() - t0
        e2e_mlups = n_site * args.steps / 1e6 / t_e2e
        d2h = h2d + 11 * 8 * S
    out = {
        "metric": "multiphase MLUPS (all lattice sites, reference definition src/main.cpp:270)", "value": mlups, "unit": "MLUPS",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if args.global_nx else "weak",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": (f"{dims} random sphere pack (radius 12, porosity ~0.4)" if args.geometry == "pack" else f"DIAGNOSTIC {dims} empty duct") +
                               f", {args.case}, velocity inlet + convective outlet, theta 45, {prec}",
                   "lattice": [NX, S, S], "fluid_nodes": n_fluid, "porosity": n_fluid / n_site, "parallelism": "1 GPU",
                   "l2": f"state {(38 * s_bytes * n_fluid) / 1e9:.2f}+ GB >> 126 MB L2 (inputs larger than L2, no flush needed)",
                   "fluid_mlups": n_fluid * args.steps / 1e6 / (ms * 1e-3), "geometry_preprocess_s": t_geo,
                   "saturation_full_domain": mon2["saturation_full_domain"], "activity_map": solver.activity},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "collide (k_collide_odd_ws / k_collide_even_tma, one launch per step)",
                     "bytes_model": "77*sizeof(real)*N_fluid per collide launch (38 PDF reads + 38 PDF writes + phi write)",
                     "bytes_per_launch": bytes_collide, "ms_per_launch": t_collide,
                     "odd": {"ms": t_odd, "achieved": bytes_collide / (t_odd * 1e-3) / 1e9, "frac": bytes_collide / (t_odd * 1e-3) / 1e9 / peak},
                     "even": {"ms": t_even, "achieved": bytes_collide / (t_even * 1e-3) / 1e9, "frac": bytes_collide / (t_even * 1e-3) / 1e9 / peak},
                     "whole_step": {"bytes_model": "78*sizeof(real)*N_fluid + N_site per step, all kernels (SURVEY.md 8d)", "bytes_per_step": bytes_step,
                                    "achieved": achieved_step, "frac": achieved_step / peak, "collide_share_of_step": t_collide / ms_step}},
        "e2e": {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "what": "upload_state from pinned host + K steps + monitor + download_state, through the C ABI"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(prec)
    solver.close()
    return out


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CUDA build (it has no CPU solver), unmodified sources, oracle/_ref/
# ----------------------------------------------------------------------------------------------------------
def run_reference(args) -> dict:
    import refcase as rc
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        sys.exit(0)
    S, prec = args.size, args.prec
    NX = args.global_nx or S
    dims = f"{S}^3" if NX == S else f"{NX}x{S}x{S}"
    exe = rc.ref_binary("gpu", prec)
    base = {"impl": "reference", "metric": "multiphase MLUPS (all lattice sites, reference definition src/main.cpp:270)", "unit": "MLUPS",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.global_nx else "weak",
            "vs_baseline": None, "dtype": prec, "data": "synthetic"}
    if not exe.exists():
        base["unavailable"] = f"{exe} missing (oracle/build_ref.sh needs /root/reference)"
        return base
    ctl = workload_control(NX, S, S, args.case)
    solid = workload_geometry(NX, S, S, seed=args.seed)
    solid_file = solid.copy()   # the reference applies the x/y walls itself (src/Misc.cpp:67-78); harmless to pre-apply
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        rc.write_case(td, ctl, solid_file)
        t0 = time.perf_counter()
        out = rc.run_ref("gpu", prec, td, td / "out", time=(args.warmup, args.steps), timeout=3000)
        wall = time.perf_counter() - t0
        tim = dict(l.split() for l in (out / "timing.txt").read_text().splitlines())
        host = dict(l.split() for l in (out / "host_timing.txt").read_text().splitlines())
    mlups = float(tim["mlups"])
    base.update({"value": mlups, "ms_per_step": float(tim["ms_per_step"]),
                 "config": {"workload": f"{dims} random sphere pack (radius 12, porosity ~0.4), {args.case}, velocity inlet + convective outlet, theta 45, {prec}",
                            "what": "reference CUDA kernels (unmodified sources, nvcc sm_100 -O3, block 128x1x1) via main_iteration_kernel_GPU()",
                            "host_setup_s": float(host["initialization_basic_multi_s"]), "wall_s": wall},
                 "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": 1, "kind": "reference",
                                  "sample": "the reference has no CPU solver: this is its own CUDA build on 1 B200; 1 host core drives it"},
                 "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000, help="timed steps (BASELINE configs 2/3: 2 k timed steps after 100 warm-up)")
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--prec", default="f64", choices=["f64", "f32"])
    ap.add_argument("--size", type=int, default=256, help="S: the lattice is S^3 per GPU (weak scaling)")
    ap.add_argument("--global-nx", type=int, default=0, help="strong scaling: the lattice is NX x S x S whatever the number of GPUs")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg: profiling runs, and --size 512 at N > 1 (the leg keeps two host copies of every rank's state, 36 GB per rank in f64)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo messages through peer memory (default) or NCCL send/recv")
    ap.add_argument("--case", default="drainage", choices=["drainage", "imbibition"], help="imbibition = BASELINE configs[4]")
    ap.add_argument("--seed", type=int, default=20240229, help="sphere-pack seed (configs[3]: 20240230)")
    ap.add_argument("--partition", default="equal", choices=["equal", "balanced"],
                    help="N > 1: equal-width x-slabs, or cuts that balance fluid nodes + halo cost per rank (slab.balanced_cuts)")
    ap.add_argument("--halo-cost", type=float, default=10.0,
                    help="--partition balanced: cost of one neighbour per face site, in fluid-node updates (f64 256^2 faces: ~104 us = 10.5)")
    ap.add_argument("--activity", type=int, default=None, choices=[0, 1],
                    help="gradient chain with the interface-activity map (kernels_activity.cuh); default: the library's (MFLBM_ACTIVITY)")
    ap.add_argument("--geometry", default="pack", choices=["pack", "open"], help="open = empty duct, diagnostic only (not the benchmark workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.activity is not None:
        os.environ["MFLBM_ACTIVITY"] = str(args.activity)   # read by the library when a solver is created
    if args.impl == "reference":
        out = run_reference(args)
    else:
        out = run_ours(args)
    if out is not None and int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
