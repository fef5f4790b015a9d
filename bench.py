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
        if self.index < 0:   # N > 1: rank 0 samples its GPU; 8 nvidia-smi pollers at once stall the driver for everybody
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
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
            "kind": "port", "sample": f"{n}^3 sphere pack (NOT the 256^3 workload: a bounded sample of the same generator and case), {steps} steps, C oracle with OpenMP (the reference has no CPU solver path)"}


def host_mem_available() -> int:
    try:
        for ln in Path("/proc/meminfo").read_text().splitlines():
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return 0


def roofline_record(prec: str, n_fluid: int, n_site: int, ms_step: float, t_odd: float | None, t_even: float | None, traffic=None) -> dict:
    """Both figures at every N (per GPU): the dominant launch (collide, 77*s*N_fluid bytes, CUDA events around the launch) and the
    whole step (78*s*N_fluid + N_site bytes, SURVEY.md 8d, all kernels and - for slabs - the halo messages)."""
    s_bytes = 8 if prec == "f64" else 4
    peak, peak_src = hbm_peak()
    bytes_step = 78 * s_bytes * n_fluid + n_site
    achieved_step = bytes_step / (ms_step * 1e-3) / 1e9
    bytes_collide = 77 * s_bytes * n_fluid
    rec = {"bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": traffic, "peak_source": peak_src,
           "kernel": "collide (k_collide_odd_ws / k_collide_even_tma, one launch per step)",
           "bytes_model": "77*sizeof(real)*N_fluid per collide launch (38 PDF reads + 38 PDF writes + phi write)",
           "bytes_per_launch": bytes_collide,
           "whole_step": {"bytes_model": "78*sizeof(real)*N_fluid + N_site per step, all kernels (SURVEY.md 8d)", "bytes_per_step": bytes_step,
                          "achieved": achieved_step, "frac": achieved_step / peak}}
    if t_odd and t_even:
        t_collide = 0.5 * (t_odd + t_even)
        achieved = bytes_collide / (t_collide * 1e-3) / 1e9
        rec.update({"achieved": achieved, "frac": achieved / peak, "ms_per_launch": t_collide,
                    "odd": {"ms": t_odd, "achieved": bytes_collide / (t_odd * 1e-3) / 1e9, "frac": bytes_collide / (t_odd * 1e-3) / 1e9 / peak},
                    "even": {"ms": t_even, "achieved": bytes_collide / (t_even * 1e-3) / 1e9, "frac": bytes_collide / (t_even * 1e-3) / 1e9 / peak}})
        rec["whole_step"]["collide_share_of_step"] = t_collide / ms_step
    else:
        rec.update({"achieved": achieved_step, "frac": achieved_step / peak})
    return rec


def time_collide_launches(solver, stream, nt: int, steps: int, step_fn=None):
    """CUDA events on the launching stream around the collide launch of 2*kpairs single steps -> (ms odd, ms even, next nt).
    step_fn(nt, phase): how the caller sequences a step (slabs interleave their halo exchanges)."""
    import torch
    kpairs = max(2, min(steps // 2, 25))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * kpairs)]
    first = nt
    for a, b in ev:
        if step_fn is None:
            a.record(stream); solver.step_phase(nt, 0); b.record(stream)
            solver.step_phase(nt, 1); solver.step_phase(nt, 2)
        else:
            step_fn(nt, a, b)
        nt += 1
    torch.cuda.synchronize()
    t_all = [a.elapsed_time(b) for a, b in ev]
    first_odd = first % 2 == 1
    t_odd = float(np.mean(t_all[0::2] if first_odd else t_all[1::2]))
    t_even = float(np.mean(t_all[1::2] if first_odd else t_all[0::2]))
    return t_odd, t_even, nt


def measure_single(args, prec: str, local: int, want_e2e: bool) -> dict:
    """one lattice resident on one GPU: device-timed steps, per-launch timing of the collide kernels, optional end-to-end leg"""
    import torch
    import mflbm
    S = args.size
    NX = args.global_nx or S
    ctl = workload_control(NX, S, S, args.case)
    solid = workload_geometry(NX, S, S, seed=args.seed, kind=args.geometry)
    stream = torch.cuda.Stream()
    solver = mflbm.Solver(mflbm.derive_params(ctl, prec), prec, device=local, stream=stream.cuda_stream)
    t0 = time.perf_counter()
    solver.preprocess_geometry(solid)
    t_geo = time.perf_counter() - t0
    del solid
    W = inlet_profile(ctl, prec)
    solver.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=W)
    n_fluid, n_site = solver.num_fluid_nodes, NX * S * S
    # ---- device-resident timing (the clock sampler covers warm-up + timed region + per-launch timing: K = 20 steps last 20 ms) ----
    with ClockSampler(local) as clk:
        solver.run(1, args.warmup)
        solver.sync()
        l0 = solver.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nt = 1 + args.warmup
        torch.cuda.synchronize()
        e0.record(stream)
        solver.run(nt, args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = solver.kernel_launches - l0
        nt += args.steps
        t_odd, t_even, nt = time_collide_launches(solver, stream, nt, args.steps)
        if ms < 500.0:   # keep the GPU busy long enough for a few 20 ms clock samples under load (not timed)
            solver.run(nt, 2 * max(1, int(250.0 / max(ms / args.steps, 1e-3)) // 2)); nt += 2 * max(1, int(250.0 / max(ms / args.steps, 1e-3)) // 2)
            solver.sync()
    mon = solver.monitor()
    assert mon["nan_detected"] == 0, "simulation produced non-finite values"
    chain_bricks = solver.chain_bricks() if solver.chain != "list" else None
    ms_step = ms / args.steps
    traffic = None
    tf = REPO / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(f"{prec}_{S}", None) if NX == S else None
        except Exception:
            traffic = None
    rec = {"prec": prec, "lattice": [NX, S, S], "n_fluid": n_fluid, "n_site": n_site, "ms": ms, "ms_per_step": ms_step,
           "mlups": n_site * args.steps / 1e6 / (ms * 1e-3), "launches": int(launches), "t_geo": t_geo, "clocks": clk.summary(),
           "saturation_full_domain": mon["saturation_full_domain"], "chain": solver.chain, "chain_bricks": chain_bricks,
           "roofline": roofline_record(prec, n_fluid, n_site, ms_step, t_odd, t_even, traffic), "e2e": None}
    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    if want_e2e:
        s_bytes = 8 if prec == "f64" else 4
        need = s_bytes * (38 + 8) * (NX + 8) * (S + 8) * (S + 8)
        avail = host_mem_available()
        if avail and need * 1.25 > avail:
            rec["e2e"] = {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                          "skipped": f"the host copy of the state needs {need / 1e9:.1f} GB of pinned memory, {avail / 1e9:.1f} GB available"}
        else:
            host = {k: torch.empty(v, dtype=torch.float64 if prec == "f64" else torch.float32).pin_memory().numpy()
                    for k, v in solver.state_shapes().items()}
            solver.download_state_into(host)   # the caller's arrays now hold the state, as after a timer step of the reference
            h2d = sum(v.nbytes for k, v in host.items() if k != "curv")   # curv is a function of cn_*: accepted, never copied (mflbm.h)
            d2h = sum(v.nbytes for v in host.values()) + 16 * 8 * S       # + the monitor's 16 doubles per z slice
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            solver.upload_state(pdf=host["pdf"], phi=host["phi"], cn_x=host["cn_x"], cn_y=host["cn_y"], cn_z=host["cn_z"], c_norm=host["c_norm"],
                                curv=host["curv"], f_convec=host.get("f_convec"), g_convec=host.get("g_convec"), phi_convec=host.get("phi_convec"))
            t1 = time.perf_counter()
            solver.run(nt, args.steps)
            mon2 = solver.monitor()
            t2 = time.perf_counter()
            solver.download_state_into(host)
            t_e2e = time.perf_counter() - t0
            rec["saturation_full_domain"] = mon2["saturation_full_domain"]
            rec["e2e"] = {"value": n_site * args.steps / 1e6 / t_e2e, "unit": "MLUPS", "h2d_bytes_per_step": h2d / args.steps,
                          "d2h_bytes_per_step": d2h / args.steps, "seconds": t_e2e, "upload_seconds": t1 - t0, "steps_and_monitor_seconds": t2 - t1,
                          "download_seconds": t_e2e - (t2 - t0),
                          "what": "upload_state from pinned host arrays (reference layouts) + K steps + monitor + download_state, through the C ABI"}
            del host
    solver.close()
    return rec


def run_ours(args) -> dict:
    import torch
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        try:
            return bench_slabs(args, rank, world, local)
        finally:
            dist.destroy_process_group()
    S, prec = args.size, args.prec
    NX = args.global_nx or S
    dims = f"{S}^3" if NX == S else f"{NX}x{S}x{S}"
    r = measure_single(args, prec, local, want_e2e=not args.no_e2e)
    s_bytes = 8 if prec == "f64" else 4
    out = {
        "metric": "multiphase MLUPS (all lattice sites, reference definition src/main.cpp:270)", "value": r["mlups"], "unit": "MLUPS",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.global_nx else "weak",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": (f"{dims} random sphere pack (radius 12, porosity ~0.4)" if args.geometry == "pack" else f"DIAGNOSTIC {dims} empty duct") +
                               f", {args.case}, velocity inlet + convective outlet, theta 45, {prec}",
                   "lattice": [NX, S, S], "fluid_nodes": r["n_fluid"], "porosity": r["n_fluid"] / r["n_site"], "parallelism": "1 GPU",
                   "l2": f"state {(38 * s_bytes * r['n_fluid']) / 1e9:.2f}+ GB >> 126 MB L2 (inputs larger than L2, no flush needed)",
                   "fluid_mlups": r["n_fluid"] * args.steps / 1e6 / (r["ms"] * 1e-3), "geometry_preprocess_s": r["t_geo"],
                   "saturation_full_domain": r["saturation_full_domain"], "gradient_chain": r["chain"],
                   "chain_bricks_processed_of_total": r["chain_bricks"]},
        "roofline": r["roofline"],
        "e2e": r["e2e"] if r["e2e"] is not None else {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "--no-e2e"},
        "gpu_launches": r["launches"],
        "clocks": r["clocks"],
    }
    if not args.no_fp32 and prec == "f64":
        # the metric is quoted for both precisions (BASELINE configs[1] and [2]): the single-precision build path on the same
        # workload, device-resident timing only
        q = measure_single(args, "f32", local, want_e2e=False)
        out["fp32"] = {"value": q["mlups"], "unit": "MLUPS", "ms_per_step": q["ms_per_step"], "dtype": "f32", "gpu_launches": q["launches"],
                       "fluid_mlups": q["n_fluid"] * args.steps / 1e6 / (q["ms"] * 1e-3), "roofline": q["roofline"],
                       "saturation_full_domain": q["saturation_full_domain"], "clocks": q["clocks"]}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(prec)
    return out


# ----------------------------------------------------------------------------------------------------------
# N > 1 (under torchrun, one rank per GPU): the lattice is cut into x-slabs
# ----------------------------------------------------------------------------------------------------------
def verify_slabs(prec: str, rank: int, world: int, local: int, halo: str) -> dict:
    """Before anything is timed: a small sphere-pack lattice stepped (a) as `world` x-slabs over the transport the benchmark uses
    and (b) as one domain on this rank's GPU; every rank compares its own columns of every state array bit for bit."""
    import torch
    import torch.distributed as dist
    import mflbm
    from mflbm import slab as sm
    S, nsteps = 48, 41
    nxg = 24 * world
    ctl = workload_control(nxg, S, 2 * S, "drainage")
    solid = workload_geometry(nxg, S, 2 * S, seed=7)
    W = inlet_profile(ctl, prec)
    P = mflbm.derive_params(ctl, prec)
    rng = sm.partition(nxg, world, rank)
    stream = torch.cuda.Stream(device=local)
    bad = []
    with torch.cuda.stream(stream):
        cs = sm.CudaSlab(P, prec, rng, local, stream=stream)
        cs.solver.preprocess_geometry(solid)
        cs.solver.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=np.ascontiguousarray(W[:, rng.x0 - 1:rng.x1 + 2]))
        if halo == "p2p":
            cs.connect_p2p(dist)
        st = sm.SlabStepper(cs, rng)
        st.run(1, nsteps)
        st.settle()
        mine = cs.solver.download_state()
        ref = mflbm.Solver(P, prec, device=local, stream=stream.cuda_stream)
        ref.preprocess_geometry(solid)
        ref.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=W)
        ref.run(1, nsteps)
        want = ref.download_state()
        for k, g in (("pdf", 1), ("phi", 4), ("cn_x", 2), ("cn_y", 2), ("cn_z", 2), ("c_norm", 2), ("curv", 1)):
            ls, gs = sm._owned_columns(rng, g)
            if not np.array_equal(mine[k][..., ls], want[k][..., gs]):
                bad.append(k)
        transport = "p2p" if getattr(cs, "p2p", False) else "nccl"
        ref.close(); cs.close()
    flag = torch.tensor([0 if bad else 1], dtype=torch.int32, device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if bad:
        print(f"[rank {rank}] slab parity check FAILED for {bad}", file=sys.stderr, flush=True)
    return {"bit_identical": bool(int(flag[0])), "lattice": [nxg, S, 2 * S], "steps": nsteps, "prec": prec, "transport": transport,
            "what": "x-slab run vs single-domain run of the same library on every rank's own columns: pdf, phi, cn_*, c_norm, curv"}


def bench_slabs(args, rank: int, world: int, local: int) -> dict | None:
    import torch
    import torch.distributed as dist
    import mflbm
    from mflbm import slab as sm
    S, prec = args.size, args.prec
    strong = int(args.global_nx or 0)
    nxg = strong if strong else S * world      # strong scaling: fixed NX x S x S lattice; weak: S^3 per GPU
    case, seed = args.case, int(args.seed)
    parity = None if args.no_verify else verify_slabs(prec, rank, world, local, args.halo)
    if parity is not None and not parity["bit_identical"]:
        raise SystemExit("x-slab parity check failed: refusing to time a wrong result")
    ctl = workload_control(nxg, S, S, case)
    rng = sm.partition(nxg, world, rank)
    if args.partition == "balanced" and world > 1:
        # every rank counts the fluid nodes of its equal-width columns, the counts are gathered, and all ranks derive the same
        # cost-balanced cuts (fluid nodes + a per-neighbour halo cost expressed in fluid-node updates per face site)
        w = workload_geometry_window(nxg, S, S, rng.x0, rng.x1, seed=seed, kind=args.geometry)
        mine = (w[:, :, rng.x0 - 1:rng.x1] == 0).sum(axis=(0, 1)).astype(np.float64)
        del w
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        rng = sm.partition_balanced(np.concatenate(parts), world, rank, side_cost=float(args.halo_cost) * S * S)
    params = mflbm.derive_params(ctl, prec)
    stream = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        slab = sm.CudaSlab(params, prec, rng, local, stream=stream)
        t0 = time.perf_counter()
        solid = workload_geometry_window(nxg, S, S, rng.x0 - 12, rng.x1 + 12, seed=seed, kind=args.geometry)
        slab.solver.preprocess_geometry(solid)
        t_geo = time.perf_counter() - t0
        del solid
        W = inlet_profile(ctl, prec)
        W_local = np.ascontiguousarray(W[:, rng.x0 - 1:rng.x1 + 2])   # local columns 0..nx+1 of the global profile
        slab.solver.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=W_local)
        if args.halo == "p2p":
            slab.connect_p2p(dist)
        stepper = sm.SlabStepper(slab, rng)
        with ClockSampler(local if rank == 0 else -1) as clk:
            stepper.run(1, args.warmup)
            nt = 1 + args.warmup
            torch.cuda.synchronize()
            dist.barrier()
            l0 = slab.solver.kernel_launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            stepper.run(nt, args.steps)
            e1.record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            nt += args.steps
            ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            ms = float(ms_t[0])
            launches = slab.solver.kernel_launches - l0
            # the dominant launch, timed on every rank with CUDA events around the collide phase of single (ungraphed) steps
            def one(nt_, a, b):
                a.record(stream); slab.step_phase(nt_, 0); b.record(stream)
                stepper.exchange(sm.KIND_PDF_ODD if nt_ % 2 else sm.KIND_PDF_EVEN)
                slab.step_phase(nt_, 1)
                stepper.exchange(sm.KIND_PHI)
                slab.step_phase(nt_, 2)
            t_odd, t_even, nt = time_collide_launches(slab.solver, stream, nt, args.steps, step_fn=one)
            # where a slab's step goes, phase by phase (serial schedule, CUDA events on the launching stream, mean over 20 steps;
            # the timed run above overlaps the PDF message with the interior collide tiles): diagnostic, reported as config.phase_ms
            names = ("collide", "halo_pdf", "boundaries", "halo_phi", "chain")
            evs = []
            for _ in range(20):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                e[0].record(stream); slab.step_phase(nt, 0)
                e[1].record(stream); stepper.exchange(sm.KIND_PDF_ODD if nt % 2 else sm.KIND_PDF_EVEN)
                e[2].record(stream); slab.step_phase(nt, 1)
                e[3].record(stream); stepper.exchange(sm.KIND_PHI)
                e[4].record(stream); slab.step_phase(nt, 2)
                e[5].record(stream)
                evs.append(e); nt += 1
            torch.cuda.synchronize()
            ph = torch.tensor([float(np.mean([e[i].elapsed_time(e[i + 1]) for e in evs])) for i in range(5)], dtype=torch.float64, device=f"cuda:{local}")
            ph_max = ph.clone()
            dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
            phase_ms = {n: {"rank0": float(a), "max": float(b)} for n, a, b in zip(names, ph.tolist(), ph_max.tolist())}
            stepper.last_ntime = nt - 1
            tk = torch.tensor([t_odd, t_even], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            t_odd, t_even = float(tk[0]), float(tk[1])
        nf = torch.tensor([slab.solver.num_fluid_nodes], dtype=torch.float64, device=f"cuda:{local}")
        nf_max = nf.clone()
        dist.all_reduce(nf, op=dist.ReduceOp.SUM)
        dist.all_reduce(nf_max, op=dist.ReduceOp.MAX)
        n_fluid = int(nf[0])
        mon = sm.reduce_monitor(slab.solver.monitor(), rng, params, dist, device=f"cuda:{local}")
        # ---- end to end through the C ABI: pinned host state -> device, K steps, monitor + state back ----------
        e2e = {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "skipped": "--no-e2e"}
        if not args.no_e2e:
            s_bytes = 8 if prec == "f64" else 4
            need = s_bytes * (38 + 8) * (rng.nx_local + 8) * (S + 8) * (S + 8)
            avail = host_mem_available()
            fits = torch.tensor([1 if (not avail or need * world * 1.25 < avail) else 0], dtype=torch.int32, device=f"cuda:{local}")
            dist.all_reduce(fits, op=dist.ReduceOp.MIN)
            if not int(fits[0]):
                e2e["skipped"] = f"the host copies of the state need {need * world / 1e9:.0f} GB of pinned memory on this box, {avail / 1e9:.0f} GB available"
            else:
                host = {k: torch.empty(v, dtype=torch.float64 if prec == "f64" else torch.float32).pin_memory().numpy()
                        for k, v in slab.solver.state_shapes().items()}
                slab.solver.download_state_into(host)
                h2d = sum(v.nbytes for k, v in host.items() if k != "curv")
                d2h = sum(v.nbytes for v in host.values())
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                slab.solver.upload_state(**{k: host[k] for k in ("pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv")},
                                         f_convec=host.get("f_convec"), g_convec=host.get("g_convec"), phi_convec=host.get("phi_convec"))
                stepper.run(nt, args.steps)
                sm.reduce_monitor(slab.solver.monitor(), rng, params, dist, device=f"cuda:{local}")
                slab.solver.download_state_into(host)
                torch.cuda.synchronize()
                te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
                tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local}")
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
                e2e = {"value": nxg * S * S * args.steps / 1e6 / float(te[0]), "unit": "MLUPS", "h2d_bytes_per_step": float(tot[0]) / args.steps,
                       "d2h_bytes_per_step": float(tot[1]) / args.steps, "seconds": float(te[0]),
                       "what": "per rank: upload_state from pinned host + K steps + monitor (all_reduce) + download_state, through the C ABI"}
                del host
    n_site = nxg * S * S
    ms_step = ms / args.steps
    out = None
    if rank == 0:
        assert mon["nan_detected"] == 0, "simulation produced non-finite values"
        # per-GPU roofline of the slowest rank: the slab with the most fluid nodes
        roof = roofline_record(prec, int(nf_max[0]), n_site // world, ms_step, t_odd, t_even)
        roof["per"] = "GPU (the slab with the most fluid nodes; launch times are the max over ranks)"
        out = {
            "metric": "multiphase MLUPS (all lattice sites, reference definition src/main.cpp:270)", "value": n_site * args.steps / 1e6 / (ms * 1e-3),
            "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": f"{nxg}x{S}x{S} random sphere pack (radius 12, porosity ~0.4)" + ("" if strong else f" = {S}^3 per GPU") + f", {case}, velocity inlet + convective outlet, theta 45, {prec}",
                       "lattice": [nxg, S, S], "fluid_nodes": n_fluid, "porosity": n_fluid / n_site, "parallelism": f"{world} x-slabs, halos " + ("pushed into peer memory over NVLink (CUDA IPC), arrival flags, no collective" if getattr(slab, "p2p", False) else "NCCL send/recv"),
                       "l2": "state per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "fluid_mlups": n_fluid * args.steps / 1e6 / (ms * 1e-3), "geometry_preprocess_s": t_geo,
                       "saturation_full_domain": mon["saturation_full_domain"], "halo_exchanges_per_step": 2,
                       "partition": args.partition, "slab_width_rank0": rng.nx_local, "parity_checked": parity, "phase_ms": phase_ms},
            "roofline": roof,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "clocks": clk.summary(),
        }
    slab.close()
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
        block = [int(b) for b in args.ref_block.split(",")] if args.ref_block else None
        out = rc.run_ref("gpu", prec, td, td / "out", time=(args.warmup, args.steps), timeout=3000, e2e=0 if args.no_e2e else args.steps, block=block)
        wall = time.perf_counter() - t0
        tim = dict(l.split() for l in (out / "timing.txt").read_text().splitlines())
        host = dict(l.split() for l in (out / "host_timing.txt").read_text().splitlines())
        e2e = dict(l.split() for l in (out / "e2e.txt").read_text().splitlines()) if (out / "e2e.txt").exists() else None
    mlups = float(tim["mlups"])
    base.update({"value": mlups, "ms_per_step": float(tim["ms_per_step"]),
                 "config": {"workload": f"{dims} random sphere pack (radius 12, porosity ~0.4), {args.case}, velocity inlet + convective outlet, theta 45, {prec}",
                            "what": f"reference CUDA kernels (unmodified sources, nvcc sm_100 -O3, block {args.ref_block or '128,1,1 (shipped)'}) via main_iteration_kernel_GPU()",
                            "host_setup_s": float(host["initialization_basic_multi_s"]), "wall_s": wall,
                            "note": None if args.gpus <= 1 else f"the reference is single-GPU (README.md:119): at --gpus {args.gpus} this is one GPU's share of our arm's lattice"
                                    + (" (weak scaling)" if not args.global_nx else " - here the whole strong-scaling lattice on one GPU")},
                 "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": 1, "kind": "reference",
                                  "sample": "the reference has no CPU solver: this is its own CUDA build on 1 B200; 1 host core drives it"},
                 "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    if e2e is not None:
        # the same post-condition as our e2e leg, through the reference's own transfer code: MemAllocate_multi_GPU(1)
        # (cudaMalloc + H2D of the state from its calloc'ed host arrays, src/Init_multiphase_GPU.cu:72-97), K steps, the D2H
        # block of main_iteration_kernel_GPU() on the last one (src/main_iteration_GPU.cu:2058-2076); see oracle/ref_harness.cpp
        base["e2e"] = {"value": float(e2e["mlups"]), "unit": "MLUPS", "h2d_bytes_per_step": int(e2e["h2d_bytes"]) / args.steps,
                       "d2h_bytes_per_step": int(e2e["d2h_bytes"]) / args.steps, "seconds": float(e2e["seconds"]),
                       "upload_seconds": float(e2e["upload_seconds"]), "host_monitor_seconds_not_counted": float(e2e["host_monitor_seconds"]),
                       "what": "MemAllocate_multi_GPU(1) state upload + K steps + the reference's D2H block on the last step (host arrays pageable, as the reference allocates them)"}
    return base


def resolve_workload(args, world: int = 1) -> None:
    """fill in the workload defaults: N = 1 -> BASELINE configs[1] (256^3 drainage; the metric is quoted on it), N > 1 -> configs[4]
    (512^3 per GPU, imbibition, weak scaling); --config strong -> configs[3] (1024 x 512 x 512 drainage with the next seed)"""
    args.warmup = max(args.warmup, 3)
    if world > 1 and world != args.gpus:
        args.gpus = world
    multi = args.gpus > 1
    if args.config == "strong":
        args.global_nx, args.size, args.case, args.seed = 1024, args.size or 512, args.case or "drainage", args.seed + 1
    elif args.config == "weak":
        args.size, args.case = args.size or 512, args.case or "imbibition"
    if not args.size:
        args.size = 512 if multi else 256
    if not args.case:
        args.case = "imbibition" if multi else "drainage"


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000, help="timed steps (BASELINE configs 2/3: 2 k timed steps after 100 warm-up)")
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--prec", default="f64", choices=["f64", "f32"])
    ap.add_argument("--size", type=int, default=0, help="S: the lattice is S^3 per GPU (weak scaling); default 256 at N = 1 (BASELINE configs[1]), 512 at N > 1 (configs[4])")
    ap.add_argument("--config", default="", choices=["", "weak", "strong"],
                    help="weak = BASELINE configs[4] (512^3 per GPU, imbibition; the default at N > 1), strong = configs[3] (1024x512x512 drainage, seed + 1, whatever N)")
    ap.add_argument("--no-fp32", action="store_true", help="N = 1: skip the single-precision sub-record")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the slab-vs-single-domain parity check that precedes the timing")
    ap.add_argument("--global-nx", type=int, default=0, help="strong scaling: the lattice is NX x S x S whatever the number of GPUs")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg: profiling runs, and --size 512 at N > 1 (the leg keeps two host copies of every rank's state, 36 GB per rank in f64)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo messages through peer memory (default) or NCCL send/recv")
    ap.add_argument("--case", default="", choices=["", "drainage", "imbibition"], help="default: drainage at N = 1 (configs[1]), imbibition at N > 1 (configs[4])")
    ap.add_argument("--seed", type=int, default=20240229, help="sphere-pack seed (configs[3]: 20240230)")
    ap.add_argument("--partition", default="equal", choices=["equal", "balanced"],
                    help="N > 1: equal-width x-slabs, or cuts that balance fluid nodes + halo cost per rank (slab.balanced_cuts)")
    ap.add_argument("--halo-cost", type=float, default=10.0,
                    help="--partition balanced: cost of one neighbour per face site, in fluid-node updates (f64 256^2 faces: ~104 us = 10.5)")
    ap.add_argument("--chain", default=None, choices=["brick", "csr", "list"],
                    help="gradient chain: brick by brick where an interface can be (kernels_chain.cuh; csr = per-brick site lists built once per geometry) or the four list kernels (kernels_step.cuh)")
    ap.add_argument("--geometry", default="pack", choices=["pack", "open"], help="open = empty duct, diagnostic only (not the benchmark workload)")
    ap.add_argument("--ref-block", default="", help="--impl reference: block_Threads_X,Y,Z override (the shipped control file has 128,1,1)")
    return ap


def main():
    args = build_parser().parse_args()
    resolve_workload(args, int(os.environ.get("WORLD_SIZE", 1)))
    if args.chain is not None:
        os.environ["MFLBM_CHAIN"] = args.chain   # read by the library when a solver is created
    if args.impl == "reference":
        out = run_reference(args)
    else:
        out = run_ours(args)
    if out is not None and int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
