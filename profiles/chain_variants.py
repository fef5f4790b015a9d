#!/usr/bin/env python
"""GPU-box measurement of the gradient-chain variants on the benchmark workload (one process, one geometry, every variant):
whole step from CUDA-graph replay (K steps after W) and the chain phase alone (CUDA events around step_phase(nt, 2)).

    python profiles/chain_variants.py [--size 256] [--steps 20] [--warmup 5] > gpurun_out/<tag>/chain_variants.json

Variants are environment switches the library reads when a solver is created: MFLBM_CHAIN (earlier runs of this
script also swept MFLBM_VARIANT digits that selected kernel versions since removed: profiles/r03a/c/e_chain_variants*.json).  Results of all variants are bit-identical (tests/test_gpu_parity.py); the saturation printed per line shows it.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "mf-lbm-cuda_b200"))

VARIANTS = [
    ("brick", {"MFLBM_CHAIN": "brick"}),
    ("csr", {"MFLBM_CHAIN": "csr"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precs", default="f64,f32")
    ap.add_argument("--only", default="", help="comma-separated variant names")
    a = ap.parse_args()
    import torch
    import bench
    import mflbm
    S = a.size
    ctl = bench.workload_control(S, S, S, "drainage")
    solid = bench.workload_geometry(S, S, S)
    stream = torch.cuda.Stream()
    out = []
    for prec in a.precs.split(","):
        W = bench.inlet_profile(ctl, prec)
        for name, env in VARIANTS:
            if a.only and name not in a.only.split(","):
                continue
            os.environ.update(env)
            t0 = time.perf_counter()
            solver = mflbm.Solver(mflbm.derive_params(ctl, prec), prec, device=0, stream=stream.cuda_stream)
            solver.preprocess_geometry(solid)
            solver.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=W)
            t_setup = time.perf_counter() - t0
            solver.run(1, a.warmup)
            solver.sync()
            nt = 1 + a.warmup
            reps = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                solver.run(nt, a.steps)
                e1.record(stream)
                torch.cuda.synchronize()
                reps.append(e0.elapsed_time(e1) / a.steps)
                nt += a.steps
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for p, q in ev:
                solver.step_phase(nt, 0)
                solver.step_phase(nt, 1)
                p.record(stream)
                solver.step_phase(nt, 2)
                q.record(stream)
                nt += 1
            torch.cuda.synchronize()
            chain_us = float(np.median([p.elapsed_time(q) for p, q in ev])) * 1e3
            mon = solver.monitor()
            rec = {"prec": prec, "variant": name, "env": env, "ms_per_step": reps, "ms_per_step_best": min(reps), "chain_phase_us_median": chain_us,
                   "mlups_best": S ** 3 / 1e6 / (min(reps) * 1e-3), "chain_bricks": solver.chain_bricks() if solver.chain != "list" else None,
                   "saturation_full_domain": mon["saturation_full_domain"], "nan": mon["nan_detected"], "setup_s": t_setup, "steps_done": nt - 1}
            print(json.dumps(rec), flush=True)
            out.append(rec)
            solver.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
