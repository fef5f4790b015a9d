"""Generates tests/golden/ref_gpu_<case>_<prec>.npz from the REAL reference GPU kernels
(oracle/_ref/ref_gpu_*, the unmodified /root/reference sources built for sm_100 by oracle/build_ref.sh).
Needs a GPU:   gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'   then copy the npz files here."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent / "oracle"))
import common  # noqa: E402
import refgpu  # noqa: E402

outdir = Path(sys.argv[1]) if len(sys.argv) > 1 else HERE
outdir.mkdir(parents=True, exist_ok=True)
for name in sorted(common.CASES):
    for prec in ("f64", "f32"):
        meta, geom, states, _ = refgpu.run_reference_gpu(name, prec)
        blob = {}
        for step in refgpu.STEPS:
            for k, v in refgpu.fingerprint(name, states[step]).items():
                blob[f"s{step}_{k}"] = v
        np.savez_compressed(outdir / f"ref_gpu_{name}_{prec}.npz", **blob)
        print(name, prec, "ok", flush=True)
