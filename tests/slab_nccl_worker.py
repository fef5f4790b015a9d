"""torchrun worker of tests/test_gpu_slab.py::test_two_gpus_nccl_equal_single_domain (also usable by hand:
python -m torch.distributed.run --nproc-per-node 2 tests/slab_nccl_worker.py /tmp/out).  One rank per GPU, NCCL."""
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for p in (REPO / "tests", REPO / "oracle", REPO / "mf-lbm-cuda_b200"):
    sys.path.insert(0, str(p))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import common  # noqa: E402
import mflbm  # noqa: E402
from mflbm import slab  # noqa: E402
from test_gpu_slab import owned  # noqa: E402


def main():
    out = Path(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, prec, nsteps, halo in (("pack_velocity", "f64", 20, "nccl"), ("tube_pressure", "f32", 21, "nccl"), ("periodic_drop", "f64", 10, "nccl"),
                                     ("pack_velocity", "f64", 20, "p2p"), ("tube_pressure", "f32", 21, "p2p"), ("periodic_drop", "f32", 10, "p2p")):
        o, ctl, solid = common.make_oracle(name, prec)
        P = mflbm.derive_params(ctl, prec)
        interior = (o.arr("walls_global") != 0).astype(np.int8)
        opt, z0, W = ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], o.arr("W_in")
        rng = slab.partition(o.nx, world, rank)
        stream = torch.cuda.Stream(device=local)
        with torch.cuda.stream(stream):
            cs = slab.CudaSlab(P, prec, rng, local, stream=stream)
            cs.solver.preprocess_geometry(interior)
            cs.solver.init_state(opt, z0, W_in=np.ascontiguousarray(W[:, rng.x0 - 1:rng.x1 + 2]))
            if halo == "p2p":   # CUDA IPC handles of the receive buffers / arrival flags, halo messages over NVLink stores
                cs.connect_p2p(dist)
            st = slab.SlabStepper(cs, rng)
            st.run(1, nsteps)
            st.settle()
            mine = owned(cs.solver.download_state(), rng, o.nx)
            mon = slab.reduce_monitor(cs.solver.monitor(), rng, P, dist, device=f"cuda:{local}") if nsteps % 2 == 0 else None
        np.savez(out / f"{name}_{rank}.npz", **mine)
        dist.barrier()
        if rank == 0:
            ref = mflbm.Solver(P, prec, device=local)
            ref.preprocess_geometry(interior)
            ref.init_state(opt, z0, W_in=W)
            ref.run(1, nsteps)
            want = ref.download_state()
            parts = [np.load(out / f"{name}_{r}.npz") for r in range(world)]
            for k in ("pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"):
                got = np.concatenate([p[k] for p in parts], axis=-1)
                if got.shape != want[k].shape or not np.array_equal(got, want[k]):
                    ok = False
                    print(f"MISMATCH {name} {prec} {halo} {k}", flush=True)
            if mon is not None:
                m = ref.monitor()
                for k in ("saturation", "saturation_full_domain", "ca"):
                    if abs(mon[k] - m[k]) > 1e-9 * max(1.0, abs(m[k])):
                        ok = False
                        print(f"MONITOR MISMATCH {name} {k} {mon[k]} {m[k]}", flush=True)
            ref.close()
        cs.solver.close()
        dist.barrier()
    if rank == 0 and ok:
        print("SLAB_NCCL_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
