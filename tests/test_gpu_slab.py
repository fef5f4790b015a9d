"""x-slab path on the GPU: CUDA halo pack/unpack kernels, extended kernel ranges and the exchange schedule.

  * one GPU: the lattice is cut into 2 or 3 slabs that live on the same device; the messages are plain device copies
    between the slabs' halo buffers.  The union of the slabs must equal the single-domain CUDA run bit for bit.
  * two GPUs (skipped when fewer are visible): the same through torch.distributed/NCCL, launched with torchrun.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


class LocalChain:
    """all slabs in one process: exchange = pack everywhere, copy send -> neighbour's recv, unpack everywhere"""

    def __init__(self, slabs):
        self.slabs = slabs

    def exchange(self, kind):
        from mflbm.slab import LEFT, RIGHT
        for s in self.slabs:
            s.halo_pack(kind)
        for r, s in enumerate(self.slabs):
            if s.rng.has_left:
                s.halo_tensors(kind, LEFT)[1].copy_(self.slabs[r - 1].halo_tensors(kind, RIGHT)[0])
            if s.rng.has_right:
                s.halo_tensors(kind, RIGHT)[1].copy_(self.slabs[r + 1].halo_tensors(kind, LEFT)[0])
        for s in self.slabs:
            s.halo_unpack(kind)

    def step(self, ntime):
        for s in self.slabs:
            s.step_phase(ntime, 0)
        self.exchange(1 if ntime % 2 else 0)
        for s in self.slabs:
            s.step_phase(ntime, 1)
        self.exchange(2)
        for s in self.slabs:
            s.step_phase(ntime, 2)


def owned(st, rng, nx_global):
    """cut the columns a slab owns (plus the global ghost columns at the chain ends) out of its local arrays"""
    out = {}
    for k, g in (("pdf", 1), ("phi", 4), ("cn_x", 2), ("cn_y", 2), ("cn_z", 2), ("c_norm", 2), ("curv", 1)):
        lo = g if rng.has_left else 0                       # local column 1 sits at index g
        hi = g + rng.nx_local if rng.has_right else None
        out[k] = st[k][..., lo:hi]
    return out


@pytest.mark.parametrize("name,prec,world,nsteps", [
    ("tube_pressure", "f64", 2, 12),
    ("pack_velocity", "f64", 3, 12),
    ("periodic_drop", "f64", 2, 11),
    ("imbibition_plate2", "f32", 2, 12),
    ("rect_quirk", "f32", 4, 7),
])
def test_cuda_slabs_on_one_gpu_equal_single_domain(gpu_lib, name, prec, world, nsteps):
    import torch
    import mflbm
    from mflbm import slab
    o, ctl, solid = common.make_oracle(name, prec)
    P = mflbm.derive_params(ctl, prec)
    interior = (o.arr("walls_global") != 0).astype(np.int8)
    opt, z0 = ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"]
    W = o.arr("W_in")
    ref = mflbm.Solver(P, prec)
    ref.preprocess_geometry(interior)
    ref.init_state(opt, z0, W_in=W)
    for n in range(nsteps):
        ref.step(1 + n)
    want = ref.download_state()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        slabs = []
        for r in range(world):
            rng = slab.partition(o.nx, world, r)
            cs = slab.CudaSlab(P, prec, rng, 0, stream=stream)
            cs.solver.preprocess_geometry(interior)
            cs.solver.init_state(opt, z0, W_in=np.ascontiguousarray(W[:, rng.x0 - 1:rng.x1 + 2]))
            slabs.append(cs)
        assert sum(s.solver.num_fluid_nodes for s in slabs) == ref.num_fluid_nodes
        chain = LocalChain(slabs)
        for n in range(nsteps):
            chain.step(1 + n)
        if nsteps % 2 == 0:
            chain.exchange(1)      # SlabStepper.settle(): owner columns complete before gathering
        parts = [owned(s.solver.download_state(), s.rng, o.nx) for s in slabs]
    for k in ("pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"):
        got = np.concatenate([p[k] for p in parts], axis=-1)
        assert got.shape == want[k].shape, (k, got.shape, want[k].shape)
        assert np.array_equal(got, want[k]), (k, float(np.nanmax(np.abs(got - want[k]))))
    # monitor: per-slab sums add up to the single-domain figures
    if nsteps % 2 == 0:
        m = ref.monitor()
        ms = [s.solver.monitor() for s in slabs]
        for k in ("vol1_full", "mass1_full", "mass2_full"):
            assert abs(sum(x[k] for x in ms) - m[k]) <= 1e-9 * abs(m[k]) + 1e-12, k
    for s in slabs:
        s.solver.close()
    ref.close()


def test_two_gpus_nccl_equal_single_domain(gpu_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(REPO / "tests" / "slab_nccl_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SLAB_NCCL_OK" in r.stdout, r.stdout[-3000:]
