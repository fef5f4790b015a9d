"""N > 1 path on the CPU: the x-slab stepper (mflbm/slab.py: partition, exchange schedule, neighbour wiring) driven
over torch.distributed/gloo with world_size 2 and 3.  The compute backend is the C oracle restricted to each slab
(tests/slab_oracle.py); the union of the slabs must equal a single-domain oracle run bit for bit, and every column a
rank does not own or receive is NaN-poisoned, so a missing or mistimed exchange cannot go unnoticed."""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent
for p in (REPO / "tests", REPO / "oracle", REPO / "mf-lbm-cuda_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, prec, nsteps, outdir, balanced=False):
    import torch
    import torch.distributed as dist
    import common
    from mflbm import slab
    from slab_oracle import OracleSlab
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        o, ctl, solid = common.make_oracle(name, prec)
        rng = slab.partition(o.nx, world, rank)
        if balanced:
            # the bench's --partition balanced protocol: every rank counts the fluid nodes of its equal-width columns, the counts
            # are gathered, all ranks derive the same cost-balanced cuts
            fluid = o.arr("walls")[2:-2, 2:-2, 2:-2] == 0
            mine = fluid[:, :, rng.x0 - 1:rng.x1].sum(axis=(0, 1)).astype(np.float64)
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            rng = slab.partition_balanced(np.concatenate(parts), world, rank, side_cost=0.6 * o.ny * o.nz)
        backend = OracleSlab(o, rng)
        st = slab.SlabStepper(backend, rng)
        st.run(1, nsteps)
        assert st.exchanges == 2 * nsteps
        st.settle()
        mine = backend.owned()
        np.savez(Path(outdir) / f"rank{rank}.npz", **mine)
        # checkpoint in the reference's single-file layout, every rank writing its own columns: the local arrays are this slab's
        # window of the (NaN-poisoned) full-domain arrays, i.e. what a slab solver's download_state returns
        win = lambda a, g: a[..., rng.x0 - 1:rng.x0 - 1 + rng.nx_local + 2 * g]
        local = dict(pdf=win(backend.pdf, 1), phi=win(backend.phi, 4))
        if ctl["outlet_BC"] == 1:
            local.update({k: win(o.arr(k).reshape((-1, o.ny + 2, o.nx + 2)) if k != "phi_convec" else o.arr(k).reshape(o.ny + 2, o.nx + 2), 1)
                          for k in ("f_convec", "g_convec", "phi_convec")})
        slab.write_checkpoint_slabs(Path(outdir) / "id0000", rng, o.nx, local, nsteps + 1, 0.25, 1.5, dist)
        solid_local = (o.arr("walls_global") != 0)[:, :, rng.x0 - 1:rng.x1]
        slab.write_vtk_phase_slabs(Path(outdir) / "small.vtk", rng, o.nx, win(backend.phi, 4), solid_local, dist)
        # monitor reduction: per-slab sums -> global (gloo all_reduce)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,prec,world,nsteps", [
    ("tube_pressure", "f64", 2, 10),       # Zou-He inlet/outlet + porous plate
    ("pack_velocity", "f64", 2, 10),       # velocity inlet + convective outlet
    ("periodic_drop", "f64", 2, 9),        # y/z periodic kernels on the ghost columns; ends on an odd step
    ("pack_velocity", "f32", 3, 6),        # uneven partition (24 columns over 3 ranks -> interior rank with two neighbours)
    ("rect_quirk", "f64", 4, 5),           # 22 columns over 4 ranks: widths 6,6,5,5
])
def test_slabs_over_gloo_equal_single_domain(tmp_path, name, prec, world, nsteps):
    import common
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, prec, nsteps, str(tmp_path)), nprocs=world, join=True)
    ref, ctl, solid = common.make_oracle(name, prec)
    ref.run(1, nsteps)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for k in ("pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"):
        got = np.concatenate([p[k] for p in parts], axis=-1)
        want = ref.arr(k)
        assert got.shape == want.shape, (k, got.shape, want.shape)
        assert not np.isnan(got).any(), f"{k}: NaN - a column was used before it was exchanged"
        assert np.array_equal(got, want), (k, float(np.abs(got - want).max()))


def test_balanced_slabs_over_gloo_equal_single_domain(tmp_path):
    """cost-balanced cuts (slab.balanced_cuts: wide end slabs, narrow interior ones) through the same protocol"""
    import common
    name, prec, world, nsteps = "pack_velocity", "f64", 3, 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, prec, nsteps, str(tmp_path), True), nprocs=world, join=True)
    ref, ctl, solid = common.make_oracle(name, prec)
    ref.run(1, nsteps)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    widths = [p["phi"].shape[-1] for p in parts]
    assert widths[1] < widths[0] and widths[1] < widths[2], widths      # the interior slab pays for two neighbours
    for k in ("pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"):
        got = np.concatenate([p[k] for p in parts], axis=-1)
        assert np.array_equal(got, ref.arr(k)), k


@pytest.mark.parametrize("name,prec,world,nsteps", [("pack_velocity", "f64", 3, 6), ("tube_pressure", "f32", 2, 5)])
def test_slab_checkpoint_equals_single_domain_file(tmp_path, name, prec, world, nsteps):
    """N slabs write one checkpoint file in the reference layout (src/IO_multiphase.cpp:252-305), each rank its own columns through
    a memory map: byte-identical to the file a single domain writes, and readable slab by slab by any other decomposition."""
    import common
    from mflbm import slab
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, prec, nsteps, str(tmp_path)), nprocs=world, join=True)
    ref, ctl, solid = common.make_oracle(name, prec)
    ref.run(1, nsteps)
    conv = ctl["outlet_BC"] == 1
    full = dict(pdf=ref.arr("pdf"), phi=ref.arr("phi"))
    if conv:
        full.update(f_convec=ref.arr("f_convec").reshape(19, ref.ny + 2, ref.nx + 2), g_convec=ref.arr("g_convec").reshape(19, ref.ny + 2, ref.nx + 2),
                    phi_convec=ref.arr("phi_convec").reshape(ref.ny + 2, ref.nx + 2))
    single = tmp_path / "single_id0000"
    slab.write_checkpoint_slabs(single, slab.partition(ref.nx, 1, 0), ref.nx, full, nsteps + 1, 0.25, 1.5)
    assert single.read_bytes() == (tmp_path / "id0000").read_bytes()
    # the host driver's reader (tests/test_host_driver.py) understands it
    from test_host_driver import read_checkpoint
    c = read_checkpoint(single, ref.nx, ref.ny, ref.nz, full["pdf"].dtype, conv)
    assert c["ntime"] == nsteps + 1 and np.array_equal(c["pdf"], full["pdf"]) and np.array_equal(c["phi"], full["phi"])
    # scatter: a different decomposition reads its slabs, ghost columns included
    for w2 in (1, 2, 4):
        for r in range(w2):
            rng = slab.partition(ref.nx, w2, r)
            got = slab.read_checkpoint_slab(tmp_path / "id0000", rng, ref.nx, ref.ny, ref.nz, full["pdf"].dtype, conv)
            assert got["ntime_next"] == nsteps + 1 and got["force_z"] == 0.25 and got["rho_in"] == 1.5
            for k, a in full.items():
                g = 4 if k == "phi" else 1
                assert np.array_equal(got[k], a[..., rng.x0 - 1:rng.x0 - 1 + rng.nx_local + 2 * g]), (w2, r, k)
    # phase-field VTK file written by the slabs: the reference's "small_" format, phi zeroed in solids
    from test_host_driver import read_vtk
    header, npts, fields = read_vtk(tmp_path / "small.vtk")
    assert header[4] == f"DIMENSIONS {ref.nx} {ref.ny} {ref.nz}" and npts == ref.nx * ref.ny * ref.nz
    typ, payload = fields["phi"]
    assert typ == "float" and len(payload) == 4 * npts
    got = np.frombuffer(payload, ">f4").reshape(ref.nz, ref.ny, ref.nx)
    solid_g = ref.arr("walls_global") != 0
    want = np.where(solid_g, 0.0, ref.arr("phi")[4:-4, 4:-4, 4:-4]).astype(np.float32)
    assert np.array_equal(got, want)
    with pytest.raises(IOError):
        slab.read_checkpoint_slab(tmp_path / "id0000", slab.partition(ref.nx + 1, 1, 0), ref.nx + 1, ref.ny, ref.nz, full["pdf"].dtype, conv)


def _monitor_worker(rank, world, port, outdir):
    import json
    import types
    import torch.distributed as dist
    from mflbm import slab
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        r = rank + 1.0
        m = dict(vol1_sum=1.0 * r, vol2_sum=3.0 * r, mass1_sum=r, mass2_sum=r, vol1_full=2.0 * r, vol2_full=2.0 * r, mass1_full=r, mass2_full=r,
                 fl1_avg=0.5 * r, fl2_avg=0.25 * r, fl1_avg_whole=r, fl2_avg_whole=r, kinetic_energy=[r, 2 * r], umax=0.1 * r, nan_detected=int(rank == 1),
                 pre_w_sum=10.0 * r, pre_nw_sum=20.0 * r, n_w=4 * (rank + 1), n_nw=2 * (rank + 1), outlet_phase1_count=rank,
                 profiles={k: np.full(5, r) * (i + 1) for i, k in enumerate(("fl1", "fl2", "pre", "mass1", "mass2", "vol1", "vol2"))})
        params = types.SimpleNamespace(A_xy=4.0, la_nu1=0.5, lbm_gamma=0.25)
        out = slab.reduce_monitor(m, slab.partition(16, world, rank), params, dist)
        out["profiles"] = {k: v.tolist() for k, v in out["profiles"].items()}
        (Path(outdir) / f"mon{rank}.json").write_text(json.dumps(out))
    finally:
        dist.destroy_process_group()


def test_monitor_reduction_over_gloo(tmp_path):
    """per-slab monitor sums -> global figures (SUM / MAX all_reduce), the per-slice profiles and the capillary-pressure sums"""
    import json
    world = 3
    mp.spawn(_monitor_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [json.loads((tmp_path / f"mon{r}.json").read_text()) for r in range(world)]
    assert all(o == outs[0] for o in outs)            # every rank holds the same global figures
    o = outs[0]
    assert o["vol1_sum"] == 6.0 and o["vol2_sum"] == 18.0 and o["saturation"] == 0.25 and o["saturation_full_domain"] == 0.5
    assert o["kinetic_energy"] == [6.0, 12.0] and abs(o["umax"] - 0.3) < 1e-15 and o["nan_detected"] == 1
    assert o["pre_w_sum"] == 60.0 and o["n_w"] == 24.0 and o["n_nw"] == 12.0 and o["outlet_phase1_count"] == 3.0
    assert o["ca"] == ((3.0 + 1.5) / 4.0) * 0.5 / 0.25
    assert o["profiles"]["fl1"] == [6.0] * 5 and o["profiles"]["vol2"] == [42.0] * 5


def test_balanced_cuts():
    from mflbm import slab
    rng = np.random.default_rng(3)
    for nx, world in ((64, 2), (257, 4), (2048, 8), (40, 8)):
        col = rng.uniform(0.5, 1.5, nx) * 1000.0
        for side in (0.0, 5000.0):
            cuts = slab.balanced_cuts(col, world, side)
            assert cuts[0] == 0 and cuts[-1] == nx and len(cuts) == world + 1
            w = np.diff(cuts)
            assert w.min() >= 4
            rs = [slab.partition_balanced(col, world, r, side) for r in range(world)]
            assert rs[0].x0 == 1 and rs[-1].x1 == nx and all(a.x1 + 1 == b.x0 for a, b in zip(rs, rs[1:]))
            if nx >= 64:   # costs equal to within two columns' worth
                cost = [col[cuts[r]:cuts[r + 1]].sum() + side * ((r > 0) + (r < world - 1)) for r in range(world)]
                assert max(cost) - min(cost) <= 2 * col.max() + 1e-9, (nx, world, side, cost)
    # uniform columns, no halo cost: the equal-width partition
    assert slab.balanced_cuts(np.ones(1024), 4) == [0, 256, 512, 768, 1024]
    # the 8-slab weak-scaling shape: end slabs wider than interior ones
    w = np.diff(slab.balanced_cuts(np.full(2048, 29000.0), 8, 10 * 65536.0))
    assert w[0] == w[-1] and w[0] > w[1] and len(set(w[1:-1])) <= 2
    with pytest.raises(ValueError):
        slab.balanced_cuts(np.ones(10), 4)


def test_partition_covers_the_lattice():
    from mflbm import slab
    for nx in (16, 22, 257, 1024):
        for world in (1, 2, 3, 4, 8):
            if nx // world < 4:
                continue
            rs = [slab.partition(nx, world, r) for r in range(world)]
            assert rs[0].x0 == 1 and rs[-1].x1 == nx
            assert all(a.x1 + 1 == b.x0 for a, b in zip(rs, rs[1:]))
            assert max(r.nx_local for r in rs) - min(r.nx_local for r in rs) <= 1
            assert not rs[0].has_left and not rs[-1].has_right and all(r.has_right for r in rs[:-1])
    with pytest.raises(ValueError):
        slab.partition(10, 4, 0)
