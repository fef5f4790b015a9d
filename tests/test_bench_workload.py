"""CPU checks of the synthetic benchmark workload (bench.py): the slab-window rasteriser must reproduce the full geometry
on the columns a slab reads, and the numpy inlet profile must equal the oracle's (itself pinned against the reference)."""
import numpy as np
import pytest

import bench
import common


def test_window_geometry_equals_full_geometry():
    nx, ny, nz = 96, 48, 40
    full = bench.workload_geometry(nx, ny, nz)
    assert full.dtype == np.int8 and set(np.unique(full)) <= {0, 1}
    assert 0.3 < 1.0 - full[:, 1:-1, 1:-1].mean() < 0.6     # porosity of the pack
    for lo, hi in ((1, 30), (19, 66), (60, 96), (-11, 36), (70, 108)):
        win = bench.workload_geometry_window(nx, ny, nz, lo, hi)
        a, b = max(1, lo), min(nx, hi)
        assert np.array_equal(win[:, :, a - 1:b], full[:, :, a - 1:b]), (lo, hi)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_numpy_inlet_profile_matches_oracle(prec):
    """bench.inlet_profile vs the oracle's W_in for the same duct.  The oracle evaluates the series with libm in the
    solver precision (bit-exact against the reference CPU code); numpy's float32 transcendental functions may differ in
    the last bits, so FP32 gets a tolerance."""
    from oracle import Oracle
    n = 24
    ctl = bench.workload_control(n, n, n)
    full = dict(common.rc.DEFAULT_CONTROL)
    full.update(ctl)
    full["external_geometry_read_cmd"] = 1
    o = Oracle(full, prec)
    o.setup(bench.workload_geometry(n, n, n))
    W = bench.inlet_profile(ctl, prec)
    Wo = o.arr("W_in")
    assert W.shape == Wo.shape
    if prec == "f64":
        assert np.abs(W - Wo).max() <= 1e-13 * np.abs(Wo).max()   # numpy sums the series pairwise, the reference sequentially
    else:
        assert np.abs(W - Wo).max() <= 2e-5 * np.abs(Wo).max()


def test_workload_defaults_follow_the_baseline_configs():
    """N = 1: BASELINE configs[1] (256^3 drainage); N > 1: configs[4] (512^3 per GPU, imbibition); --config strong: configs[3]"""
    import bench
    def resolved(argv, world=1):
        a = bench.build_parser().parse_args(argv)
        bench.resolve_workload(a, world)
        return a
    a = resolved([])
    assert (a.gpus, a.size, a.case, a.global_nx, a.prec) == (1, 256, "drainage", 0, "f64")
    a = resolved(["--steps", "2", "--warmup", "1"])
    assert a.warmup == 3 and a.steps == 2          # at least three warm-up steps
    a = resolved(["--gpus", "8"], world=8)
    assert (a.size, a.case, a.global_nx) == (512, "imbibition", 0)
    a = resolved(["--gpus", "2", "--impl", "reference"])      # the reference arm of an N-GPU run: one GPU's share of that lattice
    assert (a.size, a.case) == (512, "imbibition")
    a = resolved(["--gpus", "4", "--config", "strong"], world=4)
    assert (a.size, a.case, a.global_nx, a.seed) == (512, "drainage", 1024, 20240230)
    a = resolved(["--gpus", "4", "--size", "256", "--case", "drainage"], world=4)
    assert (a.size, a.case) == (256, "drainage")
    a = resolved([], world=2)                      # launched under torchrun without --gpus
    assert a.gpus == 2 and a.size == 512
