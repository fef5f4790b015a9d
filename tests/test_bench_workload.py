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
