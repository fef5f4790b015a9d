"""CPU check of the interface-activity criterion (mf-lbm-cuda_b200/csrc/kernels_chain.cuh) against the oracle.

The CUDA chain may skip the normals / cn-extrapolation stencils at the sites of a "quiet" brick - one whose 27-brick
neighbourhood holds no non-solid site with |phi - s| > eps (s = +1 or -1; eps = 1e-7 in double, 0 in single precision) -
and store zeros there.  Here the same flags are built in numpy from the oracle's phi after a number of steps, and the
oracle's own cn_* / c_norm (full evaluation of every stencil, /root/reference/src/main_iteration_GPU.cu:757-906) must be
exactly zero wherever the criterion says quiet.  Also checks that the criterion is worth having: most of a lattice with a
distant interface is quiet."""
import numpy as np
import pytest

import common

BX, BY, BZ = 8, 4, 4
EPS = {"f64": 1e-7, "f32": 0.0}


def quiet_sites(phi4: np.ndarray, wtype4: np.ndarray, eps: float) -> np.ndarray:
    """boolean array on the 4-ghost grid: the site lies in a quiet brick"""
    nonsolid = wtype4 <= 0
    with np.errstate(invalid="ignore"):
        p = nonsolid & ~(np.abs(phi4.astype(np.float64) - 1.0) <= eps)
        m = nonsolid & ~(np.abs(phi4.astype(np.float64) + 1.0) <= eps)
    nz, ny, nx = phi4.shape
    gz, gy, gx = -(-nz // BZ), -(-ny // BY), -(-nx // BX)

    def bricks(a):
        pad = np.zeros((gz * BZ, gy * BY, gx * BX), dtype=bool)
        pad[:nz, :ny, :nx] = a
        return pad.reshape(gz, BZ, gy, BY, gx, BX).any(axis=(1, 3, 5))

    def dilate(b):
        out = np.zeros_like(b)
        padded = np.pad(b, 1)
        for dz in range(3):
            for dy in range(3):
                for dx in range(3):
                    out |= padded[dz:dz + gz, dy:dy + gy, dx:dx + gx]
        return out

    active = dilate(bricks(p)) & dilate(bricks(m))
    return np.repeat(np.repeat(np.repeat(~active, BZ, axis=0), BY, axis=1), BX, axis=2)[:nz, :ny, :nx]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name,steps", [("pack_velocity", 40), ("tube_pressure", 40), ("periodic_drop", 25), ("imbibition_plate2", 40)])
def test_quiet_bricks_hold_zero_normals(name, steps, prec):
    o, ctl, solid = common.make_oracle(name, prec)
    for nsteps in (0, steps):
        if nsteps:
            o.run(1, nsteps)
        quiet2 = quiet_sites(o.arr("phi"), o.arr("walls_type"), EPS[prec])[2:-2, 2:-2, 2:-2]   # on the 2-ghost grid of cn_*, c_norm
        for k in ("c_norm", "cn_x", "cn_y", "cn_z"):
            a = o.arr(k)
            assert a.shape == quiet2.shape
            assert not np.any(a[quiet2] != 0), (name, prec, nsteps, k, int(np.count_nonzero(a[quiet2])))


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_criterion_on_the_benchmark_workload(prec):
    """64 x 64 x 160 cut of the benchmark workload (bench.py): exact zeros in quiet bricks after 300 steps, and at least
    half of the non-solid sites quiet (the interface sits at z = 8)."""
    import bench
    from oracle import Oracle
    nx, ny, nz = 64, 64, 160
    ctl = dict(common.rc.DEFAULT_CONTROL)
    ctl.update(bench.workload_control(nx, ny, nz))
    ctl["external_geometry_read_cmd"] = 1
    o = Oracle(ctl, prec)
    o.setup(bench.workload_geometry(nx, ny, nz))
    o.run(1, 300)
    wt = o.arr("walls_type")
    quiet4 = quiet_sites(o.arr("phi"), wt, EPS[prec])
    quiet2 = quiet4[2:-2, 2:-2, 2:-2]
    for k in ("c_norm", "cn_x", "cn_y", "cn_z"):
        a = o.arr(k)
        assert not np.any(a[quiet2] != 0), (k, int(np.count_nonzero(a[quiet2])))
    nonsolid = wt <= 0
    frac = np.count_nonzero(quiet4 & nonsolid) / np.count_nonzero(nonsolid)
    assert frac > 0.5, frac


def test_python_constants_mirror_the_kernel_header():
    """the numpy emulation above must test the criterion the CUDA code applies: same brick extents, same tolerances"""
    import re
    from pathlib import Path
    src = (Path(__file__).resolve().parent.parent / "mf-lbm-cuda_b200" / "csrc" / "kernels_chain.cuh").read_text()
    m = re.search(r"BR_X = (\d+), BR_Y = (\d+), BR_Z = (\d+)", src)
    assert m and tuple(int(v) for v in m.groups()) == (BX, BY, BZ)
    d = re.search(r"act_eps<double>\(\) \{ return ([0-9.e+-]+); \}", src)
    f = re.search(r"act_eps<float>\(\) \{ return ([0-9.e+-]+)f; \}", src)
    assert d and float(d.group(1)) == EPS["f64"]
    assert f and float(f.group(1)) == EPS["f32"]
