"""CUDA path (through the C ABI) against the C oracle on the same seeded inputs.  All tests need a GPU."""
import os

import numpy as np
import pytest

import common
from common import TOL, relerr

pytestmark = pytest.mark.gpu

ALL = sorted(common.CASES)
DYN = ["tube_pressure", "pack_velocity", "imbibition_plate2", "periodic_drop", "duct_no_geometry", "rect_quirk"]


def interior_walls(o):
    return (o.arr("walls_global") != 0).astype(np.int8)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ALL)
def test_geometry_preprocessing_bit_exact(gpu_lib, name, prec):
    import mflbm
    o, ctl, solid = common.make_oracle(name, prec)
    s = mflbm.Solver(mflbm.derive_params(ctl, prec), prec)
    s.preprocess_geometry(interior_walls(o))
    g = s.download_geometry()
    for k in ("walls", "walls_type"):
        assert np.array_equal(g[k], o.arr(k)), k
    for k in ("s_nx", "s_ny", "s_nz"):
        assert np.array_equal(g[k].view(np.uint8), o.arr(k).view(np.uint8)), f"{k} not bit-exact"
    want = [int(o.scalar(k)) for k in ("num_solid_boundary_global", "num_fluid_boundary_global", "num_solid_boundary", "num_fluid_boundary")]
    assert g["counts"].tolist() == want
    assert s.num_fluid_nodes == int(o.scalar("pore_sum"))
    s.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ALL)
def test_init_state(gpu_lib, name, prec):
    import mflbm
    o, ctl, solid = common.make_oracle(name, prec)
    s = mflbm.Solver(mflbm.derive_params(ctl, prec), prec)
    s.preprocess_geometry(interior_walls(o))
    s.init_state(ctl["initial_fluid_distribution_option"], ctl["initial_interface_position"], W_in=o.arr("W_in"))
    st = s.download_state()
    assert np.array_equal(st["pdf"], o.arr("pdf")), "initial pdf must be bit-exact"
    if ctl["outlet_BC"] == 1:
        for k in ("f_convec", "g_convec", "phi_convec"):
            assert np.array_equal(st[k], o.arr(k)), k
    # phi / cn / curv went through the gradient chain on the GPU (FMA contraction) -> tolerance
    for k in ("phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"):
        assert relerr(st[k], o.arr(k)) <= TOL[prec], (k, relerr(st[k], o.arr(k)))
    s.close()


@pytest.mark.parametrize("nsteps", [1, 2, 100])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", DYN)
def test_steps_match_oracle(gpu_lib, name, prec, nsteps):
    o, ctl, solid = common.make_oracle(name, prec)
    s = common.solver_from_oracle(o, ctl, prec)
    for n in range(nsteps):
        s.step(1 + n)
    o.run(1, nsteps)
    st = s.download_state()
    tol = TOL[prec]
    assert np.isfinite(st["pdf"]).all()
    assert relerr(st["pdf"], o.arr("pdf")) <= tol, ("pdf", relerr(st["pdf"], o.arr("pdf")))
    assert relerr(st["phi"], o.arr("phi")) <= tol, ("phi", relerr(st["phi"], o.arr("phi")))
    # derived fields: threshold branches (c_norm < 1e-6, secant solver) may amplify rounding locally -> looser
    for k in ("cn_x", "cn_y", "cn_z", "c_norm", "curv"):
        ok, msg = common.derived_close(st[k], o.arr(k), tol, common.derived_weight(o.arr("c_norm"), k))
        assert ok, (k, msg)
    if ctl["outlet_BC"] == 1:
        for k in ("f_convec", "g_convec"):
            assert relerr(st[k], o.arr(k), scale_of=o.arr("pdf")) <= tol, k
        assert relerr(st["phi_convec"], o.arr("phi_convec")) <= tol
    s.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_run_with_graph_equals_step_loop(gpu_lib, prec):
    o, ctl, solid = common.make_oracle("tube_pressure", prec)
    a = common.solver_from_oracle(o, ctl, prec)
    b = common.solver_from_oracle(o, ctl, prec)
    for n in range(21):
        a.step(1 + n)
    b.run(1, 21)
    sa, sb = a.download_state(), b.download_state()
    for k in ("pdf", "phi", "cn_x", "c_norm", "curv"):
        assert np.array_equal(sa[k], sb[k]), k
    assert b.kernel_launches == a.kernel_launches
    a.close(); b.close()


@pytest.mark.parametrize("mrt", [1, 2, 3, 4])
def test_mrt_variants(gpu_lib, mrt):
    import mflbm
    from oracle import Oracle
    ctl, solid = common.full_control("tube_pressure")
    o = Oracle(ctl, "f64", mrt=mrt).setup(solid)
    P = mflbm.derive_params(ctl, "f64", mrt=mrt)
    s = mflbm.Solver(P, "f64")
    s.upload_geometry(o.arr("walls"), o.arr("walls_type"), o.arr("s_nx"), o.arr("s_ny"), o.arr("s_nz"))
    s.upload_state(pdf=o.arr("pdf"), phi=o.arr("phi"), cn_x=o.arr("cn_x"), cn_y=o.arr("cn_y"), cn_z=o.arr("cn_z"), c_norm=o.arr("c_norm"),
                   curv=o.arr("curv"))
    s.run(1, 10)
    o.run(1, 10)
    st = s.download_state()
    assert relerr(st["pdf"], o.arr("pdf")) <= TOL["f64"]
    s.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["tube_pressure", "pack_velocity"])
def test_monitor_matches_oracle(gpu_lib, name, prec):
    o, ctl, solid = common.make_oracle(name, prec)
    s = common.solver_from_oracle(o, ctl, prec)
    s.run(1, 50)
    o.run(1, 50)
    m = s.monitor(profiles=True)
    mo, prof = o.monitor()
    # the oracle sums sequentially in T like the reference; the device sums in double -> 1e-6 (north_star) in f64
    tol = 1e-9 if prec == "f64" else 2e-4
    assert abs(m["saturation"] - mo["saturation"]) <= tol
    assert abs(m["saturation_full_domain"] - mo["saturation_full_domain"]) <= tol
    assert abs(m["umax"] - mo["umax_global"]) <= tol * max(1.0, mo["umax_global"])
    assert abs(m["ca"] - mo["ca"]) <= (1e-9 if prec == "f64" else 1e-4) * max(abs(mo["ca"]), 1e-6) + 1e-12
    names = ["fl1", "fl2", "pre", "mass1", "mass2", "vol1", "vol2"]
    for n, k in enumerate(names):
        scale = np.abs(prof[n]).max() + 1e-30
        assert np.abs(m["profiles"][k] - prof[n]).max() / scale <= (1e-9 if prec == "f64" else 1e-3), k
    assert m["nan_detected"] == 0
    s.close()


def test_errors_are_reported_not_fatal(gpu_lib):
    import ctypes as C
    import mflbm
    ctl, solid = common.full_control("tube_pressure")
    P = mflbm.derive_params(ctl, "f64")
    P.iper = 1
    with pytest.raises(mflbm.MflbmError, match="x-periodic"):
        mflbm.Solver(P, "f64")
    P.iper = 0
    s = mflbm.Solver(P, "f64")
    with pytest.raises(mflbm.MflbmError, match="before geometry"):
        s.step(1)
    assert gpu_lib.mflbm_f64_step(None, 1) != 0
    assert b"null solver" in gpu_lib.mflbm_last_error()
    s.close()




@pytest.mark.parametrize("cap", [1, 3])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["pack_velocity", "periodic_drop"])
def test_stage_rings_wrap_on_small_lattices(gpu_lib, monkeypatch, name, prec, cap):
    """The collide kernels are persistent: a CTA walks tiles blockIdx, blockIdx + grid, ... through rings of shared-memory
    stages.  On the small parity lattices every CTA gets at most one tile, so the rings never wrap; MFLBM_MAX_CTAS caps
    the grid so that they do (one CTA then walks every tile), and the result must still match the oracle."""
    monkeypatch.setenv("MFLBM_MAX_CTAS", str(cap))
    o, ctl, solid = common.make_oracle(name, prec)
    s = common.solver_from_oracle(o, ctl, prec)
    s.run(1, 30)
    o.run(1, 30)
    st = s.download_state()
    for k in ("pdf", "phi"):
        assert relerr(st[k], o.arr(k)) <= TOL[prec], (k, relerr(st[k], o.arr(k)))
    s.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_full_size_run_is_deterministic(gpu_lib, prec):
    """Two solvers, same 256^3 benchmark input, stepped side by side: every PDF must agree bit for bit after each of the
    first steps.  A stage of the TMA ring that is refilled before all of its readers are done shows up here as a handful
    of differing nodes (it did, once per ~10^5 warp-tiles, before the stage release was ordered behind the phi store)."""
    import bench
    import mflbm
    n = 256
    ctl = bench.workload_control(n, n, n)
    solid = bench.workload_geometry(n, n, n)
    W = bench.inlet_profile(ctl, prec)
    P = mflbm.derive_params(ctl, prec)
    pair = []
    for _ in range(2):
        s = mflbm.Solver(P, prec)
        s.preprocess_geometry(solid)
        s.init_state(1, ctl["initial_interface_position"], W_in=W)
        pair.append(s)
    for k in range(6):
        for s in pair:
            s.step(1 + k)
        a, b = (s.download_state(fields=("pdf",))["pdf"] for s in pair)
        assert np.array_equal(a, b), (k + 1, int((a != b).sum()))
    for s in pair:
        s.close()


def _pair_with_and_without_activity(monkeypatch, make, scan=False, chain="brick"):
    """two solvers over the same input: list gradient chain (every site every step, kernels_step.cuh) / brick chain
    (kernels_chain.cuh: only where an interface can be).  MFLBM_CHAIN is read when a solver is created.  scan: the brick
    flags come from a scan of phi instead of the collide kernels (MFLBM_ACT_SCAN).  chain: "brick" = k_chain_normals (node-type
    tile, sites collected per step), "csr" = k_chain_normals_csr (per-brick site lists built once per geometry, the default)"""
    monkeypatch.setenv("MFLBM_CHAIN", "list")
    plain = make()
    monkeypatch.setenv("MFLBM_CHAIN", chain)
    if scan:
        monkeypatch.setenv("MFLBM_ACT_SCAN", "1")
    act = make()
    monkeypatch.delenv("MFLBM_CHAIN", raising=False)
    monkeypatch.delenv("MFLBM_ACT_SCAN", raising=False)
    return plain, act


def _assert_states_identical(plain, act, where):
    a, b = plain.download_state(), act.download_state()
    assert sorted(a) == sorted(b)
    for k in sorted(a):
        assert np.array_equal(a[k], b[k], equal_nan=True), (where, k, int((a[k] != b[k]).sum()))


@pytest.mark.parametrize("chain", ["brick", "csr"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["pack_velocity", "tube_pressure", "periodic_drop", "imbibition_plate2", "rect_quirk", "duct_no_geometry"])
def test_brick_chain_is_bit_identical(gpu_lib, monkeypatch, name, prec, chain):
    """kernels_chain.cuh: where every non-solid phi of a 27-brick neighbourhood is +1 (or -1) to within eps the chain
    skips its stencils; elsewhere it evaluates them from a TMA-staged tile.  Every array - phi at the solid-boundary sites and
    the zeroed normals included - must equal the list chain's bit for bit, step loop and graph replay alike."""
    o, ctl, solid = common.make_oracle(name, prec)
    plain, act = _pair_with_and_without_activity(monkeypatch, lambda: common.solver_from_oracle(o, ctl, prec), chain=chain)
    _assert_states_identical(plain, act, "initial state")
    for s in (plain, act):
        for n in range(1, 4):
            s.step(n)
    _assert_states_identical(plain, act, "3 steps")
    for s in (plain, act):
        s.run(4, 57)
    _assert_states_identical(plain, act, "60 steps")
    o.run(1, 60)
    st = act.download_state()
    for k in ("pdf", "phi"):
        assert relerr(st[k], o.arr(k)) <= TOL[prec], (k, relerr(st[k], o.arr(k)))
    plain.close(); act.close()


@pytest.mark.parametrize("chain", ["brick", "csr"])
@pytest.mark.parametrize("scan", [False, True])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_brick_chain_follows_a_moving_interface(gpu_lib, monkeypatch, prec, scan, chain):
    """A lattice large enough for most bricks to be quiet (64 x 48 x 128 pack, interface at z = 40), driven hard enough
    (Ca 0.05) for the front to sweep bricks from quiet to active and back: sites whose normals were non-zero must be zeroed
    when their brick falls quiet again, held extrapolated values must be dropped when it wakes up."""
    import bench
    import mflbm
    nx, ny, nz = 64, 48, 128
    ctl = bench.workload_control(nx, ny, nz)
    ctl.update(initial_interface_position=40.0, capillary_number=5e-2)
    solid = bench.workload_geometry(nx, ny, nz, seed=5)
    W = bench.inlet_profile(ctl, prec)
    P = mflbm.derive_params(ctl, prec)

    def make():
        s = mflbm.Solver(P, prec)
        s.preprocess_geometry(solid)
        s.init_state(1, ctl["initial_interface_position"], W_in=W)
        return s

    plain, act = _pair_with_and_without_activity(monkeypatch, make, scan=scan, chain=chain)
    _assert_states_identical(plain, act, "initial state")
    nt = 1
    for chunk in (1, 1, 98, 300):
        for s in (plain, act):
            s.run(nt, chunk)
        nt += chunk
        _assert_states_identical(plain, act, f"{nt - 1} steps")
    assert plain.monitor()["nan_detected"] == 0
    plain.close(); act.close()
