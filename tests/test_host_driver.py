"""The C++ host driver (mf-lbm-cuda_b200/bin/mflbm_run) as a drop-in for the reference program.

  * CPU (`-m "not gpu"`): `--check-input` parses a reference-format case directory and prints what the host derives before
    the GPU layer is touched; scalars must be bit-identical to the ones the oracle-checked Python mirror derives
    (mflbm.derive_params), the velocity-inlet profile bit-identical to the oracle's (itself pinned against the
    reference's CPU code), and the reference's fatal input errors must be fatal here too.
  * GPU (`-m gpu`): the driver and the reference's STOCK program (oracle/_ref/MF_LBM_CUDA_*, src/main.cpp compiled
    unmodified) run the same case directory; monitor files, checkpoint and VTK output are compared.
"""
import shutil
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

import common
import refcase as rc

REPO = Path(__file__).resolve().parent.parent
EXE = REPO / "mf-lbm-cuda_b200" / "bin" / "mflbm_run"


def run_driver(case: Path, *args, check=True, timeout=900):
    if not EXE.exists():
        subprocess.run(["make", "-C", str(REPO / "mf-lbm-cuda_b200" / "host"), "all"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([str(EXE), "--dir", str(case), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if check and r.returncode != 0:
        raise AssertionError(f"mflbm_run failed ({r.returncode}):\n{r.stdout[-3000:]}")
    return r


def parse_check(out: str) -> dict:
    d = {}
    for line in out.splitlines():
        if line.startswith("CHECK "):
            k, *v = line.split()[1:]
            d[k] = v
    return d


def fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["tube_pressure", "pack_velocity", "periodic_drop", "rect_quirk"])
def test_check_input_matches_python_mirror_and_oracle(tmp_path, name, prec):
    import mflbm
    ctl, solid = common.CASES[name]()
    full = rc.write_case(tmp_path, ctl, solid)
    chk = parse_check(run_driver(tmp_path, "--prec", prec, "--check-input").stdout)
    o, octl, _ = common.make_oracle(name, prec)
    assert [int(x) for x in chk["dims"]] == [o.nx, o.ny, o.nz]
    # interior walls after set_walls (incl. the swapped-stride read quirk for nx != ny): the oracle's array is pinned
    # against the reference CPU code (tests/test_oracle_vs_reference.py)
    want_walls = (o.arr("walls_global") != 0).astype(np.int8)
    assert int(chk["walls_fnv1a"][0]) == fnv1a(want_walls.tobytes())
    assert int(chk["pore_sum"][0]) == int(o.scalar("pore_sum"))
    P = mflbm.derive_params(full, prec)
    R = np.float32 if prec == "f32" else np.float64
    for k in ("cos_theta", "la_nui1", "la_nui2", "phi_inlet", "force_z", "rho_in", "rho_out", "uin_avg", "A_xy"):
        got = R(float.fromhex(chk[k][0]))
        assert got == R(getattr(P, k)), (k, got, getattr(P, k))
    if full["inlet_BC"] == 1 and full["kper"] == 0:
        assert int(chk["W_in_fnv1a"][0]) == fnv1a(np.ascontiguousarray(o.arr("W_in")).tobytes()), "W_in must be bit-identical to the reference's"


def test_fatal_inputs_are_fatal(tmp_path):
    ctl, solid = common.CASES["tube_pressure"]()
    # a trailing newline in job_status.txt (src/IO_multiphase.cpp:25-36 compares the whole file)
    a = tmp_path / "a"
    rc.write_case(a, ctl, solid, job_status="new_simulation\n")
    assert run_driver(a, "--check-input", check=False).returncode == 1
    # contact angle above 90 degrees (src/IO_multiphase.cpp:200-203)
    b = tmp_path / "b"
    rc.write_case(b, dict(ctl, theta=120), solid)
    r = run_driver(b, "--check-input", check=False)
    assert r.returncode == 1 and "contact angle" in r.stdout
    # a missing key (src/utils.cpp:67)
    c = tmp_path / "c"
    rc.write_case(c, ctl, solid)
    f = c / "input" / "simulation_control.txt"
    f.write_text("\n".join(l for l in f.read_text().splitlines() if not l.startswith("RK_beta")) + "\n")
    r = run_driver(c, "--check-input", check=False)
    assert r.returncode == 1 and "RK_beta" in r.stdout
    # a geometry file name that does not end in a digit (src/Misc.cpp:31-34)
    d = tmp_path / "d"
    rc.write_case(d, ctl, solid)
    (d / "input" / "Geometry_File_Path.txt").write_text("geo_file_path_\tinput/geometry/nodigits\ngeo_boundary_file_path_\tinput/geometry/nodigits\n")
    assert run_driver(d, "--check-input", check=False).returncode == 1


def test_slab_windows_and_owned_columns(tmp_path):
    """host/slabs.hpp (what mflbm_run --gpus N uploads to / gathers from every slab), compiled and run on the CPU"""
    exe = tmp_path / "slabs_check"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", str(REPO / "tests" / "host_slabs_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "SLABS_OK" in r.stdout, r.stdout


# ----------------------------------------------------------------------------------------------------------
# GPU: against the reference's stock program
# ----------------------------------------------------------------------------------------------------------
def read_checkpoint(path: Path, nx, ny, nz, rt, convective):
    raw = path.read_bytes()
    s = np.dtype(rt).itemsize
    nt = struct.unpack_from("<i", raw, 0)[0]
    force_z, rho_in = np.frombuffer(raw, rt, 2, 4)
    off = 4 + 2 * s
    n1, n4, npl = (nx + 2) * (ny + 2) * (nz + 2), (nx + 8) * (ny + 8) * (nz + 8), (nx + 2) * (ny + 2)
    pdf = np.frombuffer(raw, rt, 38 * n1, off).reshape(2, 19, nz + 2, ny + 2, nx + 2); off += 38 * n1 * s
    phi = np.frombuffer(raw, rt, n4, off).reshape(nz + 8, ny + 8, nx + 8); off += n4 * s
    out = dict(ntime=nt, force_z=force_z, rho_in=rho_in, pdf=pdf, phi=phi)
    if convective:
        out["f_convec"] = np.frombuffer(raw, rt, 19 * npl, off); off += 19 * npl * s
        out["g_convec"] = np.frombuffer(raw, rt, 19 * npl, off); off += 19 * npl * s
        out["phi_convec"] = np.frombuffer(raw, rt, npl, off); off += npl * s
    assert off == len(raw), "checkpoint size"
    return out


def read_vtk(path: Path):
    raw = path.read_bytes()
    pos, header, fields = 0, [], {}
    def line():
        nonlocal pos
        e = raw.index(b"\n", pos); l = raw[pos:e].decode(); pos = e + 1; return l
    for _ in range(8):
        header.append(line())
    n = int(header[-1].split()[1])
    while pos < len(raw):
        _, name, typ = line().split()
        line()   # LOOKUP_TABLE default
        fields[name] = (typ, raw[pos:])   # payload cut by the caller (the declared type is not the stored one)
        # find the next "SCALARS" marker to delimit
        nxt = raw.find(b"SCALARS ", pos)
        end = len(raw) if nxt < 0 else nxt
        fields[name] = (typ, raw[pos:end]); pos = end
    return header, n, fields


def rows(path: Path):
    return np.array([[float(x) for x in l.split() if x != "sec"] for l in path.read_text().splitlines() if l.strip()])


STOCK_CASES = {
    # label: (case of tests/common.py, control overrides)
    "tube_pressure": ("tube_pressure", dict(max_time_step=39, monitor_timer=10, monitor_profile_timer_ratio=2, computation_time_timer=20, display_steps_timer=20,
                                            ntime_animation=20, ntime_visual=40, benchmark_cmd=0, steady_state_option=3, convergence_criteria=1e-12)),
    "pack_velocity": ("pack_velocity", dict(max_time_step=29, monitor_timer=10, monitor_profile_timer_ratio=1, computation_time_timer=30, display_steps_timer=10,
                                            ntime_animation=1000000, ntime_visual=1000000, benchmark_cmd=1, steady_state_option=0)),
    # change_inlet_fluid_phase (src/Misc.cpp:277-384) and the capillary-pressure steady-state monitor (src/Monitor.cpp:357-441)
    "tube_change_inlet": ("tube_pressure", dict(max_time_step=19, monitor_timer=10, computation_time_timer=20, display_steps_timer=1000, benchmark_cmd=1,
                                                change_inlet_fluid_phase_cmd=2, steady_state_option=1, convergence_criteria=1e-12)),
    # y/z periodic body-force case with the phase-field steady-state monitor (src/Monitor.cpp:279-352)
    "periodic_phasefield": ("periodic_drop", dict(max_time_step=19, monitor_timer=10, computation_time_timer=20, display_steps_timer=1000, benchmark_cmd=1,
                                                  steady_state_option=2, convergence_criteria=1e-12)),
}


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(STOCK_CASES))
def test_driver_matches_stock_reference_program(gpu_lib, tmp_path, name, prec):
    _compare_with_stock_program(tmp_path, name, prec, lambda case: run_driver(case, "--prec", prec))


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["tube_pressure", "pack_velocity", "periodic_phasefield"])
def test_shim_program_matches_stock_reference_program(gpu_lib, tmp_path, name, prec):
    """The drop-in, compiled: the reference's own src/main.cpp + CPU sources (unmodified) with integration/mflbm_shim.cpp in
    place of src/main_iteration_GPU.cu and src/Init_multiphase_GPU.cu, linked against libmflbm.so (oracle/build_ref.sh,
    target shim), against the stock program on the same case directories: every output file."""
    shim = rc.REF_BIN_DIR / f"MF_LBM_CUDA_shim_{prec}"
    if not shim.exists():
        pytest.skip(f"{shim} not built (oracle/build_ref.sh shim needs /root/reference)")

    def run(case):
        r = subprocess.run([str(shim)], cwd=str(case), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:]
    _compare_with_stock_program(tmp_path, name, prec, run)


def _compare_with_stock_program(tmp_path, name, prec, run_ours):
    stock = rc.REF_BIN_DIR / f"MF_LBM_CUDA_{prec}"
    if not stock.exists():
        pytest.skip(f"{stock} not built (oracle/build_ref.sh needs /root/reference)")
    label, (name, over) = name, STOCK_CASES[name]
    ctl, solid = common.CASES[name]()
    ctl = dict(ctl, **over)
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    full = rc.write_case(ours, ctl, solid)
    shutil.copytree(ours, ref)
    run_ours(ours)
    r = subprocess.run([str(stock)], cwd=str(ref), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    nz, ny, nx = solid.shape
    rt = np.float32 if prec == "f32" else np.float64
    tol = common.TOL[prec]
    # job status (written to ./job_status.txt by both)
    assert (ours / "job_status.txt").read_text() == (ref / "job_status.txt").read_text() == "simulation_reached_max_step\n"
    # monitor files: same rows, same columns (printed with 6 significant digits)
    out_o, out_r = ours / "results" / "out1.output", ref / "results" / "out1.output"
    names = sorted(p.name for p in out_r.iterdir() if p.is_file() and p.suffix == ".dat")
    assert names == sorted(p.name for p in out_o.iterdir() if p.is_file() and p.suffix == ".dat")
    for n in names:
        a, b = rows(out_o / n), rows(out_r / n)
        assert a.shape == b.shape, n
        if n == "time.dat":
            assert np.array_equal(a[:, 0], b[:, 0])
            continue
        # columns that are sums of signed values (flow rates at rest) are compared on the scale of the file
        scale = np.maximum(np.abs(b), np.abs(b).max(axis=0, keepdims=True))
        rtol = 2e-5 if prec == "f64" else 2e-3
        bad = np.abs(a - b) > rtol * scale + 1e-30
        if n in ("Ca_number.dat", "flowrate_time.dat", "steady_monitor_saturation_error.dat"):
            bad &= np.abs(a - b) > (1e-9 if prec == "f64" else 1e-6)   # cancelling sums / differences of nearly equal numbers
        assert not bad.any(), (n, a[bad][:4], b[bad][:4])
    assert sorted(p.name for p in (out_o / "profile").iterdir()) == sorted(p.name for p in (out_r / "profile").iterdir())
    for p in (out_r / "profile").iterdir():
        a, b = rows(out_o / "profile" / p.name), rows(p)
        assert a.shape == b.shape
        assert np.allclose(a[1:], b[1:], rtol=2e-5 if prec == "f64" else 2e-3, atol=1e-9 if prec == "f64" else 1e-5), p.name   # row 0 is out of bounds in the reference
    info_o, info_r = (out_o / "info.txt").read_text().splitlines(), (out_r / "info.txt").read_text().splitlines()
    assert info_o == info_r
    # checkpoint
    conv = full["outlet_BC"] == 1
    co = read_checkpoint(ours / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, rt, conv)
    cr = read_checkpoint(ref / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, rt, conv)
    assert co["ntime"] == cr["ntime"] and co["force_z"] == cr["force_z"] and co["rho_in"] == cr["rho_in"]
    assert common.relerr(co["pdf"], cr["pdf"]) <= tol
    # fluid nodes as the programs see them: domain walls added, geometry read with the reference's swapped strides when
    # nx != ny - the oracle's wall array, pinned against the reference CPU code
    o_walls = common.make_oracle(name, prec)[0]
    walls_seen = o_walls.arr("walls_global").copy()   # (arr() is a view into the oracle's memory)
    fluid = np.zeros(cr["phi"].shape, bool)
    fluid[4:-4, 4:-4, 4:-4] = walls_seen == 0
    # the reference's host phi was zeroed inside solids by its monitor / VTK writer before the checkpoint was written
    assert np.abs(co["phi"][fluid].astype(np.float64) - cr["phi"][fluid]).max() <= tol * max(1.0, float(np.abs(cr["phi"]).max()))
    if conv:
        # Reference defect: its step function never copies the convective-outlet buffers back to the host
        # (src/main_iteration_GPU.cu:2059-2076 downloads phi, curv, c_norm, cn_*, pdf only), so its checkpoint carries the
        # INITIAL buffers whatever the step.  We write the device state; the checker for it is the oracle.
        o, _, _ = common.make_oracle(name, prec)
        for k in ("f_convec", "g_convec", "phi_convec"):
            assert np.array_equal(cr[k], o.arr(k).reshape(-1)), f"reference {k} is expected to be the step-0 buffer"
        o.run(1, cr["ntime"] - 1)
        for k in ("f_convec", "g_convec"):
            assert common.relerr(co[k], o.arr(k).reshape(-1), scale_of=cr["pdf"]) <= tol, k
        assert common.relerr(co["phi_convec"], o.arr("phi_convec").reshape(-1)) <= tol
    # VTK files: same set, same headers, same payload sizes, values within tolerance
    fo, fr = ours / "results" / "out3.field_data", ref / "results" / "out3.field_data"
    vt_r = sorted(str(p.relative_to(fr)) for p in fr.rglob("*.vtk"))
    assert vt_r == sorted(str(p.relative_to(fo)) for p in fo.rglob("*.vtk"))
    for rel in vt_r:
        ho, no_, fo_ = read_vtk(fo / rel)
        hr, nr_, fr_ = read_vtk(fr / rel)
        assert ho == hr and list(fo_) == list(fr_), rel
        for k in fr_:
            (to, bo), (tr, br) = fo_[k], fr_[k]
            assert to == tr and len(bo) == len(br), (rel, k)
            if k == "walls":
                assert bo == br
                continue
            dt = np.dtype(">f4") if len(br) == 4 * nr_ else np.dtype(">f8")
            a, b = np.frombuffer(bo, dt).astype(np.float64), np.frombuffer(br, dt).astype(np.float64)
            if k in ("phi", "density"):   # solid nodes included: both writers store 0 there
                assert np.abs(a - b).max() <= max(tol, 1e-6 if dt.itemsize == 4 else 0) * max(1.0, np.abs(b).max()), (rel, k)
            else:   # velocities: O(1e-3 .. 1e-6) numbers built from cancelling sums
                assert np.abs(a - b).max() <= (1e-9 if prec == "f64" else 1e-5), (rel, k)


@pytest.mark.gpu
def test_driver_restart_continues_like_the_reference(gpu_lib, tmp_path):
    """continue_simulation from one and the same reference-written checkpoint: both programs re-execute the last step
    index and must end in the same state."""
    prec = "f64"
    stock = rc.REF_BIN_DIR / f"MF_LBM_CUDA_{prec}"
    if not stock.exists():
        pytest.skip(f"{stock} not built")
    ctl, solid = common.CASES["tube_pressure"]()
    base = dict(ctl, max_time_step=19, monitor_timer=10, computation_time_timer=20, display_steps_timer=1000, benchmark_cmd=1)
    first = tmp_path / "first"
    rc.write_case(first, base, solid)
    r = subprocess.run([str(stock)], cwd=str(first), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    # the second leg must END on a monitor step: the reference checkpoints its last downloaded host copy, which is only
    # refreshed on timer steps
    second = dict(base, max_time_step=20)
    for d in (ours, ref):
        shutil.copytree(first, d)
        (d / "input" / "simulation_control.txt").write_text(rc.control_text(dict(second, nxGlobal=solid.shape[2], nyGlobal=solid.shape[1], nzGlobal=solid.shape[0],
                                                                                  external_geometry_read_cmd=1)))
        (d / "input" / "job_status.txt").write_text("continue_simulation")
    run_driver(ours, "--prec", prec)
    r = subprocess.run([str(stock)], cwd=str(ref), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    nz, ny, nx = solid.shape
    co = read_checkpoint(ours / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, np.float64, False)
    cr = read_checkpoint(ref / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, np.float64, False)
    assert co["ntime"] == cr["ntime"] == 20 + 20 + 1
    assert common.relerr(co["pdf"], cr["pdf"]) <= common.TOL[prec]


@pytest.mark.gpu
def test_driver_random_initial_distribution(gpu_lib, tmp_path):
    """initial_fluid_distribution_option 6: fluid 1 with probability target_fluid1_saturation per node, drawn with rand()
    seeded by the wall clock in the reference too, so only statistics can be checked"""
    ctl, solid = common.CASES["imbibition_plate2"]()
    ctl = dict(ctl, initial_fluid_distribution_option=6, target_fluid1_saturation=0.4, max_time_step=9, monitor_timer=10, computation_time_timer=10,
               display_steps_timer=1000, benchmark_cmd=1)
    rc.write_case(tmp_path, ctl, solid)
    r = run_driver(tmp_path, "--prec", "f64")
    sat0 = float([l for l in r.stdout.splitlines() if l.startswith("Initial saturation:")][0].split(":")[1])
    assert abs(sat0 - 0.4) < 0.03, sat0
    sat = rows(tmp_path / "results" / "out1.output" / "saturation_full_domain.dat")
    assert sat.shape[0] == 1 and np.isfinite(sat).all() and 0.3 < sat[0, 1] < 0.5
    assert (tmp_path / "job_status.txt").read_text() == "simulation_reached_max_step\n"


# the shipped control file (/root/reference/input/simulation_control.txt) with the changes SURVEY.md 8d lists for BASELINE
# configs[0]: the bundled geometry's own dimensions (the shipped 240 x 240 x 260 do not match the 60 x 60 x 80 file), drainage
# (option 1, saturation_injection 1), the porous plate inside the sample, 2 000 steps monitored every 200
SHIPPED = dict(initial_fluid_distribution_option=1, benchmark_cmd=0, extreme_large_sim_cmd=0, breakthrough_check=0, steady_state_option=3,
               convergence_criteria=1e-4, output_fieldData_precision_cmd=0, modify_geometry_cmd=0, geometry_preprocess_cmd=0,
               porous_plate_cmd=2, Z_porous_plate=72, change_inlet_fluid_phase_cmd=0, n_exclude_inlet=10, n_exclude_outlet=10,
               fluid1_viscosity=0.04, fluid2_viscosity=0.4, surface_tension=0.03, theta=65, RK_beta=0.95, inlet_BC=2, outlet_BC=2,
               saturation_injection=1.0, target_inject_pore_volume=-1.0, initial_interface_position=8.0, capillary_number=100e-6,
               body_force_0=1e-4, rho_in_new=1, rho_out_BC=0, target_fluid1_saturation=0.4, max_time_step=1999, max_time_step_benchmark=100,
               ntime_visual=5000000, ntime_animation=10000, monitor_timer=200, monitor_profile_timer_ratio=5, computation_time_timer=2000,
               display_steps_timer=10000, checkpoint_save_timer=1.0, checkpoint_2rd_save_timer=5.5, simulation_duration_timer=168.0,
               d_vol_animation=-0.05, d_vol_detail=-1.0, d_vol_monitor=-0.01)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_driver_matches_stock_program_on_the_shipped_case(gpu_lib, tmp_path, prec):
    """BASELINE configs[0]: the reference's shipped drainage case on its bundled geometry (tube_sphere 60 x 60 x 80, fixture
    made from /root/reference/input/geometry/tube_sphere.dat by tests/golden/make_tube_sphere_fixture.py), 2 000 steps, run by
    the reference's stock program and by mflbm_run.  Saturation within 1e-6 at every monitor step (north_star); the checkpoint
    is compared past the 100-step horizon of the 1e-12 / 1e-5 field tolerance, hence looser."""
    stock = rc.REF_BIN_DIR / f"MF_LBM_CUDA_{prec}"
    if not stock.exists():
        pytest.skip(f"{stock} not built (oracle/build_ref.sh needs /root/reference)")
    solid = np.load(REPO / "tests" / "golden" / "tube_sphere_60_60_80.npz")["solid"]
    nz, ny, nx = solid.shape
    ours, ref = tmp_path / "ours", tmp_path / "ref"
    rc.write_case(ours, SHIPPED, solid)
    shutil.copytree(ours, ref)
    run_driver(ours, "--prec", prec)
    r = subprocess.run([str(stock)], cwd=str(ref), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    assert (ours / "job_status.txt").read_text() == (ref / "job_status.txt").read_text() == "simulation_reached_max_step\n"
    out_o, out_r = ours / "results" / "out1.output", ref / "results" / "out1.output"
    for n in ("saturation.dat", "saturation_full_domain.dat"):
        a, b = rows(out_o / n), rows(out_r / n)
        assert a.shape == b.shape and a.shape[0] >= 9, (n, a.shape, b.shape)     # one row per monitor step (200, 400, ...)
        assert np.array_equal(a[:, 0], b[:, 0])
        # column 1 is the saturation, printed with 6 significant digits by both programs: half a unit of the last place on top of
        # the 1e-6 criterion.  The other columns are the volume / mass sums behind it (~1e5): the reference accumulates them
        # sequentially in T_P (src/Monitor.cpp:52-80), which in single precision moves ITS sums by a few units
        assert np.abs(a[:, 1] - b[:, 1]).max() <= 1e-6 + 5e-7, (n, a[:, 1], b[:, 1])
        assert (np.abs(a[:, 2:] - b[:, 2:]) <= (1e-8 if prec == "f64" else 1e-4) * np.maximum(1.0, np.abs(b[:, 2:])) + 5e-6 * np.abs(b[:, 2:])).all(), (n, a, b)
    rt = np.float32 if prec == "f32" else np.float64
    co = read_checkpoint(ours / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, rt, False)
    cr = read_checkpoint(ref / "results" / "out2.checkpoint" / "id0000", nx, ny, nz, rt, False)
    assert co["ntime"] == cr["ntime"] and co["force_z"] == cr["force_z"] and co["rho_in"] == cr["rho_in"]
    assert common.relerr(co["pdf"], cr["pdf"]) <= (1e-9 if prec == "f64" else 1e-3), common.relerr(co["pdf"], cr["pdf"])


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["pack_velocity", "tube_pressure", "periodic_phasefield"])
def test_driver_on_two_gpus_writes_what_one_gpu_writes(gpu_lib, tmp_path, name, prec):
    """mflbm_run --gpus 2: the lattice cut into two x-slabs, one per GPU, driven from the one host process through peer pointers
    (host/domain.hpp).  Checkpoint and VTK files must be byte-identical to the single-GPU run's (the kernels are bit-identical
    under decomposition), the monitor files equal to the printed precision (sums combined in a different order)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case, over = STOCK_CASES[name]
    ctl, solid = common.CASES[case]()
    ctl = dict(ctl, **over)
    one, two = tmp_path / "one", tmp_path / "two"
    rc.write_case(one, ctl, solid)
    shutil.copytree(one, two)
    run_driver(one, "--prec", prec)
    r = run_driver(two, "--prec", prec, "--gpus", "2")
    assert "2 x-slabs" in r.stdout
    assert (one / "job_status.txt").read_text() == (two / "job_status.txt").read_text()
    files = sorted(str(q.relative_to(one)) for q in (one / "results").rglob("*") if q.is_file())
    assert files == sorted(str(q.relative_to(two)) for q in (two / "results").rglob("*") if q.is_file())
    for rel in files:
        a, b = (one / rel).read_bytes(), (two / rel).read_bytes()
        if rel.endswith(".vtk") or "out2.checkpoint" in rel or rel.endswith("info.txt"):
            assert a == b, rel
        elif rel.endswith("time.dat"):
            continue
        else:
            ra, rb = rows(one / rel), rows(two / rel)
            assert ra.shape == rb.shape, rel
            assert np.allclose(ra, rb, rtol=1e-5 if prec == "f64" else 1e-3, atol=1e-9 if prec == "f64" else 1e-5), rel


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_driver_restart_on_two_gpus(gpu_lib, tmp_path, prec):
    """continue_simulation on two x-slabs from a checkpoint written by a single-GPU run (and the other way round): the windows
    cut out of the global checkpoint arrays (own + ghost columns, host/domain.hpp) must reproduce the single-GPU continuation
    byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctl, solid = common.CASES["pack_velocity"]()
    base = dict(ctl, max_time_step=19, monitor_timer=10, computation_time_timer=20, display_steps_timer=1000, benchmark_cmd=1)
    nz, ny, nx = solid.shape
    second = dict(base, max_time_step=40, nxGlobal=nx, nyGlobal=ny, nzGlobal=nz, external_geometry_read_cmd=1)
    finals = {}
    for first_gpus in ("1", "2"):
        first = tmp_path / f"first{first_gpus}"
        rc.write_case(first, base, solid)
        run_driver(first, "--prec", prec, "--gpus", first_gpus)
        for gpus in ("1", "2"):
            d = tmp_path / f"cont{first_gpus}{gpus}"
            shutil.copytree(first, d)
            (d / "input" / "simulation_control.txt").write_text(rc.control_text(second))
            (d / "input" / "job_status.txt").write_text("continue_simulation")
            run_driver(d, "--prec", prec, "--gpus", gpus)
            finals[(first_gpus, gpus)] = (d / "results" / "out2.checkpoint" / "id0000").read_bytes()
    ref = finals[("1", "1")]
    for k, v in finals.items():
        assert v == ref, k
