"""The C oracle against the real reference (CPU side): bit-exact set-up.

Two layers:
  * live  - when oracle/_ref/ref_cpu_* exists (built by oracle/build_ref.sh from /root/reference) the reference
            program is run on the same case directory and every array is compared bit for bit;
  * golden - SHA-256 digests of the same reference arrays, committed under tests/golden/ref_cpu_digests.json by
            tests/golden/make_golden_cpu.py, so the check also runs where the reference binaries are absent.
"""
import hashlib
import json
import tempfile
from pathlib import Path

import numpy as np
import pytest

import common
import refcase as rc

GOLDEN = Path(__file__).parent / "golden" / "ref_cpu_digests.json"
GEOM_KEYS = ["walls", "walls_type", "walls_global", "pore_profile_z", "s_nx", "s_ny", "s_nz", "W_in"]
STATE_KEYS = ["pdf", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"]
SCALARS = ["la_nui1", "la_nui2", "cos_theta", "force_z", "rho_in", "rho_out", "phi_inlet", "uin_avg", "A_xy", "A_xy_effective"]
COUNTS = ["num_solid_boundary_global", "num_fluid_boundary_global", "num_solid_boundary", "num_fluid_boundary", "pore_sum", "pore_sum_effective"]


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_digests(name: str, prec: str) -> dict:
    o, ctl, solid = common.make_oracle(name, prec)
    d = {k: digest(o.arr(k)) for k in GEOM_KEYS + STATE_KEYS}
    if ctl["outlet_BC"] == 1:
        for k in ("f_convec", "g_convec", "phi_convec"):
            d[k] = digest(o.arr(k))
    for k in SCALARS:
        d[k] = repr(o.scalar(k))
    for k in COUNTS:
        d[k] = int(o.scalar(k))
    d["saturation_full_domain"] = repr(o.cal_saturation())
    return d


def reference_digests(name: str, prec: str, workdir: Path) -> dict:
    ctl, solid = common.CASES[name]()
    full = rc.write_case(workdir, ctl, solid)
    out = rc.run_ref("cpu", prec, workdir, workdir / f"out_{prec}")
    meta = rc.read_meta(out)
    g = rc.load_geometry(out, meta)
    s = rc.load_state(out, 0, meta)
    d = {k: digest(g[k]) for k in GEOM_KEYS}
    d.update({k: digest(s[k]) for k in STATE_KEYS})
    if full["outlet_BC"] == 1:
        d.update(f_convec=digest(s["f_convec_bc"]), g_convec=digest(s["g_convec_bc"]), phi_convec=digest(s["phi_convec_bc"]))
    for k in SCALARS:
        d[k] = repr(float(meta[k]))
    for k in COUNTS:
        d[k] = int(meta[k])
    d["saturation_full_domain"] = repr(float(meta["saturation_full_domain"]))
    return d


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", sorted(common.CASES))
def test_oracle_setup_matches_golden_reference_digests(name, prec):
    gold = json.loads(GOLDEN.read_text())[f"{name}/{prec}"]
    mine = oracle_digests(name, prec)
    bad = [k for k in gold if mine.get(k) != gold[k]]
    assert not bad, f"oracle differs from the reference in {bad}"


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["tube_pressure", "pack_velocity", "rect_quirk"])
def test_oracle_setup_matches_live_reference(name, prec):
    if not rc.ref_binary("cpu", prec).exists():
        pytest.skip("oracle/_ref/ref_cpu_* not built (needs /root/reference)")
    with tempfile.TemporaryDirectory() as td:
        ref = reference_digests(name, prec, Path(td))
    mine = oracle_digests(name, prec)
    bad = [k for k in ref if mine.get(k) != ref[k]]
    assert not bad, f"oracle differs from the live reference in {bad}"
