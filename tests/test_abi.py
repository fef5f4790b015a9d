"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol include/mflbm.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def declared_symbols():
    h = (REPO / "include" / "mflbm.h").read_text()
    names = set(re.findall(r"mflbm_##P##_(\w+)\(", h))
    syms = [f"mflbm_{p}_{n}" for p in ("f32", "f64") for n in sorted(names)]
    syms += ["mflbm_last_error", "mflbm_version"]
    return syms


def test_library_exports_every_declared_symbol():
    import mflbm
    lib = C.CDLL(str(mflbm.LIB_PATH))   # no compute calls: there is no GPU on the CPU test box
    syms = declared_symbols()
    assert len(syms) >= 2 * 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.mflbm_version.restype = C.c_int
    assert lib.mflbm_version() == 100


def test_binding_covers_the_header():
    import mflbm
    h = (REPO / "include" / "mflbm.h").read_text()
    names = set(re.findall(r"mflbm_##P##_(\w+)\(", h))
    assert names == set(mflbm.EXPORTED), names ^ set(mflbm.EXPORTED)


def test_params_struct_layout_matches_header_order():
    import mflbm
    h = (REPO / "include" / "mflbm.h").read_text()
    body = h[h.index("typedef struct mflbm_##P##_params {"):h.index("} mflbm_##P##_params;")]
    body = re.sub(r"/\*.*?\*/", "", body)
    fields = []
    for decl in re.findall(r"(?:int64_t|int32_t|REAL)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f[0] for f in mflbm.PARAMS["f64"]._fields_]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_derived_parameters_match_oracle(prec):
    """host-side derivation (binding) == oracle == reference (the oracle is pinned by test_oracle_vs_reference)"""
    import common
    import mflbm
    for name in sorted(common.CASES):
        o, ctl, solid = common.make_oracle(name, prec)
        P = mflbm.derive_params(ctl, prec)
        for k in ("la_nui1", "la_nui2", "cos_theta", "force_z", "rho_in", "rho_out", "phi_inlet", "uin_avg", "A_xy"):
            assert float(getattr(P, k)) == o.scalar(k), (name, k, float(getattr(P, k)), o.scalar(k))
