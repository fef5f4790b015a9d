"""Generates tests/golden/ref_cpu_digests.json from the REAL reference (oracle/_ref/ref_cpu_*, built from
/root/reference by oracle/build_ref.sh).  Run here (no GPU needed):  python tests/golden/make_golden_cpu.py"""
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent / "oracle"))
import common  # noqa: E402
from test_oracle_vs_reference import reference_digests  # noqa: E402

out = {}
for name in sorted(common.CASES):
    for prec in ("f64", "f32"):
        with tempfile.TemporaryDirectory() as td:
            out[f"{name}/{prec}"] = reference_digests(name, prec, Path(td))
        print(name, prec, "ok")
(HERE / "ref_cpu_digests.json").write_text(json.dumps(out, indent=1, sort_keys=True))
