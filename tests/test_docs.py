"""Documentation that is meant to be compiled stays in step with what is compiled."""
import re
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def test_integration_listing_is_the_compiled_shim():
    """INTEGRATION.md section 2 shows integration/mflbm_shim.cpp (the reference-side binding oracle/build_ref.sh builds and
    tests/test_host_driver.py runs): the listing and the file must not drift apart (header comment aside)."""
    md = (REPO / "INTEGRATION.md").read_text()
    listing = re.search(r"```cpp\n(.*?)```", md, re.S).group(1)
    src = (REPO / "integration" / "mflbm_shim.cpp").read_text()
    body = lambda t: [l.rstrip() for l in t.splitlines() if l.strip() and not l.lstrip().startswith("//")]
    assert body(listing) == body(src)


def test_every_documented_switch_is_read_by_the_library():
    md = (REPO / "INTEGRATION.md").read_text()
    cu = (REPO / "mf-lbm-cuda_b200" / "csrc" / "mflbm.cu").read_text()
    documented = set(re.findall(r"`(MFLBM_[A-Z_]+)", md))
    read = set(re.findall(r'getenv\("(MFLBM_[A-Z_]+)"\)', cu))
    assert documented == read, (documented ^ read)
