"""CPU: the C-ABI library builds, loads and exports every symbol declared in
include/mcs.h; argument validation works without a GPU."""
import ctypes
import os
import re

from multicamera_stitching_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    decl = declared_symbols()
    assert len(decl) >= 10
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(_cabi.EXPORTS) == decl


def test_binding_loads_and_reports_abi():
    lib = _cabi.load()
    assert lib.mcs_abi_version() == _cabi.ABI_VERSION
    assert _cabi.launch_count() >= 0


def test_argument_validation_needs_no_gpu():
    lib = _cabi.load()
    handle = ctypes.c_void_p()
    rc = lib.mcs_plan_create(ctypes.byref(handle), 0, 3, None, None, None, None, None, 10, 10)
    assert rc == -1 and b"n_layers" in lib.mcs_last_error()
    rc = lib.mcs_match_hamming_top2(None, None, 10, None, None, 10, 33, 0.75, None, None, None, 1, None)
    assert rc == -1 and b"desc_bytes" in lib.mcs_last_error()
    rc = lib.mcs_stitch_u8(None, None, None, None, 1, None, 0, 0, None)
    assert rc == -1 and b"plan is NULL" in lib.mcs_last_error()
    rc = lib.mcs_ransac_homography(None, None, None, 8, None, 4, -1.0, None, None, None, None, 1, None)
    assert rc == -1
    assert lib.mcs_plan_destroy(None) == 0
