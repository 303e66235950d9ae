"""
The C-ABI shared library: builds, loads, exports every symbol include/gfb200.h declares, and its
struct layouts agree with the ctypes binding.  No compute calls (no GPU needed).
"""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from genesis_forge_b200 import _native

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "gfb200.h"


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|char\s*\*|const char\s*\*)\s*\*?\s*(gfb_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    declared = declared_functions()
    assert len(declared) >= 15, declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gfb200.h but not exported by libgfb200.so"
    assert set(_native.EXPORTS) == set(declared), set(_native.EXPORTS) ^ set(declared)


def test_dynamic_symbol_table_matches_header():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line and "gfb_" in line}
    assert set(declared_functions()) <= exported


def test_abi_version_and_struct_layouts():
    lib = _native.lib()
    assert lib.gfb_abi_version() == _native.K["GFB_ABI_VERSION"]
    assert lib.gfb_abi_sizeof(0) == ctypes.sizeof(_native.Program)
    assert lib.gfb_abi_sizeof(1) == ctypes.sizeof(_native.Buffers)
    assert lib.gfb_abi_sizeof(2) == ctypes.sizeof(_native.Report)
    assert lib.gfb_abi_sizeof(3) == ctypes.sizeof(_native.ProgramHead)
    assert lib.gfb_abi_sizeof(4) == _native.B_COUNT
    assert lib.gfb_abi_sizeof(5) == ctypes.sizeof(_native.Spawn)
    assert lib.gfb_abi_sizeof(99) == -1


def test_header_enums_are_parsed():
    K = _native.K
    assert K["GFB_B_POS"] == 0 and K["GFB_B_COUNT"] > 60
    assert K["GFB_PHASE_ALL"] == 127
    assert K["GFB_R_FEET_SLIDE"] > K["GFB_R_IS_ALIVE"] >= 1
    assert K["GFB_T_TIMEOUT"] == 1


def test_invalid_arguments_are_reported_not_crashed():
    lib = _native.lib()
    assert lib.gfb_create(0, 0, ctypes.byref(ctypes.c_void_p())) == _native.K["GFB_ERR_INVALID"] % (1 << 32) - (1 << 32)
    assert lib.gfb_set_program(None, None) < 0
    assert lib.gfb_launch_count(None) == 0
    assert lib.gfb_spawn_pose(None, None, None, 0, 0, *([None] * 12)) < 0
    host = _native.Handle(8, -1)  # host-only handle: launches are refused with a message, not a crash
    assert lib.gfb_spawn_pose(host.ptr, ctypes.byref(_native.Spawn()), None, 4, 8, *([None] * 12)) == \
        _native.K["GFB_ERR_NO_DEVICE"] % (1 << 32) - (1 << 32)
    assert b"host-only" in lib.gfb_last_error(host.ptr)
    lib.gfb_destroy(None)  # no-op


def test_sm100a_code_is_what_was_built():
    """The shipped library carries sm_100a SASS with TMA bulk copies and mbarrier instructions."""
    try:
        sass = subprocess.run(["cuobjdump", "-sass", str(_native.LIB_PATH)], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "UBLKCP" in sass, "no bulk async copy (TMA) instructions in the SASS"
    assert "SYNCS" in sass, "no mbarrier instructions in the SASS"
    assert "HMMA" not in sass and "UTCHMMA" not in sass  # nothing here is a dense contraction
