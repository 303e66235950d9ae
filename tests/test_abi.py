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


def test_documents_track_the_header():
    """DESIGN.md / INTEGRATION.md quote the ABI version and name every entry point of the header."""
    declared = declared_functions()
    version = _native.K["GFB_ABI_VERSION"]
    design = (ROOT / "DESIGN.md").read_text()
    integration = (ROOT / "INTEGRATION.md").read_text()
    assert f"{len(declared)} entry points, ABI v{version}" in design
    assert f"gfb_abi_version() == {version}" in integration
    missing = [name for name in declared if name not in integration]
    assert not missing, f"entry points without a row in INTEGRATION.md: {missing}"


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


def test_sass_obeys_the_uniform_datapath_rule():
    """No mbarrier / bulk-copy instruction sits in a lane-divergent region whose sibling lanes write a
    uniform register before the warp reconverges (csrc/device_utils.cuh; the bug it guards against --
    an mbarrier init value overwritten in ~15 % of the blocks -- is recorded in
    profiles/r2_01_mbarrier_init_clobber_evidence.txt).  The main library and a sample of the
    specialised kernels are checked."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_lint
    assert sass_lint.lint(_native.LIB_PATH) == []
    specs = sorted((ROOT / "genesis_forge_b200" / "_spec").glob("spec_*.so"))[:3]
    for so in specs:
        assert sass_lint.lint(so) == [], so.name


def test_sass_lint_recognises_the_hazard():
    """The lint flags the shape that failed on the GPU: thread 0 alone executes SYNCS.EXCH with uniform
    operands while the sibling lanes' path reloads one of those uniform registers."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_lint
    code = [
        (0x00, "S2R R0, SR_TID.X"),
        (0x10, "ISETP.NE.AND P1, PT, R0, RZ, PT"),
        (0x20, "@P1 BRA 0x70"),
        (0x30, "UMOV UR4, 0x1ffffe"),
        (0x40, "FENCE.VIEW.ASYNC.S"),
        (0x50, "SYNCS.EXCH.64 URZ, [UR6], UR4"),
        (0x60, "BRA 0xa0"),
        (0x70, "LDCU UR4, c[0x0][0x1944]"),
        (0x80, "UISETP.NE.AND UP0, UPT, UR4, URZ, UPT"),
        (0x90, "BRA.U UP0, 0xa0"),
        (0xa0, "BAR.SYNC.DEFER_BLOCKING 0x0"),
    ]
    problems = sass_lint.lint_kernel("k", code)
    assert len(problems) == 1 and "UR4" in problems[0]
    safe = code[:7] + [(0x70, "BSYNC.RECONVERGENT B0"), (0x80, "LDCU UR4, c[0x0][0x1944]"), (0xa0, "BAR.SYNC.DEFER_BLOCKING 0x0")]
    assert sass_lint.lint_kernel("k", safe) == []
    warp_uniform = [(0x00, "S2R R0, SR_TID.X"), (0x10, "ISETP.GT.U32.AND P1, PT, R0, 0x1f, PT")] + code[2:]
    assert sass_lint.lint_kernel("k", warp_uniform) == []
