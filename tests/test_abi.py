"""The C-ABI library loads and exports every symbol include/pww_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from paint_with_words_sd_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pww_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pww_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared_functions()
    assert set(names) == set(_native.EXPORTS), (names, _native.EXPORTS)
    L = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} missing from libpww_b200.so"


def test_metadata_calls():
    L = _native.lib()
    assert L.pww_version() >= 100
    assert L.pww_status_str(0) == b"ok" and b"unsupported" in L.pww_status_str(-2)
    assert L.pww_xattn_workspace_bytes(2, 8, 4096, 77, 40) > 0
    assert L.pww_xattn_workspace_bytes(0, 8, 4096, 77, 40) == 0


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so it is testable on the CPU box."""
    L = _native.lib()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p16 = (p + 15) // 16 * 16
    # null q
    assert L.pww_xattn_fwd_f16(None, p16, p16, p16, 1, 8, 64, 77, 40, 20480, 320, 24640, 320, 20480, 320,
                               None, 0, None, None, None, 0.158, None) == -1
    # unsupported head dim
    assert L.pww_xattn_fwd_f16(p16, p16, p16, p16, 1, 8, 64, 77, 48, 24576, 384, 29568, 384, 24576, 384,
                               None, 0, None, None, None, 0.158, None) == -2
    # misaligned pointer
    assert L.pww_xattn_fwd_f16(p16 + 2, p16, p16, p16, 1, 8, 64, 77, 40, 20480, 320, 24640, 320, 20480, 320,
                               None, 0, None, None, None, 0.158, None) == -1
    # T too large
    assert L.pww_xattn_fwd_f16(p16, p16, p16, p16, 1, 8, 64, 200, 40, 20480, 320, 64000, 320, 20480, 320,
                               None, 0, None, None, None, 0.158, None) == -2
    # stats: workspace too small
    assert L.pww_xattn_stats_f16(p16, p16, 1, 8, 64, 77, 40, 20480, 320, 24640, 320, 0, None, p16, p16, 8, None) == -4
