"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/xfusion.h declares (no compute calls here)."""
import ctypes
import os
import re

from transfusion_b200 import _lib
from transfusion_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "xfusion.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build_library()
    assert os.path.isfile(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in xfusion.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "transfusion_b200/_lib.EXPORTS out of sync with include/xfusion.h"


def test_version_and_error_string():
    L = _lib.lib()
    assert L.xf_version() == 2
    assert isinstance(L.xf_last_error(), bytes)
    assert L.xf_launch_count() == 0


def test_invalid_arguments_are_rejected_without_a_gpu():
    L = _lib.lib()
    g = _lib.XfGemm()  # all-null descriptor
    assert L.xf_gemm(ctypes.byref(g), None) < 0
    assert b"null" in L.xf_last_error()
    a = _lib.XfAttnFwd()
    assert L.xf_attn_fwd(ctypes.byref(a), None) < 0


def test_struct_sizes_match_header_layout():
    # 64-bit pointers / int64 + packed int32/float fields, natural alignment
    assert ctypes.sizeof(_lib.XfGemm) % 8 == 0
    assert ctypes.sizeof(_lib.XfAttnFwd) % 8 == 0
    assert ctypes.sizeof(_lib.XfAttnBwd) % 8 == 0
