"""CPU: the C-ABI library loads and exports every symbol include/nbp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from nextbestpath_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nbp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nbp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 8
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in nbp_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES table out of sync with the header"


def test_version_and_argument_errors_need_no_gpu():
    L = _lib.lib()
    assert L.nbp_version() == 1
    assert L.nbp_raster_workspace_bytes(4, 1000) > 1000 * 2 * 96
    # invalid arguments are rejected before any CUDA call
    rc = L.nbp_grid_scatter(None, None, 0, None, None, 0, None, None, None, 0, 3, 4, 256, -40.0, 40.0, 0, None, None)
    assert rc == -1 and b"null pointer" in L.nbp_last_error()
