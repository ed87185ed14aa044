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


def test_conv_desc_layout_matches_the_header(tmp_path):
    """struct nbp_conv_desc as the C compiler lays it out (the header compiled with gcc) == the ctypes mirror the host side fills:
    every field's offset and the total size.  New trailing fields (dot epilogue, gate) must keep a zero-filled round-1 caller valid."""
    import subprocess
    fields = [f for f, _ in _lib.ConvDesc._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "nbp_b200.h")}"', "int main(void) {"]
    prog += [f'  printf("{f} %zu\\n", offsetof(nbp_conv_desc, {f}));' for f in fields]
    prog += ['  printf("sizeof %zu\\n", sizeof(nbp_conv_desc));', "  return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().split("\n"))
    for f in fields:
        assert int(out[f]) == getattr(_lib.ConvDesc, f).offset, f"offset of nbp_conv_desc.{f}"
    assert int(out["sizeof"]) == ctypes.sizeof(_lib.ConvDesc)
    # the round-2 additions sit behind the round-1 fields
    assert getattr(_lib.ConvDesc, "dot_w").offset > getattr(_lib.ConvDesc, "pool_lo_off").offset
    # argument errors of the new epilogue modes are reported without a GPU
    d = _lib.ConvDesc()
    d.precise = 1; d.src0 = 64; d.weight = 64; d.scale = 64; d.shift = 64; d.c0 = 64; d.ld0 = 128; d.lo0 = 64
    d.n = 1; d.h = 16; d.w = 16; d.taps = 1; d.c_out = 256; d.dot_out = 64; d.dot_w = 64
    rc = _lib.lib().nbp_conv_fwd(ctypes.byref(d), None)
    assert rc == -1 and b"dot epilogue" in _lib.lib().nbp_last_error()


def test_workspace_size_queries():
    """Size queries are pure host arithmetic: the back-projection key scratch holds 4 bytes per pixel and frame on top of the
    selection records, and the combined workspace is what ops.backproject_append allocates for the sub-sampling path."""
    L = _lib.lib()
    base = L.nbp_backproject_workspace_bytes(1024)
    keys = L.nbp_backproject_key_cache_bytes(1024, 256, 456)
    assert base >= 32 * 1024 and 4 * 1024 * 256 * 456 <= keys <= 4 * 1024 * 256 * 456 + 512
    assert L.nbp_backproject_key_cache_bytes(-1, 256, 456) == 0 and L.nbp_backproject_key_cache_bytes(4, 0, 456) == 0
