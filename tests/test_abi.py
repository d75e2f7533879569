"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/segclip_b200.h declares, and the ctypes struct layouts equal the C compiler's."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from segclip_b200 import build
    build.build()
    from segclip_b200 import _lib
    return _lib


def test_library_loads_and_exports_all_declared_symbols(L):
    lib = L.lib()          # raises if a declared prototype is missing from the .so
    assert lib.sc_abi_version() == 1
    assert len(L.FUNCS) >= 30
    for name in L.FUNCS:
        assert hasattr(lib, name)


def test_struct_layouts_match_c_compiler(L, tmp_path):
    names = sorted(L.STRUCTS)
    src = '#include <stdio.h>\n#include "segclip_b200.h"\nint main(void){\n'
    for n in names:
        src += '  printf("%s %%zu\\n", sizeof(%s));\n' % (n, n)
    src += "  return 0; }\n"
    c = tmp_path / "sizes.c"
    c.write_text(src)
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(out[n]) == ctypes.sizeof(L.STRUCTS[n]), n


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    from segclip_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsegclip_b200.so")
    with pytest.raises(_lib.SegclipB200Error):
        _lib.lib()


def test_error_reporting_is_textual(L):
    import ctypes as C
    d = L.GemmDesc()          # all-null descriptor
    rc = L.lib().sc_gemm(C.byref(d), None)
    assert rc < 0 and b"null" in L.lib().sc_last_error()
