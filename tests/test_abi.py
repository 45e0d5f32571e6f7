"""CPU checks of the drop-in boundary: the C-ABI shared library builds/loads here (nvcc cross-compiles for sm_100a)
and exports exactly the symbols include/sibrar_b200.h declares; the ctypes prototypes cover all of them."""
import ctypes
import os
import re

from sibrar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sibrar_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sbr_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.lib()
    assert lib.sbr_version() >= 100
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms, set(_lib.EXPORTS) ^ set(syms)


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    n = ctypes.c_int64(0)
    assert lib.sbr_topk_workspace_bytes(0, 10, 64, 10, 1, ctypes.byref(n)) == 1
    assert b"sbr_topk_workspace_bytes" in lib.sbr_last_error()
    assert lib.sbr_topk_workspace_bytes(1000, 5000, 64, 10, 2, ctypes.byref(n)) == 0 and n.value > 0


def test_sass_uses_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        return
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS of {_lib.LIB_PATH}"
