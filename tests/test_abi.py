"""CPU: the C-ABI shared library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports
every symbol include/cwfa_b200.h declares.  No compute calls."""
import ctypes
import os
import re

from cwfa_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cwfa_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cwfa_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cwfa_b200.h but not exported"


def test_python_binding_covers_header():
    assert set(declared_symbols()) == set(_lib.exported_symbols())


def test_version_and_no_cpu_fallback():
    lib = _lib.load()
    assert b"sm_100a" in lib.cwfa_version()
    import pytest
    import torch
    from cwfa_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.haar1d_forward(torch.zeros(1, 4, 2, 2))


def test_sass_is_sm100():
    """The built objects carry sm_100a SASS only (no PTX-JIT fallback targets)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
