"""CPU checks of the drop-in boundary: the shared library loads (no GPU needed
for that), exports every symbol declared in include/b200vec.h and
include/nvector_b200.h, and fails loudly -- never falls back -- without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: Path, prefix: str):
    txt = header.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from sundials_b200 import _lib

    return _lib.load()


def test_library_exports_every_b200vec_symbol(lib):
    from sundials_b200 import _lib

    names = _declared(ROOT / "include" / "b200vec.h", "b200vec_")
    assert len(names) >= 55
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200vec.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_nvector_symbol(lib):
    names = _declared(ROOT / "include" / "nvector_b200.h", "N_V")
    assert len(names) >= 80
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nvector_b200.h but not exported"


def test_version_and_no_silent_cpu_fallback(lib):
    import torch

    assert b"sm_100a" in lib.b200vec_version()
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is for GPU-less hosts")
    h = C.c_void_p()
    rc = lib.b200vec_ctx_create(C.byref(h), -1, None)
    assert rc < 0 and not h.value          # loud failure, no CPU path
    assert lib.b200vec_last_error()
    from sundials_b200 import nvector as nv
    from sundials_b200._lib import B200VecError

    with pytest.raises(B200VecError):
        nv.Context()


def test_library_has_only_sm100a_code():
    import shutil
    import subprocess

    from sundials_b200 import _lib

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
