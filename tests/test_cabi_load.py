"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/b200vs.h declares.  No compute call is made without a GPU; the entry points that
need one must fail loudly (no CPU fallback)."""
import ctypes as C
import subprocess

import pytest
import torch


def test_every_declared_symbol_is_exported(native_lib):
    from b200vs import _cabi
    declared = _cabi.declared_symbols()
    assert len(declared) >= 17, declared
    out = subprocess.run(["nm", "-D", "--defined-only", str(_cabi.LIB_PATH)], capture_output=True,
                         text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in b200vs.h but not exported: {missing}"
    for s in declared:
        assert hasattr(native_lib, s)


def test_version_and_error_string(native_lib):
    assert b"sm_100a" in native_lib.vs_version()
    assert native_lib.vs_last_error() is not None
    assert native_lib.vs_launch_count() >= 0


def test_library_contains_sm100a_sass_only(native_lib):
    from b200vs import _cabi
    out = subprocess.run(["cuobjdump", "--list-elf", str(_cabi.LIB_PATH)], capture_output=True,
                         text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_argument_validation_needs_no_gpu(native_lib):
    from b200vs import _cabi
    h = C.c_void_p()
    with pytest.raises(ValueError):
        _cabi.check(native_lib.vs_create(0, 0, 0, 0, 0, C.byref(h)))          # dim = 0
    with pytest.raises(ValueError):
        _cabi.check(native_lib.vs_create(0, 8, 7, 0, 0, C.byref(h)))          # bad metric
    with pytest.raises(ValueError):
        _cabi.check(native_lib.vs_search(None, None, 1, 1, 0, None, -1, None, None, None))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(native_lib, tmp_path):
    """Without a CUDA device the engine raises; it never computes on the CPU."""
    from b200vs import _cabi, MLXVectorStore, MLXVectorStoreConfig
    h = C.c_void_p()
    with pytest.raises(RuntimeError):
        _cabi.check(native_lib.vs_create(0, 8, 0, 0, 0, C.byref(h)))
    with pytest.raises(RuntimeError):
        MLXVectorStore(str(tmp_path / "s"), MLXVectorStoreConfig(dimension=8))
