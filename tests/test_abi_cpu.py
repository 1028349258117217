"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from kvmatch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "kvmatch_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kvm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = header_functions()
    assert set(names) == set(_lib.EXPORTS)
    for name in names:
        assert hasattr(L, name), name
    assert L.kvm_abi_version() == 4


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.load()
    h = ctypes.c_void_p()
    rc = L.kvm_create(ctypes.byref(h), 0)
    assert rc == _lib.KVM_E_NODEVICE
    assert b"no CPU fallback" in L.kvm_last_error(None)
    from kvmatch_b200 import GpuSeries, KvmError
    with pytest.raises(KvmError):
        GpuSeries(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kvmatch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "kvm_oracle" not in text and "libkvm_oracle" not in text, f
