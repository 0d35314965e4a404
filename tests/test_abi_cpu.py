"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol
include/vidmem.h declares; argument validation and the no-fallback rule hold without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    import vidmem_b200
    lib = vidmem_b200._lib.load()
    hdr = open(os.path.join(ROOT, "include", "vidmem.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in vidmem.h but not exported"
    assert declared == set(vidmem_b200._lib.EXPORTS)
    assert lib.vm_version() == 1


def test_badarg_and_error_message():
    import vidmem_b200
    lib = vidmem_b200._lib.load()
    h = C.c_void_p()
    rc = lib.vm_store_create(C.byref(h), 0, 0, 0, 10)     # dim 0 is rejected before any CUDA call
    assert rc == vidmem_b200._lib.VM_ERR_BADARG
    assert b"dim" in lib.vm_last_error()
    assert lib.vm_ld(384) == 384 and lib.vm_ld(100) == 104 and lib.vm_ld(1) == 8


def test_no_cpu_fallback():
    import torch
    import vidmem_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        vidmem_b200.EmbeddingStore(8, 16)
    lib = vidmem_b200._lib.load()
    h = C.c_void_p()
    rc = lib.vm_store_create(C.byref(h), 0, 8, 0, 16)
    assert rc in (vidmem_b200._lib.VM_ERR_CUDA, vidmem_b200._lib.VM_ERR_UNSUPPORTED)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "real-time-brain-inspired-video-memory_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libvm_oracle" not in src, f
