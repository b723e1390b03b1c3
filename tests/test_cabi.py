"""The C-ABI library builds, loads and exports every symbol include/slm_holo.h declares
(no compute calls: there is no GPU on the CPU test box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "slm_holo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slm_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_declarations_agree():
    from spatial_light_modulator_module_b200 import _ffi
    assert header_functions() == sorted(_ffi.SIGNATURES)


def test_cuda_library_exports_every_symbol():
    from spatial_light_modulator_module_b200 import _ffi, build
    lib_path = build.build()
    assert lib_path == _ffi.LIB_PATH and os.path.exists(lib_path)
    lib = _ffi.declare(ctypes.CDLL(lib_path))
    for name in header_functions():
        assert hasattr(lib, name)
    assert lib.slm_version() >= 100
    buf = (ctypes.c_int * 32)()
    n = lib.slm_supported_lengths(buf, 32)
    assert {768, 1024, 512, 2048} <= set(buf[:n])


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    from spatial_light_modulator_module_b200 import _ffi, build
    build.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cuda_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from spatial_light_modulator_module_b200 import _ffi
    from spatial_light_modulator_module_b200.engine import Engine
    with pytest.raises(_ffi.EngineError):
        Engine((128, 128), "fp32", 1)


def test_product_never_imports_oracle_or_emulation():
    pkg = os.path.join(ROOT, "spatial_light_modulator_module_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f
                assert "libslmholo_emu" not in src, f
