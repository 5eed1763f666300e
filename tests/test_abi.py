"""The C-ABI library loads on a box without a GPU and exports every symbol include/haslr_b200.h declares
(no compute call is made here). The header is parsed, so a declaration added to it without an export fails this test."""
import ctypes
import os
import re

import haslr_b200
from haslr_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "haslr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return sorted(set(re.findall(r"\b(hgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert header_functions() == sorted(ffi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(haslr_b200.lib_path())
    for name in header_functions():
        assert hasattr(lib, name), f"{name} is declared in include/haslr_b200.h but not exported"


def test_abi_version_and_error_strings_without_a_gpu():
    lib = ctypes.CDLL(haslr_b200.lib_path())
    lib.hgpu_abi_version.restype = ctypes.c_int
    assert lib.hgpu_abi_version() >= 1
    lib.hgpu_strerror.restype = ctypes.c_char_p
    lib.hgpu_strerror.argtypes = [ctypes.c_int]
    for code in (0, -1, -2, -3, -4, -5, -6):
        assert lib.hgpu_strerror(code)


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    try:
        haslr_b200.Context(0)
    except haslr_b200.HgpuError:
        return
    raise AssertionError("Context(0) must raise when there is no sm_100 device: the product has no CPU path")
