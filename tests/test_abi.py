"""CPU: the C-ABI shared library loads and exports every symbol include/ucf_vit_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import os
import re

from ucf_vit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ucf_vit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ucf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _header_functions()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_header():
    assert _header_functions() == _lib.declared_symbols()


def test_abi_version_and_error_string():
    lib = _lib.lib()
    assert lib.ucf_abi_version() == 4
    assert isinstance(lib.ucf_last_error(), (bytes, type(None)))
    assert lib.ucf_launch_count() == 0 or lib.ucf_launch_count() > 0


def test_bad_arguments_are_rejected_without_a_device():
    lib = _lib.lib()
    # M = 0 is rejected before any CUDA call is made
    rc = lib.ucf_gemm_bf16(0, 0, 0, 0, 0, 0, 8, 8, 8, 8, 8, 0, 0, 0, 0, 0, 1, 0, None, None)
    assert rc == -1 and b"empty problem" in lib.ucf_last_error()
    rc = lib.ucf_attention_fwd(1, 1, 1, 1, 1, 1, 1, 4, 4, 36, *([8] * 12), 1.0, None)
    assert rc == -4 and b"head_dim 36" in lib.ucf_last_error()


def test_masking_entry_points_reject_bad_arguments_without_a_device():
    lib = _lib.lib()
    assert lib.ucf_mask_plan(1, 2, 8, 9, 1, 1, 1, None) == -1 and b"len_keep" in lib.ucf_last_error()
    assert lib.ucf_mask_plan(1, 2, 20000, 5, 1, 1, 1, None) == -4 and b"12288" in lib.ucf_last_error()
    assert lib.ucf_mask_plan(None, 0, 8, 2, None, None, None, None) == 0          # empty batch: nothing to do
    assert lib.ucf_gather_tokens(16, 16, None, None, 16, 1, 4, 4, 12, 0, 0, None) == -1
    assert b"multiple of 8" in lib.ucf_last_error()
    assert lib.ucf_gather_tokens(16, 16, None, None, 24, 1, 4, 4, 16, 0, 0, None) == -1
    assert b"16-byte aligned" in lib.ucf_last_error()
    assert lib.ucf_scatter_tokens(16, 16, 16, None, 1, 4, 4, 12, 0, None) == -1


def test_add_bcast_rejects_bad_arguments_without_a_device():
    lib = _lib.lib()
    assert lib.ucf_add_bcast(16, 16, 16, 2, 3, 4, 12, 0, 0, 0, 0, None) == -1 and b"multiples of 8" in lib.ucf_last_error()
    assert lib.ucf_add_bcast(16, 16, 16, 2, 3, 4, 16, 4, 0, 0, 0, None) == -1
    assert lib.ucf_add_bcast(16, 8, 16, 2, 3, 4, 16, 0, 0, 0, 0, None) == -1 and b"16-byte aligned" in lib.ucf_last_error()
    assert lib.ucf_add_bcast(None, None, None, 0, 3, 4, 16, 0, 0, 0, 0, None) == 0
