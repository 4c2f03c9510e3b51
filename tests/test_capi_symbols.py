"""The C-ABI library loads and exports every symbol include/gfpp.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from genefaceplusplus_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gfpp.h")).read()
    return sorted(set(re.findall(r"GFPP_API\s+[\w\s\*]+?\b(gfpp_\w+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    names = _declared()
    assert len(names) >= 14, names
    L = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"libgfpp.so does not export {n}"
    assert set(names) == set(_capi.EXPORTS)


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "gfpp.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # strip comments: only declarations matter
    assert "at::" not in code and "torch" not in code.lower() and "#include <torch" not in src


def test_version_and_error_string_callable_without_gpu():
    L = _capi.lib()
    assert L.gfpp_version() == 100
    assert isinstance(L.gfpp_last_error(), bytes)


def test_argument_validation_happens_before_any_cuda_call():
    L = _capi.lib()
    # null pointers => GFPP_ERR_INVALID (-1) with a message, no crash, no GPU needed
    rc = L.gfpp_near_far_from_aabb(None, None, None, 0, 0.05, None, None, None)
    assert rc == -1 and b"null" in L.gfpp_last_error()
    rc = L.gfpp_sh_encode_forward(ctypes.c_void_p(8), ctypes.c_void_p(8), 4, 3, 9, None)
    assert rc == -4
    rc = L.gfpp_freq_encode_forward(ctypes.c_void_p(8), 4, 2, 10, 41, ctypes.c_void_p(8), None)
    assert rc == -1


def test_struct_sizes_match_header():
    # gfpp_model is 512 x uint64
    assert ctypes.sizeof(_capi.Model) == 4096
    assert ctypes.sizeof(_capi.GridDesc) == 48
