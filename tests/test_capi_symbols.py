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


def test_round2_entry_points_validate_before_any_cuda_call():
    """SR / torso-SR / training-side entry points (gfpp.h sections C and D): null pointers, unsupported shapes, foreign handles and
    short buffers are reported through the status code + gfpp_last_error(), without a GPU."""
    L = _capi.lib()
    P = ctypes.c_void_p
    assert L.gfpp_sr_packed_bytes() > 900_000 and L.gfpp_sr_workspace_bytes(1, 256) > 60_000_000
    assert L.gfpp_torso_sr_packed_bytes() > 40_000 and L.gfpp_torso_sr_workspace_bytes(8) >= 8 * 96 * 4
    assert L.gfpp_march_rays_train_scratch_bytes(4096) >= 8
    assert L.gfpp_sr_pack(None, None, 0, None, None) == -1 and b"null" in L.gfpp_last_error()
    d, m = _capi.SrDesc(), _capi.SrModel()
    assert L.gfpp_sr_pack(ctypes.byref(d), P(4096), 10, ctypes.byref(m), None) == -1            # null weights
    assert L.gfpp_sr_forward(ctypes.byref(m), 1, 256, P(4096), None, 0, P(4096), 0, P(4096), 1 << 30, None) == -1   # handle never packed
    assert b"gfpp_sr_pack" in L.gfpp_last_error()
    td, tm, tf = _capi.TorsoSrDesc(), _capi.TorsoSrModel(), _capi.TorsoSrFrames()
    assert L.gfpp_torso_sr_pack(ctypes.byref(td), P(4096), 1 << 20, ctypes.byref(tm), None) == -1
    assert L.gfpp_torso_sr_composite(ctypes.byref(tm), ctypes.byref(tf), P(4096), None, None, None, None, P(4096), 1 << 20, None) == -1
    assert b"gfpp_torso_sr_pack" in L.gfpp_last_error()
    assert L.gfpp_march_rays_train(*([None] * 3), 1.0, 0.0, 16, 8, 1, 128, 64, *([None] * 9), 0, None) == -1
    assert L.gfpp_composite_rays_train_forward(*([None] * 5), 0, 0, 1e-4, *([None] * 5)) == -1
    assert L.gfpp_grid_encode_backward(P(8), P(8), P(8), P(8), P(8), 4, 3, 4, 16, 0.5, 16, None, None, 1, 0, 0, None) == -4      # C must be 2
    assert L.gfpp_grid_encode_backward(P(8), P(8), P(8), P(8), P(8), 4, 3, 2, 16, 0.5, 16, P(8), None, 1, 0, 0, None) == -1     # dy_dx without grad_inputs
    assert L.gfpp_packbits(P(8), 4, 0.5, P(8), None) == -1 and b"aligned" in L.gfpp_last_error()
    assert L.gfpp_morton3D_dilation(P(16), 9, 128, P(16), None) == -1
    assert ctypes.sizeof(_capi.SrModel) == 256 and ctypes.sizeof(_capi.TorsoSrModel) == 1024
