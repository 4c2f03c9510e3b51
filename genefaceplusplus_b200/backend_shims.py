"""Level-B drop-in (SURVEY.md 8(b) "B-native"): register ctypes-backed modules under the names the reference's Python
imports its CUDA extensions by -- `_raymarching_face`, `_gridencoder`, `_shencoder`, `_freqencoder`
(modules/radnerfs/raymarching/raymarching.py:9-12, encoders/*/: `try: import _x as _backend`) -- so the UNMODIFIED reference
wrappers run on libgfpp's per-op kernels.  Call `install()` before `import modules.radnerfs`.

Only the inference exports exist (near_far_from_aabb, march_rays, composite_rays, grid_encode_forward, sh_encode_forward,
freq_encode_forward); training-only functions raise, exactly because they are out of scope.
"""
import ctypes
import sys
import types

import numpy as np
import torch

from . import _capi


def _P(t, dtype=None):
    """Device pointer with the checks the reference's pybind layer makes (CHECK_CUDA / CHECK_CONTIGUOUS / dtype): a CPU, strided
    or wrongly typed tensor raises RuntimeError here instead of faulting asynchronously inside a kernel."""
    if not torch.is_tensor(t) or not t.is_cuda:
        raise RuntimeError("libgfpp: expected a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("libgfpp: expected a contiguous tensor")
    want = dtype if dtype is not None else _DTYPES.get(t.dtype)
    if want is None or t.dtype != want:
        raise RuntimeError(f"libgfpp: unsupported dtype {t.dtype} (fp32 / int32 / uint8 buffers only; run the encoders with autocast disabled)")
    return ctypes.c_void_p(t.data_ptr())


_DTYPES = {torch.float32: torch.float32, torch.int32: torch.int32, torch.uint8: torch.uint8}
_OFFSETS_HOST = {}


def _offsets_host(offsets):
    """Host copy of a grid's level offsets, cached by (data_ptr, numel, version): no D2H sync per encoder call."""
    key = (offsets.data_ptr(), offsets.numel(), int(offsets._version))
    off = _OFFSETS_HOST.get(key)
    if off is None:
        if len(_OFFSETS_HOST) > 64:
            _OFFSETS_HOST.clear()
        off = _OFFSETS_HOST[key] = np.ascontiguousarray(offsets.detach().cpu().numpy().astype(np.int32))
    return off


def _ck(rc, what):
    if rc != 0:
        # the reference raises RuntimeError (TORCH_CHECK / std::runtime_error)
        raise RuntimeError(f"{what}: {_capi.lib().gfpp_last_error().decode()}")


def _training_only(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is a training-only op: keep the stock extension for training (SURVEY.md 2.2)")
    return f


def make_modules():
    L = _capi.lib()
    S = _capi.stream_ptr
    cf = ctypes.c_float

    rm = types.ModuleType("_raymarching_face")
    rm.near_far_from_aabb = lambda ro, rd, aabb, N, min_near, nears, fars: _ck(
        L.gfpp_near_far_from_aabb(_P(ro), _P(rd), _P(aabb), N, cf(min_near), _P(nears), _P(fars), S()), "near_far_from_aabb")
    rm.march_rays = lambda n_alive, n_step, alive, t, ro, rd, bound, dt_gamma, max_steps, C, H, grid, near, far, xyzs, dirs, deltas, noises: _ck(
        L.gfpp_march_rays(n_alive, n_step, _P(alive), _P(t), _P(ro), _P(rd), cf(bound), cf(dt_gamma), max_steps, C, H, _P(grid), _P(near),
                          _P(far), _P(xyzs), _P(dirs), _P(deltas), _P(noises), S()), "march_rays")
    rm.composite_rays = lambda n_alive, n_step, T_thresh, alive, t, sig, rgb, deltas, ws, depth, image: _ck(
        L.gfpp_composite_rays(n_alive, n_step, cf(T_thresh), _P(alive), _P(t), _P(sig), _P(rgb), _P(deltas), _P(ws), _P(depth), _P(image), S()),
        "composite_rays")
    for n in ("packbits", "sph_from_ray", "morton3D", "morton3D_invert", "morton3D_dilation", "march_rays_train", "march_rays_train_backward",
              "composite_rays_train_forward", "composite_rays_train_backward"):
        setattr(rm, n, _training_only(n))

    ge = types.ModuleType("_gridencoder")

    def grid_encode_forward(inputs, emb, offsets, outputs, B, D, C, Lv, S_, H, dy_dx, gridtype, align_corners, interp):
        if dy_dx is not None:
            raise NotImplementedError("grid_encode_forward with dy_dx is a training path")
        if emb.dtype != torch.float32:
            raise RuntimeError("libgfpp grid tables are fp32 (run the encoder with autocast disabled)")
        off = _offsets_host(offsets)
        _ck(L.gfpp_grid_encode_forward(_P(inputs, torch.float32), _P(emb), off.ctypes.data_as(ctypes.c_void_p), _P(outputs, torch.float32), B, D, C, Lv, cf(S_), H, gridtype,
                                       int(align_corners), interp, S()), "grid_encode_forward")

    ge.grid_encode_forward = grid_encode_forward
    ge.grid_encode_backward = _training_only("grid_encode_backward")
    ge.grad_total_variation = _training_only("grad_total_variation")

    sh = types.ModuleType("_shencoder")

    def sh_encode_forward(inp, out, B, D, C, dy_dx):
        if dy_dx is not None:
            raise NotImplementedError("sh_encode_forward with dy_dx is a training path")
        _ck(L.gfpp_sh_encode_forward(_P(inp), _P(out), B, D, C, S()), "sh_encode_forward")

    sh.sh_encode_forward = sh_encode_forward
    sh.sh_encode_backward = _training_only("sh_encode_backward")

    fr = types.ModuleType("_freqencoder")
    fr.freq_encode_forward = lambda inp, B, D, deg, C, out: _ck(L.gfpp_freq_encode_forward(_P(inp), B, D, deg, C, _P(out), S()), "freq_encode_forward")
    fr.freq_encode_backward = _training_only("freq_encode_backward")
    return {m.__name__: m for m in (rm, ge, sh, fr)}


def install():
    mods = make_modules()
    sys.modules.update(mods)
    return mods
