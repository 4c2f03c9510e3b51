"""oracle/gpu_ref_ops.py -- TEST / BASELINE INFRASTRUCTURE: the reference's OWN compiled CUDA extensions behind the
wrapper-level API of oracle/ops.py, so that oracle/render.py's restatement of the reference's host loop can be driven by the
reference's kernels on a GPU.  This is SURVEY.md 8(d)(ii), "the kernel to beat": the reference's Python cannot travel to the
GPU box, its four extensions (built unmodified by oracle/build_ref.py into oracle/_ref/) can.

Wrapper semantics restate the reference's Python wrappers (paths relative to /root/reference):
  near_far_from_aabb   modules/radnerfs/raymarching/raymarching.py:18-48
  march_rays           :347-398   (128-row padding, zero-initialised outputs, zero noise)
  composite_rays       :401-423
  grid_encode          modules/radnerfs/encoders/gridencoder/grid.py:24-63 (L-major output, permute; half tables under autocast)
  sh_encode            modules/radnerfs/encoders/shencoder/sphere_harmonics.py:16-37
  freq_encode          modules/radnerfs/encoders/freqencoder/freq.py:14-40
Only tools/ref_gpu_baseline.py and GPU tests may import this module; the product never does.
"""
import importlib.util
import os

import numpy as np
import torch

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def _load(name):
    if name not in _mods:
        so = os.path.join(_REF, name, name + ".so")
        if not os.path.exists(so):
            raise FileNotFoundError(f"{so}: build it with `python -m oracle.build_ref` where /root/reference exists")
        spec = importlib.util.spec_from_file_location(name, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


linear = None   # dense layers stay torch (cuBLAS), exactly like the reference


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o = rays_o.contiguous().view(-1, 3)
    rays_d = rays_d.contiguous().view(-1, 3)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    _load("_raymarching_face").near_far_from_aabb(rays_o, rays_d, aabb.contiguous(), N, min_near, nears, fars)
    return nears, fars


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024):
    assert not perturb
    rays_o = rays_o.contiguous().view(-1, 3)
    rays_d = rays_d.contiguous().view(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
    noises = torch.zeros(n_alive, dtype=rays_o.dtype, device=dev)
    _load("_raymarching_face").march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                                          density_bitfield, near, far, xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    _load("_raymarching_face").composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas.float().contiguous(),
                                              rgbs.float().contiguous(), deltas, weights_sum, depth, image)
    return tuple()


def grid_encode(inputs01, embeddings, offsets, per_level_scale, base_resolution, gridtype_id=1, align_corners=False, interp_id=0):
    inputs01 = inputs01.float().contiguous()
    B, D = inputs01.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    if torch.is_autocast_enabled() and C % 2 == 0:        # grid.py:43-44
        embeddings = embeddings.to(torch.half)
    out = torch.empty(L, B, C, device=inputs01.device, dtype=embeddings.dtype)
    _load("_gridencoder").grid_encode_forward(inputs01, embeddings.contiguous(), offsets.int().contiguous(), out, B, D, C, L,
                                              float(np.log2(per_level_scale)), base_resolution, None, gridtype_id, align_corners, interp_id)
    return out.permute(1, 0, 2).reshape(B, L * C)


def sh_encode(dirs, degree=4):
    dirs = dirs.float().contiguous().view(-1, 3)
    out = torch.empty(dirs.shape[0], degree * degree, dtype=dirs.dtype, device=dirs.device)
    _load("_shencoder").sh_encode_forward(dirs, out, dirs.shape[0], 3, degree, None)
    return out


def freq_encode(x, degree):
    x = x.float().contiguous()
    B, D = x.shape
    C = D + D * 2 * degree
    out = torch.empty(B, C, dtype=x.dtype, device=x.device)
    _load("_freqencoder").freq_encode_forward(x, B, D, degree, C, out)
    return out
