"""oracle/ops.py -- TEST INFRASTRUCTURE: ctypes wrappers over oracle/_build/liborc.so.

Wrapper-level semantics follow the reference's Python wrappers:
  * march_rays: 128-row padding + zero-initialised outputs + zero noise
    (modules/radnerfs/raymarching/raymarching.py:376-392)
  * grid_encode: L-major output then permute to [B, L*C]
    (modules/radnerfs/encoders/gridencoder/grid.py:47-57)
All tensors are CPU float32/int32/uint8 torch tensors.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_b = ctypes.POINTER(ctypes.c_uint8)
u32 = ctypes.c_uint32


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc only).  Building the checker is not using it."""
    src = os.path.join(_HERE, "native_ops.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_grid_encode_forward.restype = ctypes.c_int
        _lib.orc_sh_encode_forward.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu", (t.dtype, t.is_contiguous())
    return ctypes.cast(t.data_ptr(), c_f)


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), c_i)


def _b(t):
    assert t.dtype == torch.uint8 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), c_b)


# ---------------------------------------------------------------- raymarching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32)
    fars = torch.empty(N, dtype=torch.float32)
    lib().orc_near_far_from_aabb(_f(rays_o), _f(rays_d), _f(aabb.float().contiguous()), u32(N),
                                 ctypes.c_float(min_near), _f(nears), _f(fars))
    return nears, fars


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024):
    assert not perturb, "oracle covers the inference path (perturb=False) only"
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32)
    dirs = torch.zeros(M, 3, dtype=torch.float32)
    deltas = torch.zeros(M, 2, dtype=torch.float32)
    noises = torch.zeros(n_alive, dtype=torch.float32)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _i(rays_alive), _f(rays_t), _f(rays_o), _f(rays_d),
                         ctypes.c_float(bound), ctypes.c_float(dt_gamma), u32(max_steps), u32(C), u32(H),
                         _b(density_bitfield), _f(near), _f(far), _f(xyzs), _f(dirs), _f(deltas), _f(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    sigmas = sigmas.float().contiguous()
    rgbs = rgbs.float().contiguous()
    lib().orc_composite_rays(u32(n_alive), u32(n_step), ctypes.c_float(T_thresh), _i(rays_alive), _f(rays_t),
                             _f(sigmas), _f(rgbs), _f(deltas), _f(weights_sum), _f(depth), _f(image))
    return tuple()


def packbits(grid, thresh):
    grid = grid.float().contiguous().view(-1)
    N = grid.numel() // 8
    out = torch.zeros(N, dtype=torch.uint8)
    lib().orc_packbits(_f(grid), u32(N), ctypes.c_float(thresh), _b(out))
    return out


def morton3D(coords):
    coords = coords.int().contiguous().view(-1, 3)
    out = torch.empty(coords.shape[0], dtype=torch.int32)
    lib().orc_morton3D(_i(coords), u32(coords.shape[0]), _i(out))
    return out


# ---------------------------------------------------------------- encoders
def grid_encode_forward_raw(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, gridtype, align_corners, interp):
    rc = lib().orc_grid_encode_forward(_f(inputs), _f(embeddings), _i(offsets), _f(outputs), u32(B), u32(D), u32(C),
                                       u32(L), ctypes.c_float(S), u32(H), u32(gridtype), ctypes.c_int(int(align_corners)),
                                       u32(interp))
    if rc != 0:
        # gridencoder.cu:380,397 throw std::runtime_error for unsupported C / D
        raise RuntimeError("GridEncoding: D must be 2..5 and C in {1,2,4,8}")


def grid_encode(inputs01, embeddings, offsets, per_level_scale, base_resolution, gridtype_id=1, align_corners=False,
                interp_id=0):
    """[B,D] in [0,1] -> [B, L*C]  (grid.py:24-63 without autocast / grads)."""
    inputs01 = inputs01.float().contiguous()
    B, D = inputs01.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    out = torch.empty(L, B, C, dtype=torch.float32)
    grid_encode_forward_raw(inputs01, embeddings.float().contiguous(), offsets.int().contiguous(), out, B, D, C, L, S,
                            base_resolution, gridtype_id, align_corners, interp_id)
    return out.permute(1, 0, 2).reshape(B, L * C)


def sh_encode_forward_raw(inputs, outputs, B, degree):
    rc = lib().orc_sh_encode_forward(_f(inputs), _f(outputs), u32(B), u32(degree))
    if rc != 0:
        raise RuntimeError("oracle SH encoder covers degree 1..4")


def sh_encode(dirs, degree=4):
    dirs = dirs.float().contiguous().view(-1, 3)
    out = torch.empty(dirs.shape[0], degree * degree, dtype=torch.float32)
    sh_encode_forward_raw(dirs, out, dirs.shape[0], degree)
    return out


def freq_encode_forward_raw(inputs, B, D, deg, C, outputs):
    lib().orc_freq_encode_forward(_f(inputs), u32(B), u32(D), u32(deg), u32(C), _f(outputs))


def freq_encode(x, degree):
    x = x.float().contiguous()
    B, D = x.shape
    C = D + D * 2 * degree
    out = torch.empty(B, C, dtype=torch.float32)
    freq_encode_forward_raw(x, B, D, degree, C, out)
    return out


def linear(x, W, relu=False):
    """Fixed-order fp32 bias-free linear (cond_encoder.py:183-202); x [M,K], W [N,K]."""
    x = x.float().contiguous()
    W = W.float().contiguous()
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, dtype=torch.float32)
    lib().orc_linear(_f(x), _f(W), _f(y), u32(M), u32(K), u32(N), ctypes.c_int(int(relu)))
    return y


# ---------------------------------------------------------------- training side
def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, M=None, noises=None, dt_gamma=0.0, max_steps=1024,
                     counter=None):
    """raymarching.py:184-260 at wrapper level, without the mean_count / align logic: returns (xyzs [M,3], dirs [M,3],
    deltas [M,2], rays [N,3], counter [2]); point offsets in ray order (the checker's layout, see native_ops.c)."""
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    N = rays_o.shape[0]
    M = N * max_steps if M is None else M
    xyzs = torch.zeros(M, 3); dirs = torch.zeros(M, 3); deltas = torch.zeros(M, 2)
    rays = torch.empty(N, 3, dtype=torch.int32)
    counter = torch.zeros(2, dtype=torch.int32) if counter is None else counter
    noises = torch.zeros(N) if noises is None else noises.float().contiguous()
    lib().orc_march_rays_train(_f(rays_o), _f(rays_d), _b(density_bitfield), ctypes.c_float(bound), ctypes.c_float(dt_gamma), u32(max_steps),
                               u32(N), u32(C), u32(H), u32(M), _f(nears), _f(fars), _f(xyzs), _f(dirs), _f(deltas), _i(rays), _i(counter),
                               _f(noises))
    return xyzs, dirs, deltas, rays, counter


def march_rays_train_backward(grad_xyzs, grad_dirs, rays, deltas):
    N, M = rays.shape[0], grad_xyzs.shape[0]
    go, gd = torch.zeros(N, 3), torch.zeros(N, 3)
    lib().orc_march_rays_train_backward(_f(grad_xyzs.float().contiguous()), _f(grad_dirs.float().contiguous()), _i(rays), _f(deltas), u32(N), u32(M),
                                        _f(go), _f(gd))
    return go, gd


def composite_rays_train_forward(sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
    M, N = sigmas.shape[0], rays.shape[0]
    ws, asum, depth, image = torch.empty(N), torch.empty(N), torch.empty(N), torch.empty(N, 3)
    lib().orc_composite_rays_train_forward(_f(sigmas.float().contiguous()), _f(rgbs.float().contiguous()), _f(ambient.float().contiguous()), _f(deltas),
                                           _i(rays), u32(M), u32(N), ctypes.c_float(T_thresh), _f(ws), _f(asum), _f(depth), _f(image))
    return ws, asum, depth, image


def composite_rays_train_backward(grad_ws, grad_asum, grad_image, sigmas, rgbs, ambient, deltas, rays, ws, asum, image, T_thresh=1e-4):
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr, ga = torch.zeros(M), torch.zeros(M, 3), torch.zeros(M)
    lib().orc_composite_rays_train_backward(_f(grad_ws.float().contiguous()), _f(grad_asum.float().contiguous()), _f(grad_image.float().contiguous()),
                                            _f(sigmas.float().contiguous()), _f(rgbs.float().contiguous()), _f(ambient.float().contiguous()), _f(deltas),
                                            _i(rays), _f(ws), _f(asum), _f(image), u32(M), u32(N), ctypes.c_float(T_thresh), _f(gs), _f(gr), _f(ga))
    return gs, gr, ga


def grid_encode_dydx(inputs01, embeddings, offsets, per_level_scale, base_resolution, gridtype_id=1, align_corners=False, interp_id=0):
    """dy_dx [B, L, D, C] of the forward (gridencoder.cu:198-243)."""
    inputs01 = inputs01.float().contiguous()
    B, D = inputs01.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    out = torch.empty(B, L, D, C)
    rc = lib().orc_grid_encode_dydx(_f(inputs01), _f(embeddings.float().contiguous()), _i(offsets.int().contiguous()), _f(out), u32(B), u32(D), u32(C),
                                    u32(L), ctypes.c_float(float(np.log2(per_level_scale))), u32(base_resolution), u32(gridtype_id),
                                    ctypes.c_int(int(align_corners)), u32(interp_id))
    assert rc == 0
    return out


def grid_encode_backward(grad_BLC, inputs01, embeddings, offsets, per_level_scale, base_resolution, gridtype_id=1, align_corners=False, interp_id=0,
                         dy_dx=None):
    """grad_BLC [B, L*C] (the autograd layout, grid.py:72-74) -> (grad_embeddings [sum,C], grad_inputs [B,D] or None)."""
    inputs01 = inputs01.float().contiguous()
    B, D = inputs01.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    grad = grad_BLC.float().view(B, L, C).permute(1, 0, 2).contiguous()
    ge = torch.zeros_like(embeddings, dtype=torch.float32)
    gi = torch.zeros(B, D) if dy_dx is not None else None
    rc = lib().orc_grid_encode_backward(_f(grad), _f(inputs01), _i(offsets.int().contiguous()), _f(ge), u32(B), u32(D), u32(C), u32(L),
                                        ctypes.c_float(float(np.log2(per_level_scale))), u32(base_resolution),
                                        _f(dy_dx.float().contiguous()) if dy_dx is not None else None, _f(gi) if gi is not None else None,
                                        u32(gridtype_id), ctypes.c_int(int(align_corners)), u32(interp_id))
    assert rc == 0
    return ge, gi


def grad_total_variation(inputs01, embeddings, offsets, weight, per_level_scale, base_resolution, gridtype_id=1, align_corners=False, grad=None):
    inputs01 = inputs01.float().contiguous()
    B, D = inputs01.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    grad = torch.zeros_like(embeddings, dtype=torch.float32) if grad is None else grad
    rc = lib().orc_grad_total_variation(_f(inputs01), _f(embeddings.float().contiguous()), _f(grad), _i(offsets.int().contiguous()), ctypes.c_float(weight),
                                        u32(B), u32(D), u32(C), u32(L), ctypes.c_float(float(np.log2(per_level_scale))), u32(base_resolution),
                                        u32(gridtype_id), ctypes.c_int(int(align_corners)))
    assert rc == 0
    return grad


def morton3D_invert(indices):
    indices = indices.int().contiguous().view(-1)
    out = torch.empty(indices.shape[0], 3, dtype=torch.int32)
    lib().orc_morton3D_invert(_i(indices), u32(indices.shape[0]), _i(out))
    return out


def morton3D_dilation(grid):
    grid = grid.float().contiguous()
    C, H3 = grid.shape
    H = int(round(H3 ** (1.0 / 3.0)))
    out = torch.empty_like(grid)
    lib().orc_morton3D_dilation(_f(grid), u32(C), u32(H), _f(out))
    return out


def sph_from_ray(rays_o, rays_d, radius):
    rays_o = rays_o.float().contiguous().view(-1, 3)
    rays_d = rays_d.float().contiguous().view(-1, 3)
    out = torch.empty(rays_o.shape[0], 2)
    lib().orc_sph_from_ray(_f(rays_o), _f(rays_d), ctypes.c_float(radius), u32(rays_o.shape[0]), _f(out))
    return out
