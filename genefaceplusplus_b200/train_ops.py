"""Training-side native ops on libgfpp: the wrapper-level API of the reference's
modules/radnerfs/raymarching/raymarching.py:50-344 and encoders/gridencoder/grid.py:24-164 -- same function names, argument
meaning and autograd behaviour -- over csrc/train_kernels.cu through the C-ABI (include/gfpp.h section D).

    from genefaceplusplus_b200 import train_ops as raymarching
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, counter, mean_count,
                                                            perturb, 128, force_all_rays, dt_gamma, max_steps)
    weights_sum, ambient_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, ambient, deltas, rays, T_thresh)

fp32 only (the reference casts with custom_fwd(cast_inputs=float32) as well); CUDA tensors only -- there is no CPU path.
Differences from the reference, both deliberate: march_rays_train lays the samples out in ray order (deterministic; the
reference's atomicAdd order changes from run to run, every consumer goes through `rays`), and nothing here calls
torch.cuda.empty_cache().
"""
import numpy as np
import torch
from torch.autograd import Function

from . import backend_shims

_mods = None


def _backend(name):
    global _mods
    if _mods is None:
        _mods = backend_shims.make_modules()
    return _mods[name]


def _cuda_f32(t):
    if not t.is_cuda:
        raise RuntimeError("libgfpp: expected a CUDA tensor (there is no CPU path)")
    return t.float().contiguous()


# ------------------------------------------------------------------------------------------------ update_extra_state helpers
def sph_from_ray(rays_o, rays_d, radius):
    """raymarching.py:50-78: [N,3] x2 -> [N,2] (theta, phi) in [-1,1] on the sphere of `radius`."""
    rays_o, rays_d = _cuda_f32(rays_o).view(-1, 3), _cuda_f32(rays_d).view(-1, 3)
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
    _backend("_raymarching_face").sph_from_ray(rays_o, rays_d, radius, N, coords)
    return coords


def morton3D(coords):
    """raymarching.py:81-101: int32 [N,3] -> int32 [N]."""
    coords = coords.int().contiguous()
    indices = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
    _backend("_raymarching_face").morton3D(coords, coords.shape[0], indices)
    return indices


def morton3D_invert(indices):
    """raymarching.py:103-123: int32 [N] -> int32 [N,3]."""
    indices = indices.int().contiguous()
    coords = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=indices.device)
    _backend("_raymarching_face").morton3D_invert(indices, indices.shape[0], coords)
    return coords


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:126-152: float [C, H^3] -> uint8 [C*H^3/8], bit i of byte n = grid[8n+i] > thresh."""
    grid = _cuda_f32(grid)
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    _backend("_raymarching_face").packbits(grid, N, thresh, bitfield)
    return bitfield


def morton3D_dilation(grid):
    """raymarching.py:155-178: 6-neighbour max pooling of a Morton-ordered [C, H^3] grid."""
    grid = _cuda_f32(grid)
    C, H3 = grid.shape
    H = int(round(H3 ** (1.0 / 3.0)))
    out = torch.empty_like(grid)
    _backend("_raymarching_face").morton3D_dilation(grid, C, H, out)
    return out


# ------------------------------------------------------------------------------------------------ marching + compositing
class _march_rays_train(Function):
    """raymarching.py:184-273."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024):
        rays_o, rays_d = _cuda_f32(rays_o).view(-1, 3), _cuda_f32(rays_d).view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        dev = rays_o.device
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else torch.zeros(N, dtype=torch.float32, device=dev)
        _backend("_raymarching_face").march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, N, C, H, M, _cuda_f32(nears),
                                                       _cuda_f32(fars), xyzs, dirs, deltas, rays, step_counter, noises)
        if force_all_rays or mean_count <= 0:
            m = int(step_counter[0].item())
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        ctx.save_for_backward(rays, deltas)
        ctx.mark_non_differentiable(rays)
        return xyzs, dirs, deltas, rays

    @staticmethod
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        rays, deltas = ctx.saved_tensors
        N, M = rays.shape[0], grad_xyzs.shape[0]
        grad_rays_o = torch.zeros(N, 3, device=rays.device)
        grad_rays_d = torch.zeros(N, 3, device=rays.device)
        _backend("_raymarching_face").march_rays_train_backward(_cuda_f32(grad_xyzs), _cuda_f32(grad_dirs), rays, deltas.contiguous(), N, M,
                                                                grad_rays_o, grad_rays_d)
        return (grad_rays_o, grad_rays_d) + (None,) * 13


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    """raymarching.py:276-341 (grad_depth is not propagated there either)."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, ambient, deltas = _cuda_f32(sigmas), _cuda_f32(rgbs), _cuda_f32(ambient), _cuda_f32(deltas)
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        ambient_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _backend("_raymarching_face").composite_rays_train_forward(sigmas, rgbs, ambient, deltas, rays, M, N, T_thresh, weights_sum, ambient_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, ambient_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_ambient_sum, grad_depth, grad_image):
        sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas, grad_rgbs, grad_ambient = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(ambient)
        _backend("_raymarching_face").composite_rays_train_backward(_cuda_f32(grad_weights_sum), _cuda_f32(grad_ambient_sum), _cuda_f32(grad_image), sigmas, rgbs,
                                                                    ambient, deltas, rays, weights_sum, ambient_sum, image, M, N, T_thresh, grad_sigmas,
                                                                    grad_rgbs, grad_ambient)
        return grad_sigmas, grad_rgbs, grad_ambient, None, None, None


composite_rays_train = _composite_rays_train.apply


# ------------------------------------------------------------------------------------------------ grid encoder with gradients
class _grid_encode(Function):
    """grid.py:24-88: inputs [B,D] in [0,1], embeddings [sum,C] -> [B, L*C]; gradients w.r.t. the table (vector reductions) and,
    with calc_grad_inputs, w.r.t. the inputs (through dy_dx)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0, align_corners=False,
                interpolation=0):
        inputs, embeddings = _cuda_f32(inputs), _cuda_f32(embeddings)
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=torch.float32)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=torch.float32) if calc_grad_inputs else None
        _backend("_gridencoder").grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, base_resolution, dy_dx, gridtype, align_corners,
                                                     interpolation)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, base_resolution, gridtype, interpolation, align_corners]
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, align_corners = ctx.dims
        grad = _cuda_f32(grad).view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs) if dy_dx is not None else None
        _backend("_gridencoder").grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                                                      align_corners, interpolation)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


@torch.no_grad()
def grad_total_variation(embeddings, offsets, per_level_scale, base_resolution, input_dim, weight=1e-7, inputs=None, bound=1, B=1000000, gridtype=0,
                         align_corners=False):
    """GridEncoder.grad_total_variation (grid.py:166-185): adds the TV gradient of the cells `inputs` fall into (random points when
    None) to embeddings.grad, which must exist."""
    if embeddings.grad is None:
        raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
    D, C, L = input_dim, embeddings.shape[1], offsets.shape[0] - 1
    if inputs is None:
        inputs = torch.rand(B, D, device=embeddings.device)
    else:
        inputs = (inputs + bound) / (2 * bound)
        inputs = inputs.view(-1, D)
        B = inputs.shape[0]
    _backend("_gridencoder").grad_total_variation(_cuda_f32(inputs), _cuda_f32(embeddings), embeddings.grad, offsets, weight, B, D, C, L,
                                                  float(np.log2(per_level_scale)), base_resolution, gridtype, align_corners)
