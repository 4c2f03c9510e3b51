"""oracle/sr_emulate.py -- TEST INFRASTRUCTURE: CPU emulation of the data flow of the native SR-head kernels.

Restates, in plain torch on the CPU, exactly what genefaceplusplus_b200/csrc/sr_kernel.cu computes from the folded GEMM
matrices of `Superresolution.folded_weights()` (the gfpp_sr_desc layouts, include/gfpp.h): NHWC activations rounded to fp16
between the layers, K ordered (tap, channel) with tap = ky*3+kx <-> pixel offset (ky-1, kx-1) and zeros outside the image,
the four output phases of the merged up-sampling layer scattered to (2y+py, 2x+px), fp32 toRGB on the unrounded
activations, the closed-form [1,3,3,1] up-sampling of the rgb skip.  Two uses:

  * on the CPU (tests/test_host_logic.py): against `Superresolution.forward` (the fp32 convolutions pinned by the reference
    golden tests/golden/sr_head.npz) -- proves the folding algebra and every layout convention of the host half;
  * on the GPU (tests/test_gpu_sr_native.py): the kernels must agree with this rounding model to ~1e-5, which separates a
    layout / descriptor / pipeline bug from the expected fp16 operand rounding (~5e-4 against the fp32 path).

Never imported by the package.
"""
import math

import torch
import torch.nn.functional as F


def _r16(x, on):
    return x.half().float() if on else x


def _act(x, noise, bias):
    if noise is not None:
        x = x + noise.reshape(noise.shape[0] if noise.dim() == 3 else 1, noise.shape[-2], noise.shape[-1], 1)
    x = F.leaky_relu(x + bias.view(1, 1, 1, -1), 0.2) * math.sqrt(2.0)
    return x.clamp(-256.0, 256.0)


def _im2col(x):
    """x [F,H,W,C] -> [F,H,W,9*C] with k = (ky*3+kx)*C + c and zeros outside the image."""
    Fn, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    return torch.cat([xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)], dim=-1)


def _upsample_skip(img):
    """img [F,h,w,3] -> [F,2h,2w,3]: even outputs 0.25*in[i-1] + 0.75*in[i], odd 0.75*in[i] + 0.25*in[i+1] per axis."""
    def up(t, dim):
        n = t.shape[dim]
        z = torch.zeros_like(t.narrow(dim, 0, 1))
        prev = torch.cat([z, t.narrow(dim, 0, n - 1)], dim)
        nxt = torch.cat([t.narrow(dim, 1, n - 1), z], dim)
        even = 0.25 * prev + 0.75 * t
        odd = 0.75 * t + 0.25 * nxt
        return torch.stack([even, odd], dim + 1).flatten(dim, dim + 1)
    return up(up(img, 1), 2)


@torch.no_grad()
def emulate(fw, rgb_flat, R, noise=(None, None, None, None), fp16=True, clamp=False, intermediates=None):
    """fw: folded_weights() dict (CPU tensors); rgb_flat [F, R*R, 3]; noise[i]: None or [res,res] / [F,res,res] planes already
    multiplied by the layer's strength.  Returns [F,3,2R,2R].  `intermediates`: optional dict that receives the tensors the
    kernels keep in their workspace (x0a, x0b: [F,R,R,128]; img0 [F,R,R,3]; x1a [F,2R,2R,64]) for layer-by-layer diagnosis."""
    x = rgb_flat.reshape(-1, R, R, 3).float()
    Fn = x.shape[0]
    h = _im2col(x) @ fw["conv_in_w"]                                        # fp32 FFMA layer, [F,R,R,128]
    h = _r16(_act(h, noise[0], fw["bias"][0]), fp16)                        # stored as fp16 NHWC
    x0a = h
    a = _act(_im2col(h) @ _r16(fw["conv0_w"], fp16).t(), noise[1], fw["bias"][1])
    img0 = x + (a @ fw["rgb_w"][0].t() + fw["rgb_b"][0]).clamp(-256.0, 256.0)   # toRGB on the unrounded activations
    h = _r16(a, fp16)
    u = _im2col(h) @ _r16(fw["up_w"], fp16).t()                             # [F,R,R,256], column (py*2+px)*64 + co
    u = u.reshape(Fn, R, R, 2, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(Fn, 2 * R, 2 * R, 64)
    x0b = h
    h = _r16(_act(u, noise[2], fw["bias"][2]), fp16)
    if intermediates is not None:
        intermediates.update(x0a=x0a, x0b=x0b, img0=img0, x1a=h)
    a = _act(_im2col(h) @ _r16(fw["conv1_w"], fp16).t(), noise[3], fw["bias"][3])
    img = _upsample_skip(img0) + (a @ fw["rgb_w"][1].t() + fw["rgb_b"][1]).clamp(-256.0, 256.0)
    if clamp:
        img = img.clamp(0.0, 1.0)
    return img.permute(0, 3, 1, 2).contiguous()
