"""Super-resolution head of the SR checkpoints: 256x256 NeRF image -> 512x512.

Drop-in for `Superresolution` of modules/radnerfs/radnerf_sr.py:15-48, which the reference assembles from StyleGAN2
synthesis blocks (modules/eg3ds/models/superresolution.py:159-257 `SynthesisBlockNoUp`,
modules/eg3ds/models/networks_stylegan2.py:286-475 `SynthesisLayer` / `ToRGBLayer` / `SynthesisBlock`).  Same parameter and
buffer names and shapes, so `load_state_dict(strict=True)` takes an SR checkpoint as is.  Written from the math, not from
the reference's code:

  * every layer is a style-modulated convolution  y = conv(x, W * s_in * d_out)  with  s = A w + b  (A scaled by
    1/sqrt(w_dim)) and, for the 3x3 layers, the demodulation  d_o = rsqrt(sum_{i,k} (W_oik s_i)^2 + 1e-8);
  * the network always feeds the constant latent w = 1 (radnerf_sr.py:33-34), so s and d are constants of the checkpoint:
    the modulated kernels are built once per call and shared by the whole batch (one plain conv2d per layer for a clip,
    instead of the reference's per-sample grouped convolutions);
  * 3x3 layers: + noise * noise_strength, + bias, leaky-ReLU(0.2) * sqrt(2), clamp to +-256;  1x1 toRGB layers: weights
    scaled by 1/sqrt(C_in), + bias, clamp; the rgb skip path is the input image (block 0) / its 2x FIR up-sampling (block 1);
  * 2x up-sampling convolution (block1.conv0): transposed convolution with stride 2, then the separable binomial filter
    [1,3,3,1]^2 / 64 with gain 4 on the (2H+1)-sized result padded by one pixel -- the polyphase identity the reference's
    conv2d_resample uses (torch_utils/ops/conv2d_resample.py:118-133).

Arithmetic is fp32 on every device (the reference switches these blocks to fp16 on CUDA; fp32 is the oracle's precision).
This is host-side PyTorch (library convolutions): plumbing around the NeRF hot path, not a hand-written kernel yet.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _binomial_filter() -> torch.Tensor:
    """[1,3,3,1] (x) [1,3,3,1], normalised to unit DC gain (upfirdn2d.setup_filter([1,3,3,1]))."""
    f = torch.tensor([1.0, 3.0, 3.0, 1.0])
    f = torch.outer(f, f)
    return f / f.sum()


def _fir(x: torch.Tensor, f: torch.Tensor, pad, gain: float) -> torch.Tensor:
    """Depth-wise FIR: pad = (left, right, top, bottom); true convolution with f * gain."""
    c = x.shape[1]
    x = F.pad(x, pad)
    k = (f * gain).flip([0, 1]).to(x.dtype)[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(x, k, groups=c)


def _upsample2x(x: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """Zero-insertion by 2 followed by the low-pass f with gain 4 (upfirdn2d.upsample2d): [N,C,H,W] -> [N,C,2H,2W]."""
    n, c, h, w = x.shape
    z = torch.zeros(n, c, h, 2, w, 2, dtype=x.dtype, device=x.device)
    z[:, :, :, 0, :, 0] = x
    return _fir(z.reshape(n, c, 2 * h, 2 * w), f, (2, 1, 2, 1), 4.0)


class _Affine(nn.Module):
    """Style affine: FullyConnectedLayer(w_dim, C, bias_init=1) of networks_stylegan2.py:99-131 with linear activation."""

    def __init__(self, w_dim, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(channels, w_dim))
        self.bias = nn.Parameter(torch.ones(channels))
        self.weight_gain = 1.0 / math.sqrt(w_dim)

    def forward(self, w):  # [w_dim] -> [C]
        return torch.addmv(self.bias, self.weight * self.weight_gain, w)


class _ModConv(nn.Module):
    """3x3 modulated + demodulated convolution layer (SynthesisLayer), optionally 2x up-sampling."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, up=1, conv_clamp=256.0):
        super().__init__()
        self.up, self.resolution, self.conv_clamp = up, resolution, conv_clamp
        self.affine = _Affine(w_dim, in_channels)
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, 3, 3))
        self.register_buffer("resample_filter", _binomial_filter())
        self.register_buffer("noise_const", torch.randn(resolution, resolution))
        self.noise_strength = nn.Parameter(torch.zeros([]))
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def modulated_weight(self, w):
        s = self.affine(w)                                                   # [I]
        k = self.weight * s.view(1, -1, 1, 1)
        d = (k.square().sum(dim=[1, 2, 3]) + 1e-8).rsqrt()                   # [O]
        return k * d.view(-1, 1, 1, 1)

    def forward(self, x, w, noise_mode="random"):
        k = self.modulated_weight(w).to(x.dtype)
        if self.up == 1:
            x = F.conv2d(x, k, padding=1)
        else:
            x = F.conv_transpose2d(x, k.transpose(0, 1), stride=2)           # [.., 2H+1, 2W+1]
            x = _fir(x, self.resample_filter, (1, 1, 1, 1), 4.0)             # -> [.., 2H, 2W]
        if noise_mode == "random":
            x = x + torch.randn(x.shape[0], 1, self.resolution, self.resolution, device=x.device, dtype=x.dtype) * self.noise_strength
        elif noise_mode == "const":
            x = x + self.noise_const * self.noise_strength
        elif noise_mode != "none":
            raise ValueError(f"noise_mode must be random / const / none, not {noise_mode!r}")
        x = F.leaky_relu(x + self.bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2.0)
        return x.clamp(-self.conv_clamp, self.conv_clamp) if self.conv_clamp is not None else x


class _ToRGB(nn.Module):
    """1x1 modulated convolution without demodulation (ToRGBLayer)."""

    def __init__(self, in_channels, out_channels, w_dim, conv_clamp=256.0):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = _Affine(w_dim, in_channels)
        self.weight = nn.Parameter(torch.randn(out_channels, in_channels, 1, 1))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.weight_gain = 1.0 / math.sqrt(in_channels)

    def forward(self, x, w):
        s = self.affine(w) * self.weight_gain
        y = F.conv2d(x, (self.weight * s.view(1, -1, 1, 1)).to(x.dtype)) + self.bias.view(1, -1, 1, 1)
        return y.clamp(-self.conv_clamp, self.conv_clamp) if self.conv_clamp is not None else y


class _Block(nn.Module):
    """conv0 -> conv1 -> toRGB added to the (possibly up-sampled) rgb skip: the 'skip' architecture of SynthesisBlock(NoUp)."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, up):
        super().__init__()
        self.up = up
        self.register_buffer("resample_filter", _binomial_filter())
        self.conv0 = _ModConv(in_channels, out_channels, w_dim, resolution, up=up)
        self.conv1 = _ModConv(out_channels, out_channels, w_dim, resolution)
        self.torgb = _ToRGB(out_channels, 3, w_dim)

    def forward(self, x, img, w, noise_mode="random"):
        x = self.conv0(x, w, noise_mode)
        x = self.conv1(x, w, noise_mode)
        if self.up == 2:
            img = _upsample2x(img, self.resample_filter)
        return x, img + self.torgb(x, w)


class Superresolution(nn.Module):
    """radnerf_sr.py:15-48.  forward(rgb [B,3,h,w] in [0,1]) -> [B,3,512,512] (not clamped; the caller clamps)."""

    def __init__(self, channels=3, img_resolution=512, sr_antialias=True):
        super().__init__()
        if img_resolution != 512:
            raise NotImplementedError("the reference's SR head is 256 -> 512 only")
        self.sr_antialias = sr_antialias
        self.input_resolution = 256
        self.w_dim = 16
        self.block0 = _Block(channels, 128, self.w_dim, 256, up=1)
        self.block1 = _Block(128, 64, self.w_dim, 512, up=2)
        self.register_buffer("resample_filter", _binomial_filter())

    def forward(self, rgb, noise_mode="random"):
        x = rgb.float()
        if x.shape[-1] < self.input_resolution:
            x = F.interpolate(x, size=(self.input_resolution, self.input_resolution), mode="bilinear", align_corners=False,
                              antialias=self.sr_antialias)
        if x.shape[-2:] != (self.input_resolution, self.input_resolution):
            raise ValueError(f"SR input must be at most {self.input_resolution}x{self.input_resolution}, got {tuple(rgb.shape)}")
        w = torch.ones(self.w_dim, dtype=x.dtype, device=x.device)
        feat, img = self.block0(x, x, w, noise_mode)
        _, img = self.block1(feat, img, w, noise_mode)
        return img
