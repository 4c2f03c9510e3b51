"""Shared helpers of the GPU parity tests."""
import math

import torch

from genefaceplusplus_b200 import scene as scn
from genefaceplusplus_b200.renderer import RADNeRF, RADNeRFTorso


def lively_state(state, gain=4.0):
    """Scale the three head MLPs so the field has O(1) dynamic range (sigma 0.1..5, colours 0..1, ambient
    coordinates spanning [-1,1]).  The default-init scene of SURVEY.md 8(d) is very bland (colours 0.49..0.51),
    which would let real bugs hide below 1e-3; the lively variant makes them glaring."""
    out = dict(state)
    for k in list(out):
        if k.startswith(("ambient_net", "sigma_net", "color_net", "torso_deform_net", "torso_canonicial_net")):
            out[k] = out[k] * gain
    return out


def build_model(sc: scn.Scene, state=None, device="cuda", precision="fp32"):
    cls = RADNeRFTorso if sc.torso else RADNeRF
    m = cls(sc.hparams)
    m.mlp_precision = precision
    m.load_state_dict(state if state is not None else sc.state, strict=True)
    m.density_scale = sc.density_scale
    return m.to(device).eval()


def psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10 * math.log10(1.0 / mse)


def parity_report(mine, ref, knife=None, knife_tol=1e-3, tol=1e-3):
    """max-abs / PSNR of `mine` vs the oracle `ref` ([N,3] or [N]).

    Rays whose transmittance came within `knife_tol` (relative) of T_thresh in the oracle are "knife-edge":
    an fp32 reordering can legitimately flip their termination, which adds or drops one sample of weight
    <= T_thresh.  They are REPORTED (count) and excluded from the max-abs bar, never silently masked;
    the PSNR is computed over ALL pixels."""
    d = (mine.float().cpu() - ref.float().cpu()).abs()
    if d.dim() > 1:
        d = d.reshape(d.shape[0], -1).max(-1).values
    d = torch.nan_to_num(d, nan=0.0)  # NaN == NaN positions (depth of rays that miss the aabb) are checked separately
    excl = torch.zeros_like(d, dtype=torch.bool) if knife is None else (knife.cpu().reshape(-1) < knife_tol)
    worst = d[~excl].max().item() if (~excl).any() else 0.0
    return {"max_abs": worst, "max_abs_all": d.max().item(), "n_knife": int(excl.sum()), "n_over": int((d > tol).sum()),
            "psnr": psnr(torch.nan_to_num(mine.float().cpu()), torch.nan_to_num(ref.float().cpu()))}
