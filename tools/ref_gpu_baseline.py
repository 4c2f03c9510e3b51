"""The reference's GPU path on this box ("the kernel to beat"), next to libgfpp on the same frames.

The reference's Python cannot travel to the GPU box, its CUDA extensions can (oracle/_ref, built unmodified by
oracle/build_ref.py).  This drives oracle/render.py's restatement of the reference's host loop (renderer.py:340-384: ~40
launches and a host sync per round) with those kernels, dense layers on cuBLAS through torch -- fp32, and under
torch.autocast(fp16) as the reference ships it (inference/genefacepp_infer.py:458) -- and reports frames/s for both beside
libgfpp's, plus image differences.  The reference kernels march with nvcc's FMA contraction and `__expf`, libgfpp with the
source-level rounding of the CPU oracle, so a handful of rays gain or lose ONE boundary sample ("cell flips":
up to alpha x |colour - background| ~ 0.05 at density_scale 8): differences are therefore reported as max-abs, the number of
pixels beyond 1e-3, and PSNR.

    python tools/ref_gpu_baseline.py [--size 512] [--frames 8]

`measure()` is also what bench.py calls for the `gpu_reference` object of its JSON line.
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def available():
    ref = os.path.join(ROOT, "oracle", "_ref")
    return all(os.path.exists(os.path.join(ref, n, n + ".so")) for n in ("_raymarching_face", "_gridencoder", "_shencoder", "_freqencoder"))


def _cmp(a, b):
    d = (a.float() - b.float()).abs().reshape(-1, 3).max(-1).values
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return {"max_abs": d.max().item(), "n_over_1e-3": int((d > 1e-3).sum().item()), "n_pixels": d.numel(),
            "psnr": 999.0 if mse == 0 else 10 * math.log10(1.0 / mse)}


def measure(size=512, frames=8, density_scale=8.0, precisions=("fp32", "fp16"), torso=True):
    from genefaceplusplus_b200 import scene as scn
    from genefaceplusplus_b200.renderer import RADNeRF, RADNeRFTorso
    from oracle import gpu_ref_ops
    from oracle.render import OracleModel
    dev = torch.device("cuda", torch.cuda.current_device())
    sc = scn.Scene(H=size, W=size, T=max(frames, 8), torso=torso, density_scale=density_scale)
    ref = OracleModel(sc.state, sc.hparams, backend=gpu_ref_ops, device=dev, collect_stats=False)
    ref.density_scale = density_scale
    fis = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.frame_inputs(t).items()} for t in range(frames)]

    def run_ref(autocast):
        outs = []
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            for fi in fis:
                outs.append(ref.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"],
                                       T_thresh=sc.T_thresh, **sc.hparams)["rgb_map"].float().reshape(-1, 3))
        return torch.stack(outs)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        return out, frames / (e0.elapsed_time(e1) / 1000.0)

    img32, fps32 = timed(lambda: run_ref(False))
    img16, fps16 = timed(lambda: run_ref(True))
    line = {"what": "the reference's own CUDA extensions (oracle/_ref, built unmodified) under its host loop restated in oracle/render.py, "
                    "dense layers on cuBLAS via torch; same scene, same box, same run",
            "size": size, "frames": frames, "density_scale": density_scale,
            "fp32_fps": fps32, "fp16_autocast_fps": fps16, "ref_fp16_vs_ref_fp32": _cmp(img16, img32)}
    poses = torch.stack([sc.pose(t) for t in range(frames)])
    kw = dict(cond_seq=sc.cond[:max(frames, 8)], bg_color=sc.bg_color, bg_coords=sc.bg_coords, T_thresh=sc.T_thresh)
    for prec in precisions:
        m = (RADNeRFTorso if torso else RADNeRF)(sc.hparams)
        m.load_state_dict(sc.state); m.density_scale = density_scale; m.mlp_precision = prec
        m = m.to(dev).eval()
        out, fps = timed(lambda: m.render_clip(poses, sc.intrinsics, size, size, **kw)[:frames])
        line[f"ours_{prec}_fps_same_{frames}_frames"] = fps
        line[f"ours_{prec}_vs_ref_fp32"] = _cmp(out.reshape(frames, -1, 3), img32)
        if prec != "fp32":
            line[f"ours_{prec}_vs_ref_fp16_autocast"] = _cmp(out.reshape(frames, -1, 3), img16)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--density-scale", type=float, default=8.0)
    a = ap.parse_args()
    print(json.dumps(measure(a.size, a.frames, a.density_scale)))


if __name__ == "__main__":
    main()
