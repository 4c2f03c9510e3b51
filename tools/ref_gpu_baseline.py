"""The reference's GPU path on this box ("the kernel to beat"), next to libgfpp on the same frames.

The reference's Python cannot travel to the GPU box, its CUDA extensions can (oracle/_ref, built unmodified by
oracle/build_ref.py).  This drives oracle/render.py's restatement of the reference's host loop (renderer.py:340-384: ~40
launches and a host sync per round) with those kernels, dense layers on cuBLAS through torch -- fp32, and under
torch.autocast(fp16) as the reference ships it (inference/genefacepp_infer.py:458) -- and prints frames/s for both beside
libgfpp's, plus the max-abs difference of the images.

    python tools/ref_gpu_baseline.py [--size 512] [--frames 8]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genefaceplusplus_b200 import scene as scn  # noqa: E402
from genefaceplusplus_b200.renderer import RADNeRFTorso  # noqa: E402
from oracle import gpu_ref_ops  # noqa: E402
from oracle.render import OracleModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--density-scale", type=float, default=8.0)
    a = ap.parse_args()
    dev = torch.device("cuda")
    sc = scn.Scene(H=a.size, W=a.size, T=max(a.frames, 8), torso=True, density_scale=a.density_scale)
    ref = OracleModel(sc.state, sc.hparams, backend=gpu_ref_ops, device=dev, collect_stats=False)
    ref.density_scale = a.density_scale
    frames = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.frame_inputs(t).items()} for t in range(a.frames)]

    def run_ref(autocast):
        outs = []
        with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
            for fi in frames:
                outs.append(ref.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"],
                                       T_thresh=sc.T_thresh, **sc.hparams)["rgb_map"].float())
        return torch.stack(outs)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        return out, a.frames / (e0.elapsed_time(e1) / 1000.0)

    img32, fps32 = timed(lambda: run_ref(False))
    img16, fps16 = timed(lambda: run_ref(True))
    line = {"size": a.size, "frames": a.frames, "density_scale": a.density_scale,
            "reference_kernels_fp32_fps": fps32, "reference_kernels_fp16_autocast_fps": fps16,
            "max_abs_fp16_vs_fp32": (img16 - img32).abs().max().item()}
    for prec in ("fp32", "fp16"):
        m = RADNeRFTorso(sc.hparams); m.load_state_dict(sc.state); m.density_scale = a.density_scale; m.mlp_precision = prec
        m = m.cuda().eval()
        poses = torch.stack([sc.pose(t) for t in range(a.frames)])
        kw = dict(cond_seq=sc.cond[:max(a.frames, 8)], bg_color=sc.bg_color, bg_coords=sc.bg_coords, T_thresh=sc.T_thresh)
        out, fps = timed(lambda: m.render_clip(poses, sc.intrinsics, a.size, a.size, **kw)[:a.frames])
        line[f"libgfpp_{prec}_fps"] = fps
        line[f"max_abs_libgfpp_{prec}_vs_reference_fp32"] = (out.view_as(img32) - img32).abs().max().item()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
