"""BASELINE config 5: roofline sweep 256^2 -> 1024^2, 16 -> 128 samples/ray, translucent / opaque scene, 1 GPU.
Prints one JSON line per configuration: fps, valid samples S per frame, head-kernel ms/frame (CUDA events inside
libgfpp), achieved algorithmic GB/s and its fraction of the measured HBM peak, fp32-equivalent TFLOP/s."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genefaceplusplus_b200 import _capi, scene as scn  # noqa: E402
from genefaceplusplus_b200.renderer import RADNeRFTorso  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
L = _capi.lib()
L.gfpp_profile_enable(1)
for size in (256, 512, 1024):
    for ms in (16, 32, 64, 128):
        for ds in (1.0, 64.0):
            T = 16 if size < 1024 else 8
            sc = scn.Scene(H=size, W=size, T=T, torso=True, max_steps=ms, density_scale=ds)
            m = RADNeRFTorso(sc.hparams); m.load_state_dict(sc.state); m.density_scale = ds; m.mlp_precision = prec
            m = m.cuda().eval()
            poses = torch.stack([sc.pose(t) for t in range(T)]).cuda()
            feat = m.cal_cond_feat_clip(sc.cond.cuda())
            pose6 = scn.convert_poses(poses.cpu()).cuda()
            kw = dict(poses_c2w=poses, intrinsics=sc.intrinsics, H=size, W=size, pose6=pose6, bg_coords=sc.bg_coords.cuda(), bg_color=sc.bg_color.cuda(),
                      dt_gamma=sc.hparams["dt_gamma"], max_steps=ms, T_thresh=0.01, want_torso_maps=False, want_stats=True)
            for _ in range(3):
                res = m.render_frames(feat, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                res = m.render_frames(feat, **kw)
            e1.record(); torch.cuda.synchronize()
            buf = (ctypes.c_float * 4)(); L.gfpp_profile_read(buf)
            st = res["stats"].cpu()
            S = st[:, 2].float().mean().item(); N = size * size
            head = buf[0] / 1000.0 / T
            bytes_ = S * 2048 + N * 20
            print(json.dumps({"size": size, "max_steps": ms, "density_scale": ds, "precision": prec, "fps": 3 * T / (e0.elapsed_time(e1) / 1000.0),
                              "S_per_frame": S, "B_total": int(st[0, 0]), "head_ms_per_frame": head * 1000, "achieved_GBps": bytes_ / head / 1e9,
                              "hbm_frac": bytes_ / head / 1e9 / peak, "fp32_equiv_TFLOPs": S * 178944 / head / 1e12}))
            del m
            torch.cuda.empty_cache()
