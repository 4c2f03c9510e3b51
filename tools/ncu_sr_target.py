"""tools/ncu_sr_target.py -- the smallest program that launches the SR-variant kernels, for ncu:
    default : one SR-head forward of 8 frames (k_sr_conv_in, k_sr_conv<0,1,2>) + one torso-SR composite (k_torso_sr)
    --clip  : one torso-SR clip of 16 frames with every stage in libgfpp (the launch list of that path)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from genefaceplusplus_b200 import scene as scn  # noqa: E402
from genefaceplusplus_b200.config import may_hparams  # noqa: E402
from genefaceplusplus_b200.renderer import RADNeRFTorsowithSR  # noqa: E402

T = 16 if "--clip" in sys.argv else 8
hp = may_hparams(with_sr=True, add_eye_blink_cond=True, eye_blink_dim=4, smo_win_size=3, torso_head_aware=True)
m = RADNeRFTorsowithSR(hp)
m.load_state_dict(scn.make_torso_sr_state(hp), strict=True)
m.density_scale = 8.0
m.mlp_precision = "fp16"
m = m.cuda().eval()
sc = scn.Scene(H=256, W=256, T=T, torso=True, density_scale=8.0)
lm = scn.lm68_sequence(T).cuda()
if "--clip" in sys.argv:
    poses = torch.stack([sc.pose(t) for t in range(T)]).cuda()
    for _ in range(2):
        m.render_clip(poses, sc.intrinsics, 256, 256, cond_seq=sc.cond.cuda(), bg_color=sc.bg_color.cuda(), bg_coords=sc.bg_coords.cuda(), lm68_seq=lm,
                      eye_area_percent=torch.full((T,), 0.37), max_steps=16, T_thresh=sc.T_thresh, sr_noise_mode="const")
        torch.cuda.synchronize()
else:
    rgb = torch.rand(T, 256 * 256, 3, device="cuda")
    ws = torch.rand(T, 256 * 256, device="cuda")
    for _ in range(2):
        m.sr_net.forward_native(rgb, noise_mode="const", clamp=True)
        m.torso_composite_native(rgb * 0.5, ws, lm, sc.bg_coords.cuda(), sc.bg_color.cuda(), want_maps=False)
        torch.cuda.synchronize()
