"""oracle/make_head_sr_golden.py -- pins the oracle's head-SR path against the reference's RADNeRFwithSR.render (radnerf_sr.py).

Runs only where /root/reference exists: egs/datasets/May/lm3d_radnerf_sr.yaml (with_sr, add_eye_blink_cond, eye_blink_dim 2, smo_win_size 3),
state genefaceplusplus_b200.scene.make_head_sr_state loaded with strict=True, one 256x256 frame on CPU, crops + sums of the
reference's outputs into tests/golden/head_sr256.npz.

Usage:  python -m oracle.make_head_sr_golden
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from genefaceplusplus_b200 import scene as scn  # noqa: E402
from genefaceplusplus_b200.config import may_hparams  # noqa: E402
from oracle import ops, ref_shim  # noqa: E402
from oracle.render import OracleModel  # noqa: E402

FRAME, EYE, DS = 5, 0.61, 8.0
CROPS = {"rgb_map": (96, 160, 96, 160), "sr_rgb_map": (224, 288, 224, 288)}
OVERRIDES = dict(with_sr=True, add_eye_blink_cond=True, eye_blink_dim=2, smo_win_size=3)


def main():
    cwd = os.getcwd()
    ops.build()
    set_hparams = ref_shim.install(ops)
    ref_hp = set_hparams("egs/datasets/May/lm3d_radnerf_sr.yaml", print_hparams=False)
    from modules.radnerfs.radnerf_sr import RADNeRFwithSR
    with torch.no_grad():
        model = RADNeRFwithSR(ref_hp).eval()
    hp = may_hparams(**OVERRIDES)
    for k, v in OVERRIDES.items():
        assert ref_hp[k] == v, (k, ref_hp[k], v)
    sc = scn.Scene(H=256, W=256, T=8, torso=False, density_scale=DS)
    fi = sc.frame_inputs(FRAME)
    fi["cond"] = scn.cond_window(sc.cond, FRAME, 3)
    state = scn.make_head_sr_state(hp)
    model.load_state_dict(state, strict=True)
    model.density_scale = DS
    kw = dict(ref_hp); kw["max_steps"] = 16
    eye = torch.tensor([[EYE]])
    with torch.no_grad():
        ref = model.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=FRAME, staged=False,
                           bg_color=fi["bg_color"], perturb=False, force_all_rays=False, T_thresh=sc.T_thresh, eye_area_percent=eye, **kw)
    orc = OracleModel(state, hp); orc.density_scale = DS
    mine = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=FRAME, bg_color=fi["bg_color"],
                      T_thresh=sc.T_thresh, eye_area_percent=eye, **{**hp, "max_steps": 16})
    out, worst = {}, 0.0
    for k in ("rgb_map", "sr_rgb_map", "depth_map"):
        d = (ref[k].float() - mine[k].float()).abs().max().item()
        worst = max(worst, d)
        print(f"  {k:12s} {tuple(ref[k].shape)}  max|ref-oracle| = {d:.3e}")
    for k, (a, b, c, d) in CROPS.items():
        out[f"{k}_crop"] = ref[k][0, :, a:b, c:d].numpy().astype(np.float32)
        out[f"{k}_sum"] = ref[k].double().sum(dim=(0, 2, 3)).numpy()
    meta = dict(source="reference RADNeRFwithSR.render on CPU via oracle/ref_shim.py", frame=FRAME, eye=EYE, density_scale=DS,
                overrides=OVERRIDES, crops=CROPS, stats=mine["stats"], n_state_keys=len(state), torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.chdir(cwd)
    print(f"WORST max|reference - oracle| = {worst:.3e}")
    if worst > 2e-6:
        return 1
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "head_sr256.npz"), **out)
    print("wrote tests/golden/head_sr256.npz")
    return 0


if __name__ == "__main__":
    sys.exit(main())
