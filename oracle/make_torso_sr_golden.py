"""oracle/make_torso_sr_golden.py -- pins the oracle's torso-SR path against the reference.

Runs only where /root/reference exists.  Builds the reference's RADNeRFTorsowithSR (modules/radnerfs/radnerf_torso_sr.py) from
egs/datasets/May/lm3d_radnerf_torso_sr.yaml (with_sr, torso_head_aware, add_eye_blink_cond, eye_blink_dim 4, smo_win_size 3),
load_state_dict(strict=True)s genefaceplusplus_b200.scene.make_torso_sr_state, renders one 256x256 frame on CPU through the
reference's unmodified `render()` (native ops served by the C restatement, oracle/ref_shim.py), compares with
oracle.render.OracleModel on the same inputs and stores crops + sums of the reference's outputs in tests/golden/torso_sr256.npz.

Usage:  python -m oracle.make_torso_sr_golden
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from genefaceplusplus_b200 import scene as scn  # noqa: E402
from genefaceplusplus_b200.config import may_hparams  # noqa: E402
from oracle import ops, ref_shim  # noqa: E402
from oracle.render import OracleModel  # noqa: E402

FRAME, EYE, DS = 2, 0.37, 8.0
CROPS = {"rgb_map": (96, 160, 96, 160), "sr_rgb_map": (224, 288, 224, 288), "torso_rgb_map": (176, 240, 96, 160)}
OVERRIDES = dict(with_sr=True, torso_head_aware=True, add_eye_blink_cond=True, eye_blink_dim=4, smo_win_size=3)


def inputs():
    hp = may_hparams(**OVERRIDES)
    sc = scn.Scene(H=256, W=256, T=8, torso=True, density_scale=DS)
    fi = sc.frame_inputs(FRAME)
    fi["cond"] = scn.cond_window(sc.cond, FRAME, 3)
    lm68 = scn.lm68_sequence(8)[FRAME].reshape(1, 136)
    return hp, sc, fi, lm68


def main():
    cwd = os.getcwd()
    ops.build()
    set_hparams = ref_shim.install(ops)
    ref_hp = set_hparams("egs/datasets/May/lm3d_radnerf_torso_sr.yaml", print_hparams=False)
    from modules.radnerfs.radnerf_torso_sr import RADNeRFTorsowithSR
    with torch.no_grad():
        model = RADNeRFTorsowithSR(ref_hp).eval()
    hp, sc, fi, lm68 = inputs()
    for k, v in OVERRIDES.items():
        assert ref_hp[k] == v, (k, ref_hp[k], v)
    state = scn.make_torso_sr_state(hp)
    model.load_state_dict(state, strict=True)                      # pins key names and shapes of the torso-SR model
    model.density_scale = DS
    kw = dict(ref_hp); kw["max_steps"] = 16
    t0 = time.time()
    with torch.no_grad():
        ref = model.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=FRAME, staged=False,
                           bg_color=fi["bg_color"], perturb=False, force_all_rays=False, T_thresh=sc.T_thresh, lm68=lm68,
                           eye_area_percent=torch.tensor([[EYE]]), upscale_torso=True, **kw)
    t1 = time.time()
    orc = OracleModel(state, hp); orc.density_scale = DS
    mine = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=FRAME, bg_color=fi["bg_color"],
                      T_thresh=sc.T_thresh, lm68=lm68, eye_area_percent=torch.tensor([[EYE]]), upscale_torso=True,
                      **{**hp, "max_steps": 16})
    t2 = time.time()
    out, worst = {}, 0.0
    for k in ("rgb_map", "sr_rgb_map", "torso_rgb_map", "sr_torso_rgb_map", "torso_alpha_map", "depth_map", "deform"):
        d = (ref[k].float() - mine[k].float()).abs().max().item()
        worst = max(worst, d)
        print(f"  {k:18s} {tuple(ref[k].shape)}  max|ref-oracle| = {d:.3e}")
    for k, (a, b, c, d) in CROPS.items():
        out[f"{k}_crop"] = ref[k][0, :, a:b, c:d].numpy().astype(np.float32)
        out[f"{k}_sum"] = ref[k].double().sum(dim=(0, 2, 3)).numpy()
    out["torso_alpha_sum"] = np.asarray([ref["torso_alpha_map"].double().sum().item()])
    out["deform_abssum"] = np.asarray([ref["deform"].double().abs().sum().item()])
    meta = dict(source="reference RADNeRFTorsowithSR.render on CPU via oracle/ref_shim.py", frame=FRAME, eye=EYE, density_scale=DS,
                overrides=OVERRIDES, crops=CROPS, stats=mine["stats"], torch=torch.__version__,
                n_state_keys=len(state), ref_seconds=round(t1 - t0, 1), oracle_seconds=round(t2 - t1, 1))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.chdir(cwd)
    print(f"WORST max|reference - oracle| = {worst:.3e}; stats {mine['stats']['S']} samples, P={mine['stats']['P']}; ref {t1-t0:.1f}s oracle {t2-t1:.1f}s")
    if worst > 2e-6:
        return 1
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "torso_sr256.npz"), **out)
    print("wrote tests/golden/torso_sr256.npz")
    return 0


if __name__ == "__main__":
    sys.exit(main())
