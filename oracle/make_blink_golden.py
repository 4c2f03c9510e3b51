"""oracle/make_blink_golden.py -- golden vectors for the eye-blink conditioning branch (radnerf.py:40-47, 88-106).

Runs only where /root/reference exists.  Builds the reference's RADNeRF with `add_eye_blink_cond=True, eye_blink_dim=2`
(the values of egs/datasets/May/lm3d_radnerf_sr.yaml:28-29) on top of the May head config, gives the blink modules a seeded
random state, and records the reference's own `cal_cond_feat(cond_window, eye_area_percent)` for a few frames, together
with the conditioning sub-state, in tests/golden/cond_blink.npz.  tests/test_host_logic.py replays it through
genefaceplusplus_b200.renderer (per frame and batched over the clip) and through oracle.render.

Usage:  python -m oracle.make_blink_golden
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
from oracle import ops, ref_shim  # noqa: E402

T = 6
EYE = [0.0, 0.13, 0.5, 0.92, 0.31, 0.07]


def main():
    cwd = os.getcwd()
    ops.build()
    set_hparams = ref_shim.install(ops)
    model, hp = ref_shim.build_reference_model(set_hparams, torso=False, overrides="add_eye_blink_cond=True,eye_blink_dim=2")
    assert hp["add_eye_blink_cond"] and hp["eye_blink_dim"] == 2
    sc = scn.Scene(H=16, W=16, T=T, torso=False)
    state = dict(sc.state)
    g = torch.Generator().manual_seed(11)
    for k, v in model.state_dict().items():
        if k.startswith("blink_"):
            state[k] = torch.randn(v.shape, generator=g) * 0.3
    model.load_state_dict(state, strict=True)
    out = {}
    with torch.no_grad():
        for t in range(T):
            fi = sc.frame_inputs(t)
            out[f"f{t}_with_eye"] = model.cal_cond_feat(fi["cond"], eye_area_percent=torch.tensor([[EYE[t]]])).numpy().astype(np.float32)
            out[f"f{t}_no_eye"] = model.cal_cond_feat(fi["cond"]).numpy().astype(np.float32)
            out[f"f{t}_cond_win"] = fi["cond"].numpy().astype(np.float32)
    for k, v in state.items():
        if k.startswith(("cond_prenet.", "cond_att_net.", "blink_")):
            out["state/" + k] = v.numpy()
    out["cond_seq"] = sc.cond.numpy().astype(np.float32)
    out["eye"] = np.asarray(EYE, dtype=np.float32)
    meta = dict(source="reference RADNeRF.cal_cond_feat on CPU via oracle/ref_shim.py", overrides="add_eye_blink_cond=True,eye_blink_dim=2",
                torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cond_blink.npz"), **out)
    print("wrote tests/golden/cond_blink.npz;", "f2_with_eye[:4] =", out["f2_with_eye"][:4], " f2_no_eye[:4] =", out["f2_no_eye"][:4])


if __name__ == "__main__":
    main()
