"""First-light check of the row-owner head kernel (head_v2_kernel.cu) against the CPU oracle: default and lively scenes, fp16 and robust.
    timeout 300 python tools/v2_check.py [size]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from genefaceplusplus_b200 import scene as scn
from helpers import build_model, lively_state, parity_report
from oracle import ops
from oracle.render import OracleModel
ops.build()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, torso, ds, gain, kw in (("default", True, 8.0, None, {}), ("default_ds1", False, 1.0, None, {}),
                                  ("lively", False, 1.0, 4.0, dict(table_decay=1.0, table_amp=1.0)),
                                  ("lively_ds16", False, 16.0, 4.0, dict(table_decay=1.0, table_amp=1.0))):
    sc = scn.Scene(H=size, W=size, T=4, torso=torso, density_scale=ds, **kw)
    state = lively_state(sc.state, gain) if gain else sc.state
    fi = sc.frame_inputs(1)
    orc = OracleModel(state, sc.hparams); orc.density_scale = sc.density_scale
    ref = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"], T_thresh=sc.T_thresh, **sc.hparams)
    for prec in ("fp16", "robust"):
        for v1 in ((False, True) if prec == "fp16" else (False,)):
            if v1: os.environ["GFPP_HEAD_V1"] = "1"
            else: os.environ.pop("GFPP_HEAD_V1", None)
            m = build_model(sc, state, precision=prec)
            t0 = time.time()
            out = m.render(fi["rays_o"].cuda(), fi["rays_d"].cuda(), fi["cond"].cuda(), fi["bg_coords"].cuda(), fi["poses"].cuda(),
                           bg_color=fi["bg_color"].cuda(), T_thresh=sc.T_thresh, **sc.hparams)
            torch.cuda.synchronize()
            rep = parity_report(out["rgb_map"].view(-1, 3), ref["rgb_map"].view(-1, 3), ref["knife"], knife_tol=3e-2 if gain else 1e-3)
            repw = parity_report(out["weights_sum"].view(-1), ref["weights_sum"].view(-1), ref["knife"], knife_tol=3e-2 if gain else 1e-3)
            print(f"[{name:12s} {prec:6s} {'v1' if v1 else 'v2'}] rgb max|d|={rep['max_abs']:.2e} (all {rep['max_abs_all']:.2e}, knife {rep['n_knife']}) "
                  f"psnr={rep['psnr']:.1f} alpha {repw['max_abs']:.2e}  ({time.time()-t0:.2f}s)", flush=True)
os.environ.pop("GFPP_HEAD_V1", None)
