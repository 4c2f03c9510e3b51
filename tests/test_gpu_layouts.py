"""The sector-packed table layouts change WHERE values live, not WHAT is computed."""
import os
import subprocess
import sys

import pytest
import torch

from genefaceplusplus_b200 import scene as scn
from helpers import build_model, lively_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from genefaceplusplus_b200 import scene as scn
from helpers import build_model, lively_state
sc = scn.Scene(H=64, W=64, T=3, torso=True, density_scale=8.0, table_decay=1.0, table_amp=1.0)
m = build_model(sc, lively_state(sc.state, 3.0), precision=sys.argv[1])
fi = sc.frame_inputs(1)
out = m.render(fi["rays_o"].cuda(), fi["rays_d"].cuda(), fi["cond"].cuda(), fi["bg_coords"].cuda(), fi["poses"].cuda(), bg_color=fi["bg_color"].cuda(), T_thresh=0.01, **sc.hparams)
torch.save({k: out[k].cpu() for k in ("rgb_map", "weights_sum", "depth_map")}, sys.argv[2])
"""


def _render(precision, env, tmp):
    path = os.path.join(tmp, f"out_{precision}_{len(env)}.pt")
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests")), precision, path], check=True, env=e, timeout=300)
    return torch.load(path)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_quad_layout_is_bit_identical_to_the_reference_layout(precision, tmp_path):
    a = _render(precision, {}, str(tmp_path))
    b = _render(precision, {"GFPP_NO_QUADS": "1"}, str(tmp_path))
    for k in a:
        assert torch.equal(torch.nan_to_num(a[k]), torch.nan_to_num(b[k])), k


def test_fp16_oct_tables_stay_within_the_fp16_mode_budget(tmp_path):
    a = _render("fp16", {}, str(tmp_path))                       # fp16 octs (tables rounded to fp16, like the reference's autocast)
    b = _render("fp16", {"GFPP_NO_OCTS": "1"}, str(tmp_path))   # fp32 quads, fp16 MMA operands only
    d = (a["rgb_map"] - b["rgb_map"]).abs().max().item()
    print(f"fp16 octs vs fp32 quads (both with fp16 MMA operands): max |d rgb| = {d:.2e}")
    assert d < 5e-3      # lively scene: the same order as the fp16-operand rounding itself (tests/test_gpu_render_tc.py)
