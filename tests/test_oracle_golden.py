"""The self-contained CPU oracle vs the committed golden fixtures.

tests/golden/*.npz were produced by the REFERENCE'S OWN RADNeRF / RADNeRFTorso.render() running on CPU
(oracle/validate_against_reference.py --write-golden, through oracle/ref_shim.py).  These tests do not need
/root/reference: they regenerate the synthetic scene from its seeds and check that oracle/render.py
reproduces the reference's outputs, i.e. that the oracle is pinned."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from genefaceplusplus_b200 import scene as scn

# render fixtures only (cond_blink.npz / sr_head.npz hold conditioning vectors and SR-head outputs: tests/test_host_logic.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if os.path.basename(p).startswith(("head", "torso")))


def _meta(z):
    return json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_golden(path, oracle_ops):
    from oracle.render import OracleModel
    z = np.load(path)
    m = _meta(z)
    sc = scn.Scene(H=m["size"], W=m["size"], T=8, torso=m["torso"], max_steps=m["max_steps"], density_scale=m["density_scale"])
    orc = OracleModel(sc.state, sc.hparams)
    orc.density_scale = m["density_scale"]
    for t in m["frames"]:
        fi = sc.frame_inputs(t)
        out = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"],
                         T_thresh=m["T_thresh"], **sc.hparams)
        for key in ["rgb_map", "depth_map"] + (["torso_alpha_map", "torso_rgb_map"] if m["torso"] else []):
            ref = torch.from_numpy(z[f"f{t}_{key}"])
            d = (out[key].float() - ref).abs().max().item()
            assert d <= 2e-6, (key, t, d)
        st = json.loads(bytes(z[f"f{t}_stats"]).decode())
        assert out["stats"]["schedule"] == [tuple(x) for x in st["schedule"]]
        assert out["stats"]["S"] == st["S"]


def test_golden_covers_plumbing_config():
    """BASELINE config 1: 64x64, 8 samples/ray, 4 frames, head-only."""
    names = [os.path.basename(p) for p in GOLDEN]
    assert "head64_ms8_ds1.npz" in names
    m = _meta(np.load([p for p in GOLDEN if p.endswith("head64_ms8_ds1.npz")][0]))
    assert m["size"] == 64 and m["max_steps"] == 8 and len(m["frames"]) == 4 and not m["torso"]
