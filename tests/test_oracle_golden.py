"""The self-contained CPU oracle vs the committed golden fixtures.

tests/golden/*.npz were produced by the REFERENCE'S OWN RADNeRF / RADNeRFTorso.render() running on CPU
(oracle/validate_against_reference.py --write-golden, through oracle/ref_shim.py).  These tests do not need
/root/reference: they regenerate the synthetic scene from its seeds and check that oracle/render.py
reproduces the reference's outputs, i.e. that the oracle is pinned."""
import glob
import json
import os
import re

import numpy as np
import pytest
import torch

from genefaceplusplus_b200 import scene as scn

# render fixtures only (cond_blink.npz / sr_head.npz hold conditioning vectors and SR-head outputs: tests/test_host_logic.py)
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if re.match(r"(head|torso)\d+_ms", os.path.basename(p)))


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


def test_oracle_torso_sr_reproduces_reference_golden(oracle_ops):
    """the torso-SR model (jaw-landmark + head-aware torso field, eye-blink conditioning, SR head) --
    tests/golden/torso_sr256.npz holds the REFERENCE's RADNeRFTorsowithSR.render outputs (oracle/make_torso_sr_golden.py)."""
    from genefaceplusplus_b200.config import may_hparams
    from oracle.render import OracleModel
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "torso_sr256.npz"))
    m = _meta(z)
    hp = may_hparams(**m["overrides"])
    sc = scn.Scene(H=256, W=256, T=8, torso=True, density_scale=m["density_scale"])
    t = m["frame"]
    fi = sc.frame_inputs(t)
    fi["cond"] = scn.cond_window(sc.cond, t, 3)
    lm68 = scn.lm68_sequence(8)[t].reshape(1, 136)
    state = scn.make_torso_sr_state(hp)
    assert len(state) == m["n_state_keys"]
    orc = OracleModel(state, hp)
    orc.density_scale = m["density_scale"]
    out = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=t, bg_color=fi["bg_color"],
                     T_thresh=sc.T_thresh, lm68=lm68, eye_area_percent=torch.tensor([[m["eye"]]]), upscale_torso=True,
                     **{**hp, "max_steps": 16})
    assert out["stats"]["S"] == m["stats"]["S"] and out["stats"]["P"] == m["stats"]["P"] and out["stats"]["schedule"] == [tuple(x) for x in m["stats"]["schedule"]]
    for k, (a, b, c, d) in m["crops"].items():
        assert (out[k][0, :, a:b, c:d] - torch.from_numpy(z[f"{k}_crop"])).abs().max().item() < 5e-6, k
        assert np.allclose(out[k].double().sum(dim=(0, 2, 3)).numpy(), z[f"{k}_sum"], rtol=1e-6, atol=5e-2), k
    assert abs(out["torso_alpha_map"].double().sum().item() - float(z["torso_alpha_sum"][0])) < 5e-2
    assert abs(out["deform"].double().abs().sum().item() - float(z["deform_abssum"][0])) < 5e-2
    # the blink and landmark inputs are live: changing them changes the image
    out2 = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], index=t, bg_color=fi["bg_color"],
                      T_thresh=sc.T_thresh, lm68=-lm68, eye_area_percent=None, **{**hp, "max_steps": 16})
    assert (out2["rgb_map"] - out["rgb_map"]).abs().max().item() > 1e-3


def test_head_sr_oracle_and_product_reproduce_reference_golden(oracle_ops, monkeypatch):
    """RADNeRFwithSR (radnerf_sr.py:203-210): tests/golden/head_sr256.npz holds the reference's own render() outputs
    (oracle/make_head_sr_golden.py).  Checked: the oracle's head-SR path, and the product class with its NeRF render (libgfpp,
    GPU-only, covered elsewhere) stood in for by the oracle."""
    from genefaceplusplus_b200.config import may_hparams
    from genefaceplusplus_b200.renderer import RADNeRF, RADNeRFwithSR
    from oracle.render import OracleModel
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "head_sr256.npz"))
    m = _meta(z)
    hp = may_hparams(**m["overrides"])
    sc = scn.Scene(H=256, W=256, T=8, torso=False, density_scale=m["density_scale"])
    t = m["frame"]
    fi = sc.frame_inputs(t)
    fi["cond"] = scn.cond_window(sc.cond, t, 3)
    state = scn.make_head_sr_state(hp)
    assert len(state) == m["n_state_keys"]
    eye = torch.tensor([[m["eye"]]])
    orc = OracleModel(state, hp)
    orc.density_scale = m["density_scale"]
    kw = dict(index=t, bg_color=fi["bg_color"], T_thresh=sc.T_thresh, eye_area_percent=eye)
    out = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], **kw, **{**hp, "max_steps": 16})
    assert out["stats"]["S"] == m["stats"]["S"]
    for k, (a, b, c, d) in m["crops"].items():
        assert (out[k][0, :, a:b, c:d] - torch.from_numpy(z[f"{k}_crop"])).abs().max().item() < 5e-6, k
        assert np.allclose(out[k].double().sum(dim=(0, 2, 3)).numpy(), z[f"{k}_sum"], rtol=1e-6, atol=5e-2), k
    # product class: same state with strict=True; render() = parent's NeRF render + SR head
    prod = RADNeRFwithSR(hp).eval()
    prod.load_state_dict(state, strict=True)
    plain = OracleModel(state, {**hp, "with_sr": False}, torso=False)
    plain.density_scale = m["density_scale"]

    def fake_render(self, rays_o, rays_d, cond, bg_coords, poses, *a, eye_area_percent=None, **k):
        assert eye_area_percent is not None                      # renderer.py:308: the head classes forward it
        r = plain.render(rays_o, rays_d, cond, bg_coords, poses, bg_color=fi["bg_color"], T_thresh=sc.T_thresh,
                         eye_area_percent=eye_area_percent, **{**hp, "with_sr": False, "max_steps": 16})
        return {"rgb_map": r["rgb_map"], "depth_map": r["depth_map"]}

    monkeypatch.setattr(RADNeRF, "render", fake_render)
    res = prod.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], eye_area_percent=eye)
    for k, (a, b, c, d) in m["crops"].items():
        assert (res[k][0, :, a:b, c:d] - torch.from_numpy(z[f"{k}_crop"])).abs().max().item() < 5e-6, k
