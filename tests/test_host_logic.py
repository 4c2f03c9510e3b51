"""Host-side logic on CPU: scene determinism, level layout, state_dict contract, round-schedule lemma, failure modes."""
import numpy as np
import pytest
import torch

from genefaceplusplus_b200 import scene as scn
from genefaceplusplus_b200.config import GridLayout, may_hparams, may_intrinsics
from genefaceplusplus_b200.renderer import RADNeRF, RADNeRFTorso


def test_grid_layout_matches_reference_offsets():
    # offsets probed from the reference's GridEncoder
    pos = GridLayout(3)
    assert pos.n_entries == 903480 and pos.offsets[:6].tolist() == [0, 4920, 18744, 51512, 117048, 182584]
    tor = GridLayout(2)
    assert tor.n_entries == 555520 and tor.offsets[:11].tolist() == [0, 296, 872, 1896, 3832, 7432, 14160, 26936, 50968, 96768, 162304]
    assert abs(pos.per_level_scale - 1.381912879967776) < 1e-12


def test_scene_is_deterministic_and_may_shaped():
    a, b = scn.make_state(torso=True), scn.make_state(torso=True)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert a["position_embedder.embeddings"].shape == (903480, 2)
    assert a["sigma_net.net.2.weight"].shape == (129, 128) and a["color_net.net.0.weight"].shape == (128, 148)
    occ = (a["density_grid"] > 0.5).float().mean().item()
    assert 0.017 < occ < 0.022        # ellipsoid ~1.93 % of the 128^3 cells
    n_params = sum(a[k].numel() for k in a if k.endswith("net.0.weight") or ".net." in k)
    assert a["density_bitfield"].dtype == torch.uint8 and a["density_bitfield"].numel() == 128 ** 3 // 8
    fx, fy, cx, cy = may_intrinsics(512, 512)
    assert abs(fx - 2320.0) < 1e-9 and cx == 256


def test_dropin_modules_accept_the_reference_state_dict():
    hp = may_hparams()
    st = scn.make_state(torso=True, hparams=hp)
    m = RADNeRFTorso(hp)
    missing, unexpected = m.load_state_dict(st, strict=True)
    assert not missing and not unexpected
    h = RADNeRF(hp)
    h.load_state_dict(scn.make_state(torso=False, hparams=hp), strict=True)
    # same key set as the reference module (golden meta was produced with strict=True into the reference classes)
    assert set(m.state_dict().keys()) == set(st.keys())


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU / on CPU tensors."""
    from genefaceplusplus_b200 import _capi
    sc = scn.Scene(H=8, W=8, T=2, torso=False)
    m = RADNeRF(sc.hparams).eval()
    m.load_state_dict(sc.state, strict=True)
    fi = sc.frame_inputs(0)
    with pytest.raises(_capi.GfppError):
        m.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], **sc.hparams)
    m.train()
    with pytest.raises((NotImplementedError, _capi.GfppError)):
        m.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], **sc.hparams)


def test_unsupported_configs_raise_like_the_reference_would():
    with pytest.raises(NotImplementedError):
        RADNeRF(may_hparams(cond_type="nope"))
    with pytest.raises(NotImplementedError):
        RADNeRF(may_hparams(hidden_dim_sigma=64))
    with pytest.raises(NotImplementedError):
        RADNeRFTorso(may_hparams(torso_head_aware=True))
    RADNeRF(may_hparams(add_eye_blink_cond=True, eye_blink_dim=2))   # supported since the blink branch was added


def test_cond_feat_clip_equals_per_frame(oracle_ops):
    """Batched conditioning (one launch set per clip) == the reference's per-frame windows incl. zero padding at the edges."""
    from oracle.render import OracleModel
    sc = scn.Scene(H=8, W=8, T=7, torso=False)
    m = RADNeRF(sc.hparams).eval()
    m.load_state_dict(sc.state, strict=True)
    feat = m.cal_cond_feat_clip(sc.cond)
    orc = OracleModel(sc.state, sc.hparams)
    for t in range(7):
        ref = orc.cal_cond_feat(scn.cond_window(sc.cond, t))
        assert (feat[t] - ref).abs().max().item() < 1e-5, t
        assert (m.cal_cond_feat(scn.cond_window(sc.cond, t)) - ref).abs().max().item() < 1e-5


def test_eye_blink_cond_matches_the_reference_golden():
    """add_eye_blink_cond (radnerf.py:40-47, 97-103): the golden vectors are the REFERENCE's own
    cal_cond_feat outputs (oracle/make_blink_golden.py).  Replayed per frame, batched over the clip, and through the oracle."""
    import os
    from oracle.render import OracleModel
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cond_blink.npz"))
    hp = may_hparams(add_eye_blink_cond=True, eye_blink_dim=2)
    m = RADNeRF(hp).eval()
    sub = {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    assert {"blink_embedding.weight", "blink_encoder.0.weight", "blink_encoder.0.bias", "blink_encoder.1.weight",
            "blink_encoder.1.bias"} <= set(sub) and set(sub) <= set(m.state_dict())       # same parameter names as the reference
    missing, unexpected = m.load_state_dict(sub, strict=False)
    assert not unexpected
    cond_seq, eye = torch.from_numpy(g["cond_seq"]), torch.from_numpy(g["eye"])
    T = eye.shape[0]
    with torch.no_grad():
        clip = m.cal_cond_feat_clip(cond_seq[:T] if cond_seq.shape[0] == T else cond_seq, eye_area_percent=eye)
        clip0 = m.cal_cond_feat_clip(cond_seq)
    orc = OracleModel(sub, hp)
    worst = 0.0
    for t in range(T):
        win = torch.from_numpy(g[f"f{t}_cond_win"])
        ref_eye, ref_no = torch.from_numpy(g[f"f{t}_with_eye"]), torch.from_numpy(g[f"f{t}_no_eye"])
        with torch.no_grad():
            got = m.cal_cond_feat(win, eye_area_percent=eye[t].reshape(1, 1))
            got0 = m.cal_cond_feat(win)
        for a, b in ((got, ref_eye), (got0, ref_no), (clip[t], ref_eye), (clip0[t], ref_no),
                     (orc.cal_cond_feat(win, eye_area_percent=eye[t]), ref_eye), (orc.cal_cond_feat(win), ref_no)):
            worst = max(worst, (a - b).abs().max().item())
        assert (ref_eye - ref_no).abs().max().item() > 1e-3 or eye[t] == 0      # the branch is live in the golden data
    assert worst < 2e-6, worst
    # the non-SR torso class never forwards eye_area_percent (radnerf_torso.py:86-106); the head class does (renderer.py:308)
    assert RADNeRF.forwards_eye_area and not RADNeRFTorso.forwards_eye_area


def test_round_schedule_lemma(oracle_ops):
    """with perturb=False the multi-round image equals ONE round of n_step = B (the cap the schedule
    produced), and B can be replayed from the histogram of death indices -- the property the fused kernel relies on."""
    from oracle.render import OracleModel
    for ds, ms in ((1.0, 8), (32.0, 16)):
        sc = scn.Scene(H=40, W=40, T=2, torso=False, max_steps=ms, density_scale=ds)
        fi = sc.frame_inputs(1)
        orc = OracleModel(sc.state, sc.hparams); orc.density_scale = ds
        ref = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"], T_thresh=0.01, **sc.hparams)
        B = ref["stats"]["B_total"]
        ro, rd = fi["rays_o"].view(-1, 3), fi["rays_d"].view(-1, 3)
        N = ro.shape[0]
        nears, fars = oracle_ops.near_far_from_aabb(ro, rd, sc.state["aabb_infer"], 0.05)
        alive = torch.arange(N, dtype=torch.int32)
        rays_t = nears.clone()
        xyzs, dirs, deltas = oracle_ops.march_rays(N, B, alive, rays_t, ro, rd, 1.0, sc.state["density_bitfield"], 1, 128, nears, fars, 128, False, sc.hparams["dt_gamma"], ms)
        sig, rgb, _ = orc.forward(xyzs, dirs, orc.cal_cond_feat(fi["cond"]), sc.state["individual_embeddings"][0])
        ws = torch.zeros(N); dp = torch.zeros(N); img = torch.zeros(N, 3)
        oracle_ops.composite_rays(N, B, alive, rays_t, sig * ds, rgb, deltas, ws, dp, img, 0.01)
        one = (img + (1 - ws).unsqueeze(-1) * fi["bg_color"].view(-1, 3)).clamp(0, 1)
        assert (one - ref["rgb_map"].view(-1, 3)).abs().max().item() < 1e-6
        # replay the schedule from death indices D (1-based position of the break)
        dl = deltas[: N * B].view(N, B, 2)
        # D from a per-ray sequential pass
        D = torch.full((N,), B + 1, dtype=torch.long)
        wsum = torch.zeros(N)
        s2 = (sig * ds)[: N * B].view(N, B)
        done = torch.zeros(N, dtype=torch.bool)
        for k in range(B):
            ran_out = (dl[:, k, 0] == 0) & ~done
            D[ran_out] = k + 1; done |= ran_out
            T = 1 - wsum
            wsum = torch.where(done, wsum, wsum + (1 - torch.exp(-s2[:, k] * dl[:, k, 0])) * T)
            br = (T < 0.01) & ~done
            D[br] = k + 1; done |= br
        cum, n_alive, sched = 0, N, []
        while cum < ms and n_alive > 0:
            n_step = max(min(N // n_alive, 8), 1)
            sched.append((n_alive, n_step)); cum += n_step
            n_alive = int((D > cum).sum())
        assert sched == ref["stats"]["schedule"], (sched, ref["stats"]["schedule"])


def test_sr_head_matches_the_reference_golden():
    """the 256 -> 512 super-resolution head (genefaceplusplus_b200/superres.py) against the REFERENCE's
    own Superresolution outputs (oracle/make_sr_golden.py): same state_dict keys/shapes, same numbers."""
    import json
    import os
    from genefaceplusplus_b200.superres import Superresolution
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sr_head.npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    net = Superresolution(channels=3).eval()
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    ref = {k: tuple(v) for k, v in meta["shapes"].items()}
    assert mine == ref                                                      # drop-in contract: load_state_dict(strict=True) works
    assert (net.resample_filter - torch.from_numpy(g["resample_filter"])).abs().max().item() == 0
    state = scn.synthetic_sr_state(ref, seed=3)
    missing, unexpected = net.load_state_dict(state, strict=False)
    assert not unexpected and all(k.endswith("resample_filter") for k in missing)
    c0, c1, c2, c3 = meta["crop"]
    small = scn.hashed_uniform(3 * 64 * 64, 77, 1.0).reshape(1, 3, 64, 64) + 0.5
    full = scn.hashed_uniform(3 * 256 * 256, 78, 1.0).reshape(1, 3, 256, 256) + 0.5
    with torch.no_grad():
        for name, x in (("in64", small), ("in256", full)):
            y = net(x, noise_mode="const")
            assert y.shape == (1, 3, 512, 512)
            crop = torch.from_numpy(g[f"{name}_crop"])
            assert (y[0, :, c0:c1, c2:c3] - crop).abs().max().item() < 2e-5 * max(1.0, crop.abs().max().item()), name
            assert np.allclose(y.double().sum(dim=(0, 2, 3)).numpy(), g[f"{name}_sum"], rtol=1e-6, atol=1e-2), name
            assert np.allclose(y.double().abs().sum(dim=(0, 2, 3)).numpy(), g[f"{name}_abssum"], rtol=1e-6, atol=1e-2), name
        # a batch shares the modulated kernels: identical to frame-by-frame
        yb = net(torch.cat([full, full.flip(-1)], 0), noise_mode="const")
        assert (yb[0] - net(full, noise_mode="const")[0]).abs().max().item() < 1e-5
        # the noise term is live in the golden state, and 'random' differs from 'const' through it only
        assert (net(full, noise_mode="none") - net(full, noise_mode="const")).abs().max().item() > 1e-4


def test_audio_window_prenet_matches_the_reference_golden():
    """cond_win_size != 1 (audio-window conditioning: deepspeech 16x29, esperanto): `_AudioNet` against the REFERENCE's own AudioNet
    outputs for every window size it supports (oracle/make_condwin_golden.py), same parameter names, and the same ValueError for the
    sizes it rejects (its `win_size == [5, 8]` branch can never be taken)."""
    import json
    import os
    from genefaceplusplus_b200.renderer import _AudioNet
    from oracle.make_condwin_golden import condwin_state
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cond_win.npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    for win, din in meta["cases"]:
        net = _AudioNet(din, 64, win_size=win).eval()
        net.load_state_dict(condwin_state(net.state_dict(), win), strict=True)
        with torch.no_grad():
            y = net(torch.from_numpy(g[f"w{win}_x"]))
        assert (y - torch.from_numpy(g[f"w{win}_y"])).abs().max().item() < 2e-6, win
    for win in meta["unsupported"]:
        with pytest.raises(ValueError):
            _AudioNet(29, 64, win_size=win)
    # the model accepts the audio-window configs and batches them over a clip exactly like frame-by-frame cal_cond_feat
    hp = may_hparams(cond_type="deepspeech", cond_win_size=16, smo_win_size=8)
    m = RADNeRF(hp).eval()
    assert m.cond_prenet.encoder_conv[0].stride == (2,) and m.cond_in_dim == 29
    seq = scn.hashed_uniform(6 * 16 * 29, 31, 2.0).reshape(6, 16, 29)
    with torch.no_grad():
        clip = m.cal_cond_feat_clip(seq)
        for t in (0, 3, 5):
            win = torch.zeros(8, 16, 29)
            for j in range(8):
                i = t - 4 + j
                if 0 <= i < 6:
                    win[j] = seq[i]
            assert (clip[t] - m.cal_cond_feat(win).reshape(-1)).abs().max().item() < 1e-5


def test_sr_folded_weights_reproduce_the_fp32_head():
    """Host half of the native SR path (superres.folded_weights: modulation + demodulation folded, transposed stride-2
    convolution and its FIR merged into four 3x3 phase kernels) pushed through the CPU emulation of the kernels' data flow
    (oracle/sr_emulate.py) must give the fp32 convolutions' image; with the kernels' fp16 operand rounding it must stay
    inside the 1e-3 bar on the clamped image."""
    import json
    import os
    from genefaceplusplus_b200.superres import Superresolution
    from oracle.sr_emulate import emulate
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sr_head.npz"))
    meta = json.loads(bytes(g["meta"]).decode())
    net = Superresolution(channels=3).eval()
    net.load_state_dict(scn.synthetic_sr_state({k: tuple(v) for k, v in meta["shapes"].items()}, seed=3), strict=False)
    fw = net.folded_weights()
    assert fw["up_w"].shape == (256, 1152) and fw["conv0_w"].shape == (128, 1152) and fw["conv1_w"].shape == (64, 576)
    full = scn.hashed_uniform(3 * 256 * 256, 78, 1.0).reshape(1, 3, 256, 256) + 0.5
    flat = full.permute(0, 2, 3, 1).reshape(1, -1, 3).contiguous()
    with torch.no_grad():
        nz = [(l.noise_const * l.noise_strength).float() for l in net._layers()]
        ref = net(full, noise_mode="const")
        e32 = (emulate(fw, flat, 256, nz, fp16=False) - ref).abs().max().item()
        e16 = (emulate(fw, flat, 256, nz, fp16=True).clamp(0, 1) - ref.clamp(0, 1)).abs().max().item()
    assert e32 < 1e-5, e32
    assert e16 < 1e-3, e16
    c0, c1, c2, c3 = meta["crop"]
    crop = torch.from_numpy(g["in256_crop"])
    assert (emulate(fw, flat, 256, nz, fp16=False)[0, :, c0:c1, c2:c3] - crop).abs().max().item() < 2e-5 * max(1.0, crop.abs().max().item())


def test_sr_model_wraps_the_head_render(monkeypatch):
    """RADNeRFwithSR (radnerf_sr.py:203-210): extra state names, result keys and shapes -- with the NeRF render stubbed out
    (the real one needs the GPU and is covered by the head tests)."""
    from genefaceplusplus_b200.renderer import RADNeRFwithSR
    hp = may_hparams(add_eye_blink_cond=True, eye_blink_dim=2, with_sr=True)
    m = RADNeRFwithSR(hp).eval()
    keys = set(m.state_dict())
    assert "lambda_ambient" in keys and "sr_net.block1.torgb.affine.weight" in keys and "blink_encoder.1.bias" in keys
    fake = torch.rand(1, 256 * 256, 3)
    monkeypatch.setattr(RADNeRF, "render", lambda self, *a, **k: {"rgb_map": fake.clone(), "depth_map": torch.zeros(1, 256 * 256)})
    out = m.render(None, None, None, None, None, sr_noise_mode="none")
    assert out["rgb_map"].shape == (1, 3, 256, 256) and out["sr_rgb_map"].shape == (1, 3, 512, 512)
    assert (out["rgb_map"][0].permute(1, 2, 0).reshape(-1, 3) - fake[0]).abs().max().item() == 0
    assert out["sr_rgb_map"].min().item() >= 0 and out["sr_rgb_map"].max().item() <= 1
    with pytest.raises(ValueError):
        m.render_clip(torch.eye(4)[None], may_intrinsics(512, 512), 512, 512)


def test_torso_sr_model_host_path_matches_the_reference_golden(oracle_ops, monkeypatch):
    """RADNeRFTorsowithSR (radnerf_torso_sr.py): torso-SR field, three-way composite and SR head of the PRODUCT class against
    the reference's own render() outputs (tests/golden/torso_sr256.npz).  The two GPU-only pieces are stood in for by the
    checker: the head NeRF (libgfpp's fused kernels, covered by the head tests) and the per-op encoder kernels."""
    import json
    import os
    from genefaceplusplus_b200.renderer import RADNeRFTorsowithSR
    from oracle.render import OracleModel
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "torso_sr256.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    hp = may_hparams(**meta["overrides"])
    state = scn.make_torso_sr_state(hp)
    m = RADNeRFTorsowithSR(hp).eval()
    m.load_state_dict(state, strict=True)                       # the reference class loaded this very state with strict=True
    m.density_scale = meta["density_scale"]
    sc = scn.Scene(H=256, W=256, T=8, torso=True, density_scale=meta["density_scale"])
    t = meta["frame"]
    fi = sc.frame_inputs(t)
    cond = scn.cond_window(sc.cond, t, 3)
    lm68 = scn.lm68_sequence(8)[t].reshape(1, 136)
    eye = torch.tensor([[meta["eye"]]])
    head = OracleModel(state, {**hp, "with_sr": False}, torso=False)
    head.density_scale = meta["density_scale"]

    def fake_head(self, rays_o, rays_d, cond_feat, dt_gamma, max_steps, T_thresh):
        assert (cond_feat.reshape(-1) - head.cal_cond_feat(cond, eye).reshape(-1)).abs().max().item() < 1e-6     # product cond nets, blink incl.
        r = head.render(rays_o.view(1, -1, 3), rays_d.view(1, -1, 3), cond, fi["bg_coords"], fi["poses"], bg_color=torch.zeros(256 * 256, 3),
                        T_thresh=T_thresh, eye_area_percent=eye, **{**hp, "with_sr": False, "max_steps": max_steps, "dt_gamma": dt_gamma})
        return r["rgb_map"].view(-1, 3), r["weights_sum"].view(-1), r["depth_map"].view(-1)

    monkeypatch.setattr(RADNeRFTorsowithSR, "_head", fake_head)
    m.encoders = oracle_ops
    m.torso_backend = "torch"        # the host-side field (A/B reference of libgfpp's k_torso_sr, which needs the GPU)
    kw = {k: v for k, v in hp.items() if k not in ("max_steps", "dt_gamma", "bg_color")}
    out = m.render(fi["rays_o"], fi["rays_d"], cond, fi["bg_coords"], fi["poses"], index=t, dt_gamma=hp["dt_gamma"], bg_color=fi["bg_color"],
                   max_steps=16, T_thresh=sc.T_thresh, upscale_torso=True, lm68=lm68, eye_area_percent=eye, staged=False, **kw)
    assert out["rgb_map"].shape == (1, 3, 256, 256) and out["sr_rgb_map"].shape == (1, 3, 512, 512) and out["sr_torso_rgb_map"].shape == (1, 3, 512, 512)
    for k, (a, b, c, d) in meta["crops"].items():
        assert (out[k][0, :, a:b, c:d] - torch.from_numpy(z[f"{k}_crop"])).abs().max().item() < 5e-6, k
    assert abs(out["torso_alpha_map"].double().sum().item() - float(z["torso_alpha_sum"][0])) < 5e-2
    assert abs(out["deform"].double().abs().sum().item() - float(z["deform_abssum"][0])) < 5e-2
    with pytest.raises(ValueError):
        m.render(fi["rays_o"], fi["rays_d"], cond, fi["bg_coords"], fi["poses"], lm68=None)


def test_driver_shim_conditioning_and_plumbing():
    """genefaceplusplus_b200/driver.py: batched conditioning == per-frame cal_cond_feat on the driver's own
    window lists; the render loop hands libgfpp the right chunks (render_frames stubbed: it needs the GPU)."""
    from genefaceplusplus_b200 import driver
    hp = may_hparams(add_eye_blink_cond=True, eye_blink_dim=2)
    sc = scn.Scene(H=8, W=8, T=7, torso=True)
    m = RADNeRFTorso(hp).eval()
    m.load_state_dict(sc.state, strict=False)
    head = RADNeRF(hp).eval()
    head.load_state_dict({k: v for k, v in sc.state.items() if not k.startswith(("torso_", "density_grid_torso"))}, strict=False)
    wins = [scn.cond_window(sc.cond, t) for t in range(7)]
    eye = [torch.tensor([[0.1 * t]]) for t in range(7)]
    with torch.no_grad():
        f_head = driver.cond_feat_from_windows(head, wins, eye)
        f_torso = driver.cond_feat_from_windows(m, wins, eye)
        for t in range(7):
            assert (f_head[t] - head.cal_cond_feat(wins[t], eye_area_percent=eye[t])).abs().max().item() < 1e-6
            assert (f_torso[t] - m.cal_cond_feat(wins[t])).abs().max().item() < 1e-6       # radnerf_torso.py:106: no eye input
    calls = []

    def fake_render_frames(cond_feat, **kw):
        calls.append((cond_feat.shape[0], tuple(kw["rays_o"].shape), tuple(kw["pose6"].shape)))
        return {"rgb_map": torch.full((cond_feat.shape[0], 64, 3), 0.5)}

    m.render_frames = fake_render_frames
    fis = [sc.frame_inputs(t) for t in range(7)]
    batch = {"rays_o": [f["rays_o"] for f in fis], "rays_d": [f["rays_d"] for f in fis], "cond_wins": wins, "poses": [f["poses"] for f in fis],
             "bg_coords": sc.bg_coords, "bg_img": sc.bg_color, "eye_area_percent": eye}
    out = driver.render_driver_batch(m, batch, frames_per_call=3)
    assert out.shape == (7, 3, 8, 8) and calls == [(3, (3, 64, 3), (3, 6)), (3, (3, 64, 3), (3, 6)), (1, (1, 64, 3), (1, 6))]
    u8 = driver.video_uint8(out)
    assert u8.dtype == torch.uint8 and u8.shape == (7, 8, 8, 3) and int(u8[0, 0, 0, 0]) == 127


def test_sr_model_clip_api(monkeypatch):
    """RADNeRFwithSR.render_clip: NeRF clip at 256x256 (stubbed: GPU) -> SR head in chunks -> [T,3,512,512] in [0,1]."""
    from genefaceplusplus_b200.renderer import RADNeRFwithSR
    m = RADNeRFwithSR(may_hparams(with_sr=True)).eval()
    fake = torch.rand(5, 256 * 256, 3)
    monkeypatch.setattr(RADNeRF, "render_clip", lambda self, *a, **k: fake.clone())
    out = m.render_clip(torch.eye(4).repeat(5, 1, 1), may_intrinsics(256, 256), 256, 256, sr_noise_mode="none", sr_frames_per_call=2)
    assert out.shape == (5, 3, 512, 512) and 0 <= out.min().item() and out.max().item() <= 1
    one = m.sr_net(fake[3].view(1, 256, 256, 3).permute(0, 3, 1, 2), noise_mode="none").clamp(0, 1)
    assert (out[3] - one[0]).abs().max().item() < 1e-5          # chunking does not change a frame
