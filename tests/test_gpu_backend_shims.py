"""Level-B boundary: the modules registered under the reference's extension names take the pybind argument lists."""
import importlib

import numpy as np
import pytest
import torch

from genefaceplusplus_b200 import backend_shims, scene as scn
from genefaceplusplus_b200.config import GridLayout

pytestmark = pytest.mark.gpu


def test_shims_register_and_run_with_pybind_signatures(oracle_ops):
    backend_shims.install()
    rm = importlib.import_module("_raymarching_face")
    ge = importlib.import_module("_gridencoder")
    sh = importlib.import_module("_shencoder")
    fr = importlib.import_module("_freqencoder")
    sc = scn.Scene(H=48, W=48, T=2, torso=False)
    fi = sc.frame_inputs(0)
    ro, rd = fi["rays_o"].view(-1, 3).cuda().contiguous(), fi["rays_d"].view(-1, 3).cuda().contiguous()
    N = ro.shape[0]
    aabb = sc.state["aabb_infer"].cuda()
    nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
    rm.near_far_from_aabb(ro, rd, aabb, N, 0.05, nears, fars)                     # raymarching.py:45
    n_ref, f_ref = oracle_ops.near_far_from_aabb(ro.cpu(), rd.cpu(), aabb.cpu(), 0.05)
    assert torch.equal(nears.cpu(), n_ref)
    # one host-loop round exactly like renderer.py:366-380, on the shims
    n_step = 2
    alive = torch.arange(N, dtype=torch.int32, device="cuda")
    rays_t = nears.clone()
    M = N * n_step + 128 - (N * n_step) % 128
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    bits = sc.state["density_bitfield"].cuda()
    rm.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, 1 / 256, 16, 1, 128, bits, nears, fars, xyzs, dirs, deltas, torch.zeros(N, device="cuda"))
    x_ref, _, l_ref = oracle_ops.march_rays(N, n_step, alive.cpu(), n_ref.clone(), ro.cpu(), rd.cpu(), 1.0, bits.cpu(), 1, 128, n_ref, f_ref, 128, False, 1 / 256, 16)
    assert torch.equal(xyzs.cpu(), x_ref) and torch.equal(deltas.cpu(), l_ref)
    lay = GridLayout(3)
    emb = sc.state["position_embedder.embeddings"].cuda()
    x01 = ((xyzs + 1) / 2).contiguous()
    out = torch.empty(16, M, 2, device="cuda")
    ge.grid_encode_forward(x01, emb, sc.state["position_embedder.offsets"].cuda(), out, M, 3, 2, 16, np.log2(lay.per_level_scale), 16, None, 1, False, 0)  # grid.py:54
    ref = oracle_ops.grid_encode(x01.cpu(), emb.cpu(), sc.state["position_embedder.offsets"], lay.per_level_scale, 16, 1, False, 0)
    assert (out.permute(1, 0, 2).reshape(M, 32).cpu() - ref).abs().max().item() < 2e-6
    o16 = torch.empty(M, 16, device="cuda")
    sh.sh_encode_forward(dirs, o16, M, 3, 4, None)                                  # sphere_harmonics.py:32
    assert (o16.cpu() - oracle_ops.sh_encode(dirs.cpu(), 4)).abs().max().item() < 1e-6
    p6 = fi["poses"].cuda().contiguous()
    o54 = torch.empty(1, 54, device="cuda")
    fr.freq_encode_forward(p6, 1, 6, 4, 54, o54)                                    # freq.py:29
    assert (o54.cpu() - oracle_ops.freq_encode(p6.cpu(), 4)).abs().max().item() < 2e-6
    sig = torch.rand(M, device="cuda") * 20; rgb = torch.rand(M, 3, device="cuda")
    ws = torch.zeros(N, device="cuda"); dp = torch.zeros(N, device="cuda"); img = torch.zeros(N, 3, device="cuda")
    rm.composite_rays(N, n_step, 0.01, alive, rays_t, sig, rgb, deltas, ws, dp, img)  # raymarching.py:419
    a2 = torch.arange(N, dtype=torch.int32); t2 = n_ref.clone(); w2 = torch.zeros(N); d2 = torch.zeros(N); i2 = torch.zeros(N, 3)
    oracle_ops.composite_rays(N, n_step, a2, t2, sig.cpu(), rgb.cpu(), l_ref, w2, d2, i2, 0.01)
    assert torch.equal(alive.cpu(), a2) and (img.cpu() - i2).abs().max().item() < 2e-6
    with pytest.raises(NotImplementedError):   # input-gradient op of the SH encoder: still the stock extension's job
        sh.sh_encode_backward()
