"""Training-side native ops of libgfpp (csrc/train_kernels.cu) on the B200:
  * against the C checker (oracle/native_ops.c, second half) through the wrapper-level API genefaceplusplus_b200/train_ops.py;
  * checker AND libgfpp against the REFERENCE'S OWN training kernels (oracle/_ref, compiled unmodified), ray by ray: the
    reference hands out point offsets in atomicAdd arrival order, so layouts are compared through each implementation's `rays`
    table."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from genefaceplusplus_b200 import scene as scn
from genefaceplusplus_b200.config import GridLayout

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _load_ref(name):
    so = os.path.join(REF, name, name + ".so")
    if not os.path.exists(so):
        pytest.skip(f"{so} not built (oracle/build_ref.py needs /root/reference)")
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rays(oracle_ops, H=64):
    sc = scn.Scene(H=H, W=H, T=2, torso=False)
    fi = sc.frame_inputs(1)
    ro, rd = fi["rays_o"].view(-1, 3).contiguous(), fi["rays_d"].view(-1, 3).contiguous()
    nears, fars = oracle_ops.near_far_from_aabb(ro, rd, sc.state["aabb_infer"], 0.05)
    return sc, ro, rd, nears, fars


def _segments(N=3000, seed=0, max_len=40):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, max_len, (N,), generator=g, dtype=torch.int32)
    lens[::7] = 0
    lens[3] = 70
    offs = torch.cumsum(lens.long(), 0) - lens.long()
    perm = torch.randperm(N, generator=g).int()
    rays = torch.stack([perm, offs.int(), lens], 1).contiguous()
    M = int(lens.sum())
    sig = torch.rand(M, generator=g) * 6
    rgb = torch.rand(M, 3, generator=g)
    amb = torch.rand(M, generator=g)
    dt = torch.rand(M, generator=g) * 0.05 + 0.01
    deltas = torch.stack([dt, torch.rand(M, generator=g) * 3 + 2], 1).contiguous()
    return rays, M, sig, rgb, amb, deltas


@pytest.mark.parametrize("max_steps,dt_gamma,perturb", [(16, 1 / 256, False), (64, 0.0, True)])
def test_march_rays_train_bit_exact_and_deterministic(oracle_ops, max_steps, dt_gamma, perturb):
    from genefaceplusplus_b200 import backend_shims
    rm = backend_shims.make_modules()["_raymarching_face"]
    sc, ro, rd, nears, fars = _rays(oracle_ops)
    bits = sc.state["density_bitfield"]
    N = ro.shape[0]
    noises = torch.rand(N, generator=torch.Generator().manual_seed(4)) if perturb else torch.zeros(N)
    x_ref, d_ref, l_ref, r_ref, c_ref = oracle_ops.march_rays_train(ro, rd, 1.0, bits, 1, 128, nears, fars, noises=noises, dt_gamma=dt_gamma, max_steps=max_steps)
    M = N * max_steps

    def run(M_):
        xyzs, dirs, deltas = torch.zeros(M_, 3, device="cuda"), torch.zeros(M_, 3, device="cuda"), torch.zeros(M_, 2, device="cuda")
        rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
        counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        rm.march_rays_train(ro.cuda(), rd.cuda(), bits.cuda(), 1.0, dt_gamma, max_steps, N, 1, 128, M_, nears.cuda(), fars.cuda(), xyzs, dirs, deltas, rays,
                            counter, noises.cuda())
        torch.cuda.synchronize()
        return xyzs.cpu(), dirs.cpu(), deltas.cpu(), rays.cpu(), counter.cpu()

    xyzs, dirs, deltas, rays, counter = run(M)
    assert torch.equal(rays, r_ref) and torch.equal(counter, c_ref)                 # counts, ray-order offsets, counter
    m = int(counter[0])
    assert torch.equal(xyzs[:m], x_ref[:m]) and torch.equal(deltas[:m], l_ref[:m]) and torch.equal(dirs[:m], d_ref[:m])
    assert xyzs[m:].abs().sum().item() == 0
    again = run(M)
    assert all(torch.equal(a, b) for a, b in zip(again, (xyzs, dirs, deltas, rays, counter))), "march_rays_train must be deterministic"
    xs, _, ls, rs, _ = run(m // 2)                                                   # overflow: rows kept, samples dropped
    x2, _, l2, r2, _ = oracle_ops.march_rays_train(ro, rd, 1.0, bits, 1, 128, nears, fars, M=m // 2, noises=noises, dt_gamma=dt_gamma, max_steps=max_steps)
    assert torch.equal(rs, r2) and torch.equal(xs, x2) and torch.equal(ls, l2)


def _per_ray_err(got, ref, rays):
    """max |got - ref| per ray over its samples (sample-indexed tensors) -> [N]"""
    e = (got - ref).abs()
    if e.dim() > 1:
        e = e.max(-1).values
    out = torch.zeros(rays.shape[0])
    for n in range(rays.shape[0]):
        o, k = int(rays[n, 1]), int(rays[n, 2])
        if k:
            out[n] = e[o:o + k].max()
    return out


@pytest.mark.parametrize("T_thresh", [1e-4, 0.2])
def test_composite_rays_train_forward_backward_vs_checker(oracle_ops, T_thresh):
    from genefaceplusplus_b200 import train_ops
    rays, M, sig, rgb, amb, deltas = _segments()
    ws, asum, depth, image = oracle_ops.composite_rays_train_forward(sig, rgb, amb, deltas, rays, T_thresh)
    g = torch.Generator().manual_seed(9)
    N = rays.shape[0]
    gws, gas, gim = torch.randn(N, generator=g), torch.randn(N, generator=g), torch.randn(N, 3, generator=g)
    gs, gr, ga = oracle_ops.composite_rays_train_backward(gws, gas, gim, sig, rgb, amb, deltas, rays, ws, asum, image, T_thresh)
    s_, r_, a_ = sig.cuda().requires_grad_(), rgb.cuda().requires_grad_(), amb.cuda().requires_grad_()
    w2, a2, d2, i2 = train_ops.composite_rays_train(s_, r_, a_, deltas.cuda(), rays.cuda(), T_thresh)
    ((w2 * gws.cuda()).sum() + (a2 * gas.cuda()).sum() + (i2 * gim.cuda()).sum() + 0.0 * d2.sum()).backward()
    torch.cuda.synchronize()
    # a ray whose transmittance passes within rounding of T_thresh may be cut one sample earlier / later (product association
    # of the warp scan vs the sequential loop): such knife-edge rays are counted, never masked silently
    idx = rays[:, 0].long()
    fwd = torch.stack([(w2.cpu() - ws).abs()[idx], (a2.cpu() - asum).abs()[idx] * 0.1, (d2.cpu() - depth).abs()[idx], (i2.cpu() - image).abs().max(-1).values[idx]]).max(0).values
    bwd = torch.stack([_per_ray_err(s_.grad.cpu(), gs, rays) / max(1.0, gs.abs().max().item()), _per_ray_err(r_.grad.cpu(), gr, rays),
                       _per_ray_err(a_.grad.cpu(), ga, rays)]).max(0).values
    bad = ((fwd > 2e-5) | (bwd > 2e-5)).nonzero().view(-1)
    print(f"T_thresh={T_thresh}: forward max {fwd.max().item():.2e}, backward max {bwd.max().item():.2e}, rays over 2e-5: {bad.numel()} of {N}")
    assert bad.numel() <= 2, bad.tolist()


def test_march_rays_train_autograd_backward(oracle_ops):
    from genefaceplusplus_b200 import train_ops
    sc, ro, rd, nears, fars = _rays(oracle_ops, H=32)
    bits = sc.state["density_bitfield"]
    ro_, rd_ = ro.cuda().requires_grad_(), rd.cuda().requires_grad_()
    xyzs, dirs, deltas, rays = train_ops.march_rays_train(ro_, rd_, 1.0, bits.cuda(), 1, 128, nears.cuda(), fars.cuda(), None, -1, False, 128, True, 1 / 256, 16)
    g = torch.Generator().manual_seed(2)
    gx, gd = torch.randn(xyzs.shape[0], 3, generator=g), torch.randn(xyzs.shape[0], 3, generator=g)
    ((xyzs * gx.cuda()).sum() + (dirs * gd.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert xyzs.shape[0] % 128 == 0
    go, gdd = oracle_ops.march_rays_train_backward(gx, gd, rays.cpu(), deltas.detach().cpu().contiguous())
    assert (ro_.grad.cpu() - go).abs().max().item() < 1e-4 and (rd_.grad.cpu() - gdd).abs().max().item() < 1e-3


@pytest.mark.parametrize("D,gridtype,interp", [(3, 1, 0), (2, 1, 0), (3, 0, 1)])
def test_grid_encode_autograd_and_tv_vs_checker(oracle_ops, D, gridtype, interp):
    from genefaceplusplus_b200 import train_ops
    lay = GridLayout(D, log2_hashmap_size=14 if gridtype == 0 else 16, desired_resolution=512, gridtype="hash" if gridtype == 0 else "tiled")
    offsets = torch.from_numpy(np.asarray(lay.offsets, dtype=np.int32))
    g = torch.Generator().manual_seed(D * 10 + gridtype)
    table = torch.rand(int(offsets[-1]), 2, generator=g) - 0.5
    B = 5000
    x = torch.rand(B, D, generator=g) * 0.98 + 0.01
    x[5] = 1.5
    G = torch.randn(B, 32, generator=g)
    y_ref = oracle_ops.grid_encode(x, table, offsets, lay.per_level_scale, 16, gridtype, False, interp)
    dydx_ref = oracle_ops.grid_encode_dydx(x, table, offsets, lay.per_level_scale, 16, gridtype, False, interp)
    ge_ref, gi_ref = oracle_ops.grid_encode_backward(G, x, table, offsets, lay.per_level_scale, 16, gridtype, False, interp, dy_dx=dydx_ref)
    x_, t_ = x.cuda().requires_grad_(), table.cuda().requires_grad_()
    y = train_ops.grid_encode(x_, t_, offsets.cuda(), lay.per_level_scale, 16, True, gridtype, False, interp)
    (y * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    e_y, e_t, e_x = (y.detach().cpu() - y_ref).abs().max().item(), (t_.grad.cpu() - ge_ref).abs().max().item(), (x_.grad.cpu() - gi_ref).abs().max().item()
    print(f"D={D} gridtype={gridtype} interp={interp}: forward {e_y:.2e}, table grad {e_t:.2e} (max {ge_ref.abs().max().item():.2e}), input grad {e_x:.2e} (max {gi_ref.abs().max().item():.2e})")
    assert e_y < 2e-6 and e_t < 2e-5 * max(1.0, ge_ref.abs().max().item()) and e_x < 1e-4 * max(1.0, gi_ref.abs().max().item())
    xb = x * 2 - 1                                     # the wrapper maps [-bound, bound] back to [0,1] (grid.py:180): feed both the same numbers
    tv_ref = oracle_ops.grad_total_variation((xb + 1) / 2, table, offsets, 0.5, lay.per_level_scale, 16, gridtype, False)
    t2 = table.cuda().requires_grad_()
    t2.grad = torch.zeros_like(t2)
    train_ops.grad_total_variation(t2, offsets.cuda(), lay.per_level_scale, 16, D, weight=0.5, inputs=xb.cuda(), bound=1, gridtype=gridtype)
    torch.cuda.synchronize()
    e_tv = (t2.grad.cpu() - tv_ref).abs().max().item()
    print(f"TV grad {e_tv:.2e} (max {tv_ref.abs().max().item():.2e})")
    assert e_tv < 1e-4 * max(1.0, tv_ref.abs().max().item())


def test_update_extra_state_helpers_vs_checker(oracle_ops):
    from genefaceplusplus_b200 import train_ops
    g = torch.Generator().manual_seed(3)
    coords = torch.randint(0, 128, (5000, 3), generator=g, dtype=torch.int32)
    idx = train_ops.morton3D(coords.cuda())
    assert torch.equal(idx.cpu(), oracle_ops.morton3D(coords)) and torch.equal(train_ops.morton3D_invert(idx).cpu(), coords)
    grid = torch.rand(2, 32 ** 3, generator=g)
    assert torch.equal(train_ops.morton3D_dilation(grid.cuda()).cpu(), oracle_ops.morton3D_dilation(grid))
    assert torch.equal(train_ops.packbits(grid.cuda(), 0.5).cpu(), oracle_ops.packbits(grid, 0.5))
    ro = torch.randn(4000, 3, generator=g) * 0.3
    rd = torch.nn.functional.normalize(torch.randn(4000, 3, generator=g), dim=-1)
    assert (train_ops.sph_from_ray(ro.cuda(), rd.cuda(), 2.0).cpu() - oracle_ops.sph_from_ray(ro, rd, 2.0)).abs().max().item() < 2e-6


# ------------------------------------------------------------------------------------------------ pin against the reference's kernels
def test_training_ops_vs_the_reference_kernels(oracle_ops):
    """The reference's own march_rays_train / composite_rays_train / grid_encode_backward / grad_total_variation kernels
    (unmodified, oracle/_ref) against the checker and libgfpp.  FMA contraction in the reference build may flip an occupancy
    decision for a handful of rays: those are counted and bounded."""
    from genefaceplusplus_b200 import backend_shims
    ref_rm, ref_ge = _load_ref("_raymarching_face"), _load_ref("_gridencoder")
    ours = backend_shims.make_modules()
    sc, ro, rd, nears, fars = _rays(oracle_ops)
    bits = sc.state["density_bitfield"]
    N, max_steps = ro.shape[0], 16
    M = N * max_steps
    x_o, d_o, l_o, r_o, c_o = oracle_ops.march_rays_train(ro, rd, 1.0, bits, 1, 128, nears, fars, dt_gamma=1 / 256, max_steps=max_steps)
    xyzs, dirs, deltas = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ref_rm.march_rays_train(ro.cuda(), rd.cuda(), bits.cuda(), 1.0, 1 / 256, max_steps, N, 1, 128, M, nears.cuda(), fars.cuda(), xyzs, dirs, deltas, rays, counter,
                            torch.zeros(N, device="cuda"))
    torch.cuda.synchronize()
    rays, xyzs, deltas = rays.cpu(), xyzs.cpu(), deltas.cpu()
    assert counter.cpu().tolist()[1] == N
    order = torch.argsort(rays[:, 0].long())
    rr = rays[order]                                                                  # reference rows by ray id
    same_count = rr[:, 2] == r_o[:, 2]
    flips = int((~same_count).sum())
    worst = 0.0
    for n in torch.nonzero(same_count & (r_o[:, 2] > 0)).view(-1).tolist():
        a, b, k = int(rr[n, 1]), int(r_o[n, 1]), int(r_o[n, 2])
        worst = max(worst, (xyzs[a:a + k] - x_o[b:b + k]).abs().max().item(), (deltas[a:a + k] - l_o[b:b + k]).abs().max().item())
    print(f"march_rays_train: reference kernel vs checker: {flips} of {N} rays with a different sample count, max |d| on the rest {worst:.2e}")
    assert flips <= max(2, N // 2000) and worst <= 2e-6
    assert abs(int(counter.cpu()[0]) - int(c_o[0])) <= 16 * max(1, flips)
    # compositing: same inputs in the checker's layout through the reference kernels
    rays_s, Ms, sig, rgb, amb, dl = _segments()
    Ns = rays_s.shape[0]
    ws, asum, depth, image = [torch.empty(Ns, device="cuda") for _ in range(3)] + [torch.empty(Ns, 3, device="cuda")]
    ref_rm.composite_rays_train_forward(sig.cuda(), rgb.cuda(), amb.cuda(), dl.cuda(), rays_s.cuda(), Ms, Ns, 1e-4, ws, asum, depth, image)
    w_o, a_o, d_o2, i_o = oracle_ops.composite_rays_train_forward(sig, rgb, amb, dl, rays_s, 1e-4)
    e_f = max((ws.cpu() - w_o).abs().max().item(), (depth.cpu() - d_o2).abs().max().item(), (image.cpu() - i_o).abs().max().item())
    g = torch.Generator().manual_seed(9)
    gws, gas, gim = torch.randn(Ns, generator=g), torch.randn(Ns, generator=g), torch.randn(Ns, 3, generator=g)
    gs, gr, ga = torch.zeros(Ms, device="cuda"), torch.zeros(Ms, 3, device="cuda"), torch.zeros(Ms, device="cuda")
    ref_rm.composite_rays_train_backward(gws.cuda(), gas.cuda(), gim.cuda(), sig.cuda(), rgb.cuda(), amb.cuda(), dl.cuda(), rays_s.cuda(), ws, asum, image, Ms, Ns,
                                         1e-4, gs, gr, ga)
    gs_o, gr_o, ga_o = oracle_ops.composite_rays_train_backward(gws, gas, gim, sig, rgb, amb, dl, rays_s, w_o, a_o, i_o, 1e-4)
    e_b = max((gs.cpu() - gs_o).abs().max().item() / max(1.0, gs_o.abs().max().item()), (gr.cpu() - gr_o).abs().max().item(), (ga.cpu() - ga_o).abs().max().item())
    print(f"composite_rays_train: reference kernels vs checker: forward {e_f:.2e}, backward {e_b:.2e}")
    assert e_f <= 2e-5 and e_b <= 2e-5
    # grid backward + TV: the reference's kernels vs the checker vs libgfpp
    lay = GridLayout(3, log2_hashmap_size=16, desired_resolution=2048, gridtype="tiled")
    offsets = torch.from_numpy(np.asarray(lay.offsets, dtype=np.int32))
    table = torch.rand(int(offsets[-1]), 2, generator=g) - 0.5
    B = 4096
    x = torch.rand(B, 3, generator=g)
    G = torch.randn(B, 32, generator=g)
    S = float(np.log2(lay.per_level_scale))
    grad = G.view(B, 16, 2).permute(1, 0, 2).contiguous().cuda()
    out_r, dy_r = torch.empty(16, B, 2, device="cuda"), torch.empty(B, 16 * 3 * 2, device="cuda")
    ref_ge.grid_encode_forward(x.cuda(), table.cuda(), offsets.cuda(), out_r, B, 3, 2, 16, S, 16, dy_r, 1, False, 0)
    ge_r, gi_r = torch.zeros_like(table).cuda(), torch.zeros(B, 3, device="cuda")
    ref_ge.grid_encode_backward(grad, x.cuda(), table.cuda(), offsets.cuda(), ge_r, B, 3, 2, 16, S, 16, dy_r, gi_r, 1, False, 0)
    out_g, dy_g = torch.empty(16, B, 2, device="cuda"), torch.empty(B, 16 * 3 * 2, device="cuda")
    ours["_gridencoder"].grid_encode_forward(x.cuda(), table.cuda(), offsets.cuda(), out_g, B, 3, 2, 16, S, 16, dy_g, 1, False, 0)
    ge_g, gi_g = torch.zeros_like(table).cuda(), torch.zeros(B, 3, device="cuda")
    ours["_gridencoder"].grid_encode_backward(grad, x.cuda(), table.cuda(), offsets.cuda(), ge_g, B, 3, 2, 16, S, 16, dy_g, gi_g, 1, False, 0)
    dy_o = oracle_ops.grid_encode_dydx(x, table, offsets, lay.per_level_scale, 16, 1, False, 0)
    ge_o, gi_o = oracle_ops.grid_encode_backward(G, x, table, offsets, lay.per_level_scale, 16, 1, False, 0, dy_dx=dy_o)
    torch.cuda.synchronize()
    # The reference derives every level scale with the DEVICE exp2f (gridencoder.cu:137, <= 2 ulp), the checker and libgfpp with
    # the host libm (tests/test_gpu_ref_pin.py): a sample within ~1e-4 cells of a cell face then sits in the neighbouring cell.
    # Table gradients are continuous across that face; dy_dx (a per-cell slope), the input gradient built from it and the TV term
    # (added to the cell's own entry) are not -- for those the disagreeing entries are COUNTED and bounded, the rest held to 1e-4.
    tv_r, tv_g = torch.zeros_like(table).cuda(), torch.zeros_like(table).cuda()
    ref_ge.grad_total_variation(x.cuda(), table.cuda(), tv_r, offsets.cuda(), 0.5, B, 3, 2, 16, S, 16, 1, False)
    ours["_gridencoder"].grad_total_variation(x.cuda(), table.cuda(), tv_g, offsets.cuda(), 0.5, B, 3, 2, 16, S, 16, 1, False)
    tv_o = oracle_ops.grad_total_variation(x, table, offsets, 0.5, lay.per_level_scale, 16, 1, False)
    torch.cuda.synchronize()

    def cmp(name, a, b, scale, frac_allowed):
        e = (a.cpu().reshape(-1) - b.cpu().reshape(-1)).abs() / scale
        frac = (e > 1e-4).float().mean().item()
        print(f"  {name}: max {e.max().item():.2e}, median {e.median().item():.2e}, entries over 1e-4: {frac:.2e} (allowed {frac_allowed:.0e})")
        assert frac <= frac_allowed and e.median().item() <= 1e-5, name

    print("grid encoder (errors relative to the largest entry):")
    s_dy, s_t, s_x, s_tv = max(1.0, dy_o.abs().max().item()), max(1.0, ge_o.abs().max().item()), max(1.0, gi_o.abs().max().item()), max(1e-3, tv_o.abs().max().item())
    cmp("dy_dx        reference vs checker", dy_r.view(B, 16, 3, 2), dy_o, s_dy, 2e-3)
    cmp("dy_dx        libgfpp vs checker  ", dy_g.view(B, 16, 3, 2), dy_o, s_dy, 0.0)
    cmp("table grad   reference vs checker", ge_r, ge_o, s_t, 5e-4)
    cmp("table grad   libgfpp vs checker  ", ge_g, ge_o, s_t, 0.0)
    cmp("input grad   reference vs checker", gi_r, gi_o, s_x, 2e-2)
    cmp("input grad   libgfpp vs checker  ", gi_g, gi_o, s_x, 0.0)
    cmp("TV grad      reference vs checker", tv_r, tv_o, s_tv, 1e-3)
    cmp("TV grad      libgfpp vs checker  ", tv_g, tv_o, s_tv, 0.0)
