"""Deterministic synthetic "May-shaped" scene.

No checkpoint or dataset ships with the reference, so every BASELINE config is rendered from a
synthetic model with the reference's exact state_dict keys/shapes (verified by
`load_state_dict(strict=True)` into the reference's own RADNeRFTorso in tests/golden/make_golden.py).
All randomness comes from numpy's PCG64 with fixed seeds so that the scene is identical in this
container and on the GPU box.

Host-side restatements used to build inputs (plumbing, not the hot path):
  get_rays         modules/radnerfs/utils.py:283-364   (full-image branch, N=-1)
  get_bg_coords    modules/radnerfs/utils.py:274-279
  convert_poses    modules/radnerfs/utils.py:264-270  (+ matrix_to_euler_angles :169-204, 'XYZ')
  get_audio_features(att_mode=2)  modules/radnerfs/utils.py:86-102
"""
import math

import numpy as np
import torch

from .config import GridLayout, cascade_count, may_hparams, may_intrinsics


# ----------------------------------------------------------------------------- model state
def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _linear_w(rng, out_f, in_f):
    # same distribution as nn.Linear's default kaiming_uniform(a=sqrt(5)): U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    return _uniform(rng, (out_f, in_f), 1.0 / math.sqrt(in_f))


def _conv_wb(rng, out_c, in_c, k):
    b = 1.0 / math.sqrt(in_c * k)
    return _uniform(rng, (out_c, in_c, k), b), _uniform(rng, (out_c,), b)


def spread3(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3(x, y, z):
    """10-bit-per-axis Morton code (modules/radnerfs/raymarching/src/raymarching.cu:57-72)."""
    return spread3(x) | (spread3(y) << np.uint32(1)) | (spread3(z) << np.uint32(2))


def ellipsoid_density_grid(grid_size=128, radii=(0.35, 0.30, 0.35)):
    """density_grid [1, H^3] in Morton order: 1 inside the ellipsoid, 0 outside (cascade 0 only)."""
    H = grid_size
    idx = np.arange(H, dtype=np.uint32)
    c = (idx.astype(np.float64) + 0.5) / H * 2 - 1
    X, Y, Z = np.meshgrid(idx, idx, idx, indexing="ij")
    wx, wy, wz = c[X], c[Y], c[Z]
    inside = (wx / radii[0]) ** 2 + (wy / radii[1]) ** 2 + (wz / radii[2]) ** 2 < 1.0
    grid = np.zeros(H ** 3, dtype=np.float32)
    grid[morton3(X.ravel(), Y.ravel(), Z.ravel())] = inside.ravel().astype(np.float32)
    return torch.from_numpy(grid).view(1, -1)


def pack_bitfield(density_grid, thresh=0.5):
    """little-endian 8 cells per byte (raymarching.cu:267-300)."""
    g = (density_grid.view(-1).numpy() > thresh).astype(np.uint8).reshape(-1, 8)
    weights = (1 << np.arange(8)).astype(np.uint8)
    return torch.from_numpy((g * weights).sum(1).astype(np.uint8))


def torso_density_grid(grid_size=128, frac=0.45):
    """density_grid_torso [H*H]: 1 for the bottom `frac` of the IMAGE, else 0.

    grid_sample(grid[1,1,H,W], bg_coords) reads bg_coords[...,0] as x (width index of the grid) and
    bg_coords[...,0] is the image ROW coordinate (utils.py:274-279), so image rows map to the grid's
    last axis: the torso band is `grid[:, w >= (1-frac)*H]` (the transpose noted in radnerf_torso.py:223).
    """
    H = grid_size
    g = np.zeros((H, H), dtype=np.float32)
    g[:, int(round((1.0 - frac) * H)):] = 1.0
    return torch.from_numpy(g.reshape(-1))


def make_state(torso=True, hparams=None, table_amp=0.5, table_decay=0.0, seed=0):
    """Synthetic state_dict with the reference's key names and shapes.

    table_amp / table_decay: level l of every grid is U(-a_l, a_l) with a_l = table_amp * 2^(-table_decay*l*S)
    (decay 0 = SURVEY 8(d): flat 0.5 at every level; decay 1 = amplitude ~ 1/resolution, i.e. a field
    whose gradient is level-independent, closer to a trained model).
    """
    hp = hparams or may_hparams()
    rng = np.random.Generator(np.random.PCG64(seed))
    G = hp["grid_size"]
    st = {}
    st["individual_embeddings"] = torch.from_numpy((rng.standard_normal((hp["individual_embedding_num"], hp["individual_embedding_dim"])) * 0.1).astype(np.float32))
    if torso:
        st["torso_individual_codes"] = torch.from_numpy((rng.standard_normal((hp["individual_embedding_num"], hp["torso_individual_embedding_dim"])) * 0.1).astype(np.float32))
    b = float(hp["bound"])
    aabb = torch.tensor([-b, -b / 2, -b, b, b / 2, b], dtype=torch.float32)
    st["aabb_train"] = aabb.clone()
    st["aabb_infer"] = aabb.clone()
    dg = ellipsoid_density_grid(G)
    assert cascade_count(hp["bound"]) == 1
    st["density_grid"] = dg
    st["density_bitfield"] = pack_bitfield(dg)
    st["step_counter"] = torch.zeros(16, 2, dtype=torch.int32)
    if torso:
        st["density_grid_torso"] = torso_density_grid(G)

    cin = 68 * 3
    for i, (o, c) in zip((0, 2, 4, 6), ((32, cin), (32, 32), (64, 32), (64, 64))):
        w, bb = _conv_wb(rng, o, c, 3)
        st[f"cond_prenet.encoder_conv.{i}.weight"], st[f"cond_prenet.encoder_conv.{i}.bias"] = w, bb
    for i in (0, 2):
        st[f"cond_prenet.encoder_fc1.{i}.weight"] = _linear_w(rng, 64, 64)
        st[f"cond_prenet.encoder_fc1.{i}.bias"] = _uniform(rng, (64,), 1 / 8.0)
    for i, (o, c) in zip((0, 2, 4, 6, 8), ((16, 64), (8, 16), (4, 8), (2, 4), (1, 2))):
        w, bb = _conv_wb(rng, o, c, 3)
        st[f"cond_att_net.attentionConvNet.{i}.weight"], st[f"cond_att_net.attentionConvNet.{i}.bias"] = w, bb
    S = hp["smo_win_size"]
    st["cond_att_net.attentionNet.0.weight"] = _linear_w(rng, S, S)
    st["cond_att_net.attentionNet.0.bias"] = _uniform(rng, (S,), 1 / math.sqrt(S))

    def table(layout, sub_seed):
        r = np.random.Generator(np.random.PCG64(seed * 1000 + sub_seed))
        t = r.uniform(-1.0, 1.0, size=(layout.n_entries, layout.level_dim)).astype(np.float32)
        for l in range(layout.num_levels):
            a = table_amp * 2.0 ** (-table_decay * l * float(layout.S))
            t[layout.offsets[l]:layout.offsets[l + 1]] *= np.float32(a)
        return torch.from_numpy(t)

    gt = {"tiledgrid": "tiled", "hashgrid": "hash"}[hp["grid_type"]]
    pos = GridLayout(3, log2_hashmap_size=hp["log2_hashmap_size"], desired_resolution=hp["desired_resolution"] * hp["bound"], gridtype=gt, interpolation=hp["grid_interpolation_type"])
    amb = GridLayout(hp["ambient_coord_dim"], log2_hashmap_size=hp["log2_hashmap_size"], desired_resolution=hp["desired_resolution"], gridtype=gt, interpolation=hp["grid_interpolation_type"])
    st["position_embedder.embeddings"] = table(pos, 1)
    st["position_embedder.offsets"] = torch.from_numpy(pos.offsets.copy())
    hd = hp["hidden_dim_ambient"]
    st["ambient_net.net.0.weight"] = _linear_w(rng, hd, pos.output_dim + hp["cond_out_dim"])
    st["ambient_net.net.1.weight"] = _linear_w(rng, hd, hd)
    st["ambient_net.net.2.weight"] = _linear_w(rng, hp["ambient_coord_dim"], hd)
    st["ambient_embedder.embeddings"] = table(amb, 2)
    st["ambient_embedder.offsets"] = torch.from_numpy(amb.offsets.copy())
    hs = hp["hidden_dim_sigma"]
    st["sigma_net.net.0.weight"] = _linear_w(rng, hs, pos.output_dim + amb.output_dim)
    st["sigma_net.net.1.weight"] = _linear_w(rng, hs, hs)
    st["sigma_net.net.2.weight"] = _linear_w(rng, 1 + hp["geo_feat_dim"], hs)
    hc = hp["hidden_dim_color"]
    st["color_net.net.0.weight"] = _linear_w(rng, hc, 16 + hp["geo_feat_dim"] + hp["individual_embedding_dim"])
    st["color_net.net.1.weight"] = _linear_w(rng, 3, hc)
    if torso:
        tor = GridLayout(2, log2_hashmap_size=16, desired_resolution=2048, gridtype="tiled")
        st["torso_embedder.embeddings"] = table(tor, 3)
        st["torso_embedder.offsets"] = torch.from_numpy(tor.offsets.copy())
        din = (2 + 2 * 2 * 10) + (6 + 6 * 2 * 4) + hp["torso_individual_embedding_dim"]
        st["torso_deform_net.net.0.weight"] = _linear_w(rng, 64, din)
        st["torso_deform_net.net.1.weight"] = _linear_w(rng, 64, 64)
        st["torso_deform_net.net.2.weight"] = _linear_w(rng, 2, 64)
        st["torso_canonicial_net.net.0.weight"] = _linear_w(rng, 32, tor.output_dim + din)
        st["torso_canonicial_net.net.1.weight"] = _linear_w(rng, 32, 32)
        st["torso_canonicial_net.net.2.weight"] = _linear_w(rng, 4, 32)
    return st


# ----------------------------------------------------------------------------- camera / inputs
def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def camera_pose(t: int) -> torch.Tensor:
    """c2w [4,4] fp32: camera ~4 units from the origin looking at it, slow head-like sway."""
    R = _rot_y(math.radians(10.0) * math.sin(2 * math.pi * t / 100)) @ _rot_x(math.radians(5.0) * math.sin(2 * math.pi * t / 70)) @ np.diag([1.0, -1.0, -1.0])
    # the camera looks along +z of camera space (get_rays: dir = (x, y, 1) @ R^T), so it sits at -4 * R[:, 2]
    pos = -(R @ np.array([0.0, 0.0, 4.0]))
    pos = pos + 0.05 * np.array([math.sin(2 * math.pi * t / 50), math.cos(2 * math.pi * t / 80), 0.0])
    P = np.eye(4, dtype=np.float64)
    P[:3, :3] = R
    P[:3, 3] = pos
    return torch.from_numpy(P.astype(np.float32))


def get_rays(pose4x4: torch.Tensor, intrinsics, H: int, W: int):
    """rays_o, rays_d [1, H*W, 3] for one pose (utils.py:283-364, N=-1 branch; same op order)."""
    fx, fy, cx, cy = intrinsics
    poses = pose4x4.view(1, 4, 4)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t().reshape([1, H * W]) + 0.5
    j = j.t().reshape([1, H * W]) + 0.5
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def get_bg_coords(H: int, W: int):
    X = torch.arange(H) / (H - 1) * 2 - 1
    Y = torch.arange(W) / (W - 1) * 2 - 1
    xs, ys = torch.meshgrid(X, Y, indexing="ij")
    return torch.cat([xs.reshape(-1, 1), ys.reshape(-1, 1)], dim=-1).unsqueeze(0)


def convert_poses(poses: torch.Tensor) -> torch.Tensor:
    """[B,4,4] -> [B,6] = euler 'XYZ' (pytorch3d convention) + translation."""
    M = poses[:, :3, :3]
    out = torch.empty(poses.shape[0], 6, dtype=torch.float32)
    out[:, 0] = torch.atan2(-M[:, 1, 2], M[:, 2, 2])
    out[:, 1] = torch.asin(M[:, 0, 2])
    out[:, 2] = torch.atan2(-M[:, 0, 1], M[:, 0, 0])
    out[:, 3:] = poses[:, :3, 3]
    return out


def cond_sequence(T: int, seed=4, dim=204) -> torch.Tensor:
    """[T,1,dim] landmark-like conditioning, clamped to +-1.5 (infer_lm3d_clamp_std, base.yaml:121)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = np.clip(rng.standard_normal((T, 1, dim)), -1.5, 1.5).astype(np.float32)
    return torch.from_numpy(c)


def cond_window(features: torch.Tensor, index: int, smo_win_size=5) -> torch.Tensor:
    """get_audio_features(features, att_mode=2, index): centred window, zero-padded at the clip edges."""
    left = index - smo_win_size // 2
    right = index + (smo_win_size - smo_win_size // 2)
    pad_l = max(0, -left)
    pad_r = max(0, right - features.shape[0])
    w = features[max(left, 0):min(right, features.shape[0])]
    if pad_l:
        w = torch.cat([torch.zeros_like(w[:1]).expand(pad_l, *w.shape[1:]), w], 0)
    if pad_r:
        w = torch.cat([w, torch.zeros_like(w[:1]).expand(pad_r, *w.shape[1:])], 0)
    return w.contiguous()


def bg_image(H: int, W: int, seed=5) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.uniform(0, 1, size=(1, H * W, 3)).astype(np.float32))


class Scene:
    """Bundle of everything one BASELINE config needs, generated lazily per frame."""

    def __init__(self, H=512, W=512, T=250, torso=True, max_steps=16, density_scale=1.0, table_amp=0.5,
                 table_decay=0.0, T_thresh=0.01, seed=0):
        self.H, self.W, self.T, self.torso = H, W, T, torso
        self.hparams = may_hparams(max_steps=max_steps)
        self.state = make_state(torso=torso, hparams=self.hparams, table_amp=table_amp, table_decay=table_decay, seed=seed)
        self.density_scale = float(density_scale)
        self.T_thresh = float(T_thresh)
        self.intrinsics = may_intrinsics(H, W)
        self.cond = cond_sequence(T)
        self.bg_color = bg_image(H, W)
        self.bg_coords = get_bg_coords(H, W)

    def pose(self, t):
        return camera_pose(t)

    def frame_inputs(self, t):
        """The exact argument set the driver passes to render() for frame t (genefacepp_infer.py:476-479)."""
        P = self.pose(t)
        rays_o, rays_d = get_rays(P, self.intrinsics, self.H, self.W)
        return {
            "rays_o": rays_o, "rays_d": rays_d,
            "cond": cond_window(self.cond, t, self.hparams["smo_win_size"]),
            "bg_coords": self.bg_coords, "poses": convert_poses(P.view(1, 4, 4)),
            "bg_color": self.bg_color, "pose4x4": P,
        }


# ------------------------------------------------------------------------------------------------ SR head
def hashed_uniform(n: int, salt: int, scale: float = 1.0) -> torch.Tensor:
    """n reproducible values in [-scale/2, scale/2): an integer hash of the index, no RNG and no libm, so the golden
    generator (which runs next to the reference) and the tests (which run anywhere) build bit-identical tensors."""
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    return torch.from_numpy(((x.astype(np.float64) / 2.0 ** 32 - 0.5) * scale).astype(np.float32))


def synthetic_sr_state(shapes: dict, seed: int = 0) -> dict:
    """A lively synthetic state for the SR head: `shapes` maps state_dict keys (as the module reports them) to shapes.
    Convolution / affine weights ~U(-1,1), biases and noise strengths small but non-zero so every term of the layers is live;
    the FIR buffers keep their constructor values (they are constants of the architecture)."""
    out = {}
    for n, (k, shape) in enumerate(sorted(shapes.items())):
        if k.endswith("resample_filter"):
            continue
        numel = int(np.prod(shape)) if len(shape) else 1
        if k.endswith("noise_strength"):
            v = hashed_uniform(1, seed * 1000 + n, 0.2).reshape(())
        elif k.endswith("affine.bias"):
            v = 1.0 + hashed_uniform(numel, seed * 1000 + n, 0.5).reshape(shape)
        elif k.endswith("bias"):
            v = hashed_uniform(numel, seed * 1000 + n, 0.4).reshape(shape)
        else:
            v = hashed_uniform(numel, seed * 1000 + n, 2.0).reshape(shape)
        out[k] = v
    return out


def make_torso_sr_state(hparams, seed: int = 0) -> dict:
    """Synthetic state of the torso-SR model (modules/radnerfs/radnerf_torso_sr.py:17-61): the head + torso-grid state of
    make_state, torso MLPs with the SR variant's input widths (freq-encoded 2-D position 42 + code 8 + freq-encoded 7 jaw
    landmarks 126 [+ 16 head-aware features]), the head-colour encoder, the eye-blink modules and the SR head (noise
    strengths zero, so the reference's default noise_mode='random' is deterministic)."""
    hp = hparams
    st = make_state(torso=True, hparams=hp, seed=seed)
    salt = [seed * 100000 + 500]

    def lin(out_f, in_f, bias=False):
        salt[0] += 1
        w = hashed_uniform(out_f * in_f, salt[0], 2.0 / math.sqrt(in_f)).reshape(out_f, in_f)
        if not bias:
            return w
        salt[0] += 1
        return w, hashed_uniform(out_f, salt[0], 2.0 / math.sqrt(in_f))

    din = (2 + 2 * 2 * 10) + (14 + 14 * 2 * 4) + hp["torso_individual_embedding_dim"] + (16 if hp.get("torso_head_aware") else 0)
    st["torso_deform_net.net.0.weight"] = lin(64, din)
    st["torso_deform_net.net.1.weight"] = lin(64, 64)
    st["torso_deform_net.net.2.weight"] = lin(2, 64)
    st["torso_canonicial_net.net.0.weight"] = lin(32, 32 + din)
    st["torso_canonicial_net.net.1.weight"] = lin(32, 32)
    st["torso_canonicial_net.net.2.weight"] = lin(4, 32)
    if hp.get("torso_head_aware"):
        for i, (o, c) in zip((0, 2, 4), ((16, 4), (32, 16), (16, 32))):
            st[f"head_color_weights_encoder.{i}.weight"], st[f"head_color_weights_encoder.{i}.bias"] = lin(o, c, bias=True)
    if hp.get("add_eye_blink_cond"):
        half = hp["cond_out_dim"] // 2
        salt[0] += 1
        st["blink_embedding.weight"] = hashed_uniform(half, salt[0], 1.0).reshape(1, half)
        st["blink_encoder.0.weight"], st["blink_encoder.0.bias"] = lin(half, half, bias=True)
        st["blink_encoder.1.weight"], st["blink_encoder.1.bias"] = lin(hp["eye_blink_dim"], half, bias=True)
    from .superres import Superresolution
    net = Superresolution(channels=3)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sr = synthetic_sr_state(shapes, seed=seed + 5)
    for k, v in net.state_dict().items():
        v = sr.get(k, v)
        st["sr_net." + k] = torch.zeros_like(v) if k.endswith("noise_strength") else v.clone()
    return st


def lm68_sequence(T: int, salt: int = 91) -> torch.Tensor:
    """[T, 136] synthetic 2-D landmarks in [-1, 1] (the torso-SR model reads landmarks 5..11, radnerf_torso_sr.py:84)."""
    return hashed_uniform(T * 136, salt, 2.0).reshape(T, 136)


def make_head_sr_state(hparams, seed: int = 0) -> dict:
    """Synthetic state of the head-SR model (modules/radnerfs/radnerf_sr.py:50-115): make_state's head + blink modules + SR head
    (noise strengths zero) + `lambda_ambient`."""
    full = make_torso_sr_state({**hparams, "torso_head_aware": False}, seed=seed)
    drop = ("torso_", "density_grid_torso", "head_color_weights_encoder.")
    st = {k: v for k, v in full.items() if not k.startswith(drop)}
    st["lambda_ambient"] = torch.tensor([1.0])
    return st
