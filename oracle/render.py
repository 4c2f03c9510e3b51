"""oracle/render.py -- TEST INFRASTRUCTURE (CPU checker; never imported by the product path).

Self-contained CPU fp32 restatement of the reference's per-frame render path, built on the C
restatement of the native ops (oracle/native_ops.c via oracle/ops.py) and plain torch CPU ops for
the dense layers.  Needs neither /root/reference nor a GPU, so it travels to the GPU box.

Follows (paths relative to /root/reference):
  NeRFRenderer.render   inference branch      modules/radnerfs/renderer.py:286-399
  RADNeRF.cal_cond_feat / forward             modules/radnerfs/radnerf.py:88-141
  AudioNet / AudioAttNet / MLP                modules/radnerfs/cond_encoder.py:98-202
  RADNeRFTorso.forward_torso / render         modules/radnerfs/radnerf_torso.py:51-199
  GridEncoder.forward ([0,1] mapping)         modules/radnerfs/encoders/gridencoder/grid.py:148-164

Pinned against the reference's own classes by oracle/validate_against_reference.py (bit-exact on
CPU in this container; see DESIGN.md "oracle pinning").

`OracleModel.render` keeps the reference's host-driven multi-round loop *verbatim in structure*
(n_step schedule, 128-row padding, boolean-mask compaction) because the round schedule is part of
the result.  It additionally reports statistics the kernels and bench need:
  S         number of valid evaluated samples (rows with delta != 0 that the compositor consumed)
  schedule  [(n_alive, n_step), ...]
  B_total   sum of n_step  (the per-ray sample cap the schedule produced)
  knife     per-ray min relative distance of T to T_thresh over consumed samples (termination margin)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


def _mlp(x, weights, linear=None):
    """cond_encoder.py:197-202 -- bias-free Linear, ReLU between layers, none after the last."""
    for i, w in enumerate(weights):
        x = F.linear(x, w) if linear is None else linear(x, w)
        if i != len(weights) - 1:
            x = F.relu(x)
    return x


class OracleModel:
    def __init__(self, state, hparams, torso=None, fixed_order_linear=False, backend=None, device="cpu", collect_stats=True):
        """backend: object with the wrapper-level native-op API of oracle/ops.py (default: the C restatement on CPU;
        oracle/gpu_ref_ops.py serves the same loop from the reference's own compiled CUDA kernels for the GPU baseline)."""
        self.ops = backend if backend is not None else ops
        self.dev = torch.device(device)
        self.collect_stats = collect_stats
        self.st = {k: v.detach().to(self.dev) for k, v in state.items()}
        self.hp = dict(hparams)
        self.torso = ("torso_embedder.embeddings" in self.st) if torso is None else torso
        self.bound = self.hp["bound"]
        self.cascade = 1 + math.ceil(math.log2(self.hp["bound"]))
        self.grid_size = self.hp["grid_size"]
        self.min_near = self.hp["min_near"]
        self.density_scale = 1
        self.mean_density_torso = 0  # plain attribute in the reference, not in the checkpoint (radnerf_torso.py:22)
        self.density_thresh_torso = self.hp["density_thresh_torso"]
        self.torso_shrink = self.hp["torso_shrink"]
        self._linear = self.ops.linear if fixed_order_linear else None
        self.per_level_scale = float(np.exp2(np.log2(self.hp["desired_resolution"] * self.bound / 16) / 15))
        self.per_level_scale_amb = float(np.exp2(np.log2(self.hp["desired_resolution"] / 16) / 15))
        self.per_level_scale_torso = float(np.exp2(np.log2(2048 / 16) / 15))
        self.gridtype_id = {"tiledgrid": 1, "hashgrid": 0}[self.hp["grid_type"]]
        self.interp_id = {"linear": 0, "smoothstep": 1}[self.hp["grid_interpolation_type"]]
        # SR variants (radnerf_sr.py / radnerf_torso_sr.py): NeRF at 256x256 + the 256 -> 512 head; the torso-SR model also
        # swaps the torso field's inputs (jaw landmarks instead of the pose, optional head-colour features)
        self.with_sr = bool(self.hp.get("with_sr", False)) and any(k.startswith("sr_net.") for k in self.st)
        self._sr_net = None

    # ---- conditioning (radnerf.py:88-106; cond_encoder.py:98-180) ----
    def cal_cond_feat(self, cond, eye_area_percent=None):
        st = self.st
        x = cond.float().permute(0, 2, 1)  # [b, c, t=1]
        for i in (0, 2, 4, 6):
            x = F.conv1d(x, st[f"cond_prenet.encoder_conv.{i}.weight"], st[f"cond_prenet.encoder_conv.{i}.bias"], stride=1, padding=1)
            x = F.leaky_relu(x, 0.02)
        x = x.squeeze(-1)
        x = F.leaky_relu(F.linear(x, st["cond_prenet.encoder_fc1.0.weight"], st["cond_prenet.encoder_fc1.0.bias"]), 0.02)
        x = F.linear(x, st["cond_prenet.encoder_fc1.2.weight"], st["cond_prenet.encoder_fc1.2.bias"])  # [b, 64]
        if self.hp.get("add_eye_blink_cond", False):   # radnerf.py:97-103
            E = self.hp["eye_blink_dim"]
            pct = (torch.zeros(1, 1, device=x.device) if eye_area_percent is None
                   else torch.as_tensor(eye_area_percent, dtype=torch.float32, device=x.device).reshape(1, 1))
            b = st["blink_embedding.weight"][0].reshape(1, -1) * pct
            b = F.linear(b, st["blink_encoder.0.weight"], st["blink_encoder.0.bias"])
            b = F.linear(b, st["blink_encoder.1.weight"], st["blink_encoder.1.bias"])
            x = x.clone()
            x[..., :E] = x[..., :E] + b.expand(x[..., :E].shape)
        if not self.hp["with_att"]:
            return x
        seq = x.shape[0]
        y = x[:, :x.shape[1]].permute(1, 0).unsqueeze(0)  # [1, c, b]
        for i in (0, 2, 4, 6, 8):
            y = F.conv1d(y, st[f"cond_att_net.attentionConvNet.{i}.weight"], st[f"cond_att_net.attentionConvNet.{i}.bias"], stride=1, padding=1)
            y = F.leaky_relu(y, 0.02)
        y = F.linear(y.view(1, seq), st["cond_att_net.attentionNet.0.weight"], st["cond_att_net.attentionNet.0.bias"])
        y = torch.softmax(y, dim=1).view(seq, 1)
        return torch.sum(y * x, dim=0)  # [64]

    # ---- field query (radnerf.py:108-141) ----
    def _grid(self, x, bound, emb, off, pls):
        x01 = (x + bound) / (2 * bound)
        return self.ops.grid_encode(x01, self.st[emb], self.st[off], pls, 16, self.gridtype_id, False, self.interp_id)

    def forward(self, position, direction, cond_feat, individual_code):
        st = self.st
        n = position.shape[0]
        cond = cond_feat.repeat([n, 1])
        pos_feat = self._grid(position, self.bound, "position_embedder.embeddings", "position_embedder.offsets", self.per_level_scale)
        amb_in = torch.cat([pos_feat, cond], dim=1)
        amb_logit = _mlp(amb_in, [st[f"ambient_net.net.{i}.weight"] for i in range(self.hp["num_layers_ambient"])], self._linear).float()
        amb_pos = torch.tanh(amb_logit)
        amb_feat = self._grid(amb_pos, 1, "ambient_embedder.embeddings", "ambient_embedder.offsets", self.per_level_scale_amb)
        h = torch.cat([pos_feat, amb_feat], dim=-1)
        h = _mlp(h, [st[f"sigma_net.net.{i}.weight"] for i in range(self.hp["num_layers_sigma"])], self._linear)
        sigma = torch.exp(h[..., 0].float())
        geo = h[..., 1:]
        dfeat = self.ops.sh_encode(direction, 4)
        if individual_code is not None:
            cin = torch.cat([dfeat, geo, individual_code.repeat(n, 1)], dim=-1)
        else:
            cin = torch.cat([dfeat, geo], dim=-1)
        color = torch.sigmoid(_mlp(cin, [st[f"color_net.net.{i}.weight"] for i in range(self.hp["num_layers_color"])], self._linear))
        return sigma, color, amb_pos

    # ---- torso field (radnerf_torso.py:51-84) ----
    def forward_torso(self, x, poses, c):
        st = self.st
        x = x * self.torso_shrink
        enc_pose = self.ops.freq_encode(poses.reshape(-1, 6), 4)
        enc_x = self.ops.freq_encode(x, 10)
        parts = [enc_x, enc_pose.repeat(x.shape[0], 1)]
        if c is not None:
            parts.append(c.repeat(x.shape[0], 1))
        h = torch.cat(parts, dim=-1)
        dx = _mlp(h, [st[f"torso_deform_net.net.{i}.weight"] for i in range(3)], self._linear)
        x = (x + dx).clamp(-1, 1).float()
        xf = self._grid(x, 1, "torso_embedder.embeddings", "torso_embedder.offsets", self.per_level_scale_torso)
        h = torch.cat([xf, h], dim=-1)
        h = _mlp(h, [st[f"torso_canonicial_net.net.{i}.weight"] for i in range(3)], self._linear)
        return torch.sigmoid(h[..., :1]), torch.sigmoid(h[..., 1:]), dx

    # ---- torso field of the SR model (radnerf_torso_sr.py:73-113) ----
    def forward_torso_sr(self, x, poses, c, image, weights_sum, lm68):
        st = self.st
        x = x * self.torso_shrink
        enc_x = self.ops.freq_encode(x, 10)
        jaw = lm68.reshape(1, 68, 2)[:, [5, 6, 7, 8, 9, 10, 11]].reshape(1, -1)
        enc_lm = self.ops.freq_encode(jaw.float(), 4)
        parts = [enc_x]
        if c is not None:
            parts.append(c.repeat(x.shape[0], 1))
        parts.append(enc_lm.repeat(x.shape[0], 1))
        h = torch.cat(parts, dim=-1)
        if self.hp.get("torso_head_aware", False):
            e = torch.cat([image, weights_sum], dim=-1)
            for i in (0, 2, 4):
                e = F.linear(e, st[f"head_color_weights_encoder.{i}.weight"], st[f"head_color_weights_encoder.{i}.bias"])
                if i != 4:
                    e = F.leaky_relu(e, 0.02)
            h = torch.cat([h, e], dim=-1)
        dx = _mlp(h, [st[f"torso_deform_net.net.{i}.weight"] for i in range(3)], self._linear)
        x = (x + dx).clamp(-1, 1).float()
        xf = self._grid(x, 1, "torso_embedder.embeddings", "torso_embedder.offsets", self.per_level_scale_torso)
        h = torch.cat([xf, h], dim=-1)
        h = _mlp(h, [st[f"torso_canonicial_net.net.{i}.weight"] for i in range(3)], self._linear)
        return torch.sigmoid(h[..., :1]), torch.sigmoid(h[..., 1:]), dx

    def sr_net(self):
        if self._sr_net is None:
            from genefaceplusplus_b200.superres import Superresolution   # host-side torch module, itself pinned by tests/golden/sr_head.npz
            net = Superresolution(channels=3).to(self.dev).eval()
            net.load_state_dict({k[len("sr_net."):]: v for k, v in self.st.items() if k.startswith("sr_net.")}, strict=True)
            self._sr_net = net
        return self._sr_net

    # ---- render (renderer.py:286-399 / radnerf_torso.py:86-199), inference branch ----
    @torch.no_grad()
    def render(self, rays_o, rays_d, cond, bg_coords, poses, index=0, dt_gamma=0, bg_color=None, perturb=False,
               force_all_rays=False, max_steps=1024, T_thresh=1e-4, eye_area_percent=None, lm68=None, upscale_torso=False, **kwargs):
        assert not perturb
        st = self.st
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        bg_coords = bg_coords.contiguous().view(-1, 2)
        N = rays_o.shape[0]
        results = {}
        dev = self.dev
        nears, fars = self.ops.near_far_from_aabb(rays_o, rays_d, st["aabb_infer"], self.min_near)
        # renderer.py:308 and the SR classes forward eye_area_percent; the plain torso class does not (radnerf_torso.py:106)
        cond_feat = self.cal_cond_feat(cond, eye_area_percent if (self.with_sr or not self.torso) else None)
        ind_code = st["individual_embeddings"][0] if self.hp["individual_embedding_dim"] > 0 else None

        weights_sum = torch.zeros(N, device=dev)
        depth = torch.zeros(N, device=dev)
        image = torch.zeros(N, 3, device=dev)
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        knife = torch.full((N,), float("inf"), device=dev)
        n_samples = torch.zeros(N, dtype=torch.int32, device=dev)
        schedule = []
        step = 0
        S_valid = 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            schedule.append((n_alive, n_step))
            xyzs, dirs, deltas = self.ops.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                st["density_bitfield"], self.cascade, self.grid_size, nears, fars, 128,
                                                False, dt_gamma, max_steps)
            sigmas, rgbs, _ = self.forward(xyzs, dirs, cond_feat, ind_code)
            sigmas = self.density_scale * sigmas
            if self.collect_stats:
                ids = rays_alive.long().clone()  # statistics only: ray ids before composite kills them
                ws_before = weights_sum[ids].clone()
            self.ops.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
            if self.collect_stats:
                self._stats(ids, n_step, sigmas, deltas, ws_before, T_thresh, knife, n_samples)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        S_valid = int(n_samples.sum().item())

        if bg_color is None:
            bg_color = 1
        if self.torso:
            tcode = st["torso_individual_codes"][0] if self.hp["torso_individual_embedding_dim"] > 0 else None
            thr = min(self.density_thresh_torso, self.mean_density_torso)
            G = self.grid_size
            occ = F.grid_sample(st["density_grid_torso"].view(1, 1, G, G), bg_coords.view(1, -1, 1, 2), align_corners=True).view(-1)
            mask = occ > thr
            torso_alpha = torch.zeros(N, 1, device=dev)
            torso_color = torch.zeros(N, 3, device=dev)
            if mask.any():
                if self.with_sr:
                    a, c, deform = self.forward_torso_sr(bg_coords[mask], poses, tcode, image[mask], weights_sum.unsqueeze(-1)[mask], lm68)
                else:
                    a, c, deform = self.forward_torso(bg_coords[mask], poses, tcode)
                torso_alpha[mask] = a.float()
                torso_color[mask] = c.float()
                results["deform"] = deform
            bg_color = torso_color * torso_alpha + bg_color * (1 - torso_alpha)
            results["torso_alpha_map"] = torso_alpha
            results["torso_rgb_map"] = bg_color
            results["torso_mask"] = mask
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        image = image.view(*prefix, 3).clamp(0, 1)
        d = torch.clamp(depth - nears, min=0) / (fars - nears)
        results["depth_map"] = d.view(*prefix)
        results["rgb_map"] = image
        if self.with_sr:   # radnerf_sr.py:203-210, radnerf_torso_sr.py:229-246 (the 256 is hard-coded there too)
            rgb_image = image.reshape(1, 256, 256, 3).permute(0, 3, 1, 2)
            noise_mode = kwargs.get("sr_noise_mode", "random")
            results["rgb_map"] = rgb_image
            results["sr_rgb_map"] = self.sr_net()(rgb_image, noise_mode=noise_mode).clamp(0, 1)
            if self.torso:
                tb = results["torso_rgb_map"].reshape(1, 256, 256, 3).permute(0, 3, 1, 2)
                results["torso_rgb_map"] = tb
                if upscale_torso:
                    results["sr_torso_rgb_map"] = self.sr_net()(tb, noise_mode=noise_mode).clamp(0, 1)
        # oracle-only extras
        results["weights_sum"] = weights_sum
        results["stats"] = {"S": S_valid, "schedule": schedule, "B_total": int(sum(s for _, s in schedule)),
                            "N": N, "P": int(results["torso_mask"].sum().item()) if self.torso else 0}
        results["knife"] = knife
        results["n_samples"] = n_samples
        return results

    @staticmethod
    def _stats(ids, n_step, sigmas, deltas, ws_before, T_thresh, knife, n_samples):
        """Replay the compositor on the side (statistics only; never feeds back into the image):
        counts the samples each ray consumed and records how close T came to T_thresh."""
        n_alive = ids.shape[0]
        sg = sigmas[: n_alive * n_step].view(n_alive, n_step).float()
        dl = deltas[: n_alive * n_step].view(n_alive, n_step, 2)
        ws = ws_before.clone()
        alive = torch.ones(n_alive, dtype=torch.bool, device=ids.device)
        row_knife = torch.full((n_alive,), float("inf"), device=ids.device)
        row_cnt = torch.zeros(n_alive, dtype=torch.int32, device=ids.device)
        for k in range(n_step):
            valid = alive & (dl[:, k, 0] != 0)
            T = 1 - ws
            alpha = 1 - torch.exp(-sg[:, k] * dl[:, k, 0])
            ws = torch.where(valid, ws + alpha * T, ws)
            margin = (T - T_thresh).abs() / T_thresh
            row_knife = torch.where(valid, torch.minimum(row_knife, margin), row_knife)
            row_cnt = row_cnt + valid.int()
            alive = valid & ~(T < T_thresh)
        knife[ids] = torch.minimum(knife[ids], row_knife)
        n_samples[ids] += row_cnt
