"""Driver-level shim: the render loop of `GeneFace2Infer.forward_secc2video`
(inference/genefacepp_infer.py:433-486) without its T serial `render()` calls, per-frame `.cpu()` and fp32 frame stack.

`render_driver_batch` takes the very `batch` dict `prepare_batch_from_inp` builds (:203-269: lists `rays_o`, `rays_d`,
`cond_wins`, `poses`, `eye_area_percent`, plus `bg_coords`, `bg_img`) and returns the frames the driver would have produced,
rendering `frames_per_call` frames per libgfpp call with the conditioning nets batched over the clip.  The maintainer's
change in forward_secc2video is

    pred_rgbs = render_driver_batch(self.secc2video_model, batch, T_thresh=inp['raymarching_end_threshold'])   # [T,3,H,W] in [0,1], on the GPU
    out_imgs = video_uint8(pred_rgbs).cpu().numpy()                                                           # what imageio gets (:505)

For drivers that can hand over c2w poses instead of materialised rays (1.5 GB for 250 frames), `model.render_clip` generates
the rays in-kernel; this shim is for the unmodified batch layout.  Non-SR head / head+torso models (SR models render per frame).
"""
import torch

from .dist import to_uint8


def cond_feat_from_windows(model, cond_wins, eye_area_percent=None):
    """All frames' `cal_cond_feat` in one pass: cond_wins = list/stack of T windows [S,cond_win_size,C] (what the driver passes per frame,
    genefacepp_infer.py:420-422) -> [T,64].  Equals calling model.cal_cond_feat window by window."""
    wins = torch.stack([w.reshape(w.shape[0], -1) for w in cond_wins]) if not torch.is_tensor(cond_wins) else cond_wins.reshape(cond_wins.shape[0], cond_wins.shape[1], -1)
    T, S, C = wins.shape
    dev = model.density_bitfield.device
    wins = wins.to(dev).float()
    with torch.autocast(dev.type, enabled=False):
        feat = model.cond_prenet(wins.reshape(T * S, model.cond_win_size, C // model.cond_win_size))
        if model.add_eye_blink_cond:
            eye = None
            if eye_area_percent is not None and model.forwards_eye_area:
                eye = torch.as_tensor(eye_area_percent, dtype=torch.float32).reshape(T)
            feat = model._add_blink(feat, eye, T)
        feat = feat.view(T, S, -1)
        feat = model.cond_att_net.forward_batched(feat) if model.with_att else feat[:, S // 2]   # centre row, like cal_cond_feat_clip
    return feat


@torch.no_grad()
def render_driver_batch(model, batch, T_thresh=1e-2, frames_per_call=32, dt_gamma=None, max_steps=None):
    """-> [T,3,H,W] fp32 in [0,1] on the model's device (H = W = sqrt(N))."""
    hp = model.hparams
    T = len(batch["poses"])
    dev = model.density_bitfield.device
    feat = cond_feat_from_windows(model, batch["cond_wins"], batch.get("eye_area_percent"))
    N = batch["rays_o"][0].numel() // 3
    H = int(round(N ** 0.5))
    out = torch.empty(T, N, 3, device=dev, dtype=torch.float32)
    for s in range(0, T, frames_per_call):
        e = min(T, s + frames_per_call)
        rays_o = torch.stack([batch["rays_o"][i].reshape(N, 3) for i in range(s, e)])
        rays_d = torch.stack([batch["rays_d"][i].reshape(N, 3) for i in range(s, e)])
        pose6 = torch.stack([batch["poses"][i].reshape(6) for i in range(s, e)]) if model.has_torso else None
        res = model.render_frames(feat[s:e], rays_o=rays_o, rays_d=rays_d, pose6=pose6,
                                  bg_coords=batch["bg_coords"] if model.has_torso else None, bg_color=batch.get("bg_img"),
                                  dt_gamma=hp["dt_gamma"] if dt_gamma is None else dt_gamma,
                                  max_steps=hp["max_steps"] if max_steps is None else max_steps, T_thresh=T_thresh, want_torso_maps=False)
        out[s:e] = res["rgb_map"]
    return out.view(T, H, N // H, 3).permute(0, 3, 1, 2)


def video_uint8(pred_rgbs):
    """[T,3,H,W] in [0,1] -> uint8 [T,H,W,3]: the driver's `((x*2-1).clamp(-1,1) + 1)/2 * 255 -> int` (:488-505) in one step."""
    return to_uint8(pred_rgbs.permute(0, 2, 3, 1))
