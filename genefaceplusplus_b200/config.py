"""Configuration surface of the render path.

The reference reads a yaml chain (egs/datasets/May/lm3d_radnerf_torso.yaml ->
egs/egs_bases/radnerf/lm3d_radnerf.yaml -> base.yaml) into a global `hparams` dict and forwards
every key to `render(**hparams)` (inference/genefacepp_infer.py:476-479).  The renderer only
needs the keys below; `may_hparams()` restates their May values so the GPU box (no reference
tree, no yaml) can build a May-shaped model.  A real hparams dict loaded by the reference's
`set_hparams` can be passed instead -- extra keys are ignored exactly as `**kwargs` does.
"""
import math

import numpy as np

# values from egs/egs_bases/radnerf/base.yaml:57-103 with the overrides of
# egs/egs_bases/radnerf/lm3d_radnerf.yaml:6-13 and egs/datasets/May/lm3d_radnerf*.yaml
_MAY = {
    "cuda_ray": True,
    "max_steps": 16,
    "min_near": 0.05,
    "bound": 1,
    "grid_size": 128,
    "desired_resolution": 2048,
    "log2_hashmap_size": 16,
    "dt_gamma": 0.00390625,
    "density_thresh": 10,
    "density_thresh_torso": 0.01,
    "torso_shrink": 0.8,
    "grid_type": "tiledgrid",
    "grid_interpolation_type": "linear",
    "with_att": True,
    "torso_head_aware": False,
    "num_layers_sigma": 3,
    "hidden_dim_sigma": 128,
    "geo_feat_dim": 128,
    "num_layers_color": 2,
    "hidden_dim_color": 128,
    "cond_out_dim": 64,
    "num_layers_ambient": 3,
    "hidden_dim_ambient": 128,
    "ambient_coord_dim": 3,
    "individual_embedding_num": 13000,
    "individual_embedding_dim": 4,
    "torso_individual_embedding_dim": 8,
    "cond_type": "idexp_lm3d_normalized",
    "nerf_keypoint_mode": "lm68",
    "cond_win_size": 1,
    "smo_win_size": 5,
    "cond_dropout_rate": 0.0,
}


def may_hparams(**overrides):
    hp = dict(_MAY)
    hp.update(overrides)
    return hp


# camera of the May dataset: data_gen/runs/binarizer_nerf.py:332-333 (focal 1015, centre 112 at 224 px),
# rescaled to the render size in tasks/radnerfs/dataset_utils.py:216-230
FOCAL_224 = 1015.0
CENTER_224 = 112.0


def may_intrinsics(H: int, W: int):
    fx = FOCAL_224 * (H / 2) / CENTER_224
    fy = FOCAL_224 * (W / 2) / CENTER_224
    return (fx, fy, H / 2, W / 2)


class GridLayout:
    """Level layout of one multi-resolution tiled/hash grid.

    Restates GridEncoder.__init__ (modules/radnerfs/encoders/gridencoder/grid.py:98-136): per-level
    table sizes `min(2^log2_hashmap, (res+1)^D)` rounded up to 8, and the per-level `scale`/`resolution`
    the kernel derives at run time (gridencoder.cu:137-139):  scale = exp2f(l*S)*H - 1, res = ceil(scale)+1
    with S = float32(log2(per_level_scale)).
    """

    def __init__(self, input_dim, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16,
                 desired_resolution=2048, gridtype="tiled", align_corners=False, interpolation="linear"):
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.base_resolution = base_resolution
        self.per_level_scale = float(np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1)))
        self.gridtype_id = {"hash": 0, "tiled": 1}[gridtype]
        self.interp_id = {"linear": 0, "smoothstep": 1}[interpolation]
        self.align_corners = align_corners
        max_params = 2 ** log2_hashmap_size
        offsets, off = [], 0
        for i in range(num_levels):
            res = int(np.ceil(base_resolution * self.per_level_scale ** i))
            n = min(max_params, (res if align_corners else res + 1) ** input_dim)
            n = int(np.ceil(n / 8) * 8)
            offsets.append(off)
            off += n
        offsets.append(off)
        self.offsets = np.asarray(offsets, dtype=np.int32)
        self.n_entries = off
        self.S = np.float32(np.log2(self.per_level_scale))

    @property
    def output_dim(self):
        return self.num_levels * self.level_dim


def cascade_count(bound) -> int:
    # modules/radnerfs/renderer.py:70
    return 1 + math.ceil(math.log2(bound))
