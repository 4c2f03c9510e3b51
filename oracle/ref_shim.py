"""oracle/ref_shim.py -- TEST INFRASTRUCTURE (see oracle/README.md).

Import shim that lets the *unmodified* reference Python (`/root/reference/modules/radnerfs`)
run on CPU in this container, with its four CUDA-only native backends replaced by the
C restatement in oracle/native_ops.c.  Used only by oracle/validate_against_reference.py
and tests/golden/make_golden.py (both run where /root/reference exists; never on the GPU box).

Recipe follows SURVEY.md appendix A:
  * stub the seven third-party modules that modules/radnerfs/utils.py:7-33 and
    renderer.py:2 import at top level but the inference path never calls;
  * pre-register fake `_raymarching_face/_gridencoder/_shencoder/_freqencoder` modules
    (the wrappers do `try: import _x as _backend`, raymarching.py:9-12, grid.py:9-12, ...);
  * replace the wrapper-level callables that force `.cuda()`
    (raymarching.py:33-34, 373-374; freq.py:22).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GFPP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules", "radnerfs"))


def install(ops):
    """Make `import modules.radnerfs...` work on CPU.  `ops` is oracle.ops (the ctypes wrappers).

    Returns the reference's hparams loader `set_hparams`.  Side effects: chdir to the reference
    root (yaml base_config paths are cwd-relative, utils/commons/hparams.py:53-57) and
    sys.path insertion.
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ["trimesh", "mcubes", "lpips", "tensorboardX", "matplotlib", "matplotlib.pyplot", "imageio"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    # native backends with the pybind signatures of SURVEY.md section 8(b)
    grid_mod = types.ModuleType("_gridencoder")

    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
        assert dy_dx is None
        ops.grid_encode_forward_raw(inputs, embeddings, offsets, outputs, B, D, C, L, float(S), H, gridtype, align_corners, interp)

    grid_mod.grid_encode_forward = grid_encode_forward
    sys.modules["_gridencoder"] = grid_mod

    sh_mod = types.ModuleType("_shencoder")

    def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
        assert dy_dx is None and D == 3
        ops.sh_encode_forward_raw(inputs, outputs, B, C)

    sh_mod.sh_encode_forward = sh_encode_forward
    sys.modules["_shencoder"] = sh_mod

    freq_mod = types.ModuleType("_freqencoder")
    freq_mod.freq_encode_forward = lambda inputs, B, D, deg, C, outputs: ops.freq_encode_forward_raw(inputs, B, D, deg, C, outputs)
    sys.modules["_freqencoder"] = freq_mod
    sys.modules["_raymarching_face"] = types.ModuleType("_raymarching_face")

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)

    import modules.radnerfs.raymarching as rm  # noqa: E402
    import modules.radnerfs.raymarching.raymarching as rm_inner  # noqa: E402

    for target in (rm, rm_inner):
        target.near_far_from_aabb = ops.near_far_from_aabb
        target.march_rays = ops.march_rays
        target.composite_rays = ops.composite_rays

    # FreqEncoder's autograd.Function forces .cuda(): override the module-level forward
    from modules.radnerfs.encoders.freqencoder.freq import FreqEncoder  # noqa: E402

    def freq_forward(self, inputs, **kwargs):
        prefix = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim).float().contiguous()
        return ops.freq_encode(flat, self.degree).reshape(prefix + [self.output_dim])

    FreqEncoder.forward = freq_forward

    # torch>=2.4 deprecates torch.cuda.amp.custom_fwd; on CPU they are no-ops already.
    from utils.commons.hparams import set_hparams  # noqa: E402

    return set_hparams


def build_reference_model(set_hparams, torso: bool, overrides: str = ""):
    """Instantiate the reference's RADNeRF / RADNeRFTorso from the May yaml chain."""
    cfg = "egs/datasets/May/lm3d_radnerf_torso.yaml" if torso else "egs/datasets/May/lm3d_radnerf.yaml"
    hp = set_hparams(cfg, hparams_str=overrides, print_hparams=False)
    if torso:
        from modules.radnerfs.radnerf_torso import RADNeRFTorso as cls
    else:
        from modules.radnerfs.radnerf import RADNeRF as cls
    with torch.no_grad():
        model = cls(hp).eval()
    return model, hp
