"""oracle/make_sr_golden.py -- golden vectors for the super-resolution head.

Runs only where /root/reference exists.  Builds the reference's `Superresolution(channels=3)` (radnerf_sr.py:15-48: StyleGAN2
synthesis blocks from modules/eg3ds) on CPU, loads a reproducible synthetic state (genefaceplusplus_b200.scene.synthetic_sr_state:
integer-hash values, so nothing but the key/shape list has to be stored), runs it with `noise_mode='const'` on a 64x64 image
(exercises the antialiased up-sampling to 256) and on a 256x256 image, and records in tests/golden/sr_head.npz:
the reference's state_dict key/shape list (the drop-in contract), a 64x64 crop of each 512x512 output and per-channel sums.

Usage:  python -m oracle.make_sr_golden
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from genefaceplusplus_b200 import scene as scn  # noqa: E402
from oracle import ops, ref_shim  # noqa: E402

CROP = (slice(200, 264), slice(300, 364))


def sr_inputs():
    small = scn.hashed_uniform(3 * 64 * 64, 77, 1.0).reshape(1, 3, 64, 64) + 0.5
    full = scn.hashed_uniform(3 * 256 * 256, 78, 1.0).reshape(1, 3, 256, 256) + 0.5
    return {"in64": small, "in256": full}


def main():
    cwd = os.getcwd()
    ops.build()
    ref_shim.install(ops)
    from modules.radnerfs.radnerf_sr import Superresolution
    with torch.no_grad():
        net = Superresolution(channels=3).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(scn.synthetic_sr_state(shapes, seed=3), strict=False)        # FIR buffers keep their values
    out = {}
    with torch.no_grad():
        for name, x in sr_inputs().items():
            y = net(x.clone(), noise_mode="const")
            assert y.shape == (1, 3, 512, 512)
            out[f"{name}_crop"] = y[0, :, CROP[0], CROP[1]].numpy().astype(np.float32)
            out[f"{name}_sum"] = y.double().sum(dim=(0, 2, 3)).numpy()
            out[f"{name}_abssum"] = y.double().abs().sum(dim=(0, 2, 3)).numpy()
            print(name, "range", float(y.min()), float(y.max()), "sum", out[f"{name}_sum"])
    out["resample_filter"] = net.resample_filter.numpy()
    meta = dict(source="reference Superresolution (radnerf_sr.py:15-48) on CPU via oracle/ref_shim.py, noise_mode='const'",
                state="genefaceplusplus_b200.scene.synthetic_sr_state(shapes, seed=3)", crop=[200, 264, 300, 364],
                shapes={k: list(v) for k, v in shapes.items()}, torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sr_head.npz"), **out)
    print("wrote tests/golden/sr_head.npz")


if __name__ == "__main__":
    main()
