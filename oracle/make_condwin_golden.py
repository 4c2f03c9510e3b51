"""oracle/make_condwin_golden.py -- golden vectors for the audio-window conditioning pre-net (cond_win_size != 1).

Runs only where /root/reference exists.  Builds the reference's own `AudioNet(dim_in, 64, win_size)`
(modules/radnerfs/cond_encoder.py:98-143: four Conv1d k3 p1 with the win-size dependent strides, then two Linear layers) for the
window sizes it supports (1, 2, 3, 4, 16 -- its `win_size == [5, 8]` branch can never be taken, so 5 and 8 raise), gives it a
reproducible state (genefaceplusplus_b200.scene.hashed_uniform: integer-hash values, so only outputs have to be stored) and records
its outputs on hashed windows in tests/golden/cond_win.npz.
tests/test_host_logic.py replays them through genefaceplusplus_b200.renderer._AudioNet (same parameter names).

Usage:  python -m oracle.make_condwin_golden
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

CASES = [(1, 204), (2, 29), (3, 44), (4, 29), (16, 29)]   # (win_size, dim_in): lm3d / deepspeech / esperanto widths


def condwin_state(shapes, win):
    """Reproducible AudioNet state: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from integer hashes, salted by key order and window size."""
    st = {}
    for i, (k, v) in enumerate(shapes.items()):
        fan = max(1, v[0].numel()) if v.dim() > 1 else 8
        st[k] = scn.hashed_uniform(v.numel(), 9000 + 100 * win + i, 2.0 / float(np.sqrt(fan))).reshape(v.shape)
    return st


def main():
    cwd = os.getcwd()
    ops.build()
    ref_shim.install(ops)
    from modules.radnerfs.cond_encoder import AudioNet
    out, unsupported = {}, []
    for win, din in CASES:
        net = AudioNet(din, 64, win_size=win).eval()
        state = condwin_state(net.state_dict(), win)
        net.load_state_dict(state, strict=True)
        x = scn.hashed_uniform(5 * win * din, 7000 + win, 3.0).reshape(5, win, din)
        with torch.no_grad():
            y = net(x)
        assert y.shape == (5, 64), y.shape
        out[f"w{win}_x"] = x.numpy().astype(np.float32)
        out[f"w{win}_y"] = y.numpy().astype(np.float32)
    for win in (5, 8, 7):
        try:
            AudioNet(29, 64, win_size=win)
        except ValueError:
            unsupported.append(win)
    meta = dict(source="reference AudioNet (cond_encoder.py:98-143) on CPU", cases=CASES, unsupported=unsupported, torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cond_win.npz"), **out)
    print("wrote tests/golden/cond_win.npz; win sizes the reference rejects:", unsupported)


if __name__ == "__main__":
    main()
