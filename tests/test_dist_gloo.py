"""N>1 path on CPU: frame sharding + the single all-gather, world_size 2, gloo backend (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genefaceplusplus_b200 import dist as gdist


def test_frame_block_partition():
    for T in (1, 7, 250, 2000, 2001):
        for world in (1, 2, 3, 8):
            blocks = [gdist.frame_block(T, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == T
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == gdist.padded_block_len(T, world)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _FakeRenderer:
    """Stands in for the GPU model: frame t renders to the constant image t (so order errors are visible)."""
    density_bitfield = torch.zeros(1)

    def cal_cond_feat_clip(self, cond_seq):
        return cond_seq.reshape(cond_seq.shape[0], -1)[:, :4].clone()

    def render_clip(self, poses, intrinsics, H, W, cond_feat=None, **kw):
        t = poses[:, 0, 3]
        assert torch.allclose(cond_feat[:, 0], t)      # each rank got ITS slice of the conditioning
        return t.view(-1, 1, 1).expand(-1, H * W, 3).contiguous()


class _FakeSRRenderer(_FakeRenderer):
    """SR-model shape of the clip API: blink-conditioned features from the FULL sequence, per-frame landmarks sliced to the block,
    [t,3,4,4] frames out, extra keyword arguments passed through."""

    def cal_cond_feat_clip(self, cond_seq, eye_area_percent=None):
        assert eye_area_percent is not None and eye_area_percent.shape[0] == cond_seq.shape[0]
        return cond_seq.reshape(cond_seq.shape[0], -1)[:, :4].clone() + eye_area_percent.view(-1, 1)

    def render_clip(self, poses, intrinsics, H, W, cond_feat=None, lm68_seq=None, sr_noise_mode=None, **kw):
        t = poses[:, 0, 3]
        assert torch.allclose(cond_feat[:, 0], t + 0.25) and torch.allclose(lm68_seq[:, 0], t) and sr_noise_mode == "const"
        return t.view(-1, 1, 1, 1).expand(-1, 3, 4, 4).contiguous()


def _worker(rank, world, port, T, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    poses = torch.eye(4).repeat(T, 1, 1)
    poses[:, 0, 3] = torch.arange(T, dtype=torch.float32)
    cond = torch.arange(T, dtype=torch.float32).view(T, 1, 1).expand(T, 1, 8).contiguous()
    out = gdist.render_clip_sharded(_FakeRenderer(), poses, (1, 1, 0, 0), 2, 2, cond, as_uint8=False)
    u8 = gdist.gather_frames(gdist.to_uint8(torch.full((gdist.frame_block(T, rank, world)[1] - gdist.frame_block(T, rank, world)[0], 4, 3), 0.5)), T)
    lm = torch.arange(T, dtype=torch.float32).view(T, 1).expand(T, 136).contiguous()
    sr = gdist.render_clip_sharded(_FakeSRRenderer(), poses, (1, 1, 0, 0), 2, 2, cond, eye_area_percent=torch.full((T,), 0.25), lm68_seq=lm,
                                   sr_noise_mode="const")
    assert sr.shape == (T, 3, 4, 4) and sr[:, 0, 0, 0].tolist() == [float(t) for t in range(T)]     # SR-model clips shard the same way
    q.put((rank, out[:, 0, 0].tolist(), tuple(u8.shape), int(u8.max())))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [6, 7])
def test_sharded_render_and_gather_world2(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, frames, shape, mx in res:
        assert frames == [float(t) for t in range(T)], (rank, frames)   # every rank holds the whole clip in video order
        assert shape == (T, 4, 3) and mx == 127


def test_to_uint8_matches_the_driver_conversion():
    x = torch.tensor([0.0, 0.5, 0.999, 1.0])
    assert gdist.to_uint8(x).tolist() == [0, 127, 254, 255]      # (x * 255).int(), inference/genefacepp_infer.py:469


def _worker_og(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 6                                                    # frames per rank (weak scaling), chunks of 4 + 2
    local = (torch.arange(T).view(T, 1, 1) + 10 * rank).expand(T, 5, 3).to(torch.uint8).contiguous()
    og = gdist.OverlappedGather(T, (5, 3), torch.uint8, "cpu")
    for a, b in ((0, 4), (4, 6)):
        og.push(local, a, b)                                 # chunk-wise, asynchronous
    full = og.finish()
    q.put((rank, tuple(full.shape), full[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_overlapped_chunked_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_og, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shape, vals in res:
        assert shape == (12, 5, 3)
        assert vals == [0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 14, 15], (rank, vals)   # rank-major = video order, on every rank
