#!/usr/bin/env python
"""bench.py -- frames/sec of the GeneFace++ render hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's sm_100a kernels
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's path on the host CPU cores

One "step" = one pass of the hot path over one batch of synthetic input = rendering this rank's shard of a
512x512 head+torso clip (BASELINE config "May head+torso two-pass 512x512, 250 frames, 1xB200"; at N GPUs every
rank renders its own 250 frames -- config 4, 2000 frames over 8 GPUs -- and one NCCL all-gather of the RGB
follows: weak scaling).  Random-init May-shaped weights, synthetic poses/conditioning (no checkpoint or dataset
ships with the reference): "data": "synthetic".

Timed numbers:
  value      frames/s with inputs resident in HBM (poses, conditioning sequence, background), CUDA events,
             barrier + synchronize on both sides, max over ranks
  e2e        the same clip through the public clip API starting from pinned HOST buffers: H2D of poses +
             conditioning inside the timed region, D2H of the rendered uint8 frames
  roofline   dominant kernel (k_head, pass 1) timed live with CUDA events inside libgfpp on the launching
             stream: algorithmic bytes / duration vs the measured HBM peak
  cpu_baseline  the CPU oracle port timed on rank 0's host cores on a bounded sample of the same workload
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=250, help="frames per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--density-scale", type=float, default=8.0)
    ap.add_argument("--head-only", action="store_true")
    ap.add_argument("--frames-per-call", type=int, default=50)
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the bounded CPU-baseline sample (~3.5 s each on the box)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=os.environ.get("GFPP_BENCH_PRECISION", "fp16"), choices=["fp32", "fp16", "bf16x3", "bf16"],
                    help="arithmetic of the head MLP GEMMs (marching/gather/compositing are fp32 in every mode)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d
    return 6650.0, "fallback", {}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if "Active" in r[5 + k] and "Not" not in r[5 + k]:
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


_CPU_THREADS = None


def cpu_reference_fps(args, n_frames):
    """The reference's path on the host cores: the reference's PyTorch-eager modules restated in oracle/render.py
    over the C restatement of its CUDA-only native ops (kind "port": the reference has no CPU implementation of
    those ops and its Python cannot travel to the GPU box).  All host threads."""
    import torch
    from genefaceplusplus_b200 import scene as scn
    from oracle import ops
    from oracle.render import OracleModel
    ops.build()
    cores = os.cpu_count() or 1
    sc = scn.Scene(H=args.size, W=args.size, T=max(n_frames, 8), torso=not args.head_only, density_scale=args.density_scale)
    orc = OracleModel(sc.state, sc.hparams)
    orc.density_scale = sc.density_scale
    # "all the host threads it can use": more threads than the op sizes can feed makes eager PyTorch slower, so the
    # thread count is calibrated on a small frame and the fastest setting is used (and reported)
    global _CPU_THREADS
    if _CPU_THREADS is None:
        cal = scn.Scene(H=96, W=96, T=8, torso=not args.head_only, density_scale=args.density_scale)
        co = OracleModel(cal.state, cal.hparams); co.density_scale = cal.density_scale
        fi = cal.frame_inputs(0)
        best = (1e30, cores)
        for nt in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
            torch.set_num_threads(nt); ops.set_num_threads(nt)
            co.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"], T_thresh=cal.T_thresh, **cal.hparams)
            t0 = time.time()
            co.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"], T_thresh=cal.T_thresh, **cal.hparams)
            best = min(best, (time.time() - t0, nt))
        _CPU_THREADS = best[1]
    cores = _CPU_THREADS
    torch.set_num_threads(cores); ops.set_num_threads(cores)
    t0 = time.time()
    S = 0
    for t in range(n_frames):
        fi = sc.frame_inputs(t)
        out = orc.render(fi["rays_o"], fi["rays_d"], fi["cond"], fi["bg_coords"], fi["poses"], bg_color=fi["bg_color"],
                         T_thresh=sc.T_thresh, **sc.hparams)
        S += out["stats"]["S"]
    dt = time.time() - t0
    return n_frames / dt, cores, dt, S


def workload_name(args):
    return (f"May {'head-NeRF' if args.head_only else 'head+torso two-pass'} {args.size}x{args.size}, {args.frames}-frame driving clip per GPU, "
            f"max_steps=16, T_thresh=0.01, density_scale={args.density_scale:g}")


def run_reference(args, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps_list = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_fps(args, 1)
    cores = os.cpu_count() or 1
    t_total = 0.0
    for _ in range(args.steps):
        fps, cores, dt, _ = cpu_reference_fps(args, args.cpu_frames)
        fps_list.append(fps); t_total += dt
    v = args.cpu_frames * args.steps / t_total
    line = {"impl": "reference", "metric": "frames/sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": f"{args.cpu_frames} frame(s) of the clip per step"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{args.cpu_frames} frame(s) at {args.size}x{args.size} per step, {args.steps} steps"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out)


def _claim_stdout():
    """Keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner to stdout) are redirected to stderr."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    args = parse()
    out = _claim_stdout()
    try:
        return _main(args, out)
    finally:
        out.flush()


def _main(args, out):
    if args.impl == "reference":
        return run_reference(args, out)
    import torch
    import torch.distributed as dist
    from genefaceplusplus_b200 import _capi, scene as scn
    from genefaceplusplus_b200 import dist as gdist
    from genefaceplusplus_b200.renderer import RADNeRF, RADNeRFTorso

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()
    _capi.check(L.gfpp_check_device(), "gfpp_check_device")

    H = W = args.size
    T = args.frames
    sc = scn.Scene(H=H, W=W, T=T * world, torso=not args.head_only, density_scale=args.density_scale)
    model = (RADNeRF if args.head_only else RADNeRFTorso)(sc.hparams)
    model.load_state_dict(sc.state, strict=True)
    model.density_scale = sc.density_scale
    model.mlp_precision = args.precision
    model = model.to(dev).eval()
    s, e = gdist.frame_block(T * world, rank, world)
    poses_host = torch.stack([sc.pose(t) for t in range(s, e)]).pin_memory()
    cond_host = sc.cond.clone().pin_memory()            # full sequence on every rank (window halo)
    bg_color = sc.bg_color.to(dev)
    bg_coords = sc.bg_coords.to(dev)
    poses_dev = poses_host.to(dev)
    cond_dev = cond_host.to(dev)
    N = H * W
    rgb = torch.empty(T, N, 3, device=dev, dtype=torch.float32)
    u8_host = torch.empty(T, N, 3, dtype=torch.uint8).pin_memory()
    launches = 0

    def step_device():
        nonlocal launches
        feat = model.cal_cond_feat_clip(cond_dev)[s:e]
        n0 = 0
        for a in range(0, T, args.frames_per_call):
            b = min(T, a + args.frames_per_call)
            res = model.render_frames(feat[a:b], poses_c2w=poses_dev[a:b], intrinsics=sc.intrinsics, H=H, W=W,
                                      pose6=pose6_dev[a:b] if not args.head_only else None, bg_coords=bg_coords, bg_color=bg_color,
                                      dt_gamma=sc.hparams["dt_gamma"], max_steps=sc.hparams["max_steps"], T_thresh=sc.T_thresh,
                                      want_torso_maps=False, want_stats=True)
            rgb[a:b].copy_(res["rgb_map"])
            stats_acc.append(res["stats"])
            n0 += model.last_launch_count
        launches += n0
        if world > 1:
            return gdist.gather_frames(gdist.to_uint8(rgb), T * world)
        return rgb

    def step_e2e():
        p = poses_host.to(dev, non_blocking=True)
        c = cond_host.to(dev, non_blocking=True)
        out = model.render_clip(p, sc.intrinsics, H, W, cond_seq=c, bg_color=bg_color, bg_coords=bg_coords, pose6=pose6_dev,
                                T_thresh=sc.T_thresh, frames_per_call=args.frames_per_call, out=rgb)
        cf = None
        u8_host.copy_(gdist.to_uint8(out), non_blocking=True)
        return cf

    pose6_dev = scn.convert_poses(poses_host).to(dev) if not args.head_only else None
    stats_acc = []

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    sync_all()
    stats_acc.clear(); launches = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else {}
    ms = ev0.elapsed_time(ev1)
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = tms.item()
    timed_launches = launches
    st = torch.cat(stats_acc, 0).cpu()   # [steps*T, 4]: B_total, n_survivors, S, P
    S_per_frame = st[:, 2].float().mean().item()
    P_per_frame = st[:, 3].float().mean().item()
    fps = world * T * args.steps / (ms / 1000.0)

    # ---------------- dominant-kernel roofline (live, CUDA events inside libgfpp) ----------------
    L.gfpp_profile_enable(1)
    head_ms, pass2_ms, epi_ms, pre_ms = [], [], [], []
    feat = model.cal_cond_feat_clip(cond_dev)[s:e]
    Fc = min(args.frames_per_call, T)
    for it in range(5):
        model.render_frames(feat[:Fc], poses_c2w=poses_dev[:Fc], intrinsics=sc.intrinsics, H=H, W=W,
                            pose6=pose6_dev[:Fc] if not args.head_only else None, bg_coords=bg_coords, bg_color=bg_color,
                            dt_gamma=sc.hparams["dt_gamma"], max_steps=sc.hparams["max_steps"], T_thresh=sc.T_thresh, want_torso_maps=False)
        buf = (ctypes.c_float * 4)()
        _capi.check(L.gfpp_profile_read(buf), "profile_read")
        if it >= 2:
            head_ms.append(buf[0]); pass2_ms.append(buf[1]); epi_ms.append(buf[2]); pre_ms.append(buf[3])
    L.gfpp_profile_enable(0)
    hbm, peak_kind, _ = peaks()
    head_t = statistics.mean(head_ms) / 1000.0
    # 2 grids x 16 levels x 8 corners x 8 B per valid sample; per ray 12 B colour + 4 B alpha + 4 B depth out
    # (+24 B when rays are supplied: here they are generated in-kernel); + the packed weights once
    alg_bytes = Fc * (S_per_frame * 2048 + N * (12 + 4 + 4)) + 0.36e6
    alg_flops = Fc * S_per_frame * 178944
    achieved = alg_bytes / head_t / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp) and args.precision == "fp16":
        traffic = json.load(open(tp))["dram_bytes_per_frame"] * Fc     # from one ncu --set full capture, scaled to this launch size
    roofline = {"bound": "hbm", "kernel": "k_head (pass 1)" if args.precision == "fp32" else "k_head_tc (pass 1)", "achieved": achieved, "peak": hbm, "peak_source": peak_kind, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "launch_ms": head_t * 1000.0, "frames_per_launch": Fc,
                "share_of_step": head_t / (head_t + (statistics.mean(pass2_ms) + statistics.mean(epi_ms) + statistics.mean(pre_ms)) / 1000),
                "ray_setup_ms": statistics.mean(pre_ms),
                "fp32_tflops": alg_flops / head_t / 1e12, "pass2_ms": statistics.mean(pass2_ms), "epilogue_ms": statistics.mean(epi_ms),
                "alg_bytes_per_frame": alg_bytes / Fc, "alg_flops_per_frame": alg_flops / Fc}

    # ---------------- end to end (host buffers in, uint8 frames out) ----------------
    for _ in range(2):
        step_e2e()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_fps = world * T * args.steps / (ems.item() / 1000.0)
    h2d = poses_host.numel() * 4 + cond_host.numel() * 4
    d2h = u8_host.numel()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, cores, dt, _ = cpu_reference_fps(args, args.cpu_frames)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_frames} frame(s) of the same clip at {H}x{W} ({dt:.1f} s)"}

    if rank == 0:
        line = {"metric": "frames/sec at 512x512 head+torso" if not args.head_only else "frames/sec at 512x512 head", "value": fps,
                "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "fp16": "f16 operands, f32 accumulate (tcgen05)", "bf16x3": "bf16 hi/lo split x3, f32 accumulate (tcgen05)", "bf16": "bf16 operands, f32 accumulate (tcgen05)"}[args.precision],
                "data": "synthetic",
                "config": {"workload": workload_name(args), "mlp_precision": args.precision, "frames_per_gpu_per_step": T, "frames_per_call": args.frames_per_call,
                           "S_valid_samples_per_frame": S_per_frame, "P_torso_pixels_per_frame": P_per_frame, "B_total": int(st[0, 0]),
                           "parallelism": f"frame-sharded x{world}, 1 all-gather of uint8 RGB" if world > 1 else "single GPU",
                           "l2": "per-step working set (786 MB fp32 frames out + 1.8 GB workspace) >> 126 MB L2; grid tables (14.4 MB) are L2-resident by design"},
                "clocks": clocks, "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                          "note": "pinned host poses+conditioning in, uint8 [T,H,W,3] frames out"},
                "gpu_launches": timed_launches, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), file=out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
