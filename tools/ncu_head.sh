#!/usr/bin/env bash
# One `ncu --set full` capture (with source counters) of the pass-1 head kernel: 10 frames of 512x512 head+torso per launch.
#   gpurun --timeout 600 -- 'bash tools/ncu_head.sh <tag> [precision] [kernel-regex]'
# -> gpurun_out/prof_<tag>.ncu-rep ; read here with `ncu -i ... --page raw --csv` / `--page source --csv`.
set -u
cd "$(dirname "$0")/.."
tag=${1:-head}; prec=${2:-fp16}; kre=${3:-k_head}
mkdir -p gpurun_out
cat > /tmp/ncu_drive.py <<PY
import sys, os, torch
sys.path.insert(0, os.getcwd())
from genefaceplusplus_b200 import scene as scn
from genefaceplusplus_b200.renderer import RADNeRFTorso
sc = scn.Scene(H=512, W=512, T=10, torso=True, density_scale=8.0)
m = RADNeRFTorso(sc.hparams); m.load_state_dict(sc.state); m.density_scale = 8.0; m.mlp_precision = "$prec"; m = m.cuda().eval()
poses = torch.stack([sc.pose(t) for t in range(10)])
kw = dict(cond_seq=sc.cond, bg_color=sc.bg_color, bg_coords=sc.bg_coords, T_thresh=0.01, frames_per_call=10)
for _ in range(3):
    m.render_clip(poses, sc.intrinsics, 512, 512, **kw); torch.cuda.synchronize()
PY
# the pass-1 launch of the third clip: each clip launches the head kernel twice (pass 1, pass 2)
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$kre" --launch-skip 4 --launch-count 1 \
    -f -o gpurun_out/prof_$tag python /tmp/ncu_drive.py > gpurun_out/prof_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_$tag.log
