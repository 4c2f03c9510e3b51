#!/usr/bin/env bash
# Round-2 evidence run on one B200 (one gpurun call): full GPU suite, A/B of the issue-order knob, ncu captures of both
# second-generation kernels, launch list, phase stamps, single-GPU roofline sweep.  Everything under its own timeout.
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/r02a; mkdir -p $o
echo "== 1. GPU suite (default)"; timeout 420 python -m pytest tests -m gpu -x -q > $o/pytest_default.log 2>&1; echo "rc=$?"; tail -3 $o/pytest_default.log
echo "== 2. v2 tests with GFPP_V2_RELAXED=1"; GFPP_V2_RELAXED=1 timeout 300 python -m pytest tests/test_gpu_v2.py tests/test_gpu_render_tc.py tests/test_gpu_render.py -m gpu -x -q > $o/pytest_relaxed.log 2>&1; echo "rc=$?"; tail -3 $o/pytest_relaxed.log
echo "== 3. A/B"; for r in 0 1 0 1; do GFPP_V2_RELAXED=$r timeout 100 python tools/quick_bench.py fp16 2>&1 | tail -1 | sed "s/^/relaxed=$r /"; done | tee $o/ab.txt
GFPP_V2_RELAXED=0 timeout 100 python tools/quick_bench.py robust 2>&1 | tail -1 | sed "s/^/relaxed=0 /" | tee -a $o/ab.txt
GFPP_V2_RELAXED=1 timeout 100 python tools/quick_bench.py robust 2>&1 | tail -1 | sed "s/^/relaxed=1 /" | tee -a $o/ab.txt
echo "== 4. ncu"; bash tools/ncu_head.sh headv2_r02_fp16 fp16 k_head_v2 | tail -1; bash tools/ncu_head.sh headv2_r02_robust robust k_head_v2 | tail -1
echo "== 5. launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_fp16.csv python bench.py --steps 1 --warmup 3 --frames 100 --no-cpu-baseline --no-gpu-reference --no-other-mode > $o/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== 6. phases"; timeout 100 python tools/phase_breakdown.py fp16 > $o/phase_v2_fp16.txt 2>&1; timeout 100 python tools/phase_breakdown.py robust > $o/phase_v2_robust.txt 2>&1; tail -21 $o/phase_v2_fp16.txt
echo "== 7. sweep (1 GPU)"; timeout 300 python tools/roofline_sweep.py fp16 > $o/sweep_n1_fp16.jsonl 2> $o/sweep_n1.err; echo "rc=$?"; wc -l $o/sweep_n1_fp16.jsonl
