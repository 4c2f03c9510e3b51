#!/usr/bin/env bash
# tools/gpu_pending_checks.sh -- ONE gpurun call that runs, for the first time on a B200, everything written without a GPU:
# training-side ops, the torso-SR field kernel, the tcgen05 SR head, the all-native torso-SR clip path, and the SR timing.
# Every step has its own timeout and log under gpurun_out/pending/; a failure or hang of one step never blocks the others,
# and the kernels most likely to hang (new tcgen05 pipelines) run after the plain-CUDA ones.
#   /usr/local/graft/bin/gpurun --timeout 480 -- 'bash tools/gpu_pending_checks.sh'
set -u
OUT=gpurun_out/pending
mkdir -p "$OUT"
export GFPP_PENDING=1 PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
run() {  # name, timeout, command...
    local name=$1 to=$2; shift 2
    ( timeout "$to" "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" >> "$OUT/$name.log" ) &
}
T0=$SECONDS
# wave 1: plain-CUDA kernels + the regression of what the shims / build changes could have touched
run train_ops   200 python -m pytest tests/test_gpu_train_ops.py -q -rA -s -p no:cacheprovider
run torso_field 200 python -m pytest tests/test_gpu_sr_native.py -q -rA -s -p no:cacheprovider -k "torso_sr_field"
run regression  200 python -m pytest tests/test_gpu_backend_shims.py tests/test_gpu_sr.py tests/test_gpu_v2.py -q -rA -p no:cacheprovider
run smoke       200 python -c "import __graft_entry__ as g; g.smoke()"
wait
echo "wave 1 done at $((SECONDS - T0)) s" > "$OUT/timing.txt"
# wave 2: the tcgen05 SR head
run sr_head     200 python -m pytest tests/test_gpu_sr_native.py -q -rA -s -p no:cacheprovider -k "native_sr_head or repacks"
wait
echo "wave 2 done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
run sr_clip     150 python -m pytest tests/test_gpu_sr_native.py -q -rA -s -p no:cacheprovider -k "clip_all_native"
wait
echo "wave 3 done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
run sr_bench    120 python tools/sr_bench.py --frames 32 --reps 3 --out "$OUT/sr_bench.jsonl"
wait
echo "all done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
tail -n 4 "$OUT"/*.log
