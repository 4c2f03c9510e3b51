#!/usr/bin/env bash
# tools/gpu_pending_checks.sh -- second (confirming) B200 run of the round-2 additions after the first run's fixes: the un-gated
# new test files, the SR tests now on the native kernels by default, ncu captures of the SR / torso-SR kernels, the SR timing.
# Every step has its own timeout and log under gpurun_out/pending2/.
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/gpu_pending_checks.sh'
set -u
OUT=gpurun_out/pending2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
run() {  # name, timeout, command...
    local name=$1 to=$2; shift 2
    ( timeout "$to" "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" >> "$OUT/$name.log" ) &
}
T0=$SECONDS
run train_ops   150 python -m pytest tests/test_gpu_train_ops.py -q -rA -s -p no:cacheprovider
run sr_native   150 python -m pytest tests/test_gpu_sr_native.py -q -rA -s -p no:cacheprovider
run sr_models   150 python -m pytest tests/test_gpu_sr.py tests/test_gpu_backend_shims.py -q -rA -s -p no:cacheprovider
wait
echo "tests done at $((SECONDS - T0)) s" > "$OUT/timing.txt"
run ncu_full    100 ncu --set full --clock-control none --import-source on -k "regex:k_sr_conv|k_torso_sr" --launch-skip 5 --launch-count 5 -f -o "$OUT/sr_kernels" python tools/ncu_sr_target.py
wait
echo "ncu full done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
run ncu_list    100 ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 400 --csv --log-file "$OUT/launches_torso_sr_clip.csv" python tools/ncu_sr_target.py --clip
wait
echo "ncu list done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
run sr_bench    100 python tools/sr_bench.py --frames 64 --reps 5 --out "$OUT/sr_bench.jsonl"
wait
echo "all done at $((SECONDS - T0)) s" >> "$OUT/timing.txt"
tail -n 3 "$OUT"/*.log
