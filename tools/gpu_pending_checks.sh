#!/usr/bin/env bash
# tools/gpu_pending_checks.sh -- third (last) B200 run of round 2: the SR-head test with its rounding-aware workspace bounds, and a
# regression of the core per-op / render tests against the rebuilt library.  Logs under gpurun_out/pending3/.
#   /usr/local/graft/bin/gpurun --timeout 170 -- 'bash tools/gpu_pending_checks.sh'
set -u
OUT=gpurun_out/pending3
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; ( timeout "$to" "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" >> "$OUT/$name.log" ) & }
run sr_head  120 python -m pytest tests/test_gpu_sr_native.py -q -rA -s -p no:cacheprovider -k "native_sr_head or repacks"
run core     150 python -m pytest tests/test_gpu_ops.py tests/test_gpu_render.py tests/test_gpu_layouts.py -q -rA -p no:cacheprovider
run core2    150 python -m pytest tests/test_gpu_render_tc.py tests/test_gpu_tc.py tests/test_gpu_ref_pin.py -q -rA -p no:cacheprovider
wait
tail -n 4 "$OUT"/*.log
