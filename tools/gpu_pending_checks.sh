#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out and still needs its first run on a B200.
# One gpurun call (~4-6 GPU-minutes), every step under its own timeout, logs under gpurun_out/pending/:
#
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_pending_checks.sh'
#
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/pending
mkdir -p "$out"

echo "== 1. MLP-layer chain probe (1-3 batch slots in flight; docs/HEAD_V2_PLAN.md step 1)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I genefaceplusplus_b200/csrc -o tools/microbench/mma_chain.bin tools/microbench/mma_chain.cu \
    > "$out/mma_chain_build.log" 2>&1
timeout 40 tools/microbench/mma_chain.bin > "$out/mma_chain.txt" 2>&1; echo "   rc=$?"; tail -8 "$out/mma_chain.txt"

echo "== 2. SR drop-ins against the reference goldens (tests/test_gpu_sr_pending.py)"
GFPP_RUN_PENDING=1 timeout 200 python -m pytest tests/test_gpu_sr_pending.py -m gpu -q -s > "$out/sr_pending.log" 2>&1; echo "   rc=$?"; tail -12 "$out/sr_pending.log"

echo "== 3. the reference's own CUDA kernels under the restated host loop (SURVEY 8(d)(ii))"
timeout 240 python tools/ref_gpu_baseline.py --size 512 --frames 8 > "$out/ref_gpu_baseline.json" 2> "$out/ref_gpu_baseline.err"; echo "   rc=$?"
tail -3 "$out/ref_gpu_baseline.json"; tail -3 "$out/ref_gpu_baseline.err"

echo "== 4. regression: full GPU suite + bench"
timeout 300 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1; echo "   rc=$?"; tail -3 "$out/pytest_gpu.log"
timeout 200 python bench.py --no-cpu-baseline > "$out/bench_fp16.json" 2> "$out/bench_fp16.err"; echo "   rc=$?"; cut -c1-200 "$out/bench_fp16.json"
