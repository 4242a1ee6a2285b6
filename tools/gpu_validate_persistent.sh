# One gpurun call that validates the experimental persistent GEMM kernel and measures its effect:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_validate_persistent.sh'
# 1. bitwise parity against the one-tile-per-CTA kernel (gated tests), 2. per-shape GEMM times of a real step
# with and without EA_GEMM_PERSIST=1, 3. the full GPU suite and the bench line with the variant enabled.
set -x
cd /root/repo
mkdir -p gpurun_out
( time EA_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py -m gpu -q -x ) > gpurun_out/pytest_persistent.log 2>&1
tail -8 gpurun_out/pytest_persistent.log
if grep -q "failed\|error" gpurun_out/pytest_persistent.log; then echo "PERSISTENT KERNEL NOT VALID - stopping"; exit 0; fi
timeout 500 python tools/gemm_breakdown.py gpurun_out/gemm_breakdown_base.json 2>&1 | head -12
EA_GEMM_PERSIST=1 timeout 500 python tools/gemm_breakdown.py gpurun_out/gemm_breakdown_persist.json 2>&1 | head -12
EA_GEMM_PERSIST=2 timeout 500 python tools/gemm_breakdown.py gpurun_out/gemm_breakdown_persist2.json 2>&1 | head -12
( time EA_GEMM_PERSIST=1 timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_persist.log 2>&1
tail -4 gpurun_out/pytest_gpu_persist.log
EA_GEMM_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-700
EA_GEMM_PERSIST=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-700
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-400
