# round 2, GPU session 18: L2 eviction hint on the W operand's TMA loads (EA_GEMM_W_EVICT = 0 none / 1 evict-first / 2 evict-last)
set -x
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py -m gpu -q -x ) > gpurun_out/s18_pytest.log 2>&1
tail -3 gpurun_out/s18_pytest.log
n=0
for v in 0 1 2 0 1; do
  n=$((n+1))
  EA_GEMM_W_EVICT=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s18_bench_${n}.json 2> gpurun_out/s18_bench_${n}.err || tail -3 gpurun_out/s18_bench_${n}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s18_bench_${n}.json'))
    print('W_EVICT $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('W_EVICT $v run $n FAILED', e)
PY
done
