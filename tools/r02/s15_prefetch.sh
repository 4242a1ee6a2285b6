# round 2, GPU session 15: L2 weight look-ahead prefetch (EA_WEIGHT_PREFETCH = 0 off / 1 next launch / 2 two ahead), same library
set -x
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s15_pytest.log 2>&1
tail -3 gpurun_out/s15_pytest.log
n=0
for v in 0 1 2 0 1 3; do
  n=$((n+1))
  EA_WEIGHT_PREFETCH=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s15_bench_${n}_pf$v.json 2> gpurun_out/s15_bench_${n}_pf$v.err || tail -3 gpurun_out/s15_bench_${n}_pf$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s15_bench_${n}_pf$v.json'))
    print('PREFETCH $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('PREFETCH $v run $n FAILED', e)
PY
done
