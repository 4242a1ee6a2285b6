# round 2, GPU session 26: single TMEM register buffer also in the one-tile kernel (X) vs persistent kernel only (V)
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_X.so timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s26_pytest_V.log 2>&1
tail -3 gpurun_out/s26_pytest_V.log
n=0
for v in V X V X; do
  n=$((n+1))
  EA_LIB_PATH=$AB/libea_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s26_bench_${n}_$v.json 2> gpurun_out/s26_bench_${n}_$v.err || tail -3 gpurun_out/s26_bench_${n}_$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s26_bench_${n}_$v.json') if l.startswith('{')][-1])
    print('VARIANT $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('VARIANT $v run $n FAILED', e)
PY
done
