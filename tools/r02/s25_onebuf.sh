# round 2, GPU session 25: persistent epilogue with ONE TMEM register buffer reloaded as soon as it is consumed (V) vs two buffers + copy (W)
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_V.so timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s25_pytest_V.log 2>&1
tail -3 gpurun_out/s25_pytest_V.log
n=0
for v in W V W V; do
  n=$((n+1))
  EA_LIB_PATH=$AB/libea_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s25_bench_${n}_$v.json 2> gpurun_out/s25_bench_${n}_$v.err || tail -3 gpurun_out/s25_bench_${n}_$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/s25_bench_${n}_$v.json') if l.startswith('{')][-1])
    print('VARIANT $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('VARIANT $v run $n FAILED', e)
PY
done
