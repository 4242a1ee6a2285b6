# round 2, GPU session 13: row-major blocked tile order for LayerNorm-folded GEMMs + N = 320 tile pin (R) vs the kept build (H)
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_R.so timeout 600 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s13_pytest_R.log 2>&1
ok=$?
tail -3 gpurun_out/s13_pytest_R.log
EA_LIB_PATH=$AB/libea_R.so timeout 300 python tools/r02/gemm_ln_cost.py 2>&1 | tail -12
n=0
for v in H R "R EA_PL_N320_BN=96" H R; do
  set -- $v
  n=$((n+1))
  env $2 EA_LIB_PATH=$AB/libea_$1.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s13_bench_${n}_$1.json 2> gpurun_out/s13_bench_${n}_$1.err || tail -3 gpurun_out/s13_bench_${n}_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s13_bench_${n}_$1.json'))
    print('VARIANT $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('VARIANT $v run $n FAILED', e)
PY
done
