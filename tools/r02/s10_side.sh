# round 2, GPU session 10: side-data warps (S) vs the kept build (H) on the same box, then the per-shape sweep of the small-K linears
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_S.so timeout 400 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py -m gpu -q -x ) > gpurun_out/s10_pytest_S.log 2>&1
ok=$?
tail -3 gpurun_out/s10_pytest_S.log
if [ $ok -eq 0 ]; then VARS="H S H S"; SW=S; else VARS="H"; SW=H; fi
n=0
for v in $VARS; do
  n=$((n+1))
  EA_LIB_PATH=$AB/libea_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s10_bench_${n}_$v.json 2> gpurun_out/s10_bench_${n}_$v.err || tail -3 gpurun_out/s10_bench_${n}_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s10_bench_${n}_$v.json'))
    print('VARIANT ${v} run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('VARIANT ${v} run $n FAILED', e)
PY
done
EA_LIB_PATH=$AB/libea_$SW.so timeout 600 python tools/r02/gemm_shape_sweep.py gpurun_out/s10_gemm_shape_sweep.json > gpurun_out/s10_gemm_shape_sweep.txt 2>&1
tail -5 gpurun_out/s10_gemm_shape_sweep.txt
