# round 2, GPU session 8: same-box A/B of library variants (tools/r02/build_variants.sh)
#   A = ea_gemm.cu of the r02d build (6.87 ms)   B = HEAD (ping-pong drain, 168 registers, spills)
#   C = setmaxnreg warp-group layout (232-register epilogue)   D = C + exp2 polynomial in attention   E = C at 80 / 216
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_D.so timeout 900 python -m pytest tests/test_gpu_gemm_persistent.py tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s8_pytest_D.log 2>&1
tail -4 gpurun_out/s8_pytest_D.log
for rep in 1 2; do
for v in A B C D E; do
  EA_LIB_PATH=$AB/libea_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s8_bench_${v}_$rep.json 2> gpurun_out/s8_bench_${v}_$rep.err || tail -3 gpurun_out/s8_bench_${v}_$rep.err
  python - <<PY
import json
d=json.load(open('gpurun_out/s8_bench_${v}_$rep.json'))
print('VARIANT ${v} rep $rep ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
PY
done
done
