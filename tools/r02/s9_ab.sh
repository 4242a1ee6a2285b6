# round 2, GPU session 9: same-box A/B of the epilogue pieces (tools/r02/build_variants.sh; attention exp MUFU-only except K)
#   A = r02d ea_gemm.cu   F = all new + setmaxnreg 40/232   G = F, rolled flush   H = F, no ping-pong   I = F, rolled + no ping-pong
#   J = I on the 320-thread layout (~A + float4 GEGLU loads + flush32)   K = F + exp2 polynomial in attention
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_F.so timeout 240 python -m pytest tests/test_gpu_gemm_persistent.py -m gpu -q -x ) > gpurun_out/s9_pytest_F.log 2>&1
ok=$?
tail -3 gpurun_out/s9_pytest_F.log
if [ $ok -eq 0 ]; then VARS="A F G H I J K A F"; else VARS="A J A J"; fi
n=0
for v in $VARS; do
  n=$((n+1))
  EA_LIB_PATH=$AB/libea_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s9_bench_${n}_$v.json 2> gpurun_out/s9_bench_${n}_$v.err || tail -3 gpurun_out/s9_bench_${n}_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s9_bench_${n}_$v.json'))
    print('VARIANT ${v} run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('VARIANT ${v} run $n FAILED', e)
PY
done
if [ $ok -eq 0 ]; then
( EA_LIB_PATH=$AB/libea_K.so timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/s9_pytest_K.log 2>&1
tail -3 gpurun_out/s9_pytest_K.log
fi
