# round 2, GPU session 14: GroupNorm with cp.async pixel staging + early affine loads (P) vs the old kernel (Q); both with the N = 320 tile pin
set -x
cd /root/repo
mkdir -p gpurun_out
AB=/root/repo/editanything_b200/lib/ab
( EA_LIB_PATH=$AB/libea_P.so timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_vae.py -m gpu -q -x ) > gpurun_out/s14_pytest_P.log 2>&1
tail -3 gpurun_out/s14_pytest_P.log
n=0
for v in Q P Q P; do
  n=$((n+1))
  EA_LIB_PATH=$AB/libea_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s14_bench_${n}_$v.json 2> gpurun_out/s14_bench_${n}_$v.err || tail -3 gpurun_out/s14_bench_${n}_$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s14_bench_${n}_$v.json'))
    print('VARIANT $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'], 'frac', d['roofline']['frac'], 'vae', d['config'].get('vae_decode_ms_per_image'))
except Exception as e:
    print('VARIANT $v run $n FAILED', e)
PY
done
