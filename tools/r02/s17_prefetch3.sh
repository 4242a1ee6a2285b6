# round 2, GPU session 17: L2 weight look-ahead through prefetch.global.L2 (load-store path), weight-bound targets or all
set -x
cd /root/repo
mkdir -p gpurun_out
n=0
for v in "0 512" "1 512" "1 100000" "0 512" "1 512" "1 100000"; do
  set -- $v
  n=$((n+1))
  EA_WEIGHT_PREFETCH=$1 EA_WEIGHT_PREFETCH_MAXM=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --no-sam --no-e2e > gpurun_out/s17_bench_${n}.json 2> gpurun_out/s17_bench_${n}.err || tail -3 gpurun_out/s17_bench_${n}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s17_bench_${n}.json'))
    print('PREFETCH $v run $n ms_step', d['ms_per_step'], 'gemm_ms', d['roofline']['gemm_ms_per_step'])
except Exception as e:
    print('PREFETCH $v run $n FAILED', e)
PY
done
