# round 2, GPU session 3: lockstep (grouped) encoder + grouped GEMM + bench with the batch-4 block
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s3_pytest_gpu.log 2>&1
tail -12 gpurun_out/s3_pytest_gpu.log
grep -h "max-abs" gpurun_out/s3_pytest_gpu.log | head -12
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; cut -c1-300 gpurun_out/s3_bench.json; tail -5 gpurun_out/s3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s3_bench.json'))
print('ms_step', d['ms_per_step'], 'launches', d['launches_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['launches'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'])
print('batch4', d['config'].get('batch4'))
PY
EA_LOCKSTEP=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e --no-batch4 > gpurun_out/s3_bench_nolockstep.json 2>/dev/null; cut -c150-330 gpurun_out/s3_bench_nolockstep.json
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --images-per-gpu 4 > gpurun_out/s3_bench_ipg4.json 2> gpurun_out/s3_bench_ipg4.err; cut -c1-330 gpurun_out/s3_bench_ipg4.json; tail -3 gpurun_out/s3_bench_ipg4.err
