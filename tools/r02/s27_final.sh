# round 2, GPU session 27: final build - full GPU suite + default bench
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/s27_pytest_gpu.log 2>&1
tail -4 gpurun_out/s27_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err; tail -2 gpurun_out/s27_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s27_bench.json') if l.startswith('{')][-1])
print('FINAL ms_step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], 'value', d['value'], 'image_ms', d['config']['image_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], d['e2e']['pass_ms'], 'batch4', d['config']['batch4']['ms_per_step'], d['config']['batch4']['frac_of_sustained_peak'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
PY
wc -l gpurun_out/s27_bench.json
