# round 2, GPU session 5: LN-stats load fix, conv_in rewrite, guess mode / scale maps, fused UniPC, app on CUDA
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/s5_pytest_gpu.log 2>&1
tail -14 gpurun_out/s5_pytest_gpu.log
grep -h "max-abs" gpurun_out/s5_pytest_gpu.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; tail -3 gpurun_out/s5_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5_bench.json'))
print('ms_step', d['ms_per_step'], 'launches', d['launches_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['launches'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'], 'value', d['value'])
print('batch4', d['config'].get('batch4'))
PY
timeout 500 python tools/gemm_breakdown.py gpurun_out/s5_gemm_breakdown.json 2>&1 | head -30
