# round 2, GPU session 2: LayerNorm fold + default persistent kernel + device step tables + pipeline on CUDA
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s2_pytest_gpu.log 2>&1
tail -15 gpurun_out/s2_pytest_gpu.log
grep -h "max-abs" gpurun_out/s2_pytest_gpu.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; cut -c1-300 gpurun_out/s2_bench.json; tail -3 gpurun_out/s2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s2_bench.json'))
print('ms_step', d['ms_per_step'], 'launches', d['launches_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'])
PY
EA_LN_FOLD=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s2_bench_nofold.json 2>/dev/null; cut -c150-330 gpurun_out/s2_bench_nofold.json
EA_GEMM_PERSIST=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s2_bench_nopersist.json 2>/dev/null; cut -c150-330 gpurun_out/s2_bench_nopersist.json
EA_GEMM_PERSIST=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s2_bench_persist2all.json 2>/dev/null; cut -c150-330 gpurun_out/s2_bench_persist2all.json
