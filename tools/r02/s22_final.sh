# round 2, GPU session 22: final build - full GPU suite, full bench (default and without the W evict-first hint for the SAM / VAE legs), smoke()
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s22_pytest_gpu.log 2>&1
tail -4 gpurun_out/s22_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for v in 1 0; do
EA_GEMM_W_EVICT=$v timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s22_bench_evict$v.json 2> gpurun_out/s22_bench_evict$v.err; tail -2 gpurun_out/s22_bench_evict$v.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s22_bench_evict$v.json') if l.startswith('{')][-1])
print('W_EVICT $v ms_step', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'], 'value', d['value'], 'sam', round(d['config']['sam_ms_per_image'],3), 'vae dec/enc', round(d['config']['vae_decode_ms_per_image'],3), round(d['config']['vae_encode_ms_per_image'],3), 'batch4', d['config']['batch4']['ms_per_step'], d['config']['batch4']['frac_of_sustained_peak'])
PY
done
