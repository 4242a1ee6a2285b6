# round 2, GPU session 24: e2e stability (3 runs of the default bench without the CPU leg), final library
set -x
cd /root/repo
mkdir -p gpurun_out
for n in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s24_bench_$n.json 2> gpurun_out/s24_bench_$n.err; tail -2 gpurun_out/s24_bench_$n.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s24_bench_$n.json') if l.startswith('{')][-1])
print('RUN $n ms_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'value', d['value'], 'image_ms', d['config']['image_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], d['e2e']['pass_ms'], 'vae', round(d['config']['vae_decode_ms_per_image'],3), 'batch4', d['config']['batch4']['ms_per_step'])
PY
done
