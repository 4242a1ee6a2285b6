# round 2, GPU session 23 (2 GPUs): the driver's multi-GPU launch line at N = 2, 1 and 4 images per GPU (+ the reference arm's N > 1 contract)
set -x
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
for ipg in 1 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 --images-per-gpu $ipg > gpurun_out/s23_bench_2gpu_ipg$ipg.json 2> gpurun_out/s23_bench_2gpu_ipg$ipg.err; tail -3 gpurun_out/s23_bench_2gpu_ipg$ipg.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/s23_bench_2gpu_ipg$ipg.json') if l.startswith('{')][-1])
print('N=2 ipg $ipg value', d['value'], 'ms_step', d['ms_per_step'], 'e2e', d['e2e'], 'n_gpus', d['n_gpus'], d['config'].get('parallelism'))
PY
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-batch4 > gpurun_out/s23_bench_1gpu.json 2> gpurun_out/s23_bench_1gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s23_bench_1gpu.json') if l.startswith('{')][-1])
print('N=1 value', d['value'], 'ms_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'vae', d['config']['vae_decode_ms_per_image'], 'sam', d['config']['sam_ms_per_image'])
PY
