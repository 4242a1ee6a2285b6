# round 2, GPU session 28 (2 GPUs): the driver's default multi-GPU command line (batch4 block + all-reduce included)
cd /root/repo
mkdir -p gpurun_out
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s28_bench_2gpu_default.json 2> gpurun_out/s28_bench_2gpu_default.err
tail -3 gpurun_out/s28_bench_2gpu_default.err
wc -l gpurun_out/s28_bench_2gpu_default.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s28_bench_2gpu_default.json").read().strip().splitlines()[-1])
print("N=2 default value", d["value"], "ms_step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["pass_ms"], "batch4", d["config"]["batch4"]["ms_per_step"], d["config"]["batch4"]["images_per_s_denoise_only"], "cpu_baseline" in d)
PY
