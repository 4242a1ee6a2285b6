set -x
cd /root/repo
mkdir -p gpurun_out
for c in attn_db_long attn_d40_self attn_d64; do timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-300; done
EA_BENCH_TAG=attn_r01m timeout 300 python tools/bench_ops.py attn 2>&1 | tail -11
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-2600
