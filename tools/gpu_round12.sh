set -x
cd /root/repo
mkdir -p gpurun_out
for c in attn_sam_window attn_sam_global sam_helpers attn_d40_self attn_cross77; do timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-300; done
timeout 600 python -m pytest tests/test_gpu_sam.py -m gpu -q -x 2>&1 | tail -5
EA_BENCH_TAG=attn_r01l timeout 300 python tools/bench_ops.py attn 2>&1 | tail -11
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-900
