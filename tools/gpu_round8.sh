set -x
cd /root/repo
mkdir -p gpurun_out
timeout 100 python tools/exp_attn_timing.py 2>&1 | tail -30 | grep -E "kv8|kv9|kv10" 
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_r01g timeout 600 python tools/bench_ops.py attn gemm norm > gpurun_out/bench_ops_r01g.log 2>&1; cat gpurun_out/bench_ops_r01g.log | tail -45
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
