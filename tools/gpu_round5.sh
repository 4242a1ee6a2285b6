set -x
cd /root/repo
mkdir -p gpurun_out
timeout 100 python tools/exp_attn_timing.py 2>&1 | tail -30
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_r01e timeout 400 python tools/bench_ops.py attn gemm > gpurun_out/bench_ops_r01e.log 2>&1; cat gpurun_out/bench_ops_r01e.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
