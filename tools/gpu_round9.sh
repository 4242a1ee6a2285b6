set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_r01i timeout 600 python tools/bench_ops.py gemm norm > gpurun_out/bench_ops_r01i.log 2>&1; cat gpurun_out/bench_ops_r01i.log | tail -34
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-1700; tail -3 gpurun_out/bench.err
