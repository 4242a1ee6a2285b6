set -x
cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/exp_gemm_timing.py 2>&1 | tail -70
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_r01f timeout 400 python tools/bench_ops.py gemm > gpurun_out/bench_ops_r01f.log 2>&1; cat gpurun_out/bench_ops_r01f.log | tail -25
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
