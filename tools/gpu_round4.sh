set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_r01d timeout 400 python tools/bench_ops.py attn gemm > gpurun_out/bench_ops_r01d.log 2>&1; cat gpurun_out/bench_ops_r01d.log | tail -40
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
NCU_COMMON="--clock-control none --profile-from-start off"
timeout 900 ncu --metrics gpu__time_duration.sum $NCU_COMMON --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_bench.log 2>&1
