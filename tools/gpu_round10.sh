set -x
cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/gpu_probe_ops.py --one concurrent_variants 2>&1 | tail -1 | cut -c1-400
timeout 120 python tools/gpu_probe_ops.py --one groupnorm 2>&1 | tail -1 | cut -c1-300
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sam > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-1700; tail -3 gpurun_out/bench.err
EA_CONCURRENT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sam > gpurun_out/bench_serial.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_serial.json | cut -c1-400
EA_BENCH_TAG=ops_r01j timeout 600 python tools/bench_ops.py norm 2>&1 | tail -9
