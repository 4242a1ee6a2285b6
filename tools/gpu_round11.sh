set -x
cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/gpu_probe_ops.py --one groupnorm 2>&1 | tail -1 | cut -c1-200
timeout 120 python tools/gpu_probe_ops.py --one concurrent_variants 2>&1 | tail -1 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-3000; tail -3 gpurun_out/bench.err
EA_BENCH_TAG=ops_r01k timeout 600 python tools/bench_ops.py norm 2>&1 | tail -9
