set -x
cd /root/repo
mkdir -p gpurun_out
for c in gemm_2cta conv_2cta gemm_plain conv_64 conv_16_skip gemm_geglu; do timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-700; done
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_BENCH_TAG=ops_2cta timeout 400 python tools/bench_ops.py gemm > gpurun_out/bench_ops_2cta.log 2>&1; cat gpurun_out/bench_ops_2cta.log | tail -25
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
EA_GEMM_2CTA=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sam > gpurun_out/bench_no2cta.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_no2cta.json
