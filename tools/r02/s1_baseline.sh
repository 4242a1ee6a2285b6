# round 2, GPU session 1: parity on the new full-size goldens + A/B measurements that decide the perf plan.
set -x
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/s1_pytest_gpu.log 2>&1
tail -8 gpurun_out/s1_pytest_gpu.log
grep -h "eps max-abs" gpurun_out/s1_pytest_gpu.log
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "full" ) > gpurun_out/s1_pytest_full.log 2>&1
grep -h "eps max-abs\|passed\|failed" gpurun_out/s1_pytest_full.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_base.json 2> gpurun_out/s1_bench_base.err; cut -c1-400 gpurun_out/s1_bench_base.json
EA_CONCURRENT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s1_bench_serial.json 2>/dev/null; cut -c1-300 gpurun_out/s1_bench_serial.json
timeout 500 python tools/gemm_breakdown.py gpurun_out/s1_gemm_breakdown_base.json 2>&1 | head -40
EA_GEMM_PERSIST=2 timeout 500 python tools/gemm_breakdown.py gpurun_out/s1_gemm_breakdown_persist2.json 2>&1 | head -40
EA_GEMM_PERSIST=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s1_bench_persist2.json 2>/dev/null; cut -c1-300 gpurun_out/s1_bench_persist2.json
EA_GEMM_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sam --no-vae --no-e2e > gpurun_out/s1_bench_persist1.json 2>/dev/null; cut -c1-300 gpurun_out/s1_bench_persist1.json
