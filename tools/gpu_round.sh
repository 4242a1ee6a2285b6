#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/host.txt 2>&1; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/host.txt
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
NCU_COMMON="--clock-control none --profile-from-start off"
timeout 900 ncu --metrics gpu__time_duration.sum $NCU_COMMON --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on $NCU_COMMON -k regex:ea_gemm_kernel -s 60 -c 6 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
