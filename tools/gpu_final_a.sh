set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
EA_BENCH_TAG=ops_r01h timeout 600 python tools/bench_ops.py attn gemm norm > gpurun_out/bench_ops_r01h.log 2>&1
NCU_COMMON="--clock-control none --profile-from-start off"
timeout 900 ncu --metrics gpu__time_duration.sum $NCU_COMMON --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on $NCU_COMMON -k regex:ea_gemm_kernel -s 30 -c 4 -o gpurun_out/prof_gemm_step -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
