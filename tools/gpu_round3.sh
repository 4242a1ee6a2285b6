set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
EA_PDL=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-sam > gpurun_out/bench_nopdl.json 2> gpurun_out/bench.err; cat gpurun_out/bench_nopdl.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2>> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
NCU_COMMON="--clock-control none --profile-from-start off"
timeout 900 ncu --metrics gpu__time_duration.sum $NCU_COMMON --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ea_attn_kernel -c 1 -o gpurun_out/prof_attn40 -f python tools/gpu_probe_ops.py --one attn_d40_self > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
