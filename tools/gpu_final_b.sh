set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | head -2
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01q.json 2> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench_r01q.json; tail -3 gpurun_out/bench.err
NCU_COMMON="--clock-control none --profile-from-start off"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum $NCU_COMMON --csv --log-file gpurun_out/launches_r01q.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --import-source on $NCU_COMMON -k regex:ea_gemm_kernel -s 30 -c 4 -o gpurun_out/prof_gemm_r01q -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full $NCU_COMMON -k regex:ea_attn_db_kernel -c 1 -o gpurun_out/prof_attn_db_r01q -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --profiler-range > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
