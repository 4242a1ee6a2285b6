# round 2, GPU session 4: profiles of the lockstep step (launch list with DRAM bytes, ncu --set full of the heaviest
# kernels, per-shape GEMM breakdown) + the fixed grouped-operator cases
set -x
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "grouped or ln_fold or step_gather or groupnorm" ) > gpurun_out/s4_pytest_ops.log 2>&1
tail -4 gpurun_out/s4_pytest_ops.log
NCU_COMMON="--clock-control none --profile-from-start off"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --no-vae --no-batch4 --profiler-range"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum $NCU_COMMON --csv --log-file gpurun_out/s4_launches.csv $BENCH > gpurun_out/s4_ncu_bench.log 2>&1
tail -2 gpurun_out/s4_ncu_bench.log
python tools/summarize_launches.py gpurun_out/s4_launches.csv --traffic-json gpurun_out/s4_gemm_traffic.json > gpurun_out/s4_launches_summary.txt 2>&1; head -50 gpurun_out/s4_launches_summary.txt
timeout 900 ncu --set full --import-source on $NCU_COMMON -k regex:ea_gemm -c 22 -o gpurun_out/s4_gemm_full $BENCH > gpurun_out/s4_ncu_gemm.log 2>&1; tail -2 gpurun_out/s4_ncu_gemm.log
timeout 900 ncu --set full --import-source on $NCU_COMMON -k regex:"gn_fused|attn_db|conv_smallcin|step_gather" -c 8 -o gpurun_out/s4_misc_full $BENCH > gpurun_out/s4_ncu_misc.log 2>&1; tail -2 gpurun_out/s4_ncu_misc.log
ls -la gpurun_out/*.ncu-rep
timeout 500 python tools/gemm_breakdown.py gpurun_out/s4_gemm_breakdown_lockstep.json 2>&1 | head -45
