# round 2, GPU session 19: validation of the final build - full GPU suite, bench (both arms), launch list, ncu --set full of the GEMMs
set -x
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/s19_pytest_gpu.log 2>&1
tail -6 gpurun_out/s19_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; tail -3 gpurun_out/s19_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s19_bench.json'))
print('ms_step', d['ms_per_step'], 'launches', d['launches_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_image'], 'image_ms', d['config']['image_ms'], 'value', d['value'], 'cpu', d.get('cpu_baseline'))
print('batch4', d['config'].get('batch4'))
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/s19_bench_reference.json 2> gpurun_out/s19_bench_reference.err; tail -2 gpurun_out/s19_bench_reference.json | cut -c1-600
NCU_COMMON="--clock-control none --profile-from-start off"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sam --no-vae --no-batch4 --profiler-range"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum $NCU_COMMON --csv --log-file gpurun_out/s19_launches.csv $BENCH > gpurun_out/s19_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/s19_launches.csv --traffic-json gpurun_out/s19_gemm_traffic.json > gpurun_out/s19_launches_summary.txt 2>&1; head -30 gpurun_out/s19_launches_summary.txt
timeout 900 ncu --set full --import-source on $NCU_COMMON -k regex:ea_gemm -c 24 -o gpurun_out/s19_gemm_full $BENCH > gpurun_out/s19_ncu_gemm.log 2>&1; tail -2 gpurun_out/s19_ncu_gemm.log
timeout 500 python tools/gemm_breakdown.py gpurun_out/s19_gemm_breakdown.json 2>&1 | head -30
