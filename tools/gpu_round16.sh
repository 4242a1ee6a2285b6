set -x
cd /root/repo
mkdir -p gpurun_out
EA_NVCC_EXTRA=-DEA_GEMM_TIMING python -m editanything_b200.csrc.build > /dev/null 2>&1
timeout 300 python tools/exp_gemm_chain.py 2>&1 | grep -v "raw launch" | tail -40
python -m editanything_b200.csrc.build --force > /dev/null 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-2600
timeout 500 python tools/gemm_breakdown.py gpurun_out/gemm_breakdown_r01n.json 2>&1 | head -30
