set -x
cd /root/repo
mkdir -p gpurun_out
EA_NVCC_EXTRA=-DEA_GEMM_TIMING python -m editanything_b200.csrc.build > /dev/null 2>&1
timeout 400 python tools/exp_step_chain.py gpurun_out/step_chain_r01q.json 2>&1 | head -24
python -m editanything_b200.csrc.build --force > /dev/null 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | head -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-2600
