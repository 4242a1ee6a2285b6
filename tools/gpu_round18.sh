set -x
cd /root/repo
mkdir -p gpurun_out
EA_NVCC_EXTRA=-DEA_GEMM_TIMING python -m editanything_b200.csrc.build > /dev/null 2>&1
timeout 400 python tools/exp_gemm_chain.py cold 2>&1 | grep -v "raw launch" | tail -60
