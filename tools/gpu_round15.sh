set -x
cd /root/repo
EA_NVCC_EXTRA=-DEA_GEMM_TIMING python -m editanything_b200.csrc.build > /dev/null 2>&1
timeout 300 python tools/exp_gemm_chain.py 2>&1 | tail -60
