set -x
cd /root/repo
for c in attn_d64 attn_d40_self attn_d80_self attn_d160_self attn_cross77 attn_cross77_d40 attn_sam_window attn_sam_global; do timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-300; done > gpurun_out/attn_tmem.log 2>&1
cat gpurun_out/attn_tmem.log
for c in attn_d64 attn_d40_self attn_d80_self attn_d160_self attn_cross77 attn_sam_window; do EA_ATTN_P_SMEM=1 timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-300; done > gpurun_out/attn_smem.log 2>&1
cat gpurun_out/attn_smem.log
for c in attn_d64 attn_d40_self attn_d80_self attn_cross77 attn_sam_window; do EA_ATTN_NQT=1 timeout 120 python tools/gpu_probe_ops.py --one $c 2>&1 | tail -1 | cut -c1-300; done > gpurun_out/attn_nqt1.log 2>&1
cat gpurun_out/attn_nqt1.log
EA_BENCH_TAG=attn_tmem timeout 300 python tools/bench_ops.py attn 2>&1 | tail -12
EA_ATTN_P_SMEM=1 EA_BENCH_TAG=attn_smem timeout 300 python tools/bench_ops.py attn 2>&1 | tail -12
EA_BENCH_TAG=gemm_base timeout 300 python tools/bench_ops.py gemm norm 2>&1 | tail -40
