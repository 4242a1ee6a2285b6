#!/bin/bash
# Builds A/B variants of libea_b200.so into editanything_b200/lib/ab/ (git-ignored, shipped by gpurun) so one GPU
# session can time them on the SAME box (box-to-box spread is 3-6 %, larger than most single changes).
# usage: tools/r02/build_variants.sh NAME GEMM_REV ATTN_FLAGS [GEMM_FLAGS]   (GEMM_REV: git rev of ea_gemm.cu or "wt")
set -e
cd "$(dirname "$0")/../.."
name=$1; rev=$2; attn_flags=$3; gemm_flags=$4
out=editanything_b200/lib/ab; mkdir -p $out/$name
src=editanything_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DEA_PRECISE_MATH -I$src -Iinclude"
if [ "$rev" = wt ]; then cp $src/ea_gemm.cu $out/$name/ea_gemm.cu; else git show $rev:$src/ea_gemm.cu > $out/$name/ea_gemm.cu; fi
nvcc $F $gemm_flags -c $out/$name/ea_gemm.cu -o $out/$name/ea_gemm.o &
nvcc $F $attn_flags -c $src/ea_attn.cu -o $out/$name/ea_attn.o &
nvcc $F -c $src/ea_api.cu -o $out/$name/ea_api.o &
nvcc $F -c $src/ea_pointwise.cu -o $out/$name/ea_pointwise.o &
wait
nvcc -shared -o $out/libea_$name.so $out/$name/*.o -cudart static
rm -rf $out/$name
ls -la $out/libea_$name.so
