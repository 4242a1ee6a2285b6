// Microbenchmark: tcgen05.mma issue/complete rate on one CTA per SM as a function of N, the number
// of independent accumulators the K-steps are spread over, and the A operand source (smem / TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I editanything_b200/csrc -o mma_rate tools/exp/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "ea_common.cuh"
using namespace ea;

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int N, int n_acc, int n_mma, int a_tmem, int b_mn, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tslot, 512u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc(128, (uint32_t)N, 0, (uint32_t)b_mn);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const int acc_stride = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const int acc = i % n_acc;
        const uint32_t d = tmem + (uint32_t)(acc * acc_stride);
        const uint64_t db = b_mn ? umma_desc_mn_sw128(sb + (uint32_t)((i & 7) * 2048), 16384, 1024)
                                 : umma_desc_k_sw128(sb + (uint32_t)((i & 3) * 32), 1024);
        if (a_tmem) umma_f16_ts(d, tmem + 448u + (uint32_t)((i & 3) * 8), db, idesc, i >= n_acc ? 1u : 0u);
        else umma_f16_ss(d, umma_desc_k_sw128(sa + (uint32_t)((i & 3) * 32), 1024), db, idesc, i >= n_acc ? 1u : 0u);
      }
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, (uint32_t)(rep & 1));
      long long t2 = clock64();
      if (blockIdx.x == 0 && rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512u); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 64;
  printf("%5s %5s %6s %5s | %10s %12s %10s\n", "N", "n_acc", "a_tmem", "b_mn", "issue_cyc", "complete_cyc", "cyc/MMA");
  int Ns[] = {48, 64, 128, 160, 256};
  for (int N : Ns)
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int n_acc = 1; n_acc <= 4; n_acc *= 2) {
          if (n_acc * (N <= 64 ? 64 : (N <= 128 ? 128 : 256)) > 448) continue;
          mma_rate_kernel<<<148, 128, 64 * 1024>>>(N, n_acc, n_mma, a_tmem, b_mn, d_out);
          long long h[2];
          cudaError_t e = cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%5d %5d %6d %5d | %10lld %12lld %10.1f\n", N, n_acc, a_tmem, b_mn, h[0], h[1], (double)h[1] / n_mma);
        }
  return 0;
}
