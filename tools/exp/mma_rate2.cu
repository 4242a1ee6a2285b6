// Microbenchmark 2: tcgen05.mma issue cost per instruction for different ways of writing the issue
// loop (single-lane branch vs warp-uniform elect_one; precomputed descriptors; unrolled).
#include <cstdio>
#include <cstdlib>
#include "ea_common.cuh"
using namespace ea;

template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_issue_kernel(int N, int n_iter, long long* out) {
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
  if (warp == 1) {
    const uint32_t idesc = umma_idesc(128, (uint32_t)N, 0, 0);
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem), 1024);
    const uint64_t db = umma_desc_k_sw128(smem_u32(smem) + 16384, 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (MODE == 0) {            // single-lane branch, 4 MMAs per iteration, descriptor add per MMA
      if (lane == 0) {
        t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
        t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t2 = clock64();
      }
    } else {                    // warp-uniform control flow; one elected lane issues
      t0 = clock64();
      for (int it = 0; it < n_iter; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      t2 = clock64();
    }
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512u); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(mma_issue_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(mma_issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_iter = 32;
  printf("%5s %5s | %10s %12s %12s %10s\n", "mode", "N", "issue_cyc", "complete_cyc", "cyc/MMA", "ideal");
  int Ns[] = {48, 128, 160, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int N : Ns) {
      if (mode == 0) mma_issue_kernel<0><<<148, 128, 64 * 1024>>>(N, n_iter, d_out);
      else mma_issue_kernel<1><<<148, 128, 64 * 1024>>>(N, n_iter, d_out);
      long long h[2];
      cudaError_t e = cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("%5d %5d | %10lld %12lld %12.1f %10.1f\n", mode, N, h[0], h[1], (double)h[1] / (4 * n_iter), N / 2.0);
    }
  return 0;
}
