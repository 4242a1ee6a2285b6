// Microbenchmark 3: cost of tcgen05.commit in the issue stream, and of TS (A in TMEM) / MN-major B.
#include <cstdio>
#include "ea_common.cuh"
using namespace ea;

// mode 0: G MMAs then commit, repeated;  mode 1: same MMAs, single commit at the end
// ts: A from TMEM; bmn: B MN-major
__global__ void __launch_bounds__(128, 1)
k(int N, int G, int n_grp, int commit_each, int ts, int bmn, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[8];
  __shared__ uint64_t fin;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&fin, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tslot, 512u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 3 && lane == 0) {
    const uint32_t idesc = umma_idesc(128, (uint32_t)N, 0, (uint32_t)bmn);
    const uint64_t da = umma_desc_k_sw128(smem_u32(smem), 1024);
    const uint64_t db = bmn ? umma_desc_mn_sw128(smem_u32(smem) + 16384, 16384, 1024)
                            : umma_desc_k_sw128(smem_u32(smem) + 16384, 1024);
    const uint32_t step_b = bmn ? 128u : 2u;
    long long t0 = clock64();
    int bi = 0;
    for (int g = 0; g < n_grp; ++g) {
      uint64_t a = da, b = db;
      uint32_t pa = tmem + 448u;
      for (int i = 0; i < G; ++i) {
        if (ts) umma_f16_ts(tmem, pa, b, idesc, 1u); else umma_f16_ss(tmem, a, b, idesc, 1u);
        a += 2; b += step_b; pa += 8;
        if ((i & 3) == 3) { a = da; b = db; pa = tmem + 448u; }
      }
      if (commit_each) { umma_commit(&bars[bi]); bi = (bi + 1) & 7; }
    }
    long long t1 = clock64();
    umma_commit(&fin);
    mbar_wait(&fin, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512u); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("%4s %3s %6s %3s %4s | %10s %12s %12s %12s\n", "N", "G", "commit", "ts", "bmn", "issue_cyc", "complete", "cyc/group", "ideal/group");
  struct C { int N, G, ce, ts, bmn; } cs[] = {
      {160, 4, 0, 0, 0}, {160, 4, 1, 0, 0}, {160, 8, 1, 0, 0}, {128, 4, 0, 0, 0}, {128, 4, 1, 0, 0}, {64, 4, 0, 0, 0}, {64, 4, 1, 0, 0},
      {64, 8, 1, 0, 0}, {256, 4, 1, 0, 0}, {48, 8, 0, 1, 1}, {48, 8, 1, 1, 1}, {48, 8, 0, 0, 1}, {48, 8, 0, 1, 0}, {128, 3, 0, 0, 0}, {128, 3, 1, 0, 0}};
  for (auto c : cs) {
    const int n_grp = 32;
    k<<<148, 128, 64 * 1024>>>(c.N, c.G, n_grp, c.ce, c.ts, c.bmn, d_out);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("%4d %3d %6d %3d %4d | %10lld %12lld %12.1f %12.1f\n", c.N, c.G, c.ce, c.ts, c.bmn, h[0], h[1], (double)h[1] / n_grp,
           c.G * c.N / 2.0);
  }
  return 0;
}
