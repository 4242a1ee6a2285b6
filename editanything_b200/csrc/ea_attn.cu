// ea_attn.cu — fused attention on tcgen05 (sm_100a): O = softmax(Q K^T * scale [+ bias]) V.
//
// Replaces the einsum -> softmax -> einsum of CrossAttention.forward
// (ldm/modules/attention.py:170-193; QK^T accumulated in fp32 like the reference's
// _ATTN_PRECISION="fp32" path) and SAM's windowed / global attention with decomposed relative
// position bias.  The N x N logits never touch HBM.
//
// One CTA = one (batch, head, 128-query tile); 192 threads:
//   warp 0    TMA producer: Q once, then K/V tiles of 128 keys into a 1-2 stage ring.  Head
//             dims that are not a multiple of 64 (40, 80, 160) are zero-padded for free by TMA
//             out-of-bounds fill: the tensor map's innermost extent is d, the box is 64 wide.
//   warp 1    TMEM allocator + MMA issuer: S = Q K^T (M=128, N=128, K=d) into TMEM columns
//             [0,128); O += P V (M=128, N=d, K=128) into columns [128,128+d); V is consumed as
//             an MN-major B operand straight from its row-major [key, d] tile.
//   warps 2-5 softmax: thread <-> query row.  tcgen05.ld the logits, online softmax in fp32
//             (exp2 with folded scale, lazy rescale of O in TMEM only when the running max
//             moves by more than 2^8), P written to shared memory as the K-major SWIZZLE_128B
//             A operand of the second MMA.  Final 1/l normalisation and 16-byte stores.
#include "ea_common.cuh"
#include "ea_internal.h"

namespace ea {

static constexpr int AT_BQ = 128;
static constexpr int AT_BKV = 128;
static constexpr int AT_ATOM = 128 * 128;  // bytes of one 128-row x 64-half swizzled atom
static constexpr int AT_THREADS = 192;

struct AttnKParams {
  int Nq, Nkv, d, heads;
  int nd;      // ceil(d / 64) atoms along the head dim
  int ksteps;  // ceil(d / 16) MMA K-steps for Q K^T
  int dpad16;  // d rounded up to 16: MMA N of the PV product
  int stages;
  int tmem_cols;
  float scale_log2;
  ea_half* out;
  long long o_bs, o_ns;
  const float* rel_h;
  const float* rel_w;
  int rel_s;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
ea_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  // layout: Q[nd] | P[2] | stages x (K[nd] | V[nd]) | barriers
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + p.nd * AT_ATOM;
  uint8_t* sKV = sP + 2 * AT_ATOM;
  const int kv_stage_bytes = 2 * p.nd * AT_ATOM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + p.stages * kv_stage_bytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_final = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Nkv + AT_BKV - 1) / AT_BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_final, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_expect_tx(q_full, (uint32_t)(p.nd * AT_ATOM));
      for (int a = 0; a < p.nd; ++a)
        tma_load_4d(sQ + a * AT_ATOM, &tmQ, q_full, a * 64, head, qt * AT_BQ, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        uint8_t* sK = sKV + stage * kv_stage_bytes;
        uint8_t* sV = sK + p.nd * AT_ATOM;
        mbar_expect_tx(&kv_full[stage], (uint32_t)kv_stage_bytes);
        for (int a = 0; a < p.nd; ++a) {
          tma_load_4d(sK + a * AT_ATOM, &tmK, &kv_full[stage], a * 64, head, j * AT_BKV, b);
          tma_load_4d(sV + a * AT_ATOM, &tmV, &kv_full[stage], a * 64, head, j * AT_BKV, b);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, (uint32_t)p.dpad16, 0, 1);  // B (=V) MN-major
      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
        const uint32_t aK = smem_u32(sKV + stage * kv_stage_bytes);
        const uint32_t aV = aK + p.nd * AT_ATOM;
        const uint32_t aQ = smem_u32(sQ);
        // S = Q K^T
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * AT_ATOM + (ks & 3) * 32);
          umma_f16_ss(tmem_S, umma_desc_k_sw128(aQ + off, 1024), umma_desc_k_sw128(aK + off, 1024),
                      idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        // O += P V   (after the softmax warps published P for this tile)
        mbar_wait(p_full, (uint32_t)(j & 1));
        tc_fence_after();
        const uint32_t aP = smem_u32(sP);
        for (int ks = 0; ks < AT_BKV / 16; ++ks) {
          const uint32_t offP = (uint32_t)((ks >> 2) * AT_ATOM + (ks & 3) * 32);
          const uint32_t offV = (uint32_t)(ks * 16 * 128);  // 16 key rows of 128 B
          umma_f16_ss(tmem_O, umma_desc_k_sw128(aP + offP, 1024),
                      umma_desc_mn_sw128(aV + offV, AT_ATOM, 1024), idesc_o,
                      (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[stage]);
        if (j == n_tiles - 1) umma_commit(o_final);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ======================= softmax / correction / epilogue ==============
    const int wq = warp & 3;  // TMEM lane quarter accessible to this warp
    const int r = wq * 32 + lane;
    const int q = qt * AT_BQ + r;
    const bool row_ok = q < p.Nq;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const long long bh = (long long)b * p.heads + head;
    const float* rh = p.rel_h ? p.rel_h + (bh * p.Nq + (row_ok ? q : 0)) * p.rel_s : nullptr;
    const float* rw = p.rel_w ? p.rel_w + (bh * p.Nq + (row_ok ? q : 0)) * p.rel_s : nullptr;
    float m_used = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      const int kv_left = p.Nkv - j * AT_BKV;  // valid keys in this tile (>= 1)
      // pass 1: row max
      float m_tile = -INFINITY;
      for (int c = 0; c < AT_BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = __uint_as_float(v[i]) * p.scale_log2;
          if (rh) {
            int kk = j * AT_BKV + c + i;
            if (kk < p.Nkv) {
              int kh = kk / p.rel_s;
              t += (__ldg(rh + kh) + __ldg(rw + (kk - kh * p.rel_s))) * LOG2E;
            }
          }
          if (c + i < kv_left) m_tile = fmaxf(m_tile, t);
        }
      }
      // lazy rescale (exact: the final normalisation uses the same m_used)
      float alpha = 1.f;
      bool need = (m_tile > m_used + 8.f);
      if (need) {
        alpha = (m_used == -INFINITY) ? 0.f : exp2f(m_used - m_tile);
        m_used = m_tile;
        l *= alpha;
      }
      const bool warp_need = __any_sync(0xffffffffu, need) && (j > 0);
      if (warp_need) {
        for (int c = 0; c < p.dpad16; c += 16) {
          uint32_t o[16];
          tmem_ld16(tmem_O + lane_off + (uint32_t)c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tmem_O + lane_off + (uint32_t)c, o);
        }
        tmem_st_wait();
      }
      // pass 2: P = exp2(t - m_used), row sum, write swizzled K-major tile
      for (int c = 0; c < AT_BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float t = __uint_as_float(v[i]) * p.scale_log2;
          if (rh) {
            int kk = j * AT_BKV + c + i;
            if (kk < p.Nkv) {
              int kh = kk / p.rel_s;
              t += (__ldg(rh + kh) + __ldg(rw + (kk - kh * p.rel_s))) * LOG2E;
            }
          }
          float e = (c + i < kv_left) ? exp2f(t - m_used) : 0.f;
          pv[i] = e;
          l += e;
        }
        const int atom = c >> 6;
        uint8_t* prow = sP + atom * AT_ATOM + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          int c16 = ((c & 63) >> 3) + g;
          uint4 o = make_uint4(ea_pack2(pv[g * 8 + 0], pv[g * 8 + 1]),
                               ea_pack2(pv[g * 8 + 2], pv[g * 8 + 3]),
                               ea_pack2(pv[g * 8 + 4], pv[g * 8 + 5]),
                               ea_pack2(pv[g * 8 + 6], pv[g * 8 + 7]));
          *reinterpret_cast<uint4*>(prow + ((c16 ^ (r & 7)) << 4)) = o;
        }
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue
    mbar_wait(o_final, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    ea_half* orow = p.out + (long long)b * p.o_bs + (long long)(row_ok ? q : 0) * p.o_ns +
                    (long long)head * p.d;
    for (int c = 0; c < p.dpad16; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_O + lane_off + (uint32_t)c, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 u = make_uint4(
                ea_pack2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l),
                ea_pack2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l),
                ea_pack2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l),
                ea_pack2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l));
            *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int encode_qkv(CUtensorMap* m, const void* base, int d, int heads, int N, int B,
                      long long ns, long long bs) {
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)ns * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[4] = {64, 1, 128, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ea_tmap_encode()(m, EA_TMAP_DTYPE, 4, const_cast<void*>(base), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace ea

using namespace ea;

extern "C" int ea_attention(const ea_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->out) return EA_ERR_ARG;
  if (a->d % 8 != 0 || a->d < 8 || a->d > 192) return EA_ERR_SHAPE;
  if (a->Nq <= 0 || a->Nkv <= 0 || a->B <= 0 || a->heads <= 0) return EA_ERR_SHAPE;
  if (a->q_ns % 8 || a->k_ns % 8 || a->v_ns % 8 || a->q_bs % 8 || a->k_bs % 8 || a->v_bs % 8 ||
      a->o_ns % 8 || a->o_bs % 8)
    return EA_ERR_SHAPE;
  if ((a->rel_h != nullptr) != (a->rel_w != nullptr)) return EA_ERR_ARG;
  if (a->rel_h && a->rel_s <= 0) return EA_ERR_ARG;

  AttnKParams p;
  memset(&p, 0, sizeof(p));
  p.Nq = a->Nq; p.Nkv = a->Nkv; p.d = a->d; p.heads = a->heads;
  p.nd = (a->d + 63) / 64;
  p.ksteps = (a->d + 15) / 16;
  p.dpad16 = p.ksteps * 16;
  p.stages = p.nd <= 2 ? 2 : 1;
  p.tmem_cols = (128 + p.nd * 64) <= 256 ? 256 : 512;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<ea_half*>(a->out);
  p.o_bs = a->o_bs; p.o_ns = a->o_ns;
  p.rel_h = a->rel_h; p.rel_w = a->rel_w; p.rel_s = a->rel_s;

  CUtensorMap tq, tk, tv;
  if (encode_qkv(&tq, a->q, a->d, a->heads, a->Nq, a->B, a->q_ns, a->q_bs)) return EA_ERR_TMAP;
  if (encode_qkv(&tk, a->k, a->d, a->heads, a->Nkv, a->B, a->k_ns, a->k_bs)) return EA_ERR_TMAP;
  if (encode_qkv(&tv, a->v, a->d, a->heads, a->Nkv, a->B, a->v_ns, a->v_bs)) return EA_ERR_TMAP;

  const int smem_bytes =
      p.nd * AT_ATOM + 2 * AT_ATOM + p.stages * 2 * p.nd * AT_ATOM + 9 * 8 + 16 + 1024;
  static int max_set = 0;
  if (smem_bytes > max_set) {
    if (cudaFuncSetAttribute(ea_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes) != cudaSuccess)
      return EA_ERR_CUDA;
    max_set = smem_bytes;
  }
  dim3 grid((unsigned)((a->Nq + AT_BQ - 1) / AT_BQ), (unsigned)a->heads, (unsigned)a->B);
  ea_attn_kernel<<<grid, AT_THREADS, smem_bytes, stream>>>(tq, tk, tv, p);
  ea_count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : EA_ERR_CUDA;
}
