// Flash-attention backward on tcgen05 (head_dim 64): one CTA per (128-key block j, kv head, batch) loops over the
// query blocks (and, for GQA, over the q heads of the group) that attend to it.
//   MMA1  S   = Q_i K_j^T            (128 x 128, K = 64)      MMA2  dP  = dO_i V_j^T
//   softmax warps:  P = exp2(S*scale - LSE_i),  dS = P o (dP - delta_i) * scale   -> smem (bf16, UMMA layouts)
//   MMA3  dV_j += P^T dO_i           (A = P as MN-major operand, accumulates in TMEM over the whole loop)
//   MMA4  dK_j += dS^T Q_i
//   MMA5  dQ_i  = dS K_j             -> TMEM -> fp32 atomics (vectorised red.global) into dQ
// Serves the bidirectional InternViT attention (packed qkv) and the causal GQA Qwen2 attention (KV cache layout,
// key-padding mask).  Warps: 0 = TMA producer, 1 = MMA issuer, 2-5 = softmax / reduction warpgroup.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int HD = 64, BQ = 128, BKV = 128;
constexpr int BWD_THREADS = 192;
constexpr int kTile = 128 * HD * 2;  // 16 KB
constexpr int kSmK = 0, kSmV = kTile;
constexpr int kSmQ = 2 * kTile;             // 2 stages
constexpr int kSmdO = kSmQ + 2 * kTile;     // 2 stages
constexpr int kSmP = kSmdO + 2 * kTile;     // 32 KB
constexpr int kSmdS = kSmP + 2 * kTile;     // 32 KB
constexpr int kSmBar = kSmdS + 2 * kTile;
constexpr int kSmTotal = kSmBar + 256;

struct BwdParams {
  int lq, lkv, past, causal;
  int hq, group;
  int q_col0, k_col0, v_col0;
  int kv_head_col_stride, kv_batch_stride, kv_head_batch_stride;
  const uint8_t* key_valid;
  int key_valid_ld;
  const float* lse;    // [B, hq, lq]
  const float* delta;  // [B, hq, lq]
  float* dq;           // [B*lq, hq*64] fp32, zero-initialised, accumulated with atomics
  float* dk;           // [B, hkv, lkv, 64] fp32
  float* dv;           // [B, hkv, lkv, 64] fp32
  float scale, scale_log2;
};

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do, BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* kv_full = bars;        // 1
  uint64_t* qdo_full = bars + 1;   // 2
  uint64_t* qdo_empty = bars + 3;  // 2
  uint64_t* sdp_full = bars + 5;   // 1
  uint64_t* p_ready = bars + 6;    // 4 arrivals
  uint64_t* dq_full = bars + 7;    // 1
  uint64_t* dq_free = bars + 8;    // 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int jb = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * BKV;
  const int nqb = (p.lq + BQ - 1) / BQ;
  // first query block that can see this key block
  int i_first = 0;
  if (p.causal) i_first = max(0, (k0 - p.past) / BQ);
  const int n_i = max(0, nqb - i_first);
  const int n_items = n_i * p.group;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023) != 0) __trap();
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320, tm_dq = tmem_base + 384;

  if (warp == 0) {
    const int kc = p.k_col0 + hk * p.kv_head_col_stride, vc = p.v_col0 + hk * p.kv_head_col_stride;
    const int kb = b * p.kv_batch_stride + hk * p.kv_head_batch_stride;
    if (elect_one_sync()) {
      mbar_expect_tx(kv_full, 2 * kTile);
      tma_load_3d(smem + kSmK, &tmap_k, kv_full, kc, k0, kb);
      tma_load_3d(smem + kSmV, &tmap_v, kv_full, vc, k0, kb);
    }
    __syncwarp();
    for (int it = 0; it < n_items; ++it) {
      const int s = it & 1;
      const int h = hk * p.group + it / n_i, i = i_first + it % n_i;
      mbar_wait(&qdo_empty[s], ((it >> 1) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&qdo_full[s], 2 * kTile);
        tma_load_3d(smem + kSmQ + s * kTile, &tmap_q, &qdo_full[s], p.q_col0 + h * HD, i * BQ, b);
        tma_load_3d(smem + kSmdO + s * kTile, &tmap_do, &qdo_full[s], h * HD, i * BQ, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_kv = umma_idesc_bf16(BKV, HD, 1, 1);  // P^T dO, dS^T Q: both operands MN-major
    constexpr uint32_t idesc_dq = umma_idesc_bf16(BQ, HD, 0, 1);   // dS K: A K-major, B MN-major
    const uint32_t sk = smem_u32(smem + kSmK), sv = smem_u32(smem + kSmV), sp = smem_u32(smem + kSmP), sds = smem_u32(smem + kSmdS);
    const uint64_t dk_k = umma_desc_kmajor_sw128(sk), dv_k = umma_desc_kmajor_sw128(sv);
    const uint64_t dk_mn = umma_desc_mnmajor_sw128(sk, kTile);
    const uint64_t dp_mn = umma_desc_mnmajor_sw128(sp, kTile), dds_mn = umma_desc_mnmajor_sw128(sds, kTile);
    const uint64_t dds_k0 = umma_desc_kmajor_sw128(sds), dds_k1 = umma_desc_kmajor_sw128(sds + kTile);
    mbar_wait(kv_full, 0);
    for (int it = 0; it < n_items; ++it) {
      const int s = it & 1;
      mbar_wait(&qdo_full[s], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t sq = smem_u32(smem + kSmQ + s * kTile), sdo = smem_u32(smem + kSmdO + s * kTile);
      const uint64_t dq_k = umma_desc_kmajor_sw128(sq), ddo_k = umma_desc_kmajor_sw128(sdo);
      const uint64_t dq_mn = umma_desc_mnmajor_sw128(sq, kTile), ddo_mn = umma_desc_mnmajor_sw128(sdo, kTile);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tm_s, dq_k + 2 * k, dk_k + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tm_dp, ddo_k + 2 * k, dv_k + 2 * k, idesc_s, k != 0);
        tc_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(p_ready, it & 1);
      if (it > 0) mbar_wait(dq_free, (it - 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc_mma_bf16(tm_dv, dp_mn + (uint64_t)k * 128, ddo_mn + (uint64_t)k * 128, idesc_kv, (it | k) != 0);
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc_mma_bf16(tm_dk, dds_mn + (uint64_t)k * 128, dq_mn + (uint64_t)k * 128, idesc_kv, (it | k) != 0);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          tc_mma_bf16(tm_dq, (k < 4 ? dds_k0 : dds_k1) + 2 * (k & 3), dk_mn + (uint64_t)k * 128, idesc_dq, k != 0);
        tc_commit(dq_full);
        tc_commit(&qdo_empty[s]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint8_t* prow = smem + kSmP + r * 128;
    uint8_t* dsrow = smem + kSmdS + r * 128;
    const int rsw = r & 7;
    const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)b * p.key_valid_ld : nullptr;
    const float lg2e = 1.4426950408889634f;

    for (int it = 0; it < n_items; ++it) {
      const int h = hk * p.group + it / n_i, i = i_first + it % n_i;
      const int row = i * BQ + r;
      const bool row_ok = row < p.lq;
      const int qpos = p.past + row;
      float lse2 = 0.f, dlt = 0.f;
      if (row_ok) {
        const float l = p.lse[((size_t)b * p.hq + h) * p.lq + row];
        lse2 = (l == -INFINITY) ? INFINITY : l * lg2e;  // fully masked row -> P = 0
        dlt = p.delta[((size_t)b * p.hq + h) * p.lq + row];
      }
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t sr[32], dpr[32];
        tmem_ld_32x32(tm_s + lane_off + c, sr);
        tmem_ld_32x32(tm_dp + lane_off + c, dpr);
        tmem_ld_wait();
        uint32_t pk[16], dk16[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float pv[2], dsv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = k0 + c + e + u;
            bool ok = row_ok && col < p.lkv && (!p.causal || col <= qpos);
            if (ok && kvalid) ok = kvalid[col] != 0;
            const float pe = ok ? ex2_approx(__uint_as_float(sr[e + u]) * p.scale_log2 - lse2) : 0.f;
            pv[u] = pe;
            dsv[u] = pe * (__uint_as_float(dpr[e + u]) - dlt) * p.scale;
          }
          pk[e >> 1] = pack_bf16(pv[0], pv[1]);
          dk16[e >> 1] = pack_bf16(dsv[0], dsv[1]);
        }
        const int hoff = (c >> 6) * kTile, chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int off = hoff + (((chunk0 + q4) ^ rsw) << 4);
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(dsrow + off) = make_uint4(dk16[4 * q4], dk16[4 * q4 + 1], dk16[4 * q4 + 2], dk16[4 * q4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // dQ_i tile -> global fp32 accumulation
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      float* dqrow = p.dq + ((size_t)b * p.lq + row) * (size_t)(p.hq * HD) + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld_32x32(tm_dq + lane_off + c, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            atomicAdd(reinterpret_cast<float4*>(dqrow + c + e),
                      make_float4(__uint_as_float(o[e]), __uint_as_float(o[e + 1]), __uint_as_float(o[e + 2]), __uint_as_float(o[e + 3])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    }
    // dK_j / dV_j: thread <-> key row
    if (n_items > 0) {
      mbar_wait(dq_full, (n_items - 1) & 1);  // the last commit covers every MMA3 / MMA4 issued
      tc_fence_after();
    }
    const int key = k0 + r;
    const bool key_ok = key < p.lkv;
    float* dkrow = p.dk + (((size_t)b * (p.hq / p.group) + hk) * p.lkv + (key_ok ? key : 0)) * HD;
    float* dvrow = p.dv + (((size_t)b * (p.hq / p.group) + hk) * p.lkv + (key_ok ? key : 0)) * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t a[32], bb[32];
      if (n_items > 0) {  // warp-uniform: the TMEM loads stay convergent for partially valid warps
        tmem_ld_32x32(tm_dk + lane_off + c, a);
        tmem_ld_32x32(tm_dv + lane_off + c, bb);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) { a[e] = 0; bb[e] = 0; }
      }
      if (key_ok) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          *reinterpret_cast<float4*>(dkrow + c + e) = make_float4(__uint_as_float(a[e]), __uint_as_float(a[e + 1]), __uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
          *reinterpret_cast<float4*>(dvrow + c + e) = make_float4(__uint_as_float(bb[e]), __uint_as_float(bb[e + 1]), __uint_as_float(bb[e + 2]), __uint_as_float(bb[e + 3]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, const BwdParams& p, int batch,
               int hkv, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SLB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.lkv, BKV), hkv, batch);
  attn_bwd_kernel<<<grid, BWD_THREADS, kSmTotal, stream>>>(tq, tk, tv, tdo, p);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

}  // namespace

extern "C" int slb_attn_vit_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, float* dq, float* dk, float* dv,
                                int tiles, int n_tokens, int heads, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && n_tokens > 0 && heads > 0 && lse && delta && dq && dk && dv, "attn_vit_bwd: bad args");
  const int C = heads * HD;
  CUtensorMap tm, tdo;
  int rc = slb_make_tmap_3d(&tm, qkv, (uint64_t)3 * C, (uint64_t)n_tokens, (uint64_t)tiles, (uint64_t)3 * C * 2,
                            (uint64_t)n_tokens * 3 * C * 2, HD, BQ, 1);
  if (rc) return rc;
  rc = slb_make_tmap_3d(&tdo, dout, (uint64_t)C, (uint64_t)n_tokens, (uint64_t)tiles, (uint64_t)C * 2, (uint64_t)n_tokens * C * 2, HD, BQ, 1);
  if (rc) return rc;
  BwdParams p{};
  p.lq = n_tokens; p.lkv = n_tokens; p.past = 0; p.causal = 0;
  p.hq = heads; p.group = 1;
  p.q_col0 = 0; p.k_col0 = C; p.v_col0 = 2 * C;
  p.kv_head_col_stride = HD; p.kv_batch_stride = 1; p.kv_head_batch_stride = 0;
  p.key_valid = nullptr; p.key_valid_ld = 0;
  p.lse = lse; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_bwd(tm, tm, tm, tdo, p, tiles, heads, (cudaStream_t)stream);
}

extern "C" int slb_attn_gqa_bwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid,
                                int key_valid_ld, const void* dout, const float* lse, const float* delta, float* dq, float* dk, float* dv,
                                int batch, int lq, int lmax, int hq, int hkv, void* stream) {
  SLB_CHECK_ARG(batch > 0 && lq > 0 && lq <= lmax && hq % hkv == 0 && (ldq % 8) == 0, "attn_gqa_bwd: bad shape");
  CUtensorMap tq, tk, tv, tdo;
  int rc = slb_make_tmap_3d(&tq, q, (uint64_t)ldq, (uint64_t)lq, (uint64_t)batch, (uint64_t)ldq * 2, (uint64_t)lq * ldq * 2, HD, BQ, 1);
  if (rc) return rc;
  rc = slb_make_tmap_3d(&tk, kcache, HD, (uint64_t)lmax, (uint64_t)batch * hkv, HD * 2, (uint64_t)lmax * HD * 2, HD, BKV, 1);
  if (rc) return rc;
  rc = slb_make_tmap_3d(&tv, vcache, HD, (uint64_t)lmax, (uint64_t)batch * hkv, HD * 2, (uint64_t)lmax * HD * 2, HD, BKV, 1);
  if (rc) return rc;
  const int C = hq * HD;
  rc = slb_make_tmap_3d(&tdo, dout, (uint64_t)C, (uint64_t)lq, (uint64_t)batch, (uint64_t)C * 2, (uint64_t)lq * C * 2, HD, BQ, 1);
  if (rc) return rc;
  BwdParams p{};
  p.lq = lq; p.lkv = lq; p.past = 0; p.causal = 1;
  p.hq = hq; p.group = hq / hkv;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0;
  p.kv_head_col_stride = 0; p.kv_batch_stride = hkv; p.kv_head_batch_stride = 1;
  p.key_valid = key_valid; p.key_valid_ld = key_valid_ld;
  p.lse = lse; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_bwd(tq, tk, tv, tdo, p, batch, hkv, (cudaStream_t)stream);
}
