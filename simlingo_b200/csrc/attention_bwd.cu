// Flash-attention backward on tcgen05 (head_dim 64): one CTA per (128-key block j, kv head, batch) loops over the
// query blocks (and, for GQA, over the q heads of the group) that attend to it.
//   MMA1  S   = Q_i K_j^T            (128 x 128, K = 64)      MMA2  dP  = dO_i V_j^T
//   softmax warps:  P = exp2(S*scale - LSE_i),  dS = P o (dP - delta_i) * scale   -> smem (bf16, UMMA layouts)
//   MMA3  dV_j += P^T dO_i           (A = P as MN-major operand, accumulates in TMEM over the whole loop)
//   MMA4  dK_j += dS^T Q_i
//   MMA5  dQ_i  = dS K_j             -> TMEM -> bf16 smem staging -> TMA store into this key block's partial-dQ slab;
//                                       dq_reduce_kernel sums the slabs (fp32) afterwards.  (Accumulating dQ in global
//                                       memory instead - fp32 atomics or TMA reduce-add - is limited by the L2 reduction
//                                       rate: 680 MB of read-modify-write per ViT layer, 0.8 ms measured.)
// Serves the bidirectional InternViT attention (packed qkv) and the causal GQA Qwen2 attention (KV cache layout,
// key-padding mask).  Warps: 0 = TMA producer, 1 = MMA issuer, 2-5 / 6-9 = two softmax warpgroups that split the
// 128 key columns of S / dP (and the 64 dQ columns) between them.  The MMA warp issues S/dP of item i+1 right behind
// dV/dK/dQ of item i, so the tensor pipe only idles while the softmax of one tile runs.
#include <cstdlib>
#include "common.cuh"
#include "../../include/simlingo_b200.h"

long long* slb_debug_trace_ptr();  // attention_vit.cu (slb_debug_set_trace)

namespace {

constexpr int HD = 64, BQ = 128, BKV = 128;
constexpr int BWD_THREADS = 576;  // TMA warp, MMA warp, 4 softmax warpgroups (32 key columns of S / dP each)
constexpr int kSoftmaxThreads = BWD_THREADS - 64;
constexpr int kTile = 128 * HD * 2;  // 16 KB
constexpr int kSmK = 0, kSmV = kTile;
constexpr int kQS = 3;                      // Q / dO ring depth
constexpr int kSmQ = 2 * kTile;
constexpr int kSmdO = kSmQ + kQS * kTile;
constexpr int kSmP = kSmdO + kQS * kTile;   // 32 KB
constexpr int kSmdS = kSmP + 2 * kTile;     // 32 KB
constexpr int kSmdQ = kSmdS + 2 * kTile;    // [128 rows x 64 bf16] staging of the partial dQ tile (128-byte swizzled rows)
constexpr int kSmBar = kSmdQ + kTile;
constexpr int kSmTotal = kSmBar + 256;

// optional timeline trace (debug): CTA (0,0,0) records (tag, clock64) pairs per role; enabled by slb_debug_set_trace()
constexpr int kTraceMax = 256;
struct Trace {
  long long* buf;  // [8 roles][kTraceMax][2]
  int role;
  int n;
  __device__ __forceinline__ void rec(int tag) {
    if (buf && n < kTraceMax) {
      buf[((size_t)role * kTraceMax + n) * 2] = tag;
      buf[((size_t)role * kTraceMax + n) * 2 + 1] = clock64();
      ++n;
    }
  }
};

struct BwdParams {
  long long* trace;
  int variant;  // always 0 unless built with -DSLB_ABLATION (SLB_BWD_VARIANT timing experiments: 1 no dV MMAs, 2 no dK, 4 no dQ, 8 no softmax math)
  int lq, lkv, past, causal;
  int hq, group;
  int q_col0, k_col0, v_col0;
  int kv_head_col_stride, kv_batch_stride, kv_head_batch_stride;
  const uint8_t* key_valid;
  int key_valid_ld;
  const float* lse;    // [B, hq, lq]
  const float* delta;  // [B, hq, lq]
  int batch;           // partial dQ slabs: bf16 [n_key_blocks, B, lq, hq*64], slab index = jb * batch + b (via tmap_dq)
  float* dk;           // [B, hkv, lkv, 64] fp32   (dkv_bf16 == nullptr)
  float* dv;           // [B, hkv, lkv, 64] fp32
  bf16* dkv_bf16;      // packed output (ViT): row (b*lkv + key) * dkv_ld, dK at dk_col0 + hk*64, dV at dv_col0 + hk*64
  int dkv_ld, dk_col0, dv_col0;
  float scale, scale_log2;
};

// P = exp2(S * scale_log2 - lse2), dS = P * (dP - delta) * scale for 32 columns of one row; MASKED adds the column
// validity bitmask (key bound / key-padding) and the causal test - only boundary and diagonal tiles need it.
template <bool MASKED>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&sr)[32], const uint32_t (&dpr)[32], uint32_t (&pk)[16], uint32_t (&dk16)[16],
                                              float scale_log2, float neg_lse2, float scale, float neg_dlt_s, uint32_t cmask, int col0, int qmax) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float pv[2], dsv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float pe = ex2_approx(fmaf(__uint_as_float(sr[e + u]), scale_log2, neg_lse2));
      if (MASKED) pe = (((cmask >> (e + u)) & 1u) && (col0 + e + u <= qmax)) ? pe : 0.f;
      pv[u] = pe;
      dsv[u] = pe * fmaf(__uint_as_float(dpr[e + u]), scale, neg_dlt_s);
    }
    pk[e >> 1] = pack_bf16(pv[0], pv[1]);
    dk16[e >> 1] = pack_bf16(dsv[0], dsv[1]);
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* kv_full = bars;        // 1
  uint64_t* qdo_full = bars + 1;   // kQS
  uint64_t* qdo_empty = bars + 5;  // kQS
  uint64_t* sdp_full = bars + 9;   // 1
  uint64_t* p_ready = bars + 10;   // 8 arrivals
  uint64_t* dq_full = bars + 11;   // 1
  uint64_t* dq_free = bars + 12;   // 8 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

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
    tma_prefetch_desc(&tmap_dq);
    mbar_init(kv_full, 1);
    for (int s = 0; s < kQS; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(p_ready, kSoftmaxThreads / 32);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, kSoftmaxThreads / 32);
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
      const int s = it % kQS;
      const int h = hk * p.group + it / n_i, i = i_first + it % n_i;
      mbar_wait(&qdo_empty[s], ((it / kQS) & 1) ^ 1);
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
    Trace tr{(blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace : nullptr, 0, 0};
    auto issue_sdp = [&](int it) {
      const int s = it % kQS;
      mbar_wait(&qdo_full[s], (it / kQS) & 1);
      tc_fence_after();
      const uint32_t sq = smem_u32(smem + kSmQ + s * kTile), sdo = smem_u32(smem + kSmdO + s * kTile);
      const uint64_t dq_k = umma_desc_kmajor_sw128(sq), ddo_k = umma_desc_kmajor_sw128(sdo);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tm_s, dq_k + 2 * k, dk_k + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tm_dp, ddo_k + 2 * k, dv_k + 2 * k, idesc_s, k != 0);
        tc_commit(sdp_full);
      }
      __syncwarp();
    };
    if (n_items > 0) issue_sdp(0);
    for (int it = 0; it < n_items; ++it) {
      const int s = it % kQS;
      const uint32_t sq = smem_u32(smem + kSmQ + s * kTile), sdo = smem_u32(smem + kSmdO + s * kTile);
      const uint64_t dq_mn = umma_desc_mnmajor_sw128(sq, kTile), ddo_mn = umma_desc_mnmajor_sw128(sdo, kTile);
      tr.rec(10 + 4 * it);
      mbar_wait(p_ready, it & 1);   // P / dS of item `it` are in smem and its S / dP have been read out of TMEM
      tr.rec(11 + 4 * it);
      tc_fence_after();
      // S / dP of the next item first: the softmax warps start on them while dV / dK / dQ of this item are running
      if (it + 1 < n_items) issue_sdp(it + 1);
      tr.rec(12 + 4 * it);
      if (it > 0) mbar_wait(dq_free, (it - 1) & 1);
      tr.rec(13 + 4 * it);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        if (!(p.variant & 1))
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc_mma_bf16(tm_dv, dp_mn + (uint64_t)k * 128, ddo_mn + (uint64_t)k * 128, idesc_kv, (it | k) != 0);
        if (!(p.variant & 2))
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc_mma_bf16(tm_dk, dds_mn + (uint64_t)k * 128, dq_mn + (uint64_t)k * 128, idesc_kv, (it | k) != 0);
        if (!(p.variant & 4))
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          tc_mma_bf16(tm_dq, (k < 4 ? dds_k0 : dds_k1) + 2 * (k & 3), dk_mn + (uint64_t)k * 128, idesc_dq, k != 0);
        tc_commit(dq_full);       // dQ tile ready; P / dS smem may be overwritten
        tc_commit(&qdo_empty[s]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int wg = (warp - 2) >> 2;  // 0..3: key columns 32 wg .. 32 wg + 31 of S / dP, columns 16 wg .. + 15 of dQ / dK / dV
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint8_t* prow = smem + kSmP + (wg >> 1) * kTile + r * 128;
    uint8_t* dsrow = smem + kSmdS + (wg >> 1) * kTile + r * 128;
    uint8_t* dqstage = smem + kSmdQ;
    const int rsw = r & 7;
    const bool issuer = (warp == 2) && (lane == 0);
    const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)b * p.key_valid_ld : nullptr;
    const float lg2e = 1.4426950408889634f;
    // validity of this warpgroup's 32 key columns (fixed for the whole CTA): inside the sequence and not padding
    uint32_t cmask0 = 0u;
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
      const int col = k0 + wg * 32 + c;
      cmask0 |= ((col < p.lkv && (!kvalid || kvalid[col] != 0)) ? 1u : 0u) << c;
    }
    const bool cols_all = cmask0 == 0xffffffffu;

    // partial dQ tile of item `jt` (this warpgroup's 32 columns): TMEM -> bf16 -> swizzled smem staging -> TMA store
    auto dq_out = [&](int hh, int ii) {
      uint32_t o[16];
      tmem_ld_32x16(tm_dq + lane_off + wg * 16, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (issuer) tma_store_wait_read<0>();  // the previous store has finished reading the staging tile
      __syncwarp();
      named_bar_sync(1, kSoftmaxThreads);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(o[8 * j]), __uint_as_float(o[8 * j + 1]));
        v.y = pack_bf16(__uint_as_float(o[8 * j + 2]), __uint_as_float(o[8 * j + 3]));
        v.z = pack_bf16(__uint_as_float(o[8 * j + 4]), __uint_as_float(o[8 * j + 5]));
        v.w = pack_bf16(__uint_as_float(o[8 * j + 6]), __uint_as_float(o[8 * j + 7]));
        *reinterpret_cast<uint4*>(dqstage + r * 128 + (((2 * wg + j) ^ rsw) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      named_bar_sync(1, kSoftmaxThreads);
      if (issuer) {
        tma_store_3d(&tmap_dq, dqstage, hh * HD, ii * BQ, jb * p.batch + b);
        tma_store_commit();
      }
    };
    // (head, query block) of an item advance incrementally: a runtime integer division costs ~500 cycles here
    auto load_stats = [&](int jt, int hh, int ii, float& l, float& d) {
      const int row = ii * BQ + r;
      l = -INFINITY; d = 0.f;
      if (jt < n_items && row < p.lq) {
        l = p.lse[((size_t)b * p.hq + hh) * p.lq + row];
        d = p.delta[((size_t)b * p.hq + hh) * p.lq + row];
      }
    };
    float l_next, d_next;
    int h_cur = hk * p.group, i_cur = i_first, h_prev = 0, i_prev = 0;
    load_stats(0, h_cur, i_cur, l_next, d_next);
    Trace ts{(blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && quad == 2 && wg < 2) ? p.trace : nullptr, 1 + wg, 0};

    for (int it = 0; it < n_items; ++it) {
      const int i = i_cur;
      const int row = i * BQ + r;
      // rows past the end / fully masked rows carry lse = -inf: P = exp2(-inf) = 0
      const float lse2 = (l_next == -INFINITY) ? INFINITY : l_next * lg2e, dlt = d_next;
      int h_nxt = h_cur, i_nxt = i_cur + 1;
      if (i_nxt == i_first + n_i) { i_nxt = i_first; ++h_nxt; }
      load_stats(it + 1, h_nxt, i_nxt, l_next, d_next);  // prefetch: the global-load latency hides behind this item's math
      const float neg_lse2 = -lse2, neg_dlt_s = -dlt * p.scale;
      const int qmax = p.causal ? p.past + row : 0x7fffffff;
      // warp-uniform: does any (row, column) of this tile need the mask?
      const bool masked = !cols_all || (p.causal && (k0 + BKV - 1 > p.past + i * BQ));
      ts.rec(10 + 6 * it);
      mbar_wait(sdp_full, it & 1);
      ts.rec(11 + 6 * it);
      tc_fence_after();
      {
        const int c = wg * 32;
        uint32_t sr[32], dpr[32];
        tmem_ld_32x32(tm_s + lane_off + c, sr);
        tmem_ld_32x32(tm_dp + lane_off + c, dpr);
        tmem_ld_wait();
        uint32_t pk[16], dk16[16];
        if (p.variant & 8) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { pk[e] = sr[e]; dk16[e] = dpr[e]; }
        } else if (masked) softmax_chunk<true>(sr, dpr, pk, dk16, p.scale_log2, neg_lse2, p.scale, neg_dlt_s, cmask0, k0 + c, qmax);
        else softmax_chunk<false>(sr, dpr, pk, dk16, p.scale_log2, neg_lse2, p.scale, neg_dlt_s, 0u, 0, 0);
        if (it > 0) {
          // dV / dK / dQ of the previous item have completed: P / dS smem is free again and its dQ tile is in TMEM
          ts.rec(12 + 6 * it);
          mbar_wait(dq_full, (it - 1) & 1);
          ts.rec(13 + 6 * it);
          tc_fence_after();
        }
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int off = ((chunk0 + q4) ^ rsw) << 4;
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(dsrow + off) = make_uint4(dk16[4 * q4], dk16[4 * q4 + 1], dk16[4 * q4 + 2], dk16[4 * q4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      ts.rec(14 + 6 * it);
      if (it > 0) dq_out(h_prev, i_prev);  // overlaps with S / dP of the next item and dV / dK / dQ of this one
      ts.rec(15 + 6 * it);
      h_prev = h_cur; i_prev = i_cur;
      h_cur = h_nxt; i_cur = i_nxt;
    }
    if (n_items > 0) {
      mbar_wait(dq_full, (n_items - 1) & 1);  // the last commit covers every MMA issued
      tc_fence_after();
      dq_out(h_prev, i_prev);
    }
    if (issuer) tma_store_wait<0>();
    // dK_j / dV_j: thread <-> key row (all MMAs have completed: dq_full of the last item was awaited above)
    const int key = k0 + r;
    const bool key_ok = key < p.lkv;
    {
      const int c = wg * 16;
      uint32_t a[16], bb[16];
      if (n_items > 0) {  // warp-uniform: the TMEM loads stay convergent for partially valid warps
        tmem_ld_32x16(tm_dk + lane_off + c, a);
        tmem_ld_32x16(tm_dv + lane_off + c, bb);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) { a[e] = 0; bb[e] = 0; }
      }
      if (key_ok && p.dkv_bf16) {
        bf16* base = p.dkv_bf16 + ((size_t)b * p.lkv + key) * p.dkv_ld + hk * HD + c;
#pragma unroll
        for (int e = 0; e < 16; e += 8) {
          uint4 vk, vv;
          vk.x = pack_bf16(__uint_as_float(a[e]), __uint_as_float(a[e + 1])); vk.y = pack_bf16(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
          vk.z = pack_bf16(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])); vk.w = pack_bf16(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7]));
          vv.x = pack_bf16(__uint_as_float(bb[e]), __uint_as_float(bb[e + 1])); vv.y = pack_bf16(__uint_as_float(bb[e + 2]), __uint_as_float(bb[e + 3]));
          vv.z = pack_bf16(__uint_as_float(bb[e + 4]), __uint_as_float(bb[e + 5])); vv.w = pack_bf16(__uint_as_float(bb[e + 6]), __uint_as_float(bb[e + 7]));
          *reinterpret_cast<uint4*>(base + p.dk_col0 + e) = vk;
          *reinterpret_cast<uint4*>(base + p.dv_col0 + e) = vv;
        }
      } else if (key_ok) {
        float* dkrow = p.dk + (((size_t)b * (p.hq / p.group) + hk) * p.lkv + key) * HD;
        float* dvrow = p.dv + (((size_t)b * (p.hq / p.group) + hk) * p.lkv + key) * HD;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
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

// out[b, i, :] = sum over the key blocks that touched query row i of partial[jb, b, i, :]   (fp32 accumulation)
__global__ void dq_reduce_kernel(const bf16* __restrict__ partial, int nkb, int batch, int lq, int C, int causal, bf16* __restrict__ out_bf16,
                                 long long ld_out, float* __restrict__ out_f32) {
  const int vpr = C >> 3;
  const size_t total = (size_t)batch * lq * vpr;
  const size_t slab = (size_t)batch * lq * C;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vpr);
    const size_t row = idx / vpr;
    const int i = (int)(row % lq);
    const int n = causal ? min(nkb, i / BQ + 1) : nkb;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bf16* src = partial + row * C + v * 8;
    for (int jb = 0; jb < n; ++jb) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + jb * slab);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
    if (out_bf16) {
      uint4 o;
      o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(out_bf16 + row * ld_out + v * 8) = o;
    } else {
      float4* o = reinterpret_cast<float4*>(out_f32 + row * C + v * 8);
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, BwdParams p, int batch, int hkv,
               void* workspace, size_t workspace_bytes, bf16* dq_bf16, long long dq_ld, float* dq_f32, cudaStream_t stream) {
  const int nkb = ceil_div(p.lkv, BKV);
  const int C = p.hq * HD;
  const size_t need = (size_t)nkb * batch * p.lq * C * 2;
  SLB_CHECK_ARG(workspace && workspace_bytes >= need, "attn_bwd: workspace of %zu bytes needed, got %zu", need, workspace_bytes);
  SLB_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "attn_bwd: workspace must be 16-byte aligned");
  CUtensorMap tdq;  // partial dQ slabs [nkb*B, lq, C] bf16: rows past lq are clipped per slab by the TMA store
  int rc = slb_make_tmap_3d(&tdq, workspace, (uint64_t)C, (uint64_t)p.lq, (uint64_t)nkb * batch, (uint64_t)C * 2, (uint64_t)p.lq * C * 2, HD, BQ, 1);
  if (rc) return rc;
  p.batch = batch;
  p.trace = slb_debug_trace_ptr();
#ifdef SLB_ABLATION  // timing experiments only (results are wrong by construction): not compiled into the product library
  static int variant = -1;
  if (variant < 0) { const char* ev = getenv("SLB_BWD_VARIANT"); variant = ev ? atoi(ev) : 0; }
  p.variant = variant;
#else
  p.variant = 0;
#endif
  static bool attr_set = false;
  if (!attr_set) {
    SLB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal));
    attr_set = true;
  }
  dim3 grid(nkb, hkv, batch);
  attn_bwd_kernel<<<grid, BWD_THREADS, kSmTotal, stream>>>(tq, tk, tv, tdo, tdq, p);
  SLB_LAUNCH_CHECK();
  const size_t vecs = (size_t)batch * p.lq * (C / 8);
  int rgrid = (int)((vecs + 255) / 256);
  const int cap = slb_num_sms() * 8;
  if (rgrid > cap) rgrid = cap;
  dq_reduce_kernel<<<rgrid, 256, 0, stream>>>((const bf16*)workspace, nkb, batch, p.lq, C, p.causal, dq_bf16, dq_ld, dq_f32);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

}  // namespace

extern "C" size_t slb_attn_bwd_workspace(int batch, int lq, int hq) {
  return (size_t)ceil_div(lq, BKV) * batch * lq * hq * HD * 2;
}

extern "C" int slb_attn_vit_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, void* workspace,
                                size_t workspace_bytes, int tiles, int n_tokens, int heads, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && n_tokens > 0 && heads > 0 && lse && delta && dqkv, "attn_vit_bwd: bad args");
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
  p.lse = lse; p.delta = delta; p.dk = nullptr; p.dv = nullptr;
  p.dkv_bf16 = (bf16*)dqkv; p.dkv_ld = 3 * C; p.dk_col0 = C; p.dv_col0 = 2 * C;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_bwd(tm, tm, tm, tdo, p, tiles, heads, workspace, workspace_bytes, (bf16*)dqkv, 3 * C, nullptr, (cudaStream_t)stream);
}

extern "C" int slb_attn_gqa_bwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid,
                                int key_valid_ld, const void* dout, const float* lse, const float* delta, float* dq, float* dk, float* dv,
                                void* workspace, size_t workspace_bytes, int batch, int lq, int lmax, int hq, int hkv, void* stream) {
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
  p.lse = lse; p.delta = delta; p.dk = dk; p.dv = dv; p.dkv_bf16 = nullptr;
  p.scale = 0.125f; p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_bwd(tq, tk, tv, tdo, p, batch, hkv, workspace, workspace_bytes, nullptr, 0, dq, (cudaStream_t)stream);
}
