// Flash-style attention for sm_100a.
//   attn_fwd_kernel: one CTA per (128-query block, head, batch).  S = Q K^T and O_j = P V_j run on tcgen05
//   (accumulators in TMEM), K/V tiles stream through a 2-stage TMA ring, online softmax in registers of four
//   softmax warps (thread <-> query row <-> TMEM lane).  Two CTAs are co-resident per SM (112 KB smem,
//   256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//   Serves both the bidirectional InternViT attention (packed qkv, N=1025, 16 heads) and the causal GQA
//   Qwen2 prefill (14 q heads / 2 kv heads over the KV cache, key-padding mask).
//   attn_small_kernel: decode / 30-query append (Lq <= 32): CUDA-core kernel, one block per (b, head, query).
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int ATT_THREADS = 192;
constexpr int kMaxKvBlocks = 32;            // key-validity bitmasks in static smem (512 B: keeps two CTAs per SM): up to 4096 keys
constexpr int kQBytes = BQ * HD * 2;        // 16 KB
constexpr int kKVBytes = BKV * HD * 2;      // 16 KB
constexpr int kPBytes = BQ * BKV * 2;       // 32 KB (two K-major halves of 64 columns)
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kQBytes;            // 2 stages
constexpr int kSmemV = kSmemK + 2 * kKVBytes;       // 2 stages
constexpr int kSmemP = kSmemV + 2 * kKVBytes;
constexpr int kSmemBar = kSmemP + kPBytes;
constexpr int kSmemMask = kSmemBar + 128;           // key-validity bitmasks, 16 B per key block
constexpr int kSmemTotal = kSmemMask + kMaxKvBlocks * 16;
constexpr uint32_t kTmemCols = 256;  // S: [0,128)  O: [128,192)

struct AttnParams {
  int lq, lkv;           // valid query rows per batch item, valid keys per batch item
  int past;              // absolute position of query row 0 (causal offset)
  int causal;
  int hq, group;         // q heads, q heads per kv head
  int q_col0, k_col0, v_col0;
  int kv_head_col_stride, kv_batch_stride, kv_head_batch_stride;
  const uint8_t* key_valid;  // [B, key_valid_ld] or null
  int key_valid_ld;
  bf16* out; long long ldo;  // out[(b*lq + i) * ldo + h*64 + d]
  float* lse;                // [B, hq, lq] or null
  float scale_log2;          // softmax scale * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * BQ;
  const int hk = h / p.group;
  // number of kv blocks this q block needs
  int kv_end = p.lkv;
  if (p.causal) kv_end = min(kv_end, p.past + q0 + BQ);
  const int nkv = (kv_end + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023) != 0) __trap();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(smem + kSmemQ, &tmap_q, q_full, p.q_col0 + h * HD, q0, b);
      const int kc = p.k_col0 + hk * p.kv_head_col_stride, vc = p.v_col0 + hk * p.kv_head_col_stride;
      const int kb = b * p.kv_batch_stride + hk * p.kv_head_batch_stride;
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * kKVBytes);
        tma_load_3d(smem + kSmemK + s * kKVBytes, &tmap_k, &kv_full[s], kc, j * BKV, kb);
        tma_load_3d(smem + kSmemV + s * kKVBytes, &tmap_v, &kv_full[s], vc, j * BKV, kb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);  // B = V is MN-major (d contiguous)
      const uint32_t sq = smem_u32(smem + kSmemQ), sp = smem_u32(smem + kSmemP);
      const uint64_t dq = umma_desc_kmajor_sw128(sq);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(smem + kSmemK + s * kKVBytes));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      if (nkv > 0) issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(smem + kSmemV + s * kKVBytes), kKVBytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t dp = umma_desc_kmajor_sw128(sp + (k >> 2) * (BQ * 128)) + 2 * (k & 3);
          tc_mma_bf16(tmem_o, dp, dv + (uint64_t)k * (16 * 128 >> 4), idesc_o, k != 0);
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[s]);
        if (j + 1 < nkv) issue_s(j + 1);
      }
    }
  } else {
    // ---------------- softmax warps: thread <-> query row ----------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;            // row within the q block
    const int row = q0 + r;                    // query index within the batch item
    const int qpos = p.past + row;             // absolute position (causal)
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint8_t* kvalid = p.key_valid ? p.key_valid + (size_t)b * p.key_valid_ld : nullptr;
    // per key block, a 128-bit validity mask (inside the sequence and not padding), built once per CTA: the softmax
    // loops then test a register bit instead of loading key_valid per element, and unmasked blocks take the fast path
    uint32_t* kmask = reinterpret_cast<uint32_t*>(smem + kSmemMask);
    for (int j = 0; j < nkv; ++j) {
      const int col = j * BKV + r;
      const bool ok = col < p.lkv && (!kvalid || kvalid[col] != 0);
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) kmask[j * 4 + quad] = bal;
    }
    named_bar_sync(2, 128);
    float o_acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = smem + kSmemP + r * 128;
    const int rsw = r & 7;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int col_base = j * BKV;
      const uint32_t km0 = kmask[j * 4], km1 = kmask[j * 4 + 1], km2 = kmask[j * 4 + 2], km3 = kmask[j * 4 + 3];
      const bool need_mask = ((km0 & km1 & km2 & km3) != 0xffffffffu) || (p.causal && col_base + BKV - 1 > p.past + q0);
      // pass 1: row max
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t sr[32];
        tmem_ld_32x32(tmem_s + lane_off + c, sr);
        tmem_ld_wait();
        if (need_mask) {
          const uint32_t km = c == 0 ? km0 : (c == 32 ? km1 : (c == 64 ? km2 : km3));
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = ((km >> i) & 1u) && (!p.causal || col_base + c + i <= qpos);
            m_blk = fmaxf(m_blk, ok ? __uint_as_float(sr[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(sr[i]));
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f((m_run - m_use) * p.scale_log2);  // m_run = -inf -> 0
      const float moff = m_use * p.scale_log2;
      float l_blk = 0.f;
      // pass 2: p = exp2(s*scale - m*scale) -> bf16 -> smem (K-major, 128B swizzle)
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t sr[32];
        tmem_ld_32x32(tmem_s + lane_off + c, sr);
        tmem_ld_wait();
        float pv[32];
        const uint32_t km = c == 0 ? km0 : (c == 32 ? km1 : (c == 64 ? km2 : km3));
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = ex2_approx(fmaf(__uint_as_float(sr[i]), p.scale_log2, -moff));
          if (need_mask) e = (((km >> i) & 1u) && (!p.causal || col_base + c + i <= qpos)) ? e : 0.f;
          pv[i] = e;
        }
        // round to bf16 first so the row sum matches what the PV MMA sees
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
          float2 back = unpack_bf16(pk[i]);
          l_blk += back.x + back.y;
        }
        uint8_t* half = prow + (c >> 6) * (BQ * 128);
        const int chunk0 = (c & 63) >> 3;  // 16-byte chunk index of column c within the 64-column half
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(half + (((chunk0 + q4) ^ rsw) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      // O_j from TMEM
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t orr[32];
        tmem_ld_32x32(tmem_o + lane_off + c, orr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(orr[i]);
      }
      tc_fence_before();
    }
    if (row < p.lq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* o = p.out + ((size_t)b * p.lq + row) * p.ldo + h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 8) {
        uint4 u;
        u.x = pack_bf16(o_acc[d] * inv, o_acc[d + 1] * inv);
        u.y = pack_bf16(o_acc[d + 2] * inv, o_acc[d + 3] * inv);
        u.z = pack_bf16(o_acc[d + 4] * inv, o_acc[d + 5] * inv);
        u.w = pack_bf16(o_acc[d + 6] * inv, o_acc[d + 7] * inv);
        *reinterpret_cast<uint4*>(o + d) = u;
      }
      if (p.lse) {
        const float mm = (m_run == -INFINITY) ? 0.f : m_run;
        p.lse[((size_t)b * p.hq + h) * p.lq + row] = (l_run > 0.f) ? (mm * p.scale_log2 + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// small-Lq attention (decode / query append): block per (query, head, batch), 256 threads.
//   phase 1: one thread per key computes q.k (16-byte loads of the key row)           -> scores in smem, block max
//   phase 2: exp and block sum
//   phase 3: each warp takes a contiguous chunk of keys; lanes own two output dims, so a key's V row is one coalesced
//            128-byte warp load; 4 keys in flight; the 8 partial outputs are combined through smem
// ------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 1024, kSmallWarps = kSmallThreads / 32;  // one key per thread for the 545..700-key agent prompts
// ROPE (decode step, lq == 1): q still holds the UN-rotated projection output; the kernel rotates q on the fly (fp32), rotates the new
// key, writes it and the new value row into the caches (one block per kv head does the writes) and takes the newest key / value from
// shared memory instead of the cache - the separate RoPE + KV-write launch of every layer disappears (24 launches per token).
template <bool ROPE>
__global__ void __launch_bounds__(kSmallThreads)
attn_small_kernel(const bf16* __restrict__ q, long long ldq, const bf16* kc, const bf16* vc,
                  const uint8_t* __restrict__ key_valid, int key_valid_ld, bf16* __restrict__ out, long long ldo, int lq, int past,
                  int lmax, int hq, int hkv, float scale, const int* __restrict__ past_dev, float log2_theta) {
  extern __shared__ float sm[];
  __shared__ float rcs[32], rsn[32], knew[64], vnew[64];
  pdl_trigger();
  pdl_wait();   // q / the newest K, V rows are written by the kernel right before us
  if (past_dev) past = *past_dev;  // position counter kept on the device: the same CUDA graph serves every decode step
  float* qs = sm;                        // 64
  float* red = sm + 64;                  // 2 * kSmallWarps
  float* part = sm + 64 + 2 * kSmallWarps;                 // kSmallWarps * 64
  float* sc = part + kSmallWarps * 64;   // lmax
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (hq / hkv);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkeys = past + i + 1;
  const int nk_c = ROPE ? nkeys - 1 : nkeys;   // keys taken from the cache
  const bf16* qr = q + ((size_t)b * lq + i) * ldq + h * 64;
  const bf16* kbase = kc + ((size_t)b * hkv + hk) * lmax * 64;
  const bf16* vbase = vc + ((size_t)b * hkv + hk) * lmax * 64;
  if (ROPE) {
    if (tid < 32) {   // same arithmetic as rope_kv_write_kernel
      const float inv_freq = exp2f(-(float)(2 * tid) / 64.0f * log2_theta);
      sincosf((float)(past + i) * inv_freq, &rsn[tid], &rcs[tid]);
    }
    __syncthreads();
    const bf16* krow = q + ((size_t)b * lq + i) * ldq + (hq + hk) * 64;
    const bf16* vrow = krow + hkv * 64;
    const bool writer = (h % (hq / hkv)) == 0;
    if (tid < 64) {
      const int d = tid & 31;
      const float x0 = __bfloat162float(qr[d]), x1 = __bfloat162float(qr[d + 32]);
      qs[tid] = (tid < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]) * scale;
    } else if (tid < 128) {
      const int t = tid - 64, d = t & 31;
      const float x0 = __bfloat162float(krow[d]), x1 = __bfloat162float(krow[d + 32]);
      const bf16 kr16 = __float2bfloat16(t < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]);
      knew[t] = __bfloat162float(kr16);   // the value later steps will read back from the cache
      if (writer) const_cast<bf16*>(kbase)[(size_t)(nkeys - 1) * 64 + t] = kr16;
    } else if (tid < 192) {
      const int t = tid - 128;
      const bf16 v16 = vrow[t];
      vnew[t] = __bfloat162float(v16);
      if (writer) const_cast<bf16*>(vbase)[(size_t)(nkeys - 1) * 64 + t] = v16;
    }
  } else {
    if (tid < 64) qs[tid] = __bfloat162float(qr[tid]) * scale;
  }
  __syncthreads();
  const uint8_t* kv_ok = key_valid ? key_valid + (size_t)b * key_valid_ld : nullptr;
  float mx = -INFINITY;
  for (int j = tid; j < nk_c; j += kSmallThreads) {
    float s = -INFINITY;
    if (!kv_ok || kv_ok[j]) {
      s = 0.f;
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)j * 64);
      uint4 u[8];
#pragma unroll
      for (int v8 = 0; v8 < 8; ++v8) u[v8] = kr[v8];
#pragma unroll
      for (int v8 = 0; v8 < 8; ++v8) {
        float2 a = unpack_bf16(u[v8].x), bb = unpack_bf16(u[v8].y), c = unpack_bf16(u[v8].z), d = unpack_bf16(u[v8].w);
        const float* qq = qs + v8 * 8;
        s += a.x * qq[0] + a.y * qq[1] + bb.x * qq[2] + bb.y * qq[3] + c.x * qq[4] + c.y * qq[5] + d.x * qq[6] + d.y * qq[7];
      }
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  if (ROPE && tid == 0) {   // the newest key (always valid) comes from shared memory
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 64; ++e) s += knew[e] * qs[e];
    sc[nkeys - 1] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < kSmallWarps; ++w) mx = fmaxf(mx, red[w]);
  if (mx == -INFINITY) mx = 0.f;
  float sum = 0.f;
  for (int j = tid; j < nkeys; j += kSmallThreads) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[kSmallWarps + warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kSmallWarps; ++w) sum += red[kSmallWarps + w];
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  const int chunk = (nk_c + kSmallWarps - 1) / kSmallWarps;
  const int j0 = warp * chunk, j1 = min(nk_c, j0 + chunk);
  float a0 = 0.f, a1 = 0.f;
  const uint32_t* v32 = reinterpret_cast<const uint32_t*>(vbase) + lane;  // dims 2*lane, 2*lane+1 of every row (32 words per row)
  int j = j0;
  for (; j + 16 <= j1; j += 16) {  // 16 independent row loads in flight per lane
    uint32_t wv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) wv[u] = v32[(size_t)(j + u) * 32];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float2 f = unpack_bf16(wv[u]);
      const float pj = sc[j + u];
      a0 += pj * f.x;
      a1 += pj * f.y;
    }
  }
  for (; j + 4 <= j1; j += 4) {
    uint32_t w0 = v32[(size_t)j * 32], w1 = v32[(size_t)(j + 1) * 32], w2 = v32[(size_t)(j + 2) * 32], w3 = v32[(size_t)(j + 3) * 32];
    const float p0 = sc[j], p1 = sc[j + 1], p2 = sc[j + 2], p3 = sc[j + 3];
    float2 f0 = unpack_bf16(w0), f1 = unpack_bf16(w1), f2 = unpack_bf16(w2), f3 = unpack_bf16(w3);
    a0 += p0 * f0.x + p1 * f1.x + p2 * f2.x + p3 * f3.x;
    a1 += p0 * f0.y + p1 * f1.y + p2 * f2.y + p3 * f3.y;
  }
  for (; j < j1; ++j) {
    const float2 f = unpack_bf16(v32[(size_t)j * 32]);
    a0 += sc[j] * f.x;
    a1 += sc[j] * f.y;
  }
  part[warp * 64 + 2 * lane] = a0;
  part[warp * 64 + 2 * lane + 1] = a1;
  __syncthreads();
  if (tid < 64) {
    float o = ROPE ? sc[nkeys - 1] * vnew[tid] : 0.f;
#pragma unroll
    for (int w = 0; w < kSmallWarps; ++w) o += part[w * 64 + tid];
    out[((size_t)b * lq + i) * ldo + h * 64 + tid] = __float2bfloat16(o * inv);
  }
}

// Batched decode: one block per (query, kv head, batch) serving all q heads of the GQA group, so a key / value row is
// read once for the 7 heads that share it (the per-head kernel above re-reads it 7 times: 69 MB per layer at batch 32).
constexpr int kGrpThreads = 512, kGrpWarps = kGrpThreads / 32, kMaxGroup = 8;
template <bool ROPE>   // see attn_small_kernel: fused RoPE + KV write for the decode step (lq == 1)
__global__ void __launch_bounds__(kGrpThreads)
attn_decode_group_kernel(const bf16* __restrict__ q, long long ldq, const bf16* kc, const bf16* vc,
                         const uint8_t* __restrict__ key_valid, int key_valid_ld, bf16* __restrict__ out, long long ldo, int lq, int past,
                         int lmax, int hq, int hkv, float scale, const int* __restrict__ past_dev, float log2_theta) {
  extern __shared__ float sm[];
  __shared__ float rcs[32], rsn[32], knew[64], vnew[64];
  pdl_trigger();
  pdl_wait();
  if (past_dev) past = *past_dev;
  const int G = hq / hkv;
  float* qs = sm;                                   // [G][64]
  float* red = qs + kMaxGroup * 64;                 // [2][kGrpWarps][kMaxGroup]
  float* part = red + 2 * kGrpWarps * kMaxGroup;    // [kGrpWarps][G][64]
  float* sc = part + kGrpWarps * kMaxGroup * 64;    // [G][lmax]
  const int i = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkeys = past + i + 1;
  const int nk_c = ROPE ? nkeys - 1 : nkeys;   // keys taken from the cache
  const bf16* kbase = kc + ((size_t)b * hkv + hk) * lmax * 64;
  const bf16* vbase = vc + ((size_t)b * hkv + hk) * lmax * 64;
  if (ROPE) {
    if (tid < 32) {
      const float inv_freq = exp2f(-(float)(2 * tid) / 64.0f * log2_theta);
      sincosf((float)(past + i) * inv_freq, &rsn[tid], &rcs[tid]);
    }
    __syncthreads();
    const bf16* row = q + ((size_t)b * lq + i) * ldq;
    for (int t = tid; t < G * 64; t += kGrpThreads) {
      const int hh = t / 64, e = t & 63, d = e & 31;
      const bf16* qr = row + (hk * G + hh) * 64;
      const float x0 = __bfloat162float(qr[d]), x1 = __bfloat162float(qr[d + 32]);
      qs[t] = (e < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]) * scale;
    }
    if (tid < 64) {
      const bf16* krow = row + (hq + hk) * 64;
      const int d = tid & 31;
      const float x0 = __bfloat162float(krow[d]), x1 = __bfloat162float(krow[d + 32]);
      const bf16 kr16 = __float2bfloat16(tid < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]);
      knew[tid] = __bfloat162float(kr16);
      const_cast<bf16*>(kbase)[(size_t)(nkeys - 1) * 64 + tid] = kr16;
    } else if (tid < 128) {
      const int t = tid - 64;
      const bf16 v16 = row[(hq + hkv + hk) * 64 + t];
      vnew[t] = __bfloat162float(v16);
      const_cast<bf16*>(vbase)[(size_t)(nkeys - 1) * 64 + t] = v16;
    }
  } else {
    for (int t = tid; t < G * 64; t += kGrpThreads)
      qs[t] = __bfloat162float(q[((size_t)b * lq + i) * ldq + (hk * G + t / 64) * 64 + (t & 63)]) * scale;
  }
  __syncthreads();
  const uint8_t* kv_ok = key_valid ? key_valid + (size_t)b * key_valid_ld : nullptr;
  float mx[kMaxGroup];
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) mx[h] = -INFINITY;
  if (ROPE && tid < G) {   // the newest key (always valid) comes from shared memory: thread h scores head h
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 64; ++e) s += knew[e] * qs[tid * 64 + e];
    sc[(size_t)tid * lmax + nkeys - 1] = s;
#pragma unroll
    for (int h = 0; h < kMaxGroup; ++h) mx[h] = (h == tid) ? s : mx[h];
  }
  for (int j = tid; j < nk_c; j += kGrpThreads) {
    const bool ok = !kv_ok || kv_ok[j];
    float kf[64];
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)j * 64);
#pragma unroll
    for (int v8 = 0; v8 < 8; ++v8) {
      const uint4 u = kr[v8];
      const float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      kf[v8 * 8] = a.x; kf[v8 * 8 + 1] = a.y; kf[v8 * 8 + 2] = bb.x; kf[v8 * 8 + 3] = bb.y;
      kf[v8 * 8 + 4] = c.x; kf[v8 * 8 + 5] = c.y; kf[v8 * 8 + 6] = d.x; kf[v8 * 8 + 7] = d.y;
    }
#pragma unroll
    for (int h = 0; h < kMaxGroup; ++h) {
      if (h < G) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 64; ++e) s += kf[e] * qs[h * 64 + e];
        s = ok ? s : -INFINITY;
        sc[(size_t)h * lmax + j] = s;
        mx[h] = fmaxf(mx[h], s);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) {
    mx[h] = warp_max(mx[h]);
    if (lane == 0) red[warp * kMaxGroup + h] = mx[h];
  }
  __syncthreads();
  float sum[kMaxGroup];
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) {
    float m = red[h];
    for (int w = 1; w < kGrpWarps; ++w) m = fmaxf(m, red[w * kMaxGroup + h]);
    mx[h] = (m == -INFINITY) ? 0.f : m;
    sum[h] = 0.f;
  }
  for (int j = tid; j < nkeys; j += kGrpThreads) {
#pragma unroll
    for (int h = 0; h < kMaxGroup; ++h) {
      if (h < G) {
        const float e = __expf(sc[(size_t)h * lmax + j] - mx[h]);
        sc[(size_t)h * lmax + j] = e;
        sum[h] += e;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) {
    sum[h] = warp_sum(sum[h]);
    if (lane == 0) red[(kGrpWarps + warp) * kMaxGroup + h] = sum[h];
  }
  __syncthreads();
  // weighted V: each warp a contiguous chunk of keys, lanes own dims (2 lane, 2 lane + 1), all heads of the group at once
  const int chunk = (nk_c + kGrpWarps - 1) / kGrpWarps;
  const int j0 = warp * chunk, j1 = min(nk_c, j0 + chunk);
  const uint32_t* v32 = reinterpret_cast<const uint32_t*>(vbase) + lane;
  float a0[kMaxGroup], a1[kMaxGroup];
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) { a0[h] = 0.f; a1[h] = 0.f; }
  int j = j0;
  for (; j + 8 <= j1; j += 8) {
    uint32_t wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = v32[(size_t)(j + u) * 32];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float2 f = unpack_bf16(wv[u]);
#pragma unroll
      for (int h = 0; h < kMaxGroup; ++h) {
        if (h < G) {
          const float pj = sc[(size_t)h * lmax + j + u];
          a0[h] += pj * f.x;
          a1[h] += pj * f.y;
        }
      }
    }
  }
  for (; j < j1; ++j) {
    const float2 f = unpack_bf16(v32[(size_t)j * 32]);
#pragma unroll
    for (int h = 0; h < kMaxGroup; ++h) {
      if (h < G) {
        const float pj = sc[(size_t)h * lmax + j];
        a0[h] += pj * f.x;
        a1[h] += pj * f.y;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxGroup; ++h) {
    if (h < G) {
      part[(warp * kMaxGroup + h) * 64 + 2 * lane] = a0[h];
      part[(warp * kMaxGroup + h) * 64 + 2 * lane + 1] = a1[h];
    }
  }
  __syncthreads();
  for (int t = tid; t < G * 64; t += kGrpThreads) {
    const int h = t / 64, d = t & 63;
    float o = ROPE ? sc[(size_t)h * lmax + nkeys - 1] * vnew[d] : 0.f, sm_ = 0.f;
    for (int w = 0; w < kGrpWarps; ++w) { o += part[(w * kMaxGroup + h) * 64 + d]; sm_ += red[(kGrpWarps + w) * kMaxGroup + h]; }
    out[((size_t)b * lq + i) * ldo + (hk * G + h) * 64 + d] = __float2bfloat16(sm_ > 0.f ? o / sm_ : 0.f);
  }
}

int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int batch,
                cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SLB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    // two CTAs per SM need the full 228 KB carve-out (2 x (112.6 KB + 1 KB reserved)); the default picks 132 KB = one CTA
    SLB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.lq, BQ), p.hq, batch);
  attn_fwd_kernel<<<grid, ATT_THREADS, kSmemTotal, stream>>>(tq, tk, tv, p);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

}  // namespace

int slb_attn_vit2_try(const void* qkv, void* out, float* lse, int tiles, int n_tokens, int heads, cudaStream_t stream, int* rc_out);

extern "C" int slb_attn_vit_fwd(const void* qkv, void* out, float* lse, int tiles, int n_tokens, int heads, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && n_tokens > 0 && heads > 0, "attn_vit: bad shape");
  {
    int rc2 = SLB_OK;
    if (slb_attn_vit2_try(qkv, out, lse, tiles, n_tokens, heads, (cudaStream_t)stream, &rc2)) return rc2;
  }
  const int C = heads * HD;
  CUtensorMap tm;
  int rc = slb_make_tmap_3d(&tm, qkv, (uint64_t)3 * C, (uint64_t)n_tokens, (uint64_t)tiles, (uint64_t)3 * C * 2,
                            (uint64_t)n_tokens * 3 * C * 2, HD, BQ, 1);
  if (rc) return rc;
  AttnParams p{};
  p.lq = n_tokens; p.lkv = n_tokens; p.past = 0; p.causal = 0;
  p.hq = heads; p.group = 1;
  p.q_col0 = 0; p.k_col0 = C; p.v_col0 = 2 * C;
  p.kv_head_col_stride = HD; p.kv_batch_stride = 1; p.kv_head_batch_stride = 0;
  p.key_valid = nullptr; p.key_valid_ld = 0;
  p.out = (bf16*)out; p.ldo = C; p.lse = lse;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_attn(tm, tm, tm, p, tiles, (cudaStream_t)stream);
}

size_t slb_attn_gqa2_workspace(int batch, int lq);  // attention_gqa.cu
int slb_attn_gqa2_try(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid, int key_valid_ld,
                      void* out, float* lse, int batch, int lq, int past, int lmax, int hq, int hkv, uint32_t* kmask_ws, cudaStream_t stream,
                      int* rc_out);

extern "C" size_t slb_attn_gqa_fwd_workspace(int batch, int lq) { return slb_attn_gqa2_workspace(batch, lq); }

// decode / query-append attention over the KV cache (lq <= 32, no lse); rope_log2_theta > 0 (lq == 1 only): fused RoPE + KV write
static int launch_decode_attn(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid, int key_valid_ld,
                              void* out, int batch, int lq, int past, const int32_t* past_dev, int lmax, int hq, int hkv, float rope_log2_theta,
                              void* stream) {
  const float scale = 0.125f;
  const bool rope = rope_log2_theta > 0.f;
  if (batch >= 4 && hq / hkv <= kMaxGroup) {
    // batched decode / query append: one block per kv head serves its whole GQA group
    const size_t smem = ((size_t)kMaxGroup * 64 + 2 * kGrpWarps * kMaxGroup + (size_t)kGrpWarps * kMaxGroup * 64 + (size_t)(hq / hkv) * lmax) * sizeof(float);
    SLB_CHECK_ARG(smem <= 200 * 1024, "attn_gqa: lmax=%d too long for the grouped decode kernel", lmax);
    static size_t smem_set[2] = {0, 0};
    if (smem > smem_set[rope]) {
      if (rope) SLB_CUDA(cudaFuncSetAttribute(attn_decode_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else SLB_CUDA(cudaFuncSetAttribute(attn_decode_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[rope] = smem;
    }
    auto kern = rope ? attn_decode_group_kernel<true> : attn_decode_group_kernel<false>;
    SLB_CUDA(slb_launch_pdl(true, kern, dim3(lq, hkv, batch), dim3(kGrpThreads), smem, (cudaStream_t)stream, (const bf16*)q, (long long)ldq,
                            (const bf16*)kcache, (const bf16*)vcache, key_valid, key_valid_ld, (bf16*)out, (long long)hq * HD, lq, past, lmax, hq,
                            hkv, scale, (const int*)past_dev, rope_log2_theta));
    return SLB_OK;
  }
  dim3 grid(lq, hq, batch);
  const size_t smem = (64 + 2 * kSmallWarps + kSmallWarps * 64 + (size_t)lmax) * sizeof(float);
  SLB_CHECK_ARG(smem <= 46 * 1024, "attn_gqa: lmax=%d too long for the small-Lq kernel", lmax);
  auto kern = rope ? attn_small_kernel<true> : attn_small_kernel<false>;
  SLB_CUDA(slb_launch_pdl(true, kern, grid, dim3(kSmallThreads), smem, (cudaStream_t)stream, (const bf16*)q, (long long)ldq, (const bf16*)kcache,
                          (const bf16*)vcache, key_valid, key_valid_ld, (bf16*)out, (long long)hq * HD, lq, past, lmax, hq, hkv, scale,
                          (const int*)past_dev, rope_log2_theta));
  return SLB_OK;
}

extern "C" int slb_attn_decode_rope(void* qkv, int64_t ldq, void* kcache, void* vcache, const uint8_t* key_valid, int key_valid_ld, void* out,
                                    int batch, int past, const int32_t* past_dev, int lmax, int hq, int hkv, float theta, void* stream) {
  SLB_CHECK_ARG(qkv && kcache && vcache && out && batch > 0 && past >= 0 && past + 1 <= lmax && hkv > 0 && hq % hkv == 0, "attn_decode_rope: bad shape");
  SLB_CHECK_ARG((ldq % 8) == 0 && ldq >= (int64_t)(hq + 2 * hkv) * HD && theta > 1.f, "attn_decode_rope: qkv rows must hold q | k | v (ldq=%lld), theta > 1",
                (long long)ldq);
  return launch_decode_attn(qkv, ldq, kcache, vcache, key_valid, key_valid_ld, out, batch, 1, past, past_dev, lmax, hq, hkv, log2f(theta), stream);
}

extern "C" int slb_attn_gqa_fwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid,
                                int key_valid_ld, void* out, float* lse, int batch, int lq, int past, const int32_t* past_dev, int lmax,
                                int hq, int hkv, void* workspace, size_t workspace_bytes, void* stream) {
  SLB_CHECK_ARG(batch > 0 && lq > 0 && past >= 0 && past + lq <= lmax && hq % hkv == 0, "attn_gqa: bad shape");
  SLB_CHECK_ARG((ldq % 8) == 0, "attn_gqa: ldq must be a multiple of 8");
  SLB_CHECK_ARG(past + lq <= kMaxKvBlocks * BKV, "attn_gqa: at most %d keys", kMaxKvBlocks * BKV);
  if (lq <= 32 && lse == nullptr)
    return launch_decode_attn(q, ldq, kcache, vcache, key_valid, key_valid_ld, out, batch, lq, past, past_dev, lmax, hq, hkv, 0.f, stream);
  SLB_CHECK_ARG(past_dev == nullptr, "attn_gqa: a device-side position is only supported for chunks of <= 32 queries without lse");
  {  // prefill / teacher-forced pass: the persistent head-pair kernel (attention_gqa.cu)
    int rc2 = SLB_OK;
    uint32_t* ws = (workspace && workspace_bytes >= slb_attn_gqa2_workspace(batch, lq)) ? (uint32_t*)workspace : nullptr;
    if (slb_attn_gqa2_try(q, ldq, kcache, vcache, key_valid, key_valid_ld, out, lse, batch, lq, past, lmax, hq, hkv, ws, (cudaStream_t)stream, &rc2))
      return rc2;
  }
  CUtensorMap tq, tk, tv;
  int rc = slb_make_tmap_3d(&tq, q, (uint64_t)ldq, (uint64_t)lq, (uint64_t)batch, (uint64_t)ldq * 2, (uint64_t)lq * ldq * 2, HD, BQ, 1);
  if (rc) return rc;
  rc = slb_make_tmap_3d(&tk, kcache, HD, (uint64_t)lmax, (uint64_t)batch * hkv, HD * 2, (uint64_t)lmax * HD * 2, HD, BKV, 1);
  if (rc) return rc;
  rc = slb_make_tmap_3d(&tv, vcache, HD, (uint64_t)lmax, (uint64_t)batch * hkv, HD * 2, (uint64_t)lmax * HD * 2, HD, BKV, 1);
  if (rc) return rc;
  AttnParams p{};
  p.lq = lq; p.lkv = past + lq; p.past = past; p.causal = 1;
  p.hq = hq; p.group = hq / hkv;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0;
  p.kv_head_col_stride = 0; p.kv_batch_stride = hkv; p.kv_head_batch_stride = 1;
  p.key_valid = key_valid; p.key_valid_ld = key_valid_ld;
  p.out = (bf16*)out; p.ldo = (long long)hq * HD; p.lse = lse;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  return launch_attn(tq, tk, tv, p, batch, (cudaStream_t)stream);
}
