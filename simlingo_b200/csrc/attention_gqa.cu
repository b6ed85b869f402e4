// Causal GQA attention forward for the Qwen2 prefill / teacher-forced pass (Lq >= 128, past = 0), second generation:
// the InternViT kernel's architecture (attention_vit.cu) applied to 14 query heads sharing 2 KV heads.
//
//   Persistent grid (one CTA per SM, 12 warps, setmaxnreg re-balanced); work item = (batch b, kv head g, 128-query block qb,
//   pair of query heads of the GQA group).  The two heads of a pair are the two "halves" of the CTA: they read the SAME
//   K_j / V_j stream (one TMA load serves two S MMAs and two PV MMAs; the 7 heads of a group re-read a K/V tile from L2
//   4 times instead of 7), and while one half's softmax runs the other half's MMAs are in flight:
//     warp 0 / 3  TMA producers: Q (two 128 x 64 tiles, one per head) per item + the K_j stream / the V_j stream, 4-deep rings
//     warps 1,2   MMA issuers, one per half: S = Q K_j^T (128 x 128 x 64), O += P V_j (128 x 64 x 128), O resident in TMEM
//     warps 4-7   softmax warpgroup of half 0, warps 8-11 of half 1: one TMEM pass (the row's 128 scores in registers),
//                 lazy running maximum (rescale O / l only when the maximum grows by > 2^8), exp2 on packed fp32x2 FMA +
//                 MUFU, P -> shared memory in the K-major 128B-swizzled UMMA layout
//   Causality: a query block only visits key blocks 0..qb (block skipping); the mask is applied in registers on the diagonal
//   block only.  Key padding (left-padded prompts): 128-bit validity words per key block, precomputed by a tiny kernel,
//   consulted only when a block is not all-valid.  Items are ordered longest (qb large) first for load balance.
//   An odd group size (7) leaves one single-head item per (b, g, qb): its second half idles through the rings.
//   Fully masked rows (padding queries) produce out = 0, lse = -inf, as the first-generation kernel did.
#include "common.cuh"
#include "attn_common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int HD = 64, BQ = 128, BKV = 128, NSTAGE = 4;
constexpr int G2_THREADS = 384;
constexpr int kTile = BQ * HD * 2;            // 16 KB
constexpr int kSmQ = 0;                       // 2 tiles (one per head of the pair)
constexpr int kSmK = kSmQ + 2 * kTile;
constexpr int kSmV = kSmK + NSTAGE * kTile;
constexpr int kSmP = kSmV + NSTAGE * kTile;   // [half]: 2 x 32 KB
constexpr int kPBuf = BQ * BKV * 2;
constexpr int kSmBar = kSmP + 2 * kPBuf;
constexpr int kSmTotal = kSmBar + 256;
constexpr float kLazyThreshold = 8.0f;

struct Gqa2Params {
  bf16* out;               // [B * L, hq * 64]
  float* lse;              // [B, hq, L] or null
  const uint32_t* kmask;   // [B, nqb, 4] validity words per key block, or null (all keys valid)
  int L, lmax, batch, hq, hkv, group, npairs, nqb;
  float scale_log2;
};

// validity words: bit i of word w of block j of batch b = key j*128 + 32*w + i is inside the sequence and not padding
__global__ void gqa2_kmask_kernel(const uint8_t* __restrict__ key_valid, int ld, int L, int nqb, uint32_t* __restrict__ out) {
  const int b = blockIdx.y, j = blockIdx.x, t = threadIdx.x;   // 128 threads
  const int col = j * BKV + t;
  const bool ok = col < L && key_valid[(size_t)b * ld + col] != 0;
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  if ((t & 31) == 0) out[((size_t)b * nqb + j) * 4 + (t >> 5)] = bal;
}

__global__ void __launch_bounds__(G2_THREADS, 1)
attn_gqa2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, Gqa2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // NSTAGE
  uint64_t* k_empty = k_full + NSTAGE;           // NSTAGE (one arrival per half)
  uint64_t* v_full = k_empty + NSTAGE;           // NSTAGE
  uint64_t* v_empty = v_full + NSTAGE;           // NSTAGE (one arrival per half)
  uint64_t* s_full = v_empty + NSTAGE;           // [half]
  uint64_t* p_full = s_full + 2;                 // [half]
  uint64_t* s_free = p_full + 2;                 // [half]
  uint64_t* o_done = s_free + 2;                 // [half][cnt & 1]
  uint64_t* q_empty = o_done + 4;                // one arrival per half
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int per_qb = p.batch * p.hkv * p.npairs;
  const int n_items = p.nqb * per_qb;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023) != 0) __trap();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); mbar_init(&s_free[g], 4);
      mbar_init(&o_done[2 * g], 1); mbar_init(&o_done[2 * g + 1], 1);
    }
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
  pdl_trigger();
  pdl_wait();   // q / K / V come from the RoPE kernel right before us

  // item k of this CTA -> (batch, kv head, query block, first head of the pair, number of heads in the pair); heaviest first
  auto decode = [&](int k, int& b, int& g, int& qb, int& h0, int& nh) {
    const int w = (int)blockIdx.x + k * (int)gridDim.x;
    qb = p.nqb - 1 - w / per_qb;
    int r = w % per_qb;
    const int pr = r % p.npairs;
    r /= p.npairs;
    g = r % p.hkv;
    b = r / p.hkv;
    h0 = g * p.group + 2 * pr;
    nh = (2 * pr + 1 < p.group) ? 2 : 1;
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      int blk = 0;
      for (int k = 0; k < my_items; ++k) {
        int b, g, qb, h0, nh;
        decode(k, b, g, qb, h0, nh);
        if (k > 0) mbar_wait(q_empty, (k - 1) & 1);
        if (elect_one_sync()) {
          mbar_expect_tx(q_full, nh * kTile);
          tma_load_3d(smem + kSmQ, &tmap_q, q_full, h0 * HD, qb * BQ, b);
          if (nh == 2) tma_load_3d(smem + kSmQ + kTile, &tmap_q, q_full, (h0 + 1) * HD, qb * BQ, b);
        }
        __syncwarp();
        const int kv_row0 = (b * p.hkv + g) * p.lmax;
        for (int j = 0; j <= qb; ++j, ++blk) {
          const int st = blk % NSTAGE, use = blk / NSTAGE;
          mbar_wait(&k_empty[st], (use & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(&k_full[st], kTile);
            tma_load_2d(smem + kSmK + st * kTile, &tmap_k, &k_full[st], 0, kv_row0 + j * BKV);
          }
          __syncwarp();
        }
      }
    } else if (warp == 3) {
      int blk = 0;
      for (int k = 0; k < my_items; ++k) {
        int b, g, qb, h0, nh;
        decode(k, b, g, qb, h0, nh);
        const int kv_row0 = (b * p.hkv + g) * p.lmax;
        for (int j = 0; j <= qb; ++j, ++blk) {
          const int st = blk % NSTAGE, use = blk / NSTAGE;
          mbar_wait(&v_empty[st], (use & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(&v_full[st], kTile);
            tma_load_2d(smem + kSmV + st * kTile, &tmap_v, &v_full[st], 0, kv_row0 + j * BKV);
          }
          __syncwarp();
        }
      }
    } else {
      // MMA issuer of half g.  The K / V rings are indexed by the CTA-wide block counter `blk` (both halves see every tile);
      // the per-half barriers (s_full, s_free, p_full, o_done) by `cnt`, which only counts the blocks this half processes.
      const int g = warp - 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);
      const uint32_t sk = smem_u32(smem + kSmK), sv = smem_u32(smem + kSmV), sp = smem_u32(smem + kSmP) + g * kPBuf;
      const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(smem + kSmQ) + g * kTile);
      const uint32_t tm_s = tmem_base + g * BKV, tm_o = tmem_base + 256 + g * HD;
      // cursor over the flat sequence of (item, key block) tiles
      struct Cur { int k, j, nkv, blk, active; };
      auto first = [&](Cur& c) {
        c.k = 0; c.j = 0; c.blk = 0;
        if (my_items > 0) { int b, gg, qb, h0, nh; decode(0, b, gg, qb, h0, nh); c.nkv = qb + 1; c.active = g < nh; }
      };
      auto next = [&](Cur& c) -> bool {   // false at the end
        ++c.blk;
        if (++c.j < c.nkv) return true;
        if (++c.k >= my_items) return false;
        int b, gg, qb, h0, nh;
        decode(c.k, b, gg, qb, h0, nh);
        c.j = 0; c.nkv = qb + 1; c.active = g < nh;
        return true;
      };
      auto issue_s = [&](const Cur& c) {
        const int st = c.blk % NSTAGE;
        if (c.j == 0) mbar_wait(q_full, c.k & 1);
        mbar_wait(&k_full[st], (c.blk / NSTAGE) & 1);
        if (c.active) {
          tc_fence_after();
          const uint64_t dk = umma_desc_kmajor_sw128(sk + st * kTile);
          if (elect_one_sync()) {
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) tc_mma_bf16(tm_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
            tc_commit(&s_full[g]);
            tc_commit(&k_empty[st]);
            if (c.j == c.nkv - 1) tc_commit(q_empty);
          }
        } else if (elect_one_sync()) {   // idle half of a single-head item: pass the tile through
          mbar_arrive(&k_empty[st]);
          if (c.j == c.nkv - 1) mbar_arrive(q_empty);
        }
        __syncwarp();
      };
      if (my_items > 0) {
        Cur cur, nxt;
        first(cur);
        issue_s(cur);
        int cnt = 0;   // blocks processed by this half so far
        bool more = true;
        while (more) {
          nxt = cur;
          more = next(nxt);
          if (more) {
            if (cur.active && nxt.active) mbar_wait(&s_free[g], cnt & 1);            // softmax has pulled S(cur) out of TMEM
            else if (!cur.active && nxt.active && cnt > 0) mbar_wait(&s_free[g], (cnt - 1) & 1);  // ... S of the last active block
            issue_s(nxt);
          }
          const int st = cur.blk % NSTAGE;
          if (cur.active) {
            mbar_wait(&p_full[g], cnt & 1);
            mbar_wait(&v_full[st], (cur.blk / NSTAGE) & 1);
            tc_fence_after();
            const uint64_t dv = umma_desc_mnmajor_sw128(sv + st * kTile, kTile);
            const uint64_t dp0 = umma_desc_kmajor_sw128(sp), dp1 = umma_desc_kmajor_sw128(sp + BQ * 128);
            if (elect_one_sync()) {
#pragma unroll
              for (int kk = 0; kk < BKV / 16; ++kk)
                tc_mma_bf16(tm_o, (kk < 4 ? dp0 : dp1) + 2 * (kk & 3), dv + (uint64_t)kk * (16 * 128 >> 4), idesc_o, (cur.j | kk) != 0);
              tc_commit(&o_done[2 * g + (cnt & 1)]);
              tc_commit(&v_empty[st]);
            }
            __syncwarp();
            ++cnt;
          } else {
            mbar_wait(&v_full[st], (cur.blk / NSTAGE) & 1);
            if (elect_one_sync()) mbar_arrive(&v_empty[st]);
            __syncwarp();
          }
          cur = nxt;
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ---------------- softmax warpgroups: thread <-> query row <-> TMEM lane ----------------
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * BKV + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + g * HD + lane_off;
    uint8_t* prow = smem + kSmP + g * kPBuf + r * 128;
    const int rsw = r & 7;
    const float scale = p.scale_log2;
    const uint64_t sc2 = pack2f(scale, scale);
    int cnt = 0;

    for (int k = 0; k < my_items; ++k) {
      int b, kvh, qb, h0, nh;
      decode(k, b, kvh, qb, h0, nh);
      if (g >= nh) continue;
      const int h = h0 + g;
      const int row = qb * BQ + r;           // query position inside the sequence (past = 0)
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j <= qb; ++j, ++cnt) {
        mbar_wait(&s_full[g], cnt & 1);
        tc_fence_after();
        uint32_t s0[32], s1[32], s2[32], s3[32];
        tmem_ld_32x32(tmem_s + 0, s0);
        tmem_ld_32x32(tmem_s + 32, s1);
        tmem_ld_32x32(tmem_s + 64, s2);
        tmem_ld_32x32(tmem_s + 96, s3);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
        // the (single) P buffer is still being read by the previous PV of this half until that MMA retires
        if (cnt >= 1) {
          mbar_wait(&o_done[2 * g + ((cnt - 1) & 1)], ((cnt - 1) >> 1) & 1);
          tc_fence_after();
        }
        // masks: causal on the diagonal block, key padding where the block is not all-valid
        uint32_t km[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (p.kmask) {
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.kmask + ((size_t)b * p.nqb + j) * 4));
          km[0] = w.x; km[1] = w.y; km[2] = w.z; km[3] = w.w;
        }
        if (j == qb) {   // key c of the block visible iff c <= r
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int hi = r - 32 * w;   // bits 0..hi stay
            const uint32_t cm = hi >= 31 ? 0xffffffffu : (hi < 0 ? 0u : ((2u << hi) - 1u));
            km[w] &= cm;
          }
        }
        if ((km[0] & km[1] & km[2] & km[3]) != 0xffffffffu) {
          const uint32_t ninf = 0xff800000u;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s0[i] = ((km[0] >> i) & 1u) ? s0[i] : ninf;
            s1[i] = ((km[1] >> i) & 1u) ? s1[i] : ninf;
            s2[i] = ((km[2] >> i) & 1u) ? s2[i] : ninf;
            s3[i] = ((km[3] >> i) & 1u) ? s3[i] : ninf;
          }
        }
        float m_blk = max32(s0, -INFINITY);
        m_blk = max32(s1, m_blk);
        m_blk = max32(s2, m_blk);
        m_blk = max32(s3, m_blk);
        const float m_new = fmaxf(m_ref, m_blk * scale);
        const bool grow = m_new > m_ref + kLazyThreshold;   // also true when the first visible key appears (m_ref = -inf)
        const float f = grow ? ex2_approx(m_ref - m_new) : 1.0f;
        if (grow) { m_ref = m_new; l_run *= f; }
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // (the previous PV of this half has retired, see above: O may be rescaled in place)
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32(tmem_o + c, o);
          }
          tmem_st_wait();
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;   // no visible key yet: every p is exp2(-inf) = 0
        const uint64_t nm2 = pack2f(-m_use, -m_use);
        uint64_t sum2 = pack2f(0.f, 0.f);
        exp_store32_plain(s0, sc2, nm2, sum2, prow, 0, rsw);
        exp_store32_plain(s1, sc2, nm2, sum2, prow, 4, rsw);
        exp_store32_plain(s2, sc2, nm2, sum2, prow + BQ * 128, 0, rsw);
        exp_store32_plain(s3, sc2, nm2, sum2, prow + BQ * 128, 4, rsw);
        {
          float lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(sum2));
          l_run += lo + hi;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // epilogue: O / l
      mbar_wait(&o_done[2 * g + ((cnt - 1) & 1)], ((cnt - 1) >> 1) & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      const bool row_ok = row < p.L;
      bf16* orow = p.out + ((size_t)b * p.L + row) * ((size_t)p.hq * HD) + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld_32x32(tmem_o + c, o);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(o[v8 * 8 + 0]) * inv, __uint_as_float(o[v8 * 8 + 1]) * inv);
            u.y = pack_bf16(__uint_as_float(o[v8 * 8 + 2]) * inv, __uint_as_float(o[v8 * 8 + 3]) * inv);
            u.z = pack_bf16(__uint_as_float(o[v8 * 8 + 4]) * inv, __uint_as_float(o[v8 * 8 + 5]) * inv);
            u.w = pack_bf16(__uint_as_float(o[v8 * 8 + 6]) * inv, __uint_as_float(o[v8 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + v8 * 8) = u;
          }
        }
      }
      if (p.lse && row_ok)
        p.lse[((size_t)b * p.hq + h) * p.L + row] = l_run > 0.f ? (m_ref + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
      tc_fence_before();  // O has been read out: the next item's first PV of this half may overwrite it (ordered by p_full)
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

size_t slb_attn_gqa2_workspace(int batch, int lq) { return (size_t)batch * ((lq + BQ - 1) / BQ) * 4 * sizeof(uint32_t); }

// returns 1 if the kernel applies (and was launched; rc in *rc_out), 0 if the caller should use the first-generation kernel.
// kmask_ws: device scratch of slb_attn_gqa2_workspace bytes, needed only when key_valid is given.
int slb_attn_gqa2_try(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid, int key_valid_ld,
                      void* out, float* lse, int batch, int lq, int past, int lmax, int hq, int hkv, uint32_t* kmask_ws, cudaStream_t stream,
                      int* rc_out) {
  *rc_out = SLB_OK;
  if (past != 0 || lq < BQ || hq % hkv != 0 || (key_valid && !kmask_ws)) return 0;
  const int nqb = (lq + BQ - 1) / BQ;
  if (lmax < nqb * BKV) return 0;   // the cache must hold whole key blocks (rows beyond lq are zero or masked by causality)
  CUtensorMap tq, tk, tv;
  int rc = slb_make_tmap_3d(&tq, q, (uint64_t)ldq, (uint64_t)lq, (uint64_t)batch, (uint64_t)ldq * 2, (uint64_t)lq * ldq * 2, HD, BQ, 1);
  if (!rc) rc = slb_make_tmap_2d(&tk, kcache, HD, (uint64_t)batch * hkv * lmax, HD * 2, HD, BKV);
  if (!rc) rc = slb_make_tmap_2d(&tv, vcache, HD, (uint64_t)batch * hkv * lmax, HD * 2, HD, BKV);
  if (rc) { *rc_out = rc; return 1; }
  Gqa2Params p;
  p.out = (bf16*)out; p.lse = lse; p.kmask = nullptr;
  p.L = lq; p.lmax = lmax; p.batch = batch; p.hq = hq; p.hkv = hkv; p.group = hq / hkv; p.npairs = (p.group + 1) / 2; p.nqb = nqb;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  if (key_valid) {
    gqa2_kmask_kernel<<<dim3(nqb, batch), 128, 0, stream>>>(key_valid, key_valid_ld, lq, nqb, kmask_ws);
    p.kmask = kmask_ws;
  }
  const int n_items = nqb * batch * hkv * p.npairs;
  const int grid = n_items < slb_num_sms() ? n_items : slb_num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_gqa2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal);
    if (e != cudaSuccess) { *rc_out = slb_fail(SLB_ECUDA, "attn_gqa2 attribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set = true;
  }
  cudaError_t e = slb_launch_pdl((long long)batch * lq <= 4096, attn_gqa2_kernel, dim3(grid), dim3(G2_THREADS), (size_t)kSmTotal, stream, tq, tk, tv, p);
  if (e != cudaSuccess) *rc_out = slb_fail(SLB_ECUDA, "attn_gqa2 launch: %s", cudaGetErrorString(e));
  return 1;
}
