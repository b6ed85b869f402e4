// InternViT attention, second generation (bidirectional, head_dim 64, n_tokens = 1 + 256*k, CLS at row 0).
//
//   One CTA per (256 patch-token queries, head, tile), one CTA per SM, 12 warps (setmaxnreg re-balanced):
//     warp 0 / 3  TMA producers: Q0,Q1 per item + the K_j stream / the V_j stream (128 keys per tile, 4-deep rings;
//                 a K tile is recycled right after its S MMAs, a V tile after its PV MMAs)
//     warps 1,2   MMA issuers (one per query half g): S_g = Q_g K_j^T (128x128x64), O_g += P_g V_j (128x64x128),
//                 ping-ponging between the two query halves g = 0,1 so that one half's softmax overlaps the other
//                 half's MMAs.  O_g accumulates in TMEM across all key blocks.
//     warps 4-7   softmax warpgroup 0 (thread <-> query row <-> TMEM lane), warps 8-11 softmax warpgroup 1:
//                 single TMEM pass (the 128 scores of the row live in registers), exp2 with a *lazy* running
//                 maximum (O/l are rescaled only when the maximum grows by more than 2^8, FA4-style), P written
//                 to shared memory in the K-major 128B-swizzled UMMA layout.
//   The CLS token is peeled off the tensor-core tiling: as a *key* it is folded in by the softmax threads
//   (one 64-long dot product per row, added to O in the epilogue); as a *query* it is handled by a tiny
//   CUDA-core kernel (one block per (tile, head)).  That keeps every MMA tile full: 1025 = 1 + 8*128.
#include "common.cuh"
#include "attn_common.cuh"
#include "../../include/simlingo_b200.h"

#include <cstdlib>

namespace {

constexpr int HD = 64, BQ = 128, BKV = 128, NSTAGE = 4;
constexpr int V2_THREADS = 384;  // 3 warpgroups: {TMA, MMA, 2 idle}, softmax 0, softmax 1
constexpr int kTile = BQ * HD * 2;            // 16 KB: one 128 x 64 bf16 operand tile
constexpr int kSmQ = 0;                       // 2 tiles
constexpr int kSmK = kSmQ + 2 * kTile;        // NSTAGE tiles
constexpr int kSmV = kSmK + NSTAGE * kTile;   // NSTAGE tiles
constexpr int kSmP = kSmV + NSTAGE * kTile;   // [g] : 2 x 32 KB
constexpr int kPBuf = BQ * BKV * 2;
constexpr int kSmBar = kSmP + 2 * kPBuf;
constexpr int kSmTotal = kSmBar + 256;
constexpr float kLazyThreshold = 8.0f;        // log2 units
constexpr int kDefaultPTmem = 1;              // 1: P operand of the PV product in tensor memory (tcgen05.st + TS-mode MMA)
constexpr int kDefaultEmuPairs = 4;           // pairs per 16 whose exp2 runs on the FMA / ALU pipes (ex2_emu2)
#ifndef VIT2_NO_SPEC
#define VIT2_NO_SPEC 1  /* speculative exp pass disabled: no measurable gain, see DESIGN.md */
#endif

// optional timeline trace (debug): CTA 0 records (tag, clock64) pairs per role; enabled by slb_debug_set_trace()
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

struct Vit2Params {
  long long* trace;
  const bf16* qkv;   // [tiles * n_tokens, 3C]
  bf16* out;         // [tiles * n_tokens, C]
  float* lse;        // [tiles, heads, n_tokens] or null
  int n_tokens, heads, nkv, tiles;
  float scale_log2;
};

// p = exp2(s * scale - mref) for 32 scores; accumulates the (packed) row sum; writes 64 bytes (4 x 16 B chunks) of P.
// EMU: which of the 16 pairs take the software exp2 (FMA / ALU pipes) instead of MUFU, see ex2_emu2.
// PT: P goes to tensor memory (16 columns of packed bf16 pairs at tp) instead of the swizzled smem tile.
template <int VAR, uint32_t EMU, int PT>
__device__ __forceinline__ void exp_store32(const uint32_t (&s)[32], uint64_t sc2, uint64_t nm2, uint64_t& sum2, uint8_t* half_row,
                                            int chunk0, int rsw, uint32_t tp) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x2 = f2fma(f2pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
    uint64_t e2 = x2;
    if (!(VAR & 1)) e2 = ex2_pair<EMU>(x2, i);
    sum2 = fadd2(sum2, e2);
    float a, b;
    f2unpack(e2, a, b);
    pk[i] = pack_bf16(a, b);
  }
  if (VAR & 4) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= pk[i];
    if (x == 0x12345678u) *reinterpret_cast<uint32_t*>(half_row) = x;
    return;
  }
  if (PT) {
    tmem_st_32x16(tp, pk);
    return;
  }
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    *reinterpret_cast<uint4*>(half_row + (((chunk0 + q4) ^ rsw) << 4)) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
}

// speculative variant of exp_store32: also tracks the block maximum of the raw scores (FMNMX3 on the ALU pipe runs
// under the shadow of the MUFU-bound exp2 stream)
__device__ __forceinline__ void exp_store32_max(const uint32_t (&s)[32], uint64_t sc2, uint64_t nm2, uint64_t& sum2, float& mx,
                                                uint8_t* half_row, int chunk0, int rsw) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = __uint_as_float(s[2 * i]), b = __uint_as_float(s[2 * i + 1]);
    mx = fmax3(mx, a, b);
    ffma2(a, b, sc2, nm2);
    a = ex2_approx(a);
    b = ex2_approx(b);
    sum2 = fadd2(sum2, pack2f(a, b));
    pk[i] = pack_bf16(a, b);
  }
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    *reinterpret_cast<uint4*>(half_row + (((chunk0 + q4) ^ rsw) << 4)) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
}

// Persistent: grid = #SMs, CTA c processes work items c, c + gridDim.x, ... (item = (query block of 256, head, tile),
// query block fastest so that concurrently running CTAs share K/V in L2).  Every mbarrier keeps counting across
// items (global block counter `blk`), so the TMA / MMA pipelines stay warm over item boundaries: the next item's
// Q, K_0 and S_0 are in flight while the softmax warps finish the current item's epilogue.
template <int VAR, uint32_t EMU, int PT>
__global__ void __launch_bounds__(V2_THREADS, 1)
attn_vit2_kernel(const __grid_constant__ CUtensorMap tmap, Vit2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // NSTAGE
  uint64_t* k_empty = k_full + NSTAGE;           // NSTAGE (both MMA warps commit after their S)
  uint64_t* v_full = k_empty + NSTAGE;           // NSTAGE
  uint64_t* v_empty = v_full + NSTAGE;           // NSTAGE (both MMA warps commit after their PV)
  uint64_t* s_full = v_empty + NSTAGE;           // [g]
  uint64_t* p_full = s_full + 2;                 // [g]
  uint64_t* s_free = p_full + 2;                 // [g]: softmax g has pulled S_g(blk) into registers
  uint64_t* o_done = s_free + 2;                 // [g][blk & 1]: PV_g(blk) (and everything issued before it) has retired
  uint64_t* q_empty = o_done + 4;                // both MMA warps have issued (and retired) the item's last S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int C = p.heads * HD;
  const int nkv = p.nkv;
  const int nqb = (p.n_tokens - 1) / (2 * BQ);
  const int n_items = nqb * p.heads * p.tiles;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_blk = my_items * nkv;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023) != 0) __trap();
    tma_prefetch_desc(&tmap);
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
  pdl_wait();   // qkv comes from the projection GEMM right before us

  // item index -> (first patch-token row, head, tile)
  auto decode = [&](int k, int& q0, int& h, int& t) {
    const int w = (int)blockIdx.x + k * (int)gridDim.x;
    q0 = 1 + (w % nqb) * 2 * BQ;
    h = (w / nqb) % p.heads;
    t = w / (nqb * p.heads);
  };

  // register re-balancing between warpgroups (the kernel is launched at 168 regs/thread = 65536 / 384)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      {  // Q per item + the K stream (a K tile is recycled as soon as both S MMAs that read it retire)
        int blk = 0;
        Trace tr{(blockIdx.x == 0 && lane == 0) ? p.trace : nullptr, 2, 0};
        for (int k = 0; k < my_items; ++k) {
          int q0, h, t;
          decode(k, q0, h, t);
          if (k > 0) mbar_wait(q_empty, (k - 1) & 1);
          tr.rec(1000 + k);  // issuing Q(k)
          if (elect_one_sync()) {
            mbar_expect_tx(q_full, 2 * kTile);
            tma_load_3d(smem + kSmQ, &tmap, q_full, h * HD, q0, t);
            tma_load_3d(smem + kSmQ + kTile, &tmap, q_full, h * HD, q0 + BQ, t);
          }
          __syncwarp();
          for (int j = 0; j < nkv; ++j, ++blk) {
            const int st = blk % NSTAGE, use = blk / NSTAGE;
            mbar_wait(&k_empty[st], (use & 1) ^ 1);
            tr.rec(1100 + blk);  // issuing K(blk)
            if (elect_one_sync()) {
              mbar_expect_tx(&k_full[st], kTile);
              tma_load_3d(smem + kSmK + st * kTile, &tmap, &k_full[st], C + h * HD, 1 + j * BKV, t);
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == 3) {
      {  // the V stream (a V tile is held until both PV MMAs retire)
        int blk = 0;
        for (int k = 0; k < my_items; ++k) {
          int q0, h, t;
          decode(k, q0, h, t);
          for (int j = 0; j < nkv; ++j, ++blk) {
            const int st = blk % NSTAGE, use = blk / NSTAGE;
            mbar_wait(&v_empty[st], (use & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_expect_tx(&v_full[st], kTile);
              tma_load_3d(smem + kSmV + st * kTile, &tmap, &v_full[st], 2 * C + h * HD, 1 + j * BKV, t);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // one MMA-issuing warp per query half g: S_g(blk+1) is issued as soon as softmax g has pulled S_g(blk) into
      // registers (s_free), i.e. long before P_g(blk) is ready, so the softmax warps never wait for a score tile
      {
        const int g = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);
        const uint32_t sk = smem_u32(smem + kSmK), sv = smem_u32(smem + kSmV), sp = smem_u32(smem + kSmP) + g * kPBuf;
        const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(smem + kSmQ) + g * kTile);
        const uint32_t tm_s = tmem_base + g * BKV, tm_o = tmem_base + 256 + g * HD, tm_p = tmem_base + 384 + g * (BKV / 2);
        Trace tr{(blockIdx.x == 0 && g == 0 && lane == 0) ? p.trace : nullptr, 0, 0};
        auto issue_s = [&](int blk) {
          const int st = blk % NSTAGE, j = blk % nkv;
          if (j == 0) mbar_wait(q_full, (blk / nkv) & 1);
          mbar_wait(&k_full[st], (blk / NSTAGE) & 1);
          tr.rec(100 + blk);  // K ready, issuing S(blk)
          tc_fence_after();
          const uint64_t dk = umma_desc_kmajor_sw128(sk + st * kTile);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) tc_mma_bf16(tm_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            tc_commit(&s_full[g]);
            tc_commit(&k_empty[st]);
            if (j == nkv - 1) tc_commit(q_empty);
          }
          __syncwarp();
          tr.rec(150 + blk);  // S(blk) issued + committed
        };
        if (total_blk > 0) issue_s(0);
        for (int blk = 0; blk < total_blk; ++blk) {
          const int st = blk % NSTAGE, j = blk % nkv;
          if (blk + 1 < total_blk) {
            mbar_wait(&s_free[g], blk & 1);
            tr.rec(200 + blk);  // s_free(blk) seen
            issue_s(blk + 1);
          }
          mbar_wait(&p_full[g], blk & 1);
          tr.rec(300 + blk);  // p_full(blk) seen
          mbar_wait(&v_full[st], (blk / NSTAGE) & 1);
          tr.rec(400 + blk);  // V ready, issuing PV(blk)
          tc_fence_after();
          const uint64_t dv = umma_desc_mnmajor_sw128(sv + st * kTile, kTile);
          const uint64_t dp0 = umma_desc_kmajor_sw128(sp), dp1 = umma_desc_kmajor_sw128(sp + BQ * 128);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) {
              if (PT) tc_mma_bf16_ts(tm_o, tm_p + 8 * k, dv + (uint64_t)k * (16 * 128 >> 4), idesc_o, (j | k) != 0);
              else tc_mma_bf16(tm_o, (k < 4 ? dp0 : dp1) + 2 * (k & 3), dv + (uint64_t)k * (16 * 128 >> 4), idesc_o, (j | k) != 0);
            }
            tc_commit(&o_done[2 * g + (blk & 1)]);
            tc_commit(&v_empty[st]);
          }
          __syncwarp();
          tr.rec(450 + blk);  // PV(blk) issued + committed
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ---------------- softmax warpgroups ----------------
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * BKV + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + g * HD + lane_off;
    const uint32_t tmem_p = tmem_base + 384 + g * (BKV / 2) + lane_off;   // PT: P_g as packed bf16 pairs, 64 columns
    uint8_t* prow = smem + kSmP + g * kPBuf + r * 128;
    const int rsw = r & 7;
    const float scale = p.scale_log2;
    const uint64_t sc2 = pack2f(scale, scale);
    int blk = 0;
    Trace tr{(blockIdx.x == 0 && g == 0 && lane == 0) ? p.trace : nullptr, warp == 4 ? 1 : warp - 2, 0};  // roles 1,3,4,5

    for (int k = 0; k < my_items; ++k) {
      int q0, h, t;
      decode(k, q0, h, t);
      const int row = q0 + g * BQ + r;  // token index inside the tile
      const bf16* kcls = p.qkv + (size_t)t * p.n_tokens * 3 * C + C + h * HD;
      const bf16* vcls = kcls + C;

      // CLS key: s_cls = q_row . k_cls (log2 domain), folded in as the initial state of the online softmax
      mbar_wait(q_full, k & 1);
      float s_cls = 0.f;
      {
        const uint8_t* qrow = smem + kSmQ + g * kTile + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ rsw) << 4));
          const uint4 kv = __ldg(reinterpret_cast<const uint4*>(kcls) + c);
          const float2 q0f = unpack_bf16(qv.x), q1f = unpack_bf16(qv.y), q2f = unpack_bf16(qv.z), q3f = unpack_bf16(qv.w);
          const float2 k0f = unpack_bf16(kv.x), k1f = unpack_bf16(kv.y), k2f = unpack_bf16(kv.z), k3f = unpack_bf16(kv.w);
          s_cls += q0f.x * k0f.x + q0f.y * k0f.y + q1f.x * k1f.x + q1f.y * k1f.y + q2f.x * k2f.x + q2f.y * k2f.y + q3f.x * k3f.x + q3f.y * k3f.y;
        }
        s_cls *= scale;
      }
      float m_ref = s_cls;  // running (lazy) maximum, log2 domain
      float l_run = 1.0f;   // exp2(s_cls - m_ref)

      tr.rec(900 + k);  // item prologue done
      for (int j = 0; j < nkv; ++j, ++blk) {
        mbar_wait(&s_full[g], blk & 1);
        tr.rec(500 + blk);  // s_full(blk) seen
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
        tr.rec(600 + blk);  // S in registers
        // the (single) P buffer is still being read by PV_g(blk-1) until that MMA retires
        if (blk >= 1) {
          mbar_wait(&o_done[2 * g + ((blk - 1) & 1)], ((blk - 1) >> 1) & 1);
          tc_fence_after();
        }

        uint64_t sum2 = pack2f(0.f, 0.f);
        bool redo = (j == 0) || VIT2_NO_SPEC;  // the first block of an item establishes the reference maximum first
        if (!redo) {
          // speculative pass with the current reference maximum; the block maximum is tracked on the side
          const uint64_t nm2 = pack2f(-m_ref, -m_ref);
          float m_blk = -INFINITY;
          exp_store32_max(s0, sc2, nm2, sum2, m_blk, prow, 0, rsw);
          exp_store32_max(s1, sc2, nm2, sum2, m_blk, prow, 4, rsw);
          exp_store32_max(s2, sc2, nm2, sum2, m_blk, prow + BQ * 128, 0, rsw);
          exp_store32_max(s3, sc2, nm2, sum2, m_blk, prow + BQ * 128, 4, rsw);
          redo = __any_sync(0xffffffffu, m_blk * scale > m_ref + kLazyThreshold);
        }
        if (redo) {
          // (re)establish the reference maximum, rescale O / l in place (rare after the first block), recompute P
          float m_blk = max32(s0, -INFINITY);
          if (!(VAR & 8)) {
            m_blk = max32(s1, m_blk);
            m_blk = max32(s2, m_blk);
            m_blk = max32(s3, m_blk);
          }
          const float m_new = fmaxf(m_ref, m_blk * scale);
          const bool grow = m_new > m_ref + kLazyThreshold;
          const float f = grow ? ex2_approx(m_ref - m_new) : 1.0f;
          if (grow) { m_ref = m_new; l_run *= f; }
          if (j > 0 && __any_sync(0xffffffffu, grow)) {
            // (PV_g(blk-1) has retired, see above: O may be rescaled in place)
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
          const uint64_t nm2 = pack2f(-m_ref, -m_ref);
          sum2 = pack2f(0.f, 0.f);
          // The exp2 stream saturates the MUFU pipe: let the two warpgroups take turns (named barriers 1 / 2) so
          // that one warpgroup's TMEM loads / max / fences / barrier latencies hide under the other's exp phase.
          if (VAR & 16) {
            if (g == 0) { if (blk > 0) asm volatile("bar.sync 1, 256;" ::: "memory"); }
            else        { asm volatile("bar.sync 2, 256;" ::: "memory"); }
          }
          exp_store32<VAR, EMU, PT>(s0, sc2, nm2, sum2, prow, 0, rsw, tmem_p);
          exp_store32<VAR, EMU, PT>(s1, sc2, nm2, sum2, prow, 4, rsw, tmem_p + 16);
          exp_store32<VAR, EMU, PT>(s2, sc2, nm2, sum2, prow + BQ * 128, 0, rsw, tmem_p + 32);
          exp_store32<VAR, EMU, PT>(s3, sc2, nm2, sum2, prow + BQ * 128, 4, rsw, tmem_p + 48);
          if (VAR & 16) {
            if (g == 0) { asm volatile("bar.arrive 2, 256;" ::: "memory"); }
            else if (blk + 1 < total_blk) { asm volatile("bar.arrive 1, 256;" ::: "memory"); }
          }
        }
        {
          float lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(sum2));
          l_run += lo + hi;
        }
        tr.rec(700 + blk);  // P written
        if (PT) tmem_st_wait();
        else if (!(VAR & 2)) fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        tr.rec(800 + blk);  // p_full arrived
      }
      // epilogue: O / l (+ the CLS key's contribution)
      mbar_wait(&o_done[2 * g + ((blk - 1) & 1)], ((blk - 1) >> 1) & 1);
      tr.rec(950 + k);  // last PV retired
      tc_fence_after();
      const float p_cls = ex2_approx(s_cls - m_ref);
      const float inv = 1.0f / l_run;
      bf16* orow = p.out + ((size_t)t * p.n_tokens + row) * C + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld_32x32(tmem_o + c, o);
        tmem_ld_wait();
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vcls + c) + v8);
          const float2 a = unpack_bf16(vv.x), b = unpack_bf16(vv.y), cc = unpack_bf16(vv.z), d = unpack_bf16(vv.w);
          const float vf[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
          float of[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) of[e] = (__uint_as_float(o[v8 * 8 + e]) + p_cls * vf[e]) * inv;
          uint4 u;
          u.x = pack_bf16(of[0], of[1]); u.y = pack_bf16(of[2], of[3]); u.z = pack_bf16(of[4], of[5]); u.w = pack_bf16(of[6], of[7]);
          *reinterpret_cast<uint4*>(orow + c + v8 * 8) = u;
        }
      }
      if (p.lse) p.lse[((size_t)t * p.heads + h) * p.n_tokens + row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
      tc_fence_before();  // O_g has been read out: the next item's first PV may overwrite it (ordered by p_full)
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// CLS query row.  One block per (tile, group of 4 heads), 512 threads.  A warp takes one key / value row at a time and its
// 32 lanes read the 512 contiguous bytes of that row which belong to the 4 heads (lane = 16-byte chunk: head lane/8, dims
// 8 (lane%8) ..+7), so every load is one fully coalesced 512-byte segment (the per-head kernel touched 32 different rows
// with every load: 225 us per layer at 128 tiles, 4 % of the offline step).  Scores: 8-lane butterfly reduction per head;
// P.V: each lane accumulates its 8 dims over the warp's keys, the 16 partials are combined through shared memory.
constexpr int kClsThreads = 512, kClsWarps = kClsThreads / 32, kClsHeads = 4;
__global__ void __launch_bounds__(kClsThreads)
attn_vit_cls_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int n_tokens, int heads, float scale) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* qs = sm;                                   // [4][64]
  float* red = qs + kClsHeads * 64;                 // [2][kClsWarps][4]
  float* part = red + 2 * kClsWarps * kClsHeads;    // [kClsWarps][256]
  float* sc = part + kClsWarps * 256;               // [4][n_tokens]
  const int hg = blockIdx.x, t = blockIdx.y;
  const int C = heads * HD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = lane >> 3, sub = lane & 7;
  const bf16* base = qkv + (size_t)t * n_tokens * 3 * C;
  if (tid < kClsHeads * 64) qs[tid] = __bfloat162float(base[hg * kClsHeads * HD + tid]) * scale;   // CLS = token 0 of the tile
  __syncthreads();
  float qf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qf[e] = qs[head * 64 + sub * 8 + e];
  const size_t rs = (size_t)3 * C / 8;  // row stride in uint4
  const uint4* kp = reinterpret_cast<const uint4*>(base + C + hg * kClsHeads * HD) + lane;
  const uint4* vp = reinterpret_cast<const uint4*>(base + 2 * C + hg * kClsHeads * HD) + lane;
  float mx = -INFINITY;
  for (int j0 = warp; j0 < n_tokens; j0 += 4 * kClsWarps) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int j = j0 + k * kClsWarps; u[k] = kp[(size_t)(j < n_tokens ? j : 0) * rs]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + k * kClsWarps;
      const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
      float s = a.x * qf[0] + a.y * qf[1] + b.x * qf[2] + b.y * qf[3] + c.x * qf[4] + c.y * qf[5] + d.x * qf[6] + d.y * qf[7];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (j < n_tokens) {
        if (sub == 0) sc[head * n_tokens + j] = s;
        mx = fmaxf(mx, s);
      }
    }
  }
  if (sub == 0) red[warp * kClsHeads + head] = mx;
  __syncthreads();
  float mxh[kClsHeads], sum[kClsHeads];
#pragma unroll
  for (int h = 0; h < kClsHeads; ++h) {
    float m = red[h];
    for (int w = 1; w < kClsWarps; ++w) m = fmaxf(m, red[w * kClsHeads + h]);
    mxh[h] = m;
    sum[h] = 0.f;
  }
  for (int j = tid; j < n_tokens; j += kClsThreads) {
#pragma unroll
    for (int h = 0; h < kClsHeads; ++h) {
      const float e = __expf(sc[h * n_tokens + j] - mxh[h]);
      sc[h * n_tokens + j] = e;
      sum[h] += e;
    }
  }
#pragma unroll
  for (int h = 0; h < kClsHeads; ++h) {
    sum[h] = warp_sum(sum[h]);
    if (lane == 0) red[(kClsWarps + warp) * kClsHeads + h] = sum[h];
  }
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = warp; j0 < n_tokens; j0 += 4 * kClsWarps) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int j = j0 + k * kClsWarps; u[k] = vp[(size_t)(j < n_tokens ? j : 0) * rs]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + k * kClsWarps;
      const float pj = j < n_tokens ? sc[head * n_tokens + j] : 0.f;
      const float2 a = unpack_bf16(u[k].x), b = unpack_bf16(u[k].y), c = unpack_bf16(u[k].z), d = unpack_bf16(u[k].w);
      acc[0] += pj * a.x; acc[1] += pj * a.y; acc[2] += pj * b.x; acc[3] += pj * b.y;
      acc[4] += pj * c.x; acc[5] += pj * c.y; acc[6] += pj * d.x; acc[7] += pj * d.y;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[warp * 256 + lane * 8 + e] = acc[e];
  __syncthreads();
  if (tid < 256) {
    const int h = tid >> 6;
    float o = 0.f, tot = 0.f;
    for (int w = 0; w < kClsWarps; ++w) { o += part[w * 256 + tid]; tot += red[(kClsWarps + w) * kClsHeads + h]; }
    out[(size_t)t * n_tokens * C + hg * kClsHeads * HD + tid] = __float2bfloat16(o / tot);
    if ((tid & 63) == 0 && lse) lse[((size_t)t * heads + hg * kClsHeads + h) * n_tokens] = mxh[h] + logf(tot);
  }
}

}  // namespace

static long long* g_vit2_trace = nullptr;
// debug hook (not part of the public header): device buffer of 3 * 256 * 2 int64 receiving CTA 0's timeline
extern "C" int slb_debug_set_trace(void* dev_buf) { g_vit2_trace = (long long*)dev_buf; return 0; }
long long* slb_debug_trace_ptr() { return g_vit2_trace; }

// returns 1 if the specialised kernel applies (and was launched), 0 if the caller should use the generic kernel
int slb_attn_vit2_try(const void* qkv, void* out, float* lse, int tiles, int n_tokens, int heads, cudaStream_t stream, int* rc_out) {
  *rc_out = SLB_OK;
  if (n_tokens < 257 || ((n_tokens - 1) % 256) != 0 || (heads % 4) != 0) return 0;
  const int C = heads * HD;
  CUtensorMap tm;
  int rc = slb_make_tmap_3d(&tm, qkv, (uint64_t)3 * C, (uint64_t)n_tokens, (uint64_t)tiles, (uint64_t)3 * C * 2,
                            (uint64_t)n_tokens * 3 * C * 2, HD, BQ, 1);
  if (rc) { *rc_out = rc; return 1; }
  Vit2Params p;
  p.trace = g_vit2_trace;
  p.qkv = (const bf16*)qkv; p.out = (bf16*)out; p.lse = lse;
  p.n_tokens = n_tokens; p.heads = heads; p.nkv = (n_tokens - 1) / BKV; p.tiles = tiles;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  const int n_items = ((n_tokens - 1) / (2 * BQ)) * heads * tiles;
  const int grid = n_items < slb_num_sms() ? n_items : slb_num_sms();
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal);
    if (e != cudaSuccess) return e;
    return slb_launch_pdl(tiles <= 8, kern, dim3(grid), dim3(V2_THREADS), (size_t)kSmTotal, stream, tm, p);
  };
  cudaError_t e;
  // software-exp2 share (pairs out of 16 per 32-score chunk): SLB_VIT2_EMU selects 0 / 4 / 5 / 6 / 8 for tuning runs; valid
  // results in every setting
  static int emu = -1;
  if (emu < 0) { const char* ev = getenv("SLB_VIT2_EMU"); emu = ev ? atoi(ev) : kDefaultEmuPairs; }
#ifdef SLB_ABLATION
  // SLB_VIT2_VARIANT (timing experiments only, results are wrong; not compiled into the product library): 1 no exp2, 2 no proxy
  // fence, 4 no P stores, 8 short max; 16 = warpgroup turn-taking around the exp phase (valid results; measured: no gain)
  static int variant = -1;
  if (variant < 0) { const char* ev = getenv("SLB_VIT2_VARIANT"); variant = ev ? atoi(ev) : 0; }
  switch (variant) {
    case 1: e = launch(attn_vit2_kernel<1, 0, 0>); break;
    case 2: e = launch(attn_vit2_kernel<2, 0, 0>); break;
    case 4: e = launch(attn_vit2_kernel<4, 0, 0>); break;
    case 6: e = launch(attn_vit2_kernel<6, 0, 0>); break;
    case 7: e = launch(attn_vit2_kernel<7, 0, 0>); break;
    case 8: e = launch(attn_vit2_kernel<8, 0, 0>); break;
    case 15: e = launch(attn_vit2_kernel<15, 0, 0>); break;
    case 16: e = launch(attn_vit2_kernel<16, 0, 0>); break;
    default: e = launch(attn_vit2_kernel<0, 0, 0>); break;
  }
#else
  static int pt = -1;   // SLB_VIT2_PT=1: P through tensor memory (TS-mode PV product), tuning switch
  if (pt < 0) { const char* ev = getenv("SLB_VIT2_PT"); pt = ev ? atoi(ev) : kDefaultPTmem; }
  switch (emu + 100 * pt) {
    case 0: e = launch(attn_vit2_kernel<0, 0x0000u, 0>); break;
    case 4: e = launch(attn_vit2_kernel<0, 0x2222u, 0>); break;
    case 5: e = launch(attn_vit2_kernel<0, 0x2492u, 0>); break;
    case 8: e = launch(attn_vit2_kernel<0, 0xAAAAu, 0>); break;
    case 100: e = launch(attn_vit2_kernel<0, 0x0000u, 1>); break;
    case 104: e = launch(attn_vit2_kernel<0, 0x2222u, 1>); break;
    case 105: e = launch(attn_vit2_kernel<0, 0x2492u, 1>); break;
    case 108: e = launch(attn_vit2_kernel<0, 0xAAAAu, 1>); break;
    case 106: e = launch(attn_vit2_kernel<0, 0xA492u, 1>); break;
    default: e = launch(attn_vit2_kernel<0, 0xA492u, 0>); break;   // 6 of 16
  }
#endif
  if (e != cudaSuccess) { *rc_out = slb_fail(SLB_ECUDA, "attn_vit2 launch: %s", cudaGetErrorString(e)); return 1; }
  const size_t smem = ((size_t)kClsHeads * 64 + 2 * kClsWarps * kClsHeads + kClsWarps * 256 + (size_t)kClsHeads * n_tokens) * sizeof(float);
  e = slb_launch_pdl(tiles <= 8, attn_vit_cls_kernel, dim3(heads / kClsHeads, tiles), dim3(kClsThreads), smem, stream, (const bf16*)qkv, (bf16*)out, lse, n_tokens, heads, 0.125f);
  if (e != cudaSuccess) *rc_out = slb_fail(SLB_ECUDA, "attn_vit_cls launch: %s", cudaGetErrorString(e));
  return 1;
}
