// Greedy decode loop of the Qwen2 decoder as ONE persistent kernel (llm.py:217-248 per generated token: embed the last sampled
// token, 24 decoder layers against the KV cache, final RMSNorm, LM head, argmax, EOS bookkeeping).
//
// Why: a decode step is ~1 GB of weights streamed once (0.15 ms at the HBM roofline) but, as a chain of ~125 dependent launches of
// 4-20 us each, it took 1.0 ms (batch 1) to 2.1 ms (batch 32) per token (profiles/r02_kernel_breakdown_language_v1.log).  Here one
// CTA per SM stays resident for the whole generation; the phases of a layer are separated by grid-wide barriers (~1 us) instead of
// kernel boundaries, every phase's weights are pulled towards L2 (prefetch.global.L2) while the previous phase computes, and the
// token loop, the position counter and the EOS test live on the device (no host round trip per token).
//
// Phases per layer (grid barrier after each):
//   1 RMSNorm(x) -> bf16 rows in shared memory, q|k|v projection + bias           -> qkv   (144 n8 weight tiles over the CTAs)
//   2 RoPE(q, new k), KV-cache write, attention over positions 0..pos, split over  -> att   (batch x kv heads x key segments;
//     key segments so that batch 1 still spreads over CTAs                                   last finisher of a (b, kv head) merges)
//   3 o projection, += into the fp32 residual stream                               -> x     (112 tiles)
//   4 RMSNorm(x) -> smem, gate|up projection + SwiGLU                              -> act   (608 gate/up tile pairs)
//   5 down projection, += into the residual stream                                 -> x     (112 tiles, K = 4864)
// then final RMSNorm + LM head with a running arg-max per CTA (151 655 rows, 272 MB: the largest stream of the step), a barrier,
// the cross-CTA arg-max + bookkeeping (sampled tokens, finished flags, counts; semantics of Engine._generate_graphed.sample) by CTA 0,
// a barrier.  All products run on mma.sync.m16n8k16 (bf16, fp32 accumulate) with the batch rows as the M operand (<= 32): the work is
// weight-streaming, a 128-row tcgen05 tile would be >= 75 % padding.  Same rounding points as the per-kernel chain: bf16 normalised
// rows, bf16 qkv / attention output / SwiGLU output, fp32 residual stream, fp32 logits.
//
// Cross-CTA data (x, qkv, att, act, partial results, token ids, flags) is read with ld.global.cg (L2) only - never through the
// non-coherent path - and every producer/consumer pair is separated by a grid barrier (release add + acquire spin) or by the
// fence + counter hand-off of the attention merge.  Weights and the embedding table are read-only for the kernel's lifetime.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

#include <climits>
#include <cstdlib>

long long* slb_debug_trace_ptr();  // attention_vit.cu (slb_debug_set_trace)

namespace {

constexpr int kDecThreads = 512, kDecWarps = kDecThreads / 32;
constexpr int kMaxG = 8;        // q heads per kv head
constexpr int kSegMax = 1024;   // keys per attention work item
constexpr int kAPad = 32;       // bf16 elements of row padding of the staged rows: row stride = 64 B mod 128 B, conflict-free uint4 fragment loads
constexpr int kPartStride = 66; // floats per (item, head) attention partial: 64 outputs, running max, sum

struct DecLayerDev { const bf16 *qkv, *bqkv, *o, *gu, *d, *ln1, *ln2; };   // = slb_decode_layer

struct DecParams {
  const DecLayerDev* layers;
  int n_layers, M, D, I, V, hq, hkv, lmax, n_steps, max_new, n_seg;
  long long emb_rows, eos, layer_stride, ld_sampled;
  const bf16 *emb, *norm_w, *lm_head;
  bf16 *kc, *vc;
  int* pos; long long* nxt; long long* sampled; long long* step; unsigned char* done; long long* n_gen;
  float log2_theta, eps;
  float* x; bf16* qkv; bf16* att; bf16* act;
  float* attn_part; unsigned* attn_cnt;
  float* best_val; int* best_idx;
  unsigned* bar; unsigned* err;
  int flags;          // tuning switches (SLB_DECODE_FLAGS): 1 = fence-heavy barrier (explicit __threadfence around the release / acquire), 2 = no L2 prefetch
  long long* trace;   // debug hook (slb_debug_set_trace): [gridDim][kTraceBarriers][2] SM clocks at barrier entry / exit, or NULL
};
constexpr int kTraceBarriers = 320;

// ---- coherent (L2) loads of data produced by other CTAs during this kernel ----
__device__ __forceinline__ uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg_i(const int* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  const unsigned short v = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Grid-wide barrier: every CTA adds 1 (release) to a counter that the host zeroed before the launch and spins (acquire) until the
// counter reaches epoch * gridDim.x.  All CTAs are co-resident (cooperative launch, one CTA per SM).  A wait of more than 2 s sets the
// error flag and makes every CTA leave the kernel (a hung barrier must not hang the GPU); the host reports it.
__device__ __forceinline__ bool grid_barrier(const DecParams& p, unsigned& epoch, int* s_abort) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    if (p.trace && epoch <= (unsigned)kTraceBarriers) p.trace[((size_t)blockIdx.x * kTraceBarriers + epoch - 1) * 2] = clock64();
    // release: the add is ordered after every write this CTA made before the __syncthreads above (cumulativity through the CTA barrier);
    // acquire: relaxed polling, one acquire fence once the count is reached - no fence inside the spin loop
    if (p.flags & 1) __threadfence();
    red_release_add(p.bar, 1u);
    const unsigned target = epoch * gridDim.x;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    int abort = 0;
    while (ld_relaxed_u32(p.bar) < target) {
      if ((++spins & 1023u) == 0) {
        if (ld_relaxed_u32(p.err)) { abort = 1; break; }
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) { atomicExch(p.err, 1u); abort = 1; break; }
      }
    }
    fence_acq_rel_gpu();
    if (p.flags & 1) __threadfence();
    *s_abort = abort;
    if (p.trace && epoch <= (unsigned)kTraceBarriers) p.trace[((size_t)blockIdx.x * kTraceBarriers + epoch - 1) * 2 + 1] = clock64();
  }
  __syncthreads();
  return *s_abort == 0;
}

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// pull `nrows` weight rows of K bf16 each towards L2 (fire and forget)
__device__ __forceinline__ void prefetch_rows(const DecParams& p, const bf16* W, long long ldw, int row0, int nrows, int K) {
  if (p.flags & 2) return;
  const int lines = (K * 2 + 127) >> 7;
  for (int i = threadIdx.x; i < nrows * lines; i += kDecThreads) {
    const int r = i / lines, ln = i - r * lines;
    prefetch_l2(reinterpret_cast<const uint8_t*>(W + (size_t)(row0 + r) * ldw) + (size_t)ln * 128);
  }
}

// contiguous share [a, b) of n units for this CTA
__device__ __forceinline__ void cta_share(int n, int& a, int& b) {
  a = (int)(((long long)blockIdx.x * n) / gridDim.x);
  b = (int)(((long long)(blockIdx.x + 1) * n) / gridDim.x);
}

// first weight row of the gate (nt = 0) / up (nt = 1) n8 tile of SwiGLU output group `pt` (outputs 8 pt .. 8 pt + 7) in the
// 128-row interleaved gate|up layout of Engine._interleave_gate_up
__device__ __forceinline__ int gu_row(int pt, int nt) {
  const int j0 = pt * 8;
  return 256 * (j0 >> 7) + nt * 128 + (j0 & 127);
}

// RMSNorm of the M residual rows into shared memory (bf16, one rounding; x * rstd * w as norm_fwd_kernel): one warp per row.
// from_emb: the rows are the embeddings of the last sampled tokens (start of a step); CTA 0 then also seeds the fp32 residual stream.
__device__ __forceinline__ void stage_norm(const DecParams& p, bf16* As, int SA, const bf16* __restrict__ w, bool from_emb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  for (int m = warp; m < p.M; m += kDecWarps) {
    float v[8][4];
    float ss = 0.f;
    const bf16* er = nullptr;
    if (from_emb) {
      long long id = __ldcg(p.nxt + m);
      id = id < 0 ? 0 : (id >= p.emb_rows ? p.emb_rows - 1 : id);
      er = p.emb + (size_t)id * D;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = (i * 32 + lane) * 4;
      if (k < D) {
        if (from_emb) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(er + k));
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
          if (blockIdx.x == 0) *reinterpret_cast<float4*>(p.x + (size_t)m * D + k) = make_float4(a.x, a.y, b.x, b.y);
        } else {
          const float4 f = ldcg_f4(p.x + (size_t)m * D + k);
          v[i][0] = f.x; v[i][1] = f.y; v[i][2] = f.z; v[i][3] = f.w;
        }
        ss += v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2] + v[i][3] * v[i][3];
      }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / D + p.eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = (i * 32 + lane) * 4;
      if (k < D) {
        const uint2 wu = __ldg(reinterpret_cast<const uint2*>(w + k));
        const float2 wa = unpack_bf16(wu.x), wb = unpack_bf16(wu.y);
        uint2 o;
        o.x = pack_bf16(v[i][0] * rstd * wa.x, v[i][1] * rstd * wa.y);
        o.y = pack_bf16(v[i][2] * rstd * wb.x, v[i][3] * rstd * wb.y);
        *reinterpret_cast<uint2*>(As + (size_t)m * SA + k) = o;
      }
    }
  }
  __syncthreads();
}

// Skinny product of this CTA's weight-tile groups [g_begin, g_end) against the M activation rows.  A group = NT n8 tiles (8 weight
// rows each) whose first rows come from row_of(group, nt).  Up to 16 groups are processed per round; the 16 warps are split into
// KS = 16 / groups K-slices per group (interleaved 32-element chunks), partial tiles are summed through shared memory in a fixed order,
// and epi(group, m, c, v[NT]) receives the NT sums of row m, column c (0..7) of the group's tiles.
// Fragment trick (as skinny_gemm_kernel): lane (g = lane / 4, q = lane % 4) loads the 8 consecutive k values k0 + 8q .. 8q + 7 of
// weight row g and of activation rows g, g + 8 (one uint4 each): the same permutation of k on both operands of the two MMAs of a chunk.
template <int MT, int NT, bool A_SMEM, typename RowFn, typename EpiFn>
__device__ __forceinline__ void gemm_phase(int g_begin, int g_end, int K, const bf16* __restrict__ W, long long ldw, const bf16* A, long long lda,
                                           int M, float* part, RowFn row_of, EpiFn epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int nch = K >> 5;
  constexpr int PS = NT * 8 + 1;     // row stride of a partial tile (floats)
  constexpr int PT = MT * 16 * PS;   // floats per warp
  // chunks in flight per warp: weight fragments always; activation fragments too when they come from global memory (L2), which
  // costs 2 MT more 16-byte registers per chunk - shared-memory fragments are loaded right before their MMAs
  constexpr int U = A_SMEM ? 4 : (MT == 1 ? 4 : 2);
  for (int g0 = g_begin; g0 < g_end; g0 += kDecWarps) {
    const int ng = min(kDecWarps, g_end - g0);
    const int KS = kDecWarps / ng;
    const int grp = warp / KS, ks = warp - grp * KS;
    const bool active = grp < ng;
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    if (active) {
      const bf16* wr[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) wr[nt] = W + (size_t)(row_of(g0 + grp, nt) + g) * ldw + q * 8;
      const bf16* ar[MT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        ar[mt][0] = A + (size_t)(A_SMEM ? r0 : min(r0, M - 1)) * lda + q * 8;
        ar[mt][1] = A + (size_t)(A_SMEM ? r1 : min(r1, M - 1)) * lda + q * 8;
      }
      for (int c0 = ks; c0 < nch; c0 += KS * U) {
        uint4 wv[U][NT], av[A_SMEM ? 1 : U][MT][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + u * KS;
          const int cc = c < nch ? c : c0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) wv[u][nt] = __ldg(reinterpret_cast<const uint4*>(wr[nt] + cc * 32));
          if (!A_SMEM) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              av[u][mt][0] = ldcg_u4(ar[mt][0] + cc * 32);
              av[u][mt][1] = ldcg_u4(ar[mt][1] + cc * 32);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + u * KS;
          if (c < nch) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint4 lo, hi;
              if (A_SMEM) {
                lo = *reinterpret_cast<const uint4*>(ar[mt][0] + c * 32);
                hi = *reinterpret_cast<const uint4*>(ar[mt][1] + c * 32);
              } else {
                lo = av[u][mt][0];
                hi = av[u][mt][1];
              }
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                mma_16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, wv[u][nt].x, wv[u][nt].y);
                mma_16816(acc[mt][nt], lo.z, hi.z, lo.w, hi.w, wv[u][nt].z, wv[u][nt].w);
              }
            }
          }
        }
      }
      // C fragment: c0, c1 -> (row g, cols 2q, 2q + 1); c2, c3 -> (row g + 8, same cols)
      float* pw = part + warp * PT;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          pw[(mt * 16 + g) * PS + nt * 8 + 2 * q] = acc[mt][nt][0];
          pw[(mt * 16 + g) * PS + nt * 8 + 2 * q + 1] = acc[mt][nt][1];
          pw[(mt * 16 + 8 + g) * PS + nt * 8 + 2 * q] = acc[mt][nt][2];
          pw[(mt * 16 + 8 + g) * PS + nt * 8 + 2 * q + 1] = acc[mt][nt][3];
        }
    }
    __syncthreads();
    constexpr int per_grp = MT * 16 * 8;
    for (int i = threadIdx.x; i < ng * per_grp; i += kDecThreads) {
      const int gi = i / per_grp, r = i - gi * per_grp, m = r >> 3, c = r & 7;
      if (m >= M) continue;
      float v[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float s = 0.f;
        for (int k = 0; k < KS; ++k) s += part[(gi * KS + k) * PT + m * PS + nt * 8 + c];
        v[nt] = s;
      }
      epi(g0 + gi, m, c, v);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// attention of the new position over the KV cache, RoPE and cache write fused (attn_decode_group_kernel<true> arithmetic), with the
// keys of one (batch row, kv head) split into n_seg segments handled by different CTAs.  Every segment writes its un-normalised
// output, running max and sum; the CTA that finishes last (counter hand-off) merges the segments in index order and writes att.
// ------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void attn_item_range(const DecParams& p, int item, int pos, int& b, int& hk, int& j0, int& n, bool& has_new) {
  const int S = p.n_seg;
  const int seg = item % S, bh = item / S;
  hk = bh % p.hkv;
  b = bh / p.hkv;
  const int nkeys = pos + 1;
  const int per = (nkeys + S - 1) / S;
  j0 = min(seg * per, nkeys);
  const int j1 = min(nkeys, j0 + per);
  n = j1 - j0;
  has_new = n > 0 && j1 == nkeys;
}

__device__ __forceinline__ void attn_prefetch(const DecParams& p, int layer, int pos) {
  if (p.flags & 2) return;
  const int n_items = p.M * p.hkv * p.n_seg;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int b, hk, j0, n;
    bool has_new;
    attn_item_range(p, item, pos, b, hk, j0, n, has_new);
    const int nc = n - (has_new ? 1 : 0);
    const size_t off = (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64 + (size_t)j0 * 64;
    for (int i = threadIdx.x; i < 2 * nc; i += kDecThreads) {
      const bf16* base = (i < nc) ? p.kc : p.vc;
      const int r = (i < nc) ? i : i - nc;
      prefetch_l2(base + off + (size_t)r * 64);
    }
  }
}

__device__ __forceinline__ void attn_phase(const DecParams& p, int layer, int pos, float* scr, int* s_flag) {
  const int G = p.hq / p.hkv;
  const int S = p.n_seg;
  const int n_items = p.M * p.hkv * S;
  constexpr int SCS = kSegMax + 1;
  float* qs = scr;                          // [kMaxG][65]
  float* ml = qs + kMaxG * 65;              // [2][kMaxG]: max, sum
  float* rcs = ml + 2 * kMaxG;              // 32
  float* rsn = rcs + 32;                    // 32
  float* knew = rsn + 32;                   // 64
  float* vnew = knew + 64;                  // 64
  float* part_o = vnew + 64;                // [kDecWarps][kMaxG][64]
  float* sc = part_o + kDecWarps * kMaxG * 64;  // [kMaxG][SCS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ldq = (long long)(p.hq + 2 * p.hkv) * 64;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int b, hk, j0, n;
    bool has_new;
    attn_item_range(p, item, pos, b, hk, j0, n, has_new);
    const int bh = item / S;
    const int nc = n - (has_new ? 1 : 0);   // keys taken from the cache
    bf16* kbase = p.kc + (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64;
    bf16* vbase = p.vc + (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64;
    const bf16* row = p.qkv + (size_t)b * ldq;
    if (tid < 32) {   // same arithmetic as rope_kv_write_kernel
      const float inv_freq = exp2f(-(float)(2 * tid) / 64.0f * p.log2_theta);
      sincosf((float)pos * inv_freq, &rsn[tid], &rcs[tid]);
    }
    __syncthreads();
    for (int t = tid; t < G * 64; t += kDecThreads) {
      const int hh = t >> 6, e = t & 63, d = e & 31;
      const bf16* qr = row + (hk * G + hh) * 64;
      const float x0 = ldcg_bf16(qr + d), x1 = ldcg_bf16(qr + d + 32);
      qs[hh * 65 + e] = (e < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]) * 0.125f;
    }
    if (has_new) {   // the segment that ends at the new position rotates the new key and writes both new rows into the caches
      if (tid < 64) {
        const bf16* krow = row + (p.hq + hk) * 64;
        const int d = tid & 31;
        const float x0 = ldcg_bf16(krow + d), x1 = ldcg_bf16(krow + d + 32);
        const bf16 kr16 = __float2bfloat16(tid < 32 ? x0 * rcs[d] - x1 * rsn[d] : x1 * rcs[d] + x0 * rsn[d]);
        knew[tid] = __bfloat162float(kr16);   // the value later steps read back from the cache
        kbase[(size_t)pos * 64 + tid] = kr16;
      } else if (tid < 128) {
        const int t = tid - 64;
        const unsigned short vb = __ldcg(reinterpret_cast<const unsigned short*>(row + (p.hq + p.hkv + hk) * 64 + t));
        vnew[t] = __uint_as_float((uint32_t)vb << 16);
        reinterpret_cast<unsigned short*>(vbase)[(size_t)pos * 64 + t] = vb;
      }
    }
    __syncthreads();
    // scores: one (key, head) pair per thread and iteration; the 8 threads of a key share its row
    for (int idx = tid; idx < nc * 8; idx += kDecThreads) {
      const int jj = idx >> 3, h = idx & 7;
      if (h < G) {
        const bf16* kr = kbase + (size_t)(j0 + jj) * 64;
        const float* qh = qs + h * 65;
        float s = 0.f;
#pragma unroll
        for (int v8 = 0; v8 < 8; ++v8) {
          const uint4 u = ldcg_u4(kr + v8 * 8);
          const float2 a = unpack_bf16(u.x), bb = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
          const float* qq = qh + v8 * 8;
          s += a.x * qq[0] + a.y * qq[1] + bb.x * qq[2] + bb.y * qq[3] + c.x * qq[4] + c.y * qq[5] + d.x * qq[6] + d.y * qq[7];
        }
        sc[h * SCS + jj] = s;
      }
    }
    if (has_new && tid < G) {   // the newest key comes from shared memory
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 64; ++e) s += knew[e] * qs[tid * 65 + e];
      sc[tid * SCS + n - 1] = s;
    }
    __syncthreads();
    if (warp < G) {   // softmax statistics of this segment: warp h owns head h
      float* sh = sc + warp * SCS;
      float m = -INFINITY;
      for (int jj = lane; jj < n; jj += 32) m = fmaxf(m, sh[jj]);
      m = warp_max(m);
      float l = 0.f;
      for (int jj = lane; jj < n; jj += 32) {
        const float e = __expf(sh[jj] - m);
        sh[jj] = e;
        l += e;
      }
      l = warp_sum(l);
      if (lane == 0) { ml[warp] = m; ml[kMaxG + warp] = l; }
    }
    __syncthreads();
    // weighted V over the cached keys: each warp a contiguous chunk, lanes own dims (2 lane, 2 lane + 1), all heads of the group at once
    {
      const int chunk = (nc + kDecWarps - 1) / kDecWarps;
      const int ja = min(nc, warp * chunk), jb = min(nc, ja + chunk);
      const uint32_t* v32 = reinterpret_cast<const uint32_t*>(vbase) + (size_t)j0 * 32 + lane;
      float a0[kMaxG], a1[kMaxG];
#pragma unroll
      for (int h = 0; h < kMaxG; ++h) { a0[h] = 0.f; a1[h] = 0.f; }
      int jj = ja;
      for (; jj + 4 <= jb; jj += 4) {
        uint32_t wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = ldcg_u32(v32 + (size_t)(jj + u) * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = unpack_bf16(wv[u]);
#pragma unroll
          for (int h = 0; h < kMaxG; ++h) {
            if (h < G) {
              const float pj = sc[h * SCS + jj + u];
              a0[h] += pj * f.x;
              a1[h] += pj * f.y;
            }
          }
        }
      }
      for (; jj < jb; ++jj) {
        const float2 f = unpack_bf16(ldcg_u32(v32 + (size_t)jj * 32));
#pragma unroll
        for (int h = 0; h < kMaxG; ++h) {
          if (h < G) {
            const float pj = sc[h * SCS + jj];
            a0[h] += pj * f.x;
            a1[h] += pj * f.y;
          }
        }
      }
#pragma unroll
      for (int h = 0; h < kMaxG; ++h) {
        if (h < G) {
          part_o[(warp * kMaxG + h) * 64 + 2 * lane] = a0[h];
          part_o[(warp * kMaxG + h) * 64 + 2 * lane + 1] = a1[h];
        }
      }
    }
    __syncthreads();
    float* gp = p.attn_part + (size_t)item * kMaxG * kPartStride;
    for (int t = tid; t < G * 64; t += kDecThreads) {
      const int h = t >> 6, d = t & 63;
      float o = has_new ? sc[h * SCS + n - 1] * vnew[d] : 0.f;
#pragma unroll
      for (int w = 0; w < kDecWarps; ++w) o += part_o[(w * kMaxG + h) * 64 + d];
      gp[h * kPartStride + d] = o;
    }
    if (tid < G) {
      gp[tid * kPartStride + 64] = ml[tid];
      gp[tid * kPartStride + 65] = ml[kMaxG + tid];
    }
    __syncthreads();
    // hand-off: the add releases this CTA's partials (cumulative through the CTA barrier above) and acquires those of the segments
    // that arrived earlier
    if (tid == 0) *s_flag = (atom_add_acq_rel(p.attn_cnt + bh, 1u) == (unsigned)(S - 1)) ? 1 : 0;
    __syncthreads();
    if (*s_flag) {   // last segment of this (batch row, kv head) to finish: merge in segment order
      const float* g0p = p.attn_part + (size_t)bh * S * kMaxG * kPartStride;
      for (int t = tid; t < G * 64; t += kDecThreads) {
        const int h = t >> 6, d = t & 63;
        float L = 0.f, O = 0.f, mm = -INFINITY;
        for (int s0 = 0; s0 < S; s0 += 8) {   // 8 segments per round: 24 independent loads in flight, then the running-max merge in segment order
          float ms[8], ls[8], os[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int s = min(s0 + u, S - 1);
            const float* sp = g0p + ((size_t)s * kMaxG + h) * kPartStride;
            ms[u] = ldcg_f(sp + 64);
            ls[u] = ldcg_f(sp + 65);
            os[u] = ldcg_f(sp + d);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (s0 + u < S && ms[u] != -INFINITY) {   // -inf: empty segment
              const float mn = fmaxf(mm, ms[u]);
              const float wo = __expf(mm - mn), wn = __expf(ms[u] - mn);   // mm = -inf on the first segment: wo = 0
              L = L * wo + ls[u] * wn;
              O = O * wo + os[u] * wn;
              mm = mn;
            }
          }
        }
        p.att[(size_t)b * (p.hq * 64) + (hk * G + h) * 64 + d] = __float2bfloat16(L > 0.f ? O / L : 0.f);
      }
      if (tid == 0) p.attn_cnt[bh] = 0;   // next use is behind at least one grid barrier
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// final RMSNorm (already staged in As) x LM head with a running arg-max: warp w streams n8 tiles t0 + w, t0 + w + 16, ... of this
// CTA's contiguous tile range, K unsplit, the next tile of the warp prefetched to L2 while the current one is multiplied.
// Ties -> lowest index (argmax_kernel / torch.argmax).
// ------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void best_merge(float& bv, int& bi, float v, int i) {
  if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}

template <int MT>
__device__ __forceinline__ void lm_head_phase(const DecParams& p, const bf16* As, int SA, float* scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int D = p.D, V = p.V, nch = D >> 5;
  const int T = (V + 7) >> 3;
  int t0, t1;
  cta_share(T, t0, t1);
  const int lines = (D * 2 + 127) >> 7;
  float bv[MT][2];
  int bi[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) { bv[mt][0] = bv[mt][1] = -INFINITY; bi[mt][0] = bi[mt][1] = INT_MAX; }
  const bf16* ar[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    ar[mt][0] = As + (size_t)(mt * 16 + g) * SA + q * 8;
    ar[mt][1] = As + (size_t)(mt * 16 + 8 + g) * SA + q * 8;
  }
  if (t0 + warp < t1) {   // first tile of the warp
    for (int i = lane; i < 8 * lines; i += 32) {
      const int r = i / lines, ln = i - r * lines;
      prefetch_l2(reinterpret_cast<const uint8_t*>(p.lm_head + (size_t)min((t0 + warp) * 8 + r, V - 1) * D) + (size_t)ln * 128);
    }
  }
  constexpr int U = 7;
  for (int t = t0 + warp; t < t1; t += kDecWarps) {
    if (t + kDecWarps < t1) {
      for (int i = lane; i < 8 * lines; i += 32) {
        const int r = i / lines, ln = i - r * lines;
        prefetch_l2(reinterpret_cast<const uint8_t*>(p.lm_head + (size_t)min((t + kDecWarps) * 8 + r, V - 1) * D) + (size_t)ln * 128);
      }
    }
    const bf16* wr = p.lm_head + (size_t)min(t * 8 + g, V - 1) * D + q * 8;
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][e] = 0.f;
    for (int c0 = 0; c0 < nch; c0 += U) {
      uint4 wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u;
        wv[u] = __ldg(reinterpret_cast<const uint4*>(wr + (c < nch ? c : c0) * 32));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u;
        if (c < nch) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint4 lo = *reinterpret_cast<const uint4*>(ar[mt][0] + c * 32), hi = *reinterpret_cast<const uint4*>(ar[mt][1] + c * 32);
            mma_16816(acc[mt], lo.x, hi.x, lo.y, hi.y, wv[u].x, wv[u].y);
            mma_16816(acc[mt], lo.z, hi.z, lo.w, hi.w, wv[u].z, wv[u].w);
          }
        }
      }
    }
    const int n0 = t * 8 + 2 * q;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (n0 < V) { best_merge(bv[mt][0], bi[mt][0], acc[mt][0], n0); best_merge(bv[mt][1], bi[mt][1], acc[mt][2], n0); }
      if (n0 + 1 < V) { best_merge(bv[mt][0], bi[mt][0], acc[mt][1], n0 + 1); best_merge(bv[mt][1], bi[mt][1], acc[mt][3], n0 + 1); }
    }
  }
  // lanes of a quad hold different columns of the same rows
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv[mt][r], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[mt][r], o);
        best_merge(bv[mt][r], bi[mt][r], ov, oi);
      }
  float* wbv = scr;                                          // [kDecWarps][32]
  int* wbi = reinterpret_cast<int*>(scr + kDecWarps * 32);   // [kDecWarps][32]
  if (q == 0) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      wbv[warp * 32 + mt * 16 + g] = bv[mt][0];     wbi[warp * 32 + mt * 16 + g] = bi[mt][0];
      wbv[warp * 32 + mt * 16 + 8 + g] = bv[mt][1]; wbi[warp * 32 + mt * 16 + 8 + g] = bi[mt][1];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < p.M) {
    float v = -INFINITY;
    int i = INT_MAX;
    for (int w = 0; w < kDecWarps; ++w) best_merge(v, i, wbv[w * 32 + threadIdx.x], wbi[w * 32 + threadIdx.x]);
    p.best_val[(size_t)blockIdx.x * 32 + threadIdx.x] = v;
    p.best_idx[(size_t)blockIdx.x * 32 + threadIdx.x] = i;
  }
  __syncthreads();
}

// cross-CTA arg-max and the bookkeeping of Engine._generate_graphed.sample (llm.py:232-248): CTA 0, one warp per batch row
__device__ __forceinline__ void sample_phase(const DecParams& p, long long st) {
  if (blockIdx.x != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = warp; m < p.M; m += kDecWarps) {
    float v = -INFINITY;
    int i = INT_MAX;
    for (int c = lane; c < (int)gridDim.x; c += 32) best_merge(v, i, ldcg_f(p.best_val + (size_t)c * 32 + m), ldcg_i(p.best_idx + (size_t)c * 32 + m));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      best_merge(v, i, ov, oi);
    }
    if (lane == 0) {
      if (i == INT_MAX) i = 0;   // a row of NaN logits: keep the ids in range
      const unsigned char d = p.done[m];
      if (!d && st < p.max_new) p.sampled[(size_t)m * p.ld_sampled + st] = i;   // torch.where(done, cur, nxt)
      p.n_gen[m] += d ? 0 : 1;
      if (p.eos >= 0 && (long long)i == p.eos) p.done[m] = 1;
      p.nxt[m] = i;
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(kDecThreads, 1)
decode_loop_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ int s_abort, s_flag;
  const int D = p.D, I = p.I, M = p.M;
  const int SA = D + kAPad;
  bf16* As = reinterpret_cast<bf16*>(smem_raw);                                    // [MT * 16][SA] normalised rows
  float* scr = reinterpret_cast<float*>(smem_raw + (size_t)MT * 16 * SA * 2);       // phase scratch
  for (int i = threadIdx.x; i < MT * 16 * SA / 8; i += kDecThreads) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { s_abort = 0; s_flag = 0; }
  __syncthreads();
  unsigned epoch = 0;
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    p.trace[(size_t)gridDim.x * kTraceBarriers * 2] = (long long)globaltimer_ns();
    p.trace[(size_t)gridDim.x * kTraceBarriers * 2 + 1] = clock64();
  }
  const int pos0 = *p.pos;
  const long long step0 = *p.step;
  const int NQKV = (p.hq + 2 * p.hkv) * 64;
  const int T_qkv = NQKV / 8, T_o = D / 8, T_gu = I / 8, T_d = D / 8;
  int qa, qb, oa, ob, ga, gb, da, db;
  cta_share(T_qkv, qa, qb);
  cta_share(T_o, oa, ob);
  cta_share(T_gu, ga, gb);
  cta_share(T_d, da, db);
  int steps_done = 0;
  for (int t = 0; t < p.n_steps; ++t) {
    const int pos = pos0 + t;
    if (pos >= p.lmax) break;
    if (p.eos >= 0) {   // every sequence finished: stop (the host loop's `done.all()` test, llm.py:245, without the round trip)
      bool all = true;
      for (int m = 0; m < M; ++m) all = all && (__ldcg(p.done + m) != 0);
      if (all) break;
    }
    for (int l = 0; l < p.n_layers; ++l) {
      const DecLayerDev L = p.layers[l];
      // ---- 1: RMSNorm + q|k|v ----
      attn_prefetch(p, l, pos);
      stage_norm(p, As, SA, L.ln1, l == 0);
      {
        bf16* qkv = p.qkv;
        const bf16* bias = L.bqkv;
        gemm_phase<MT, 1, true>(qa, qb, D, L.qkv, D, As, SA, M, scr, [](int grp, int) { return grp * 8; },
                                [=](int grp, int m, int c, const float (&v)[1]) {
                                  const int n = grp * 8 + c;
                                  qkv[(size_t)m * NQKV + n] = __float2bfloat16(v[0] + __bfloat162float(bias[n]));
                                });
      }
      if (!grid_barrier(p, epoch, &s_abort)) return;
      // ---- 2: RoPE + KV write + attention ----
      prefetch_rows(p, L.o, D, oa * 8, (ob - oa) * 8, D);
      attn_phase(p, l, pos, scr, &s_flag);
      if (!grid_barrier(p, epoch, &s_abort)) return;
      // ---- 3: o projection into the residual stream ----
      for (int pt = ga; pt < gb; ++pt) {
        prefetch_rows(p, L.gu, D, gu_row(pt, 0), 8, D);
        prefetch_rows(p, L.gu, D, gu_row(pt, 1), 8, D);
      }
      {
        float* x = p.x;
        gemm_phase<MT, 1, false>(oa, ob, D, L.o, D, p.att, D, M, scr, [](int grp, int) { return grp * 8; },
                                 [=](int grp, int m, int c, const float (&v)[1]) {
                                   float* xp = x + (size_t)m * D + grp * 8 + c;
                                   *xp = ldcg_f(xp) + v[0];
                                 });
      }
      if (!grid_barrier(p, epoch, &s_abort)) return;
      // ---- 4: RMSNorm + gate|up + SwiGLU ----
      prefetch_rows(p, L.d, I, da * 8, (db - da) * 8, I);
      stage_norm(p, As, SA, L.ln2, false);
      {
        bf16* act = p.act;
        gemm_phase<MT, 2, true>(ga, gb, D, L.gu, D, As, SA, M, scr, [](int grp, int nt) { return gu_row(grp, nt); },
                                [=](int grp, int m, int c, const float (&v)[2]) {
                                  act[(size_t)m * I + grp * 8 + c] = __float2bfloat16(silu(v[0]) * v[1]);
                                });
      }
      if (!grid_barrier(p, epoch, &s_abort)) return;
      // ---- 5: down projection into the residual stream ----
      if (l + 1 < p.n_layers) prefetch_rows(p, p.layers[l + 1].qkv, D, qa * 8, (qb - qa) * 8, D);
      {
        float* x = p.x;
        gemm_phase<MT, 1, false>(da, db, I, L.d, I, p.act, I, M, scr, [](int grp, int) { return grp * 8; },
                                 [=](int grp, int m, int c, const float (&v)[1]) {
                                   float* xp = x + (size_t)m * D + grp * 8 + c;
                                   *xp = ldcg_f(xp) + v[0];
                                 });
      }
      if (!grid_barrier(p, epoch, &s_abort)) return;
    }
    // ---- final norm + LM head + arg-max ----
    prefetch_rows(p, p.layers[0].qkv, D, qa * 8, (qb - qa) * 8, D);
    stage_norm(p, As, SA, p.norm_w, false);
    lm_head_phase<MT>(p, As, SA, scr);
    if (!grid_barrier(p, epoch, &s_abort)) return;
    sample_phase(p, step0 + t);
    steps_done = t + 1;
    if (!grid_barrier(p, epoch, &s_abort)) return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *p.pos = pos0 + steps_done;
    *p.step = step0 + steps_done;
    if (p.trace) {
      p.trace[(size_t)gridDim.x * kTraceBarriers * 2 + 2] = (long long)globaltimer_ns();
      p.trace[(size_t)gridDim.x * kTraceBarriers * 2 + 3] = clock64();
    }
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DecWorkspace {
  size_t x, qkv, att, act, attn_part, attn_cnt, best_val, best_idx, bar, total;
};
DecWorkspace decode_layout(int batch, int hidden, int mlp, int hq, int hkv, int grid) {
  DecWorkspace w;
  size_t o = 0;
  w.x = o; o = align_up(o + (size_t)batch * hidden * 4, 256);
  w.qkv = o; o = align_up(o + (size_t)batch * (hq + 2 * hkv) * 64 * 2, 256);
  w.att = o; o = align_up(o + (size_t)batch * hq * 64 * 2, 256);
  w.act = o; o = align_up(o + (size_t)batch * mlp * 2, 256);
  // batch * hkv * n_seg <= max(grid, batch * hkv) work items
  const size_t items = (size_t)((batch * hkv > grid) ? batch * hkv : grid);
  w.attn_part = o; o = align_up(o + items * kMaxG * kPartStride * 4, 256);
  w.attn_cnt = o; o = align_up(o + (size_t)batch * hkv * 4, 256);
  w.best_val = o; o = align_up(o + (size_t)grid * 32 * 4, 256);
  w.best_idx = o; o = align_up(o + (size_t)grid * 32 * 4, 256);
  w.bar = o; o = align_up(o + 256, 256);
  w.total = o;
  return w;
}

size_t decode_smem_bytes(int mt, int hidden) {
  const size_t as = (size_t)mt * 16 * (hidden + kAPad) * 2;
  const size_t gemm = (size_t)kDecWarps * mt * 16 * 17 * 4;
  const size_t attn = (size_t)(kMaxG * 65 + 2 * kMaxG + 32 + 32 + 64 + 64 + kDecWarps * kMaxG * 64 + kMaxG * (kSegMax + 1)) * 4;
  const size_t lm = (size_t)kDecWarps * 32 * 8;
  size_t scr = gemm > attn ? gemm : attn;
  if (lm > scr) scr = lm;
  return as + align_up(scr, 16);
}

}  // namespace

extern "C" size_t slb_decode_workspace_bytes(int batch, int hidden, int mlp, int hq, int hkv) {
  if (batch <= 0 || hidden <= 0 || mlp <= 0 || hq <= 0 || hkv <= 0) return 0;
  return decode_layout(batch, hidden, mlp, hq, hkv, slb_num_sms()).total;
}

extern "C" int slb_decode_loop(const slb_decode_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SLB_CHECK_ARG(a != nullptr, "decode_loop: null args");
  SLB_CHECK_ARG(a->layers && a->n_layers > 0 && a->emb && a->norm_w && a->lm_head && a->kcache && a->vcache, "decode_loop: null weights / caches");
  SLB_CHECK_ARG(a->pos && a->nxt && a->sampled && a->step && a->done && a->n_gen && a->workspace && a->status, "decode_loop: null state pointers");
  SLB_CHECK_ARG(a->batch >= 1 && a->batch <= 32, "decode_loop: batch must be 1..32 (got %d)", a->batch);
  SLB_CHECK_ARG(a->hidden % 128 == 0 && a->hidden <= 1024, "decode_loop: hidden must be a multiple of 128, <= 1024 (got %d)", a->hidden);
  SLB_CHECK_ARG(a->mlp % 128 == 0, "decode_loop: mlp width must be a multiple of 128 (got %d)", a->mlp);
  SLB_CHECK_ARG(a->hkv > 0 && a->hq % a->hkv == 0 && a->hq / a->hkv <= kMaxG && a->hq * 64 == a->hidden,
                "decode_loop: heads %d / %d with head_dim 64 must match hidden %d (group <= %d)", a->hq, a->hkv, a->hidden, kMaxG);
  SLB_CHECK_ARG(a->vocab > 0 && a->emb_rows > 0 && a->lmax > 0 && a->n_steps >= 0 && a->max_new > 0 && a->rope_theta > 1.f, "decode_loop: bad sizes");
  if (a->n_steps == 0) return SLB_OK;
  const int grid = slb_num_sms();
  const DecWorkspace w = decode_layout(a->batch, a->hidden, a->mlp, a->hq, a->hkv, grid);
  SLB_CHECK_ARG(a->workspace_bytes >= w.total, "decode_loop: workspace of %zu bytes, need %zu", a->workspace_bytes, w.total);
  SLB_CHECK_ARG(((uintptr_t)a->workspace & 255) == 0, "decode_loop: workspace must be 256-byte aligned");
  DecParams p;
  p.layers = reinterpret_cast<const DecLayerDev*>(a->layers);
  p.n_layers = a->n_layers; p.M = a->batch; p.D = a->hidden; p.I = a->mlp; p.V = a->vocab; p.hq = a->hq; p.hkv = a->hkv;
  p.lmax = a->lmax; p.n_steps = a->n_steps; p.max_new = a->max_new;
  // key segments per (batch row, kv head): spread small batches over the CTAs, at least ~128 keys per segment, at most kSegMax
  int seg = grid / (a->batch * a->hkv);
  const int by_len = a->lmax / 128 > 0 ? a->lmax / 128 : 1;
  if (seg > by_len) seg = by_len;
  if (seg < 1) seg = 1;
  SLB_CHECK_ARG(ceil_div(a->lmax, seg) <= kSegMax, "decode_loop: lmax=%d too long (%d segments of at most %d keys)", a->lmax, seg, kSegMax);
  p.n_seg = seg;
  p.emb_rows = a->emb_rows; p.eos = a->eos; p.ld_sampled = a->ld_sampled;
  p.layer_stride = (long long)a->batch * a->hkv * a->lmax * 64;
  p.emb = (const bf16*)a->emb; p.norm_w = (const bf16*)a->norm_w; p.lm_head = (const bf16*)a->lm_head;
  p.kc = (bf16*)a->kcache; p.vc = (bf16*)a->vcache;
  p.pos = a->pos; p.nxt = (long long*)a->nxt; p.sampled = (long long*)a->sampled; p.step = (long long*)a->step;
  p.done = a->done; p.n_gen = (long long*)a->n_gen;
  p.log2_theta = log2f(a->rope_theta); p.eps = a->rms_eps;
  uint8_t* ws = (uint8_t*)a->workspace;
  p.x = (float*)(ws + w.x); p.qkv = (bf16*)(ws + w.qkv); p.att = (bf16*)(ws + w.att); p.act = (bf16*)(ws + w.act);
  p.attn_part = (float*)(ws + w.attn_part); p.attn_cnt = (unsigned*)(ws + w.attn_cnt);
  p.best_val = (float*)(ws + w.best_val); p.best_idx = (int*)(ws + w.best_idx);
  p.bar = (unsigned*)(ws + w.bar);
  p.err = (unsigned*)a->status;
  {
    const char* e = getenv("SLB_DECODE_FLAGS");
    p.flags = e ? atoi(e) : 0;
    long long* tr = slb_debug_trace_ptr();
    p.trace = tr ? tr + 8192 : nullptr;   // the first entries of the debug buffer belong to the GEMM / attention timelines
  }
  const int mt = a->batch <= 16 ? 1 : 2;
  const size_t smem = decode_smem_bytes(mt, a->hidden);
  auto kern = mt == 1 ? decode_loop_kernel<1> : decode_loop_kernel<2>;
  static bool attr_set[2] = {false, false};
  static size_t smem_set[2] = {0, 0};
  if (!attr_set[mt - 1] || smem > smem_set[mt - 1]) {
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    SLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kDecThreads, smem));
    SLB_CHECK_ARG(nb >= 1, "decode_loop: the persistent CTA (%zu bytes of shared memory) does not fit on an SM", smem);
    attr_set[mt - 1] = true;
    smem_set[mt - 1] = smem;
  }
  // barrier counter, attention hand-off counters and the error flag start at zero for every launch
  SLB_CUDA(cudaMemsetAsync(p.bar, 0, 256, stream));
  SLB_CUDA(cudaMemsetAsync(p.attn_cnt, 0, (size_t)a->batch * a->hkv * 4, stream));
  SLB_CUDA(cudaMemsetAsync(p.err, 0, 4, stream));
  static int coop = -1;
  if (coop < 0) { const char* e = getenv("SLB_DECODE_COOP"); coop = (e && atoi(e) == 0) ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kDecThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = coop ? 1 : 0;   // SLB_DECODE_COOP=0: plain launch (grid = SM count, one CTA per SM: co-resident once the stream's earlier work drains)
  SLB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return SLB_OK;
}
