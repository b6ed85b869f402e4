// Greedy decode loop of the Qwen2 decoder as ONE persistent kernel (llm.py:217-248 per generated token: embed the last sampled
// token, 24 decoder layers against the KV cache, final RMSNorm, LM head, argmax, EOS bookkeeping).
//
// STATUS: opt-in (Engine.decode_mega / SLB_DECODE_MEGA=1).  Correct - tokens identical to the per-kernel chain and to the fp32 oracle,
// tests/test_decode_gpu.py - but not faster: 1.14 ms per token at batch 1 against 0.93 ms for the PDL-chained kernels, level with
// them at batch 32 (1-CTA-per-SM timelines of every version under profiles/r02_trace_decode_v*.log, tools/trace_decode.py).
//
// Idea: a decode step is ~1 GB of weights streamed once (0.15 ms at the HBM roofline) but, as a chain of ~100 dependent launches,
// it takes 0.9 ms (batch 1) to 1.9 ms (batch 32) per token.  Here one CTA per SM stays resident for the whole generation; the
// phases of a layer are separated by grid-wide barriers (measured 1.2-1.4 us each) instead of kernel boundaries, every phase's
// weights (and the cached K / V rows of the attention phase) are copied into shared memory by bulk asynchronous copies issued
// BEFORE the barrier that precedes the phase - they do not depend on other CTAs - so that only the activation load is left on the
// critical path after the barrier, and the token loop, the position counter and the EOS test live on the device (no host round
// trip per token).
// What the timelines show: with one CTA of 16 warps per SM every dependent step inside a phase (an L2 round trip of coherent loads
// ~1.3 us, a shared-memory reduction + __syncthreads 0.3-0.5 us, the issue of one bulk copy ~0.3 us, the counter hand-off of the
// split attention ~1 us + 2 us merge) is exposed; a layer costs 5 x 1.4 us of barriers + 34 us of such chains, where the launch
// chain, whose kernels prefetch their weights under the predecessor's tail (programmatic dependent launch) and run many CTAs per SM,
// needs 38 us for the same layer.  The LM head streams at 2.0 TB/s here (272 MB in 138 us; 71 us in the first version, before the
// 200 KB of shared memory left no L1 for the weight loads), the chain's skinny GEMM + arg-max take 178 us.
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
constexpr int kSegMax = 320;    // keys per attention work item (their K and V rows are staged in shared memory: 256 B per key)
__host__ __device__ constexpr int gu_slots(int mt) { return mt == 1 ? 5 : 3; }   // gate|up tile pairs staged in shared memory per round
constexpr int kAPad = 32;       // bf16 elements of row padding of the staged rows: row stride = 64 B mod 128 B, conflict-free uint4 fragment loads
constexpr int kPartStride = 66; // floats per (item, head) attention partial: 64 outputs, running max, sum

struct DecLayerDev { const bf16 *qkv, *bqkv, *o, *gu, *d, *ln1, *ln2; };   // = slb_decode_layer

struct DecParams {
  const DecLayerDev* layers;
  int n_layers, M, D, I, V, hq, hkv, lmax, n_steps, max_new, n_seg, wbuf_bytes, scr_bytes;
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
constexpr int kFineStamps = 32;   // per-CTA stamps inside the phases of (second token, second layer)
__device__ __forceinline__ void fine_stamp(long long* ft, int k) {
  if (ft && threadIdx.x == 0) ft[k] = clock64();
}

// ---- coherent (L2) loads of data produced by other CTAs during this kernel ----
__device__ __forceinline__ uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg_i(const int* p) { return __ldcg(p); }
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
__device__ __forceinline__ bool grid_barrier(const DecParams& p, unsigned epoch, int* s_abort) {
  fence_proxy_async_smem();   // this phase's generic shared-memory writes vs the bulk copies that will overwrite those regions
  __syncthreads();
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

// ---- bulk asynchronous global -> shared copies (one instruction per row, completion on an mbarrier; no registers, no per-thread
// issue loops: with per-thread cp.async the issue of a 5-pair gate|up prefetch alone cost ~5 us of the attention phase) ----
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// One use of a staging barrier (arrival count 1): lane 0 of warp 0 announces the byte count, the lanes of warp 0 issue one copy per
// row (row r: src + r * src_stride -> dst + r * dst_stride).  Callers run this after a __syncthreads that retired every reader of
// the destination.  Out of line: the kernel's instruction footprint per layer must stay inside the instruction cache (the fully
// inlined version was 195 KB of SASS and ran every phase 3-4 x slower than its instruction count explains).
__device__ __forceinline__ void stage_strided(uint64_t* bar, int nrows, uint32_t row_bytes, const uint8_t* src, long long src_stride, uint8_t* dst,
                                           int dst_stride) {
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)nrows * row_bytes);
    __syncwarp();
    for (int r = threadIdx.x; r < nrows; r += 32) bulk_g2s(dst + (size_t)r * dst_stride, src + (size_t)r * src_stride, row_bytes, bar);
  }
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

// one n8 weight tile (8 rows x K bf16) -> shared memory, dense (rows ws = 2 K bytes apart)
__device__ __forceinline__ void stage_tile(uint64_t* bar, uint8_t* dst, const bf16* W, long long ldw, int row0, int K, int ws) {
  // the 8 rows are contiguous in global memory (ldw == K) and dense in shared memory (ws == 2 K): ONE copy - issuing a bulk copy
  // costs ~0.3 us of its warp, row-wise staging of a tile was slower than the 2-way bank conflict of the dense layout
  stage_strided(bar, 1, (uint32_t)K * 16, reinterpret_cast<const uint8_t*>(W + (size_t)row0 * ldw), 0, dst, 0);
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

// gate|up tile pairs [pa, pb) (at most gu_slots(MT)) -> the slots of the G region (tile 2 i = gate, 2 i + 1 = up of pair pa + i)
__device__ __forceinline__ void stage_gu(uint64_t* bar, uint8_t* gbuf, const bf16* W, int D, int ws, int pa, int pb) {
  if (threadIdx.x < 32) {   // one copy per n8 tile (8 contiguous weight rows), issued by different lanes
    const int ntiles = 2 * (pb - pa);
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)ntiles * 8 * D * 2);
    __syncwarp();
    if ((int)threadIdx.x < ntiles) {
      const int tile = threadIdx.x;
      bulk_g2s(gbuf + (size_t)tile * 8 * ws, W + (size_t)gu_row(pa + (tile >> 1), tile & 1) * D, (uint32_t)8 * D * 2, bar);
    }
  }
}

// RMSNorm of the M residual rows into shared memory (bf16, one rounding; x * rstd * w as norm_fwd_kernel): one warp per row, 16 rows
// per pass (one pass up to batch 16, two at 32); the norm weights travel to shared memory (`wn`, D bf16) in the same round trip as
// the first rows.  from_emb: the rows are the embeddings of the last sampled tokens (start of a step); CTA 0 then also seeds the
// fp32 residual stream.
__device__ __forceinline__ void stage_norm(const DecParams& p, bf16* As, int SA, const bf16* __restrict__ w, bool from_emb, float* scr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D;
  bf16* wn = reinterpret_cast<bf16*>(scr);
  for (int i = threadIdx.x; i < (D >> 3); i += kDecThreads) reinterpret_cast<uint4*>(wn)[i] = __ldg(reinterpret_cast<const uint4*>(w) + i);
  for (int m0 = 0; m0 < p.M; m0 += kDecWarps) {
    const int m = m0 + warp;
    const bool on = m < p.M;
    float v[8][4];
    float rstd = 0.f;
    if (on) {
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
        }
      }
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (i * 32 + lane) * 4;
        if (k < D) ss += v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2] + v[i][3] * v[i][3];
      }
      ss = warp_sum(ss);
      rstd = rsqrtf(ss / D + p.eps);
    }
    if (m0 == 0) __syncthreads();   // wn complete
    if (on) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = (i * 32 + lane) * 4;
        if (k < D) {
          const uint2 wu = *reinterpret_cast<const uint2*>(wn + k);
          const float2 wa = unpack_bf16(wu.x), wb = unpack_bf16(wu.y);
          uint2 o;
          o.x = pack_bf16(v[i][0] * rstd * wa.x, v[i][1] * rstd * wa.y);
          o.y = pack_bf16(v[i][2] * rstd * wb.x, v[i][3] * rstd * wb.y);
          *reinterpret_cast<uint2*>(As + (size_t)m * SA + k) = o;
        }
      }
    }
  }
  __syncthreads();
}

// One round of the skinny product: `ng` (<= 16) weight-tile groups, already staged (or in flight: bulk copies) in shared memory,
// against the M activation rows.  A group = NT n8 tiles (8 weight rows each); group i, tile nt sits at wb + (i NT + nt) 8 ws, rows ws
// bytes apart.  The 16 warps are split into KS = 16 / ng K-slices per group (interleaved 32-element chunks); partial tiles are summed
// through shared memory in a fixed order.  Two out-of-line halves (one copy each per <MT, NT>, shared by all phases: instruction
// footprint): gemm_mma waits for the staged operands (mbarriers), multiplies and parks the partial tiles; the caller then issues the
// next bulk copies (the staged weights are consumed); gemm_epi reduces and applies the epilogue:
//   kEpiQkv      out bf16 [M, ldo]: acc + bias[n]            kEpiResidual  out fp32 [M, ldo]: out += acc (the residual stream)
//   kEpiSwiglu   NT = 2 (gate, up tiles): out bf16 [M, ldo]: silu(gate) * up
// The epilogue operand of the first two (bias / current residual value) is requested at the start of gemm_mma and parked in shared
// memory: off the critical path.  Activation rows: amode 0 = the normalised rows in shared memory (stride a_stride bytes, rows >= M
// zero), 1 = global memory through L2 (ld.global.cg; U chunks in flight per warp), 2 = rows staged in shared memory by bulk copies
// (row index clamped to M - 1).
// Fragment trick (as skinny_gemm_kernel): lane (g = lane / 4, q = lane % 4) loads the 8 consecutive k values k0 + 8q .. 8q + 7 of
// weight row g and of activation rows g, g + 8 (one uint4 each): the same permutation of k on both operands of the two MMAs of a chunk.
enum { kEpiQkv = 0, kEpiResidual = 1, kEpiSwiglu = 2 };
struct GemmCfg {
  int ng, K, ws, M, amode, epi;
  const uint8_t* wb;
  const uint8_t* A; long long a_stride;
  float* part; float* pvs;
  uint64_t* bar0; uint32_t ph0; uint64_t* bar1; uint32_t ph1;   // staging barriers to wait for (bar1 may be null)
  void* out; long long ldo; int n0;                            // group i covers output columns n0 + 8 i .. n0 + 8 i + 7
  const bf16* bias;
  long long* ft;
};

template <int MT, int NT>
__device__ __forceinline__ void gemm_mma(const GemmCfg& c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int nch = c.K >> 5, M = c.M, ng = c.ng;
  constexpr int PS = NT * 8 + 1;     // row stride of a partial tile (floats)
  constexpr int PT = MT * 16 * PS;   // floats per warp
  constexpr int per_grp = MT * 16 * 8;
  float pv = 0.f;
  if (c.epi != kEpiSwiglu && (int)threadIdx.x < per_grp) {   // single-group rounds only (ng == 1): thread i owns output (m, col) = (i / 8, i % 8)
    const int m = threadIdx.x >> 3, col = threadIdx.x & 7;
    if (m < M) {
      if (c.epi == kEpiQkv) pv = __bfloat162float(c.bias[c.n0 + col]);
      else pv = ldcg_f(reinterpret_cast<const float*>(c.out) + (size_t)m * c.ldo + c.n0 + col);
    }
  }
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
  const uint8_t* ar[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    ar[mt][0] = c.A + (size_t)(c.amode == 0 ? r0 : min(r0, M - 1)) * c.a_stride + q * 16;
    ar[mt][1] = c.A + (size_t)(c.amode == 0 ? r1 : min(r1, M - 1)) * c.a_stride + q * 16;
  }
  constexpr int U = MT == 1 ? 4 : 2;
  uint4 av[U][MT][2];
  if (c.amode == 1 && active) {   // first round of activation fragments: requested before the wait for the staged weights
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ch = ks + u * KS;
      const int cc = ch < nch ? ch : ks;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        av[u][mt][0] = ldcg_u4(ar[mt][0] + (size_t)cc * 64);
        av[u][mt][1] = ldcg_u4(ar[mt][1] + (size_t)cc * 64);
      }
    }
  }
  mbar_wait(c.bar0, c.ph0);   // the staged operands of this round have landed
  if (c.bar1) mbar_wait(c.bar1, c.ph1);
  __syncthreads();
  fine_stamp(c.ft, 0);
  if (active) {
    const uint8_t* wr[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) wr[nt] = c.wb + (size_t)(grp * NT + nt) * 8 * c.ws + (size_t)g * c.ws + q * 16;
    if (c.amode != 1) {
#pragma unroll 2
      for (int ch = ks; ch < nch; ch += KS) {
        uint4 wv[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) wv[nt] = *reinterpret_cast<const uint4*>(wr[nt] + (size_t)ch * 64);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint4 lo = *reinterpret_cast<const uint4*>(ar[mt][0] + (size_t)ch * 64);
          const uint4 hi = *reinterpret_cast<const uint4*>(ar[mt][1] + (size_t)ch * 64);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            mma_16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, wv[nt].x, wv[nt].y);
            mma_16816(acc[mt][nt], lo.z, hi.z, lo.w, hi.w, wv[nt].z, wv[nt].w);
          }
        }
      }
    } else {
      for (int c0 = ks; c0 < nch; c0 += KS * U) {
        if (c0 != ks) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int ch = c0 + u * KS;
            const int cc = ch < nch ? ch : c0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              av[u][mt][0] = ldcg_u4(ar[mt][0] + (size_t)cc * 64);
              av[u][mt][1] = ldcg_u4(ar[mt][1] + (size_t)cc * 64);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ch = c0 + u * KS;
          if (ch < nch) {
            uint4 wv[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) wv[nt] = *reinterpret_cast<const uint4*>(wr[nt] + (size_t)ch * 64);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                mma_16816(acc[mt][nt], av[u][mt][0].x, av[u][mt][1].x, av[u][mt][0].y, av[u][mt][1].y, wv[nt].x, wv[nt].y);
                mma_16816(acc[mt][nt], av[u][mt][0].z, av[u][mt][1].z, av[u][mt][0].w, av[u][mt][1].w, wv[nt].z, wv[nt].w);
              }
          }
        }
      }
    }
    // C fragment: c0, c1 -> (row g, cols 2q, 2q + 1); c2, c3 -> (row g + 8, same cols)
    float* pw = c.part + warp * PT;
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
  if ((int)threadIdx.x < per_grp) c.pvs[threadIdx.x] = pv;
  __syncthreads();
  fine_stamp(c.ft, 1);
}

template <int MT, int NT>
__device__ __forceinline__ void gemm_epi(const GemmCfg& c) {
  constexpr int PS = NT * 8 + 1, PT = MT * 16 * PS, per_grp = MT * 16 * 8;
  const int ng = c.ng, M = c.M, KS = kDecWarps / ng;
  fine_stamp(c.ft, 2);
  for (int i = threadIdx.x; i < ng * per_grp; i += kDecThreads) {
    const int gi = i / per_grp, rr = i - gi * per_grp, m = rr >> 3, col = rr & 7;
    if (m >= M) continue;
    float v[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float s = 0.f;
      for (int k = 0; k < KS; ++k) s += c.part[(gi * KS + k) * PT + m * PS + nt * 8 + col];
      v[nt] = s;
    }
    const size_t o = (size_t)m * c.ldo + c.n0 + gi * 8 + col;
    if (c.epi == kEpiSwiglu) reinterpret_cast<bf16*>(c.out)[o] = __float2bfloat16(silu(v[0]) * v[NT - 1]);
    else if (c.epi == kEpiQkv) reinterpret_cast<bf16*>(c.out)[o] = __float2bfloat16(v[0] + c.pvs[i]);
    else reinterpret_cast<float*>(c.out)[o] = c.pvs[i] + v[0];
  }
  __syncthreads();
  fine_stamp(c.ft, 3);
}

// ------------------------------------------------------------------------------------------------------------------------------
// attention of the new position over the KV cache, RoPE and cache write fused (attn_decode_group_kernel<true> arithmetic), with the
// keys of one (batch row, kv head) split into n_seg segments handled by different CTAs.  The cached K / V rows of a segment are
// staged in shared memory by bulk copies - for a CTA's first work item already during the q|k|v phase, i.e. before the grid barrier:
// they were written by earlier steps.  Every segment writes its un-normalised output, running max and sum; the CTA that finishes
// last (counter hand-off) merges the segments in index order and writes att.
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

// bulk copies of an item's cached key / value rows into kv (K rows at 0, V rows at kSegMax * 128 bytes): two copies, one barrier use
// (also announced when the item has no cached keys, so that every issue is matched by exactly one wait)
__device__ __forceinline__ void attn_stage(const DecParams& p, uint64_t* bar, int layer, int pos, int item, uint8_t* kv) {
  if (threadIdx.x >= 32) return;
  int b, hk, j0, n;
  bool has_new;
  attn_item_range(p, item, pos, b, hk, j0, n, has_new);
  const int nc = n - (has_new ? 1 : 0);
  const size_t off = (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64 + (size_t)j0 * 64;
  if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)nc * 256);
  __syncwarp();
  if (nc > 0) {
    if (threadIdx.x == 0) bulk_g2s(kv, p.kc + off, (uint32_t)nc * 128, bar);
    if (threadIdx.x == 1) bulk_g2s(kv + (size_t)kSegMax * 128, p.vc + off, (uint32_t)nc * 128, bar);
  }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(smem_row)));
}

// Both products of a work item run on mma.sync with the q heads of the GQA group as the M operand (7 of 16 rows used): the CUDA-core
// version spent 3.3 us on the scores and 2.9 us on the weighted V of ~100 keys - dependent FMA / shared-memory chains at one CTA of
// 16 warps per SM (profiles/r02_trace_decode_v3.log).  q (rotated, scaled by 1/8: exact) and the probabilities are rounded to bf16
// for the MMAs, as the prefill / teacher-forced attention kernels do.
//   scores  S[head][key]  = Q[16 x 64] K^T: key tiles of 8 over the warps, fragments by 16-byte loads with the same k permutation
//           on both operands (as gemm_round)
//   output  O[head][dim] += P[16 x 16 keys] V[16 keys x 64]: 16-key steps over the warps, V fragments by ldmatrix.trans, per-warp
//           partial outputs summed through shared memory
__device__ __forceinline__ void attn_phase(const DecParams& p, int layer, int pos, float* scr, float* part_o, uint8_t* kv, uint64_t* bar_g,
                                        uint32_t& ph_g, int* s_flag, const float* rope, long long* ft) {
  const int G = p.hq / p.hkv;
  const int S = p.n_seg;
  const int n_items = p.M * p.hkv * S;
  constexpr int SCS = kSegMax + 2;          // even: float2 loads of probability pairs
  constexpr int QS = 72;                    // bf16 row stride of the q tile (144 B: conflict-free 16-byte fragment loads)
  bf16* Qs = reinterpret_cast<bf16*>(scr);  // [16][QS], rows >= G zero
  float* ml = scr + 16 * QS / 2;            // [2][kMaxG]: max, sum
  float* pn = ml + 2 * kMaxG;               // [kMaxG] probability of the new key
  float* snew = pn + kMaxG;                 // [kMaxG] its score
  const float* rcs = rope;                  // cos / sin of the step's position (computed once per step)
  const float* rsn = rope + 32;
  float* knew = snew + kMaxG;               // 64
  float* vnew = knew + 64;                  // 64
  float* sc = vnew + 64;                    // [kMaxG][SCS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const long long ldq = (long long)(p.hq + 2 * p.hkv) * 64;
  const bf16* Ks = reinterpret_cast<const bf16*>(kv);
  uint8_t* Vs = kv + (size_t)kSegMax * 128;
  if (blockIdx.x < n_items)
    for (int i = tid; i < 16 * QS / 2; i += kDecThreads) reinterpret_cast<uint32_t*>(Qs)[i] = 0u;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int b, hk, j0, n;
    bool has_new;
    attn_item_range(p, item, pos, b, hk, j0, n, has_new);
    const int bh = item / S;
    const int nc = n - (has_new ? 1 : 0);   // keys taken from the cache
    const int nc16 = (nc + 15) & ~15;
    if (item != (int)blockIdx.x) attn_stage(p, bar_g, layer, pos, item, kv);   // a second item of this CTA: could not be staged ahead
    bf16* kbase = p.kc + (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64;
    bf16* vbase = p.vc + (size_t)layer * p.layer_stride + ((size_t)b * p.hkv + hk) * p.lmax * 64;
    const bf16* row = p.qkv + (size_t)b * ldq;
    // the new position's q (all heads of the group), k and v: one round trip
    float qx0[2], qx1[2], kx0 = 0.f, kx1 = 0.f;
    unsigned short vb = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int t = tid + r * kDecThreads;
      qx0[r] = qx1[r] = 0.f;
      if (t < G * 64) {
        const int hh = t >> 6, d = t & 31;
        const bf16* qr = row + (hk * G + hh) * 64;
        qx0[r] = ldcg_bf16(qr + d);
        qx1[r] = ldcg_bf16(qr + d + 32);
      }
    }
    if (has_new) {
      if (tid < 64) {
        const bf16* krow = row + (p.hq + hk) * 64;
        kx0 = ldcg_bf16(krow + (tid & 31));
        kx1 = ldcg_bf16(krow + (tid & 31) + 32);
      } else if (tid < 128) {
        vb = __ldcg(reinterpret_cast<const unsigned short*>(row + (p.hq + p.hkv + hk) * 64 + (tid - 64)));
      }
    }
    __syncthreads();   // (q tile zero fill; the previous item's readers)
    fine_stamp(ft, 1);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int t = tid + r * kDecThreads;
      if (t < G * 64) {
        const int hh = t >> 6, e = t & 63, d = e & 31;
        Qs[hh * QS + e] = __float2bfloat16((e < 32 ? qx0[r] * rcs[d] - qx1[r] * rsn[d] : qx1[r] * rcs[d] + qx0[r] * rsn[d]) * 0.125f);
      }
    }
    if (has_new) {   // the segment that ends at the new position rotates the new key and writes both new rows into the caches
      if (tid < 64) {
        const int d = tid & 31;
        const bf16 kr16 = __float2bfloat16(tid < 32 ? kx0 * rcs[d] - kx1 * rsn[d] : kx1 * rcs[d] + kx0 * rsn[d]);
        knew[tid] = __bfloat162float(kr16);   // the value later steps read back from the cache
        kbase[(size_t)pos * 64 + tid] = kr16;
      } else if (tid < 128) {
        const int t = tid - 64;
        vnew[t] = __uint_as_float((uint32_t)vb << 16);
        reinterpret_cast<unsigned short*>(vbase)[(size_t)pos * 64 + t] = vb;
      }
    }
    // value rows nc .. nc16 - 1 enter the last 16-key step with probability 0: they must be finite
    for (int i = tid; i < (nc16 - nc) * 8; i += kDecThreads) reinterpret_cast<uint4*>(Vs + (size_t)nc * 128)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    mbar_wait(bar_g, ph_g);
    ph_g ^= 1u;
    __syncthreads();
    fine_stamp(ft, 2);
    // ---- scores ----
    for (int tile = warp; tile * 8 < nc; tile += kDecWarps) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const bf16* kr = Ks + (size_t)min(tile * 8 + g, nc - 1) * 64 + q * 8;
      const bf16* qr = Qs + g * QS + q * 8;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint4 kb = *reinterpret_cast<const uint4*>(kr + c * 32);
        const uint4 qa = *reinterpret_cast<const uint4*>(qr + c * 32);
        mma_16816(acc, qa.x, 0u, qa.y, 0u, kb.x, kb.y);
        mma_16816(acc, qa.z, 0u, qa.w, 0u, kb.z, kb.w);
      }
      if (g < G) {   // C fragment: c0, c1 -> (head g, keys 2q, 2q + 1)
        const int j = tile * 8 + 2 * q;
        if (j < nc) sc[g * SCS + j] = acc[0];
        if (j + 1 < nc) sc[g * SCS + j + 1] = acc[1];
      }
    }
    if (has_new && tid < G) {   // the newest key comes from shared memory
      float s = 0.f;
#pragma unroll 8
      for (int e = 0; e < 64; ++e) s += knew[e] * __bfloat162float(Qs[tid * QS + e]);
      snew[tid] = s;
    }
    __syncthreads();
    fine_stamp(ft, 3);
    if (warp < G) {   // softmax statistics of this segment: warp h owns head h
      float* sh = sc + warp * SCS;
      float m = has_new ? snew[warp] : -INFINITY;
      for (int jj = lane; jj < nc; jj += 32) m = fmaxf(m, sh[jj]);
      m = warp_max(m);
      float l = 0.f;
      for (int jj = lane; jj < nc16; jj += 32) {
        const float e = jj < nc ? __expf(sh[jj] - m) : 0.f;   // zero probability for the padding keys of the last 16-key step
        sh[jj] = e;
        l += e;
      }
      l = warp_sum(l);
      if (lane == 0) {
        const float en = has_new ? __expf(snew[warp] - m) : 0.f;
        pn[warp] = en;
        ml[warp] = m;
        ml[kMaxG + warp] = l + en;
      }
    }
    __syncthreads();
    fine_stamp(ft, 4);
    // ---- weighted V ----
    {
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      const float* pr = sc + (g < G ? g : 0) * SCS + 2 * q;
      for (int k0 = warp * 16; k0 < nc16; k0 += kDecWarps * 16) {
        uint32_t a0 = 0u, a2 = 0u;
        if (g < G) {
          const float2 p0 = *reinterpret_cast<const float2*>(pr + k0), p1 = *reinterpret_cast<const float2*>(pr + k0 + 8);
          a0 = pack_bf16(p0.x, p0.y);
          a2 = pack_bf16(p1.x, p1.y);
        }
        // ldmatrix.x4.trans: lanes 0-7 / 8-15 address keys k0 .. k0+7 / k0+8 .. k0+15 at dims 16 j, lanes 16-31 the same keys at dims 16 j + 8
        const uint8_t* vrow = Vs + (size_t)(k0 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 128 + (size_t)(lane >> 4) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4_trans(r0, r1, r2, r3, vrow + j * 32);
          mma_16816(acc[2 * j], a0, 0u, a2, 0u, r0, r1);
          mma_16816(acc[2 * j + 1], a0, 0u, a2, 0u, r2, r3);
        }
      }
      if (g < G) {   // C fragment: c0, c1 -> (head g, dims 8 j + 2q, 8 j + 2q + 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(part_o + (warp * G + g) * 64 + j * 8 + 2 * q) = make_float2(acc[j][0], acc[j][1]);
      }
    }
    __syncthreads();
    fine_stamp(ft, 5);
    float* gp = p.attn_part + (size_t)item * kMaxG * kPartStride;
    for (int t = tid; t < G * 64; t += kDecThreads) {
      const int h = t >> 6, d = t & 63;
      float o = has_new ? pn[h] * vnew[d] : 0.f;
#pragma unroll
      for (int w = 0; w < kDecWarps; ++w) o += part_o[(w * G + h) * 64 + d];
      gp[h * kPartStride + d] = o;
    }
    if (tid < G) {
      gp[tid * kPartStride + 64] = ml[tid];
      gp[tid * kPartStride + 65] = ml[kMaxG + tid];
    }
    __syncthreads();
    fine_stamp(ft, 6);
    // hand-off: the add releases this CTA's partials (cumulative through the CTA barrier above) and acquires those of the segments
    // that arrived earlier
    if (tid == 0) *s_flag = (atom_add_acq_rel(p.attn_cnt + bh, 1u) == (unsigned)(S - 1)) ? 1 : 0;
    __syncthreads();
    fine_stamp(ft, 7);
    if (ft && tid == 0) ft[9] = *s_flag;
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
    fine_stamp(ft, 8);
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
  auto l2_ahead = [&](int t) {
    if (t < t1) {
      for (int i = lane; i < 8 * lines; i += 32) {
        const int r = i / lines, ln = i - r * lines;
        prefetch_l2(reinterpret_cast<const uint8_t*>(p.lm_head + (size_t)min(t * 8 + r, V - 1) * D) + (size_t)ln * 128);
      }
    }
  };
  l2_ahead(t0 + warp);
  constexpr int U = 14;   // 16-byte weight fragments in flight per lane (7 KB per warp)
  for (int t = t0 + warp; t < t1; t += kDecWarps) {
    l2_ahead(t + kDecWarps);
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
    for (int c0 = 0; c0 < (int)gridDim.x; c0 += 32 * 8) {   // up to 8 partials per lane requested at once
      float pv[8];
      int pi[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 32 + lane;
        const int cc = c < (int)gridDim.x ? c : 0;
        pv[u] = ldcg_f(p.best_val + (size_t)cc * 32 + m);
        pi[u] = ldcg_i(p.best_idx + (size_t)cc * 32 + m);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u * 32 + lane < (int)gridDim.x) best_merge(v, i, pv[u], pi[u]);
    }
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

// Shared memory: [ weight buffer | normalised rows As | phase scratch ].  Weight buffer (bulk-copy targets):
//   slot O  [0, 8 wsd)                  one n8 tile with K = hidden: the q|k|v tile, then the o tile
//   G       [8 wsd, 8 wsd (1 + 2 GS))  GS gate|up tile pairs; during the attention phase the K / V rows of the work item
//   the down tile (K = mlp, 8 wsi bytes) overlays both once they are consumed.
// Prefetch schedule of a layer (issue points; every copy is in flight across at least one grid barrier):
//   q|k|v MMAs done -> o tile + this CTA's K / V rows;  attention done -> first GS gate|up pairs;
//   gate|up round r MMAs done -> round r + 1, after the last round the down tile;  down MMAs done -> q|k|v tile of the next layer.
template <int MT>
__global__ void __launch_bounds__(kDecThreads, 1)
decode_loop_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ int s_abort, s_flag;
  __shared__ __align__(8) uint64_t s_bar[4];   // staging barriers: slot O, G region, down tile, SwiGLU rows
  __shared__ float s_rope[64];                 // cos[32], sin[32] of the current step's position
  const int D = p.D, I = p.I, M = p.M;
  const int SA = D + kAPad;
  const int wsd = D * 2, wsi = I * 2;                                // staged weight / SwiGLU row strides (bytes): dense
  uint8_t* wbuf = smem_raw;
  uint8_t* gbuf = wbuf + (size_t)8 * wsd;
  bf16* As = reinterpret_cast<bf16*>(smem_raw + p.wbuf_bytes);                                        // [MT * 16][SA] normalised rows
  float* scr = reinterpret_cast<float*>(smem_raw + p.wbuf_bytes + (size_t)MT * 16 * SA * 2);          // phase scratch
  // per-layer weight pointers: read from shared memory where needed instead of living in 14 registers across every phase
  const DecLayerDev* s_layers = reinterpret_cast<const DecLayerDev*>(smem_raw + p.wbuf_bytes + (size_t)MT * 16 * SA * 2 + p.scr_bytes);
  for (int i = threadIdx.x; i < p.n_layers * (int)(sizeof(DecLayerDev) / 8); i += kDecThreads)
    reinterpret_cast<unsigned long long*>(smem_raw + p.wbuf_bytes + (size_t)MT * 16 * SA * 2 + p.scr_bytes)[i] =
        reinterpret_cast<const unsigned long long*>(p.layers)[i];
  for (int i = threadIdx.x; i < MT * 16 * SA / 8; i += kDecThreads) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    s_abort = 0; s_flag = 0;
    for (int i = 0; i < 4; ++i) mbar_init(&s_bar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint64_t* bar_o = &s_bar[0];
  uint64_t* bar_g = &s_bar[1];
  uint64_t* bar_d = &s_bar[2];
  uint64_t* bar_a = &s_bar[3];
  uint32_t ph_o = 0, ph_g = 0, ph_d = 0, ph_a = 0;   // parity of the next completion to wait for (uniform over the CTA)
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
  constexpr int GS = gu_slots(MT);
  const int n_items = M * p.hkv * p.n_seg;
  const bool act_smem = (size_t)M * wsi <= (size_t)MT * 16 * SA * 2;   // the SwiGLU rows fit beside each other in the As region
  int steps_done = 0;
  bool o_pending = false;   // a copy into slot O is in flight and nobody has waited for it yet
  if (qa < qb) { stage_tile(bar_o, wbuf, s_layers[0].qkv, D, qa * 8, D, wsd); o_pending = true; }
  GemmCfg c;
  c.M = M; c.part = scr; c.pvs = scr + kDecWarps * MT * 16 * 17; c.bar1 = nullptr; c.ph1 = 0; c.bias = nullptr;
  for (int t = 0; t < p.n_steps; ++t) {
    const int pos = pos0 + t;
    if (pos >= p.lmax) break;
    if (p.eos >= 0) {   // every sequence finished: stop (the host loop's `done.all()` test, llm.py:245, without the round trip)
      bool all = true;
      for (int m = 0; m < M; ++m) all = all && (__ldcg(p.done + m) != 0);
      if (all) break;
    }
    if (threadIdx.x < 32) {   // cos / sin of this step's position, same arithmetic as rope_kv_write_kernel; read after the next __syncthreads
      const float inv_freq = exp2f(-(float)(2 * threadIdx.x) / 64.0f * p.log2_theta);
      sincosf((float)pos * inv_freq, &s_rope[32 + threadIdx.x], &s_rope[threadIdx.x]);
    }
    for (int l = 0; l < p.n_layers; ++l) {
      const DecLayerDev& L = s_layers[l];
      long long* ft = (p.trace && t == 1 && l == 1) ? p.trace + (size_t)gridDim.x * kTraceBarriers * 2 + 16 + (size_t)blockIdx.x * kFineStamps : nullptr;
      // ---- 1: RMSNorm + q|k|v ----
      fine_stamp(ft, 20);
      stage_norm(p, As, SA, L.ln1, l == 0, scr);
      fine_stamp(ft, 21);
      for (int tile = qa; tile < qb || tile == qa; ++tile) {   // (a CTA without a tile still issues the copies below once)
        if (tile < qb) {
          c.ng = 1; c.K = D; c.ws = wsd; c.amode = 0; c.epi = kEpiQkv; c.wb = wbuf;
          c.A = reinterpret_cast<const uint8_t*>(As); c.a_stride = (long long)SA * 2;
          c.bar0 = bar_o; c.ph0 = ph_o; c.bar1 = nullptr;
          c.out = p.qkv; c.ldo = NQKV; c.n0 = tile * 8; c.bias = L.bqkv; c.ft = ft ? ft + 22 : nullptr;
          o_pending = false;
          gemm_mma<MT, 1>(c);
          ph_o ^= 1u;
        }
        if (tile + 1 < qb) {
          stage_tile(bar_o, wbuf, L.qkv, D, (tile + 1) * 8, D, wsd);
        } else {   // q|k|v weights consumed: the o tile and this CTA's K / V rows travel across the barrier
          if (oa < ob) { stage_tile(bar_o, wbuf, L.o, D, oa * 8, D, wsd); o_pending = true; }
          if ((int)blockIdx.x < n_items) attn_stage(p, bar_g, l, pos, blockIdx.x, gbuf);
        }
        if (tile < qb) gemm_epi<MT, 1>(c);
      }
      if (!grid_barrier(p, ++epoch, &s_abort)) return;
      // ---- 2: RoPE + KV write + attention ----
      for (int pt = ga + GS; pt < gb; ++pt) {   // later gate|up rounds: at least in L2 by the time they are staged
        prefetch_rows(p, L.gu, D, gu_row(pt, 0), 8, D);
        prefetch_rows(p, L.gu, D, gu_row(pt, 1), 8, D);
      }
      fine_stamp(ft, 0);
      attn_phase(p, l, pos, scr, reinterpret_cast<float*>(As), gbuf, bar_g, ph_g, &s_flag, s_rope, ft);
      if (ga < gb) stage_gu(bar_g, gbuf, L.gu, D, wsd, ga, min(ga + GS, gb));
      if (!grid_barrier(p, ++epoch, &s_abort)) return;
      // ---- 3: o projection into the residual stream (the gate|up copies stay in flight) ----
      for (int tile = oa; tile < ob; ++tile) {
        c.ng = 1; c.K = D; c.ws = wsd; c.amode = 1; c.epi = kEpiResidual; c.wb = wbuf;
        c.A = reinterpret_cast<const uint8_t*>(p.att); c.a_stride = (long long)D * 2;
        c.bar0 = bar_o; c.ph0 = ph_o; c.bar1 = nullptr;
        c.out = p.x; c.ldo = D; c.n0 = tile * 8; c.ft = nullptr;
        o_pending = false;
        gemm_mma<MT, 1>(c);
        ph_o ^= 1u;
        if (tile + 1 < ob) stage_tile(bar_o, wbuf, L.o, D, (tile + 1) * 8, D, wsd);
        gemm_epi<MT, 1>(c);
      }
      if (!grid_barrier(p, ++epoch, &s_abort)) return;
      // ---- 4: RMSNorm + gate|up + SwiGLU, GS tile pairs per round ----
      fine_stamp(ft, 10);
      stage_norm(p, As, SA, L.ln2, false, scr);
      fine_stamp(ft, 11);
      for (int pg = ga; pg < gb || pg == ga; pg += GS) {
        const int ng = min(GS, gb - pg);
        if (ng > 0) {
          c.ng = ng; c.K = D; c.ws = wsd; c.amode = 0; c.epi = kEpiSwiglu; c.wb = gbuf;
          c.A = reinterpret_cast<const uint8_t*>(As); c.a_stride = (long long)SA * 2;
          c.bar0 = bar_g; c.ph0 = ph_g; c.bar1 = nullptr;
          c.out = p.act; c.ldo = I; c.n0 = pg * 8; c.ft = ft ? ft + (pg == ga ? 12 : 16) : nullptr;
          gemm_mma<MT, 2>(c);
          ph_g ^= 1u;
        }
        const int nxt = pg + GS;
        if (nxt < gb) stage_gu(bar_g, gbuf, L.gu, D, wsd, nxt, min(nxt + GS, gb));
        else if (da < db) stage_tile(bar_d, wbuf, L.d, I, da * 8, I, wsi);
        if (ng > 0) gemm_epi<MT, 2>(c);
      }
      if (!grid_barrier(p, ++epoch, &s_abort)) return;
      // ---- 5: down projection into the residual stream ----
      {
        const bool stage_act = act_smem && da < db;
        if (stage_act)   // few rows: the SwiGLU outputs are staged beside each other (stride wsi) where the normalised rows were
          stage_strided(bar_a, 1, (uint32_t)M * I * 2, reinterpret_cast<const uint8_t*>(p.act), 0, reinterpret_cast<uint8_t*>(As), 0);
        for (int tile = da; tile < db || tile == da; ++tile) {
          if (tile < db) {
            c.ng = 1; c.K = I; c.ws = wsi; c.epi = kEpiResidual; c.wb = wbuf;
            if (act_smem) { c.amode = 2; c.A = reinterpret_cast<const uint8_t*>(As); c.a_stride = wsi; }
            else { c.amode = 1; c.A = reinterpret_cast<const uint8_t*>(p.act); c.a_stride = (long long)I * 2; }
            c.bar0 = bar_d; c.ph0 = ph_d;
            const bool wait_act = stage_act && tile == da;
            c.bar1 = wait_act ? bar_a : nullptr; c.ph1 = ph_a;
            c.out = p.x; c.ldo = D; c.n0 = tile * 8; c.ft = nullptr;
            gemm_mma<MT, 1>(c);
            ph_d ^= 1u;
            if (wait_act) ph_a ^= 1u;
          }
          if (tile + 1 < db) stage_tile(bar_d, wbuf, L.d, I, (tile + 1) * 8, I, wsi);
          else if (qa < qb) { stage_tile(bar_o, wbuf, s_layers[l + 1 < p.n_layers ? l + 1 : 0].qkv, D, qa * 8, D, wsd); o_pending = true; }
          if (tile < db) gemm_epi<MT, 1>(c);
        }
      }
      if (!grid_barrier(p, ++epoch, &s_abort)) return;
    }
    // ---- final norm + LM head + arg-max ----
    stage_norm(p, As, SA, p.norm_w, false, scr);
    lm_head_phase<MT>(p, As, SA, scr);
    if (!grid_barrier(p, ++epoch, &s_abort)) return;
    sample_phase(p, step0 + t);
    steps_done = t + 1;
    if (!grid_barrier(p, ++epoch, &s_abort)) return;
  }
  if (o_pending) mbar_wait(bar_o, ph_o);   // no bulk copy may still be writing this CTA's shared memory when it exits
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

size_t decode_wbuf_bytes(int mt, int hidden, int mlp) {
  const size_t wsd = (size_t)hidden * 2, wsi = (size_t)mlp * 2;
  const size_t a = 8 * wsd * (1 + 2 * gu_slots(mt)), b = 8 * wsi;
  return align_up(a > b ? a : b, 128);
}
size_t decode_scr_bytes(int mt) {
  const size_t gemm = (size_t)kDecWarps * mt * 16 * 17 * 4 + (size_t)mt * 16 * 8 * 4;   // partial tiles + parked epilogue operands
  const size_t attn = (size_t)(16 * 72 / 2 + 4 * kMaxG + 64 + 64 + kMaxG * (kSegMax + 2)) * 4;
  const size_t lm = (size_t)kDecWarps * 32 * 8;
  size_t scr = gemm > attn ? gemm : attn;
  if (lm > scr) scr = lm;
  return align_up(scr, 16);
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
  // (the staged K / V rows of a segment must fit into the gate|up region of the weight buffer)
  int seg = grid / (a->batch * a->hkv);
  const int by_len = a->lmax / 128 > 0 ? a->lmax / 128 : 1;
  if (seg > by_len) seg = by_len;
  if (seg < ceil_div(a->lmax, kSegMax)) seg = ceil_div(a->lmax, kSegMax);
  if (seg < 1) seg = 1;
  SLB_CHECK_ARG((size_t)a->batch * a->hkv * seg <= (size_t)(a->batch * a->hkv > grid ? a->batch * a->hkv : grid) * 8,
                "decode_loop: lmax=%d needs %d key segments per sequence: too many work items", a->lmax, seg);
  p.n_seg = seg;
  const int mt = a->batch <= 16 ? 1 : 2;
  p.wbuf_bytes = (int)decode_wbuf_bytes(mt, a->hidden, a->mlp);
  SLB_CHECK_ARG((size_t)kDecWarps * a->hq / a->hkv * 64 * 4 <= (size_t)mt * 16 * (a->hidden + kAPad) * 2,
                "decode_loop: the attention partials of %d heads per kv head do not fit beside the normalised rows", a->hq / a->hkv);
  SLB_CHECK_ARG((size_t)p.wbuf_bytes - 8 * (size_t)a->hidden * 2 >= (size_t)kSegMax * 256,
                "decode_loop: hidden=%d too small for the K / V staging area", a->hidden);
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
  p.scr_bytes = (int)decode_scr_bytes(mt);
  const size_t smem = (size_t)p.wbuf_bytes + (size_t)mt * 16 * (a->hidden + kAPad) * 2 + p.scr_bytes + (size_t)a->n_layers * sizeof(DecLayerDev);
  SLB_CHECK_ARG(smem <= 227 * 1024, "decode_loop: %zu bytes of shared memory for hidden=%d mlp=%d exceed the 227 KB of an SM", smem, a->hidden, a->mlp);
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
