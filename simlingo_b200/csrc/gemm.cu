// bf16 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma (fp32 accumulators in TMEM, two accumulator stages) ->
// tcgen05.ld epilogue with fused bias / activation / layer-scale / residual / SwiGLU.  Persistent, warp-specialised:
//   warp 0     TMA producer (one elected lane)
//   warp 1     TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9  epilogue (warp w reads TMEM lanes 32*(w%4) .. +31 = accumulator rows; two warps per quadrant split the columns)
// Two kernels: gemm_bf16_kernel<BN,TA,TB> (cta_group::1, 128 x BN tiles, BN = 32 / 64 / 128 / 256) and gemm2_bf16_kernel<BN,EPI,TA,TB>
// (cta_group::2: a cluster of two CTAs per 256 x BN tile, BN = 256 / 224 / 192; EPI = 1: fp32 residual stream moved by TMA).
// Operand forms: K-major (forward), b_t (dgrad), a_t + b_t (wgrad); a second A source appended along k (un-merged LoRA).
// slb_gemm_bf16 picks the kernel and tile shape by estimated waves x tile area / kernel efficiency; M <= 32 goes to the
// weight-streaming kernels of gemv.cu.  Covers K1,K3,K5,K7,K10,K13,K14 of SURVEY.md section 2.2 and their dgrad / wgrad.
#include <cstdlib>
#include "common.cuh"
#include "../../include/simlingo_b200.h"

long long* slb_debug_trace_ptr();

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)

static int slb_gemm_res_prefetch() {
#ifdef SLB_ABLATION  // A/B timing switch, not compiled into the product library
  static int v = -1;
  if (v < 0) { const char* e = getenv("SLB_GEMM_NO_RESPF"); v = (e && atoi(e)) ? 0 : 1; }
  return v;
#else
  return 1;
#endif
}

struct EpiParams {
  long long* dbg;  // optional [16] int64: wait-cycle counters of cluster 0 (slb_debug_set_trace), else null
  int M, N, K;
  int K2;          // columns of the second A source (A2), logically appended to A along k; B spans K + K2 columns
  void* out;
  long long ldo;
  const bf16* bias;
  const bf16* scale_n;
  const void* res;
  long long ldr;
  float alpha;
  int act;
  int swiglu;
  int out_fp32;
  int num_m, num_n;
  bf16* aux;        // optional second operand of the epilogue, bf16 [M, N] with row stride ld_aux
  long long ld_aux;
  int res_prefetch; // 0 disables the residual prefetch (SLB_GEMM_NO_RESPF=1, A/B timing only)
  int aux_mode;     // 1: store the pre-activation (alpha*acc + bias) there; 2: multiply by gelu'(aux) (fc2 dgrad -> d pre-GELU)
};

template <int BN>
struct SmemLayout {
  // narrow tiles (tall-skinny products such as the LoRA down-projections, N <= 64) are bound by the A stream: deeper rings
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : (BN == 64 ? 8 : 10));
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SLB_ACT_GELU) return gelu_erf_fast(v);
  if (act == SLB_ACT_SILU) return silu_fast(v);
  if (act == SLB_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Stores 32 consecutive output columns of one row.
__device__ __forceinline__ void store_row32(const EpiParams& p, int row, int col0, const float (&v)[32], int ncols_total) {
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
    bool vec = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (col0 + 32 <= ncols_total);
    const bool vec32 = ((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) && (col0 + 32 <= ncols_total);
    if (vec32) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_global_v8(o + 8 * i, __float_as_uint(v[8 * i]), __float_as_uint(v[8 * i + 1]), __float_as_uint(v[8 * i + 2]),
                     __float_as_uint(v[8 * i + 3]), __float_as_uint(v[8 * i + 4]), __float_as_uint(v[8 * i + 5]),
                     __float_as_uint(v[8 * i + 6]), __float_as_uint(v[8 * i + 7]));
    } else if (vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) o[i] = v[i];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col0;
    bool vec = ((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (col0 + 32 <= ncols_total);
    const bool vec32 = ((p.ldo & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) && (col0 + 32 <= ncols_total);
    if (vec32) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        st_global_v8(o + 16 * i, pack_bf16(v[16 * i + 0], v[16 * i + 1]), pack_bf16(v[16 * i + 2], v[16 * i + 3]),
                     pack_bf16(v[16 * i + 4], v[16 * i + 5]), pack_bf16(v[16 * i + 6], v[16 * i + 7]),
                     pack_bf16(v[16 * i + 8], v[16 * i + 9]), pack_bf16(v[16 * i + 10], v[16 * i + 11]),
                     pack_bf16(v[16 * i + 12], v[16 * i + 13]), pack_bf16(v[16 * i + 14], v[16 * i + 15]));
    } else if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
        u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
        u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(o)[i] = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) o[i] = __float2bfloat16(v[i]);
    }
  }
}

__device__ __forceinline__ void add_residual32(const EpiParams& p, int row, int col0, float (&v)[32], int ncols_total) {
  if (p.out_fp32) {
    const float* r = reinterpret_cast<const float*>(p.res) + (long long)row * p.ldr + col0;
    bool vec = ((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) && (col0 + 32 <= ncols_total);
    if (vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 t = reinterpret_cast<const float4*>(r)[i];
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) v[i] += r[i];
    }
  } else {
    const bf16* r = reinterpret_cast<const bf16*>(p.res) + (long long)row * p.ldr + col0;
    bool vec = ((p.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) && (col0 + 32 <= ncols_total);
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u = reinterpret_cast<const uint4*>(r)[i];
        float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        v[8 * i + 0] += a.x; v[8 * i + 1] += a.y; v[8 * i + 2] += b.x; v[8 * i + 3] += b.y;
        v[8 * i + 4] += c.x; v[8 * i + 5] += c.y; v[8 * i + 6] += d.x; v[8 * i + 7] += d.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) v[i] += __bfloat162float(r[i]);
    }
  }
}

__device__ __forceinline__ void load32_bf16(const bf16* p, int col0, int n, float (&f)[32]) {
  if (col0 + 32 <= n && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + col0) + i);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      f[8 * i + 0] = a.x; f[8 * i + 1] = a.y; f[8 * i + 2] = b.x; f[8 * i + 3] = b.y;
      f[8 * i + 4] = c.x; f[8 * i + 5] = c.y; f[8 * i + 6] = d.x; f[8 * i + 7] = d.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (col0 + i < n) ? __bfloat162float(__ldg(p + col0 + i)) : 0.f;
  }
}

__device__ __forceinline__ void aux_store32(const EpiParams& p, int row, int col0, const float (&v)[32], int n) {
  bf16* o = p.aux + (long long)row * p.ld_aux + col0;
  if (((p.ld_aux & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && (col0 + 32 <= n)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
      u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
      reinterpret_cast<uint4*>(o)[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < n) o[i] = __float2bfloat16(v[i]);
  }
}
__device__ __forceinline__ void aux_gelu_grad32(const EpiParams& p, int row, int col0, float (&v)[32], int n) {
  float a[32];
  load32_bf16(p.aux + (long long)row * p.ld_aux, col0, n, a);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float cdf = 0.5f * (1.0f + erff(a[i] * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * a[i] * a[i]);
    v[i] *= cdf + a[i] * pdf;
  }
}

// Residual prefetch (bf16, 16-byte aligned rows, full 32-column chunk): the residual row segment of the NEXT chunk is
// requested before the current chunk is processed - and the first one before the accumulator barrier is awaited - so
// its DRAM latency (one ~1 us stall per chunk otherwise: the layer-scale residual GEMMs sat at 51 % tensor-pipe
// utilisation) overlaps with TMEM loads, math and stores.
struct ResPrefetch {
  uint4 v[8];   // 32 bf16 (4 vectors) or 32 fp32 (8 vectors: the fp32 residual streams of the inference path)
  bool valid;
};
__device__ __forceinline__ bool res_vec_ok(const EpiParams& p) {
  return p.res_prefetch && p.res && !p.swiglu && ((p.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0);
}
__device__ __forceinline__ void res_prefetch(const EpiParams& p, int row, int col0, ResPrefetch& r) {
  r.valid = row < p.M && col0 + 32 <= p.N;
  if (r.valid) {
    if (p.out_fp32) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res) + (long long)row * p.ldr + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) r.v[i] = src[i];
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.res) + (long long)row * p.ldr + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) r.v[i] = src[i];
    }
  }
}
template <int BN>
__device__ __forceinline__ int epi_first_chunk(int half) {
  constexpr int kChunks = BN / 32, kFirst = (kChunks + 1) / 2;
  return half ? kFirst * 32 : 0;
}

// One 32-column chunk through every epilogue option, no assumptions on alignment (see epilogue_tile for the fast path).
__device__ __forceinline__ void epilogue_chunk_generic(const EpiParams& p, uint32_t taddr, int row, bool row_ok, int col0) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
  if (p.bias) {
    float b[32];
    load32_bf16(p.bias, col0, p.N, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += b[i];
  }
  if (p.aux_mode == 1 && row_ok) aux_store32(p, row, col0, v, p.N);
  if (p.aux_mode == 2 && row_ok) aux_gelu_grad32(p, row, col0, v, p.N);
  if (p.act == SLB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(v[i]);
  } else if (p.act == SLB_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
  } else if (p.act == SLB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.scale_n) {
    float sc[32];
    load32_bf16(p.scale_n, col0, p.N, sc);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= sc[i];
  }
  if (row_ok) {
    if (p.res) add_residual32(p, row, col0, v, p.N);
    store_row32(p, row, col0, v, p.N);
  }
}

// Epilogue of one accumulator tile for one thread (= one output row), 32-column chunks [c_begin, c_end):
// TMEM -> registers -> fused epilogue -> global.  Two warps share a TMEM lane quadrant and split the chunks.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, uint32_t taddr, int row, int n0, int half, ResPrefetch& pre, bool use_pre) {
  const bool row_ok = row < p.M;
  if (!p.swiglu) {
    constexpr int kChunks = BN / 32, kFirst = (kChunks + 1) / 2;
    const int c_begin = half ? kFirst * 32 : 0, c_end = half ? BN : kFirst * 32;
    // fast path: whole 32-column chunks, 16-byte aligned bf16 bias / scale vectors, bf16 or fp32 output with aligned rows, no
    // aux operand.  The per-column vectors and the next chunk's residual are requested BEFORE the TMEM load is awaited, the
    // arithmetic runs on packed fp32x2 pairs (half the issue slots), erf-GELU on the FMA pipe.
    const bool vec_io = p.out_fp32 ? (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0))
                                   : (((p.ldo & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0));
    const bool fast = vec_io && p.aux_mode == 0 && (p.act == SLB_ACT_NONE || p.act == SLB_ACT_GELU) &&
                      (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) && (!p.scale_n || (reinterpret_cast<uintptr_t>(p.scale_n) & 15) == 0) &&
                      (!p.res || use_pre);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      if (n0 + c >= p.N) break;  // warp-uniform
      if (fast && n0 + c + 32 <= p.N) {
        const int col0 = n0 + c;
        ResPrefetch cur = pre;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        uint4 bq[4], sq[4];
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 4; ++i) bq[i] = __ldg(reinterpret_cast<const uint4*>(p.bias + col0) + i);
        }
        if (p.scale_n) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sq[i] = __ldg(reinterpret_cast<const uint4*>(p.scale_n + col0) + i);
        }
        if (use_pre && c + 32 < c_end) res_prefetch(p, row, col0 + 32, pre);
        tmem_ld_wait();
        const uint64_t alpha2 = f2splat(p.alpha);
        uint64_t v2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint64_t a = f2pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          if (p.bias) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(bq)[i];
            a = f2fma(a, alpha2, bf2_to_f2(w));
          } else if (p.alpha != 1.0f) {
            a = f2mul(a, alpha2);
          }
          v2[i] = a;
        }
        if (p.act == SLB_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = gelu_erf_poly2(v2[i]);
        }
        if (p.scale_n) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = f2mul(v2[i], bf2_to_f2(reinterpret_cast<const uint32_t*>(sq)[i]));
        }
        if (row_ok) {
          if (p.res && cur.valid) {
            if (p.out_fp32) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(cur.v);
                v2[i] = f2add(v2[i], f2pack(__uint_as_float(rw[2 * i]), __uint_as_float(rw[2 * i + 1])));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v2[i] = f2add(v2[i], bf2_to_f2(reinterpret_cast<const uint32_t*>(cur.v)[i]));
            }
          }
          if (p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t* w = reinterpret_cast<const uint32_t*>(&v2[4 * i]);
              st_global_v8(o + 8 * i, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            }
          } else {
            bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col0;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a, b;
              f2unpack(v2[i], a, b);
              pk[i] = pack_bf16(a, b);
            }
            st_global_v8(o, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            st_global_v8(o + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
          }
        }
        continue;
      }
      // generic path (edge chunks, unaligned operands, aux operands of the training epilogues, SiLU / ReLU).  Kept inline: as an
      // out-of-line subroutine it returned wrong SiLU columns on the B200 (r02_ops_tests5.log)
      if (use_pre && c + 32 < c_end) res_prefetch(p, row, n0 + c + 32, pre);
      epilogue_chunk_generic(p, taddr + c, row, row_ok, n0 + c);
    }
  } else {
    // gate columns [0,BN/2), up columns [BN/2,BN) of this tile; output column base n0/2
    const int n_out_total = p.N / 2;
#pragma unroll 1
    for (int c = half * (BN / 4); c < (half + 1) * (BN / 4); c += 32) {
      uint32_t rg[32], ru[32];
      tmem_ld_32x32(taddr + c, rg);
      tmem_ld_32x32(taddr + BN / 2 + c, ru);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = silu_fast(__uint_as_float(rg[i]) * p.alpha) * (__uint_as_float(ru[i]) * p.alpha);
      if (row_ok) store_row32(p, row, n0 / 2 + c, v, n_out_total);
    }
  }
}

template <int BN, int TA, int TB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a2, EpiParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n;
  const int num_kb1 = (p.K + BK - 1) / BK;
  const int num_kb = num_kb1 + (p.K2 + BK - 1) / BK;   // k-blocks of A, then of A2 (K is a multiple of BK when K2 > 0)
  constexpr uint32_t kTmemCols = 2 * BN;  // 256 or 512: power of two

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_trigger();
  pdl_wait();   // barriers / TMEM are set up; the operands may come from the previous kernel in the stream

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n) * BM;
        const int n0 = (tile % p.num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const int k0 = kb * BK;
          if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (TA == 0) {
            if (kb < num_kb1) tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
            else tma_load_2d(sa, &tmap_a2, &full_bar[stage], (kb - num_kb1) * BK, m0);
          } else {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) tma_load_2d(sa + g * (64 * BK * 2), &tmap_a, &full_bar[stage], m0 + g * 64, k0);
          }
          if (TB == 0) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) tma_load_2d(sb + g * (64 * BK * 2), &tmap_b, &full_bar[stage], n0 + g * 64, k0);
          }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, TA, TB);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t da = TA ? umma_desc_mnmajor_sw128(sa, 64 * BK * 2) : umma_desc_kmajor_sw128(sa);
          const uint64_t db = TB ? umma_desc_mnmajor_sw128(sb, 64 * BK * 2) : umma_desc_kmajor_sw128(sb);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // K-major: +32 bytes per UMMA_K inside the swizzle row; MN-major: +16 k-rows * 128 B
              const uint64_t ka = TA ? (uint64_t)(k * (UMMA_K * 128 >> 4)) : (uint64_t)(k * (UMMA_K * 2 >> 4));
              const uint64_t kbo = TB ? (uint64_t)(k * (UMMA_K * 128 >> 4)) : (uint64_t)(k * (UMMA_K * 2 >> 4));
              tc_mma_bf16(tmem_d, da + ka, db + kbo, idesc, (kb | k) != 0);
            }
            tc_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
            if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n) * BM;
      const int n0 = (tile % p.num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool use_pre = res_vec_ok(p);
      ResPrefetch pre;
      pre.valid = false;
      if (use_pre) res_prefetch(p, row, n0 + epi_first_chunk<BN>((warp - 2) >> 2), pre);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      epilogue_tile<BN>(p, taddr, row, n0, (warp - 2) >> 2, pre, use_pre);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int TA, int TB>
int launch_gemm(const slb_gemm_args* a, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  CUtensorMap ta, tb, ta2;
  int rc;
  const int K2 = (a->A2 && a->K2 > 0) ? a->K2 : 0;
  if (!TA) rc = slb_make_tmap_2d(&ta, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, BK, BM);
  else     rc = slb_make_tmap_2d(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64, BK);
  if (rc) return rc;
  if (!TB) rc = slb_make_tmap_2d(&tb, a->B, (uint64_t)(a->K + K2), (uint64_t)a->N, (uint64_t)a->ldb * 2, BK, BN);
  else     rc = slb_make_tmap_2d(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64, BK);
  if (rc) return rc;
  ta2 = ta;
  if (K2) {
    rc = slb_make_tmap_2d(&ta2, a->A2, (uint64_t)K2, (uint64_t)a->M, (uint64_t)a->lda2 * 2, BK, BM);
    if (rc) return rc;
  }
  EpiParams p;
  p.dbg = nullptr;
  p.M = a->M; p.N = a->N; p.K = a->K; p.K2 = K2;
  p.out = a->out; p.ldo = a->ldo;
  p.bias = (const bf16*)a->bias; p.scale_n = (const bf16*)a->scale_n;
  p.res = a->residual; p.ldr = a->ldr;
  p.alpha = a->alpha; p.act = a->act; p.swiglu = a->swiglu; p.out_fp32 = a->out_fp32;
  p.aux = (bf16*)a->aux; p.ld_aux = a->ld_aux; p.aux_mode = a->aux ? a->aux_mode : 0;
  p.res_prefetch = slb_gemm_res_prefetch();
  p.num_m = ceil_div(a->M, BM);
  p.num_n = ceil_div(a->N, BN);
  auto kern = gemm_bf16_kernel<BN, TA, TB>;
  static bool attr_set = false;
  if (!attr_set) {
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int grid = p.num_m * p.num_n;
  int sms = slb_num_sms();
  if (grid > sms) grid = sms;
  SLB_CUDA(slb_launch_pdl((double)a->M * a->N * a->K < 1e10, kern, dim3(grid), dim3(GEMM_THREADS), (size_t)L::kTotal, stream, ta, tb, ta2, p));
  return SLB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  CTA r holds A rows [128r, 128r+128)
// and B rows (= output columns) [BN/2 * r, BN/2 * (r+1)) of every k-block in its own shared memory (32 KB per stage
// instead of 48 KB -> 6 stages instead of 4, and each operand byte is fetched from L2 once per 256-row / BN-col
// tile); the leader CTA issues M=256 MMAs whose accumulator rows 128r.. live in CTA r's TMEM.  Both CTAs run their
// own epilogue over their 128 rows.  K-major operands only.
// ------------------------------------------------------------------------------------------------------------
// EPI = 1: the fp32 residual-stream epilogue moves its tiles with TMA (see epilogue_tile_tma): 5 operand stages instead of 6, the
// freed shared memory holds two 32 x 32 fp32 staging boxes per epilogue warp.
template <int BN, int EPI = 0>
struct Smem2 {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = EPI ? 5 : 6;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kBoxBytes = 32 * 32 * 4;                                         // one 32-row x 32-column fp32 box (128B swizzle)
  static constexpr int kStagingOffset = (kBarOffset + kBarBytes + 1023) / 1024 * 1024;  // 1024-byte aligned (swizzle atom)
  static constexpr int kStagingBytes = EPI ? 8 * 2 * kBoxBytes : 0;
  static constexpr int kTotal = (EPI ? kStagingOffset + kStagingBytes : kBarOffset + kBarBytes) + 1024;
};

// fp32 residual-stream epilogue through TMA (one warp, its 32 accumulator rows, 32-column chunks [c_begin, c_end)):
//   residual box (32 x 32 fp32)  --TMA load-->  smem  --lane = row, 8 x LDS.128 (128B swizzle: conflict-free)-->  registers
//   out = res + scale_n * (alpha * acc + bias)   --8 x STS.128 in place-->  smem  --TMA store-->  global
// With one row per lane, the direct global accesses of the generic epilogue touch 32 different 128-byte lines per instruction (32 L1
// wavefronts each): on the layer-scale residual GEMMs of InternViT (K = 1024: 8 K cycles of MMA per tile) the epilogue took 16 K
// cycles per tile and the kernel ran at 1.9 x the plain GEMM.  The box of chunk i + 1 is requested before chunk i is processed
// (double-buffered per warp), the first one before the accumulator barrier is awaited.  Rows beyond M: zero-filled loads, clipped stores.
struct TmaEpi {
  uint8_t* stg;      // this warp's two staging boxes
  uint64_t* rbar;    // this warp's two "residual box landed" barriers
  uint32_t ci;       // chunks processed so far (buffer = ci & 1, barrier phase = (ci >> 1) & 1)
};
template <int BN>
__device__ __forceinline__ void tma_epi_request(const EpiParams& p, const CUtensorMap* tm_res, TmaEpi& te, uint32_t k, int row0, int col0, int lane) {
  if (lane == 0) {
    tma_store_wait_read<0>();   // the store that last used this buffer has finished reading it
    mbar_expect_tx(&te.rbar[k & 1], Smem2<BN, 1>::kBoxBytes);
    tma_load_2d(te.stg + (k & 1) * Smem2<BN, 1>::kBoxBytes, tm_res, &te.rbar[k & 1], col0, row0);
  }
  __syncwarp();
}
template <int BN>
__device__ __forceinline__ void epilogue_tile_tma(const EpiParams& p, const CUtensorMap* tm_res, const CUtensorMap* tm_out, uint32_t taddr, int row0,
                                                  int n0, int half, int lane, TmaEpi& te) {
  constexpr int kChunks = BN / 32, kFirst = (kChunks + 1) / 2;
  constexpr int kBox = Smem2<BN, 1>::kBoxBytes;
  const int c_begin = half ? kFirst * 32 : 0, c_end = half ? BN : kFirst * 32;
  const uint64_t alpha2 = f2splat(p.alpha);
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    if (n0 + c >= p.N) break;  // warp-uniform
    const int col0 = n0 + c;
    if (c + 32 < c_end && col0 + 32 < p.N) tma_epi_request<BN>(p, tm_res, te, te.ci + 1, row0, col0 + 32, lane);
    uint32_t r[32];
    tmem_ld_32x32(taddr + c, r);
    uint4 bq[4], sq[4];
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 4; ++i) bq[i] = __ldg(reinterpret_cast<const uint4*>(p.bias + col0) + i);
    }
    if (p.scale_n) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sq[i] = __ldg(reinterpret_cast<const uint4*>(p.scale_n + col0) + i);
    }
    uint8_t* box = te.stg + (te.ci & 1) * kBox;
    mbar_wait(&te.rbar[te.ci & 1], (te.ci >> 1) & 1);
    uint8_t* rowp = box + lane * 128;
    uint4 rv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rv[k] = *reinterpret_cast<const uint4*>(rowp + ((k ^ (lane & 7)) << 4));
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint64_t v01 = f2pack(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]));
      uint64_t v23 = f2pack(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
      if (p.bias) {
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(bq);
        v01 = f2fma(v01, alpha2, bf2_to_f2(bw[2 * k]));
        v23 = f2fma(v23, alpha2, bf2_to_f2(bw[2 * k + 1]));
      } else if (p.alpha != 1.0f) {
        v01 = f2mul(v01, alpha2);
        v23 = f2mul(v23, alpha2);
      }
      const uint64_t r01 = f2pack(__uint_as_float(rv[k].x), __uint_as_float(rv[k].y)), r23 = f2pack(__uint_as_float(rv[k].z), __uint_as_float(rv[k].w));
      if (p.scale_n) {   // multiply, then add (not fused): bit-identical to the direct-store epilogue the small-M shapes take
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(sq);
        v01 = f2mul(v01, bf2_to_f2(sw[2 * k]));
        v23 = f2mul(v23, bf2_to_f2(sw[2 * k + 1]));
      }
      v01 = f2add(v01, r01);
      v23 = f2add(v23, r23);
      float a, b, cc, d;
      f2unpack(v01, a, b);
      f2unpack(v23, cc, d);
      *reinterpret_cast<uint4*>(rowp + ((k ^ (lane & 7)) << 4)) = make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(cc), __float_as_uint(d));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm_out, box, col0, row0);
      tma_store_commit();
    }
    ++te.ci;
  }
}

template <int BN, int EPI, int TA = 0, int TB = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_res,
                  const __grid_constant__ CUtensorMap tmap_out, EpiParams p) {
  using L = Smem2<BN, EPI>;
  constexpr int kStages = L::kStages;
  constexpr int ACC_STRIDE = 256;  // TMEM columns per accumulator stage (BN <= 256)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rbar_base = tempty_bar + 2;                                    // EPI: [8 epilogue warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar_base + 16);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = p.num_m * p.num_n;  // cluster tiles (256 x BN)
  const int num_kb1 = (p.K + BK - 1) / BK;
  const int num_kb = num_kb1 + (p.K2 + BK - 1) / BK;   // k-blocks of A, then of A2
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[i], 1);  // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // leader's: 8 epilogue warps x 2 CTAs
    }
    if (EPI) {
      tma_prefetch_desc(&tmap_res);
      tma_prefetch_desc(&tmap_out);
      for (int i = 0; i < 16; ++i) mbar_init(&rbar_base[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_trigger();
  pdl_wait();   // barriers / TMEM are set up; the operands may come from the previous kernel in the stream

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0, t_begin = clock64();
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / p.num_n) * (2 * BM) + rank * BM;
        const int n0 = (tile % p.num_n) * BN + rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          const long long c0 = clock64();
          mbar_wait(&empty_bar[stage], phase ^ 1);
          w_empty += clock64() - c0;
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            if (TA == 0) {
              if (kb < num_kb1) tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
              else tma_load_2d_2sm(sa, &tmap_a2, &full_bar[stage], (kb - num_kb1) * BK, m0);
            } else {   // MN-major A ([K, M] in memory): this CTA's 128 rows as two 64-row groups of [64 k][64 m]
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) tma_load_2d_2sm(sa + g * (64 * BK * 2), &tmap_a, &full_bar[stage], m0 + g * 64, kb * BK);
            }
            if (TB == 0) {
              tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
            } else {   // MN-major B ([K, N] in memory): this CTA's BN / 2 columns as 64-column groups
#pragma unroll
              for (int g = 0; g < BN / 2 / 64; ++g) tma_load_2d_2sm(sb + g * (64 * BK * 2), &tmap_b, &full_bar[stage], n0 + g * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.dbg && cluster_id == 0 && leader && lane == 0) { p.dbg[5] = clock64() - t_begin; p.dbg[6] = w_empty; }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, TA, TB);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_full = 0, w_tempty = 0, t_begin = clock64();
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        long long c0 = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        w_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          c0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          w_full += clock64() - c0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t da = TA ? umma_desc_mnmajor_sw128(sa, 64 * BK * 2) : umma_desc_kmajor_sw128(sa);
          const uint64_t db = TB ? umma_desc_mnmajor_sw128(sa + L::kABytes, 64 * BK * 2) : umma_desc_kmajor_sw128(sa + L::kABytes);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // K-major: +32 bytes per UMMA_K inside the swizzle row; MN-major: +16 k-rows * 128 B
              const uint64_t ka = TA ? (uint64_t)(k * (UMMA_K * 128 >> 4)) : (uint64_t)(2 * k);
              const uint64_t kbo = TB ? (uint64_t)(k * (UMMA_K * 128 >> 4)) : (uint64_t)(2 * k);
              tc_mma_bf16_2sm(tmem_d, da + ka, db + kbo, idesc, (kb | k) != 0);
            }
            tc_commit_2sm(&empty_bar[stage], 3);
            if (kb == num_kb - 1) tc_commit_2sm(&tfull_bar[acc], 3);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.dbg && cluster_id == 0 && lane == 0) {
        p.dbg[0] = clock64() - t_begin; p.dbg[1] = w_full; p.dbg[2] = w_tempty;
      }
    }
  } else {
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_tfull = 0, t_begin = clock64();
    TmaEpi te;
    te.stg = smem + L::kStagingOffset + (warp - 2) * 2 * L::kBoxBytes;
    te.rbar = rbar_base + (warp - 2) * 2;
    te.ci = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / p.num_n) * (2 * BM) + rank * BM;
      const int n0 = (tile % p.num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const int half = (warp - 2) >> 2;
      const bool use_pre = !EPI && res_vec_ok(p);
      ResPrefetch pre;
      pre.valid = false;
      if (use_pre) res_prefetch(p, row, n0 + epi_first_chunk<BN>(half), pre);
      if (EPI && n0 + epi_first_chunk<BN>(half) < p.N)   // this tile's first residual box travels while the accumulator is still being computed
        tma_epi_request<BN>(p, &tmap_res, te, te.ci, m0 + quad * 32, n0 + epi_first_chunk<BN>(half), lane);
      const long long c0 = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase);
      w_tfull += clock64() - c0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
      if (EPI) epilogue_tile_tma<BN>(p, &tmap_res, &tmap_out, taddr, m0 + quad * 32, n0, half, lane, te);
      else epilogue_tile<BN>(p, taddr, row, n0, half, pre, use_pre);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI && lane == 0) tma_store_wait<0>();   // every box has reached global memory before the CTA retires its shared memory
    if (p.dbg && cluster_id == 0 && leader && warp == 2 && lane == 0) { p.dbg[3] = clock64() - t_begin; p.dbg[4] = w_tfull; }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int BN, int EPI, int TA = 0, int TB = 0>
int launch_gemm2(const slb_gemm_args* a, cudaStream_t stream) {
  using L = Smem2<BN, EPI>;
  static_assert(!TB || (BN / 2) % 64 == 0, "MN-major B: the CTA's half tile must be whole 64-column groups");
  CUtensorMap ta, tb, ta2, tres, tout;
  const int K2 = (a->A2 && a->K2 > 0) ? a->K2 : 0;
  int rc;
  if (!TA) rc = slb_make_tmap_2d(&ta, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, BK, BM);
  else     rc = slb_make_tmap_2d(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64, BK);
  if (rc) return rc;
  if (!TB) rc = slb_make_tmap_2d(&tb, a->B, (uint64_t)(a->K + K2), (uint64_t)a->N, (uint64_t)a->ldb * 2, BK, BN / 2);
  else     rc = slb_make_tmap_2d(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64, BK);
  if (rc) return rc;
  ta2 = ta;
  if (K2) {
    rc = slb_make_tmap_2d(&ta2, a->A2, (uint64_t)K2, (uint64_t)a->M, (uint64_t)a->lda2 * 2, BK, BM);
    if (rc) return rc;
  }
  tres = ta; tout = ta;
  if (EPI) {
    rc = slb_make_tmap_2d_f32(&tres, a->residual, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldr * 4, 32, 32);
    if (rc) return rc;
    rc = slb_make_tmap_2d_f32(&tout, a->out, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldo * 4, 32, 32);
    if (rc) return rc;
  }
  EpiParams p;
  p.dbg = slb_debug_trace_ptr();
  p.M = a->M; p.N = a->N; p.K = a->K; p.K2 = K2;
  p.out = a->out; p.ldo = a->ldo;
  p.bias = (const bf16*)a->bias; p.scale_n = (const bf16*)a->scale_n;
  p.res = a->residual; p.ldr = a->ldr;
  p.alpha = a->alpha; p.act = a->act; p.swiglu = a->swiglu; p.out_fp32 = a->out_fp32;
  p.aux = (bf16*)a->aux; p.ld_aux = a->ld_aux; p.aux_mode = a->aux ? a->aux_mode : 0;
  p.res_prefetch = slb_gemm_res_prefetch();
  p.num_m = ceil_div(a->M, 2 * BM);
  p.num_n = ceil_div(a->N, BN);
  auto kern = gemm2_bf16_kernel<BN, EPI, TA, TB>;
  static bool attr_set = false;
  if (!attr_set) {
    SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  int clusters = p.num_m * p.num_n;
  const int max_clusters = slb_num_sms() / 2;
  if (clusters > max_clusters) clusters = max_clusters;
  SLB_CUDA(slb_launch_pdl((double)a->M * a->N * a->K < 1e10, kern, dim3(clusters * 2), dim3(GEMM_THREADS), (size_t)L::kTotal, stream, ta, tb, ta2, tres,
                          tout, p));
  return SLB_OK;
}

// the TMA residual-stream epilogue applies to: fp32 output + fp32 residual, no activation / aux / SwiGLU, whole 32-column chunks,
// 16-byte aligned rows and per-column vectors
static bool tma_epilogue_ok(const slb_gemm_args* a) {
#ifdef SLB_ABLATION
  static int off = -1;
  if (off < 0) { const char* e = getenv("SLB_GEMM_NO_TMA_EPI"); off = (e && atoi(e)) ? 1 : 0; }
  if (off) return false;
#endif
  return a->out_fp32 && a->residual && !a->swiglu && !a->aux && a->act == SLB_ACT_NONE && (a->N % 32) == 0 && (a->ldo % 4) == 0 &&
         (a->ldr % 4) == 0 && (((uintptr_t)a->out | (uintptr_t)a->residual) & 15) == 0 && (((uintptr_t)a->bias | (uintptr_t)a->scale_n) & 15) == 0;
}

}  // namespace

int slb_gemv_try(const slb_gemm_args* a, cudaStream_t stream, int* rc_out);    // gemv.cu
int slb_skinny_try(const slb_gemm_args* a, cudaStream_t stream, int* rc_out);  // gemv.cu

extern "C" int slb_gemm_bf16(const slb_gemm_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SLB_CHECK_ARG(a != nullptr, "gemm: null args");
  SLB_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  SLB_CHECK_ARG(a->A && a->B && a->out, "gemm: null operand");
  SLB_CHECK_ARG((a->lda % 8) == 0 && (a->ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (16 B TMA stride), got %lld %lld",
                (long long)a->lda, (long long)a->ldb);
  SLB_CHECK_ARG(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->B & 15) == 0, "gemm: A/B must be 16-byte aligned");
  SLB_CHECK_ARG(!(a->a_t && !a->b_t), "gemm: a_t=1 requires b_t=1 (wgrad form)");
  if (a->swiglu) {
    SLB_CHECK_ARG((a->N % 256) == 0, "gemm: swiglu needs N %% 256 == 0 (got %d)", a->N);
    SLB_CHECK_ARG(!a->bias && !a->scale_n && !a->residual && !a->act, "gemm: swiglu excludes other epilogue terms");
  }
  SLB_CHECK_ARG(!a->aux || (!a->swiglu && (a->aux_mode == 1 || a->aux_mode == 2)), "gemm: aux_mode must be 1 or 2 (no swiglu)");
  if (a->A2 && a->K2 > 0) {
    SLB_CHECK_ARG(!a->a_t && !a->b_t, "gemm: a second A source (A2) needs K-major operands");
    SLB_CHECK_ARG((a->K % 64) == 0 && (a->K2 % 8) == 0 && (a->lda2 % 8) == 0 && ((uintptr_t)a->A2 & 15) == 0,
                  "gemm: A2 needs K %% 64 == 0, K2 %% 8 == 0 and 16-byte aligned rows (K=%d K2=%d lda2=%lld)", a->K, a->K2, (long long)a->lda2);
  }
  if (a->block_n == 0 && !a->aux) {  // 1..4 activation rows: weight-streaming GEMV on the CUDA cores (HBM-bound)
    int rc = SLB_OK;
    if (slb_gemv_try(a, stream, &rc)) return rc;
    if (slb_skinny_try(a, stream, &rc)) return rc;  // 5..32 rows (batched decode): mma.sync weight streaming
  }
  SLB_CHECK_ARG(a->rms_weight == nullptr, "gemm: the fused RMSNorm prologue exists only on the M <= 4 weight-streaming path");
  SLB_CHECK_ARG(!a->a_fp32, "gemm: fp32 activation rows (a_fp32) are only taken by the fused RMSNorm prologue of the M <= 4 path");
  int bn = a->block_n;
  if (a->swiglu && bn != 2256 && bn != 256) bn = 0;
  if (bn == 0) {
    const int sms = slb_num_sms();
    const int mt = ceil_div(a->M, BM), mt2 = ceil_div(a->M, 2 * BM);
    const bool kmajor = !a->a_t && !a->b_t;
    if (kmajor && a->N <= 64 && a->M >= 1024) {
      bn = a->N <= 32 ? 32 : 64;  // tall-skinny (LoRA down-projection): the A stream is the whole cost
    } else {
      // Tile shape by estimated time = waves x (tile area per SM) / (relative efficiency of the kernel): the cluster kernel does
      // ~10 % more per SM-cycle than the 1-CTA 256-wide tile (6-stage ring, half the operand traffic per SM), the 128-wide tile
      // ~10 % less; what decides mid-size problems is wave quantisation (M = 4728, N = 896: 76 cluster tiles of 256 x 224 on 74
      // clusters = two waves for 1.03 waves of work, 148 tiles of 128 x 256 = exactly one).
      auto cost = [&](int tiles, int units, int area, float eff) { return (float)ceil_div(tiles, units) * (float)area / eff; };
      float best = cost(mt * ceil_div(a->N, 256), sms, 128 * 256, 1.0f);
      bn = 256;
      if (!a->swiglu) {
        const float c = cost(mt * ceil_div(a->N, 128), sms, 128 * 128, 0.9f);
        if (c < best) { best = c; bn = 128; }
      }
      // the cluster kernel also takes the dgrad (b_t) and wgrad (a_t + b_t) operand forms.  In the wgrad form M is a weight dimension
      // (a multiple of 256 for every matrix on the path), so there is no row padding to lose at 1024 <= M < 2048: the InternViT fc2
      // wgrad [1024 x 4096], K = 16 400 becomes 64 cluster tiles on 74 clusters instead of 128 tiles of the 1-CTA MN-major kernel
      const int min_m2 = (a->a_t && a->b_t && (a->M % 256) == 0) ? 1024 : 2048;
      if (a->M >= min_m2 && !(a->a_t && !a->b_t)) {
        const float c256 = cost(mt2 * ceil_div(a->N, 256), sms / 2, 128 * 256, 1.1f);
        if (c256 <= best) { best = c256; bn = 2256; }
        if (kmajor && !a->swiglu && (a->N % 224) == 0) {
          const float c224 = cost(mt2 * (a->N / 224), sms / 2, 128 * 224, 1.08f);
          if (c224 < best) { best = c224; bn = 2224; }
        }
      }
    }
  }
  if (bn == 2256 && (a->a_t || a->b_t)) {
    SLB_CHECK_ARG(!a->swiglu && !(a->A2 && a->K2 > 0), "gemm: SwiGLU / a second A source need K-major operands");
    return a->a_t ? launch_gemm2<256, 0, 1, 1>(a, stream) : launch_gemm2<256, 0, 0, 1>(a, stream);
  }
  if (bn == 2256 || bn == 2224 || bn == 2192) {
    SLB_CHECK_ARG(!a->a_t && !a->b_t, "gemm: the 224 / 192-wide cluster tiles take K-major operands only");
    SLB_CHECK_ARG(!a->swiglu || bn == 2256, "gemm: swiglu needs 256-wide tiles");
    if (tma_epilogue_ok(a)) {
      if (bn == 2256) return launch_gemm2<256, 1>(a, stream);
      if (bn == 2224) return launch_gemm2<224, 1>(a, stream);
      return launch_gemm2<192, 1>(a, stream);
    }
    if (bn == 2256) return launch_gemm2<256, 0>(a, stream);
    if (bn == 2224) return launch_gemm2<224, 0>(a, stream);
    return launch_gemm2<192, 0>(a, stream);
  }
  if (!a->a_t && !a->b_t && (bn == 32 || bn == 64)) return bn == 32 ? launch_gemm<32, 0, 0>(a, stream) : launch_gemm<64, 0, 0>(a, stream);
  SLB_CHECK_ARG(bn == 128 || bn == 256, "gemm: block_n must be 0, 32 / 64 (K-major), 128, 256 (1-CTA) or 2256 / 2224 / 2192 (2-CTA)");
  if (!a->a_t && !a->b_t) return bn == 256 ? launch_gemm<256, 0, 0>(a, stream) : launch_gemm<128, 0, 0>(a, stream);
  if (!a->a_t && a->b_t) return bn == 256 ? launch_gemm<256, 0, 1>(a, stream) : launch_gemm<128, 0, 1>(a, stream);
  return bn == 256 ? launch_gemm<256, 1, 1>(a, stream) : launch_gemm<128, 1, 1>(a, stream);
}
