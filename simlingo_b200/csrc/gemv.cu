// Weight-streaming GEMV for 1..4 activation rows (batch-1 decode: Qwen2 q|k|v, o, gate|up + SwiGLU, down, LM head).
// HBM-bound: every weight row is read exactly once with 16-byte loads, the activation rows sit in shared memory,
// products accumulate in fp32 on the CUDA cores, one warp owns 4 weight rows at a time (4 independent load streams
// per lane), warp-shuffle reduction, same epilogue as the tcgen05 GEMM (alpha, bias, activation, column scale,
// residual, fused SwiGLU over the 128-row interleaved gate|up layout, bf16 or fp32 output).
// Algorithmic bytes = N*K*2 (weights) + M*K*2 + M*N*{2,4}; replaces slb_gemm_bf16's tensor-core path when M <= 4, where
// a 128-row MMA tile would be > 96 % padding and the fixed per-launch cost of the TMA/TMEM pipeline dominates.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

struct GemvParams {
  const bf16* A; long long lda;
  const bf16* W; long long ldw;
  void* out; long long ldo;
  const bf16* bias; const bf16* scale_n;
  const void* res; long long ldr;
  int N, K;
  float alpha;
  int act, swiglu, out_fp32;
  const bf16* rms_w; float rms_eps;
  int a_fp32;  // A holds fp32 rows (the fp32 residual stream entering the fused RMSNorm prologue)
  const bf16* A2; long long lda2; int K2;  // second activation source appended along k (weights hold K + K2 columns per row)
};

constexpr int kRows = 4;      // weight rows per warp step
constexpr int kGemvWarps = 8;

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == SLB_ACT_GELU) return gelu_erf(v);
  if (act == SLB_ACT_SILU) return silu(v);
  if (act == SLB_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

template <int M>
__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_bf16_kernel(GemvParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[kGemvWarps];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);  // [M][K + K2]
  const int K = p.K, KT = p.K + p.K2, kvec = KT >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_groups = p.swiglu ? p.N / 4 : (p.N + kRows - 1) / kRows;  // groups of 4 weight rows

  auto group_rows = [&](int g, int (&rows)[kRows]) {
    if (!p.swiglu) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) rows[r] = min(g * kRows + r, p.N - 1);
    } else {  // outputs j0, j0+1 need gate rows (256 t + w) and up rows (256 t + 128 + w)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = g * 2 + r, t = j >> 7, w = j & 127;
        rows[r] = 256 * t + w;
        rows[2 + r] = 256 * t + 128 + w;
      }
    }
  };
  // 4 k-steps x 4 rows = 16 independent 16-byte loads in flight per lane
  auto load_w = [&](const int (&rows)[kRows], int v0, uint4 (&wv)[4][kRows]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u * 32 + lane;
      const int vc = v < kvec ? v : 0;
#pragma unroll
      for (int r = 0; r < kRows; ++r) wv[u][r] = __ldg(reinterpret_cast<const uint4*>(p.W + (size_t)rows[r] * p.ldw) + vc);
    }
  };

  pdl_trigger();   // the next kernel of the decode chain may become resident (it prefetches its own weights, then waits for us)
  int g = blockIdx.x * kGemvWarps + warp;
  int rows[kRows];
  uint4 wv[4][kRows];
  if (g < n_groups) {  // the first weight loads do not depend on the activations: issue them before staging x
    group_rows(g, rows);
    load_w(rows, 0, wv);
  }
  pdl_wait();      // the activations (and the residual read by the epilogue) come from the previous kernel
  if (p.a_fp32) {
    // fp32 activation rows (residual stream): RMSNorm in fp32 straight from global memory (M * K * 4 bytes, L2-resident), one
    // rounding to bf16 when the normalised row is staged
    const float* af = reinterpret_cast<const float*>(p.A);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float s = 0.f;
      for (int k = threadIdx.x; k < K; k += blockDim.x) { const float v = af[(size_t)m * p.lda + k]; s += v * v; }
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGemvWarps; ++w) t += red[w];
      const float rstd = rsqrtf(t / K + p.rms_eps);
      for (int k = threadIdx.x; k < K; k += blockDim.x)
        xs[(size_t)m * KT + k] = __float2bfloat16(af[(size_t)m * p.lda + k] * rstd * __bfloat162float(p.rms_w[k]));
      __syncthreads();
    }
  } else {
  const int kvec1 = K >> 3;
  for (int i = threadIdx.x; i < M * kvec1; i += blockDim.x) {
    const int m = i / kvec1, v = i % kvec1;
    reinterpret_cast<uint4*>(xs + (size_t)m * KT)[v] = *reinterpret_cast<const uint4*>(p.A + (size_t)m * p.lda + v * 8);
  }
  __syncthreads();
  }
  if (p.K2 > 0) {   // the appended activations (LoRA down-projection outputs) are never normalised
    const int kvec2 = p.K2 >> 3;
    for (int i = threadIdx.x; i < M * kvec2; i += blockDim.x) {
      const int m = i / kvec2, v = i % kvec2;
      reinterpret_cast<uint4*>(xs + (size_t)m * KT + K)[v] = *reinterpret_cast<const uint4*>(p.A2 + (size_t)m * p.lda2 + v * 8);
    }
    __syncthreads();
  }
  if (p.rms_w && !p.a_fp32) {
    // fused Qwen2RMSNorm of the activation rows (same arithmetic as norm_fwd_kernel: x * rstd * w in fp32, one rounding)
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float s = 0.f;
      for (int k = threadIdx.x; k < K; k += blockDim.x) { const float v = __bfloat162float(xs[(size_t)m * KT + k]); s += v * v; }
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGemvWarps; ++w) t += red[w];
      const float rstd = rsqrtf(t / K + p.rms_eps);
      for (int k = threadIdx.x; k < K; k += blockDim.x)
        xs[(size_t)m * KT + k] = __float2bfloat16(__bfloat162float(xs[(size_t)m * KT + k]) * rstd * __bfloat162float(p.rms_w[k]));
      __syncthreads();
    }
  }
  for (; g < n_groups; g += gridDim.x * kGemvWarps) {
    float acc[kRows][M];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
    for (int v0 = 0; v0 < kvec; v0 += 4 * 32) {
      if (v0 > 0) load_w(rows, v0, wv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v >= kvec) continue;
        float xf[M][8];
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const uint4 xv = reinterpret_cast<const uint4*>(xs + (size_t)m * KT)[v];
          const float2 a = unpack_bf16(xv.x), b = unpack_bf16(xv.y), c = unpack_bf16(xv.z), d = unpack_bf16(xv.w);
          xf[m][0] = a.x; xf[m][1] = a.y; xf[m][2] = b.x; xf[m][3] = b.y; xf[m][4] = c.x; xf[m][5] = c.y; xf[m][6] = d.x; xf[m][7] = d.y;
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const float2 a = unpack_bf16(wv[u][r].x), b = unpack_bf16(wv[u][r].y), c = unpack_bf16(wv[u][r].z), d = unpack_bf16(wv[u][r].w);
          const float wf[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
          for (int m = 0; m < M; ++m)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[r][m] = fmaf(wf[e], xf[m][e], acc[r][m]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
    // lane r writes output r of the group
    if (!p.swiglu) {
      const int n = g * kRows + lane;
      if (lane < kRows && n < p.N) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < kRows; ++r) v = (lane == r) ? acc[r][m] : v;
          v *= p.alpha;
          if (p.bias) v += __bfloat162float(p.bias[n]);
          v = act_apply(v, p.act);
          if (p.scale_n) v *= __bfloat162float(p.scale_n[n]);
          if (p.out_fp32) {
            if (p.res) v += reinterpret_cast<const float*>(p.res)[(size_t)m * p.ldr + n];
            reinterpret_cast<float*>(p.out)[(size_t)m * p.ldo + n] = v;
          } else {
            if (p.res) v += __bfloat162float(reinterpret_cast<const bf16*>(p.res)[(size_t)m * p.ldr + n]);
            reinterpret_cast<bf16*>(p.out)[(size_t)m * p.ldo + n] = __float2bfloat16(v);
          }
        }
      }
    } else if (lane < 2) {
      const int j = g * 2 + lane;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float gate = (lane == 0 ? acc[0][m] : acc[1][m]) * p.alpha, up = (lane == 0 ? acc[2][m] : acc[3][m]) * p.alpha;
        const float v = silu(gate) * up;
        if (p.out_fp32) reinterpret_cast<float*>(p.out)[(size_t)m * p.ldo + j] = v;
        else reinterpret_cast<bf16*>(p.out)[(size_t)m * p.ldo + j] = __float2bfloat16(v);
      }
    }
    const int gn = g + gridDim.x * kGemvWarps;
    if (gn < n_groups) {
      group_rows(gn, rows);
      load_w(rows, 0, wv);
    }
  }
}

template <int M>
int launch_gemv(const GemvParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)M * (p.K + p.K2) * 2;
  auto kern = gemv_bf16_kernel<M>;
  if (smem > 48 * 1024) SLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_groups = p.swiglu ? p.N / 4 : ceil_div(p.N, kRows);
  int grid = ceil_div(n_groups, kGemvWarps);
  const int cap = slb_num_sms() * 8;
  if (grid > cap) grid = cap;
  SLB_CUDA(slb_launch_pdl(true, kern, dim3(grid), dim3(kGemvWarps * 32), smem, stream, p));
  return SLB_OK;
}

}  // namespace

// called by slb_gemm_bf16 for M <= 4 with K-major operands; returns 1 if it took the problem (rc in *rc_out)
int slb_gemv_try(const slb_gemm_args* a, cudaStream_t stream, int* rc_out) {
  if (a->M > 4 || a->a_t || a->b_t || (a->K % 8) != 0 || (size_t)a->M * (a->K + (a->A2 ? a->K2 : 0)) * 2 > 200 * 1024) return 0;
  if (a->swiglu && (a->N % 256) != 0) return 0;
  if ((((uintptr_t)a->A) & 15) || (((uintptr_t)a->B) & 15) || (a->lda % 8) || (a->ldb % 8)) return 0;
  if (a->a_fp32 && !a->rms_weight) return 0;
  GemvParams p;
  p.a_fp32 = a->a_fp32;
  p.A2 = nullptr; p.lda2 = 0; p.K2 = 0;
  if (a->A2 && a->K2 > 0) {
    if ((a->K2 % 8) || (a->lda2 % 8) || (((uintptr_t)a->A2) & 15)) return 0;
    p.A2 = (const bf16*)a->A2; p.lda2 = a->lda2; p.K2 = a->K2;
  }
  p.A = (const bf16*)a->A; p.lda = a->lda;
  p.W = (const bf16*)a->B; p.ldw = a->ldb;
  p.out = a->out; p.ldo = a->ldo;
  p.bias = (const bf16*)a->bias; p.scale_n = (const bf16*)a->scale_n;
  p.res = a->residual; p.ldr = a->ldr;
  p.N = a->N; p.K = a->K; p.alpha = a->alpha; p.act = a->act; p.swiglu = a->swiglu; p.out_fp32 = a->out_fp32;
  p.rms_w = (const bf16*)a->rms_weight; p.rms_eps = a->rms_eps;
  switch (a->M) {
    case 1: *rc_out = launch_gemv<1>(p, stream); break;
    case 2: *rc_out = launch_gemv<2>(p, stream); break;
    case 3: *rc_out = launch_gemv<3>(p, stream); break;
    default: *rc_out = launch_gemv<4>(p, stream); break;
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------------------------------
// Skinny GEMM for 4 < M <= 32 activation rows (batched decode): weight-streaming on mma.sync.m16n8k16 (bf16, fp32 acc).
// A 128-row tcgen05 tile would be >= 75 % padding here and the whole problem is HBM-bound on the weights, so the goal is
// the same as for the GEMV: read every weight row once with 16-byte loads, many CTAs.  CTA = 8 warps = 16 weight rows
// (two n8 tiles; for SwiGLU: 8 gate rows + their 8 up rows); the warps split K in interleaved 32-element chunks and
// their partial tiles are summed through shared memory.  Within a chunk, lane (g = lane/4, q = lane%4) loads the 8
// consecutive k values k0 + 8q .. 8q+7 of weight row g (one uint4) and of activation rows g, g+8 (+16, +24): a fixed
// permutation of k applied to both operands, so the two MMAs of the chunk see consistent fragments without any shuffles.
// Activations (<= 32 x K bf16) are read through L1.  Same epilogue contract as the GEMV.
// ------------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MT, int NT>  // 16-row activation tiles (M <= 16 * MT), n8 weight tiles per CTA (8 * NT weight rows)
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(GemvParams p, int M) {
  __shared__ float part[8][MT * 16][NT * 8 + 1];  // per-warp partial tiles (padded rows)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int tile = blockIdx.x;
  // weight rows of the n8 tiles.  SwiGLU: tiles [0, NT/2) are gate rows, [NT/2, NT) the matching up rows
  const uint4* wrow[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int row;
    if (!p.swiglu) {
      row = min(tile * (8 * NT) + nt * 8 + g, p.N - 1);
    } else {
      const int j = tile * (4 * NT) + (nt % (NT / 2)) * 8 + g, t = j >> 7, w = j & 127;
      row = 256 * t + (nt >= NT / 2 ? 128 : 0) + w;
    }
    wrow[nt] = reinterpret_cast<const uint4*>(p.W + (size_t)row * p.ldw) + q;
  }
  const uint4* arow[MT][2];
  const uint4* arow2[MT][2];   // second activation source (chunks beyond K)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    arow[mt][0] = reinterpret_cast<const uint4*>(p.A + (size_t)min(mt * 16 + g, M - 1) * p.lda) + q;
    arow[mt][1] = reinterpret_cast<const uint4*>(p.A + (size_t)min(mt * 16 + 8 + g, M - 1) * p.lda) + q;
    arow2[mt][0] = p.K2 ? reinterpret_cast<const uint4*>(p.A2 + (size_t)min(mt * 16 + g, M - 1) * p.lda2) + q : arow[mt][0];
    arow2[mt][1] = p.K2 ? reinterpret_cast<const uint4*>(p.A2 + (size_t)min(mt * 16 + 8 + g, M - 1) * p.lda2) + q : arow[mt][1];
  }
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  const int nchunks1 = p.K >> 5, nchunks = (p.K + p.K2) >> 5;
  // programmatic dependent launch: pull this CTA's weight rows towards L2 while the producer of the activations is still running
  pdl_trigger();
  {
    const int lines = ((p.K + p.K2) * 2 + 127) >> 7;   // 128-byte lines per weight row
    for (int i = threadIdx.x; i < NT * 8 * lines; i += blockDim.x) {
      const int rr = i / lines, ln = i % lines;
      int row;
      if (!p.swiglu) {
        row = min(tile * (8 * NT) + rr, p.N - 1);
      } else {
        const int nt = rr >> 3, gg = rr & 7;
        const int j = tile * (4 * NT) + (nt % (NT / 2)) * 8 + gg, t = j >> 7, w = j & 127;
        row = 256 * t + (nt >= NT / 2 ? 128 : 0) + w;
      }
      prefetch_l2(reinterpret_cast<const uint8_t*>(p.W + (size_t)row * p.ldw) + (size_t)ln * 128);
    }
  }
  pdl_wait();
#pragma unroll(NT <= 2 ? 4 : 2)
  for (int c = warp; c < nchunks; c += 8) {
    uint4 bw[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bw[nt] = __ldg(wrow[nt] + c * 4);
    const bool tail = c >= nchunks1;
    const int ca = tail ? c - nchunks1 : c;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint4 lo = __ldg((tail ? arow2[mt][0] : arow[mt][0]) + ca * 4), hi = __ldg((tail ? arow2[mt][1] : arow[mt][1]) + ca * 4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma_bf16_16816(acc[mt][nt], lo.x, hi.x, lo.y, hi.y, bw[nt].x, bw[nt].y);
        mma_bf16_16816(acc[mt][nt], lo.z, hi.z, lo.w, hi.w, bw[nt].z, bw[nt].w);
      }
    }
  }
  // C fragment: c0,c1 -> (row g, cols 2q, 2q+1), c2,c3 -> (row g+8, same cols)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      part[warp][mt * 16 + g][nt * 8 + 2 * q] = acc[mt][nt][0];
      part[warp][mt * 16 + g][nt * 8 + 2 * q + 1] = acc[mt][nt][1];
      part[warp][mt * 16 + 8 + g][nt * 8 + 2 * q] = acc[mt][nt][2];
      part[warp][mt * 16 + 8 + g][nt * 8 + 2 * q + 1] = acc[mt][nt][3];
    }
  __syncthreads();
  if (!p.swiglu) {
    for (int i = threadIdx.x; i < MT * 16 * NT * 8; i += blockDim.x) {
      const int m = i / (NT * 8), c = i % (NT * 8), n = tile * (NT * 8) + c;
      if (m >= M || n >= p.N) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][m][c];
      v *= p.alpha;
      if (p.bias) v += __bfloat162float(p.bias[n]);
      v = act_apply(v, p.act);
      if (p.scale_n) v *= __bfloat162float(p.scale_n[n]);
      if (p.out_fp32) {
        if (p.res) v += reinterpret_cast<const float*>(p.res)[(size_t)m * p.ldr + n];
        reinterpret_cast<float*>(p.out)[(size_t)m * p.ldo + n] = v;
      } else {
        if (p.res) v += __bfloat162float(reinterpret_cast<const bf16*>(p.res)[(size_t)m * p.ldr + n]);
        reinterpret_cast<bf16*>(p.out)[(size_t)m * p.ldo + n] = __float2bfloat16(v);
      }
    }
  } else {
    for (int i = threadIdx.x; i < MT * 16 * NT * 4; i += blockDim.x) {
      const int m = i / (NT * 4), c = i % (NT * 4), j = tile * (NT * 4) + c;
      if (m >= M) continue;
      float gate = 0.f, up = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { gate += part[w][m][c]; up += part[w][m][NT * 4 + c]; }
      const float v = silu(gate * p.alpha) * (up * p.alpha);
      if (p.out_fp32) reinterpret_cast<float*>(p.out)[(size_t)m * p.ldo + j] = v;
      else reinterpret_cast<bf16*>(p.out)[(size_t)m * p.ldo + j] = __float2bfloat16(v);
    }
  }
}

}  // namespace

// called by slb_gemm_bf16 for 4 < M <= 32 with K-major operands; returns 1 if it took the problem
int slb_skinny_try(const slb_gemm_args* a, cudaStream_t stream, int* rc_out) {
  if (a->M <= 4 || a->M > 32 || a->a_t || a->b_t || (a->K % 32) != 0 || a->rms_weight || a->aux || a->a_fp32) return 0;
  const bool has2 = a->A2 && a->K2 > 0;
  if (has2 && ((a->K2 % 32) || (a->lda2 % 8) || (((uintptr_t)a->A2) & 15))) return 0;
  if (a->swiglu && (a->N % 256) != 0) return 0;
  if ((((uintptr_t)a->A) & 15) || (((uintptr_t)a->B) & 15) || (a->lda % 8) || (a->ldb % 8)) return 0;
  GemvParams p;
  p.A = (const bf16*)a->A; p.lda = a->lda;
  p.W = (const bf16*)a->B; p.ldw = a->ldb;
  p.out = a->out; p.ldo = a->ldo;
  p.bias = (const bf16*)a->bias; p.scale_n = (const bf16*)a->scale_n;
  p.res = a->residual; p.ldr = a->ldr;
  p.N = a->N; p.K = a->K; p.alpha = a->alpha; p.act = a->act; p.swiglu = a->swiglu; p.out_fp32 = a->out_fp32;
  p.rms_w = nullptr; p.rms_eps = 0.f; p.a_fp32 = 0;
  p.A2 = has2 ? (const bf16*)a->A2 : nullptr; p.lda2 = has2 ? a->lda2 : 0; p.K2 = has2 ? a->K2 : 0;
  // 32 weight rows per CTA (the activations are re-read from L2 by every CTA: wider tiles halve that traffic); 16 for the
  // small projections so that they still spread over > 50 CTAs
  const bool wide = a->N >= 4096;
  const int rows_per_cta = wide ? 32 : 16;
  const int grid = a->swiglu ? a->N / rows_per_cta : ceil_div(a->N, rows_per_cta);
  if (a->M <= 16) {
    if (wide) slb_launch_pdl(true, skinny_gemm_kernel<1, 4>, dim3(grid), dim3(256), 0, stream, p, a->M);
    else slb_launch_pdl(true, skinny_gemm_kernel<1, 2>, dim3(grid), dim3(256), 0, stream, p, a->M);
  } else {
    if (wide) slb_launch_pdl(true, skinny_gemm_kernel<2, 4>, dim3(grid), dim3(256), 0, stream, p, a->M);
    else slb_launch_pdl(true, skinny_gemm_kernel<2, 2>, dim3(grid), dim3(256), 0, stream, p, a->M);
  }
  cudaError_t e = cudaGetLastError();
  *rc_out = (e == cudaSuccess) ? SLB_OK : slb_fail(SLB_ECUDA, "skinny gemm launch: %s", cudaGetErrorString(e));
  return 1;
}
