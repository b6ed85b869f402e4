// Backward / training-only HBM-bound kernels: norm backward, activation backward, dropout, column reductions,
// RoPE backward, ViT embedding / pixel-shuffle backward, fused softmax cross-entropy, attention delta.
// (GEMM dgrad / wgrad reuse gemm.cu through the a_t / b_t operand forms; attention backward is attention_bwd.cu.)
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  size_t cap = (size_t)slb_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm backward.  One warp per row (grid-strided); dw / db partials live in registers and are
// flushed with one atomicAdd per column per warp at the end (ACC=true), or per row (ACC=false, wide rows).
// ------------------------------------------------------------------------------------------------
template <int MAXV, bool RMS, bool ACC>
__global__ void __launch_bounds__(kWarps * 32, ACC ? 2 : 1)
norm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ mean,
                const float* __restrict__ rstd, bf16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int rows, int cols,
                const bf16* __restrict__ dx_add) {
  // dx_add (optional): gradient arriving over the residual connection, added before the single rounding of dx.
  // ACC: per-thread fp32 partials of dw / db over the warp's rows, combined per block in shared memory (bank-conflict
  // free [e][vec] layout) and flushed with one global atomicAdd per column per block.  Pass 2 re-reads dy / x from L1
  // instead of holding them in registers, which keeps the kernel at two blocks per SM.
  __shared__ float s_acc[ACC ? 2 * MAXV * 32 * 8 : 1];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
  const int nvec = cols >> 3;
  float aw[ACC ? MAXV : 1][8], ab[ACC ? MAXV : 1][8];
  if (ACC) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) { aw[j][e] = 0.f; ab[j][e] = 0.f; }
    for (int i = threadIdx.x; i < 2 * MAXV * 32 * 8; i += blockDim.x) s_acc[i] = 0.f;
  }
  for (int row = gw; row < rows; row += nw) {
    const float mu = RMS ? 0.f : mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        float dyv[8], xv[8], wv[8];
        load8(dy + (size_t)row * cols + vi * 8, dyv);
        load8(x + (size_t)row * cols + vi * 8, xv);
        load8(w + vi * 8, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mu) * rs, g = dyv[e] * wv[e];
          s1 += g;
          s2 += g * xh;
          if (dw) {
            if (ACC) { aw[j][e] += dyv[e] * xh; if (!RMS) ab[j][e] += dyv[e]; }
            else { atomicAdd(dw + vi * 8 + e, dyv[e] * xh); if (!RMS && db) atomicAdd(db + vi * 8 + e, dyv[e]); }
          }
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const float m1 = RMS ? 0.f : s1 / cols, m2 = s2 / cols;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        float dyv[8], xv[8], wv[8], o[8];
        load8(dy + (size_t)row * cols + vi * 8, dyv);
        load8(x + (size_t)row * cols + vi * 8, xv);
        load8(w + vi * 8, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (dyv[e] * wv[e] - m1 - (xv[e] - mu) * rs * m2);
        if (dx_add) {
          float av[8];
          load8(dx_add + (size_t)row * cols + vi * 8, av);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += av[e];
        }
        store8(dx + (size_t)row * cols + vi * 8, o);
      }
    }
  }
  if (ACC && dw) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(&s_acc[e * (MAXV * 32) + vi], aw[j][e]);
          if (!RMS && db) atomicAdd(&s_acc[(8 + e) * (MAXV * 32) + vi], ab[j][e]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      atomicAdd(dw + i, s_acc[(i & 7) * (MAXV * 32) + (i >> 3)]);
      if (!RMS && db) atomicAdd(db + i, s_acc[(8 + (i & 7)) * (MAXV * 32) + (i >> 3)]);
    }
  }
}

// pixel-shuffle + LayerNorm(4096) backward: dy [T*256, 4096] -> dx scattered into the ViT stream [T*1025, 1024]
// (CLS rows receive zero).  One warp per row, grid-strided; dw / db partials are accumulated per block in shared memory
// ([e][vec] layout: conflict-free across lanes) and flushed with one global atomic per column per block.
__global__ void __launch_bounds__(kWarps * 32)
pixel_shuffle_ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                            const float* __restrict__ mean, const float* __restrict__ rstd, bf16* __restrict__ dx,
                            float* __restrict__ dw, float* __restrict__ db, int tiles) {
  constexpr int C = 1024, G = 32, G2 = 16, NT = 1025, COLS = 4096;
  __shared__ float s_dw[COLS], s_db[COLS];
  for (int i = threadIdx.x; i < COLS; i += blockDim.x) { s_dw[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * kWarps + (threadIdx.x >> 5); row < tiles * 256; row += gridDim.x * kWarps) {
    const int t = row / 256, ij = row % 256, i = ij / G2, j = ij % G2;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    // pass 1: reductions (values re-read in pass 2: the 4096-wide row does not fit in registers twice)
    for (int vi = lane; vi < 512; vi += 32) {
      const int q = vi >> 7, within = vi & 127;
      const int src_tok = (2 * i + (q >> 1)) * G + (2 * j + (q & 1));
      float dyv[8], xv[8], wv[8];
      load8(dy + (size_t)row * COLS + vi * 8, dyv);
      load8(x + ((size_t)t * NT + 1 + src_tok) * C + within * 8, xv);
      load8(w + vi * 8, wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xv[e] - mu) * rs, g = dyv[e] * wv[e];
        s1 += g;
        s2 += g * xh;
        atomicAdd(&s_dw[e * 512 + vi], dyv[e] * xh);
        atomicAdd(&s_db[e * 512 + vi], dyv[e]);
      }
    }
    s1 = warp_sum(s1) / COLS;
    s2 = warp_sum(s2) / COLS;
    for (int vi = lane; vi < 512; vi += 32) {
      const int q = vi >> 7, within = vi & 127;
      const int src_tok = (2 * i + (q >> 1)) * G + (2 * j + (q & 1));
      float dyv[8], xv[8], wv[8], o[8];
      load8(dy + (size_t)row * COLS + vi * 8, dyv);
      load8(x + ((size_t)t * NT + 1 + src_tok) * C + within * 8, xv);
      load8(w + vi * 8, wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rs * (dyv[e] * wv[e] - s1 - (xv[e] - mu) * rs * s2);
      store8(dx + ((size_t)t * NT + 1 + src_tok) * C + within * 8, o);
    }
    if (ij == 0) {  // zero the CLS row gradient of this tile
      for (int vi = lane; vi < 128; vi += 32) *reinterpret_cast<uint4*>(dx + (size_t)t * NT * C + vi * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < COLS; i += blockDim.x) {
    atomicAdd(dw + i, s_dw[(i & 7) * 512 + (i >> 3)]);
    atomicAdd(db + i, s_db[(i & 7) * 512 + (i >> 3)]);
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
__global__ void gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], r[8];
    load8(x + i * 8, a);
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = gelu_erf(a[e]);
    store8(y + i * 8, r);
  }
}
__global__ void gelu_bwd_kernel(const bf16* __restrict__ pre, const bf16* __restrict__ dout, bf16* __restrict__ dpre, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], d[8], r[8];
    load8(pre + i * 8, a);
    load8(dout + i * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float cdf = 0.5f * (1.0f + erff(a[e] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * a[e] * a[e]);
      r[e] = d[e] * (cdf + a[e] * pdf);
    }
    store8(dpre + i * 8, r);
  }
}
__global__ void silu_mul_bwd_kernel(const bf16* __restrict__ g, const bf16* __restrict__ u, const bf16* __restrict__ dout,
                                    bf16* __restrict__ dg, bf16* __restrict__ du, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8], d[8], rg[8], ru[8];
    load8(g + i * 8, a);
    load8(u + i * 8, b);
    load8(dout + i * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = 1.0f / (1.0f + __expf(-a[e]));
      ru[e] = d[e] * a[e] * sg;
      rg[e] = d[e] * b[e] * sg * (1.0f + a[e] * (1.0f - sg));
    }
    store8(dg + i * 8, rg);
    store8(du + i * 8, ru);
  }
}
// counter-based dropout (mask definition: drop_hash4 / drop_keep in common.cuh); scaled by 1 / P(keep).  The same call applied
// to the gradient (same seed) is the backward.
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t nvec, uint32_t thresh, float scale, uint64_t seed,
                               const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev << 16;
  const uint64_t sm = seed * kDropGold;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8];
    load8(x + i * 8, a);
    const uint64_t h0 = drop_hash4(sm, 2 * i), h1 = drop_hash4(sm, 2 * i + 1);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = drop_keep(e < 4 ? h0 : h1, e & 3, thresh) ? a[e] * scale : 0.f;
    store8(y + i * 8, a);
  }
}
// y += dropout(x): backward of the LoRA-input dropout, accumulated straight into the gradient of the shared input
__global__ void dropout_add_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t nvec, uint32_t thresh, float scale, uint64_t seed,
                                   const uint64_t* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev << 16;
  const uint64_t sm = seed * kDropGold;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8(x + i * 8, a);
    load8(y + i * 8, b);
    const uint64_t h0 = drop_hash4(sm, 2 * i), h1 = drop_hash4(sm, 2 * i + 1);
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] += drop_keep(e < 4 ? h0 : h1, e & 3, thresh) ? a[e] * scale : 0.f;
    store8(y + i * 8, b);
  }
}
// y (bf16) = [y +] x (fp32): flush of the fp32 small-parameter gradient accumulators into the flat bf16 gradient buffer
__global__ void flush_f32_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n, int accumulate) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i] + (accumulate ? __bfloat162float(y[i]) : 0.f));
}
__global__ void add_inplace_kernel(bf16* __restrict__ a_, const bf16* __restrict__ b_, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8(a_ + i * 8, a);
    load8(b_ + i * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    store8(a_ + i * 8, a);
  }
}
// dx_ls = dy * ls[col] (layer-scale backward for the branch), dls[col] += sum_r dy * branch
__global__ void scale_cols_kernel(const bf16* __restrict__ x, const bf16* __restrict__ s, bf16* __restrict__ y, int rows, int cols) {
  const int vpr = cols >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * vpr; i += (size_t)gridDim.x * blockDim.x) {
    const int vi = (int)(i % vpr);
    float a[8], b[8];
    load8(x + i * 8, a);
    load8(s + vi * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] *= b[e];
    store8(y + i * 8, a);
  }
}

__global__ void scale_cols_add_kernel(const bf16* __restrict__ x, const bf16* __restrict__ s, const bf16* __restrict__ res,
                                      bf16* __restrict__ y, int rows, int cols) {
  const int vpr = cols >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)rows * vpr; i += (size_t)gridDim.x * blockDim.x) {
    const int vi = (int)(i % vpr);
    float a[8], b[8], c[8];
    load8(x + i * 8, a);
    load8(s + vi * 8, b);
    load8(res + i * 8, c);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = c[e] + a[e] * b[e];
    store8(y + i * 8, a);
  }
}

// out[c] += sum_r a[r, c] * (b ? b[r, c] : 1), wide form (cols, strides multiples of 8, 16-byte aligned): block = 256 columns x a
// chunk of rows, 8 warps x (lane = 8 columns, one 16-byte load per row), 4 rows in flight per warp (the 64-column / 4-byte-load form
// below ran at 1.7 TB/s on the [16400, 3072..4096] bias gradients of the InternViT backward: 3.4 ms per step)
__global__ void __launch_bounds__(256)
col_reduce_wide_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, float* __restrict__ out, int rows, int cols, long long lda,
                       long long ldb, int rows_per_block, float alpha) {
  __shared__ float sh[8][256];
  const int c0 = blockIdx.x * 256, r0 = blockIdx.y * rows_per_block;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane * 8;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    const int r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + warp; r < r1; r += 32) {
      uint4 av[4], bv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = r + 8 * k;
        av[k] = rr < r1 ? *reinterpret_cast<const uint4*>(a + (size_t)rr * lda + c) : make_uint4(0, 0, 0, 0);
        if (b) bv[k] = rr < r1 ? *reinterpret_cast<const uint4*>(b + (size_t)rr * ldb + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(&av[k]);
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = unpack_bf16(aw[e]);
          if (b) {
            const float2 y = unpack_bf16(bw[e]);
            s[2 * e] += x.x * y.x; s[2 * e + 1] += x.y * y.y;
          } else { s[2 * e] += x.x; s[2 * e + 1] += x.y; }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[warp][lane * 8 + e] = s[e];
  __syncthreads();
  if (c0 + threadIdx.x < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
    atomicAdd(out + c0 + threadIdx.x, t * alpha);
  }
}

// out[c] += sum_r a[r, c] * (b ? b[r, c] : 1): block = 64 columns x a chunk of rows, 256 threads (8 row-lanes x 32 col-pairs)
__global__ void __launch_bounds__(256)
col_reduce_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, float* __restrict__ out, int rows, int cols, long long lda,
                  long long ldb, int rows_per_block, float alpha) {
  __shared__ float sh[8][64];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * rows_per_block;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = c0 + tx * 2;
  float s0 = 0.f, s1 = 0.f;
  if (c < cols) {
    const int r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 av = unpack_bf16(*reinterpret_cast<const uint32_t*>(a + (size_t)r * lda + c));
      if (b) {
        const float2 bv = unpack_bf16(*reinterpret_cast<const uint32_t*>(b + (size_t)r * ldb + c));
        s0 += av.x * bv.x; s1 += av.y * bv.y;
      } else { s0 += av.x; s1 += av.y; }
    }
  }
  sh[ty][tx * 2] = s0; sh[ty][tx * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
    atomicAdd(out + c0 + threadIdx.x, s * alpha);
  }
}

// Layer-scale residual backward (x_out = x + ls * branch, branch = Linear(..) + bias) in one pass:
//   dbranch = dx * ls (bf16 out);  dls[c] += sum_r dx * branch;  dbias[c] += sum_r dbranch
// (replaces col_reduce(dx, branch) + scale_cols + col_reduce(dbranch): three passes over [rows, cols])
__global__ void __launch_bounds__(256)
layerscale_bwd_kernel(const bf16* __restrict__ dx, const bf16* __restrict__ branch, const bf16* __restrict__ ls, bf16* __restrict__ dbranch,
                      float* __restrict__ dls, float* __restrict__ dbias, int rows, int cols, int rows_per_block) {
  __shared__ float sh[2][8][64];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * rows_per_block;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = c0 + tx * 2;
  float s0 = 0.f, s1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (c < cols) {
    const float2 lv = unpack_bf16(*reinterpret_cast<const uint32_t*>(ls + c));
    const int r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 dv = unpack_bf16(*reinterpret_cast<const uint32_t*>(dx + (size_t)r * cols + c));
      const float2 pv = unpack_bf16(*reinterpret_cast<const uint32_t*>(branch + (size_t)r * cols + c));
      const uint32_t o = pack_bf16(dv.x * lv.x, dv.y * lv.y);
      *reinterpret_cast<uint32_t*>(dbranch + (size_t)r * cols + c) = o;
      const float2 ov = unpack_bf16(o);  // the bias gradient sums what the wgrad / dgrad GEMMs will see
      s0 += dv.x * pv.x; s1 += dv.y * pv.y;
      b0 += ov.x; b1 += ov.y;
    }
  }
  sh[0][ty][tx * 2] = s0; sh[0][ty][tx * 2 + 1] = s1;
  sh[1][ty][tx * 2] = b0; sh[1][ty][tx * 2 + 1] = b1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, cc = threadIdx.x & 63;
    if (c0 + cc < cols) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += sh[which][k][cc];
      atomicAdd((which ? dbias : dls) + c0 + cc, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// ViT embedding backward: dpatch_out[t*1024+p] = dx[t, 1+p]; dcls += sum_t dx[t,0]; dpos[tok] += sum_t dx[t,tok]
// ------------------------------------------------------------------------------------------------
__global__ void vit_assemble_bwd_kernel(const bf16* __restrict__ dx, bf16* __restrict__ dpatch, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int tiles) {
  constexpr int C = 1024, NT = 1025, VPR = 128;
  const size_t total = (size_t)NT * VPR;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int vi = (int)(idx % VPR), tok = (int)(idx / VPR);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < tiles; ++t) {
      const uint4 u = *reinterpret_cast<const uint4*>(dx + ((size_t)t * NT + tok) * C + vi * 8);
      if (tok > 0) *reinterpret_cast<uint4*>(dpatch + ((size_t)t * 1024 + tok - 1) * C + vi * 8) = u;
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      dpos[(size_t)tok * C + vi * 8 + e] += acc[e];
      if (tok == 0) dcls[vi * 8 + e] += acc[e];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE backward + packing of the attention gradients into the fused-QKV gradient [B*L, (Hq+2Hkv)*64]:
// dq (fp32, post-RoPE space) and dk rotated back by -angle, dv copied.
// ------------------------------------------------------------------------------------------------
__global__ void rope_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                                bf16* __restrict__ dqkv, int batch, int lq, int hq, int hkv, float log2_theta) {
  const int heads = hq + 2 * hkv;
  const size_t total = (size_t)batch * lq * heads * 32;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx & 31);
    const int h = (int)((idx >> 5) % heads);
    const size_t row = (idx >> 5) / heads;
    const int i = (int)(row % lq), b = (int)(row / lq);
    bf16* dst = dqkv + row * (size_t)(heads * 64) + h * 64;
    if (h < hq + hkv) {
      const float* src = (h < hq) ? dq + (row * hq + h) * 64 : dk + (((size_t)b * hkv + (h - hq)) * lq + i) * 64;
      const float g0 = src[d], g1 = src[d + 32];
      float sn = 0.f, cs = 1.f;
      if (log2_theta > 0.f) {
        const float inv_freq = exp2f(-(float)(2 * d) / 64.0f * log2_theta);
        sincosf((float)i * inv_freq, &sn, &cs);
      }
      // forward: o0 = x0 c - x1 s ; o1 = x1 c + x0 s  =>  dx0 = g0 c + g1 s ; dx1 = g1 c - g0 s
      dst[d] = __float2bfloat16(g0 * cs + g1 * sn);
      dst[d + 32] = __float2bfloat16(g1 * cs - g0 * sn);
    } else {
      const float* src = dv + (((size_t)b * hkv + (h - hq - hkv)) * lq + i) * 64;
      dst[d] = __float2bfloat16(src[d]);
      dst[d + 32] = __float2bfloat16(src[d + 32]);
    }
  }
}

// delta[b, h, i] = sum_d dO[b*L+i, h*64+d] * O[...]  (8 lanes per (row, head): one 16-byte load of each operand per lane)
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta, int batch, int lq,
                                  int heads) {
  const size_t total = (size_t)batch * lq * heads;
  const size_t warp_units = (total + 3) / 4;  // a warp covers 4 consecutive (row, head) units; the tail is masked, not skipped
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  for (size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; w < warp_units; w += ((size_t)gridDim.x * blockDim.x) >> 5) {
    const size_t u = w * 4 + grp;
    const bool live = u < total;
    const int h = live ? (int)(u % heads) : 0;
    const size_t row = live ? u / heads : 0;
    float s = 0.f;
    if (live) {
      const size_t off = row * (size_t)(heads * 64) + h * 64 + sub * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + off);
      const uint4 b = *reinterpret_cast<const uint4*>(dout + off);
      const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
      const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
      s = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y) + (a3.x * b3.x + a3.y * b3.y);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && sub == 0) delta[((size_t)(row / lq) * heads + h) * lq + (row % lq)] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// fused softmax cross-entropy over fp32 logits: one block per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ce_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels, float* __restrict__ loss,
          bf16* __restrict__ dlogits, long long ldd, float grad_scale, int cols) {
  __shared__ float red[32];
  const float* r = logits + (size_t)blockIdx.x * ld;
  const long long label = labels[blockIdx.x];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, r[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int k = 1; k < 32; ++k) mx = fmaxf(mx, red[k]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s += __expf(r[i] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int k = 0; k < 32; ++k) s += red[k];
  const float lse = mx + logf(s);
  const bool valid = label >= 0 && label < cols;
  if (threadIdx.x == 0) loss[blockIdx.x] = valid ? lse - r[label] : 0.f;
  if (dlogits) {
    bf16* d = dlogits + (size_t)blockIdx.x * ldd;
    const float gs = valid ? grad_scale : 0.f;
    for (int i = threadIdx.x; i < (int)ldd; i += blockDim.x) {
      float g = 0.f;
      if (i < cols) g = (__expf(r[i] - lse) - (i == label ? 1.f : 0.f)) * gs;
      d[i] = __float2bfloat16(g);
    }
  }
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" int slb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                                 float* dw_accum, float* db_accum, int rows, int cols, const void* dx_add, void* stream) {
  SLB_CHECK_ARG(rows > 0 && (cols % 8) == 0 && cols <= 4096, "layernorm_bwd: bad shape %d x %d", rows, cols);
  const int grid = min(ceil_div(rows, kWarps), slb_num_sms() * (cols <= 1024 ? 4 : 2));
  if (cols <= 1024)
    norm_bwd_kernel<4, false, true><<<grid, kWarps * 32, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw_accum, db_accum, rows, cols, (const bf16*)dx_add);
  else
    norm_bwd_kernel<16, false, false><<<grid, kWarps * 32, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw_accum, db_accum, rows, cols, (const bf16*)dx_add);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw_accum, int rows,
                               int cols, const void* dx_add, void* stream) {
  SLB_CHECK_ARG(rows > 0 && (cols % 8) == 0 && cols <= 1024, "rmsnorm_bwd: bad shape %d x %d", rows, cols);
  const int grid = min(ceil_div(rows, kWarps), slb_num_sms() * 4);
  norm_bwd_kernel<4, true, true><<<grid, kWarps * 32, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, nullptr, rstd, (bf16*)dx, dw_accum, nullptr, rows, cols, (const bf16*)dx_add);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_pixel_shuffle_ln_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                                        float* dw_accum, float* db_accum, int tiles, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && dw_accum && db_accum, "pixel_shuffle_ln_bwd: bad args");
  pixel_shuffle_ln_bwd_kernel<<<min(ceil_div(tiles * 256, kWarps), slb_num_sms() * 2), kWarps * 32, 0, ST(stream)>>>(
      (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw_accum, db_accum, tiles);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_gelu_fwd(const void* x, void* y, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "gelu_fwd: n=%lld", (long long)n);
  gelu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_gelu_bwd(const void* pre, const void* dout, void* dpre, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "gelu_bwd: n=%lld", (long long)n);
  gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)pre, (const bf16*)dout, (bf16*)dpre, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_silu_mul_bwd(const void* gate, const void* up, const void* dout, void* dgate, void* dup, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "silu_mul_bwd: n=%lld", (long long)n);
  silu_mul_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)gate, (const bf16*)up, (const bf16*)dout, (bf16*)dgate, (bf16*)dup, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0 && p >= 0.f && p < 1.f, "dropout: n=%lld p=%f", (long long)n, p);
  dropout_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, n / 8, drop_thresh16(p), drop_scale(p), seed, seed_dev);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_dropout_add(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0 && p >= 0.f && p < 1.f, "dropout_add: n=%lld p=%f", (long long)n, p);
  dropout_add_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, n / 8, drop_thresh16(p), drop_scale(p), seed, seed_dev);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_flush_f32_to_bf16(const float* x, void* y, int64_t n, int accumulate, void* stream) {
  SLB_CHECK_ARG(n > 0, "flush: n=%lld", (long long)n);
  flush_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, (bf16*)y, n, accumulate);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_add_inplace_bf16(void* a, const void* b, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "add_inplace: n=%lld", (long long)n);
  add_inplace_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((bf16*)a, (const bf16*)b, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_scale_cols(const void* x, const void* s, void* out, int rows, int cols, void* stream) {
  SLB_CHECK_ARG(rows > 0 && (cols % 8) == 0, "scale_cols: %d x %d", rows, cols);
  scale_cols_kernel<<<grid_for((size_t)rows * (cols / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (const bf16*)s, (bf16*)out, rows, cols);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_scale_cols_add(const void* x, const void* s, const void* res, void* out, int rows, int cols, void* stream) {
  SLB_CHECK_ARG(rows > 0 && (cols % 8) == 0, "scale_cols_add: %d x %d", rows, cols);
  scale_cols_add_kernel<<<grid_for((size_t)rows * (cols / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (const bf16*)s, (const bf16*)res, (bf16*)out, rows, cols);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_col_reduce(const void* a, int64_t lda, const void* b, int64_t ldb, float* acc, int rows, int cols, float alpha,
                              void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 2) == 0 && (lda % 2) == 0 && (ldb % 2) == 0, "col_reduce: %d x %d", rows, cols);
  const bool wide = (cols % 8) == 0 && (lda % 8) == 0 && (ldb % 8) == 0 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0 && cols >= 256;
  if (wide) {
    int rpb = ceil_div(rows, max(1, (slb_num_sms() * 4) / ceil_div(cols, 256)));
    rpb = max(rpb, 64);
    dim3 grid(ceil_div(cols, 256), ceil_div(rows, rpb));
    col_reduce_wide_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, acc, rows, cols, lda, ldb, rpb, alpha);
    SLB_LAUNCH_CHECK();
    return SLB_OK;
  }
  int rpb = ceil_div(rows, max(1, (slb_num_sms() * 4) / ceil_div(cols, 64)));
  rpb = max(rpb, 64);
  dim3 grid(ceil_div(cols, 64), ceil_div(rows, rpb));
  col_reduce_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, acc, rows, cols, lda, ldb, rpb, alpha);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_layerscale_bwd(const void* dx, const void* branch, const void* ls, void* dbranch, float* dls_accum, float* dbias_accum,
                                  int rows, int cols, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 2) == 0 && dls_accum && dbias_accum, "layerscale_bwd: %d x %d", rows, cols);
  int rpb = ceil_div(rows, max(1, (slb_num_sms() * 4) / ceil_div(cols, 64)));
  rpb = max(rpb, 64);
  dim3 grid(ceil_div(cols, 64), ceil_div(rows, rpb));
  layerscale_bwd_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)dx, (const bf16*)branch, (const bf16*)ls, (bf16*)dbranch, dls_accum, dbias_accum,
                                                    rows, cols, rpb);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_vit_assemble_bwd(const void* dx, void* dpatch_out, float* dcls_accum, float* dpos_accum, int tiles, void* stream) {
  SLB_CHECK_ARG(tiles > 0, "vit_assemble_bwd: tiles=%d", tiles);
  vit_assemble_bwd_kernel<<<grid_for((size_t)1025 * 128, 128), 128, 0, ST(stream)>>>((const bf16*)dx, (bf16*)dpatch_out, dcls_accum, dpos_accum, tiles);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_rope_bwd(const float* dq, const float* dk, const float* dv, void* dqkv, int batch, int lq, int hq, int hkv, float theta,
                            void* stream) {
  SLB_CHECK_ARG(batch > 0 && lq > 0, "rope_bwd: bad shape");
  const size_t total = (size_t)batch * lq * (hq + 2 * hkv) * 32;
  rope_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(dq, dk, dv, (bf16*)dqkv, batch, lq, hq, hkv, theta > 1.f ? log2f(theta) : 0.f);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_attn_delta(const void* o, const void* dout, float* delta, int batch, int lq, int heads, void* stream) {
  SLB_CHECK_ARG(batch > 0 && lq > 0 && heads > 0, "attn_delta: bad shape");
  SLB_CHECK_ARG(o && dout && delta && (((uintptr_t)o | (uintptr_t)dout) & 15) == 0, "attn_delta: operands must be 16-byte aligned");
  attn_delta_kernel<<<grid_for((size_t)batch * lq * heads * 8, 256), 256, 0, ST(stream)>>>((const bf16*)o, (const bf16*)dout, delta, batch, lq, heads);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* labels, float* loss, void* dlogits, int64_t ldd,
                              float grad_scale, int rows, int cols, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (!dlogits || ldd >= cols), "ce: bad shape");
  ce_kernel<<<rows, 1024, 0, ST(stream)>>>(logits, ld, (const long long*)labels, loss, (bf16*)dlogits, ldd, grad_scale, cols);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
