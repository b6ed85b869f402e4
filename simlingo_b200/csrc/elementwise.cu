// HBM-bound kernels of the path: LayerNorm / RMSNorm, ViT embedding glue, pixel-shuffle+LN, RoPE + KV-cache
// write, embedding gather / placeholder substitution, argmax, small fused MLP heads.  All are vectorised
// (128-bit loads/stores), warp-shuffle reduced, one warp per row where a row reduction is needed.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm forward: one warp per row, row kept in registers (MAXV 8-wide vectors per lane)
// ------------------------------------------------------------------------------------------------
template <int MAXV, bool RMS, typename XT = bf16>  // XT = float: the fp32 residual stream of the Qwen2 inference path
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
norm_fwd_kernel(const XT* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b, bf16* __restrict__ y,
                int rows, int cols, float eps, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = cols >> 3;
  const XT* xr = x + (size_t)row * cols;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nvec) {
      load8(xr + vi * 8, v[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += RMS ? v[j][e] * v[j][e] : v[j][e];
    }
  }
  s = warp_sum(s);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(s / cols + eps);
  } else {
    mean = s / cols;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { float d = v[j][e] - mean; q += d * d; }
      }
    }
    q = warp_sum(q);
    rstd = rsqrtf(q / cols + eps);
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* yr = y + (size_t)row * cols;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nvec) {
      float wv[8], o[8];
      load8(w + vi * 8, wv);
      if (RMS) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = v[j][e] * rstd * wv[e];
      } else {
        float bv[8];
        load8(b + vi * 8, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[j][e] - mean) * rstd * wv[e] + bv[e];
      }
      store8(yr + vi * 8, o);
    }
  }
}

// drop CLS + pixel_shuffle(0.5,'v2') + LayerNorm(4096): out row (t, i, j) gathers the 2x2 block of
// patch tokens [(2i,2j),(2i,2j+1),(2i+1,2j),(2i+1,2j+1)] channel-concatenated.
template <typename XT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pixel_shuffle_ln_kernel(const XT* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                        bf16* __restrict__ y, int tiles, float eps, float* __restrict__ mean_out,
                        float* __restrict__ rstd_out) {
  constexpr int C = 1024, G = 32, G2 = 16, NT = 1025, COLS = 4096, MAXV = 16;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= tiles * G2 * G2) return;
  const int t = row / (G2 * G2), ij = row % (G2 * G2), i = ij / G2, j = ij % G2;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int jj = 0; jj < MAXV; ++jj) {
    const int vi = lane + 32 * jj;       // 0..511
    const int q = vi >> 7, within = vi & 127;
    const int src_tok = (2 * i + (q >> 1)) * G + (2 * j + (q & 1));
    load8(x + ((size_t)t * NT + 1 + src_tok) * C + within * 8, v[jj]);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[jj][e];
  }
  s = warp_sum(s);
  const float mean = s / COLS;
  float qv = 0.f;
#pragma unroll
  for (int jj = 0; jj < MAXV; ++jj)
#pragma unroll
    for (int e = 0; e < 8; ++e) { float d = v[jj][e] - mean; qv += d * d; }
  qv = warp_sum(qv);
  const float rstd = rsqrtf(qv / COLS + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* yr = y + (size_t)row * COLS;
#pragma unroll
  for (int jj = 0; jj < MAXV; ++jj) {
    const int vi = lane + 32 * jj;
    float wv[8], bv[8], o[8];
    load8(w + vi * 8, wv);
    load8(b + vi * 8, bv);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (v[jj][e] - mean) * rstd * wv[e] + bv[e];
    store8(yr + vi * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------
// ViT embedding glue
// ------------------------------------------------------------------------------------------------
__global__ void im2col_patch_kernel(const bf16* __restrict__ px, bf16* __restrict__ out, int tiles, int kpad) {
  // one thread per 2 output elements (k even): out[(t*1024 + py*32 + pxx), k], k = c*196 + ky*14 + kx
  const size_t total = (size_t)tiles * 1024 * (kpad / 2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % (kpad / 2)) * 2;
    const size_t rowi = idx / (kpad / 2);
    const int p = (int)(rowi % 1024), t = (int)(rowi / 1024);
    const int py = p >> 5, pxx = p & 31;
    float v0 = 0.f, v1 = 0.f;
    if (kk < 588) {
      // kk even and 14 even => kk and kk+1 are in the same kernel row
      const int c = kk / 196, r = kk % 196, ky = r / 14, kx = r % 14;
      const bf16* src = px + (((size_t)t * 3 + c) * 448 + (py * 14 + ky)) * 448 + pxx * 14 + kx;
      bf162 two = *reinterpret_cast<const bf162*>(src);
      v0 = __bfloat162float(two.x);
      v1 = __bfloat162float(two.y);
    }
    *reinterpret_cast<uint32_t*>(out + rowi * kpad + kk) = pack_bf16(v0, v1);
  }
}

template <typename OT>
__global__ void vit_assemble_kernel(const bf16* __restrict__ patch_out, const bf16* __restrict__ cls,
                                    const bf16* __restrict__ pos, OT* __restrict__ x, int tiles) {
  constexpr int C = 1024, NT = 1025, VPR = C / 8;
  const size_t total = (size_t)tiles * NT * VPR;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int vi = (int)(idx % VPR);
    const size_t r = idx / VPR;
    const int tok = (int)(r % NT), t = (int)(r / NT);
    float a[8], pz[8], o[8];
    if (tok == 0) load8(cls + vi * 8, a);
    else load8(patch_out + ((size_t)t * 1024 + tok - 1) * C + vi * 8, a);
    load8(pos + (size_t)tok * C + vi * 8, pz);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = a[e] + pz[e];
    store8(x + r * C + vi * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE (rotate_half convention) on q and k + KV-cache write.  One thread per (row, group of 8 rotation pairs): the 8
// sin / cos values depend on the position only and are reused for every head, all traffic is 16-byte vectors
// (x[d..d+7] and its rotation partner x[d+32..d+39]); q is rotated in place, rotated k and plain v go to the caches.
// ------------------------------------------------------------------------------------------------
__global__ void rope_kv_write_kernel(bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc, int batch, int lq,
                                     int past, int lmax, int hq, int hkv, float log2_theta, const int* __restrict__ past_dev, int per_head) {
  pdl_trigger();
  pdl_wait();   // qkv comes from the projection kernel right before us
  if (past_dev) past = *past_dev;
  const int heads = hq + 2 * hkv;
  // few rows (decode, query append): one thread per (row, head, 8 pairs) - with one thread per (row, 8 pairs) walking over the 18
  // heads the single-token call was a chain of 36 dependent 16-byte round trips (8 us); many rows: the head loop amortises sin / cos
  const int hsplit = per_head ? heads : 1;
  const size_t total = (size_t)batch * lq * 4 * hsplit;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 3);          // dims 8c .. 8c+7 (and +32)
    const size_t rh = idx >> 2;
    const size_t row = rh / hsplit;        // b*lq + i
    const int h_lo = per_head ? (int)(rh % hsplit) : 0, h_hi = per_head ? h_lo + 1 : heads;
    const int i = (int)(row % lq), b = (int)(row / lq);
    float sn[8], cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float inv_freq = exp2f(-(float)(2 * (c * 8 + e)) / 64.0f * log2_theta);
      sincosf((float)(past + i) * inv_freq, &sn[e], &cs[e]);
    }
    bf16* rowp = qkv + row * (size_t)(heads * 64);
    for (int h = h_lo; h < h_hi; ++h) {
      bf16* src = rowp + h * 64 + c * 8;
      const uint4 lo = *reinterpret_cast<const uint4*>(src), hi = *reinterpret_cast<const uint4*>(src + 32);
      if (h >= hq + hkv) {  // v: plain copy into the cache
        bf16* dst = vc + (((size_t)b * hkv + (h - hq - hkv)) * lmax + past + i) * 64 + c * 8;
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + 32) = hi;
        continue;
      }
      float x0[8], x1[8];
      {
        const float2 a = unpack_bf16(lo.x), bb = unpack_bf16(lo.y), cc = unpack_bf16(lo.z), d = unpack_bf16(lo.w);
        x0[0] = a.x; x0[1] = a.y; x0[2] = bb.x; x0[3] = bb.y; x0[4] = cc.x; x0[5] = cc.y; x0[6] = d.x; x0[7] = d.y;
        const float2 a2 = unpack_bf16(hi.x), b2 = unpack_bf16(hi.y), c2 = unpack_bf16(hi.z), d2 = unpack_bf16(hi.w);
        x1[0] = a2.x; x1[1] = a2.y; x1[2] = b2.x; x1[3] = b2.y; x1[4] = c2.x; x1[5] = c2.y; x1[6] = d2.x; x1[7] = d2.y;
      }
      float o0[8], o1[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o0[e] = x0[e] * cs[e] - x1[e] * sn[e];
        o1[e] = x1[e] * cs[e] + x0[e] * sn[e];
      }
      uint4 r0, r1;
      r0.x = pack_bf16(o0[0], o0[1]); r0.y = pack_bf16(o0[2], o0[3]); r0.z = pack_bf16(o0[4], o0[5]); r0.w = pack_bf16(o0[6], o0[7]);
      r1.x = pack_bf16(o1[0], o1[1]); r1.y = pack_bf16(o1[2], o1[3]); r1.z = pack_bf16(o1[4], o1[5]); r1.w = pack_bf16(o1[6], o1[7]);
      bf16* dst = (h < hq) ? src : kc + (((size_t)b * hkv + (h - hq)) * lmax + past + i) * 64 + c * 8;
      *reinterpret_cast<uint4*>(dst) = r0;
      *reinterpret_cast<uint4*>(dst + 32) = r1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Embedding gather + <IMG_CONTEXT> / <TARGET_POINT> substitution.  One block per batch row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
embed_assemble_kernel(const long long* __restrict__ ids, const bf16* __restrict__ table, const bf16* __restrict__ vit,
                      const bf16* __restrict__ wp, const int* __restrict__ wp_start, int wp_len, bf16* __restrict__ out,
                      int len, int hidden, int vocab, int img_id, int n_img) {
  extern __shared__ const bf16* src_ptr[];
  __shared__ int warp_counts[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int running = 0;
  for (int base = 0; base < len; base += 1024) {
    const int l = base + tid;
    long long id = (l < len) ? ids[(size_t)b * len + l] : -1;
    const bool is_img = (l < len) && (id == img_id);
    const unsigned bal = __ballot_sync(0xffffffffu, is_img);
    if (lane == 0) warp_counts[warp] = __popc(bal);
    __syncthreads();
    int prefix = running;
    for (int wi = 0; wi < warp; ++wi) prefix += warp_counts[wi];
    const int rank = prefix + __popc(bal & ((1u << lane) - 1));
    if (l < len) {
      const bf16* s;
      const int ws = wp_start ? wp_start[b] : -1;
      if (wp && ws >= 0 && l >= ws && l < ws + wp_len) s = wp + ((size_t)b * wp_len + (l - ws)) * hidden;
      else if (is_img && vit && rank < n_img) s = vit + ((size_t)b * n_img + rank) * hidden;
      else {
        long long c = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        s = table + (size_t)c * hidden;
      }
      src_ptr[l] = s;
    }
    int tot = 0;
    for (int wi = 0; wi < 32; ++wi) tot += warp_counts[wi];
    running += tot;
    __syncthreads();
  }
  const int vpr = hidden >> 3;
  for (int idx = tid; idx < len * vpr; idx += blockDim.x) {
    const int l = idx / vpr, vi = idx % vpr;
    *reinterpret_cast<uint4*>(out + ((size_t)b * len + l) * hidden + vi * 8) =
        *reinterpret_cast<const uint4*>(src_ptr[l] + vi * 8);
  }
}

__global__ void gather_rows_kernel(const bf16* __restrict__ src, const long long* __restrict__ idx, bf16* __restrict__ dst,
                                   int n, int cols, long long src_rows) {
  const int vpr = cols >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * vpr; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vpr), vi = (int)(i % vpr);
    long long s = idx[r];
    s = s < 0 ? 0 : (s >= src_rows ? src_rows - 1 : s);
    *reinterpret_cast<uint4*>(dst + (size_t)r * cols + vi * 8) = *reinterpret_cast<const uint4*>(src + (size_t)s * cols + vi * 8);
  }
}

__global__ void scatter_rows_kernel(bf16* __restrict__ dst, const long long* __restrict__ idx, const bf16* __restrict__ src, int n,
                                    int cols, long long dst_rows) {
  const int vpr = cols >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * vpr; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vpr), vi = (int)(i % vpr);
    const long long d = idx[r];
    if (d < 0 || d >= dst_rows) continue;
    *reinterpret_cast<uint4*>(dst + (size_t)d * cols + vi * 8) = *reinterpret_cast<const uint4*>(src + (size_t)r * cols + vi * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// simple elementwise
// ------------------------------------------------------------------------------------------------
__global__ void silu_mul_kernel(const bf16* __restrict__ g, const bf16* __restrict__ u, bf16* __restrict__ o, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8], r[8];
    load8(g + i * 8, a);
    load8(u + i * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = silu(a[e]) * b[e];
    store8(o + i * 8, r);
  }
}
__global__ void add_kernel(const bf16* __restrict__ a_, const bf16* __restrict__ b_, bf16* __restrict__ o, size_t nvec) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8], r[8];
    load8(a_ + i * 8, a);
    load8(b_ + i * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = a[e] + b[e];
    store8(o + i * 8, r);
  }
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}

// ------------------------------------------------------------------------------------------------
// argmax over fp32 logits: one block per row; ties -> lowest index; also top1 - top2 margin
// ------------------------------------------------------------------------------------------------
struct Top2 { float v1; int i1; float v2; };
__device__ __forceinline__ void top2_push(Top2& t, float v, int i) {
  if (v > t.v1 || (v == t.v1 && i < t.i1)) { t.v2 = t.v1; t.v1 = v; t.i1 = i; }
  else if (v > t.v2) t.v2 = v;
}
__device__ __forceinline__ Top2 top2_merge(Top2 a, const Top2& b) {
  top2_push(a, b.v1, b.i1);
  if (b.v2 > a.v2) a.v2 = b.v2;
  return a;
}
// optional bookkeeping of the greedy loop (llm.py:232-248) by the thread that holds the row's result: column `*pos - base` of the
// row's sampled tokens, generated-token count, finished flag - instead of ~8 dependent element-wise torch kernels per token
struct SampleState {
  long long* sampled; long long ld; const int* pos; int base, max_new;
  unsigned char* done; long long* n_gen; long long* step; long long eos;
};
__global__ void __launch_bounds__(1024)
argmax_kernel(const float* __restrict__ logits, long long ld, int cols, long long* __restrict__ out_idx, float* __restrict__ margin, SampleState st) {
  __shared__ Top2 sh[32];
  const float* r = logits + (size_t)blockIdx.x * ld;
  pdl_trigger();
  pdl_wait();
  Top2 t{-INFINITY, 0x7fffffff, -INFINITY};
  // 16-byte loads, four in flight per thread (the scalar loop was one dependent 4-byte round trip per 1024 logits: 68 us for the
  // 151 655-wide vocabulary row of a decode step); scalar head / tail around the aligned body
  const int mis = (int)((reinterpret_cast<uintptr_t>(r) & 15) >> 2);
  const int head = min(cols, mis ? 4 - mis : 0);
  const int nvec = (cols - head) >> 2;
  if ((int)threadIdx.x < head) top2_push(t, r[threadIdx.x], threadIdx.x);
  const float4* rv = reinterpret_cast<const float4*>(r + head);
  for (int v0 = 0; v0 < nvec; v0 += 4 * blockDim.x) {
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u * blockDim.x + threadIdx.x;
      q[u] = v < nvec ? rv[v] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u * (int)blockDim.x + (int)threadIdx.x;
      if (v >= nvec) continue;
      const int i0 = head + 4 * v;
      top2_push(t, q[u].x, i0); top2_push(t, q[u].y, i0 + 1); top2_push(t, q[u].z, i0 + 2); top2_push(t, q[u].w, i0 + 3);
    }
  }
  for (int i = head + 4 * nvec + threadIdx.x; i < cols; i += blockDim.x) top2_push(t, r[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Top2 u;
    u.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
    u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    u.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
    t = top2_merge(t, u);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    t = threadIdx.x < nw ? sh[threadIdx.x] : Top2{-INFINITY, 0x7fffffff, -INFINITY};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Top2 u;
      u.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
      u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
      u.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
      t = top2_merge(t, u);
    }
    if (threadIdx.x == 0) {
      out_idx[blockIdx.x] = t.i1;
      if (margin) margin[blockIdx.x] = t.v1 - t.v2;
      if (st.sampled) {
        const int row = blockIdx.x, col = *st.pos - st.base;
        const unsigned char d = st.done[row];
        if (!d && col >= 0 && col < st.max_new) st.sampled[(size_t)row * st.ld + col] = t.i1;   // torch.where(done, cur, nxt)
        st.n_gen[row] += d ? 0 : 1;
        if (st.eos >= 0 && (long long)t.i1 == st.eos) st.done[row] = 1;
        if (row == 0) *st.step = col + 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tiny fused MLPs (latency-bound): warp-per-output dot products over smem-resident activations
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot_bf16(const bf16* __restrict__ wrow, const float* __restrict__ x, int n, int lane) {
  float acc = 0.f;
  const int nvec = n >> 3;
  for (int vi = lane; vi < nvec; vi += 32) {
    float wv[8];
    load8(wrow + vi * 8, wv);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += wv[e] * x[vi * 8 + e];
  }
  return warp_sum(acc);
}
// y[j] = act(b[j] + W[j,:] . x) for j in [0, nout); all warps of the block cooperate; x, y in smem.  Each warp works on
// four output rows at a time (four independent load streams) so the chain of dependent global-load latencies is
// nout / (4 * warps) long instead of nout / warps.
__device__ __forceinline__ void block_linear(const bf16* W, const bf16* bias, const float* x, float* y, int nin, int nout,
                                             int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int nvec = nin >> 3;
  for (int j0 = warp * 4; j0 < nout; j0 += nw * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int vi = lane; vi < nvec; vi += 32) {
      float wv[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r) load8(W + (size_t)min(j0 + r, nout - 1) * nin + vi * 8, wv[r]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = x[vi * 8 + e];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] += wv[r][e] * xv;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < 4 && j0 + lane < nout) {
      float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      const int j = j0 + lane;
      if (bias) v += __bfloat162float(bias[j]);
      if (act == SLB_ACT_SILU) v = silu(v);
      else if (act == SLB_ACT_RELU) v = fmaxf(v, 0.f);
      y[j] = v;
    }
  }
  __syncthreads();
}

// one block per (batch, query row): rows 0..19 -> route head, 20..29 -> speed head; writes pre-cumsum deltas
__global__ void __launch_bounds__(1024)
heads_kernel(const bf16* __restrict__ feats, long long ld_batch, slb_heads_weights w, float* __restrict__ delta) {
  __shared__ float x[896];
  __shared__ float h1[512];
  __shared__ float h2[256];
  const int b = blockIdx.x / 30, r = blockIdx.x % 30;
  const bf16* f = feats + (size_t)b * ld_batch + (size_t)r * 896;
  for (int i = threadIdx.x; i < 896; i += blockDim.x) x[i] = __bfloat162float(f[i]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (r < 20) {
    block_linear((const bf16*)w.r0w, (const bf16*)w.r0b, x, h1, 896, 512, SLB_ACT_SILU);
    block_linear((const bf16*)w.r2w, (const bf16*)w.r2b, h1, h2, 512, 256, SLB_ACT_SILU);
    if (warp < 2) {
      float v = warp_dot_bf16((const bf16*)w.r4w + warp * 256, h2, 256, lane);
      if (lane == 0) delta[((size_t)b * 30 + r) * 2 + warp] = v;
    }
  } else {
    block_linear((const bf16*)w.s0w, (const bf16*)w.s0b, x, h2, 896, 256, SLB_ACT_SILU);
    if (warp < 2) {
      float v = warp_dot_bf16((const bf16*)w.s2w + warp * 256, h2, 256, lane);
      if (lane == 0) delta[((size_t)b * 30 + r) * 2 + warp] = v;
    }
  }
}
__global__ void heads_cumsum_kernel(const float* __restrict__ delta, float* __restrict__ route, float* __restrict__ speed, int batch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (b, xy)
  if (idx >= batch * 2) return;
  const int b = idx >> 1, c = idx & 1;
  float s = 0.f;
  for (int r = 0; r < 20; ++r) { s += delta[((size_t)b * 30 + r) * 2 + c]; route[((size_t)b * 20 + r) * 2 + c] = s; }
  s = 0.f;
  for (int r = 0; r < 10; ++r) { s += delta[((size_t)b * 30 + 20 + r) * 2 + c]; speed[((size_t)b * 10 + r) * 2 + c] = s; }
}

// wp encoder: 2 -> 256 -> 512 -> 896 (ReLU); one block per point
__global__ void __launch_bounds__(1024)
wp_encoder_kernel(const float* __restrict__ coords, slb_wp_weights w, bf16* __restrict__ out) {
  __shared__ float h1[256];
  __shared__ float h2[512];
  __shared__ float h3[896];
  const int p = blockIdx.x;
  const float cx = coords[p * 2], cy = coords[p * 2 + 1];
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    const bf16* wr = (const bf16*)w.w0 + j * 2;
    float v = __bfloat162float(wr[0]) * cx + __bfloat162float(wr[1]) * cy + __bfloat162float(((const bf16*)w.b0)[j]);
    h1[j] = fmaxf(v, 0.f);
  }
  __syncthreads();
  block_linear((const bf16*)w.w2, (const bf16*)w.b2, h1, h2, 256, 512, SLB_ACT_RELU);
  block_linear((const bf16*)w.w4, (const bf16*)w.b4, h2, h3, 512, 896, SLB_ACT_NONE);
  for (int j = threadIdx.x; j < 896; j += blockDim.x) out[(size_t)p * 896 + j] = __float2bfloat16(h3[j]);
}

inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  size_t cap = (size_t)slb_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" int slb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, int rows, int cols, float eps,
                                 float* mean, float* rstd, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 8) == 0 && cols <= 4096, "layernorm: bad shape %d x %d", rows, cols);
  const int grid = ceil_div(rows, kWarpsPerBlock);
  if (cols <= 1024)
    SLB_CUDA(slb_launch_pdl(rows <= 4096, norm_fwd_kernel<4, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, ST(stream), (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, rows, cols, eps, mean, rstd));
  else
    SLB_CUDA(slb_launch_pdl(rows <= 4096, norm_fwd_kernel<16, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, ST(stream), (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, rows, cols, eps, mean, rstd));
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_layernorm_fwd_f32(const float* x, const void* w, const void* b, void* y, int rows, int cols, float eps, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 8) == 0 && cols <= 1024, "layernorm_f32: bad shape %d x %d", rows, cols);
  SLB_CHECK_ARG(x && w && b && y && (((uintptr_t)x) & 15) == 0, "layernorm_f32: null / unaligned operand");
  const int grid = ceil_div(rows, kWarpsPerBlock);
  SLB_CUDA(slb_launch_pdl(rows <= 4096, norm_fwd_kernel<4, false, float>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, ST(stream), x, (const bf16*)w, (const bf16*)b, (bf16*)y, rows, cols, eps, nullptr, nullptr));
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_rmsnorm_fwd_f32(const float* x, const void* w, void* y, int rows, int cols, float eps, float* rstd, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 8) == 0 && cols <= 1024, "rmsnorm_f32: bad shape %d x %d", rows, cols);
  SLB_CHECK_ARG(x && w && y && (((uintptr_t)x) & 15) == 0, "rmsnorm_f32: null / unaligned operand");
  const int grid = ceil_div(rows, kWarpsPerBlock);
  SLB_CUDA(slb_launch_pdl(rows <= 4096, norm_fwd_kernel<4, true, float>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, ST(stream), x, (const bf16*)w, nullptr, (bf16*)y, rows, cols, eps, nullptr, rstd));
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_rmsnorm_fwd(const void* x, const void* w, void* y, int rows, int cols, float eps, float* rstd, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && (cols % 8) == 0 && cols <= 1024, "rmsnorm: bad shape %d x %d", rows, cols);
  const int grid = ceil_div(rows, kWarpsPerBlock);
  SLB_CUDA(slb_launch_pdl(rows <= 4096, norm_fwd_kernel<4, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, ST(stream), (const bf16*)x, (const bf16*)w, nullptr, (bf16*)y, rows, cols, eps, nullptr, rstd));
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_pixel_shuffle_ln(const void* x, const void* w, const void* b, void* y, int tiles, float eps, float* mean,
                                    float* rstd, void* stream) {
  SLB_CHECK_ARG(tiles > 0, "pixel_shuffle_ln: tiles=%d", tiles);
  const int rows = tiles * 256;
  pixel_shuffle_ln_kernel<bf16><<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, ST(stream)>>>(
      (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, tiles, eps, mean, rstd);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

// fp32 residual stream of the InternViT inference path (see slb_layernorm_fwd_f32)
extern "C" int slb_pixel_shuffle_ln_f32(const float* x, const void* w, const void* b, void* y, int tiles, float eps, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && x && w && b && y, "pixel_shuffle_ln_f32: bad args");
  const int rows = tiles * 256;
  pixel_shuffle_ln_kernel<float><<<ceil_div(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, ST(stream)>>>(
      x, (const bf16*)w, (const bf16*)b, (bf16*)y, tiles, eps, nullptr, nullptr);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_im2col_patch(const void* pixels, void* patches, int tiles, int kpad, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && kpad >= 588 && (kpad % 8) == 0, "im2col: tiles=%d kpad=%d", tiles, kpad);
  const size_t total = (size_t)tiles * 1024 * (kpad / 2);
  im2col_patch_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)pixels, (bf16*)patches, tiles, kpad);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_vit_assemble(const void* patch_out, const void* cls, const void* pos, void* x, int tiles, void* stream) {
  SLB_CHECK_ARG(tiles > 0, "vit_assemble: tiles=%d", tiles);
  const size_t total = (size_t)tiles * 1025 * 128;
  vit_assemble_kernel<bf16><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)patch_out, (const bf16*)cls, (const bf16*)pos, (bf16*)x, tiles);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_vit_assemble_f32(const void* patch_out, const void* cls, const void* pos, float* x, int tiles, void* stream) {
  SLB_CHECK_ARG(tiles > 0 && patch_out && cls && pos && x, "vit_assemble_f32: bad args");
  const size_t total = (size_t)tiles * 1025 * 128;
  vit_assemble_kernel<float><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)patch_out, (const bf16*)cls, (const bf16*)pos, x, tiles);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_rope_kv_write(void* qkv, void* kcache, void* vcache, int batch, int lq, int past, const int32_t* past_dev, int lmax,
                                 int hq, int hkv, float theta, void* stream) {
  SLB_CHECK_ARG(batch > 0 && lq > 0 && past >= 0 && past + lq <= lmax, "rope: batch=%d lq=%d past=%d lmax=%d", batch, lq, past, lmax);
  const int per_head = (size_t)batch * lq * 4 < 4096 ? 1 : 0;
  const size_t total = (size_t)batch * lq * 4 * (per_head ? hq + 2 * hkv : 1);
  SLB_CUDA(slb_launch_pdl(per_head != 0 || (size_t)batch * lq <= 4096, rope_kv_write_kernel, dim3(grid_for(total, 128)), dim3(128), 0, ST(stream), (bf16*)qkv, (bf16*)kcache, (bf16*)vcache, batch, lq,
                          past, lmax, hq, hkv, log2f(theta), (const int*)past_dev, per_head));
  return SLB_OK;
}

extern "C" int slb_embed_assemble(const int64_t* ids, const void* table, const void* vit, const void* wp, const int32_t* wp_start,
                                  int wp_len, void* out, int batch, int len, int hidden, int vocab, int img_id, int n_img,
                                  void* stream) {
  SLB_CHECK_ARG(batch > 0 && len > 0 && (hidden % 8) == 0, "embed_assemble: batch=%d len=%d hidden=%d", batch, len, hidden);
  SLB_CHECK_ARG((size_t)len * sizeof(void*) <= 200 * 1024, "embed_assemble: len=%d too long", len);
  const size_t smem = (size_t)len * sizeof(void*);
  if (smem > 48 * 1024) SLB_CUDA(cudaFuncSetAttribute(embed_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  embed_assemble_kernel<<<batch, 1024, smem, ST(stream)>>>((const long long*)ids, (const bf16*)table, (const bf16*)vit, (const bf16*)wp,
                                                          wp_start, wp_len, (bf16*)out, len, hidden, vocab, img_id, n_img);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_gather_rows(const void* src, const int64_t* idx, void* dst, int n, int cols, int64_t src_rows, void* stream) {
  SLB_CHECK_ARG(n > 0 && (cols % 8) == 0, "gather_rows: n=%d cols=%d", n, cols);
  gather_rows_kernel<<<grid_for((size_t)n * (cols / 8), 128), 128, 0, ST(stream)>>>((const bf16*)src, (const long long*)idx, (bf16*)dst, n, cols, src_rows);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_scatter_rows(void* dst, const int64_t* idx, const void* src, int n, int cols, int64_t dst_rows, void* stream) {
  SLB_CHECK_ARG(n >= 0 && (cols % 8) == 0, "scatter_rows: n=%d cols=%d", n, cols);
  if (n == 0) return SLB_OK;
  scatter_rows_kernel<<<grid_for((size_t)n * (cols / 8), 128), 128, 0, ST(stream)>>>((bf16*)dst, (const long long*)idx, (const bf16*)src, n, cols, dst_rows);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_silu_mul(const void* gate, const void* up, void* out, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "silu_mul: n=%lld", (long long)n);
  silu_mul_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)gate, (const bf16*)up, (bf16*)out, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0, "add: n=%lld", (long long)n);
  add_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  SLB_CHECK_ARG(n > 0, "cast: n=%lld", (long long)n);
  cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, (bf16*)y, n);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_argmax_f32(const float* logits, int64_t ld, int rows, int cols, int64_t* out_idx, float* out_margin, void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0, "argmax: rows=%d cols=%d", rows, cols);
  SampleState st = {};
  SLB_CUDA(slb_launch_pdl(true, argmax_kernel, dim3(rows), dim3(1024), 0, ST(stream), logits, (long long)ld, cols, (long long*)out_idx, out_margin, st));
  return SLB_OK;
}

extern "C" int slb_argmax_sample(const float* logits, int64_t ld, int rows, int cols, int64_t* nxt, int64_t* sampled, int64_t ld_sampled,
                                 int max_new, const int32_t* pos, int base, uint8_t* done, int64_t* n_gen, int64_t* step, int64_t eos,
                                 void* stream) {
  SLB_CHECK_ARG(rows > 0 && cols > 0 && max_new > 0, "argmax_sample: rows=%d cols=%d max_new=%d", rows, cols, max_new);
  SLB_CHECK_ARG(logits && nxt && sampled && pos && done && n_gen && step, "argmax_sample: null state pointer");
  SampleState st;
  st.sampled = (long long*)sampled; st.ld = ld_sampled; st.pos = pos; st.base = base; st.max_new = max_new;
  st.done = done; st.n_gen = (long long*)n_gen; st.step = (long long*)step; st.eos = eos;
  SLB_CUDA(slb_launch_pdl(true, argmax_kernel, dim3(rows), dim3(1024), 0, ST(stream), logits, (long long)ld, cols, (long long*)nxt, (float*)nullptr, st));
  return SLB_OK;
}

extern "C" int slb_driving_heads(const void* feats, int64_t ld_batch, const slb_heads_weights* w, float* route, float* speed,
                                 float* delta_ws, int batch, void* stream) {
  SLB_CHECK_ARG(batch > 0 && w && delta_ws, "driving_heads: bad args");
  heads_kernel<<<batch * 30, 1024, 0, ST(stream)>>>((const bf16*)feats, ld_batch, *w, delta_ws);
  SLB_LAUNCH_CHECK();
  heads_cumsum_kernel<<<ceil_div(batch * 2, 64), 64, 0, ST(stream)>>>(delta_ws, route, speed, batch);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_wp_encoder(const float* coords, const slb_wp_weights* w, void* out, int n_points, void* stream) {
  SLB_CHECK_ARG(n_points > 0 && w, "wp_encoder: bad args");
  wp_encoder_kernel<<<n_points, 1024, 0, ST(stream)>>>(coords, *w, (bf16*)out);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
