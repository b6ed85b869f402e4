// Side kernels of the un-merged LoRA training path (PEFT lora.Linear: y = W x + b + (alpha/r) B A dropout(x), every adapter with
// its own dropout mask; reference llm.py:106-118 wraps all seven Qwen2 linears of a layer).
//
// The tensor-core work rides in the base GEMMs: forward  y = [x | t] [W | s B]^T  (slb_gemm_bf16 A2, t = A dropout(x)),
// backward  [dx | dt] = dy [W | s B]  (one dgrad over the concatenated weight).  What is left are HBM-bound passes:
//   dropout_multi   x -> n independently masked copies in one read of x (q/k/v and gate/up share their input)
//   lora_pack       scale * B of every adapter -> its column block of the concatenated weights (one launch per step)
//   lora_dx         dx_out = dx_base + sum_j mask_j o (dt_j A_j) / (1 - p): the rank-r products on the CUDA cores (0.8 GFLOP
//                   per call), fp32 accumulation over the adapters, one rounding - instead of one [M, K] GEMM + one
//                   read-modify-write pass per adapter
//   silu_mul_cat    SwiGLU forward / backward on the [gate | up] output of the fused gate-up GEMM
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int kMaxAdapters = 4;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
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

struct MultiArgs {
  bf16* y[kMaxAdapters];
  uint64_t seed[kMaxAdapters];
  int n;
};
__global__ void dropout_multi_kernel(const bf16* __restrict__ x, MultiArgs a, size_t nvec, uint32_t thresh, float scale,
                                     const uint64_t* __restrict__ seed_dev) {
  const uint64_t add = seed_dev ? (*seed_dev << 16) : 0;
  uint64_t sm[kMaxAdapters];
#pragma unroll
  for (int j = 0; j < kMaxAdapters; ++j) sm[j] = (a.seed[j] + add) * kDropGold;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < kMaxAdapters; ++j) {
      if (j < a.n) {
        float o[8];
        const uint64_t h0 = drop_hash4(sm[j], 2 * i), h1 = drop_hash4(sm[j], 2 * i + 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = drop_keep(e < 4 ? h0 : h1, e & 3, thresh) ? v[e] * scale : 0.f;
        store8(a.y[j] + i * 8, o);
      }
    }
  }
}

struct PackEntry { long long src, dst, rows, ld_dst; };   // src: bf16 [rows, r] contiguous; dst: bf16, row stride ld_dst
__global__ void lora_pack_kernel(const PackEntry* __restrict__ tab, int r8, float scale) {
  const PackEntry e = tab[blockIdx.x];
  const bf16* src = reinterpret_cast<const bf16*>(e.src);
  bf16* dst = reinterpret_cast<bf16*>(e.dst);
  const long long chunks = e.rows * r8;
  for (long long q = blockIdx.y * (long long)blockDim.x + threadIdx.x; q < chunks; q += (long long)gridDim.y * blockDim.x) {
    const long long row = q / r8;
    const int c = (int)(q % r8);
    float v[8];
    load8(src + q * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= scale;
    store8(dst + row * e.ld_dst + c * 8, v);
  }
}

// ---- lora_dx ------------------------------------------------------------------------------------------------------------
// Block = 64 rows x 128 columns, 8 warps; warp w owns rows 16 (w & 3) .., columns 64 (w >> 2) ..  The rank-r products run on
// mma.sync.m16n8k16 (bf16, fp32 accumulators): dt tile [64, r] and A_j tile [r, 128] staged in padded shared memory (conflict-free
// 32-bit A-fragment loads / ldmatrix.trans B fragments), one pass per adapter, mask applied on the accumulator fragment, fp32 sum
// over the adapters on top of the base dgrad, one rounding.  (A first CUDA-core version was LSU / FMA bound at 51 us per call.)
constexpr int DX_ROWS = 64, DX_COLS = 128, DX_THREADS = 256, DX_RMAX = 64;   // the kernel takes rank <= DX_RMAX / 2 (double-buffered operands)
constexpr int DX_LDA = DX_COLS + 8;   // bf16 elements per sA row (272 B: 8 consecutive rows hit 32 distinct banks)
struct LoraDxArgs {
  const bf16* in; long long ld_in;   // [M, >= K + r*n]: base dgrad in columns [0, K), dt_j in columns [K + r*j, K + r*(j+1))
  bf16* out; long long ld_out;       // [M, K]
  const bf16* A[kMaxAdapters];       // [r, K] each, contiguous
  uint64_t seed[kMaxAdapters];
  int n, M, K, r;
  uint32_t thresh;
  float scale;
  int use_mask;
};
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {   // 16-byte async copy, zero-fill when !pred
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(pred ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// All global reads are asynchronous copies: the base tile and adapter 0's operands leave together, adapter j + 1 is in flight while
// adapter j is multiplied (double-buffered sA / sD) - the block's dependent global round trips drop from n + 2 to 2 (the kernel is
// latency-bound: 2 blocks per SM, a few microseconds of work each).
__global__ void __launch_bounds__(DX_THREADS)
lora_dx_kernel(LoraDxArgs a, const uint64_t* __restrict__ seed_dev) {
  __shared__ __align__(16) bf16 sA[2][DX_RMAX / 2 * DX_LDA];      // [r][128 + 8], r <= 32, double-buffered over the adapters
  __shared__ __align__(16) bf16 sD[2][DX_ROWS * (DX_RMAX / 2 + 8)];  // [64][r + 8]
  __shared__ __align__(16) bf16 sT[DX_ROWS * DX_LDA];            // base dgrad tile in, result tile out
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int col_tile = blockIdx.x * DX_COLS, row_tile = blockIdx.y * DX_ROWS;
  const int wr = (warp & 3) * 16, wc = (warp >> 2) * 64;
  const int r = a.r, r8 = r >> 3, ldd = r + 8;
  const uint64_t add = seed_dev ? (*seed_dev << 16) : 0;
  const int row0 = row_tile + wr + g, row1 = row0 + 8;
  auto stage = [&](int j, int buf) {
    for (int q = tid; q < r * (DX_COLS / 8); q += DX_THREADS) {
      const int k = q >> 4, cc = q & 15;
      const int c = col_tile + cc * 8;
      const bool ok = c < a.K;
      cp_async16(sA[buf] + k * DX_LDA + cc * 8, a.A[j] + (ok ? (long long)k * a.K + c : 0), ok);
    }
    for (int q = tid; q < DX_ROWS * r8; q += DX_THREADS) {
      const int rw = q / r8, cc = q % r8;
      const int row = row_tile + rw;
      const bool ok = row < a.M;
      cp_async16(sD[buf] + rw * ldd + cc * 8, a.in + (ok ? (long long)row * a.ld_in + a.K + j * r + cc * 8 : 0), ok);
    }
  };
  for (int q = tid; q < DX_ROWS * (DX_COLS / 8); q += DX_THREADS) {
    const int rw = q >> 4, cc = q & 15;
    const int row = row_tile + rw, c = col_tile + cc * 8;
    const bool ok = row < a.M && c < a.K;
    cp_async16(sT + rw * DX_LDA + cc * 8, a.in + (ok ? (long long)row * a.ld_in + c : 0), ok);
  }
  stage(0, 0);
  cp_async_commit();
  float tot[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) tot[nt][e] = 0.f;
  for (int j = 0; j < a.n; ++j) {
    const int buf = j & 1;
    if (j + 1 < a.n) stage(j + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait_1();     // everything but the group just committed has landed: the base tile and adapter j
    __syncthreads();
    const bf16* sa = sA[buf];
    const bf16* sd = sD[buf];
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    for (int k0 = 0; k0 < r; k0 += 16) {
      const bf16* d0 = sd + (wr + g) * ldd + k0 + 2 * t;
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(d0), a1 = *reinterpret_cast<const uint32_t*>(d0 + 8 * ldd);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(d0 + 8), a3 = *reinterpret_cast<const uint32_t*>(d0 + 8 * ldd + 8);
      // ldmatrix.x4.trans: lanes 0-7 address rows k0..k0+7 of n-tile nt, 8-15 rows k0+8.., 16-23 / 24-31 the same for n-tile nt + 1
      const uint32_t base = smem_u32(sa + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * DX_LDA + wc + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < 8; nt += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(base + nt * 8 * 2, b0, b1, b2, b3);
        mma_16816(acc[nt], a0, a1, a2, a3, b0, b1);
        mma_16816(acc[nt + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    if (a.use_mask) {
      const uint64_t sm = (a.seed[j] + add) * kDropGold;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = col_tile + wc + nt * 8 + 2 * t;    // elements c, c + 1 of rows row0 / row1; (c & 3) is 0 or 2
        const uint64_t i0 = (uint64_t)row0 * (uint64_t)a.K + (uint64_t)c, i1 = (uint64_t)row1 * (uint64_t)a.K + (uint64_t)c;
        const uint64_t h0 = drop_hash4(sm, i0 >> 2), h1 = drop_hash4(sm, i1 >> 2);
        const int e = c & 3;
        tot[nt][0] += drop_keep(h0, e, a.thresh) ? acc[nt][0] * a.scale : 0.f;
        tot[nt][1] += drop_keep(h0, e + 1, a.thresh) ? acc[nt][1] * a.scale : 0.f;
        tot[nt][2] += drop_keep(h1, e, a.thresh) ? acc[nt][2] * a.scale : 0.f;
        tot[nt][3] += drop_keep(h1, e + 1, a.thresh) ? acc[nt][3] * a.scale : 0.f;
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) tot[nt][e] += acc[nt][e];
    }
    __syncthreads();       // buffer `buf` is refilled by the stage issued at the top of iteration j + 1
  }
  // result = base + sum over adapters (fp32), rounded once, back into the tile (each thread owns its fragment words), then
  // coalesced 16-byte stores
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t* w0 = reinterpret_cast<uint32_t*>(sT + (wr + g) * DX_LDA + wc + nt * 8 + 2 * t);
    uint32_t* w1 = reinterpret_cast<uint32_t*>(sT + (wr + g + 8) * DX_LDA + wc + nt * 8 + 2 * t);
    const float2 b0 = unpack_bf16(*w0), b1 = unpack_bf16(*w1);
    *w0 = pack_bf16(b0.x + tot[nt][0], b0.y + tot[nt][1]);
    *w1 = pack_bf16(b1.x + tot[nt][2], b1.y + tot[nt][3]);
  }
  __syncthreads();
  for (int q = tid; q < DX_ROWS * (DX_COLS / 8); q += DX_THREADS) {
    const int rw = q >> 4, cc = q & 15;
    const int row = row_tile + rw, c = col_tile + cc * 8;
    if (row < a.M && c < a.K) *reinterpret_cast<uint4*>(a.out + (long long)row * a.ld_out + c) = *reinterpret_cast<const uint4*>(sT + rw * DX_LDA + cc * 8);
  }
}

// ---- SwiGLU on the concatenated [gate | up] layout ----------------------------------------------------------------------
__global__ void silu_mul_cat_kernel(const bf16* __restrict__ gu, bf16* __restrict__ out, size_t rows, int i8) {
  const size_t nvec = rows * i8;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nvec; q += (size_t)gridDim.x * blockDim.x) {
    const size_t row = q / i8;
    const int c = (int)(q % i8);
    const bf16* g = gu + row * (size_t)(2 * i8 * 8) + c * 8;
    float a[8], b[8], o[8];
    load8(g, a);
    load8(g + i8 * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = silu(a[e]) * b[e];
    store8(out + q * 8, o);
  }
}
__global__ void silu_mul_cat_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dout, bf16* __restrict__ dgu, size_t rows, int i8) {
  const size_t nvec = rows * i8;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nvec; q += (size_t)gridDim.x * blockDim.x) {
    const size_t row = q / i8;
    const int c = (int)(q % i8);
    const size_t off = row * (size_t)(2 * i8 * 8) + c * 8;
    float a[8], b[8], d[8], rgv[8], ruv[8];
    load8(gu + off, a);
    load8(gu + off + i8 * 8, b);
    load8(dout + q * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = 1.0f / (1.0f + __expf(-a[e]));
      ruv[e] = d[e] * a[e] * sg;
      rgv[e] = d[e] * b[e] * sg * (1.0f + a[e] * (1.0f - sg));
    }
    store8(dgu + off, rgv);
    store8(dgu + off + i8 * 8, ruv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Grouped parameter-gradient products of the adapters:  out_j [mo_j x no_j] (+)= alpha_j P_j^T Q_j  over the `rows` token rows
// (dB_j = s dy_j^T t_j : [out_f x r],  dA_j = dt_j^T dropout_j(x) : [r x in_f]).  As 14 single-tile tcgen05 GEMMs per layer
// (K = batch * length = 4728 deep, 1-7 CTAs each, ~25 us each) they cost 9 ms of a 75 ms training step
// (profiles/r02_kernel_breakdown_train_v3.log); here every 32 x 32 output tile of every problem of a group is one CTA of one
// launch: 8 warps = 2 row halves x 4 k16 slices of a 64-row chunk, operands [k][m] / [k][n] staged by a 4-stage cp.async ring,
// both fragments by ldmatrix.trans, mma.sync.m16n8k16, the four k-slice partial tiles summed through shared memory.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kMaxWgProbs = 16;
constexpr int WG_THREADS = 256, WG_KC = 64, WG_STAGES = 4, WG_ROWB = 80;   // 80 B per staged row (32 bf16 + 16 B): conflict-free ldmatrix
struct WgProb { const bf16* P; const bf16* Q; bf16* out; long long ldp, ldq, ldo; int tiles_n, tile0; float alpha; int acc; };
struct WgArgs { WgProb pr[kMaxWgProbs]; int n, rows; };

__device__ __forceinline__ void wg_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void wg_ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void wg_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(WG_THREADS)
lora_wgrad_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(16) uint8_t wsm[];   // [WG_STAGES][2 operands][WG_KC rows][WG_ROWB]; afterwards the partial tiles
  int j = 0;
  while (j + 1 < a.n && (int)blockIdx.x >= a.pr[j + 1].tile0) ++j;
  const bf16* Pg; const bf16* Qg; bf16* og; long long ldp, ldq, ldo; float alpha; int acc_out;
  {
    const WgProb& pb = a.pr[j];
    const int tl = blockIdx.x - pb.tile0, tm = tl / pb.tiles_n, tn = tl - tm * pb.tiles_n;
    Pg = pb.P + tm * 32; Qg = pb.Q + tn * 32; ldp = pb.ldp; ldq = pb.ldq; ldo = pb.ldo;
    og = pb.out + (size_t)tm * 32 * ldo + tn * 32; alpha = pb.alpha; acc_out = pb.acc;
  }
  const int rows = a.rows, nchunks = (rows + WG_KC - 1) / WG_KC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto load_chunk = [&](int ch, int stage) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = tid + i * WG_THREADS, op = id >> 8, r = (id & 255) >> 2, c = id & 3;
      const int row = ch * WG_KC + r;
      uint8_t* dst = wsm + (size_t)((stage * 2 + op) * WG_KC + r) * WG_ROWB + c * 16;
      if (row < rows) wg_cp_async16(dst, (op ? Qg + (size_t)row * ldq : Pg + (size_t)row * ldp) + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);   // tail rows of the last chunk
    }
  };
#pragma unroll
  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < nchunks) load_chunk(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int ks = warp & 3, mt = warp >> 2;
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  const int mat = lane >> 3, r8 = lane & 7;
  const int a_off = (ks * 16 + (mat >> 1) * 8 + r8) * WG_ROWB + (mt * 16 + (mat & 1) * 8) * 2;   // A = P^T: (k0|k0+8, m0|m0+8) blocks
  const int b_off = (ks * 16 + (mat & 1) * 8 + r8) * WG_ROWB + ((mat >> 1) * 8) * 2;             // B = Q: (k0|k0+8, n0|n0+8) blocks
  for (int ch = 0; ch < nchunks; ++ch) {
    asm volatile("cp.async.wait_group %0;" ::"n"(WG_STAGES - 2) : "memory");
    __syncthreads();
    if (ch + WG_STAGES - 1 < nchunks) load_chunk(ch + WG_STAGES - 1, (ch + WG_STAGES - 1) % WG_STAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const uint8_t* Ps = wsm + (size_t)((ch % WG_STAGES) * 2) * WG_KC * WG_ROWB;
    const uint8_t* Qs = Ps + (size_t)WG_KC * WG_ROWB;
    uint32_t a0, a1, a2, a3;
    wg_ldsm_x4_trans(a0, a1, a2, a3, Ps + a_off);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      wg_ldsm_x4_trans(b0, b1, b2, b3, Qs + b_off + np * 32);
      wg_mma(acc[2 * np], a0, a1, a2, a3, b0, b1);
      wg_mma(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float* red = reinterpret_cast<float*>(wsm);   // [8 warps][16 rows][33]
  {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      red[(warp * 16 + g) * 33 + nt * 8 + 2 * q] = acc[nt][0];
      red[(warp * 16 + g) * 33 + nt * 8 + 2 * q + 1] = acc[nt][1];
      red[(warp * 16 + 8 + g) * 33 + nt * 8 + 2 * q] = acc[nt][2];
      red[(warp * 16 + 8 + g) * 33 + nt * 8 + 2 * q + 1] = acc[nt][3];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * WG_THREADS, m = idx >> 5, n = idx & 31;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) v += red[(((m >> 4) * 4 + k) * 16 + (m & 15)) * 33 + n];
    v *= alpha;
    bf16* o = og + (size_t)m * ldo + n;
    if (acc_out) v += __bfloat162float(*o);
    *o = __float2bfloat16(v);
  }
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" int slb_dropout_multi(const void* x, void* const* ys, const uint64_t* seeds, int n_out, int64_t n, float p,
                                 const uint64_t* seed_dev, void* stream) {
  SLB_CHECK_ARG(x && ys && seeds && n_out >= 1 && n_out <= kMaxAdapters, "dropout_multi: n_out=%d (1..%d)", n_out, kMaxAdapters);
  SLB_CHECK_ARG(n > 0 && (n % 8) == 0 && p >= 0.f && p < 1.f, "dropout_multi: n=%lld p=%f", (long long)n, p);
  MultiArgs a;
  for (int j = 0; j < kMaxAdapters; ++j) {
    a.y[j] = j < n_out ? (bf16*)ys[j] : nullptr;
    a.seed[j] = j < n_out ? seeds[j] : 0;
    SLB_CHECK_ARG(j >= n_out || (ys[j] && ((uintptr_t)ys[j] & 15) == 0), "dropout_multi: output %d null or not 16-byte aligned", j);
  }
  a.n = n_out;
  dropout_multi_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)x, a, n / 8, drop_thresh16(p), drop_scale(p), seed_dev);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_lora_pack(const int64_t* table_dev, int n_entries, int rank, float scale, void* stream) {
  SLB_CHECK_ARG(table_dev && n_entries > 0 && rank > 0 && (rank % 8) == 0, "lora_pack: n=%d rank=%d", n_entries, rank);
  lora_pack_kernel<<<dim3(n_entries, 8), 256, 0, ST(stream)>>>(reinterpret_cast<const PackEntry*>(table_dev), rank / 8, scale);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_lora_dx(const void* in, int64_t ld_in, void* out, int64_t ld_out, const void* const* A, const uint64_t* seeds,
                           int n_adapters, int M, int K, int rank, float p, const uint64_t* seed_dev, void* stream) {
  SLB_CHECK_ARG(in && out && A && n_adapters >= 1 && n_adapters <= kMaxAdapters, "lora_dx: n_adapters=%d (1..%d)", n_adapters, kMaxAdapters);
  SLB_CHECK_ARG(M > 0 && K > 0 && (K % 8) == 0 && rank > 0 && (rank % 16) == 0 && rank <= DX_RMAX / 2, "lora_dx: M=%d K=%d rank=%d (multiple of 16, <= %d)", M,
                K, rank, DX_RMAX / 2);
  SLB_CHECK_ARG((ld_in % 8) == 0 && (ld_out % 8) == 0 && ld_in >= K + (int64_t)rank * n_adapters && ld_out >= K &&
                ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "lora_dx: strides / alignment (ld_in=%lld ld_out=%lld)",
                (long long)ld_in, (long long)ld_out);
  SLB_CHECK_ARG(p >= 0.f && p < 1.f, "lora_dx: p=%f", p);
  LoraDxArgs a;
  a.in = (const bf16*)in; a.ld_in = ld_in; a.out = (bf16*)out; a.ld_out = ld_out;
  for (int j = 0; j < kMaxAdapters; ++j) {
    a.A[j] = j < n_adapters ? (const bf16*)A[j] : nullptr;
    a.seed[j] = (seeds && j < n_adapters) ? seeds[j] : 0;
    SLB_CHECK_ARG(j >= n_adapters || (A[j] && ((uintptr_t)A[j] & 15) == 0), "lora_dx: A[%d] null or not 16-byte aligned", j);
  }
  a.n = n_adapters; a.M = M; a.K = K; a.r = rank;
  a.use_mask = (seeds != nullptr && p > 0.f) ? 1 : 0;
  a.thresh = drop_thresh16(p);
  a.scale = drop_scale(p);
  lora_dx_kernel<<<dim3(ceil_div(K, DX_COLS), ceil_div(M, DX_ROWS)), DX_THREADS, 0, ST(stream)>>>(a, seed_dev);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_silu_mul_cat(const void* gate_up, void* out, int rows, int inter, void* stream) {
  SLB_CHECK_ARG(gate_up && out && rows > 0 && inter > 0 && (inter % 8) == 0, "silu_mul_cat: %d x %d", rows, inter);
  silu_mul_cat_kernel<<<grid_for((size_t)rows * (inter / 8), 256), 256, 0, ST(stream)>>>((const bf16*)gate_up, (bf16*)out, (size_t)rows, inter / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
extern "C" int slb_silu_mul_cat_bwd(const void* gate_up, const void* dout, void* dgate_up, int rows, int inter, void* stream) {
  SLB_CHECK_ARG(gate_up && dout && dgate_up && rows > 0 && inter > 0 && (inter % 8) == 0, "silu_mul_cat_bwd: %d x %d", rows, inter);
  silu_mul_cat_bwd_kernel<<<grid_for((size_t)rows * (inter / 8), 256), 256, 0, ST(stream)>>>((const bf16*)gate_up, (const bf16*)dout, (bf16*)dgate_up,
                                                                                             (size_t)rows, inter / 8);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_lora_wgrad_grouped(const slb_wgrad_problem* probs, int n, int rows, void* stream) {
  SLB_CHECK_ARG(probs && n >= 1 && n <= kMaxWgProbs && rows > 0, "lora_wgrad_grouped: 1..%d problems, rows > 0 (n=%d rows=%d)", kMaxWgProbs, n, rows);
  WgArgs a;
  a.n = n; a.rows = rows;
  int tiles = 0;
  for (int j = 0; j < n; ++j) {
    const slb_wgrad_problem& q = probs[j];
    SLB_CHECK_ARG(q.P && q.Q && q.out && q.mo > 0 && q.no > 0 && (q.mo % 32) == 0 && (q.no % 32) == 0,
                  "lora_wgrad_grouped: problem %d: output %d x %d must be whole 32 x 32 tiles", j, q.mo, q.no);
    SLB_CHECK_ARG((q.ldp % 8) == 0 && (q.ldq % 8) == 0 && q.ldp >= q.mo && q.ldq >= q.no && q.ldo >= q.no &&
                  (((uintptr_t)q.P | (uintptr_t)q.Q) & 15) == 0,
                  "lora_wgrad_grouped: problem %d: operand rows must be 16-byte aligned (ldp=%lld ldq=%lld)", j, (long long)q.ldp, (long long)q.ldq);
    WgProb& w = a.pr[j];
    w.P = (const bf16*)q.P; w.Q = (const bf16*)q.Q; w.out = (bf16*)q.out; w.ldp = q.ldp; w.ldq = q.ldq; w.ldo = q.ldo;
    w.tiles_n = q.no / 32; w.tile0 = tiles; w.alpha = q.alpha; w.acc = q.accumulate;
    tiles += (q.mo / 32) * (q.no / 32);
  }
  const size_t smem = (size_t)WG_STAGES * 2 * WG_KC * WG_ROWB;   // 40 KB (>= the 16.5 KB of partial tiles)
  lora_wgrad_kernel<<<tiles, WG_THREADS, smem, ST(stream)>>>(a);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
