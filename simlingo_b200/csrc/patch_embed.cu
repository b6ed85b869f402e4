// InternViT patch embedding as an IMPLICIT GEMM (UPSTREAM InternVisionEmbeddings: Conv2d(3, 1024, kernel 14, stride 14) on a
// 448 x 448 tile + class token + position embedding; SURVEY K1 / north_star "patch-embed conv as implicit GEMM").
//
// The 14 x 14 patches do not overlap, so the convolution is x[p, :] = W[1024, 588] . patch(p) with patch(p)[(c, dy, dx)] =
// pixel[c, 14 py + dy, 14 px + dx].  A row of that operand is 42 runs of 14 pixels (28 bytes): TMA cannot fetch it (global strides
// must be multiples of 16 bytes), and materialising it ([tiles * 1024, 640] bf16: slb_im2col_patch) costs a pass over 168 MB at 128
// tiles.  Here the CTA's 256 threads gather the 128 x 64 operand tile of every k-block straight from the pixel tensor into the
// 128B-swizzled K-major shared-memory layout tcgen05.mma reads (generic-proxy stores + fence.proxy.async, the same hand-off the
// attention kernels use for P), the weight tile arrives by TMA, and the accumulator (128 patches x 256 channels, fp32 in TMEM) gets
// bias + position embedding in the epilogue and lands directly in the token stream x[t * 1025 + 1 + p, :] (fp32 residual stream of
// the inference path, or bf16); the class-token rows are written by the first CTA of each tile.  Two operand stages: the gather
// of k-block i + 1 overlaps the MMAs of k-block i; two CTAs per SM overlap each other's epilogue.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int PE_BM = 128, PE_BN = 256, PE_BK = 64, PE_K = 588, PE_KPAD = 640, PE_KB = PE_KPAD / PE_BK;
constexpr int PE_THREADS = 256;
constexpr int PE_C = 1024, PE_IMG = 448, PE_P = 14, PE_TOK = 1025;
constexpr int kABytes = PE_BM * PE_BK * 2, kBBytes = PE_BN * PE_BK * 2;
constexpr int kSmA = 0, kSmB = 2 * kABytes, kSmBar = kSmB + 2 * kBBytes, kSmTotal = kSmBar + 128 + 1024;

template <typename OT>
__global__ void __launch_bounds__(PE_THREADS, 2)
patch_embed_kernel(const bf16* __restrict__ pixels, const __grid_constant__ CUtensorMap tmap_w, const bf16* __restrict__ bias,
                   const bf16* __restrict__ cls, const bf16* __restrict__ pos, OT* __restrict__ x, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + kSmBar);   // [2] weight tile landed
  uint64_t* mma_done = full_b + 2;                                 // [2] the MMAs that read stage s have retired
  uint64_t* acc_done = mma_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
  const int tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31;
  const int m0 = blockIdx.x * PE_BM, n0 = blockIdx.y * PE_BN;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < 2; ++i) { mbar_init(&full_b[i], 1); mbar_init(&mma_done[i], 1); }
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, PE_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // gather assignment: thread -> (patch row r of the tile, 32 consecutive k of the 64-wide k-block)
  const int r = tid >> 1, half = tid & 1;
  const int P = m0 + r;
  const int ti = P >> 10, pp = P & 1023, py = pp >> 5, px = pp & 31;
  const bf16* pix = pixels + ((size_t)ti * 3 * PE_IMG + (size_t)py * PE_P) * PE_IMG + (size_t)px * PE_P;   // (c = 0, dy = 0, dx = 0) of this patch
  constexpr uint32_t idesc = umma_idesc_bf16(PE_BM, PE_BN, 0, 0);

  for (int kb = 0; kb < PE_KB; ++kb) {
    const int s = kb & 1, u = kb >> 1;
    if (kb >= 2) mbar_wait(&mma_done[s], (u - 1) & 1);   // stage s is free again
    if (tid == 0) {
      mbar_expect_tx(&full_b[s], kBBytes);
      tma_load_2d(smem + kSmB + s * kBBytes, &tmap_w, &full_b[s], kb * PE_BK, n0);
    }
    {
      uint8_t* arow = smem + kSmA + s * kABytes + r * 128;
      int k = kb * PE_BK + half * 32;
      int c = k / 196, rem = k - c * 196;
      int dy = rem / PE_P, dx = rem - dy * PE_P;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          unsigned short v = 0;
          if (k < PE_K) v = *reinterpret_cast<const unsigned short*>(pix + ((size_t)c * PE_IMG + dy) * PE_IMG + dx);
          if (e & 1) w[e >> 1] |= (uint32_t)v << 16; else w[e >> 1] = v;
          ++k;
          if (++dx == PE_P) { dx = 0; if (++dy == PE_P) { dy = 0; ++c; } }
        }
        const int chunk = half * 4 + j;
        *reinterpret_cast<uint4*>(arow + ((chunk ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      mbar_wait(&full_b[s], u & 1);
      tc_fence_after();
      const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem + kSmA + s * kABytes));
      const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smem + kSmB + s * kBBytes));
      if (elect_one_sync()) {
#pragma unroll
        for (int k16 = 0; k16 < PE_BK / 16; ++k16) tc_mma_bf16(tmem_base, da + 2 * k16, db + 2 * k16, idesc, (kb | k16) != 0);
        tc_commit(&mma_done[s]);
        if (kb == PE_KB - 1) tc_commit(acc_done);
      }
      __syncwarp();
    }
  }

  // class-token rows: the first CTA of every 448 x 448 tile (8 row tiles per image) writes x[ti * 1025, n0 .. n0 + 255]
  if ((blockIdx.x & 7) == 0) {
    const int n = n0 + tid;
    const float v = __bfloat162float(cls[n]) + __bfloat162float(pos[n]);
    OT* dst = x + (size_t)(m0 >> 10) * PE_TOK * PE_C + n;
    if constexpr (sizeof(OT) == 4) *dst = v; else *dst = __float2bfloat16(v);
  }

  mbar_wait(acc_done, 0);
  tc_fence_after();
  {
    const int quad = warp & 3, ch = warp >> 2;            // TMEM lane quadrant, 128-column half
    const int row = quad * 32 + lane;
    const int Pe = m0 + row;
    const int te = Pe >> 10, pe = Pe & 1023;
    OT* xrow = x + ((size_t)te * PE_TOK + 1 + pe) * PE_C + n0 + ch * 128;
    const bf16* prow = pos + (size_t)(1 + pe) * PE_C + n0 + ch * 128;
    const bf16* brow = bias + n0 + ch * 128;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ch * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t acc[32];
      tmem_ld_32x32(taddr + c, acc);
      uint4 pq[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pq[i] = __ldg(reinterpret_cast<const uint4*>(prow + c) + i);
        bq[i] = __ldg(reinterpret_cast<const uint4*>(brow + c) + i);
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 pf = unpack_bf16(reinterpret_cast<const uint32_t*>(pq)[i]), bf = unpack_bf16(reinterpret_cast<const uint32_t*>(bq)[i]);
        v[2 * i] = __uint_as_float(acc[2 * i]) + bf.x + pf.x;
        v[2 * i + 1] = __uint_as_float(acc[2 * i + 1]) + bf.y + pf.y;
      }
      if (Pe < tiles * 1024) {
        if constexpr (sizeof(OT) == 4) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(xrow + c)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 o;
            o.x = pack_bf16(v[8 * i], v[8 * i + 1]); o.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
            o.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]); o.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
            reinterpret_cast<uint4*>(xrow + c)[i] = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, PE_BN);
  }
}

}  // namespace

extern "C" int slb_patch_embed(const void* pixels, const void* weight, int64_t ldw, const void* bias, const void* cls, const void* pos,
                               void* x, int tiles, int out_fp32, void* stream) {
  SLB_CHECK_ARG(pixels && weight && bias && cls && pos && x && tiles > 0, "patch_embed: bad args (tiles=%d)", tiles);
  SLB_CHECK_ARG(ldw >= PE_KPAD && (ldw % 8) == 0 && ((uintptr_t)weight & 15) == 0, "patch_embed: weight must be [1024, >= 640] (zero-padded k), 16-byte aligned rows");
  SLB_CHECK_ARG((((uintptr_t)bias | (uintptr_t)pos | (uintptr_t)x) & 15) == 0 && ((uintptr_t)pixels & 1) == 0, "patch_embed: operands must be 16-byte aligned");
  CUtensorMap tw;
  int rc = slb_make_tmap_2d(&tw, weight, (uint64_t)PE_KPAD, (uint64_t)PE_C, (uint64_t)ldw * 2, PE_BK, PE_BN);
  if (rc) return rc;
  const dim3 grid(tiles * 1024 / PE_BM, PE_C / PE_BN);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_fp32) {
    static bool set = false;
    if (!set) { SLB_CUDA(cudaFuncSetAttribute(patch_embed_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal)); set = true; }
    patch_embed_kernel<float><<<grid, PE_THREADS, kSmTotal, st>>>((const bf16*)pixels, tw, (const bf16*)bias, (const bf16*)cls, (const bf16*)pos, (float*)x, tiles);
  } else {
    static bool set = false;
    if (!set) { SLB_CUDA(cudaFuncSetAttribute(patch_embed_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmTotal)); set = true; }
    patch_embed_kernel<bf16><<<grid, PE_THREADS, kSmTotal, st>>>((const bf16*)pixels, tw, (const bf16*)bias, (const bf16*)cls, (const bf16*)pos, (bf16*)x, tiles);
  }
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
