// Camera-frame pre-processing in front of the hot path (reference
// simlingo_training/utils/internvl2_utils.py:179-267 and team_code/agent_simlingo.py:483-502):
//   uint8 RGB [3, H, W]  --PIL bicubic resize-->  [3, 448*gh, 448*gw]  --crop-->  gw*gh tiles of 448x448
//   --ToTensor, Normalize(ImageNet)-->  bf16 [tiles, 3, 448, 448]   (what InternViT's patch embedding consumes)
// Bit-compatible with Pillow's two-pass 8-bit resampler (src/libImaging/Resample.c): horizontal pass into a uint8
// intermediate, then vertical; fixed-point taps (22 fractional bits) computed on the host exactly as
// precompute_coeffs / normalize_coeffs_8bpc do; ss = 2^21 + sum pix * k; clip8(ss >> 22).  The vertical pass fuses the
// tile split and the fp32 normalisation (same operation order as torch: (u8 / 255 - mean) / std), rounding once to bf16.
// HBM-bound and tiny (1.1 MB in, 2.4 MB out per frame); replaces ~10 ms of PIL + torchvision on the host per tick.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int ss) {
  const int v = ss >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// tmp[b, c, y, xx] = resample of row (b, c, y) at output column xx
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ x0,
                                  const int* __restrict__ cnt, const int* __restrict__ kk, int ks, int rows, int W, int out_w) {
  const size_t total = (size_t)rows * out_w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % out_w);
    const size_t row = idx / out_w;
    const uint8_t* s = src + row * W + x0[xx];
    const int* k = kk + (size_t)xx * ks;
    int ss = 1 << (kPrecisionBits - 1);
    const int n = cnt[xx];
    for (int x = 0; x < n; ++x) ss += (int)s[x] * k[x];
    dst[idx] = (uint8_t)clip8(ss);
  }
}

// out[b, tile, c, ty, tx] = normalise(resample of column (b, c, :, xx) at output row yy)
__global__ void resample_v_tiles_kernel(const uint8_t* __restrict__ src, bf16* __restrict__ out, const int* __restrict__ y0,
                                        const int* __restrict__ cnt, const int* __restrict__ kk, int ks, int batch, int H, int out_w,
                                        int out_h, int gw, float m0, float m1, float m2, float s0, float s1, float s2) {
  const size_t total = (size_t)batch * 3 * out_h * out_w;
  const int tiles = gw * (out_h / 448);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % out_w);
    const int yy = (int)((idx / out_w) % out_h);
    const int c = (int)((idx / ((size_t)out_w * out_h)) % 3);
    const int b = (int)(idx / ((size_t)out_w * out_h * 3));
    const uint8_t* s = src + (((size_t)b * 3 + c) * H + y0[yy]) * out_w + xx;
    const int* k = kk + (size_t)yy * ks;
    int ss = 1 << (kPrecisionBits - 1);
    const int n = cnt[yy];
    for (int y = 0; y < n; ++y) ss += (int)s[(size_t)y * out_w] * k[y];
    const float u = (float)clip8(ss);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), mean), sd);
    const int tile = (yy / 448) * gw + xx / 448;
    out[((((size_t)b * tiles + tile) * 3 + c) * 448 + (yy % 448)) * 448 + (xx % 448)] = __float2bfloat16(v);
  }
}

inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  size_t cap = (size_t)slb_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int slb_preprocess_frames(const uint8_t* frames, uint8_t* tmp, const slb_resample_table* horiz, const slb_resample_table* vert,
                                     void* tiles_out, int batch, int height, int width, int grid_w, int grid_h, void* stream) {
  SLB_CHECK_ARG(frames && tmp && horiz && vert && tiles_out && batch > 0 && height > 0 && width > 0 && grid_w > 0 && grid_h > 0,
                "preprocess_frames: bad args");
  const int out_w = 448 * grid_w, out_h = 448 * grid_h;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = batch * 3 * height;
  resample_h_kernel<<<grid_for((size_t)rows * out_w, 256), 256, 0, st>>>(frames, tmp, horiz->first, horiz->count, horiz->taps, horiz->ksize,
                                                                        rows, width, out_w);
  SLB_LAUNCH_CHECK();
  resample_v_tiles_kernel<<<grid_for((size_t)batch * 3 * out_h * out_w, 256), 256, 0, st>>>(
      tmp, (bf16*)tiles_out, vert->first, vert->count, vert->taps, vert->ksize, batch, height, out_w, out_h, grid_w, 0.485f, 0.456f, 0.406f,
      0.229f, 0.224f, 0.225f);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
