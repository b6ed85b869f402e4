// Fused multi-tensor AdamW over flat buffers (fp32 master weights + moments, bf16 gradients, bf16 model copy) with
// the global-norm clip folded in as a scale, and the squared-norm reduction that feeds it.
// Reference: torch.optim.AdamW(lr, weight_decay=0.1, betas) over all trainable tensors (driving.py:718-724) and
// Trainer(gradient_clip_val=0.3) (train.py:206).  HBM-bound: 2 (grad) + 12 (p,m,v read) + 12 (write) + 2 (bf16 copy) B/param.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

__global__ void __launch_bounds__(256)
sqnorm_kernel(const bf16* __restrict__ g, size_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  const size_t nvec = n >> 3;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(g + i * 8);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = nvec * 8; i < n; ++i) { const float v = __bfloat162float(g[i]); s += v * v; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const bf16* __restrict__ g, bf16* __restrict__ pb, size_t n,
             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, const float* __restrict__ sqnorm,
             float max_norm, float prescale) {
  // clip_grad_norm_ semantics: coef = max_norm / (||g|| + 1e-6), clamped to 1 (applied to the pre-scaled gradient)
  float coef = prescale;
  if (sqnorm && max_norm > 0.f) {
    const float norm = sqrtf(*sqnorm) * prescale;
    coef *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float decay = 1.0f - lr * wd, step = lr / bc1;
  const size_t nvec = n >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const uint2 gu = *reinterpret_cast<const uint2*>(g + i * 4);
    const float2 g01 = unpack_bf16(gu.x), g23 = unpack_bf16(gu.y);
    const float gg[4] = {g01.x * coef, g01.y * coef, g23.x * coef, g23.y * coef};
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pa[e] *= decay;
      ma[e] = beta1 * ma[e] + (1.0f - beta1) * gg[e];
      va[e] = beta2 * va[e] + (1.0f - beta2) * gg[e] * gg[e];
      pa[e] -= step * ma[e] / (sqrtf(va[e]) / bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    uint2 o;
    o.x = pack_bf16(pa[0], pa[1]);
    o.y = pack_bf16(pa[2], pa[3]);
    *reinterpret_cast<uint2*>(pb + i * 4) = o;
  }
}

}  // namespace

extern "C" int slb_grad_sqnorm(const void* grad_bf16, int64_t n, float* out_sq, void* stream) {
  SLB_CHECK_ARG(n > 0 && out_sq, "grad_sqnorm: bad args");
  int grid = (int)((n / 8 + 255) / 256);
  const int cap = slb_num_sms() * 8;
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)grad_bf16, (size_t)n, out_sq);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_adamw_fused(float* master, float* m, float* v, const void* grad_bf16, void* param_bf16, int64_t n, float lr, float beta1,
                               float beta2, float eps, float wd, int step, const float* grad_sqnorm, float max_norm, float grad_prescale,
                               void* stream) {
  SLB_CHECK_ARG(n > 0 && (n % 4) == 0 && step >= 1, "adamw: n=%lld (must be a multiple of 4) step=%d", (long long)n, step);
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  int grid = (int)((n / 4 + 255) / 256);
  const int cap = slb_num_sms() * 16;
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(master, m, v, (const bf16*)grad_bf16, (bf16*)param_bf16, (size_t)n, lr, beta1, beta2, eps, wd,
                                                      bc1, sqrtf(bc2), grad_sqnorm, max_norm, grad_prescale);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
