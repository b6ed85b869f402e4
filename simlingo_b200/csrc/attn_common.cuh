// Register-level helpers shared by the flash-attention kernels (attention_vit.cu, attention_gqa.cu): 3-input max chains,
// packed fp32x2 FMA / ADD (Blackwell FFMA2), 32-score exp2 + bf16 pack + swizzled P store.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float max32(const uint32_t (&s)[32], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // four independent FMNMX3 chains
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
    m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
    m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
    m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
  }
  return fmaxf(fmax3(m0, m1, m2), m3);
}
// (a, b) * (sc, sc) + (-m, -m) on the packed fp32x2 pipe
__device__ __forceinline__ void ffma2(float& a, float& b, uint64_t sc2, uint64_t nm2) {
  uint64_t v, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(v), "l"(sc2), "l"(nm2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d));
}
__device__ __forceinline__ uint64_t pack2f(float a, float b) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  return v;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for two values WITHOUT the MUFU pipe (FA4-style software exp2): Cody-Waite split x = floor(x) + f through the 1.5 * 2^23
// magic constant under round-toward-minus-infinity, 2^f on [0, 1) by a degree-3 polynomial (max relative error 8.8e-5, 1/40 of
// the bf16 rounding P receives next), floor(x) added into the exponent field (LEA).  x is clamped at -127 (-> ~6e-39).  With
// head_dim 64 the exp2 stream needs 2 x the tensor time on the 16-lane MUFU pipe: moving ~1/3 of the elements onto the FMA /
// ALU pipes (10 issue slots per pair) balances MUFU time against issue slots.
__device__ __forceinline__ uint64_t ex2_emu2(uint64_t x2) {
  float a, b;
  f2unpack(x2, a, b);
  a = fmaxf(a, -127.f);
  b = fmaxf(b, -127.f);
  const uint64_t xc = f2pack(a, b);
  uint64_t xr;   // floor(x) + 1.5 * 2^23: the integer sits in the low mantissa bits
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(xr) : "l"(xc), "l"(f2splat(12582912.f)));
  const uint64_t xb = f2add(xr, f2splat(-12582912.f));        // floor(x), exact
  const uint64_t f = f2fma(xb, f2splat(-1.0f), xc);            // x - floor(x) in [0, 1)
  uint64_t p = f2fma(f, f2splat(0.077119089663028717f), f2splat(0.227564394474029541f));
  p = f2fma(p, f, f2splat(0.695146143436431885f));
  p = f2fma(p, f, f2splat(1.0f));                              // 2^f in [1, 2)
  uint32_t r0, r1, p0, p1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(r0), "=r"(r1) : "l"(xr));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(p));
  r0 = (r0 << 23) + p0;
  r1 = (r1 << 23) + p1;
  uint64_t out;
  asm("mov.b64 %0, {%1, %2};" : "=l"(out) : "r"(r0), "r"(r1));
  return out;
}
// exp2 of a packed pair: software path for the pairs selected by the compile-time mask, MUFU otherwise
template <uint32_t EMU_MASK>
__device__ __forceinline__ uint64_t ex2_pair(uint64_t x2, int i) {
  if ((EMU_MASK >> i) & 1u) return ex2_emu2(x2);
  float a, b;
  f2unpack(x2, a, b);
  return f2pack(ex2_approx(a), ex2_approx(b));
}

// p = exp2(s * scale - mref) for 32 scores; accumulates the (packed) row sum; writes 64 bytes (4 x 16 B chunks) of P into the
// K-major 128B-swizzled UMMA layout (row r = 128 bytes per 64-key half, chunk index XOR (r & 7)).
__device__ __forceinline__ void exp_store32_plain(const uint32_t (&s)[32], uint64_t sc2, uint64_t nm2, uint64_t& sum2, uint8_t* half_row,
                                                  int chunk0, int rsw) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = __uint_as_float(s[2 * i]), b = __uint_as_float(s[2 * i + 1]);
    ffma2(a, b, sc2, nm2);
    a = ex2_approx(a);
    b = ex2_approx(b);
    sum2 = fadd2(sum2, pack2f(a, b));
    pk[i] = pack_bf16(a, b);
  }
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    *reinterpret_cast<uint4*>(half_row + (((chunk0 + q4) ^ rsw) << 4)) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
}
