#include <cstdlib>
// Host-side plumbing of the C ABI: error string, device query, TMA descriptor encoding.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

#include <cudaTypedefs.h>

thread_local char g_slb_err[512] = {0};

int slb_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_slb_err, sizeof(g_slb_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* slb_last_error(void) { return g_slb_err; }
extern "C" int slb_version(void) { return 100; }

extern "C" int slb_num_sms(void) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

int slb_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SLB_NO_PDL"); v = (e && atoi(e)) ? 0 : 1; }
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int slb_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)r,
                    (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)outer_stride_bytes, box_inner,
                    box_outer);
  return SLB_OK;
}

// fp32 2-D map (inner box = 32 elements = 128 bytes, 128B swizzle): the residual-stream tiles the GEMM epilogue moves with TMA
int slb_make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled(2d f32) failed: %d", (int)r);
  return SLB_OK;
}

int slb_make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return SLB_OK;
}

int slb_make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                         uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return slb_fail(SLB_ECUDA, "cuTensorMapEncodeTiled(3d f32) failed: %d", (int)r);
  return SLB_OK;
}
