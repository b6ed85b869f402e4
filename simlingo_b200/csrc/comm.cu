// Data-parallel gradient exchange behind the C ABI: one NCCL communicator per process, created / destroyed explicitly,
// and the bucket all-reduce enqueued on a caller-supplied stream (slb_comm_init / slb_comm_destroy /
// slb_allreduce_bucket).  Replaces what the reference gets from Lightning DDP / DeepSpeed ZeRO-2 (train.py:160-168).
//
// NCCL is resolved at run time from the process image (the torch-bundled libnccl.so.2 that lib.py loads with RTLD_GLOBAL,
// else dlopen by soname): the library links against nothing but the CUDA runtime, so single-GPU users never need NCCL.
// The few NCCL types used are restated here (stable ABI since NCCL 2.0: 128-byte unique id passed by value, opaque
// communicator pointer, enum values of nccl.h).
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclFloat32 = 7, kNcclBfloat16 = 9 };  // ncclDataType_t
enum { kNcclSum = 0, kNcclAvg = 4 };           // ncclRedOp_t

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
  bool ok;
};

const NcclApi* nccl() {
  static NcclApi api = [] {
    NcclApi a{};
    void* h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclAllReduce")) {
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return a;
    }
    a.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    return a;
  }();
  return &api;
}

int nccl_fail(const NcclApi* a, const char* what, int rc) {
  return slb_fail(SLB_ENCCL, "%s: %s (nccl result %d)", what, a->GetErrorString ? a->GetErrorString(rc) : "?", rc);
}

}  // namespace

#define SLB_NCCL_API(a)                                                                                        \
  const NcclApi* a = nccl();                                                                                   \
  if (!a->ok) return slb_fail(SLB_ENCCL, "NCCL (libnccl.so.2) is not loadable in this process: %s", dlerror() ? dlerror() : "symbols missing")

extern "C" int slb_comm_version(void) {
  const NcclApi* a = nccl();
  int v = 0;
  if (!a->ok || !a->GetVersion || a->GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int slb_comm_unique_id(void* out_128_bytes) {
  SLB_CHECK_ARG(out_128_bytes != nullptr, "comm_unique_id: null output");
  SLB_NCCL_API(a);
  NcclUniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return nccl_fail(a, "ncclGetUniqueId", rc);
  memcpy(out_128_bytes, &id, sizeof(id));
  return SLB_OK;
}

extern "C" int slb_comm_init(void** comm_out, const void* unique_id_128_bytes, int rank, int world) {
  SLB_CHECK_ARG(comm_out != nullptr && unique_id_128_bytes != nullptr, "comm_init: null argument");
  SLB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "comm_init: rank %d of %d", rank, world);
  SLB_NCCL_API(a);
  NcclUniqueId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  NcclComm c = nullptr;
  const int rc = a->CommInitRank(&c, world, id, rank);  // uses the calling thread's current CUDA device
  if (rc != 0) return nccl_fail(a, "ncclCommInitRank", rc);
  *comm_out = c;
  return SLB_OK;
}

extern "C" int slb_comm_destroy(void* comm) {
  if (comm == nullptr) return SLB_OK;
  SLB_NCCL_API(a);
  const int rc = a->CommDestroy((NcclComm)comm);
  if (rc != 0) return nccl_fail(a, "ncclCommDestroy", rc);
  return SLB_OK;
}

// In-place all-reduce of one gradient bucket over NVLink / NVSwitch (NCCL picks ring / tree / NVLS).
// dtype: 0 = bf16, 1 = fp32; average: 0 = SUM (the 1/world factor is folded into the fused AdamW kernel), 1 = ncclAvg.
extern "C" int slb_allreduce_bucket(void* comm, void* buf, int64_t n, int dtype, int average, void* stream) {
  SLB_CHECK_ARG(comm != nullptr, "allreduce_bucket: null communicator (call slb_comm_init first)");
  SLB_CHECK_ARG(n >= 0 && (n == 0 || buf != nullptr), "allreduce_bucket: bad buffer (n=%lld)", (long long)n);
  SLB_CHECK_ARG(dtype == 0 || dtype == 1, "allreduce_bucket: dtype must be 0 (bf16) or 1 (fp32)");
  if (n == 0) return SLB_OK;
  SLB_NCCL_API(a);
  const int rc = a->AllReduce(buf, buf, (size_t)n, dtype == 0 ? kNcclBfloat16 : kNcclFloat32, average ? kNcclAvg : kNcclSum, (NcclComm)comm,
                              (cudaStream_t)stream);
  if (rc != 0) return nccl_fail(a, "ncclAllReduce", rc);
  return SLB_OK;
}
