// nccl_dyn.cuh — NCCL is bound at run time (dlopen "libnccl.so.2" on first use) instead of at
// link time: a host process that also loads PyTorch must end up with ONE NCCL in the process
// (torch's bundled 2.28 and the system 2.27 share a SONAME), whichever library is loaded first.
#pragma once
#include <nccl.h>

namespace hb {

struct NcclApi {
   bool loaded = false;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl_api();
int nccl_load();   // 0 on success, hb200 error flag otherwise

}  // namespace hb
