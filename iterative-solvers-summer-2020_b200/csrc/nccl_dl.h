// NCCL entry points resolved at run time with dlopen("libnccl.so.2"), so that libjfnk.so has no link-time
// dependency on a particular NCCL build: inside a PyTorch process this resolves to the NCCL that torch
// already loaded (same soname), and single-GPU use never touches NCCL at all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <string>

namespace jfnk {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
};

inline const NcclApi* nccl_api(std::string& why) {
  static NcclApi api;
  static int state = 0; // 0 untried, 1 ok, -1 failed
  static std::string err;
  if (state == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); state = -1; }
    else {
      bool ok = true;
      auto sym = [&](const char* n) -> void* {
        void* p = dlsym(h, n);
        if (!p) { ok = false; err = std::string("libnccl: missing symbol ") + n; }
        return p;
      };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      state = ok ? 1 : -1;
    }
  }
  if (state != 1) { why = err; return nullptr; }
  return &api;
}

} // namespace jfnk
