// comm.cu — the one exchange step of the data-parallel path behind the C-ABI:
// a SUM all-reduce of the flat fp32 gradient buffer over NVLink 5 / NVSwitch.
//
// The reference is single-process and has no collective at all; the mean-over-batch
// objective of net.py:389 makes the training step data parallel by utterance, and this is its only
// communication.  One process per GPU: rank 0 creates an id (sl_comm_unique_id), the host side
// distributes the 128 bytes (torch.distributed is the plumbing for that, nothing else), every rank calls
// sl_comm_init_rank.  The transport is NCCL, resolved at run time with dlopen("libnccl.so.2") so that
// the library shares the copy already loaded in the process (torch's) and has no link-time dependency.
//
// Why own the communicator instead of using torch.distributed's: the collective runs CONCURRENTLY with
// the backward pass, whose tcgen05 kernels are persistent grids of one CTA per SM (148).  NCCL's default
// channel count takes 16-32 SMs away from them for the duration of a bucket; a communicator created
// with ncclConfig_t.maxCTAs = a few CTAs moves 100 MB per 5 ms step just as well over NVSwitch and leaves
// the tensor pipes alone (round-1 scaling run: +0.30 ms on the two big_conv_1 gradient kernels at 8 GPUs).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/speechless_b200.h"
#include "common.cuh"

namespace sl {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  const char* (*GetLastError)(ncclComm_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // 1. SL_NCCL_LIB, 2. the copy already mapped into the process (the Python host has imported torch,
    // whose bundled NCCL must stay the only one: a second, older libnccl.so.2 loaded first would shadow
    // it), 3. the system library.
    const char* override_path = std::getenv("SL_NCCL_LIB");
    if (override_path != nullptr) api.handle = dlopen(override_path, RTLD_NOW | RTLD_LOCAL);
    if (api.handle == nullptr) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (api.handle == nullptr) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (api.handle == nullptr) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (api.handle == nullptr) return;
#define SL_SYM(field, symbol) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, symbol))
    SL_SYM(GetUniqueId, "ncclGetUniqueId");
    SL_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
    SL_SYM(AllReduce, "ncclAllReduce");
    SL_SYM(CommDestroy, "ncclCommDestroy");
    SL_SYM(CommCount, "ncclCommCount");
    SL_SYM(GetVersion, "ncclGetVersion");
    SL_SYM(GetErrorString, "ncclGetErrorString");
    SL_SYM(GetLastError, "ncclGetLastError");
#undef SL_SYM
    api.ok = api.GetUniqueId && api.CommInitRankConfig && api.AllReduce && api.CommDestroy && api.GetErrorString;
  });
  return api;
}

int nccl_fail(ncclResult_t r, const char* what, ncclComm_t comm = nullptr) {
  NcclApi& api = nccl();
  std::string msg = std::string("NCCL error: ") + (api.GetErrorString ? api.GetErrorString(r) : "?") + " in " + what;
  if (api.GetLastError) {
    const char* detail = api.GetLastError(comm);
    if (detail != nullptr && detail[0] != '\0') msg += std::string(" (") + detail + ")";
  }
  set_error(msg);
  return SL_ERR_CUDA;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 0;
  int rank = 0;
  int max_ctas = 0;
};

}  // namespace
}  // namespace sl

using namespace sl;

extern "C" {

int sl_comm_unique_id(void* id_out) {
  SL_REQUIRE(id_out != nullptr, "null pointer");
  static_assert(sizeof(ncclUniqueId) == SL_COMM_ID_BYTES, "SL_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  NcclApi& api = nccl();
  if (!api.ok) {
    set_error("libnccl.so.2 could not be loaded (set SL_NCCL_LIB to its path)");
    return SL_ERR_CUDA;
  }
  ncclUniqueId id;
  const ncclResult_t r = api.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  std::memcpy(id_out, &id, sizeof(id));
  return SL_OK;
}

int sl_comm_init_rank(void** comm_out, const void* id, int nranks, int rank, int max_ctas) {
  SL_REQUIRE(comm_out != nullptr && id != nullptr, "null pointer");
  SL_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / world size");
  NcclApi& api = nccl();
  if (!api.ok) {
    set_error("libnccl.so.2 could not be loaded (set SL_NCCL_LIB to its path)");
    return SL_ERR_CUDA;
  }
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclConfig_t config = NCCL_CONFIG_INITIALIZER;
  config.blocking = 1;
  if (max_ctas > 0) {
    // few CTAs: the collective shares the GPU with persistent one-CTA-per-SM tensor-core kernels
    config.minCTAs = 1;
    config.maxCTAs = max_ctas;
  }
  Comm* c = new Comm();
  c->nranks = nranks;
  c->rank = rank;
  c->max_ctas = max_ctas;
  const ncclResult_t r = api.CommInitRankConfig(&c->comm, nranks, uid, rank, &config);
  if (r != ncclSuccess) {
    delete c;
    return nccl_fail(r, "ncclCommInitRankConfig");
  }
  *comm_out = c;
  return SL_OK;
}

int sl_comm_size(void* comm) {
  if (comm == nullptr) return 0;
  Comm* c = static_cast<Comm*>(comm);
  NcclApi& api = nccl();
  int n = 0;
  if (api.CommCount && api.CommCount(c->comm, &n) == ncclSuccess) return n;
  return c->nranks;
}

int sl_allreduce_sum(void* comm, float* buf, size_t count, void* stream) {
  SL_REQUIRE(comm != nullptr && buf != nullptr, "null pointer");
  if (count == 0) return SL_OK;
  Comm* c = static_cast<Comm*>(comm);
  const ncclResult_t r = nccl().AllReduce(buf, buf, count, ncclFloat32, ncclSum, c->comm, static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce", c->comm);
  return SL_OK;
}

int sl_comm_destroy(void* comm) {
  if (comm == nullptr) return SL_OK;
  Comm* c = static_cast<Comm*>(comm);
  const ncclResult_t r = nccl().CommDestroy(c->comm);
  delete c;
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommDestroy");
  return SL_OK;
}

int sl_comm_nccl_version(void) {
  NcclApi& api = nccl();
  int v = 0;
  if (api.ok && api.GetVersion && api.GetVersion(&v) == ncclSuccess) return v;
  return 0;
}

}  // extern "C"
