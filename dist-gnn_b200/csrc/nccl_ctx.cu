// nccl_ctx.cu - process-global NCCL context + CUDA-IPC tensor p2p server.
//
// Replaces
//   nccl::NCCLContext / GetUniqueId / SetNCCL / Barrier_ / NCCLTensorAllGather_
//                                                   src/nccl/nccl_context.cc:13-112
//   cache::TensorP2PServer                          src/cache/tensor_p2p_cache.cc:11-132
// NCCL is bound at run time with dlopen (RTLD_NOLOAD first, so the libnccl torch already loaded
// is reused); the library therefore loads and exports its symbols on a box with no NCCL/GPU.
// NCCL is only used at build time (unique id, barrier, handle / id-list exchange); the per-batch
// path is in-kernel NVLink peer loads through the pointer tables created here.
// Differences from the reference, on purpose: exchanges go through ordinary device buffers (the
// reference cudaHostRegister()s stack memory and hands it to NCCL, tensor_p2p_cache.cc:54-63).
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "dgs_common.cuh"
#include "p2p_server.h"

namespace dgsb {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_api;

static int load_nccl() {
  if (g_api.handle) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    const char *env = getenv("DGS_NCCL_LIB");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char *n : names) {
    if (h) break;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  DGS_REQUIRE(h != nullptr, "NCCL not found: import torch first or set DGS_NCCL_LIB (%s)", dlerror());
#define LOAD(field, sym)                                               \
  g_api.field = (decltype(g_api.field))dlsym(h, sym);                  \
  DGS_REQUIRE(g_api.field != nullptr, "NCCL symbol %s missing", sym);
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(AllGather, "ncclAllGather");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  g_api.handle = h;
  return 0;
}

#define DGS_NCCL_OK(call)                                                                   \
  do {                                                                                      \
    ncclResult_t r__ = (call);                                                              \
    if (r__ != ncclSuccess) {                                                               \
      set_error("%s:%d NCCL call %s failed: %s", __FILE__, __LINE__, #call,                 \
                g_api.GetErrorString ? g_api.GetErrorString(r__) : "?");                    \
      return 200 + (int)r__;                                                                \
    }                                                                                       \
  } while (0)

struct NcclCtx {
  bool ready = false;
  ncclComm_t comm = nullptr;
  int rank = 0;
  int world = 1;
  cudaStream_t stream = nullptr;
  char *scratch = nullptr;  // device scratch: 64 B * (world + 1), at least 4 KB
  size_t scratch_bytes = 0;
};
static NcclCtx g_ctx;

}  // namespace dgsb

using namespace dgsb;

extern "C" {

int dgs_nccl_get_unique_id(int64_t out_id[16]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes = 16 int64");
  if (load_nccl()) return 1;
  ncclUniqueId id;
  DGS_NCCL_OK(g_api.GetUniqueId(&id));
  memcpy(out_id, &id, sizeof(id));
  return 0;
}

int dgs_nccl_set(int nranks, const int64_t id[16], int rank) {
  DGS_REQUIRE(nranks >= 1 && nranks <= DGS_MAX_DEVICES, "dgs_nccl_set: nranks %d out of [1,%d]",
              nranks, DGS_MAX_DEVICES);
  DGS_REQUIRE(rank >= 0 && rank < nranks, "dgs_nccl_set: rank %d out of range", rank);
  DGS_REQUIRE(!g_ctx.ready, "dgs_nccl_set: the NCCL context is process-global and already set");
  if (load_nccl()) return 1;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  DGS_NCCL_OK(g_api.CommInitRank(&g_ctx.comm, nranks, uid, rank));
  g_ctx.rank = rank;
  g_ctx.world = nranks;
  DGS_CUDA_OK(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  g_ctx.scratch_bytes = 4096 + 64 * (size_t)(nranks + 1);
  DGS_CUDA_OK(cudaMalloc(&g_ctx.scratch, g_ctx.scratch_bytes));
  DGS_CUDA_OK(cudaMemset(g_ctx.scratch, 0, g_ctx.scratch_bytes));
  g_ctx.ready = true;
  return 0;
}

int dgs_nccl_rank(void) { return g_ctx.rank; }
int dgs_nccl_world(void) { return g_ctx.world; }

int dgs_nccl_barrier(void) {
  if (!g_ctx.ready || g_ctx.world == 1) {
    DGS_CUDA_OK(cudaDeviceSynchronize());
    return 0;
  }
  DGS_CUDA_OK(cudaDeviceSynchronize());
  DGS_NCCL_OK(g_api.AllReduce(g_ctx.scratch, g_ctx.scratch, 1, ncclFloat, ncclSum, g_ctx.comm,
                              g_ctx.stream));
  DGS_CUDA_OK(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

// all-gather of `bytes` (<= 64) per rank through device scratch -> host `out` (world * bytes)
static int allgather_small(const void *local, size_t bytes, void *out_host) {
  if (!g_ctx.ready || g_ctx.world == 1) {
    memcpy(out_host, local, bytes);
    return 0;
  }
  DGS_REQUIRE(bytes <= 64, "allgather_small: %zu bytes > 64", bytes);
  char *send = g_ctx.scratch + 2048;
  char *recv = g_ctx.scratch + 4096 - 64;  // recv area: 64 * (world + 1) bytes from here
  recv = g_ctx.scratch + 4096;
  DGS_CUDA_OK(cudaMemcpyAsync(send, local, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  DGS_NCCL_OK(g_api.AllGather(send, recv, bytes, ncclChar, g_ctx.comm, g_ctx.stream));
  DGS_CUDA_OK(cudaMemcpyAsync(out_host, recv, bytes * g_ctx.world, cudaMemcpyDeviceToHost,
                              g_ctx.stream));
  DGS_CUDA_OK(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

int dgs_nccl_allgather_i64(int64_t value, int64_t *out_host) {
  DGS_REQUIRE(out_host != nullptr, "dgs_nccl_allgather_i64: null output");
  return allgather_small(&value, sizeof(value), out_host);
}

int dgs_nccl_allgatherv(const void *send_dev, int64_t send_bytes, void *const *recv_dev,
                        const int64_t *recv_bytes) {
  DGS_REQUIRE(recv_dev && recv_bytes, "dgs_nccl_allgatherv: null arrays");
  const int world = g_ctx.world, rank = g_ctx.rank;
  DGS_REQUIRE(recv_bytes[rank] == send_bytes, "dgs_nccl_allgatherv: recv_bytes[rank] != send_bytes");
  // the caller's buffers were produced on its own stream(s)
  DGS_CUDA_OK(cudaDeviceSynchronize());
  if (recv_dev[rank] != send_dev && send_bytes > 0)
    DGS_CUDA_OK(cudaMemcpyAsync(recv_dev[rank], send_dev, (size_t)send_bytes,
                                cudaMemcpyDeviceToDevice, g_ctx.stream));
  if (g_ctx.ready && world > 1) {
    DGS_NCCL_OK(g_api.GroupStart());
    for (int i = 0; i < world; ++i) {
      if (i == rank) continue;
      if (send_bytes > 0)
        DGS_NCCL_OK(g_api.Send(send_dev, (size_t)send_bytes, ncclChar, i, g_ctx.comm, g_ctx.stream));
      if (recv_bytes[i] > 0)
        DGS_NCCL_OK(g_api.Recv(recv_dev[i], (size_t)recv_bytes[i], ncclChar, i, g_ctx.comm,
                               g_ctx.stream));
    }
    DGS_NCCL_OK(g_api.GroupEnd());
  }
  DGS_CUDA_OK(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
int dgs_p2p_server_create(const void *dev_src, int64_t nbytes, dgs_p2p_server_t **out) {
  DGS_REQUIRE(out != nullptr, "dgs_p2p_server_create: null output");
  DGS_REQUIRE(nbytes > 0, "dgs_p2p_server_create: empty tensor (the reference CHECKs size(0) > 0, "
              "tensor_p2p_cache.cc:19)");
  dgs_p2p_server *s = new dgs_p2p_server();
  memset(s, 0, sizeof(*s));
  s->world = g_ctx.world;
  s->rank = g_ctx.rank;
  s->owns_local = 1;
  void *local = nullptr;
  // raw cudaMalloc (not the torch caching allocator): IPC handles need a whole allocation.
  cudaError_t e = cudaMalloc(&local, (size_t)nbytes);
  if (e != cudaSuccess) {
    delete s;
    set_error("dgs_p2p_server_create: cudaMalloc(%lld) failed: %s", (long long)nbytes,
              cudaGetErrorString(e));
    return 100 + (int)e;
  }
  // dev_src == NULL: allocate only, the caller fills the shard in place
  e = dev_src ? cudaMemcpy(local, dev_src, (size_t)nbytes, cudaMemcpyDefault) : cudaSuccess;
  if (e != cudaSuccess) {
    cudaFree(local);
    delete s;
    set_error("dgs_p2p_server_create: copy failed: %s", cudaGetErrorString(e));
    return 100 + (int)e;
  }
  s->ptrs[s->rank] = local;
  s->nbytes[s->rank] = nbytes;
  if (s->world > 1) {
    struct Msg {
      cudaIpcMemHandle_t h;  // 64 bytes
    } mine;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle is 64 bytes");
    DGS_CUDA_OK(cudaIpcGetMemHandle(&mine.h, local));
    std::vector<Msg> all(s->world);
    if (allgather_small(&mine, sizeof(mine), all.data())) return 1;
    std::vector<int64_t> sizes(s->world);
    if (dgs_nccl_allgather_i64(nbytes, sizes.data())) return 1;
    for (int i = 0; i < s->world; ++i) {
      s->nbytes[i] = sizes[i];
      if (i == s->rank) continue;
      DGS_CUDA_OK(cudaIpcOpenMemHandle(&s->ptrs[i], all[i].h, cudaIpcMemLazyEnablePeerAccess));
    }
    s->ipc_opened = 1;
    if (dgs_nccl_barrier()) return 1;
  }
  *out = s;
  return 0;
}

int dgs_p2p_server_create_virtual(int world, int rank, void *const *ptrs, const int64_t *nbytes,
                                  dgs_p2p_server_t **out) {
  DGS_REQUIRE(out && ptrs && nbytes, "dgs_p2p_server_create_virtual: null argument");
  DGS_REQUIRE(world >= 1 && world <= DGS_MAX_DEVICES && rank >= 0 && rank < world,
              "dgs_p2p_server_create_virtual: bad world/rank %d/%d", world, rank);
  dgs_p2p_server *s = new dgs_p2p_server();
  memset(s, 0, sizeof(*s));
  s->world = world;
  s->rank = rank;
  for (int i = 0; i < world; ++i) {
    s->ptrs[i] = ptrs[i];
    s->nbytes[i] = nbytes[i];
  }
  *out = s;
  return 0;
}

void *dgs_p2p_server_ptr(const dgs_p2p_server_t *s, int dev) {
  if (!s || dev < 0 || dev >= s->world) return nullptr;
  return s->ptrs[dev];
}
int64_t dgs_p2p_server_nbytes(const dgs_p2p_server_t *s, int dev) {
  if (!s || dev < 0 || dev >= s->world) return -1;
  return s->nbytes[dev];
}
int dgs_p2p_server_world(const dgs_p2p_server_t *s) { return s ? s->world : 0; }
int dgs_p2p_server_rank(const dgs_p2p_server_t *s) { return s ? s->rank : -1; }

int dgs_p2p_server_destroy(dgs_p2p_server_t *s, int barrier) {
  if (!s) return 0;
  if (s->ipc_opened) {
    for (int i = 0; i < s->world; ++i)
      if (i != s->rank && s->ptrs[i]) cudaIpcCloseMemHandle(s->ptrs[i]);
  }
  int rc = 0;
  if (barrier && s->world > 1 && s->ipc_opened) rc = dgs_nccl_barrier();
  if (s->owns_local && s->ptrs[s->rank]) cudaFree(s->ptrs[s->rank]);
  delete s;
  return rc;
}

}  // extern "C"
