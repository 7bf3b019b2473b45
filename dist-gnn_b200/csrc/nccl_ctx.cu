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
#include <cuda.h>
#include <dlfcn.h>
#include <errno.h>
#include <poll.h>
#include <nccl.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <thread>
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
  uint64_t uid_hash = 0;   // names the fd-exchange sockets of this communicator
  int64_t serial = 0;      // p2p servers created so far (same on every rank: creation is collective)
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

extern "C" int dgs_nccl_barrier(void);
extern "C" int dgs_nccl_allgather_i64(int64_t value, int64_t *out_host);

namespace dgsb {

// ------------------------------------------------------------------------------------------
// VMM shards: cuMemCreate + POSIX-fd export, fds passed between the ranks of the box over
// abstract unix sockets (SCM_RIGHTS), cuMemMap on every rank.  Measured on B200: in-kernel random
// 512-B row reads from a 28 GB peer shard run at ~75 GB/s through a legacy cudaIpcOpenMemHandle
// mapping but at 640-740 GB/s through a cuMemMap / plain peer mapping (tools/peer_probe.py), so
// the legacy CUDA-IPC path of the reference (tensor_p2p_cache.cc:52-73) is only the fallback.
struct DrvApi {
  bool ok = false;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
};
static DrvApi g_drv;

static bool load_drv() {
  if (g_drv.ok) return true;
  cudaFree(0);
#define DRV(field, name)                                                                      \
  do {                                                                                        \
    void *fn = nullptr;                                                                       \
    cudaDriverEntryPointQueryResult q;                                                        \
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {    \
      cudaGetLastError();                                                                     \
      return false;                                                                           \
    }                                                                                         \
    g_drv.field = (decltype(g_drv.field))fn;                                                  \
  } while (0)
  DRV(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
  DRV(MemCreate, "cuMemCreate");
  DRV(MemAddressReserve, "cuMemAddressReserve");
  DRV(MemMap, "cuMemMap");
  DRV(MemSetAccess, "cuMemSetAccess");
  DRV(MemExportToShareableHandle, "cuMemExportToShareableHandle");
  DRV(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
  DRV(MemUnmap, "cuMemUnmap");
  DRV(MemRelease, "cuMemRelease");
  DRV(MemAddressFree, "cuMemAddressFree");
#undef DRV
  g_drv.ok = true;
  return true;
}

static void sock_name(sockaddr_un *addr, socklen_t *len, int64_t serial, int rank) {
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  // abstract namespace (leading NUL): nothing to unlink, private to the box
  int n = snprintf(addr->sun_path + 1, sizeof(addr->sun_path) - 1, "dgs_b200_%016llx_%lld_%d",
                   (unsigned long long)g_ctx.uid_hash, (long long)serial, rank);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

static int send_fd(int sock, int fd) {
  char dummy = 'x', ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  iovec iov{&dummy, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}

static int recv_fd(int sock) {
  char dummy, ctrl[CMSG_SPACE(sizeof(int))];
  iovec iov{&dummy, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, 0) != 1) return -1;
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd;
  memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

// every rank serves its own fd to the world-1 peers and fetches theirs; returns 0 on success.
// Collective: EVERY rank runs both barriers whatever happened locally, so a rank whose socket set-up
// or a peer connection failed cannot leave the others blocked: the accept loop polls with a
// timeout and is told to stop once every rank has finished its client side (second barrier) - by
// then no peer can still be waiting for this rank's fd.
static int exchange_fds(int my_fd, int world, int rank, int64_t serial, int *peer_fds) {
  sockaddr_un addr;
  socklen_t alen;
  sock_name(&addr, &alen, serial, rank);
  int srv = socket(AF_UNIX, SOCK_STREAM, 0);
  int err = 0;
  if (srv < 0 || bind(srv, (sockaddr *)&addr, alen) != 0 || listen(srv, world) != 0) {
    if (srv >= 0) close(srv);
    srv = -1;
    err = 1;
  }
  std::atomic<int> stop{0};
  int srv_err = 0;
  std::thread server;
  if (srv >= 0) {
    server = std::thread([&]() {
      int served = 0;
      while (served < world - 1 && !stop.load(std::memory_order_acquire)) {
        pollfd pfd{srv, POLLIN, 0};
        const int pr = poll(&pfd, 1, 100 /* ms */);
        if (pr < 0 && errno != EINTR) {
          srv_err = 1;
          break;
        }
        if (pr <= 0 || !(pfd.revents & POLLIN)) continue;
        int c = accept(srv, nullptr, nullptr);
        if (c < 0 || send_fd(c, my_fd) != 0) srv_err = 1;
        if (c >= 0) close(c);
        ++served;
      }
    });
  }
  if (dgs_nccl_barrier()) err = 1;  // every rank that could is listening
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    sockaddr_un pa;
    socklen_t pl;
    sock_name(&pa, &pl, serial, p);
    int c = socket(AF_UNIX, SOCK_STREAM, 0);
    timeval tv{30, 0};  // a peer that listens but never serves must not block us for ever
    if (c >= 0) setsockopt(c, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    if (c < 0 || connect(c, (sockaddr *)&pa, pl) != 0) {
      err = 1;
    } else {
      peer_fds[p] = recv_fd(c);
      if (peer_fds[p] < 0) err = 1;
    }
    if (c >= 0) close(c);
  }
  if (dgs_nccl_barrier()) err = 1;  // every rank has finished fetching: nobody waits for us any more
  stop.store(1, std::memory_order_release);
  if (server.joinable()) server.join();
  if (srv >= 0) close(srv);
  return (err || srv_err) ? -1 : 0;
}

static int vmm_map(dgs_p2p_server *s, int slot, int dev, CUmemGenericAllocationHandle h, size_t size,
                   size_t gran) {
  CUdeviceptr ptr = 0;
  if (g_drv.MemAddressReserve(&ptr, size, gran, 0, 0) != CUDA_SUCCESS) return -1;
  if (g_drv.MemMap(ptr, size, 0, h, 0) != CUDA_SUCCESS) {
    g_drv.MemAddressFree(ptr, size);
    return -1;
  }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (g_drv.MemSetAccess(ptr, size, &acc, 1) != CUDA_SUCCESS) {
    g_drv.MemUnmap(ptr, size);
    g_drv.MemAddressFree(ptr, size);
    return -1;
  }
  s->ptrs[slot] = (void *)ptr;
  s->vmm_handle[slot] = (unsigned long long)h;
  s->vmm_size[slot] = (int64_t)size;
  return 0;
}

static void vmm_destroy(dgs_p2p_server *s);

// returns 0 on success, 1 when the VMM path is unavailable (caller falls back to legacy IPC),
// >= 2 on a hard error (set_error called).  After the availability agreement every rank walks
// through the SAME sequence of collectives whatever fails locally; failures are accumulated and
// agreed on at the end, then everything mapped so far is released on every rank.
static int vmm_create(dgs_p2p_server *s, const void *dev_src, int64_t nbytes) {
  if (getenv("DGS_P2P_LEGACY_IPC")) return 1;  // (set it on every rank or on none)
  int dev = 0;
  size_t gran = 0, size = 0;
  CUmemGenericAllocationHandle h = 0;
  int fd = -1;
  int have = 0;
  CUmemAllocationProp prop{};
  if (load_drv() && cudaGetDevice(&dev) == cudaSuccess) {
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
        gran > 0) {
      size = ((size_t)nbytes + gran - 1) / gran * gran;
      have = (g_drv.MemCreate(&h, size, &prop, 0) == CUDA_SUCCESS) ? 1 : 0;
      if (have && g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS)
        have = 0;
    }
  }
  // all ranks must take the same path: agree on availability
  std::vector<int64_t> flags(s->world);
  if (dgs_nccl_allgather_i64(have, flags.data())) {
    if (fd >= 0) close(fd);
    if (h) g_drv.MemRelease(h);
    return 2;
  }
  bool all = true;
  for (int i = 0; i < s->world; ++i) all = all && flags[i] == 1;
  if (!all) {
    if (fd >= 0) close(fd);
    if (h) g_drv.MemRelease(h);
    return 1;
  }
  s->vmm = 1;
  int fail = 0;
  char why[256] = "";
  if (vmm_map(s, s->rank, dev, h, size, gran) != 0) {
    snprintf(why, sizeof(why), "cuMemMap of the local shard (%zu bytes) failed", size);
    g_drv.MemRelease(h);   // never mapped: vmm_destroy will not see it
    fail = 1;
  }
  if (!fail && dev_src) {
    cudaError_t e = cudaMemcpy(s->ptrs[s->rank], dev_src, (size_t)nbytes, cudaMemcpyDefault);
    if (e != cudaSuccess) {
      snprintf(why, sizeof(why), "copy into the shard failed: %s", cudaGetErrorString(e));
      fail = 1;
    }
  }
  std::vector<int64_t> sizes(s->world), msizes(s->world);
  if (dgs_nccl_allgather_i64(nbytes, sizes.data())) fail = 1;
  if (dgs_nccl_allgather_i64((int64_t)size, msizes.data())) fail = 1;
  std::vector<int> fds(s->world, -1);
  const int64_t serial = g_ctx.serial++;
  if (exchange_fds(fd, s->world, s->rank, serial, fds.data()) != 0) {
    if (!fail) snprintf(why, sizeof(why), "fd exchange over unix sockets failed");
    fail = 1;
  }
  close(fd);
  for (int i = 0; i < s->world; ++i) {
    s->nbytes[i] = sizes[i];
    if (i == s->rank || fds[i] < 0) continue;
    if (!fail) {
      CUmemGenericAllocationHandle ph = 0;
      if (g_drv.MemImportFromShareableHandle(&ph, (void *)(uintptr_t)fds[i], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) {
        snprintf(why, sizeof(why), "importing the shard of rank %d failed", i);
        fail = 1;
      } else if (vmm_map(s, i, dev, ph, (size_t)msizes[i], gran) != 0) {
        snprintf(why, sizeof(why), "mapping the shard of rank %d failed", i);
        g_drv.MemRelease(ph);
        fail = 1;
      }
    }
    close(fds[i]);
  }
  // agree on the outcome (also the barrier after which every shard may be read)
  bool any_fail = fail != 0;
  if (dgs_nccl_allgather_i64(fail, flags.data()))
    any_fail = true;
  else
    for (int i = 0; i < s->world; ++i) any_fail = any_fail || flags[i] != 0;
  if (any_fail) {
    vmm_destroy(s);   // unmaps / releases whatever was mapped, here and (same decision) on every rank
    memset(s->ptrs, 0, sizeof(s->ptrs));
    set_error("dgs_p2p_server_create: %s", why[0] ? why : "shard set-up failed on another rank");
    return 2;
  }
  return 0;
}

static void vmm_destroy(dgs_p2p_server *s) {
  for (int i = 0; i < s->world; ++i) {
    if (!s->ptrs[i]) continue;
    g_drv.MemUnmap((CUdeviceptr)s->ptrs[i], (size_t)s->vmm_size[i]);
    g_drv.MemAddressFree((CUdeviceptr)s->ptrs[i], (size_t)s->vmm_size[i]);
    g_drv.MemRelease((CUmemGenericAllocationHandle)s->vmm_handle[i]);
  }
}

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
  g_ctx.uid_hash = 1469598103934665603ull;
  for (int i = 0; i < 16; ++i) g_ctx.uid_hash = (g_ctx.uid_hash ^ (uint64_t)id[i]) * 1099511628211ull;
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
  if (s->world > 1) {
    const int rc = vmm_create(s, dev_src, nbytes);
    if (rc == 0) {
      *out = s;
      return 0;
    }
    if (rc >= 2) {
      delete s;
      return rc;
    }
    memset(s->ptrs, 0, sizeof(s->ptrs));  // VMM unavailable on some rank: legacy CUDA IPC below
    s->vmm = 0;
  }
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
    // errors below release everything this call opened / allocated before returning
    auto fail = [&](int rc) {
      for (int i = 0; i < s->world; ++i)
        if (i != s->rank && s->ptrs[i]) cudaIpcCloseMemHandle(s->ptrs[i]);
      cudaFree(local);
      delete s;
      return rc;
    };
    cudaError_t ce = cudaIpcGetMemHandle(&mine.h, local);
    if (ce != cudaSuccess) {
      set_error("dgs_p2p_server_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(ce));
      return fail(100 + (int)ce);
    }
    std::vector<Msg> all(s->world);
    if (allgather_small(&mine, sizeof(mine), all.data())) return fail(1);
    std::vector<int64_t> sizes(s->world);
    if (dgs_nccl_allgather_i64(nbytes, sizes.data())) return fail(1);
    for (int i = 0; i < s->world; ++i) {
      s->nbytes[i] = sizes[i];
      if (i == s->rank) continue;
      ce = cudaIpcOpenMemHandle(&s->ptrs[i], all[i].h, cudaIpcMemLazyEnablePeerAccess);
      if (ce != cudaSuccess) {
        s->ptrs[i] = nullptr;
        set_error("dgs_p2p_server_create: cudaIpcOpenMemHandle(rank %d) failed: %s", i,
                  cudaGetErrorString(ce));
        return fail(100 + (int)ce);
      }
    }
    s->ipc_opened = 1;
    if (dgs_nccl_barrier()) return fail(1);
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
  if (s->vmm) {
    int rc = 0;
    cudaDeviceSynchronize();
    if (barrier && s->world > 1) rc = dgs_nccl_barrier();  // nobody reads our shard any more
    vmm_destroy(s);
    delete s;
    return rc;
  }
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
