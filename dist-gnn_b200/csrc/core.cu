// core.cu - errors, launch counter, random engine, pin memory.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <random>

#include "dgs_common.cuh"

namespace dgsb {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;  // B200
  }
  return cached;
}

// process-global engine, like ctx::random_engine (reference src/context/context.cc:5-6), but
// re-seedable so runs can be reproduced.
struct RandomEngine {
  std::mutex mu;
  std::mt19937_64 gen{std::random_device()()};
};
static RandomEngine &engine() {
  static RandomEngine e;
  return e;
}

}  // namespace dgsb

extern "C" {

int dgs_abi_version(void) { return DGS_B200_ABI_VERSION; }
const char *dgs_last_error(void) { return dgsb::g_err; }
int64_t dgs_launch_count(void) { return dgsb::g_launches.load(std::memory_order_relaxed); }
int dgs_sm_count(void) { return dgsb::sm_count(); }

uint64_t dgs_randn_uint64(void) {
  auto &e = dgsb::engine();
  std::lock_guard<std::mutex> lk(e.mu);
  return e.gen();
}
void dgs_seed(uint64_t seed) {
  auto &e = dgsb::engine();
  std::lock_guard<std::mutex> lk(e.mu);
  e.gen.seed(seed);
}

int dgs_host_register(void *host_ptr, size_t nbytes) {
  DGS_REQUIRE(host_ptr != nullptr && nbytes > 0, "dgs_host_register: empty buffer");
  cudaError_t e = cudaHostRegister(host_ptr, nbytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return 0;
  }
  DGS_CUDA_OK(e);
  return 0;
}
int dgs_enable_peer_access(int peer_device) {
  int dev = 0;
  DGS_CUDA_OK(cudaGetDevice(&dev));
  if (dev == peer_device) return 0;
  int can = 0;
  DGS_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  DGS_REQUIRE(can, "device %d cannot access device %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  DGS_CUDA_OK(e);
  return 0;
}

int dgs_set_l2_fetch_granularity(int bytes) {
  DGS_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "L2 fetch granularity must be 32, 64 or 128");
  DGS_CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return 0;
}

int dgs_host_unregister(void *host_ptr) {
  DGS_REQUIRE(host_ptr != nullptr, "dgs_host_unregister: null pointer");
  cudaError_t e = cudaHostUnregister(host_ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return 0;
  }
  DGS_CUDA_OK(e);
  return 0;
}

}  // extern "C"
