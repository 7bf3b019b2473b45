// dgs_common.cuh - shared device/host helpers for the dgs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/dgs_b200.h"

namespace dgsb {

// ---------------------------------------------------------------- errors / bookkeeping
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
int sm_count();

#define DGS_CUDA_OK(call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      dgsb::set_error("%s:%d CUDA call %s failed: %s", __FILE__, __LINE__, #call,      \
                      cudaGetErrorString(e__));                                        \
      return 100 + (int)e__;                                                           \
    }                                                                                  \
  } while (0)

#define DGS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      dgsb::set_error(__VA_ARGS__);   \
      return 1;                       \
    }                                 \
  } while (0)

#define DGS_LAUNCH_CHECK()                  \
  do {                                      \
    dgsb::g_launches.fetch_add(1, std::memory_order_relaxed); \
    DGS_CUDA_OK(cudaGetLastError());        \
  } while (0)

// dispatch on id type (DGS_ID_TYPE_SWITCH, reference src/common/dgs_headers.h:47-58)
#define DGS_ITYPE_SWITCH(it, T, ...)                          \
  do {                                                        \
    if ((it) == DGS_I32) {                                    \
      typedef int32_t T;                                      \
      { __VA_ARGS__ }                                         \
    } else if ((it) == DGS_I64) {                             \
      typedef int64_t T;                                      \
      { __VA_ARGS__ }                                         \
    } else {                                                  \
      dgsb::set_error("id type must be int32(0) or int64(1), got %d", (int)(it)); \
      return 1;                                               \
    }                                                         \
  } while (0)

// ---------------------------------------------------------------- peer pointer table (by value
// in kernel params: no dependent pointer load, unlike tensor_p2p_server_wrapper::At,
// reference src/cache/tensor_p2p_cache.h:21-23)
struct PtrTable {
  const void *p[DGS_MAX_DEVICES];
};

// ---------------------------------------------------------------- location table slots
struct __align__(16) LocSlot {
  long long key;  // -1 = empty
  long long val;  // prio<<56 | dev<<48 | idx
};
static constexpr long long kEmptyKey = -1;
static constexpr int kDevShift = 48;
static constexpr long long kIdxMask = (1ll << 48) - 1;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ int4 ld_nc_v4(const void *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_v4(const void *p) {
  int4 r;
  asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_v4(void *p, const int4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// probe: returns val (>=0) or -1.  One 16-byte load per probe, linear probing.
__device__ __forceinline__ long long loc_lookup(const LocSlot *__restrict__ table, uint64_t cap_mask,
                                                long long key) {
  uint64_t pos = mix64((uint64_t)key) & cap_mask;
  while (true) {
    int4 raw = ld_v4(&table[pos]);
    long long k = ((long long)(uint32_t)raw.y << 32) | (uint32_t)raw.x;
    if (k == key) return ((long long)(uint32_t)raw.w << 32) | (uint32_t)raw.z;
    if (k == kEmptyKey) return -1;
    pos = (pos + 1) & cap_mask;
  }
}

// ---------------------------------------------------------------- Philox4x32-10 (counter RNG)
// The reference draws from curandStatePhilox4_32_10_t (rowwise_sampling.cu:62-63); this is the
// same public algorithm computed inline from (key, counter) so no per-thread state is kept and
// results depend only on (seed, item index, draw index), not on the launch geometry.
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // 4 x 32 random bits for (key, counter)
  __host__ __device__ static inline uint4 gen(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi) {
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi,
                     (uint32_t)(ctr_hi >> 32)};
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += W0;
      k1 += W1;
    }
    uint4 r;
    r.x = c[0]; r.y = c[1]; r.z = c[2]; r.w = c[3];
    return r;
  }
};

// uniform float in (0, 1]  (same open/closed convention as curand_uniform)
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t x) {
  return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}
// unbiased-enough integer in [0, n): multiply-high (bias <= n / 2^32, the reference's
// curand() % n has the same order of bias, rowwise_sampling.cu:85)
__device__ __forceinline__ uint32_t rand_below(uint32_t r, uint32_t n) { return __umulhi(r, n); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024).
// returns exclusive prefix; *total gets the block sum.  smem: 32 entries of T.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *smem, T *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? smem[lane] : (T)0;
    T winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  T res = smem[warp] + inc - v;
  __syncthreads();
  return res;
}

inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace dgsb
