// loc_table.cu - GPU open-addressing node-id -> (device, slot) table for the hot-node cache.
//
// Replaces hashmap::cuda::Hashmap (src/hashmap/cuda/hashmap.h:12-90) and
// CreateNidsP2PCacheHashMapCUDA (src/hashmap/cuda/hashmap.cu:15-77).
// B200 layout: one 16-byte slot {key, prio|dev|idx} so a probe is a single 128-bit load and the
// next probe of the linear sequence is in the same 128-byte line (the reference keeps three
// parallel int64 arrays = three dependent random sectors per hit, and re-hashes on collision).
// Ownership rule of the reference, made deterministic: for an id cached on several devices the
// local copy wins, otherwise the device inserted last, i.e. the largest cyclic offset
// (dev - rank) mod world  (hashmap.cu:37-72 inserts (rank+1)%P .. (rank+P-1)%P, then local).
// Here all devices are inserted concurrently and the winner is picked with atomicMax on
// prio<<56 | dev<<48 | idx.
#include "dgs_common.cuh"

namespace dgsb {

// set by loc_insert_kernel when a key found no slot (more distinct ids than capacity - 1) or was
// negative; read back (and cleared) by dgs_loc_table_build after the inserts
__device__ int g_loc_build_error;

template <typename IdT>
__global__ void loc_insert_kernel(LocSlot *table, uint64_t cap_mask, const IdT *__restrict__ nids,
                                  int64_t n, long long dev_bits) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    long long key = (long long)nids[i];
    if (key < 0) {   // -1 is the empty marker
      g_loc_build_error = 2;
      continue;
    }
    uint64_t pos = mix64((uint64_t)key) & cap_mask;
    // an id cached on several devices occupies ONE slot (the reference's Update() overwrites the
    // entry, hashmap.h:18-32), so only the number of DISTINCT ids is bounded by the capacity; the
    // probe sequence is cut after one full cycle instead of trusting a host-side total
    uint64_t probes = 0;
    bool placed = false;
    while (probes <= cap_mask) {
      unsigned long long prev =
          atomicCAS((unsigned long long *)&table[pos].key, (unsigned long long)kEmptyKey,
                    (unsigned long long)key);
      if (prev == (unsigned long long)kEmptyKey || prev == (unsigned long long)key) {
        placed = true;
        break;
      }
      pos = (pos + 1) & cap_mask;
      ++probes;
    }
    if (placed)
      atomicMax(&table[pos].val, dev_bits | (long long)i);
    else
      g_loc_build_error = 1;
  }
}

template <typename IdT>
__global__ void loc_lookup_kernel(const LocSlot *__restrict__ table, uint64_t cap_mask,
                                  const IdT *__restrict__ nids, int64_t n, IdT *out_dev,
                                  IdT *out_idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    long long v = loc_lookup(table, cap_mask, (long long)nids[i]);
    out_dev[i] = v < 0 ? (IdT)-1 : (IdT)((v >> kDevShift) & 0xff);
    out_idx[i] = v < 0 ? (IdT)-1 : (IdT)(v & kIdxMask);
  }
}

template <typename IdT>
__global__ void loc_unpack_kernel(const LocSlot *__restrict__ table, int64_t cap, IdT *key,
                                  IdT *idx, IdT *devid) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap;
       i += (int64_t)gridDim.x * blockDim.x) {
    LocSlot s = table[i];
    bool empty = s.key == kEmptyKey;
    key[i] = (IdT)s.key;
    idx[i] = empty ? (IdT)-1 : (IdT)(s.val & kIdxMask);
    devid[i] = empty ? (IdT)-1 : (IdT)((s.val >> kDevShift) & 0xff);
  }
}

}  // namespace dgsb

using namespace dgsb;

extern "C" int64_t dgs_loc_table_capacity(int64_t n_unique) {
  // 2 * _UpPower(n) with _UpPower(n) = 1 << (floor(log2 n) + 1)   (hashmap.h:92-95, hashmap.cu:20)
  if (n_unique < 1) n_unique = 1;
  int64_t up = 1;
  while (up <= n_unique) up <<= 1;
  return 2 * up;
}

extern "C" int dgs_loc_table_build(void *table, int64_t capacity, int itype, int world, int rank,
                                   const void *const *dev_nids, const int64_t *counts,
                                   void *stream) {
  DGS_REQUIRE(table && dev_nids && counts, "dgs_loc_table_build: null argument");
  DGS_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0,
              "dgs_loc_table_build: capacity %lld is not a power of two", (long long)capacity);
  DGS_REQUIRE(world >= 1 && world <= DGS_MAX_DEVICES && rank >= 0 && rank < world,
              "dgs_loc_table_build: bad world/rank %d/%d", world, rank);
  // ids cached on several ranks share a slot, so sum(counts) may exceed the capacity (every GPU
  // caching the same hubs - the "selfish" policy - gives sum = world * n_unique); each single list
  // holds distinct ids and must fit on its own, the union is checked by the insert kernel
  for (int d = 0; d < world; ++d) {
    DGS_REQUIRE(counts[d] >= 0 && counts[d] <= kIdxMask, "dgs_loc_table_build: bad count");
    DGS_REQUIRE(counts[d] < capacity, "dgs_loc_table_build: the %lld ids of device %d do not fit "
                "capacity %lld", (long long)counts[d], d, (long long)capacity);
  }
  cudaStream_t st = (cudaStream_t)stream;
  int zero = 0;
  DGS_CUDA_OK(cudaMemcpyToSymbolAsync(g_loc_build_error, &zero, sizeof(int), 0,
                                      cudaMemcpyHostToDevice, st));
  DGS_CUDA_OK(cudaMemsetAsync(table, 0xFF, (size_t)capacity * sizeof(LocSlot), st));
  for (int d = 0; d < world; ++d) {
    if (counts[d] == 0) continue;
    DGS_REQUIRE(dev_nids[d] != nullptr, "dgs_loc_table_build: null id list for device %d", d);
    long long prio = (d == rank) ? world : ((d - rank + world) % world);
    long long bits = (prio << 56) | ((long long)d << kDevShift);
    int grid = grid_for(counts[d], 256, 8);
    DGS_ITYPE_SWITCH(itype, IdT, {
      loc_insert_kernel<IdT><<<grid, 256, 0, st>>>((LocSlot *)table, (uint64_t)capacity - 1,
                                                   (const IdT *)dev_nids[d], counts[d], bits);
    });
    DGS_LAUNCH_CHECK();
  }
  // build time only: one round trip for the overflow flag
  int err = 0;
  DGS_CUDA_OK(cudaMemcpyFromSymbolAsync(&err, g_loc_build_error, sizeof(int), 0,
                                        cudaMemcpyDeviceToHost, st));
  DGS_CUDA_OK(cudaStreamSynchronize(st));
  DGS_REQUIRE(err != 1, "dgs_loc_table_build: more distinct ids than the capacity %lld holds",
              (long long)capacity);
  DGS_REQUIRE(err == 0, "dgs_loc_table_build: negative node id in a cache list");
  return 0;
}

extern "C" int dgs_loc_table_lookup(const void *table, int64_t capacity, int itype,
                                    const void *nids, int64_t n, void *out_dev, void *out_idx,
                                    void *stream) {
  DGS_REQUIRE(n >= 0, "dgs_loc_table_lookup: negative n");
  if (n == 0) return 0;
  DGS_REQUIRE(table && nids && out_dev && out_idx, "dgs_loc_table_lookup: null argument");
  DGS_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0,
              "dgs_loc_table_lookup: capacity must be a power of two");
  int grid = grid_for(n, 256, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    loc_lookup_kernel<IdT><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const LocSlot *)table, (uint64_t)capacity - 1, (const IdT *)nids, n, (IdT *)out_dev,
        (IdT *)out_idx);
  });
  DGS_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgs_loc_table_unpack(const void *table, int64_t capacity, int itype, void *key,
                                    void *idx, void *devid, void *stream) {
  DGS_REQUIRE(table && key && idx && devid && capacity > 0, "dgs_loc_table_unpack: bad argument");
  int grid = grid_for(capacity, 256, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    loc_unpack_kernel<IdT><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const LocSlot *)table, capacity, (IdT *)key, (IdT *)idx, (IdT *)devid);
  });
  DGS_LAUNCH_CHECK();
  return 0;
}
