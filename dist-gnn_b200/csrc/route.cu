// route.cu - request routing for the id-exchange variant of the sharded extract (north star (4):
// "NCCL used only for seed/ID exchange"; SURVEY 8e: "an optional alternative to be measured
// against pure peer loads").  The reference has no counterpart: its only per-rank routing is the
// build-time all-gather-v of cache id lists (src/nccl/nccl_context.cc:65-112).
//
// dgs_route_ids partitions n requested node ids by owner (node n lives on GPU n mod world, slot
// n / world - the layout of dgs_extract_sharded) so that ONE all-to-all can carry every owner its
// requests:
//   send_idx [n]      the owners' slot numbers, grouped by owner (owner 0's requests first)
//   inv      [n]      inv[i] = position of request i inside send_idx = position of its row in the
//                     rows that come back in the same grouped order
//   counts   [world]  requests per owner (int64, device; the host reads them for the split sizes)
// Three small launches, no host round trip: per-CTA histograms -> one-CTA scan over (owner, CTA) ->
// scatter with warp-aggregated shared-memory atomics.  The order inside an owner's group is
// arbitrary (and irrelevant: inv is its exact inverse), the extract's output is not.
#include "dgs_common.cuh"

namespace dgsb {

constexpr int kRtThreads = 256;
constexpr int kRtItems = 8;                       // ids per thread
constexpr int kRtChunk = kRtThreads * kRtItems;   // ids per CTA

template <typename IdT>
__global__ void __launch_bounds__(kRtThreads)
route_count_kernel(const IdT *__restrict__ nids, int64_t n, int world, int64_t nblocks,
                   unsigned int *__restrict__ block_counts /* [world][nblocks] */) {
  __shared__ unsigned int s_cnt[DGS_MAX_DEVICES];
  if (threadIdx.x < DGS_MAX_DEVICES) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i0 = (int64_t)blockIdx.x * kRtChunk;
#pragma unroll
  for (int u = 0; u < kRtItems; ++u) {
    const int64_t i = i0 + u * kRtThreads + threadIdx.x;
    const bool live = i < n;
    const int o = live ? (int)((unsigned long long)nids[i] % (unsigned int)world) : -1;
    // warp-aggregated: one shared-memory atomic per (warp, owner present in the warp)
    const unsigned int peers = __match_any_sync(0xffffffffu, o);
    if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[o], __popc(peers));
  }
  __syncthreads();
  if ((int)threadIdx.x < world) block_counts[(int64_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of block_counts in (owner-major, CTA-minor) order, in place; totals per owner
__global__ void __launch_bounds__(1024)
route_scan_kernel(unsigned int *__restrict__ block_counts, int64_t total_entries, int64_t nblocks,
                  int world, long long *__restrict__ counts) {
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ long long s_owner_start[DGS_MAX_DEVICES + 1];
  long long carry = 0;
  for (int64_t base = 0; base < total_entries; base += 1024) {
    const int64_t e = base + threadIdx.x;
    const long long v = e < total_entries ? (long long)block_counts[e] : 0;
    const long long ex = carry + block_exclusive_scan<long long>(v, s_scan, &s_total);
    if (e < total_entries) {
      block_counts[e] = (unsigned int)ex;
      if (e % nblocks == 0) s_owner_start[e / nblocks] = ex;
    }
    carry += s_total;
  }
  if (threadIdx.x == 0) s_owner_start[world] = carry;
  __syncthreads();
  if ((int)threadIdx.x < world) counts[threadIdx.x] = s_owner_start[threadIdx.x + 1] - s_owner_start[threadIdx.x];
}

template <typename IdT>
__global__ void __launch_bounds__(kRtThreads)
route_scatter_kernel(const IdT *__restrict__ nids, int64_t n, int world, int64_t nblocks,
                     const unsigned int *__restrict__ block_offsets, IdT *__restrict__ send_idx,
                     IdT *__restrict__ inv) {
  __shared__ unsigned int s_base[DGS_MAX_DEVICES];
  if ((int)threadIdx.x < world) s_base[threadIdx.x] = block_offsets[(int64_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  const int64_t i0 = (int64_t)blockIdx.x * kRtChunk;
  IdT id[kRtItems];
#pragma unroll
  for (int u = 0; u < kRtItems; ++u) {
    const int64_t i = i0 + u * kRtThreads + threadIdx.x;
    id[u] = i < n ? nids[i] : (IdT)0;
  }
#pragma unroll
  for (int u = 0; u < kRtItems; ++u) {
    const int64_t i = i0 + u * kRtThreads + threadIdx.x;
    const bool live = i < n;
    const unsigned long long v = (unsigned long long)id[u];
    const int o = live ? (int)(v % (unsigned int)world) : -1;
    const unsigned int peers = __match_any_sync(0xffffffffu, o);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned int base = 0;
    if (live && lane == leader) base = atomicAdd(&s_base[o], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) {
      const unsigned int pos = base + __popc(peers & ((1u << lane) - 1u));
      send_idx[pos] = (IdT)(v / (unsigned int)world);
      inv[i] = (IdT)pos;
    }
  }
}

}  // namespace dgsb

using namespace dgsb;

static int64_t route_blocks(int64_t n) { return n <= 0 ? 1 : (n + kRtChunk - 1) / kRtChunk; }

extern "C" int64_t dgs_route_ws_bytes(int64_t n, int world) {
  if (n < 0 || world < 1 || world > DGS_MAX_DEVICES) return -1;
  return route_blocks(n) * world * (int64_t)sizeof(unsigned int);
}

extern "C" int dgs_route_ids(int itype, const void *nids, int64_t n, int world, void *send_idx,
                             void *inv, int64_t *counts_dev, void *ws, int64_t ws_bytes,
                             void *stream) {
  DGS_REQUIRE(world >= 1 && world <= DGS_MAX_DEVICES, "dgs_route_ids: world %d not in 1..%d", world,
              DGS_MAX_DEVICES);
  DGS_REQUIRE(n >= 0 && n < (1ll << 32) - kRtChunk, "dgs_route_ids: bad request count %lld", (long long)n);
  DGS_REQUIRE(counts_dev && ws, "dgs_route_ids: null argument");
  DGS_REQUIRE(n == 0 || (nids && send_idx && inv), "dgs_route_ids: null argument");
  DGS_REQUIRE(itype != DGS_I32 || n < (1ll << 31), "dgs_route_ids: positions do not fit int32 ids");
  DGS_REQUIRE(ws_bytes >= dgs_route_ws_bytes(n, world), "dgs_route_ids: workspace %lld < %lld bytes",
              (long long)ws_bytes, (long long)dgs_route_ws_bytes(n, world));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = route_blocks(n);
  unsigned int *bc = (unsigned int *)ws;
  DGS_ITYPE_SWITCH(itype, IdT, {
    route_count_kernel<IdT><<<(unsigned int)nblocks, kRtThreads, 0, st>>>((const IdT *)nids, n, world, nblocks, bc);
    DGS_LAUNCH_CHECK();
    route_scan_kernel<<<1, 1024, 0, st>>>(bc, nblocks * world, nblocks, world, (long long *)counts_dev);
    DGS_LAUNCH_CHECK();
    if (n > 0) {
      route_scatter_kernel<IdT><<<(unsigned int)nblocks, kRtThreads, 0, st>>>(
          (const IdT *)nids, n, world, nblocks, bc, (IdT *)send_idx, (IdT *)inv);
      DGS_LAUNCH_CHECK();
    }
  });
  return 0;
}
