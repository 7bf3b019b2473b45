// extract.cu - feature-row gather ("extract"), the headline HBM-bound kernel.
//
// Replaces the reference's block-per-row, 4-byte-scalar copies
//   _IndexKernel / _IndexOneDimKernel / GetFeaturesCUDA      src/feature/cuda/feature_ops.cu:140-210
//   _IndexP2PCacheKernel / GetFeaturesP2PCacheCUDA           src/feature/cuda/feature_ops.cu:38-138
// with two sm_100a designs over raw row bytes:
//   algo 1  "ldg":  a CTA stages a tile of row ids in shared memory, resolves every row's source
//           pointer once (local HBM / NVLink peer shard / pinned host), then all threads copy the
//           tile as a flat stream of 128-bit vectors (4 independent loads in flight per thread,
//           L1::no_allocate on both sides, output stores fully coalesced across rows).
//   algo 2  "tma":  one warp per CTA drives a 4-stage ring of shared-memory tiles with the bulk
//           async-copy engine: lane r issues cp.async.bulk global->shared for row r (completion on
//           an mbarrier), lane 0 then writes the whole tile back with ONE cp.async.bulk
//           shared->global (the output tile is contiguous).  No register staging at all.
// The location-table probe of the cached path is fused into the row-resolve step (the reference
// runs it as a separate thrust pass that writes and re-reads a pos_list, feature_ops.cu:92-108).
#include "dgs_common.cuh"
#include "p2p_server.h"

namespace dgsb {

struct RowSource {
  // plain: table != nullptr, loc == nullptr.   cached: loc != nullptr, shards = peer table,
  // table = host fallback (may be nullptr).
  const char *table;
  const LocSlot *loc;
  uint64_t cap_mask;
  PtrTable shards;
  int mod_world;  // > 0: row n lives in shard n % mod_world at slot n / mod_world (no table)
  int peers;      // shards of other GPUs are among the sources (NVLink peer loads)
  const long long *n_dev;  // optional live row count on the device (the kernel's n is then a bound)
};

template <typename IdT>
__device__ __forceinline__ const char *resolve_row(const RowSource &src, IdT nid, int64_t row_bytes) {
  if (src.mod_world > 0) {
    const long long n = (long long)nid;
    return (const char *)src.shards.p[n % src.mod_world] + (n / src.mod_world) * row_bytes;
  }
  if (src.loc != nullptr) {
    long long v = loc_lookup(src.loc, src.cap_mask, (long long)nid);
    if (v >= 0) {
      int dev = (int)((v >> kDevShift) & 0xff);
      long long idx = v & kIdxMask;
      return (const char *)src.shards.p[dev] + idx * row_bytes;
    }
  }
  return src.table + (int64_t)nid * row_bytes;
}

// ---------------------------------------------------------------------------------------------
// algo 1: flat vector gather.  VecT = int4 (16 B), int2 (8 B) or int (4 B) ... or char.
template <typename VecT>
__device__ __forceinline__ VecT ld_stream(const VecT *p) {
  return *p;
}
template <>
__device__ __forceinline__ int4 ld_stream<int4>(const int4 *p) {
  return ld_nc_v4(p);
}
template <typename VecT>
__device__ __forceinline__ void st_stream(VecT *p, const VecT &v) {
  *p = v;
}
template <>
__device__ __forceinline__ void st_stream<int4>(int4 *p, const int4 &v) {
  st_na_v4(p, v);
}

constexpr int kGatherThreads = 256;
constexpr int kGatherRows = 64;   // rows per tile
constexpr int kGatherUnroll = 4;

template <typename IdT, typename VecT>
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(RowSource src, const IdT *__restrict__ nids, int64_t n, int64_t row_bytes,
                   uint32_t vpr /* vectors per row */, uint32_t vpr_magic, char *__restrict__ out,
                   int tile_rows /* <= kGatherRows */) {
  __shared__ const char *s_src[kGatherRows];
  if (src.n_dev != nullptr) n = min(n, (int64_t)*src.n_dev);
  const int64_t num_tiles = (n + tile_rows - 1) / tile_rows;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * tile_rows;
    const int rows = (int)min((int64_t)tile_rows, n - row0);
    __syncthreads();  // previous tile's readers are done with s_src
    if (threadIdx.x < rows) {
      IdT nid = nids[row0 + threadIdx.x];
      s_src[threadIdx.x] = resolve_row<IdT>(src, nid, row_bytes);
    }
    __syncthreads();
    const uint32_t total = (uint32_t)rows * vpr;
    VecT *otile = reinterpret_cast<VecT *>(out + row0 * row_bytes);
    for (uint32_t base = threadIdx.x; base < total; base += kGatherThreads * kGatherUnroll) {
      VecT v[kGatherUnroll];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        uint32_t e = base + u * kGatherThreads;
        if (e < total) {
          uint32_t r = vpr == 1 ? e : __umulhi(e, vpr_magic);  // e / vpr, exact for e < 2^16
          uint32_t c = e - r * vpr;
          v[u] = ld_stream<VecT>(reinterpret_cast<const VecT *>(s_src[r]) + c);
        }
      }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        uint32_t e = base + u * kGatherThreads;
        if (e < total) st_stream<VecT>(otile + e, v[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// algo 3: warp-autonomous gather.  A warp owns groups of 32 consecutive output rows: lane r
// resolves the source of row r (the ids of the NEXT group are already in flight), the pointers go
// through a per-warp shared-memory line, and the warp streams the group's rows as one flat run of
// 128-bit vectors with U independent loads per lane in flight.  No CTA barrier anywhere: a warp
// never waits for the slowest row of 7 other warps, and the id -> pointer -> row dependency of a
// group overlaps the copy of the previous one.  HINT: 0 = L1::no_allocate (as algo 1),
// 1 = evict-first (ld/st.global.cs: the gathered rows and the output have no reuse).
template <int HINT>
__device__ __forceinline__ int4 ld_row_v4(const int4 *p) {
  int4 r;
  if (HINT == 1)
    asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
  else
    r = ld_nc_v4(p);
  return r;
}
template <int HINT>
__device__ __forceinline__ void st_row_v4(int4 *p, const int4 &v) {
  if (HINT == 1)
    asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
  else
    st_na_v4(p, v);
}

constexpr int kWarpGatherWarps = 8;

template <typename IdT, int U, int HINT>
__global__ void __launch_bounds__(kWarpGatherWarps * 32)
gather_rows_warp_kernel(RowSource src, const IdT *__restrict__ nids, int64_t n, int64_t row_bytes,
                        uint32_t vpr, uint32_t vpr_magic, char *__restrict__ out) {
  __shared__ const char *s_src[kWarpGatherWarps][32];
  if (src.n_dev != nullptr) n = min(n, (int64_t)*src.n_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t groups = (n + 31) >> 5;
  const int64_t gstride = (int64_t)gridDim.x * kWarpGatherWarps;
  int64_t g = (int64_t)blockIdx.x * kWarpGatherWarps + warp;
  // (measured, not kept: giving every warp an equal share by idling the surplus warps - 0.66 vs
  // 0.71 of peak at 192 k rows)
  IdT nid = 0;
  if (g < groups && g * 32 + lane < n) nid = nids[g * 32 + lane];
  for (; g < groups; g += gstride) {
    const int64_t row0 = g << 5;
    const int rows = (int)min((int64_t)32, n - row0);
    __syncwarp();   // the previous group's readers are done with the line
    if (lane < rows) s_src[warp][lane] = resolve_row<IdT>(src, nid, row_bytes);
    const int64_t g2 = g + gstride;   // ids of the next group: in flight during this copy
    if (g2 < groups && g2 * 32 + lane < n) nid = nids[g2 * 32 + lane];
    __syncwarp();
    const uint32_t total = (uint32_t)rows * vpr;
    int4 *otile = reinterpret_cast<int4 *>(out + row0 * row_bytes);
    for (uint32_t base = lane; base < total; base += 32 * U) {
      int4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t e = base + u * 32;
        if (e < total) {
          const uint32_t r = vpr == 1 ? e : __umulhi(e, vpr_magic);  // e / vpr, exact for e < 2^16
          const uint32_t c = e - r * vpr;
          v[u] = ld_row_v4<HINT>(reinterpret_cast<const int4 *>(s_src[warp][r]) + c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t e = base + u * 32;
        if (e < total) st_row_v4<HINT>(otile + e, v[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// algo 7: row-aligned warp gather.  Same warp-autonomous structure as algo 3, but a load
// instruction never straddles two rows: the warp is cut into groups of G = 2^ceil(log2(vpr)) lanes
// (32 for 400- and 512-byte rows) and each group reads ONE row per instruction, lane c < vpr taking
// vector c.  Lanes >= vpr idle (25 of 32 active on 400-byte rows), but every 128-byte line of a
// source row is requested exactly once - in the flat scheme the line that straddles an instruction
// boundary is requested twice (ncu on NVLink peer rows: 4.66 read requests per 400-byte row, 4 is
// the minimum), and the request rate is what bounds peer reads of rows that are not line multiples.
template <typename IdT, int U>
__global__ void __launch_bounds__(kWarpGatherWarps * 32)
gather_rows_aligned_kernel(RowSource src, const IdT *__restrict__ nids, int64_t n, int64_t row_bytes,
                           uint32_t vpr, int gshift, char *__restrict__ out) {
  __shared__ const char *s_src[kWarpGatherWarps][32];
  if (src.n_dev != nullptr) n = min(n, (int64_t)*src.n_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = 1 << gshift;               // lanes per row
  const int rpi = 32 >> gshift;            // rows per instruction
  const int sub = lane >> gshift;          // which row of the instruction this lane works on
  const uint32_t c0 = (uint32_t)(lane & (G - 1));
  const int64_t groups = (n + 31) >> 5;
  const int64_t gstride = (int64_t)gridDim.x * kWarpGatherWarps;
  int64_t g = (int64_t)blockIdx.x * kWarpGatherWarps + warp;
  IdT nid = 0;
  if (g < groups && g * 32 + lane < n) nid = nids[g * 32 + lane];
  for (; g < groups; g += gstride) {
    const int64_t row0 = g << 5;
    const int rows = (int)min((int64_t)32, n - row0);
    __syncwarp();
    if (lane < rows) s_src[warp][lane] = resolve_row<IdT>(src, nid, row_bytes);
    const int64_t g2 = g + gstride;
    if (g2 < groups && g2 * 32 + lane < n) nid = nids[g2 * 32 + lane];
    __syncwarp();
    char *otile = out + row0 * row_bytes;
    for (uint32_t cb = 0; cb < vpr; cb += (uint32_t)G) {      // one pass unless a row exceeds 512 bytes
      const uint32_t c = cb + c0;
      for (int r0 = 0; r0 < rows; r0 += rpi * U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = r0 + u * rpi + sub;
          if (r < rows && c < vpr) v[u] = ld_nc_v4(reinterpret_cast<const int4 *>(s_src[warp][r]) + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = r0 + u * rpi + sub;
          if (r < rows && c < vpr) st_na_v4(reinterpret_cast<int4 *>(otile + (int64_t)r * row_bytes) + c, v[u]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// algo 2: TMA bulk-copy ring (row_bytes % 16 == 0, 16-byte aligned tables).
constexpr int kTmaRows = 32;      // rows per stage = one per lane
constexpr int kTmaPrefetch = 3;   // tiles of row loads in flight per warp
constexpr int kTmaPendingSt = 2;  // bulk stores allowed to be still reading shared memory
constexpr int kTmaStages = kTmaPrefetch + kTmaPendingSt + 1;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename IdT>
__global__ void __launch_bounds__(32)
gather_rows_tma_kernel(RowSource src, const IdT *__restrict__ nids, int64_t n, int64_t row_bytes,
                       char *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [stages][kTmaRows * row_bytes] then mbarriers
  const uint32_t stage_bytes = (uint32_t)(kTmaRows * row_bytes);
  unsigned char *bufs = smem_raw;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kTmaStages * stage_bytes);
  const int lane = threadIdx.x;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  if (src.n_dev != nullptr) n = min(n, (int64_t)*src.n_dev);
  const int64_t num_tiles = (n + kTmaRows - 1) / kTmaRows;
  // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  int64_t issue_tile = blockIdx.x;   // next tile to issue loads for
  int64_t drain_tile = blockIdx.x;   // next tile to store
  uint32_t issue_it = 0, drain_it = 0;

  auto issue = [&](int64_t tile, uint32_t it) {
    const int s = it % kTmaStages;
    const int64_t row0 = tile * kTmaRows;
    const int rows = (int)min((int64_t)kTmaRows, n - row0);
    if (it >= (uint32_t)kTmaStages) {
      // the bulk store that last read this stage must have finished reading shared memory.
      // stores are committed one group per tile, in order.  When tile `it` is issued, tiles
      // 0 .. it-prefetch-1 have been committed; allowing the newest kTmaPendingSt of them to be
      // pending guarantees the group of tile (it - stages) has finished reading its stage.
      if (lane == 0) bulk_wait_read<kTmaPendingSt>();
      __syncwarp();
    }
    const uint32_t bar = smem_u32(&bars[s]);
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(rows * row_bytes));
    __syncwarp();
    if (lane < rows) {
      IdT nid = nids[row0 + lane];
      const char *p = resolve_row<IdT>(src, nid, row_bytes);
      bulk_g2s(smem_u32(bufs + (size_t)s * stage_bytes + (size_t)lane * row_bytes), p,
               (uint32_t)row_bytes, bar);
    }
  };
  auto drain = [&](int64_t tile, uint32_t it) {
    const int s = it % kTmaStages;
    const uint32_t parity = (it / kTmaStages) & 1u;
    const int64_t row0 = tile * kTmaRows;
    const int rows = (int)min((int64_t)kTmaRows, n - row0);
    mbar_wait(smem_u32(&bars[s]), parity);
    if (lane == 0) {
      bulk_s2g(out + row0 * row_bytes, smem_u32(bufs + (size_t)s * stage_bytes),
               (uint32_t)(rows * row_bytes));
      bulk_commit();
    }
    __syncwarp();
  };

  // prologue: kTmaPrefetch tiles of loads in flight
  for (int p = 0; p < kTmaPrefetch && issue_tile < num_tiles; ++p) {
    issue(issue_tile, issue_it);
    issue_tile += gridDim.x;
    ++issue_it;
  }
  while (drain_tile < num_tiles) {
    if (issue_tile < num_tiles) {
      issue(issue_tile, issue_it);
      issue_tile += gridDim.x;
      ++issue_it;
    }
    drain(drain_tile, drain_it);
    drain_tile += gridDim.x;
    ++drain_it;
  }
  if (lane == 0) bulk_wait_all();
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Resident CTAs per SM the gather kernels ask for (default 8 = the whole SM).  A caller that runs a
// gather next to another kernel on a second stream (BatchLoader.iter_many) lowers it so that the
// other kernel finds room: an NVLink-bound gather needs ~2 CTAs per SM to keep the link busy.
static int g_gather_ctas_per_sm = 8;
static int g_gather_tile_rows = 0;   // 0 = kGatherRows (tuning knob, dgs_set_gather_tile_rows)

static inline uint32_t div_magic(uint32_t d) { return (uint32_t)(((1ull << 32) + d - 1) / d); }

template <typename IdT>
static int launch_gather(const RowSource &src, const IdT *nids, int64_t n, int64_t row_bytes,
                         char *out, int algo, bool all_aligned16, cudaStream_t st) {
  if (n == 0) return 0;
  const bool can_tma = all_aligned16 && (row_bytes % 16 == 0) && row_bytes >= 16 &&
                       (size_t)kTmaStages * kTmaRows * row_bytes + 64 <= 200 * 1024;
  if (algo == 2 && !can_tma) {
    set_error("extract: TMA path needs 16-byte aligned tables and row_bytes %% 16 == 0 "
              "(row_bytes=%lld)", (long long)row_bytes);
    return 1;
  }
  // default (measured on B200, tools/extract_probe.py): launches of >= 128 MB of rows go to the
  // warp-autonomous gather (0.82 vs 0.80 of peak at 1 M x 400 B, 0.94 vs 0.90 at 1 M x 512 B),
  // mini-batch-sized ones to the CTA-tile kernel (0.71 both at 192 k x 400 B, 0.82 vs 0.81 at 512 B)
  // Peer shards with rows that are not multiples of the 128-byte line: the row-aligned variant
  // (one read request per line; 595 vs 524 GB/s of payload over NVLink at 400-byte rows).
  if (algo == 0 && src.peers && all_aligned16 && row_bytes % 16 == 0 && row_bytes % 128 != 0 &&
      row_bytes >= 64)
    algo = 7;
  if (algo == 0)
    algo = (all_aligned16 && row_bytes % 16 == 0 && row_bytes >= 64 && row_bytes / 16 * 32 < 65536 &&
            n * row_bytes >= (128ll << 20)) ? 3 : 1;
  if (algo == 2) {
    size_t smem = (size_t)kTmaStages * kTmaRows * row_bytes + kTmaStages * sizeof(uint64_t);
    auto kern = gather_rows_tma_kernel<IdT>;
    DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int grid = grid_for(n, kTmaRows, per_sm);
    kern<<<grid, 32, smem, st>>>(src, nids, n, row_bytes, out);
    DGS_LAUNCH_CHECK();
    return 0;
  }
  if (algo == 7) {
    DGS_REQUIRE(all_aligned16 && row_bytes % 16 == 0, "extract: algo 7 needs 16-byte aligned tables and "
                "row_bytes %% 16 == 0 (row_bytes=%lld)", (long long)row_bytes);
    const uint32_t vpr = (uint32_t)(row_bytes / 16);
    int gshift = 0;
    while ((1u << gshift) < vpr && gshift < 5) ++gshift;
    const int grid = grid_for(n, 32 * kWarpGatherWarps, g_gather_ctas_per_sm);
    gather_rows_aligned_kernel<IdT, 8><<<grid, kWarpGatherWarps * 32, 0, st>>>(src, nids, n, row_bytes, vpr,
                                                                            gshift, out);
    DGS_LAUNCH_CHECK();
    return 0;
  }
  if (algo >= 3 && algo <= 6) {   // 3 = U 13; 4 / 5 / 6: probe variants (evict-first, U = 4, U = 8)
    DGS_REQUIRE(all_aligned16 && row_bytes % 16 == 0, "extract: algo %d needs 16-byte aligned tables and "
                "row_bytes %% 16 == 0 (row_bytes=%lld)", algo, (long long)row_bytes);
    const uint32_t vpr = (uint32_t)(row_bytes / 16);
    DGS_REQUIRE((uint64_t)vpr * 32 < 65536, "extract: row_bytes %lld too large for algo %d",
                (long long)row_bytes, algo);
    const int grid = grid_for(n, 32 * kWarpGatherWarps, g_gather_ctas_per_sm);
    const uint32_t magic = div_magic(vpr);
    // vectors in flight per lane (U): 13 = half of a 400-byte-row group's share per lane
#define DGS_WG(UU, HH)                                                                          \
  gather_rows_warp_kernel<IdT, UU, HH><<<grid, kWarpGatherWarps * 32, 0, st>>>(src, nids, n,   \
                                                                              row_bytes, vpr, magic, out)
    if (algo == 4) DGS_WG(8, 1);
    else if (algo == 5) DGS_WG(4, 0);
    else if (algo == 6) DGS_WG(8, 0);
    else DGS_WG(13, 0);
#undef DGS_WG
    DGS_LAUNCH_CHECK();
    return 0;
  }
  const int tile_rows = (g_gather_tile_rows >= 1 && g_gather_tile_rows <= kGatherRows) ? g_gather_tile_rows : kGatherRows;
  int grid = grid_for(n, tile_rows, g_gather_ctas_per_sm);
  if (all_aligned16 && row_bytes % 16 == 0) {
    uint32_t vpr = (uint32_t)(row_bytes / 16);
    DGS_REQUIRE((uint64_t)vpr * kGatherRows < 65536, "extract: row_bytes %lld too large for the "
                "tile index math (max %d)", (long long)row_bytes, 65535 / kGatherRows * 16);
    gather_rows_kernel<IdT, int4><<<grid, kGatherThreads, 0, st>>>(src, nids, n, row_bytes, vpr,
                                                                   div_magic(vpr), out, tile_rows);
  } else if (row_bytes % 8 == 0) {
    uint32_t vpr = (uint32_t)(row_bytes / 8);
    DGS_REQUIRE((uint64_t)vpr * kGatherRows < 65536, "extract: row_bytes %lld too large",
                (long long)row_bytes);
    gather_rows_kernel<IdT, int2><<<grid, kGatherThreads, 0, st>>>(src, nids, n, row_bytes, vpr,
                                                                   div_magic(vpr), out, tile_rows);
  } else if (row_bytes % 4 == 0) {
    uint32_t vpr = (uint32_t)(row_bytes / 4);
    DGS_REQUIRE((uint64_t)vpr * kGatherRows < 65536, "extract: row_bytes %lld too large",
                (long long)row_bytes);
    gather_rows_kernel<IdT, int><<<grid, kGatherThreads, 0, st>>>(src, nids, n, row_bytes, vpr,
                                                                  div_magic(vpr), out, tile_rows);
  } else {
    uint32_t vpr = (uint32_t)row_bytes;
    DGS_REQUIRE((uint64_t)vpr * kGatherRows < 65536, "extract: row_bytes %lld too large",
                (long long)row_bytes);
    gather_rows_kernel<IdT, char><<<grid, kGatherThreads, 0, st>>>(src, nids, n, row_bytes, vpr,
                                                                   div_magic(vpr), out, tile_rows);
  }
  DGS_LAUNCH_CHECK();
  return 0;
}

static inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace dgsb

using namespace dgsb;

extern "C" int dgs_set_gather_ctas_per_sm(int ctas) {
  DGS_REQUIRE(ctas >= 1 && ctas <= 8, "dgs_set_gather_ctas_per_sm: 1..8");
  g_gather_ctas_per_sm = ctas;
  return 0;
}

extern "C" int dgs_set_gather_tile_rows(int rows) {
  DGS_REQUIRE(rows >= 0 && rows <= kGatherRows, "dgs_set_gather_tile_rows: 0 (default) .. %d", kGatherRows);
  g_gather_tile_rows = rows;
  return 0;
}

extern "C" int dgs_index_select(const void *table, int64_t row_bytes, int itype, const void *nids,
                                int64_t n, void *out, int algo, void *stream) {
  DGS_REQUIRE(n >= 0 && row_bytes > 0, "dgs_index_select: bad sizes n=%lld row_bytes=%lld",
              (long long)n, (long long)row_bytes);
  if (n == 0) return 0;
  DGS_REQUIRE(table && nids && out, "dgs_index_select: null pointer");
  RowSource src;
  memset(&src, 0, sizeof(src));
  src.table = (const char *)table;
  bool al = aligned16(table) && aligned16(out);
  DGS_ITYPE_SWITCH(itype, IdT, {
    return launch_gather<IdT>(src, (const IdT *)nids, n, row_bytes, (char *)out, algo, al,
                              (cudaStream_t)stream);
  });
  return 0;
}

extern "C" int dgs_extract_p2p(const dgs_p2p_server_t *feat, const void *host_table,
                               int64_t row_bytes, const void *loc_table, int64_t capacity,
                               int itype, const void *nids, int64_t n, void *out, int algo,
                               void *stream) {
  DGS_REQUIRE(n >= 0 && row_bytes > 0, "dgs_extract_p2p: bad sizes");
  if (n == 0) return 0;
  DGS_REQUIRE(feat && loc_table && nids && out, "dgs_extract_p2p: null pointer");
  DGS_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0,
              "dgs_extract_p2p: capacity must be a power of two");
  RowSource src;
  memset(&src, 0, sizeof(src));
  src.table = (const char *)host_table;
  src.loc = (const LocSlot *)loc_table;
  src.cap_mask = (uint64_t)capacity - 1;
  src.peers = feat->world > 1;
  bool al = aligned16(out) && (host_table == nullptr || aligned16(host_table));
  for (int d = 0; d < feat->world; ++d) {
    src.shards.p[d] = feat->ptrs[d];
    al = al && aligned16(feat->ptrs[d]);
  }
  DGS_ITYPE_SWITCH(itype, IdT, {
    return launch_gather<IdT>(src, (const IdT *)nids, n, row_bytes, (char *)out, algo, al,
                              (cudaStream_t)stream);
  });
  return 0;
}

// Modulo-sharded gather: every row is cached, row n lives in shard n % world at slot n / world
// (the layout of the synthetic multi-GPU benchmarks) - the owner is arithmetic, so the
// location table and its random 16-byte probe per row disappear.
extern "C" int dgs_extract_sharded(const dgs_p2p_server_t *feat, int64_t row_bytes, int itype,
                                   const void *nids, int64_t n, void *out, int algo, void *stream) {
  DGS_REQUIRE(n >= 0 && row_bytes > 0, "dgs_extract_sharded: bad sizes");
  if (n == 0) return 0;
  DGS_REQUIRE(feat && nids && out, "dgs_extract_sharded: null pointer");
  RowSource src;
  memset(&src, 0, sizeof(src));
  src.mod_world = feat->world;
  src.peers = feat->world > 1;
  bool al = aligned16(out);
  for (int d = 0; d < feat->world; ++d) {
    src.shards.p[d] = feat->ptrs[d];
    al = al && aligned16(feat->ptrs[d]);
  }
  DGS_ITYPE_SWITCH(itype, IdT, {
    return launch_gather<IdT>(src, (const IdT *)nids, n, row_bytes, (char *)out, algo, al,
                              (cudaStream_t)stream);
  });
  return 0;
}

// Gather whose live row count is only known on the device (n_dev): lets the extract of a mini-batch
// be enqueued right behind the sampling kernel, before the host has read the frontier size
// (whole-batch pipeline).  n_ub bounds the grid and the output capacity.  Source:
// feat == NULL -> plain table; feat + mod_world > 0 -> modulo shards; feat + loc_table -> hash.
extern "C" int dgs_extract_dyn(const void *table, const dgs_p2p_server_t *feat, const void *loc_table,
                               int64_t capacity, int mod_world, int64_t row_bytes, int itype,
                               const void *nids, int64_t n_ub, const int64_t *n_dev, void *out,
                               int algo, void *stream) {
  DGS_REQUIRE(n_ub >= 0 && row_bytes > 0, "dgs_extract_dyn: bad sizes");
  if (n_ub == 0) return 0;
  DGS_REQUIRE(nids && out, "dgs_extract_dyn: null pointer");
  RowSource src;
  memset(&src, 0, sizeof(src));
  src.table = (const char *)table;
  src.n_dev = (const long long *)n_dev;
  bool al = aligned16(out) && (table == nullptr || aligned16(table));
  if (feat != nullptr) {
    src.peers = feat->world > 1;
    for (int d = 0; d < feat->world; ++d) {
      src.shards.p[d] = feat->ptrs[d];
      al = al && aligned16(feat->ptrs[d]);
    }
    if (mod_world > 0) {
      DGS_REQUIRE(mod_world == feat->world, "dgs_extract_dyn: mod_world must equal the p2p world size");
      src.mod_world = mod_world;
    } else {
      DGS_REQUIRE(loc_table && capacity > 0 && (capacity & (capacity - 1)) == 0,
                  "dgs_extract_dyn: cached source needs a location table with power-of-two capacity");
      src.loc = (const LocSlot *)loc_table;
      src.cap_mask = (uint64_t)capacity - 1;
    }
  } else {
    DGS_REQUIRE(table != nullptr, "dgs_extract_dyn: null table");
  }
  DGS_ITYPE_SWITCH(itype, IdT, {
    return launch_gather<IdT>(src, (const IdT *)nids, n_ub, row_bytes, (char *)out, algo, al,
                              (cudaStream_t)stream);
  });
  return 0;
}
