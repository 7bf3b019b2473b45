// blocks.cu - the fused whole-batch pipeline: L hops of (sample + relabel) in ONE cooperative
// kernel launch (or 3 launches per hop when the fan-out is too large for the tile phases), with no
// host round trip until the caller reads the 2 L hop sizes.
//
// Replaces the layer loops P2PCacheNodeClassificationSample{Uniform,Bias}
// (src/sampling/sampler.cc:14-62) which, per hop, run ~15 launches (thrust lookup, 2 cub scans,
// the sampling kernel, 2 torch::cat, 3 torch::full of the hash size, 4 thrust relabel passes,
// 2 more scan kernels) and block twice on a D2H read (rowwise_sampling_p2p.cu:226-228,
// tensor_relabel.cu:129).  At batch 1024 every one of those kernels is a few microseconds, so the
// reference's hop is launch- and sync-bound (measured on B200: 0.70 ms per 3-hop batch vs 0.14 ms
// here, host sync included); a hop here is three dependent phases separated by grid barriers:
//
//   pick  : CTA per <= 128 seeds.  Location probe, indptr pair from the owner (local HBM / NVLink
//           peer / pinned host), selection (sampling_device.cuh), neighbours written to a PADDED
//           slot array (seed i owns slots [i k, (i+1) k)) - so no prefix sum is needed before
//           sampling - and every seed / neighbour id is inserted on the fly into the hop's
//           relabel table: atomicMin of the item index = first occurrence, at slot = node id when
//           the node count is known (direct table, struct Tab), else after a CAS on the key of a
//           hashed slot.  Idle CTAs wipe the slots the previous hop touched in the other table
//           (two alternate), so no memset ever runs.
//   rank  : CTA per 128 seeds.  Flags first occurrences among the seeds (A) and among the sampled
//           neighbours (B), counts the edges (C), block scans -> per-tile totals.  The totals
//           become exclusive prefixes in every CTA's shared memory at the start of the emit phase
//           (cooperative kernel) or through the last CTA to finish (multi-kernel path); nnz = C
//           and |frontier| = A + B are published on the device.
//   emit  : thread per padded slot.  frontier[new id] = id for first occurrences, and the COO is
//           written compacted and relabelled: row = new id of the seed, col = new id of the
//           neighbour (new id = tile prefix + rank inside the tile).
// The result is bit-identical to sample -> TensorRelabelCUDA({seeds, col}, {row, col}): `frontier`
// is the first-occurrence-order unique of cat(seeds, coo_col), the COO is seed-major with the
// neighbours of a seed in selection order (CSR order on the copy path).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "dgs_common.cuh"
#include "p2p_server.h"
#include "sampling_device.cuh"

namespace cg = cooperative_groups;

namespace dgsb {

#ifndef DGS_COOP_MIN_CTAS
#define DGS_COOP_MIN_CTAS 2  // measured on B200: 2 CTAs/SM (128 regs, no spills) beats 3 and 4
#endif

constexpr int kBkWarps = 8;
constexpr int kBkThreads = 256;
constexpr int kBkTile = 128;  // seeds per rank tile

struct __align__(16) RlSlot {
  long long key;       // -1 empty
  unsigned int first;  // smallest item index holding this key
  unsigned int lrank;  // rank of that first occurrence inside its tile
};

// One relabel table.  Hashed (direct == 0): open addressing over 16-byte RlSlot, any id.  Direct
// (direct == 1): the caller told us the number of nodes, so the slot of an id IS the id - 8 bytes
// {first, lrank} per node, no key, no probing, and an insert is ONE fire-and-forget atomicMin
// instead of a CAS whose result the thread has to wait for plus an atomicMin (measured on B200,
// products shape: 93 vs 107 us per batch).  180 GB of HBM pay for 16 bytes per node.
// direct == 2 (multi-batch kernel): 4 bytes per node, {tag:8 | first item:24} only - the rank of a
// first occurrence lives in a per-item side array instead (HopState::lrank_at), which halves the
// footprint the random atomics / reads of a hop cycle through L2.
struct Tab {
  char *base;
  int direct;
  __device__ __forceinline__ unsigned int *first(uint64_t pos) const {
    return reinterpret_cast<unsigned int *>(base + (direct == 2 ? pos * 4 : (direct ? pos * 8 : pos * 16 + 8)));
  }
  __device__ __forceinline__ unsigned int *lrank(uint64_t pos) const { return first(pos) + 1; }
  __device__ __forceinline__ unsigned long long *key(uint64_t pos) const {
    return reinterpret_cast<unsigned long long *>(base + pos * 16);
  }
  __device__ __forceinline__ int2 first_lrank(uint64_t pos) const {   // L2 load of {first, lrank}
    return __ldcg(reinterpret_cast<const int2 *>(first(pos)));
  }
  __device__ __forceinline__ void wipe(uint64_t pos) const {
    if (direct)
      *reinterpret_cast<long long *>(base + pos * 8) = -1;
    else
      *reinterpret_cast<int4 *>(base + pos * 16) = make_int4(-1, -1, -1, -1);
  }
};

struct HopState {        // per-hop arrays that must survive until the next hop's cleanup
  int *cnt;              // [S_max]   edges kept for seed i
  unsigned int *pos_seed;  // [S_max]   table slot of seed i
  unsigned int *pos_col;   // [E_max]   table slot of padded neighbour slot e
  Tab table;
  // compact tables (Tab::direct == 2) only: what the rank phase learns per item, so that the emit
  // phase never goes back to the table
  unsigned int *fslot;     // [E_max]          first-occurrence item of the id in padded slot e
  unsigned int *fseed;     // [S_max]          same for seed i (hop 0 only: later seeds are distinct)
  unsigned int *lrank_at;  // [S_max + E_max]  rank inside its tile of the first occurrence at item x
};

// Weighted sampling: rows longer than kHubDeg weights ("hubs") that a pick tile defers to the
// grid-wide hub phase of the multi-batch kernel (one entry per deferred seed).
struct __align__(16) HubEnt {
  const void *row;       // the seed's neighbour ids
  const float *w;        // ... and weights
  unsigned int i;        // seed index inside its batch
  int deg;
};
constexpr int kHubChunk = 4096;     // weights per chunk = one 512-weight pass for each of the 8 warps
constexpr int kHubMaxRows = 1024;   // rows / chunks / chunks per row the chunked hub phase handles
constexpr int kHubMaxChunks = 4096; //   (beyond: one CTA per row, hub_rows)
constexpr int kHubMaxPerRow = 32;
struct HubList {
  unsigned int *count;   // entries pushed this hop (reset at the start of the rank phase)
  HubEnt *ent;           // [S_max]
  unsigned int *done;    // [kHubMaxRows] chunks of row j finished (self-resetting)
  float *ckey;           // [kHubMaxChunks][32] candidates of a chunk, best first
  int *cidx;             // [kHubMaxChunks][32]
  int *cnum;             // [kHubMaxChunks]
};

struct BlocksWs {
  unsigned int *done;
  HubList hubs;
  long long *pending_S;              // live seed count of the hop whose table is still dirty
  long long *prefA, *prefB, *prefC;  // [tiles_max + 1]
  int *loff;                         // [S_max] exclusive edge offset of seed i inside its tile
  void *pad_col;                     // [E_max] ids
  HopState hop[2];
  int64_t cap;                       // slots per table
  int direct;                        // tables are direct-addressed (see Tab)
};

struct BlocksPlan {
  int64_t S_max, E_max, tiles_max, cap, bytes, table_bytes;
  int direct;
};

static inline int64_t up256(int64_t x) { return (x + 255) / 256 * 256; }

static int blocks_plan(int itype, int64_t num_seeds, int L, const int64_t *fan_out, int64_t num_nodes,
                       BlocksPlan *p, char *base, BlocksWs *ws) {
  DGS_REQUIRE(L >= 1 && L <= 16, "sample_blocks: 1..16 layers supported");
  int64_t ub = num_seeds < 1 ? 1 : num_seeds, S_max = 1, E_max = 1, items_max = 1;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[L - 1 - l];
    DGS_REQUIRE(k >= 0, "sample_blocks: fan_out must be >= 0 (use the per-hop entry for -1)");
    const int64_t e = ub * k;
    DGS_REQUIRE(ub + e < (1ll << 32) - 2, "sample_blocks: more than 2^32 items in one hop");
    if (ub > S_max) S_max = ub;
    if (e > E_max) E_max = e;
    if (ub + e > items_max) items_max = ub + e;
    ub += e;
  }
  int64_t cap = 64;
  while (cap < 2 * items_max) cap <<= 1;
  // direct addressing when the node count is known, ids fit the 32-bit slot arrays and two tables
  // of 8 bytes per node stay within 8 GiB
  static const bool no_direct = getenv("DGS_BLOCKS_HASHED") != nullptr;
  const bool direct = !no_direct && num_nodes > 0 && num_nodes < (1ll << 32) - 2 && num_nodes <= (1ll << 29);
  if (direct) cap = (num_nodes + 1) & ~1ll;   // even: the linear wipe stores 16 bytes at a time
  p->direct = direct ? 1 : 0;
  p->table_bytes = cap * (direct ? 8 : (int64_t)sizeof(RlSlot));
  const int idb = itype == DGS_I64 ? 8 : 4;
  p->S_max = S_max;
  p->E_max = E_max;
  p->tiles_max = (S_max + kBkTile - 1) / kBkTile;
  p->cap = cap;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char *q = base ? base + off : nullptr;
    off += up256(bytes);
    return q;
  };
  char *done = take(256);
  char *pa = take((p->tiles_max + 1) * 8), *pb = take((p->tiles_max + 1) * 8),
       *pc = take((p->tiles_max + 1) * 8);
  char *loff = take(S_max * 4);
  char *pad = take(E_max * idb);
  char *h[2][4];
  for (int b = 0; b < 2; ++b) {
    h[b][0] = take(S_max * 4);
    h[b][1] = take(S_max * 4);
    h[b][2] = take(E_max * 4);
    h[b][3] = take(p->table_bytes);
  }
  p->bytes = off;
  if (ws) {
    memset(ws, 0, sizeof(*ws));
    ws->done = (unsigned int *)done;
    ws->pending_S = (long long *)(done + 64);
    ws->prefA = (long long *)pa;
    ws->prefB = (long long *)pb;
    ws->prefC = (long long *)pc;
    ws->loff = (int *)loff;
    ws->pad_col = pad;
    for (int b = 0; b < 2; ++b) {
      ws->hop[b].cnt = (int *)h[b][0];
      ws->hop[b].pos_seed = (unsigned int *)h[b][1];
      ws->hop[b].pos_col = (unsigned int *)h[b][2];
      ws->hop[b].table.base = h[b][3];
      ws->hop[b].table.direct = p->direct;
    }
    ws->cap = cap;
    ws->direct = p->direct;
  }
  return 0;
}

__device__ __forceinline__ unsigned int rl_insert(const Tab &table, uint64_t mask, long long key,
                                                  unsigned int item) {
  if (table.direct) {
    atomicMin(table.first((uint64_t)key), item);
    return (unsigned int)key;
  }
  uint64_t pos = mix64((uint64_t)key) & mask;
  while (true) {
    unsigned long long prev =
        atomicCAS(table.key(pos), (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == (unsigned long long)kEmptyKey || prev == (unsigned long long)key) break;
    pos = (pos + 1) & mask;
  }
  atomicMin(table.first(pos), item);
  return (unsigned int)pos;
}

// N independent inserts with all first-probe CAS operations in flight before any result is used
// (a single rl_insert is a dependent CAS -> atomicMin chain of ~1 us).  Measured alternative: a
// plain 16-byte load per probe first and atomics only when they can change something (about half
// of a hop's sampled ids are duplicates) - fewer atomics but two dependent round trips for every
// first insert; slower overall on B200 (118 vs 103 us per batch), so not used.
template <int N>
__device__ __forceinline__ void rl_insert_batch(const Tab &table, uint64_t mask, const long long (&key)[N],
                                                const unsigned int (&item)[N], const bool (&ok)[N],
                                                unsigned int (&out_pos)[N]) {
  if (table.direct) {
#pragma unroll
    for (int u = 0; u < N; ++u) {
      if (ok[u]) {
        atomicMin(table.first((uint64_t)key[u]), item[u]);
        out_pos[u] = (unsigned int)key[u];
      }
    }
    return;
  }
  uint64_t pos[N];
  unsigned long long prev[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    pos[u] = mix64((uint64_t)key[u]) & mask;
    if (ok[u])
      prev[u] = atomicCAS(table.key(pos[u]), (unsigned long long)kEmptyKey, (unsigned long long)key[u]);
  }
#pragma unroll
  for (int u = 0; u < N; ++u) {
    if (!ok[u]) continue;
    while (prev[u] != (unsigned long long)kEmptyKey && prev[u] != (unsigned long long)key[u]) {
      pos[u] = (pos[u] + 1) & mask;
      prev[u] = atomicCAS(table.key(pos[u]), (unsigned long long)kEmptyKey, (unsigned long long)key[u]);
    }
    atomicMin(table.first(pos[u]), item[u]);
    out_pos[u] = (unsigned int)pos[u];
  }
}

__device__ __forceinline__ void rl_wipe(const Tab &table, unsigned int pos) { table.wipe(pos); }

// Wipe every slot the given hop touched (idempotent; duplicates wipe the same slot twice).
// busy_ctas: CTAs [0, busy_ctas) have sampling work of their own in this phase; when enough idle
// CTAs exist and the wipe is small it is left to them alone (it then overlaps the sampling chains
// entirely).  A hop that touched a large part of the table is wiped linearly instead (coalesced
// 16-byte stores over the whole table beat millions of scattered ones).
__device__ __forceinline__ void wipe_hop(const HopState &h, int64_t S, int k, int64_t busy_ctas,
                                         int64_t cap) {
  const int64_t items = S * (1 + (int64_t)k);
  if (items == 0) return;
  int64_t vgrid = gridDim.x, vbid = blockIdx.x;
  const int64_t idle = (int64_t)gridDim.x - busy_ctas;
  if (idle >= 64 && items <= idle * blockDim.x * 8) {
    if ((int64_t)blockIdx.x < busy_ctas) return;
    vgrid = idle;
    vbid = (int64_t)blockIdx.x - busy_ctas;
  }
  const int64_t stride = vgrid * blockDim.x;
  const int64_t tid = vbid * blockDim.x + threadIdx.x;
  // linear wipe = cap (hashed) or cap / 2 (direct) coalesced 16-byte stores; scattered = one
  // store per item at roughly 8x the cost of a coalesced one
  const int64_t n16 = h.table.direct ? cap / 2 : cap;
  if (items * 8 > n16) {
    int4 *t = reinterpret_cast<int4 *>(h.table.base);
    const int4 e = make_int4(-1, -1, -1, -1);
    for (int64_t i = tid; i < n16; i += stride) t[i] = e;
    return;
  }
  for (int64_t i = tid; i < S; i += stride) rl_wipe(h.table, __ldcg(h.pos_seed + i));
  if (k > 0) {
    const int64_t E = S * k;
    for (int64_t e0 = tid; e0 < E; e0 += 4 * stride) {
      unsigned int p[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // 4 independent (count, slot) load pairs in flight
        const int64_t e = e0 + u * stride;
        ok[u] = false;
        if (e < E) {
          const int64_t i = e / k;
          ok[u] = (int)(e - i * k) < __ldcg(h.cnt + i);
          p[u] = __ldcg(h.pos_col + e);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) rl_wipe(h.table, p[u]);
    }
  }
}

template <typename IdT>
struct PadEmit {
  IdT *pcol;
  unsigned int *ppos;
  Tab table;
  uint64_t mask;
  unsigned int item_base;
  __device__ __forceinline__ void operator()(int j, IdT v) {
    pcol[j] = v;
    ppos[j] = rl_insert(table, mask, (long long)v, item_base + (unsigned int)j);
  }
};

template <typename IdT, typename ET, int MODE>
__global__ void __launch_bounds__(kBkThreads)
fused_pick_kernel(GraphSrc g, const IdT *__restrict__ seeds, int64_t S_ub,
                  const int64_t *__restrict__ S_dev, int k, uint64_t rng_key, IdT *pad_col,
                  HopState cur, uint64_t cap_mask, HopState prev, int64_t prev_S_ub,
                  const long long *__restrict__ prev_S_dev, int prev_k, int gmem_scratch) {
  extern __shared__ __align__(16) unsigned char pick_smem[];
  const long long pS_live = prev.table.base != nullptr ? *prev_S_dev : 0;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t S = S_dev ? min(*S_dev, S_ub) : S_ub;
  const int64_t warps_total = (int64_t)gridDim.x * kBkWarps;
  int *w_idx = nullptr;
  float *w_key = nullptr;
  if (k > 0 && !gmem_scratch) {
    if (MODE == kUniform) {
      w_idx = reinterpret_cast<int *>(pick_smem) + (size_t)warp * k;
    } else if (MODE == kBias) {
      w_key = reinterpret_cast<float *>(pick_smem) + (size_t)warp * 2 * k;
      w_idx = reinterpret_cast<int *>(w_key + k);
    }
  }
  for (int64_t i = (int64_t)blockIdx.x * kBkWarps + warp; i < S; i += warps_total) {
    const long long nid = (long long)seeds[i];
    int dev;
    long long begin, deg64;
    resolve_seed<ET>(g, nid, &dev, &begin, &deg64);
    const int deg = (int)deg64;
    const bool with_replace = (MODE == kUniformReplace || MODE == kBiasReplace);
    const bool copy_path = !with_replace && deg <= k;
    const int cnt = deg == 0 ? 0 : (copy_path ? deg : k);
    if (lane == 0) {
      cur.cnt[i] = cnt;
      cur.pos_seed[i] = rl_insert(cur.table, cap_mask, nid, (unsigned int)i);
    }
    if (cnt == 0) continue;
    const IdT *row =
        reinterpret_cast<const IdT *>(dev < 0 ? g.indices : g.sh_indices.p[dev]) + begin;
    const float *wrow = nullptr;
    if (MODE == kBias || MODE == kBiasReplace)
      wrow = (dev < 0 ? g.probs : reinterpret_cast<const float *>(g.sh_probs.p[dev])) + begin;
    PadEmit<IdT> emit{pad_col + i * k, cur.pos_col + i * k, cur.table, cap_mask,
                      (unsigned int)(S_ub + i * k)};
    if (gmem_scratch && !copy_path) {
      w_idx = reinterpret_cast<int *>(emit.pcol);
      w_key = reinterpret_cast<float *>(emit.ppos);
    }
    warp_select<IdT, MODE, PadEmit<IdT>>(row, wrow, deg, k, copy_path, rng_key, (uint64_t)i, lane,
                                        w_idx, w_key, emit);
  }
  // wipe the other table: the slots the previous hop touched
  if (prev.table.base != nullptr) wipe_hop(prev, min((int64_t)pS_live, prev_S_ub), prev_k, gridDim.x, (int64_t)cap_mask + 1);
}


// ---------------------------------------------------------------------------------------------
// Phase functions.  Each is written as a loop over tiles / slots strided by the grid, so the same
// code runs as a stand-alone kernel (multi-kernel path) or as one phase of the cooperative
// whole-batch kernel (phases separated by grid barriers).  Data produced by other CTAs in an
// earlier phase is read with ld.global.cg (L2), never through a possibly stale L1 line.
constexpr int kPkSeeds = 128;  // seeds per pick tile (256 was slower on B200: 4 B2 rounds per tile)
constexpr int kPkBatch = 4;    // padded slots per thread and pass in the pick phase (8: no gain)
#ifndef DGS_EM_BATCH
#define DGS_EM_BATCH 4
#endif
constexpr int kEmBatch = DGS_EM_BATCH;    // padded slots per thread and pass in the emit phase
constexpr int kRkItems = 8;    // padded slots per thread and pass in the rank phase
constexpr int kHubDeg = 512;   // biased sampling: rows longer than this are scanned by the whole CTA (measured: 2048 slower)

// Seeds per pick tile: up to 128, fewer when the hop is small so that every CTA of the grid gets
// a tile (the phases are latency chains - more CTAs in flight, not longer chains per CTA).
__device__ __forceinline__ int pick_tile_seeds(int64_t S) {
  int64_t per = (S + gridDim.x - 1) / gridDim.x;
  per = (per + 7) & ~7ll;
  return (int)max((int64_t)8, min((int64_t)kPkSeeds, per));
}

template <typename T>
__device__ __forceinline__ T ldcg(const T *p) {
  return __ldcg(p);
}
template <typename IdT>
struct PosEmit {
  unsigned int *p;
  __device__ __forceinline__ void operator()(int j, IdT v) { p[j] = (unsigned int)v; }
};

// Pick phase, tile version: a CTA owns up to 128 seeds.
//   A  thread per seed : probe, indptr pair, count, insert the seed id        (128 chains in flight)
//   B1 selection -> POSITIONS inside the row, kept in shared memory: Floyd's subset sampling run
//      by one thread per seed for uniform sampling (no memory traffic at all), a warp per seed
//      for the weight scans of biased sampling
//   B2 thread per slot : neighbour load, padded store, table insert - 8 independent
//      load -> CAS chains in flight per thread
// Same RNG counters as the warp-per-seed kernel => identical samples.
// Floyd's subset sampling entirely in registers (fully unrolled, no shared-memory round trips):
// draw t picks r in [0, deg-k+t], or deg-k+t itself when r was already picked.  R = register slots
// (k <= R); the positions go to dst[0..k).
template <int R>
__device__ __forceinline__ void floyd_in_registers(uint64_t rng_key, uint64_t item, int deg, int k,
                                                   unsigned int *__restrict__ dst) {
  unsigned int P[R];
#pragma unroll
  for (int q = 0; q < R / 4; ++q) {
    if (4 * q < k) {
      const uint4 r4 = Philox::gen(rng_key, item, (uint64_t)q);
      P[4 * q] = r4.x; P[4 * q + 1] = r4.y; P[4 * q + 2] = r4.z; P[4 * q + 3] = r4.w;
    }
  }
#pragma unroll
  for (int t = 0; t < R; ++t) {
    if (t < k) {
      const unsigned int J = (unsigned int)(deg - k + t);
      const unsigned int r = rand_below(P[t], J + 1);
      bool dup = false;
#pragma unroll
      for (int q = 0; q < t; ++q) dup |= (P[q] == r);
      P[t] = dup ? J : r;
    }
  }
#pragma unroll
  for (int t = 0; t < R; ++t)
    if (t < k) dst[t] = P[t];
}

// kNoPos (direct tables only): the table slot of an id IS the id, so the slot arrays pos_seed /
// pos_col are neither written here nor read by the later phases.  ts = seeds per tile; the CTA owns
// tiles tile0, tile0 + tstride, ...  tagbits: OR-ed into every item index stored in the table
// (epoch tag of the multi-batch kernel's never-wiped tables; 0 elsewhere).  FR: fan-outs up to FR
// run Floyd's sampling in registers (32 costs the 128-register cooperative kernel spills, measured
// +5 us per batch at k <= 16, so only the stand-alone pick kernel of large fan-outs uses it).
template <typename IdT, typename ET, int MODE, bool kNoPos = false, int FR = 16>
__device__ __forceinline__ void pick_tile_phase(const GraphSrc &g, const IdT *__restrict__ seeds,
                                                int64_t S_ub, int64_t S, int k, uint64_t rng_key,
                                                IdT *__restrict__ pad_col, const HopState &cur,
                                                uint64_t cap_mask, int ts, int64_t tile0,
                                                int64_t tstride, unsigned int tagbits,
                                                unsigned long long *fine = nullptr,
                                                const HubList *hubs = nullptr) {
  extern __shared__ __align__(16) unsigned char pick_smem[];
  auto fstamp = [&](int slot) {
    if (fine != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      fine[slot] = t;
    }
  };
  fstamp(0);
  __shared__ const IdT *s_row[kPkSeeds];
  __shared__ const float *s_w[kPkSeeds];
  __shared__ int s_deg[kPkSeeds];
  __shared__ int s_cnt[kPkSeeds];
  __shared__ int s_next;                       // next seed of the tile to hand to a warp
  __shared__ int s_mcount[kBkWarps];           // hub rows: finalists per warp
  unsigned int *s_pick = reinterpret_cast<unsigned int *>(pick_smem);       // [kPkSeeds * k]
  float *s_key = reinterpret_cast<float *>(s_pick + (size_t)kPkSeeds * k);  // [warps * k] (kBias)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tiles = (S + ts - 1) / ts;
  const bool with_replace = (MODE == kUniformReplace || MODE == kBiasReplace);
  for (int64_t tile = tile0; tile < tiles; tile += tstride) {
    const int64_t i0 = tile * ts;
    const int ns = (int)min((int64_t)ts, S - i0);
    __syncthreads();  // previous tile's readers are done with the shared arrays
    long long seed_nid = 0;
    uint64_t seed_pos = 0;
    unsigned long long seed_prev = 0;
    if (tid < ns) {
      const int64_t i = i0 + tid;
      const long long nid = (long long)ldcg(seeds + i);
      // first probe of the seed's own table insert: issued now, resolved after the neighbours'
      // (its round trip overlaps the indptr loads, the selection and the row loads)
      seed_nid = nid;
      if (cur.table.direct) {
        seed_pos = (uint64_t)nid;
        seed_prev = (unsigned long long)nid;
      } else {
        seed_pos = mix64((uint64_t)nid) & cap_mask;
        seed_prev = atomicCAS(cur.table.key(seed_pos), (unsigned long long)kEmptyKey,
                              (unsigned long long)nid);
      }
      int dev;
      long long begin, deg64;
      resolve_seed<ET>(g, nid, &dev, &begin, &deg64);
      const int deg = (int)deg64;
      const int cnt = deg == 0 ? 0 : ((!with_replace && deg <= k) ? deg : k);
      s_row[tid] = reinterpret_cast<const IdT *>(dev < 0 ? g.indices : g.sh_indices.p[dev]) + begin;
      if (MODE == kBias || MODE == kBiasReplace)
        s_w[tid] = (dev < 0 ? g.probs : reinterpret_cast<const float *>(g.sh_probs.p[dev])) + begin;
      s_deg[tid] = deg;
      s_cnt[tid] = cnt;
      cur.cnt[i] = cnt;
      if (MODE == kUniform && k <= FR && deg > k) {
        unsigned int *dst = s_pick + (size_t)tid * k;
        if (FR <= 16 || k <= 16)
          floyd_in_registers<16>(rng_key, (uint64_t)i, deg, k, dst);
        else
          floyd_in_registers<32>(rng_key, (uint64_t)i, deg, k, dst);   // friendster's k = 20
      }
    }
    if (tile == tile0) fstamp(1);
    if (MODE == kUniform && k > FR) {
      // random words of Floyd's draws, computed by all threads (one Philox block = 4 draws)
      const int blocks_per_seed = (k + 3) >> 2;
      for (int b = tid; b < ns * blocks_per_seed; b += kBkThreads) {
        const int si = b / blocks_per_seed, q = b - si * blocks_per_seed;
        const uint4 r4 = Philox::gen(rng_key, (uint64_t)(i0 + si), (uint64_t)q);
        unsigned int *P = s_pick + (size_t)si * k + 4 * q;
        const int left = k - 4 * q;
        P[0] = r4.x;
        if (left > 1) P[1] = r4.y;
        if (left > 2) P[2] = r4.z;
        if (left > 3) P[3] = r4.w;
      }
    }
    if (MODE != kUniform || k > FR) __syncthreads();
    if (MODE == kUniform && k > FR) {
      // Floyd's subset sampling, one THREAD per seed (O(k^2) compares against shared memory)
      if (tid < ns) {
        const int deg = s_deg[tid];
        if (deg > k) {
          unsigned int *P = s_pick + (size_t)tid * k;
          for (int t = 0; t < k; ++t) {
            const unsigned int J = (unsigned int)(deg - k + t);
            const unsigned int r = rand_below(P[t], J + 1);
            bool dup = false;
            for (int q = 0; q < t; ++q) dup |= (P[q] == r);
            P[t] = dup ? J : r;
          }
        }
      }
    } else if (MODE == kBias || MODE == kBiasReplace) {
      // The weight scan of a row costs its degree, and the degrees of sampled neighbours are
      // size-biased (hubs everywhere from the second hop on).  Rows up to kHubDeg go one per
      // warp, handed out dynamically; a longer row is shared by all warps of the CTA.
      const bool hubs_shared = MODE == kBias && k <= 32;
      if (tid == 0) s_next = kBkWarps;
      __syncthreads();
      for (int s_ = warp; s_ < ns;) {
        const int deg = s_deg[s_];
        const bool skip = deg == 0 || (!with_replace && deg <= k) || (hubs_shared && deg > kHubDeg);
        if (!skip) {   // (copy path: position j, nothing to select)
          PosEmit<IdT> pe{s_pick + (size_t)s_ * k};
          warp_select<IdT, MODE, PosEmit<IdT>, true>(nullptr, s_w[s_], deg, k, false, rng_key,
                                                     (uint64_t)(i0 + s_), lane,
                                                     reinterpret_cast<int *>(pe.p),
                                                     s_key + (size_t)warp * k, pe);
        }
        int nxt = 0;
        if (lane == 0) nxt = atomicAdd(&s_next, 1);
        s_ = __shfl_sync(0xffffffffu, nxt, 0);
      }
      if (hubs_shared && hubs != nullptr) {
        // multi-batch kernel: hub rows are not scanned here - a tile that happens to own a few of
        // them would keep the whole grid waiting at the next barrier (measured: 9-30 us of a
        // 106 us batch).  They go to a grid-wide list that ALL CTAs drain in the hub phase.
        if (tid < ns) {
          const int deg = s_deg[tid];
          if (deg > kHubDeg && deg > k) {
            const unsigned int at = atomicAdd(hubs->count, 1u);
            HubEnt e;
            e.row = s_row[tid];
            e.w = s_w[tid];
            e.i = (unsigned int)(i0 + tid);
            e.deg = deg;
            hubs->ent[at] = e;
          }
        }
      } else if (hubs_shared) {
        for (int s_ = 0; s_ < ns; ++s_) {
          const int deg = s_deg[s_];
          if (deg <= kHubDeg || deg <= k) continue;   // block-uniform
          // every warp scans every 8th pass of 512 weights into its own candidate list ...
          __syncthreads();   // the lists of the previous hub row have been read
          const AresBuf mine = ares_buf(warp);
          const int M = ares_collect(s_w[s_], deg, k, 512 * warp, 512 * kBkWarps, rng_key,
                                     (uint64_t)(i0 + s_), lane, mine);
          if (lane == 0) s_mcount[warp] = M;
          __syncthreads();
          // ... and every finalist counts how many finalists of all warps beat it: the k best of
          // the union, in order
          if (lane < M) {
            const float kv = mine.key[lane];
            const int iv = mine.idx[lane];
            int rank = 0;
            for (int w = 0; w < kBkWarps; ++w) {
              const AresBuf o = ares_buf(w);
              const int Mo = s_mcount[w];
              for (int j = 0; j < Mo; ++j) rank += ares_before(o.key[j], o.idx[j], kv, iv) ? 1 : 0;
            }
            if (rank < k) s_pick[(size_t)s_ * k + rank] = (unsigned int)iv;
          }
        }
      }
    }
    __syncthreads();
    if (tile == tile0) fstamp(2);
    const int slots = ns * k;
    for (int base = tid; base < slots; base += kBkThreads * kPkBatch) {
      long long v[kPkBatch];
      unsigned int item[kPkBatch], pos[kPkBatch];
      bool ok[kPkBatch];
#pragma unroll
      for (int u = 0; u < kPkBatch; ++u) {
        const int el = base + u * kBkThreads;
        ok[u] = false;
        v[u] = 0;
        item[u] = (unsigned int)(S_ub + i0 * k + el) | tagbits;
        if (el < slots) {
          const int si = el / k;
          const int j = el - si * k;
          const bool deferred = MODE == kBias && k <= 32 && hubs != nullptr && s_deg[si] > kHubDeg &&
                                s_deg[si] > k;   // its picks come from the hub phase
          if (j < s_cnt[si] && !deferred) {
            const int deg = s_deg[si];
            unsigned int p;
            if (MODE == kUniformReplace)
              p = rand_below(philox_u32(rng_key, (uint64_t)(i0 + si), (uint32_t)j), (uint32_t)deg);
            else
              p = (!with_replace && deg <= k) ? (unsigned int)j : s_pick[el];
            v[u] = (long long)s_row[si][p];
            ok[u] = true;
          }
        }
      }
      rl_insert_batch<kPkBatch>(cur.table, cap_mask, v, item, ok, pos);
#pragma unroll
      for (int u = 0; u < kPkBatch; ++u) {
        if (ok[u]) {
          const int64_t e = i0 * k + (base + u * kBkThreads);
          pad_col[e] = (IdT)v[u];
          if (!kNoPos) cur.pos_col[e] = pos[u];
        }
      }
    }
    if (tid < ns) {
      while (seed_prev != (unsigned long long)kEmptyKey && seed_prev != (unsigned long long)seed_nid) {
        seed_pos = (seed_pos + 1) & cap_mask;
        seed_prev = atomicCAS(cur.table.key(seed_pos), (unsigned long long)kEmptyKey,
                              (unsigned long long)seed_nid);
      }
      atomicMin(cur.table.first(seed_pos), (unsigned int)(i0 + tid) | tagbits);
      if (!kNoPos) cur.pos_seed[i0 + tid] = (unsigned int)seed_pos;
    }
    if (tile == tile0) fstamp(3);
  }
  fstamp(4);
}

// Rank phase, one tile of 128 seeds: flags first occurrences among the seeds (A) and the sampled
// neighbours (B), counts the edges (C), block scans, per-tile totals to prefA / prefB / prefC.
// kNoPos: table slot = id, read from seeds / pad_col instead of the slot arrays.
template <typename IdT, bool kNoPos>
__device__ __forceinline__ void rank_one_tile(int64_t tile, int64_t S_ub, int64_t S, int k,
                                              const HopState &cur, const BlocksWs &ws,
                                              bool unique_seeds, const IdT *__restrict__ seeds,
                                              const IdT *__restrict__ pad_col,
                                              unsigned int tagbits = 0) {
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ int s_cnt[kBkTile];
  const int tid = threadIdx.x;
  const int64_t i0 = tile * kBkTile;
  const int ns = (int)min((int64_t)kBkTile, S - i0);
  const int items = ns * k;
  const int64_t e0 = i0 * k;
  __syncthreads();
  // first wave of loads: seed slot + count, and the first pass of neighbour slots
  long long fa = 0, c = 0;
  unsigned int slot_a = 0;
  if (tid < ns) {
    if (kNoPos) {
      if (!unique_seeds) slot_a = (unsigned int)ldcg(seeds + i0 + tid);
    } else {
      slot_a = ldcg(cur.pos_seed + i0 + tid);
    }
    c = ldcg(cur.cnt + i0 + tid);
    s_cnt[tid] = (int)c;
  }
  __syncthreads();
  // seeds that are a previous frontier are distinct: each is its own first occurrence
  if (tid < ns) {
    if (unique_seeds) {
      fa = 1;
    } else {
      const unsigned int f = ldcg(cur.table.first(slot_a));
      fa = f == ((unsigned int)(i0 + tid) | tagbits) ? 1 : 0;
      if (kNoPos) cur.fseed[i0 + tid] = f & 0x00ffffffu;
    }
  }
  long long carry = 0;
  long long totA = 0, totC = 0;
  for (int base = 0; base < items || base == 0; base += kBkThreads * kRkItems) {
    const int el0 = base + tid * kRkItems;
    unsigned int slot[kRkItems];
    bool valid[kRkItems];
#pragma unroll
    for (int u = 0; u < kRkItems; ++u) {
      const int el = el0 + u;
      valid[u] = false;
      if (el < items) {
        const int si = el / k;
        valid[u] = (el - si * k) < s_cnt[si];
        if (valid[u])
          slot[u] = kNoPos ? (unsigned int)ldcg(pad_col + e0 + el) : ldcg(cur.pos_col + e0 + el);
      }
    }
    unsigned int first[kRkItems];
#pragma unroll
    for (int u = 0; u < kRkItems; ++u)
      if (valid[u]) first[u] = ldcg(cur.table.first(slot[u]));
    if (base == 0) {
      // A and C share one packed scan (A in the high half)
      const long long rac = block_exclusive_scan<long long>((fa << 32) | c, s_scan, &s_total);
      totA = s_total >> 32;
      totC = s_total & 0xffffffffll;
      if (kNoPos) {
        if (fa && !unique_seeds) cur.lrank_at[i0 + tid] = (unsigned int)(rac >> 32);
      } else if (fa) {
        *cur.table.lrank(slot_a) = (unsigned int)(rac >> 32);
      }
      if (tid < ns) ws.loff[i0 + tid] = (int)(rac & 0xffffffffll);
    }
    int mine = 0;
    bool fb[kRkItems];
#pragma unroll
    for (int u = 0; u < kRkItems; ++u) {
      fb[u] = valid[u] && first[u] == ((unsigned int)(S_ub + e0 + el0 + u) | tagbits);
      mine += fb[u] ? 1 : 0;
      if (kNoPos && valid[u]) cur.fslot[e0 + el0 + u] = first[u] & 0x00ffffffu;
    }
    long long r = carry + block_exclusive_scan<long long>((long long)mine, s_scan, &s_total);
#pragma unroll
    for (int u = 0; u < kRkItems; ++u) {
      if (fb[u]) {
        if (kNoPos)
          cur.lrank_at[S_ub + e0 + el0 + u] = (unsigned int)(r++);
        else
          *cur.table.lrank(slot[u]) = (unsigned int)(r++);
      }
    }
    carry += s_total;
  }
  if (tid == 0) {
    ws.prefA[tile] = totA;
    ws.prefB[tile] = carry;
    ws.prefC[tile] = totC;
  }
}

__device__ __forceinline__ void rank_tiles_phase(int64_t S_ub, int64_t S, int k, const HopState &cur,
                                                 const BlocksWs &ws, bool unique_seeds) {
  const int64_t tiles = (S + kBkTile - 1) / kBkTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    rank_one_tile<long long, false>(tile, S_ub, S, k, cur, ws, unique_seeds, nullptr, nullptr);
}

// Executed by ONE CTA after every tile total is visible: exclusive scans of the three per-tile
// totals (every thread owns a run of consecutive tiles), publishes nnz and |frontier|.
__device__ __forceinline__ void rank_tail(int64_t S, const BlocksWs &ws, long long *out_nnz,
                                          long long *out_nfront) {
  __shared__ long long t_scan[32];
  __shared__ long long t_total;
  const int tid = threadIdx.x;
  const int64_t tiles = (S + kBkTile - 1) / kBkTile;
  volatile long long *pa = ws.prefA, *pb = ws.prefB, *pc = ws.prefC;
  const int64_t per = (tiles + kBkThreads - 1) / kBkThreads;
  const int64_t t0 = min(tiles, (int64_t)tid * per), t1 = min(tiles, t0 + per);
  long long sa = 0, sb = 0, sc = 0;
  for (int64_t t = t0; t < t1; ++t) {
    sa += pa[t];
    sb += pb[t];
    sc += pc[t];
  }
  long long xa = block_exclusive_scan<long long>(sa, t_scan, &t_total);
  const long long totA = t_total;
  long long xb = block_exclusive_scan<long long>(sb, t_scan, &t_total);
  const long long totB = t_total;
  long long xc = block_exclusive_scan<long long>(sc, t_scan, &t_total);
  const long long totC = t_total;
  for (int64_t t = t0; t < t1; ++t) {
    const long long va = pa[t], vb = pb[t], vc = pc[t];
    pa[t] = xa; pb[t] = xb; pc[t] = xc;
    xa += va; xb += vb; xc += vc;
  }
  if (tid == 0) {
    pa[tiles] = totA;
    pb[tiles] = totB;
    pc[tiles] = totC;
    *out_nnz = totC;
    *out_nfront = totA + totB;
  }
}

// Where the exclusive per-tile prefixes live during the emit phase: in global memory (written by
// the last-CTA tail; multi-kernel path) or in the CTA's own shared memory (cooperative kernel:
// every CTA scans the per-tile totals itself right after the barrier - no atomic counter, no
// serial tail, and the emit's prefix lookups never leave the SM).
struct TilePrefix {
  const long long *gA, *gB, *gC;       // global
  const unsigned int *sA, *sB, *sC;    // shared (non-null selects them)
  __device__ __forceinline__ long long A(int64_t t) const { return sA ? (long long)sA[t] : ldcg(gA + t); }
  __device__ __forceinline__ long long B(int64_t t) const { return sB ? (long long)sB[t] : ldcg(gB + t); }
  __device__ __forceinline__ long long C(int64_t t) const { return sC ? (long long)sC[t] : ldcg(gC + t); }
};

// new id of the key stored in `s` (first occurrence f, rank inside its tile)
__device__ __forceinline__ long long new_id(const RlSlot &s, int64_t S_ub, int k, long long totA,
                                            const TilePrefix &pf, bool unique_seeds) {
  const unsigned int f = s.first;
  if ((int64_t)f < S_ub) return unique_seeds ? (long long)f : pf.A(f / kBkTile) + (long long)s.lrank;
  const int64_t e = (int64_t)f - S_ub;
  return totA + pf.B((e / k) / kBkTile) + (long long)s.lrank;
}

// Cooperative kernel, start of the emit phase: exclusive scans of the per-tile totals (ws.prefA/B/C
// hold TOTALS here) into shared memory; CTA 0 publishes nnz and |frontier|.  Entries [0, tiles].
__device__ __forceinline__ void scan_totals_to_smem(int64_t S, const BlocksWs &ws, bool unique_seeds,
                                                    unsigned int *sA, unsigned int *sB,
                                                    unsigned int *sC, long long *out_nnz,
                                                    long long *out_nfront) {
  __shared__ long long t_scan[32];
  __shared__ long long t_total;
  const int tid = threadIdx.x;
  const int64_t tiles = (S + kBkTile - 1) / kBkTile;
  const int64_t per = (tiles + kBkThreads - 1) / kBkThreads;
  const int64_t t0 = min(tiles, (int64_t)tid * per), t1 = min(tiles, t0 + per);
  long long sa = 0, sbc = 0;   // B in the high half, C in the low half (both < 2^32)
  for (int64_t t = t0; t < t1; ++t) {
    if (!unique_seeds) sa += ldcg(ws.prefA + t);
    sbc += (ldcg(ws.prefB + t) << 32) | ldcg(ws.prefC + t);
  }
  long long xbc = block_exclusive_scan<long long>(sbc, t_scan, &t_total);
  const long long totBC = t_total;
  long long xa = 0, totA = S;
  if (!unique_seeds) {
    xa = block_exclusive_scan<long long>(sa, t_scan, &t_total);
    totA = t_total;
  }
  for (int64_t t = t0; t < t1; ++t) {
    sA[t] = unique_seeds ? (unsigned int)(t * kBkTile) : (unsigned int)xa;
    sB[t] = (unsigned int)(xbc >> 32);
    sC[t] = (unsigned int)(xbc & 0xffffffffll);
    if (!unique_seeds) xa += ldcg(ws.prefA + t);
    xbc += (ldcg(ws.prefB + t) << 32) | ldcg(ws.prefC + t);
  }
  if (tid == 0) {
    sA[tiles] = (unsigned int)totA;
    sB[tiles] = (unsigned int)(totBC >> 32);
    sC[tiles] = (unsigned int)(totBC & 0xffffffffll);
    if (blockIdx.x == 0) {
      *out_nnz = totBC & 0xffffffffll;
      *out_nfront = totA + (totBC >> 32);
    }
  }
  __syncthreads();
}

// Emit phase: thread per padded slot - frontier[new id] = id for first occurrences, COO written
// compacted and relabelled.  kPkBatch slots per thread and pass, loads staged level by level.
template <typename IdT>
__device__ __forceinline__ void emit_phase(const IdT *__restrict__ seeds, int64_t S_ub, int64_t S,
                                           int k, const IdT *__restrict__ pad_col,
                                           const HopState &cur, const BlocksWs &ws,
                                           IdT *__restrict__ frontier, IdT *__restrict__ out_row,
                                           IdT *__restrict__ out_col, bool unique_seeds,
                                           const TilePrefix &pf) {
  const int64_t tiles = (S + kBkTile - 1) / kBkTile;
  const long long totA = pf.A(tiles);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) *ws.pending_S = S;  // this hop's table is dirty until the next pick phase wipes it
  if (unique_seeds) {   // new id of seed i is i
    for (int64_t i = tid; i < S; i += stride) frontier[i] = ldcg(seeds + i);
  } else {
    for (int64_t i = tid; i < S; i += stride) {
      const int2 fl = cur.table.first_lrank(ldcg(cur.pos_seed + i));
      if ((unsigned int)fl.x == (unsigned int)i)
        frontier[pf.A(i / kBkTile) + (long long)(unsigned int)fl.y] = ldcg(seeds + i);
    }
  }
  if (k <= 0) return;
  const int64_t E = S * k;
  for (int64_t base = tid; base < E; base += stride * kEmBatch) {
    int64_t si[kEmBatch];
    int jj[kEmBatch];
    bool ok[kEmBatch];
    unsigned int pc[kEmBatch], ps[kEmBatch];
#pragma unroll
    for (int u = 0; u < kEmBatch; ++u) {
      const int64_t e = base + u * stride;
      ok[u] = false;
      if (e < E) {
        si[u] = e / k;
        jj[u] = (int)(e - si[u] * k);
        ok[u] = jj[u] < ldcg(cur.cnt + si[u]);
        pc[u] = ldcg(cur.pos_col + e);
        if (!unique_seeds) ps[u] = ldcg(cur.pos_seed + si[u]);
      }
    }
    unsigned int cf[kEmBatch], cr[kEmBatch], sf[kEmBatch], sr[kEmBatch];
#pragma unroll
    for (int u = 0; u < kEmBatch; ++u) {
      if (ok[u]) {
        const int2 c2 = cur.table.first_lrank(pc[u]);
        cf[u] = (unsigned int)c2.x; cr[u] = (unsigned int)c2.y;
        if (!unique_seeds) {
          const int2 s2 = cur.table.first_lrank(ps[u]);
          sf[u] = (unsigned int)s2.x; sr[u] = (unsigned int)s2.y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kEmBatch; ++u) {
      if (ok[u]) {
        const int64_t e = base + u * stride;
        RlSlot sc;
        sc.first = cf[u];
        sc.lrank = cr[u];
        const long long cid = new_id(sc, S_ub, k, totA, pf, unique_seeds);
        if ((int64_t)cf[u] == S_ub + e) frontier[cid] = ldcg(pad_col + e);
        const long long rid =
            unique_seeds ? (long long)si[u] : pf.A(sf[u] / kBkTile) + (long long)sr[u];
        const long long o = pf.C(si[u] / kBkTile) + (long long)ldcg(ws.loff + si[u]) + jj[u];
        out_row[o] = (IdT)rid;
        out_col[o] = (IdT)cid;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-kernel path: 3 launches per hop.
template <typename IdT, typename ET, int MODE>
__global__ void __launch_bounds__(kBkThreads)
fused_pick_tile_kernel(GraphSrc g, const IdT *__restrict__ seeds, int64_t S_ub,
                       const int64_t *__restrict__ S_dev, int k, uint64_t rng_key,
                       IdT *__restrict__ pad_col, HopState cur, uint64_t cap_mask, HopState prev,
                       int64_t prev_S_ub, const long long *__restrict__ prev_S_dev, int prev_k) {
  const long long pS_live = *prev_S_dev;
  const int64_t S = S_dev ? min(*S_dev, S_ub) : S_ub;
  pick_tile_phase<IdT, ET, MODE>(g, seeds, S_ub, S, k, rng_key, pad_col, cur, cap_mask,
                                 pick_tile_seeds(S), blockIdx.x, gridDim.x, 0u);
  wipe_hop(prev, min((int64_t)pS_live, prev_S_ub), prev_k, (S + pick_tile_seeds(S) - 1) / pick_tile_seeds(S),
           (int64_t)cap_mask + 1);
}

__global__ void __launch_bounds__(kBkThreads)
fused_rank_kernel(int64_t S_ub, const int64_t *__restrict__ S_dev, int k, HopState cur,
                  BlocksWs ws, int64_t *__restrict__ out_nnz, int64_t *__restrict__ out_nfront,
                  int unique_seeds) {
  __shared__ bool s_last;
  const int64_t S = S_dev ? min(*S_dev, S_ub) : S_ub;
  rank_tiles_phase(S_ub, S, k, cur, ws, unique_seeds != 0);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ws.done, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    rank_tail(S, ws, (long long *)out_nnz, (long long *)out_nfront);
    if (threadIdx.x == 0) *ws.done = 0;
  }
}

template <typename IdT>
__global__ void __launch_bounds__(kBkThreads)
fused_emit_kernel(const IdT *__restrict__ seeds, int64_t S_ub, const int64_t *__restrict__ S_dev,
                  int k, const IdT *__restrict__ pad_col, HopState cur, BlocksWs ws,
                  IdT *__restrict__ frontier, IdT *__restrict__ out_row, IdT *__restrict__ out_col,
                  int unique_seeds) {
  const int64_t S = S_dev ? min(*S_dev, S_ub) : S_ub;
  const TilePrefix pf{ws.prefA, ws.prefB, ws.prefC, nullptr, nullptr, nullptr};
  emit_phase<IdT>(seeds, S_ub, S, k, pad_col, cur, ws, frontier, out_row, out_col, unique_seeds != 0, pf);
}

// ---------------------------------------------------------------------------------------------
// Cooperative whole-batch kernel: every hop's pick / rank / totals / emit as phases of ONE launch,
// separated by grid barriers - no kernel boundary (launch + drain + ramp, ~3-4 us each at these
// sizes) and a single CPU-side launch per mini-batch.
struct HopArgs {
  const void *seeds;
  void *frontier, *out_row, *out_col;
  long long *nnz_dev, *nf_dev;
  const long long *S_dev;   // live seed count (nullptr: S_ub is exact)
  int64_t S_ub, prev_S_ub;
  uint64_t key;
  int k, prev_k, cur;
  int smem_pref;   // emit phase keeps the tile prefixes in shared memory (they fit)
};
struct BatchArgs {
  int L;
  uint64_t cap_mask;
  unsigned long long *trace;  // debug: %globaltimer stamps of CTA 0 around every phase
  long long *host_counts;     // mapped pinned host memory for the 2 L hop sizes (or null)
  HopArgs hop[8];
};

template <typename IdT, typename ET, int MODE>
__global__ void __launch_bounds__(kBkThreads, DGS_COOP_MIN_CTAS)
fused_batch_kernel(GraphSrc g, BlocksWs ws, BatchArgs a) {
  cg::grid_group grid = cg::this_grid();
  int nt = 0;
  auto stamp = [&]() {
    if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[nt++] = t;
    }
  };
  stamp();
  for (int l = 0; l < a.L; ++l) {
    const HopArgs &h = a.hop[l];
    const HopState &cur = ws.hop[h.cur];
    const HopState &prev = ws.hop[h.cur ^ 1];
    const long long pS_live = ldcg(ws.pending_S);
    const int64_t S = h.S_dev ? min((int64_t)ldcg(h.S_dev), h.S_ub) : h.S_ub;
    pick_tile_phase<IdT, ET, MODE>(g, (const IdT *)h.seeds, h.S_ub, S, h.k, h.key,
                                   (IdT *)ws.pad_col, cur, a.cap_mask, pick_tile_seeds(S), blockIdx.x,
                                   gridDim.x, 0u, a.trace ? a.trace + 256 + 8 * l : nullptr);
    wipe_hop(prev, min((int64_t)pS_live, h.prev_S_ub), h.prev_k, (S + pick_tile_seeds(S) - 1) / pick_tile_seeds(S),
               (int64_t)a.cap_mask + 1);
    stamp();
    grid.sync();
    stamp();
    rank_tiles_phase(h.S_ub, S, h.k, cur, ws, l > 0);
    stamp();
    TilePrefix pf{ws.prefA, ws.prefB, ws.prefC, nullptr, nullptr, nullptr};
    if (h.smem_pref) {
      stamp();
      grid.sync();
      stamp();
      extern __shared__ __align__(16) unsigned char dyn_smem[];
      const int64_t tiles = (S + kBkTile - 1) / kBkTile;
      unsigned int *sA = reinterpret_cast<unsigned int *>(dyn_smem);
      unsigned int *sB = sA + (tiles + 1), *sC = sB + (tiles + 1);
      scan_totals_to_smem(S, ws, l > 0, sA, sB, sC, h.nnz_dev, h.nf_dev);
      pf.sA = sA;
      pf.sB = sB;
      pf.sC = sC;
    } else {
      __shared__ bool s_last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(ws.done, 1u) == gridDim.x - 1);
      __syncthreads();
      if (s_last) {
        __threadfence();
        rank_tail(S, ws, h.nnz_dev, h.nf_dev);
        if (threadIdx.x == 0) *ws.done = 0;
      }
      stamp();
      grid.sync();
      stamp();
    }
    emit_phase<IdT>((const IdT *)h.seeds, h.S_ub, S, h.k, (const IdT *)ws.pad_col, cur, ws,
                    (IdT *)h.frontier, (IdT *)h.out_row, (IdT *)h.out_col, l > 0, pf);
    stamp();
    grid.sync();
    stamp();
  }
  // The hop sizes go straight into the caller's pinned host memory (posted PCIe writes): entry 0
  // is written last, behind a system-scope fence, and is what the host polls - no copy engine,
  // no stream synchronisation in the one host round trip of a batch.
  if (a.host_counts != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long *dev = a.hop[0].nnz_dev;   // {nnz_0, |frontier_0|, nnz_1, ...}
    for (int i = 1; i < 2 * a.L; ++i) a.host_counts[i] = ldcg(dev + i);
    __threadfence_system();
    *(volatile long long *)a.host_counts = ldcg(dev);
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-batch cooperative kernel (direct-addressed relabel tables): B INDEPENDENT mini-batches of
// the same shape in one launch.  A batch of 1024 seeds cannot fill 148 SMs - the two small hops
// are pure latency chains (a few dependent round trips + a grid barrier per phase) - so B batches
// walk through every phase together and share each of the grid barriers; the per-batch cost of a
// barrier / a dependent round trip drops by B.  Results are bit-identical to B single launches
// with the same RNG seeds (the RNG is keyed by (batch seed, hop, item), never by geometry).
//
// Differences from fused_batch_kernel above (which stays for hashed tables):
//   * no slot arrays: with direct tables the slot of an id is the id, so pos_seed / pos_col are
//     gone (one 4-byte store + two 4-byte loads per padded slot less);
//   * the table is NEVER wiped: an entry is {tag:8 | first item:24, lrank:32} and every use of the
//     table (one per hop) takes the next smaller tag, so atomicMin makes a newer entry beat any
//     stale one and a hop only ever reads ids it inserted itself.  One table per batch instead of
//     two alternating ones, no wipe stores (they were one random store per frontier id - 50 us per
//     8-batch launch), no epoch argument.  After 255 uses the host clears the table (a memset
//     every ~85 calls).  Items per hop must fit 24 bits (15.1 M at the friendster config);
//   * tiles of all batches form one virtual tile space that is dealt round-robin to the CTAs;
//     the emit phase runs over the concatenated slot space of all batches.
// Per-batch arrays live at a fixed byte stride (workspace, outputs, counts), so the kernel
// parameters describe batch 0 only.
struct MbHop {
  const void *seeds;                  // batch 0; stride = l == 0 ? seeds_stride : out_stride
  void *frontier, *out_row, *out_col; // batch 0; stride out_stride
  int64_t S_ub;
  int k;
  int smem_pref;                      // the tile prefixes of all batches fit in shared memory
};
struct MbArgs {
  int L, B;
  unsigned int tag0;                  // hop l stores / expects tag (tag0 - l) in bits 31..24
  int scan_in_emit;                   // B = 1 cooperative launch: no last-CTA tail in the rank phase,
                                      // every CTA scans the tile totals itself after the barrier
  int64_t seeds_stride, out_stride, ws_stride;   // bytes between consecutive batches
  long long *counts_dev;              // [B][2 L]  {nnz_0, |frontier_0|, nnz_1, ...}
  long long *host_counts;             // same layout in mapped pinned host memory (or null)
  unsigned long long *trace;
  uint64_t rng[DGS_MAX_BATCHES];
  MbHop hop[8];
};

// Per-batch views of the workspace: every array of batch b sits `bo` bytes behind batch 0's.
// (No dynamic indexing of struct members here - that would push the whole struct to local memory.)
template <typename T>
__device__ __forceinline__ T *off_ptr(T *p, int64_t bo) {
  return (T *)((char *)p + bo);
}
__device__ __forceinline__ Tab tab_of(const BlocksWs &w0, int64_t bo) {
  return Tab{w0.hop[0].table.base + bo, 2};
}
__device__ __forceinline__ HubList hubs_of_batch(const BlocksWs &w0, int64_t bo) {
  return HubList{off_ptr(w0.hubs.count, bo), off_ptr(w0.hubs.ent, bo), off_ptr(w0.hubs.done, bo),
                 off_ptr(w0.hubs.ckey, bo), off_ptr(w0.hubs.cidx, bo), off_ptr(w0.hubs.cnum, bo)};
}
__device__ __forceinline__ HopState hop_of_batch(const BlocksWs &w0, int64_t bo) {
  HopState h;
  h.cnt = off_ptr(w0.hop[0].cnt, bo);
  h.pos_seed = nullptr;
  h.pos_col = nullptr;
  h.table = tab_of(w0, bo);
  h.fslot = off_ptr(w0.hop[0].fslot, bo);
  h.fseed = off_ptr(w0.hop[0].fseed, bo);
  h.lrank_at = off_ptr(w0.hop[0].lrank_at, bo);
  return h;
}
constexpr unsigned int kItemMask = 0x00ffffffu;   // low 24 bits of `first`: the item index
__device__ __forceinline__ BlocksWs ws_of_batch(const BlocksWs &w0, int64_t bo) {
  BlocksWs w;
  w.done = off_ptr(w0.done, bo);
  w.hubs = hubs_of_batch(w0, bo);
  w.pending_S = nullptr;
  w.prefA = off_ptr(w0.prefA, bo);
  w.prefB = off_ptr(w0.prefB, bo);
  w.prefC = off_ptr(w0.prefC, bo);
  w.loff = off_ptr(w0.loff, bo);
  w.pad_col = off_ptr(reinterpret_cast<char *>(w0.pad_col), bo);
  w.hop[0] = w.hop[1] = hop_of_batch(w0, bo);
  w.cap = w0.cap;
  w.direct = 1;
  return w;
}

// batch that owns virtual index v, given exclusive offsets off[0..B] (off[B] = total)
__device__ __forceinline__ int batch_of(const long long *off, int B, long long v) {
  int b = 0;
  for (int i = 1; i < B; ++i) b += (v >= off[i]) ? 1 : 0;
  return b;
}

// Hub phase (weighted sampling without replacement, k <= 32): the deferred rows of every batch form
// one list that is dealt round-robin to ALL CTAs.  A CTA scans a row with its 8 warps (every 8th
// 512-weight pass each), ranks the finalists across warps and does the row's k neighbour loads +
// table inserts - exactly what the tile phase does for a row it owns, so the sample is the same.
template <typename IdT>
__device__ __forceinline__ void hub_rows(const HubList &hubs, int64_t j0, int64_t jstride, int k,
                                         uint64_t rng_key, int64_t S_ub, IdT *__restrict__ pad_col,
                                         const Tab &table, unsigned int tagbits) {
  __shared__ int h_mcount[kBkWarps];
  __shared__ unsigned int h_pick[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = (int64_t)ldcg(hubs.count);
  for (int64_t j = j0; j < n; j += jstride) {
    const int4 raw0 = __ldcg(reinterpret_cast<const int4 *>(hubs.ent + j));
    const int2 raw1 = __ldcg(reinterpret_cast<const int2 *>(hubs.ent + j) + 2);
    const IdT *row = reinterpret_cast<const IdT *>(((unsigned long long)(unsigned int)raw0.y << 32) | (unsigned int)raw0.x);
    const float *w = reinterpret_cast<const float *>(((unsigned long long)(unsigned int)raw0.w << 32) | (unsigned int)raw0.z);
    const unsigned int i = (unsigned int)raw1.x;
    const int deg = raw1.y;
    __syncthreads();   // the candidate lists / picks of the previous row have been read
    const AresBuf mine = ares_buf(warp);
    const int M = ares_collect(w, deg, k, 512 * warp, 512 * kBkWarps, rng_key, (uint64_t)i, lane, mine);
    if (lane == 0) h_mcount[warp] = M;
    __syncthreads();
    if (lane < M) {
      const float kv = mine.key[lane];
      const int iv = mine.idx[lane];
      int rank = 0;
      for (int ww = 0; ww < kBkWarps; ++ww) {
        const AresBuf o = ares_buf(ww);
        const int Mo = h_mcount[ww];
        for (int q = 0; q < Mo; ++q) rank += ares_before(o.key[q], o.idx[q], kv, iv) ? 1 : 0;
      }
      if (rank < k) h_pick[rank] = (unsigned int)iv;
    }
    __syncthreads();
    if (tid < k) {
      const IdT v = row[h_pick[tid]];
      const int64_t e = (int64_t)i * k + tid;
      atomicMin(table.first((uint64_t)v), (unsigned int)(S_ub + e) | tagbits);
      pad_col[e] = v;
    }
  }
}

// Chunked hub phase: a row of 17 000 weights on ONE CTA is ~15 us, longer than everything else in
// the hop.  Rows are cut into chunks of kHubChunk weights; the (row, chunk) pairs of a batch are
// dealt round-robin to the CTAs.  A chunk's CTA keeps the chunk's k best candidates (ordered); the
// CTA that finishes a row's LAST chunk merges the row's candidates (every candidate counts how
// many beat it - the total order is (key descending, position ascending), so set and order equal
// the single-CTA and single-warp selections) and does the row's neighbour loads + table inserts.
// Returns false (nothing done) when the list exceeds the static limits: the caller then runs
// hub_rows.  q0 / qstride: the chunk ids this CTA owns.
template <typename IdT>
__device__ __forceinline__ bool hub_chunks(const HubList &hubs, int64_t rot, int k, uint64_t rng_key,
                                           int64_t S_ub, IdT *__restrict__ pad_col, const Tab &table,
                                           unsigned int tagbits, long long *total_chunks) {
  __shared__ int c_qoff[kHubMaxRows + 1];      // exclusive chunk offsets of the rows
  __shared__ long long c_scan[32];
  __shared__ long long c_total;
  __shared__ int c_mcount[kBkWarps];
  __shared__ float m_key[kHubMaxPerRow * 32];  // merge: candidates of all chunks of one row
  __shared__ int m_idx[kHubMaxPerRow * 32];
  __shared__ int m_n;
  __shared__ unsigned int c_pick[32];
  __shared__ bool c_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int n = (int)ldcg(hubs.count);
  *total_chunks = 0;
  if (n == 0) return true;
  if (n > kHubMaxRows) return false;
  // chunks per row -> exclusive offsets (4 rows per thread)
  __syncthreads();
  int nch[4], mine = 0;
  bool too_long = false;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid * 4 + u;
    nch[u] = 0;
    if (j < n) {
      const int deg = __ldcg(reinterpret_cast<const int *>(hubs.ent + j) + 5);
      nch[u] = (deg + kHubChunk - 1) / kHubChunk;
      too_long |= nch[u] > kHubMaxPerRow;
    }
    mine += nch[u];
  }
  long long ex = block_exclusive_scan<long long>((long long)mine, c_scan, &c_total);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid * 4 + u;
    if (j <= n && j <= kHubMaxRows) c_qoff[min(j, kHubMaxRows)] = (int)ex;
    ex += nch[u];
  }
  const bool bad = __syncthreads_or(too_long ? 1 : 0) != 0;
  const int Q = (int)c_total;
  if (bad || Q > kHubMaxChunks) return false;
  if (tid == 0) c_qoff[n] = Q;
  __syncthreads();
  *total_chunks = Q;
  for (int q = (int)(((int64_t)blockIdx.x - rot) % G + G) % G; q < Q; q += G) {
    // row of chunk q: last j with c_qoff[j] <= q (binary search, uniform)
    int lo = 0, hi = n;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (c_qoff[mid] <= q) lo = mid; else hi = mid;
    }
    const int j = lo, c = q - c_qoff[j], nchunks = c_qoff[j + 1] - c_qoff[j];
    const int4 raw0 = __ldcg(reinterpret_cast<const int4 *>(hubs.ent + j));
    const int2 raw1 = __ldcg(reinterpret_cast<const int2 *>(hubs.ent + j) + 2);
    const IdT *row = reinterpret_cast<const IdT *>(((unsigned long long)(unsigned int)raw0.y << 32) | (unsigned int)raw0.x);
    const float *w = reinterpret_cast<const float *>(((unsigned long long)(unsigned int)raw0.w << 32) | (unsigned int)raw0.z);
    const unsigned int i = (unsigned int)raw1.x;
    const int deg = raw1.y;
    const int c_end = min(deg, (c + 1) * kHubChunk);
    __syncthreads();   // shared buffers of the previous chunk have been read
    const AresBuf mine_b = ares_buf(warp);
    const int M = ares_collect(w, c_end, k, c * kHubChunk + 512 * warp, kHubChunk, rng_key, (uint64_t)i, lane,
                               mine_b);
    if (lane == 0) c_mcount[warp] = M;
    __syncthreads();
    // the chunk's k best, in order, to shared memory (single chunk) or to the row's global slots
    if (lane < M) {
      const float kv = mine_b.key[lane];
      const int iv = mine_b.idx[lane];
      int rank = 0;
      for (int ww = 0; ww < kBkWarps; ++ww) {
        const AresBuf o = ares_buf(ww);
        const int Mo = c_mcount[ww];
        for (int t = 0; t < Mo; ++t) rank += ares_before(o.key[t], o.idx[t], kv, iv) ? 1 : 0;
      }
      if (rank < k) {
        if (nchunks == 1) {
          c_pick[rank] = (unsigned int)iv;
        } else {
          hubs.ckey[(size_t)q * 32 + rank] = kv;
          hubs.cidx[(size_t)q * 32 + rank] = iv;
        }
      }
    }
    if (nchunks > 1) {
      if (tid == 0) {
        int tot = 0;
        for (int ww = 0; ww < kBkWarps; ++ww) tot += c_mcount[ww];
        hubs.cnum[q] = min(tot, k);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) c_last = atomicAdd(hubs.done + j, 1u) == (unsigned int)(nchunks - 1);
      __syncthreads();
      if (!c_last) continue;            // (uniform) another CTA will merge this row
      __threadfence();
      if (tid == 0) hubs.done[j] = 0;   // last user: leave the counter clean for the next hop
      // gather the row's candidates: chunk c2 contributes cnum entries
      const int q0 = c_qoff[j];
      if (tid == 0) m_n = 0;
      __syncthreads();
      for (int c2 = warp; c2 < nchunks; c2 += kBkWarps) {
        const int cn = __ldcg(hubs.cnum + q0 + c2);
        int base = 0;
        if (lane == 0) base = atomicAdd(&m_n, cn);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < cn) {
          m_key[base + lane] = __ldcg(hubs.ckey + (size_t)(q0 + c2) * 32 + lane);
          m_idx[base + lane] = __ldcg(hubs.cidx + (size_t)(q0 + c2) * 32 + lane);
        }
      }
      __syncthreads();
      const int T = m_n;
      for (int t = tid; t < T; t += kBkThreads) {
        const float kv = m_key[t];
        const int iv = m_idx[t];
        int rank = 0;
        for (int o = 0; o < T; ++o) rank += ares_before(m_key[o], m_idx[o], kv, iv) ? 1 : 0;
        if (rank < k) c_pick[rank] = (unsigned int)iv;
      }
    }
    __syncthreads();
    if (tid < k) {
      const IdT v = row[c_pick[tid]];
      const int64_t e = (int64_t)i * k + tid;
      atomicMin(table.first((uint64_t)v), (unsigned int)(S_ub + e) | tagbits);
      pad_col[e] = v;
    }
  }
  return true;
}

// Shared-memory scratch of the multi-batch phases.
struct MbShared {
  long long S[DGS_MAX_BATCHES];          // live seed count of every batch, this hop
  long long off[DGS_MAX_BATCHES + 1];    // exclusive offsets (seeds)
  long long off2[DGS_MAX_BATCHES + 1];   // exclusive offsets (padded slots)
  long long next_S;                      // scan_in_emit: |frontier| of this hop = seeds of the next one
  bool last;
};

// live seed counts of hop l into sh.S (every thread of the CTA must call; syncs)
__device__ __forceinline__ void mb_load_S(const MbArgs &a, int l, MbShared &sh) {
  __syncthreads();
  if ((int)threadIdx.x < a.B) {
    long long S = a.hop[l].S_ub;
    if (l > 0)
      S = min((long long)S, ldcg(a.counts_dev + (int64_t)threadIdx.x * 2 * a.L + 2 * (l - 1) + 1));
    sh.S[threadIdx.x] = S;
  }
  __syncthreads();
}

// ---------------- pick: virtual tiles of all batches, dealt round-robin to the CTAs
template <typename IdT, typename ET, int MODE, int FR = 16>
__device__ __forceinline__ void mb_pick(const GraphSrc &g, const BlocksWs &ws0, const MbArgs &a, int l,
                                        MbShared &sh) {
  const int B = a.B;
  const int64_t G = gridDim.x;
  const int k = a.hop[l].k;
  const int64_t S_ub = a.hop[l].S_ub;
  const IdT *const seeds0 = (const IdT *)a.hop[l].seeds;
  const int64_t in_stride = l == 0 ? a.seeds_stride : a.out_stride;
  const unsigned int tagbits = (a.tag0 - (unsigned int)l) << 24;
  long long S_total = 0;
  for (int b = 0; b < B; ++b) S_total += sh.S[b];
  // seeds per tile: small enough that every CTA gets a tile, large enough that the tiles of all
  // batches (each batch rounds up on its own) still fit ONE round of the grid
  int ts;
  {
    const int64_t ctas = max((int64_t)1, G - B);
    int64_t per = (S_total + ctas - 1) / ctas;
    per = (per + 7) & ~7ll;
    ts = (int)max((int64_t)8, min((int64_t)kPkSeeds, per));
  }
  long long toff = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t S = sh.S[b];
    const int64_t tiles = (S + ts - 1) / ts;
    const int64_t t0 = (((int64_t)blockIdx.x - toff) % G + G) % G;
    toff += tiles;
    if (t0 >= tiles) continue;
    const int64_t bo = (int64_t)b * a.ws_stride;
    const HopState cur = hop_of_batch(ws0, bo);
    const HubList hl = hubs_of_batch(ws0, bo);
    const uint64_t key = a.rng[b] + 0x9E3779B97F4A7C15ull * (uint64_t)(l + 1);
    pick_tile_phase<IdT, ET, MODE, true, FR>(
        g, off_ptr(seeds0, (int64_t)b * in_stride), S_ub, S, k, key,
        reinterpret_cast<IdT *>(off_ptr(reinterpret_cast<char *>(ws0.pad_col), bo)), cur, 0, ts, t0, G,
        tagbits, nullptr, (MODE == kBias && k <= 32) ? &hl : nullptr);
  }
}

// does hop l of this launch have a hub phase?  (weighted, without replacement, k <= 32)
template <int MODE>
__device__ __forceinline__ bool mb_has_hubs(const MbArgs &a, int l) {
  return MODE == kBias && a.hop[l].k <= 32;
}

// ---------------- hub phase: the deferred long rows of all batches, dealt round-robin to the CTAs
template <typename IdT>
__device__ __forceinline__ void mb_hubs(const BlocksWs &ws0, const MbArgs &a, int l) {
  const int B = a.B;
  const int64_t G = gridDim.x;
  const int k = a.hop[l].k;
  const unsigned int tagbits = (a.tag0 - (unsigned int)l) << 24;
  long long off = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t bo = (int64_t)b * a.ws_stride;
    const HubList hl = hubs_of_batch(ws0, bo);
    const uint64_t key = a.rng[b] + 0x9E3779B97F4A7C15ull * (uint64_t)(l + 1);
    IdT *pad = reinterpret_cast<IdT *>(off_ptr(reinterpret_cast<char *>(ws0.pad_col), bo));
    long long chunks = 0;
    if (hub_chunks<IdT>(hl, off, k, key, a.hop[l].S_ub, pad, tab_of(ws0, bo), tagbits, &chunks)) {
      off += chunks;
      continue;
    }
    // too many / too long rows for the chunk scratch: one CTA per row
    const long long n = (long long)ldcg(hl.count);
    const int64_t j0 = (((int64_t)blockIdx.x - off) % G + G) % G;
    off += n;
    if (j0 >= n) continue;
    hub_rows<IdT>(hl, j0, G, k, key, a.hop[l].S_ub, pad, tab_of(ws0, bo), tagbits);
  }
}

// ---------------- rank: tiles of 128 seeds; the CTA that finishes a batch's last tile turns that
// batch's tile totals into exclusive prefixes and publishes its hop sizes
template <typename IdT>
__device__ __forceinline__ void mb_rank(const BlocksWs &ws0, const MbArgs &a, int l, MbShared &sh) {
  const int B = a.B, L = a.L;
  const int tid = threadIdx.x;
  const int64_t G = gridDim.x;
  const int k = a.hop[l].k;
  const int64_t S_ub = a.hop[l].S_ub;
  const IdT *const seeds0 = (const IdT *)a.hop[l].seeds;
  const int64_t in_stride = l == 0 ? a.seeds_stride : a.out_stride;
  const bool unique_seeds = l > 0;
  const unsigned int tagbits = (a.tag0 - (unsigned int)l) << 24;
  if (blockIdx.x == 0 && tid < B) *off_ptr(ws0.hubs.count, (int64_t)tid * a.ws_stride) = 0;   // next hop's list
  long long toff = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t S = sh.S[b];
    const int64_t tiles = (S + kBkTile - 1) / kBkTile;
    const int64_t t0 = (((int64_t)blockIdx.x - toff) % G + G) % G;
    toff += tiles;
    if (t0 >= tiles) continue;
    const int64_t bo = (int64_t)b * a.ws_stride;
    const BlocksWs w = ws_of_batch(ws0, bo);
    const HopState cur = w.hop[0];
    const IdT *seeds_b = off_ptr(seeds0, (int64_t)b * in_stride);
    for (int64_t tile = t0; tile < tiles; tile += G) {
      rank_one_tile<IdT, true>(tile, S_ub, S, k, cur, w, unique_seeds, seeds_b, (const IdT *)w.pad_col,
                               tagbits);
      if (a.scan_in_emit) continue;   // (the grid barrier that follows publishes the tile totals)
      __threadfence();
      __syncthreads();
      if (tid == 0) sh.last = (atomicAdd(w.done, 1u) == (unsigned int)(tiles - 1));
      __syncthreads();
      if (sh.last) {
        __threadfence();
        long long *cd = a.counts_dev + (int64_t)b * 2 * L + 2 * l;
        rank_tail(S, w, cd, cd + 1);
        if (tid == 0) *w.done = 0;
      }
    }
  }
}

// ---------------- the hop sizes go straight into the caller's pinned host memory (posted PCIe
// writes): entry 0 is written last, behind a system-scope fence, and is what the host polls.
// All sizes are final once the last hop's rank phase is over, so they are delivered BEFORE the last
// emit phase: the host builds its views and enqueues the next kernel (the extract, ordered behind
// this one on the stream) while the last hop is still being written - ~9 us of host latency hidden
// per batch.
__device__ __forceinline__ void mb_deliver_counts(const MbArgs &a) {
  if (a.host_counts != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    const int n = 2 * a.L * a.B;
    for (int i = 1; i < n; ++i) a.host_counts[i] = ldcg(a.counts_dev + i);
    __threadfence_system();
    *(volatile long long *)a.host_counts = ldcg(a.counts_dev);
  }
}

// ---------------- emit over the concatenated seed / slot spaces of all batches.  EB = padded slots
// per thread and pass (independent load chains in flight).
template <typename IdT, int EB>
__device__ __forceinline__ void mb_emit(const BlocksWs &ws0, const MbArgs &a, int l, MbShared &sh,
                                        unsigned char *dyn_smem, bool deliver) {
  const int B = a.B;
  const int tid = threadIdx.x;
  const int64_t G = gridDim.x;
  const int k = a.hop[l].k;
  const int64_t S_ub = a.hop[l].S_ub;
  const int smem_pref = a.hop[l].smem_pref;
  const IdT *const seeds0 = (const IdT *)a.hop[l].seeds;
  IdT *const frontier0 = (IdT *)a.hop[l].frontier;
  IdT *const row0 = (IdT *)a.hop[l].out_row;
  IdT *const col0 = (IdT *)a.hop[l].out_col;
  const bool unique_seeds = l > 0;
  const int64_t in_stride = l == 0 ? a.seeds_stride : a.out_stride;
  const int64_t ws_stride = a.ws_stride, out_stride = a.out_stride;
  // tile prefixes (exclusive, entries [0, tiles]) of every batch: shared-memory copies when they
  // fit, so the three prefix lookups per slot never leave the SM
  unsigned int *sp = reinterpret_cast<unsigned int *>(dyn_smem);
  const int tiles_ub = (int)((S_ub + kBkTile - 1) / kBkTile + 1);   // entries per array
  if (a.scan_in_emit) {
    // B = 1, cooperative launch: ws.pref* hold the per-tile TOTALS; every CTA scans them into its own
    // shared memory (no atomic counter, no serial tail CTA - measured 84 -> 80 us per batch in
    // round 1) and CTA 0 publishes the hop sizes
    scan_totals_to_smem(sh.S[0], ws0, unique_seeds, sp, sp + tiles_ub, sp + 2 * tiles_ub,
                        a.counts_dev + 2 * l, a.counts_dev + 2 * l + 1);
    if (tid == 0) {   // every CTA knows the next hop's seed count without another global round trip
      const int tl = (int)((sh.S[0] + kBkTile - 1) / kBkTile);
      sh.next_S = (long long)sp[tl] + (long long)sp[tiles_ub + tl];
    }
  } else if (smem_pref) {
    // one flat loop over (batch, tile): a loop over the batches would be B dependent round trips
    // (measured: the emit phase of a tiny hop took 19 us at B = 16 instead of 4)
    const int total = B * tiles_ub;
#pragma unroll 4
    for (int i = tid; i < total; i += kBkThreads) {
      const int b = i / tiles_ub, t = i - b * tiles_ub;
      if (t < (int)((sh.S[b] + kBkTile - 1) / kBkTile + 1)) {
        const int64_t bo = (int64_t)b * ws_stride;
        unsigned int *pa = sp + (size_t)b * 3 * tiles_ub;
        const unsigned int va = (unsigned int)ldcg(off_ptr(ws0.prefA, bo) + t);
        const unsigned int vb = (unsigned int)ldcg(off_ptr(ws0.prefB, bo) + t);
        const unsigned int vc = (unsigned int)ldcg(off_ptr(ws0.prefC, bo) + t);
        pa[t] = va;
        pa[tiles_ub + t] = vb;
        pa[2 * tiles_ub + t] = vc;
      }
    }
  }
  if (tid == 0) {
    long long o1 = 0, o2 = 0;
    for (int b = 0; b < B; ++b) {
      sh.off[b] = o1;
      sh.off2[b] = o2;
      o1 += sh.S[b];
      o2 += sh.S[b] * k;
    }
    sh.off[B] = o1;
    sh.off2[B] = o2;
  }
  __syncthreads();
  if (deliver) mb_deliver_counts(a);    // all hop sizes are final here: hand them to the host now
  const int64_t stride = G * kBkThreads;
  const int64_t gtid = (int64_t)blockIdx.x * kBkThreads + tid;
  // seeds -> frontier
  if (unique_seeds) {
    const int64_t tot = sh.off[B];
    for (int64_t v0 = gtid; v0 < tot; v0 += stride * 4) {
      IdT sid[4];
      int bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t v = v0 + u * stride;
        bb[u] = -1;
        if (v < tot) {
          bb[u] = batch_of(sh.off, B, v);
          sid[u] = ldcg(off_ptr(seeds0, (int64_t)bb[u] * in_stride) + (v - sh.off[bb[u]]));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (bb[u] >= 0)
          off_ptr(frontier0, (int64_t)bb[u] * out_stride)[v0 + u * stride - sh.off[bb[u]]] = sid[u];
    }
  }
  for (int64_t v = gtid; !unique_seeds && v < sh.off[B]; v += stride) {
    const int b = batch_of(sh.off, B, v);
    const unsigned int i = (unsigned int)(v - sh.off[b]);
    const IdT sid = ldcg(off_ptr(seeds0, (int64_t)b * in_stride) + i);
    IdT *frontier = off_ptr(frontier0, (int64_t)b * out_stride);
    const int64_t bo = (int64_t)b * ws_stride;
    if (ldcg(off_ptr(ws0.hop[0].fseed, bo) + i) == i) {       // first occurrence of this seed id
      const unsigned int pa = smem_pref ? sp[(size_t)b * 3 * tiles_ub + i / kBkTile]
                                        : (unsigned int)ldcg(off_ptr(ws0.prefA, bo) + i / kBkTile);
      frontier[pa + ldcg(off_ptr(ws0.hop[0].lrank_at, bo) + i)] = sid;
    }
  }
  // padded slots -> compacted, relabelled COO (+ frontier entries of first occurrences)
  if (k <= 0) return;
  const int64_t Etot = sh.off2[B];
  const unsigned int uk = (unsigned int)k;
  for (int64_t base = gtid; base < Etot; base += stride * EB) {
    int bb[EB];
    unsigned int ee[EB], si[EB];
    bool ok[EB];
    unsigned int cf[EB], cr[EB], sf[EB], sr[EB];
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      const int64_t v = base + u * stride;
      ok[u] = false;
      if (v < Etot) {
        const int b = batch_of(sh.off2, B, v);
        const unsigned int e = (unsigned int)(v - sh.off2[b]);   // < 2^32 (plan)
        const int64_t bo = (int64_t)b * ws_stride;
        bb[u] = b;
        ee[u] = e;
        si[u] = e / uk;
        ok[u] = (e - si[u] * uk) < (unsigned int)ldcg(off_ptr(ws0.hop[0].cnt, bo) + si[u]);
        // what the rank phase recorded: first-occurrence item of this slot's id (and of its seed)
        cf[u] = ldcg(off_ptr(ws0.hop[0].fslot, bo) + e);
        if (!unique_seeds) sf[u] = ldcg(off_ptr(ws0.hop[0].fseed, bo) + si[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      if (ok[u]) {
        const int64_t bo = (int64_t)bb[u] * ws_stride;
        // rank of that first occurrence inside its tile (not needed for a distinct seed: id = index)
        cr[u] = (unique_seeds && (int64_t)cf[u] < S_ub) ? 0u : ldcg(off_ptr(ws0.hop[0].lrank_at, bo) + cf[u]);
        if (!unique_seeds) sr[u] = ldcg(off_ptr(ws0.hop[0].lrank_at, bo) + sf[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      if (ok[u]) {
        const int b = bb[u];
        const int64_t bo = (int64_t)b * ws_stride;
        const unsigned int e = ee[u], s_i = si[u], j = e - s_i * uk;
        // exclusive tile prefixes of batch b: A (seed first occurrences), B (neighbour first
        // occurrences), C (edges); entry [tiles] of A = total A
        const unsigned int *pA = sp + (size_t)b * 3 * tiles_ub;
        const long long *gA = off_ptr(ws0.prefA, bo), *gB = off_ptr(ws0.prefB, bo),
                        *gC = off_ptr(ws0.prefC, bo);
        const int tiles = (int)((sh.S[b] + kBkTile - 1) / kBkTile);
        auto PA = [&](unsigned int t) { return smem_pref ? pA[t] : (unsigned int)ldcg(gA + t); };
        auto PB = [&](unsigned int t) { return smem_pref ? pA[tiles_ub + t] : (unsigned int)ldcg(gB + t); };
        auto PC = [&](unsigned int t) { return smem_pref ? pA[2 * tiles_ub + t] : (unsigned int)ldcg(gC + t); };
        // new id of the neighbour: first occurrence f among the seeds (< S_ub) or the slots
        const unsigned int f = cf[u];
        unsigned int cid;
        if ((int64_t)f < S_ub)
          cid = unique_seeds ? f : PA(f / kBkTile) + cr[u];
        else
          cid = PA(tiles) + PB((unsigned int)((f - (unsigned int)S_ub) / uk) / kBkTile) + cr[u];
        if ((int64_t)f == S_ub + (int64_t)e)      // this slot IS the first occurrence: it names the id
          off_ptr(frontier0, (int64_t)b * out_stride)[cid] =
              ldcg(reinterpret_cast<const IdT *>(off_ptr(reinterpret_cast<char *>(ws0.pad_col), bo)) + e);
        const unsigned int rid = unique_seeds ? s_i : PA(sf[u] / kBkTile) + sr[u];
        const unsigned int o = PC(s_i / kBkTile) + (unsigned int)ldcg(off_ptr(ws0.loff, bo) + s_i) + j;
        off_ptr(row0, (int64_t)b * out_stride)[o] = (IdT)rid;
        off_ptr(col0, (int64_t)b * out_stride)[o] = (IdT)cid;
      }
    }
  }
}

// One cooperative launch: all hops, phases separated by grid barriers (lowest latency; B = 1).
template <typename IdT, typename ET, int MODE>
__global__ void __launch_bounds__(kBkThreads, DGS_COOP_MIN_CTAS)
multi_batch_kernel(GraphSrc g, BlocksWs ws0, MbArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ MbShared sh;
  int nt = 0;
  auto stamp = [&]() {
    if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[nt++] = t;
    }
  };
  stamp();
  for (int l = 0; l < a.L; ++l) {
    if (a.scan_in_emit && l > 0) {
      __syncthreads();
      if (threadIdx.x == 0) sh.S[0] = min((long long)a.hop[l].S_ub, sh.next_S);
      __syncthreads();
    } else {
      mb_load_S(a, l, sh);
    }
    mb_pick<IdT, ET, MODE>(g, ws0, a, l, sh);
    if (mb_has_hubs<MODE>(a, l)) {   // (one more barrier, weighted mode only)
      grid.sync();
      mb_hubs<IdT>(ws0, a, l);
    }
    stamp();
    grid.sync();
    stamp();
    mb_rank<IdT>(ws0, a, l, sh);
    stamp();
    grid.sync();
    stamp();
    mb_emit<IdT, kEmBatch>(ws0, a, l, sh, dyn_smem, l + 1 == a.L);
    stamp();
    if (l + 1 < a.L) grid.sync();   // (nothing follows the last emit)
    stamp();
  }
}

// The same phases as separate kernels (B >= 2): a phase of B batches is throughput work, and a
// kernel per phase gets its own register budget - the rank and emit phases run at 4 CTAs per SM
// instead of the 2 the pick phase's registers allow the fused kernel; the kernel boundaries
// (~3 us each) are shared by the B batches.
#ifndef DGS_MB_PICK_CTAS
#define DGS_MB_PICK_CTAS 3  // measured at B = 8: 37.4 (2) / 36.3 (3) / 37.0 (4) us per batch
#endif
template <typename IdT, typename ET, int MODE, int FR>
__global__ void __launch_bounds__(kBkThreads, FR > 16 ? 2 : DGS_MB_PICK_CTAS)
mb_pick_kernel(GraphSrc g, BlocksWs ws0, MbArgs a, int l) {
  __shared__ MbShared sh;
  mb_load_S(a, l, sh);
  mb_pick<IdT, ET, MODE, FR>(g, ws0, a, l, sh);
}
template <typename IdT>
__global__ void __launch_bounds__(kBkThreads, 4) mb_hub_kernel(BlocksWs ws0, MbArgs a, int l) {
  mb_hubs<IdT>(ws0, a, l);
}
template <typename IdT>
__global__ void __launch_bounds__(kBkThreads, 4) mb_rank_kernel(BlocksWs ws0, MbArgs a, int l) {
  __shared__ MbShared sh;
  mb_load_S(a, l, sh);
  mb_rank<IdT>(ws0, a, l, sh);
}
template <typename IdT>
__global__ void __launch_bounds__(kBkThreads, 4) mb_emit_kernel(BlocksWs ws0, MbArgs a, int l) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ MbShared sh;
  mb_load_S(a, l, sh);
#ifndef DGS_EM_BATCH_SPLIT
#define DGS_EM_BATCH_SPLIT 4
#endif
  mb_emit<IdT, DGS_EM_BATCH_SPLIT>(ws0, a, l, sh, dyn_smem, l + 1 == a.L);
}

__global__ void blocks_ws_init_kernel(int4 *p, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = make_int4(-1, -1, -1, -1);
}

int build_graph_src(const dgs_graph_t *g, GraphSrc *out);  // sampling.cu
// ---------------------------------------------------------------------------------------------
// Workspace plans.  Which kernel a workspace serves is decided when it is sized:
//   kPathMulti  : multi_batch_kernel - direct tables (node count known), every fan-out fits the
//                 tile pick phase, <= 8 hops, cooperative launch available; B >= 1 batches;
//   kPathLegacy : fused_batch_kernel / the 3-kernels-per-hop path (hashed tables, fan-out 0 or
//                 huge, > 8 hops); one batch.
// dgs_sample_blocks*_ws_init registers the plan under the workspace pointer; every sampling call
// looks it up and refuses a workspace that was initialised for something else (a different
// layout would silently read the relabel tables at the wrong offsets).
enum { kPathLegacy = 0, kPathMulti = 1 };

struct WsInfo {
  int path, itype, L, B;
  int64_t S, num_nodes, bytes;
  int64_t fan[16];
  int uses;                  // table uses (hops) since the tables were last cleared (multi path)
  int64_t *counts_host;      // last pinned pointer asked about ...
  long long *counts_host_dev;  // ... and its device alias (null: not mapped)
};
static std::mutex g_ws_mu;
static std::unordered_map<const void *, WsInfo> g_ws;

struct MultiPlan {
  int64_t S_max, E_max, tiles_max, cap, table_bytes, stride, bytes;
};

static bool coop_launch_supported() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, v = 1;   // no device visible (sizing a workspace on a CPU-only box): assume B200
    if (cudaGetDevice(&dev) == cudaSuccess)
      if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) v = 0;
    cudaGetLastError();
    cached = v;
  }
  return cached == 1;
}

// can this configuration run on the multi-batch kernel?
static bool multi_path_ok(int L, const int64_t *fan_out, int64_t num_nodes, int64_t num_seeds) {
  static const bool no_direct = getenv("DGS_BLOCKS_HASHED") != nullptr;
  static const char *mode_env = getenv("DGS_BLOCKS_MODE");
  if (no_direct || (mode_env && strcmp(mode_env, "multi") == 0)) return false;
  if (L < 1 || L > 8 || !coop_launch_supported()) return false;
  if (!(num_nodes > 0 && num_nodes < (1ll << 32) - 2 && num_nodes <= (1ll << 29))) return false;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[l];
    // shared memory of the tile pick phase (weighted case): 128 k positions + 8 k keys
    if (k <= 0 || (size_t)kPkSeeds * k * sizeof(int) + (size_t)kBkWarps * k * sizeof(float) > 64 * 1024)
      return false;
  }
  // items of a hop (seeds + padded slots) must fit the 24 bits next to the epoch tag
  int64_t ub = num_seeds < 1 ? 1 : num_seeds;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[L - 1 - l];
    if (ub + ub * k >= (1ll << 24)) return false;
    ub += ub * k;
  }
  return true;
}

static int multi_plan(int itype, int B, int64_t num_seeds, int L, const int64_t *fan_out,
                      int64_t num_nodes, MultiPlan *p, char *base, BlocksWs *ws) {
  DGS_REQUIRE(B >= 1 && B <= DGS_MAX_BATCHES, "sample_blocks_multi: 1..%d batches per launch",
              DGS_MAX_BATCHES);
  int64_t ub = num_seeds < 1 ? 1 : num_seeds, S_max = 1, E_max = 1;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[L - 1 - l];
    const int64_t e = ub * k;
    DGS_REQUIRE(ub + e < (1ll << 32) - 2, "sample_blocks: more than 2^32 items in one hop");
    S_max = std::max(S_max, ub);
    E_max = std::max(E_max, e);
    ub += e;
  }
  const int idb = itype == DGS_I64 ? 8 : 4;
  p->S_max = S_max;
  p->E_max = E_max;
  p->tiles_max = (S_max + kBkTile - 1) / kBkTile;
  p->cap = (num_nodes + 3) & ~3ll;
  p->table_bytes = p->cap * 4;          // {tag:8 | first item:24} per node
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char *q = base ? base + off : nullptr;
    off += up256(bytes);
    return q;
  };
  char *done = take(256);
  char *pa = take((p->tiles_max + 1) * 8), *pb = take((p->tiles_max + 1) * 8),
       *pc = take((p->tiles_max + 1) * 8);
  char *cnt = take(S_max * 4);
  char *loff = take(S_max * 4);
  char *pad = take(E_max * idb);
  char *fslot = take(E_max * 4), *fseed = take(S_max * 4), *lrank_at = take((S_max + E_max) * 4);
  char *hub_ent = take(S_max * (int64_t)sizeof(HubEnt));
  char *hub_done = take(kHubMaxRows * 4), *hub_ckey = take((int64_t)kHubMaxChunks * 32 * 4),
       *hub_cidx = take((int64_t)kHubMaxChunks * 32 * 4), *hub_cnum = take(kHubMaxChunks * 4);
  char *t0 = take(p->table_bytes);
  p->stride = off;
  p->bytes = off * B;
  if (ws) {
    memset(ws, 0, sizeof(*ws));
    ws->done = (unsigned int *)done;
    ws->hubs.count = (unsigned int *)(done + 128);
    ws->hubs.ent = (HubEnt *)hub_ent;
    ws->hubs.done = (unsigned int *)hub_done;
    ws->hubs.ckey = (float *)hub_ckey;
    ws->hubs.cidx = (int *)hub_cidx;
    ws->hubs.cnum = (int *)hub_cnum;
    ws->pending_S = (long long *)(done + 64);
    ws->prefA = (long long *)pa;
    ws->prefB = (long long *)pb;
    ws->prefC = (long long *)pc;
    ws->loff = (int *)loff;
    ws->pad_col = pad;
    for (int b = 0; b < 2; ++b) {
      ws->hop[b].cnt = (int *)cnt;
      ws->hop[b].pos_seed = nullptr;
      ws->hop[b].pos_col = nullptr;
      ws->hop[b].table.base = t0;
      ws->hop[b].table.direct = 2;
      ws->hop[b].fslot = (unsigned int *)fslot;
      ws->hop[b].fseed = (unsigned int *)fseed;
      ws->hop[b].lrank_at = (unsigned int *)lrank_at;
    }
    ws->cap = p->cap;
    ws->direct = 1;
  }
  return 0;
}

template <typename IdT, typename ET>
static int launch_multi(const GraphSrc &src, int B, const IdT *seeds, int64_t seeds_stride,
                        int64_t num_seeds, int L, const int64_t *fan_out, int replace,
                        const uint64_t *rng_seeds, void *const *out_frontier, void *const *out_row,
                        void *const *out_col, int64_t out_stride, const int64_t *cap_edges,
                        const int64_t *cap_frontier, int64_t *counts_dev, const BlocksWs &ws,
                        int64_t ws_stride, long long *host_counts, unsigned int tag0, cudaStream_t st) {
  const bool bias = src.probs != nullptr || src.sh_probs.p[0] != nullptr;
  const int mode = bias ? (replace ? kBiasReplace : kBias) : (replace ? kUniformReplace : kUniform);
  MbArgs a;
  memset(&a, 0, sizeof(a));
  a.L = L;
  a.B = B;
  a.tag0 = tag0;
  a.seeds_stride = seeds_stride;
  a.out_stride = out_stride;
  a.ws_stride = ws_stride;
  a.counts_dev = (long long *)counts_dev;
  a.host_counts = host_counts;
  for (int b = 0; b < B; ++b) a.rng[b] = rng_seeds[b];
  size_t smem_max = 0;
  int64_t ub = num_seeds;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[L - 1 - l];
    DGS_REQUIRE(cap_edges[l] >= ub * k, "sample_blocks: layer %d edge capacity %lld < %lld", l,
                (long long)cap_edges[l], (long long)(ub * k));
    DGS_REQUIRE(cap_frontier[l] >= ub * (1 + k), "sample_blocks: layer %d frontier capacity %lld < %lld",
                l, (long long)cap_frontier[l], (long long)(ub * (1 + k)));
    MbHop &h = a.hop[l];
    h.seeds = l == 0 ? (const void *)seeds : out_frontier[l - 1];
    h.frontier = out_frontier[l];
    h.out_row = out_row[l];
    h.out_col = out_col[l];
    h.S_ub = ub;
    h.k = (int)k;
    const size_t sm_pick = (size_t)kPkSeeds * k * sizeof(int) + (mode == kBias ? (size_t)kBkWarps * k * sizeof(float) : 0);
    const size_t pref_bytes = (size_t)B * 3 * ((size_t)(ub + kBkTile - 1) / kBkTile + 1) * sizeof(unsigned int);
    static const bool no_smem_pref = getenv("DGS_BLOCKS_GLOBAL_PREFIX") != nullptr;
    h.smem_pref = (!no_smem_pref && pref_bytes <= 72 * 1024) ? 1 : 0;
    smem_max = std::max(smem_max, sm_pick);
    if (h.smem_pref) smem_max = std::max(smem_max, pref_bytes);
    ub += ub * k;
  }
  static const char *split_env = getenv("DGS_MB_SPLIT");   // "0" / "1": force one / many kernels
  // B >= 2, or one batch that is big by itself (friendster: 4096 seeds, [20,15,10] = 14 M padded slots
  // in the last hop): throughput work, better served by full-occupancy phase kernels than by the
  // latency-oriented cooperative kernel
  const bool big = (int64_t)B * a.hop[L - 1].S_ub * a.hop[L - 1].k >= (4ll << 20);
  const bool split = split_env ? split_env[0] == '1' : (B >= 2 || big);
  if (split) {
    // one kernel per phase (see mb_pick_kernel): 3 L launches shared by the B batches
    size_t pref_max = 0;
    for (int l = 0; l < L; ++l) {
      const size_t pref_bytes = (size_t)B * 3 * ((size_t)(a.hop[l].S_ub + kBkTile - 1) / kBkTile + 1) * sizeof(unsigned int);
      a.hop[l].smem_pref = pref_bytes <= 48 * 1024 ? 1 : 0;     // 4 CTAs per SM
      if (a.hop[l].smem_pref) pref_max = std::max(pref_max, pref_bytes);
    }
    auto kemit = mb_emit_kernel<IdT>;
    static bool emit_attr_set = false;
    if (!emit_attr_set) {
      DGS_CUDA_OK(cudaFuncSetAttribute(kemit, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
      emit_attr_set = true;
    }
    const int sms = sm_count();
    for (int l = 0; l < L; ++l) {
      const int64_t k = a.hop[l].k, S_ub = a.hop[l].S_ub;
      const size_t sm_pick = (size_t)kPkSeeds * k * sizeof(int) + (mode == kBias ? (size_t)kBkWarps * k * sizeof(float) : 0);
#define DGS_MBP_(M, FRV, CTAS)                                                                       \
  do {                                                                                               \
    auto kp = mb_pick_kernel<IdT, ET, M, FRV>;                                                       \
    if (sm_pick > 32 * 1024)                                                                         \
      DGS_CUDA_OK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pick)); \
    kp<<<sms * (CTAS), kBkThreads, sm_pick, st>>>(src, ws, a, l);                                    \
  } while (0)
#define DGS_MBP(M) DGS_MBP_(M, 16, DGS_MB_PICK_CTAS)
      switch (mode) {
        case kUniform:   // fan-outs 17..32: Floyd in 32 registers (2 CTAs per SM)
          if (k > 16 && k <= 32) DGS_MBP_(kUniform, 32, 2); else DGS_MBP(kUniform);
          break;
        case kUniformReplace: DGS_MBP(kUniformReplace); break;
        case kBias: DGS_MBP(kBias); break;
        default: DGS_MBP(kBiasReplace); break;
      }
#undef DGS_MBP
#undef DGS_MBP_
      DGS_LAUNCH_CHECK();
      if (mode == kBias && k <= 32) {   // the long rows the pick tiles deferred
        mb_hub_kernel<IdT><<<sms * 4, kBkThreads, 0, st>>>(ws, a, l);
        DGS_LAUNCH_CHECK();
      }
      const int64_t tiles = (int64_t)B * ((S_ub + kBkTile - 1) / kBkTile);
      mb_rank_kernel<IdT><<<(int)std::min<int64_t>(tiles, (int64_t)sms * 4), kBkThreads, 0, st>>>(ws, a, l);
      DGS_LAUNCH_CHECK();
      const int64_t items = (int64_t)B * (S_ub + S_ub * k);
      const int64_t ge = (items + kBkThreads * 4 - 1) / (kBkThreads * 4);
      kemit<<<(int)std::max<int64_t>(1, std::min<int64_t>(ge, (int64_t)sms * 4)), kBkThreads,
              a.hop[l].smem_pref ? pref_max : 0, st>>>(ws, a, l);
      DGS_LAUNCH_CHECK();
    }
    return 0;
  }
  {
    static const bool no_scan = getenv("DGS_MB_TAIL") != nullptr;   // force the last-CTA tail (A/B runs)
    bool all_pref = true;
    for (int l = 0; l < L; ++l) all_pref = all_pref && a.hop[l].smem_pref;
    a.scan_in_emit = (B == 1 && all_pref && !no_scan) ? 1 : 0;
  }
  void *kern = nullptr;
#define DGS_MB(M) kern = (void *)multi_batch_kernel<IdT, ET, M>
  switch (mode) {
    case kUniform: DGS_MB(kUniform); break;
    case kUniformReplace: DGS_MB(kUniformReplace); break;
    case kBias: DGS_MB(kBias); break;
    default: DGS_MB(kBiasReplace); break;
  }
#undef DGS_MB
  if (smem_max > 32 * 1024)  // static shared memory (~12 KB) counts against the 48 KB default
    DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  int per_sm = 0;
  DGS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBkThreads, smem_max));
  DGS_REQUIRE(per_sm >= 1, "sample_blocks: the batch kernel does not fit an SM (%zu bytes of shared memory)",
              smem_max);
  if (per_sm > 4) per_sm = 4;
  const int grid = sm_count() * per_sm;
  static const bool trace = getenv("DGS_BLOCKS_TRACE") != nullptr;
  static unsigned long long *trace_dev = nullptr;
  if (trace && !trace_dev) cudaMalloc(&trace_dev, 1024 * sizeof(unsigned long long));
  a.trace = trace ? trace_dev : nullptr;
  GraphSrc src_copy = src;
  BlocksWs ws_copy = ws;
  void *params[] = {&src_copy, &ws_copy, &a};
  DGS_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kBkThreads), params, smem_max, st));
  dgsb::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (trace) {
    unsigned long long h[2 + 6 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_dev, sizeof(unsigned long long) * (1 + 6 * L), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[dgs multi trace us, grid %d, B %d] pick sync rank sync emit sync:", grid, B);
    for (int i = 1; i < 1 + 6 * L; ++i)
      fprintf(stderr, "%s%.1f", (i - 1) % 6 == 0 ? " | " : " ", (double)(h[i] - h[i - 1]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return 0;
}


template <typename IdT, typename ET>
static int launch_blocks(const GraphSrc &src, const IdT *seeds, int64_t num_seeds, int L,
                         const int64_t *fan_out, int replace, uint64_t rng_seed, int64_t epoch,
                         void *const *out_frontier, void *const *out_row, void *const *out_col,
                         const int64_t *cap_edges, const int64_t *cap_frontier,
                         int64_t *counts_dev, const BlocksWs &ws, cudaStream_t st,
                         long long *host_counts, bool *host_counts_used) {
  *host_counts_used = false;
  const bool bias = src.probs != nullptr || src.sh_probs.p[0] != nullptr;
  const int mode = bias ? (replace ? kBiasReplace : kBias) : (replace ? kUniformReplace : kUniform);
  const uint64_t cap_mask = (uint64_t)ws.cap - 1;
  // upper bounds of every hop (the previous call on this workspace had the same ones)
  int64_t ubs[16];
  {
    int64_t ub = num_seeds;
    for (int l = 0; l < L; ++l) {
      ubs[l] = ub;
      ub += ub * fan_out[L - 1 - l];
    }
  }
  // shared memory of the tile pick phase: 128 k positions (+ 8 k keys when biased)
  size_t smem_max = 0;
  bool all_tile = true;
  for (int l = 0; l < L; ++l) {
    const int64_t k = fan_out[L - 1 - l];
    size_t sm = (size_t)kPkSeeds * k * sizeof(int) + (mode == kBias ? (size_t)kBkWarps * k * sizeof(float) : 0);
    DGS_REQUIRE(cap_edges[l] >= ubs[l] * k, "sample_blocks: layer %d edge capacity %lld < %lld", l,
                (long long)cap_edges[l], (long long)(ubs[l] * k));
    DGS_REQUIRE(cap_frontier[l] >= ubs[l] * (1 + k),
                "sample_blocks: layer %d frontier capacity %lld < %lld", l,
                (long long)cap_frontier[l], (long long)(ubs[l] * (1 + k)));
    if (k <= 0 || sm > 64 * 1024) all_tile = false;
    if (sm > smem_max) smem_max = sm;
  }
  static const char *mode_env = getenv("DGS_BLOCKS_MODE");  // "multi" forces the 3-kernels-per-hop path
  static int coop_supported = -1;
  if (coop_supported < 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess)
      v = 0;
    cudaGetLastError();
    coop_supported = v;
  }
  const bool want_coop = coop_supported == 1 && !(mode_env && strcmp(mode_env, "multi") == 0);

  // ---- cooperative single-launch path
  if (want_coop && all_tile && L <= 8) {
    BatchArgs a;
    memset(&a, 0, sizeof(a));
    a.L = L;
    a.cap_mask = cap_mask;
    for (int l = 0; l < L; ++l) {
      HopArgs &h = a.hop[l];
      const int pl = l > 0 ? l - 1 : L - 1;
      h.seeds = l == 0 ? (const void *)seeds : out_frontier[l - 1];
      h.frontier = out_frontier[l];
      h.out_row = out_row[l];
      h.out_col = out_col[l];
      h.nnz_dev = (long long *)(counts_dev + 2 * l);
      h.nf_dev = (long long *)(counts_dev + 2 * l + 1);
      h.S_dev = l == 0 ? nullptr : (const long long *)(counts_dev + 2 * (l - 1) + 1);
      h.S_ub = ubs[l];
      h.prev_S_ub = ubs[pl];
      h.key = rng_seed + 0x9E3779B97F4A7C15ull * (uint64_t)(l + 1);
      h.k = (int)fan_out[L - 1 - l];
      h.prev_k = (int)fan_out[L - 1 - pl];
      h.cur = (int)((epoch * L + l) & 1);
      const size_t pref_bytes = 3 * ((size_t)(ubs[l] + kBkTile - 1) / kBkTile + 1) * sizeof(unsigned int);
      static const bool no_smem_pref = getenv("DGS_BLOCKS_GLOBAL_PREFIX") != nullptr;
      h.smem_pref = (!no_smem_pref && pref_bytes <= 64 * 1024) ? 1 : 0;
      if (h.smem_pref && pref_bytes > smem_max) smem_max = pref_bytes;
    }
    void *kern = nullptr;
#define DGS_BK(M) kern = (void *)fused_batch_kernel<IdT, ET, M>
    switch (mode) {
      case kUniform: DGS_BK(kUniform); break;
      case kUniformReplace: DGS_BK(kUniformReplace); break;
      case kBias: DGS_BK(kBias); break;
      default: DGS_BK(kBiasReplace); break;
    }
#undef DGS_BK
    if (smem_max > 32 * 1024)  // static shared memory (~8 KB) counts against the 48 KB default
      DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    int per_sm = 0;
    DGS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBkThreads, smem_max));
    if (per_sm >= 1) {
      if (per_sm > 4) per_sm = 4;
      const int grid = sm_count() * per_sm;
      GraphSrc src_copy = src;
      BlocksWs ws_copy = ws;
      static const bool trace = getenv("DGS_BLOCKS_TRACE") != nullptr;
      static unsigned long long *trace_dev = nullptr;
      if (trace && !trace_dev) cudaMalloc(&trace_dev, 1024 * sizeof(unsigned long long));
      a.trace = trace ? trace_dev : nullptr;
      a.host_counts = host_counts;
      *host_counts_used = host_counts != nullptr;
      void *params[] = {&src_copy, &ws_copy, &a};
      DGS_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kBkThreads), params, smem_max, st));
      dgsb::g_launches.fetch_add(1, std::memory_order_relaxed);
      if (trace) {
        unsigned long long h[1 + 8 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_dev, sizeof(unsigned long long) * (1 + 7 * L), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[dgs coop trace us, grid %d] pick sync rank tail sync emit sync:", grid);
        for (int i = 1; i < 1 + 7 * L; ++i)
          fprintf(stderr, "%s%.1f", (i - 1) % 7 == 0 ? " | " : " ", (double)(h[i] - h[i - 1]) * 1e-3);
        fprintf(stderr, "\n");
        unsigned long long f[8 * 8];
        cudaMemcpy(f, trace_dev + 256, sizeof(f), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[dgs coop pick detail us: A B1 B2 tail(other tiles)]");
        for (int l = 0; l < L; ++l)
          fprintf(stderr, " | %.1f %.1f %.1f %.1f", (double)(f[8 * l + 1] - f[8 * l]) * 1e-3,
                  (double)(f[8 * l + 2] - f[8 * l + 1]) * 1e-3, (double)(f[8 * l + 3] - f[8 * l + 2]) * 1e-3,
                  (double)(f[8 * l + 4] - f[8 * l + 3]) * 1e-3);
        fprintf(stderr, "\n");
      }
      return 0;
    }
  }

  // ---- multi-kernel path
  static const bool timing = getenv("DGS_BLOCKS_TIMING") != nullptr;
  cudaEvent_t evs[3 * 16 + 1];
  int nev = 0;
  auto mark = [&]() {
    if (timing) {
      cudaEventCreate(&evs[nev]);
      cudaEventRecord(evs[nev], st);
      ++nev;
    }
  };
  mark();
  const IdT *cur_seeds = seeds;
  const int64_t *cur_dev = nullptr;
  for (int l = 0; l < L; ++l) {
    const int64_t k64 = fan_out[L - 1 - l];  // walked from the back (sampler.cc:20)
    const int k = (int)k64;
    const int64_t cur_ub = ubs[l];
    const int64_t nnz_ub = cur_ub * k64;
    // Hops are numbered h = epoch L + l over the life of the workspace; hop h uses state h & 1 and
    // wipes what hop h - 1 left in the other one (for l = 0 that is the last hop of the previous
    // call, whose live seed count the emit phase parked in ws.pending_S; 0 after ws_init).
    const int64_t h = epoch * L + l;
    const HopState &cur = ws.hop[h & 1];
    const HopState &prev = ws.hop[(h + 1) & 1];
    const int pl = l > 0 ? l - 1 : L - 1;
    const int64_t prev_ub = ubs[pl];
    const int prev_k = (int)fan_out[L - 1 - pl];
    const long long *prev_dev = ws.pending_S;
    int64_t *nnz_dev = counts_dev + 2 * l, *nf_dev = counts_dev + 2 * l + 1;
    const uint64_t key = rng_seed + 0x9E3779B97F4A7C15ull * (uint64_t)(l + 1);
    size_t smem_tile = (size_t)kPkSeeds * k * sizeof(int);
    if (mode == kBias) smem_tile += (size_t)kBkWarps * k * sizeof(float);
    if (k > 0 && smem_tile <= 64 * 1024) {
      const int grid = std::max(grid_for(cur_ub, 8, 2),  // tiles shrink to 8 seeds on small hops
                                grid_for(prev_ub * (1 + (int64_t)prev_k), kBkThreads * 4, 4));
#define DGS_TPICK(M)                                                                            \
  do {                                                                                          \
    auto kern = fused_pick_tile_kernel<IdT, ET, M>;                                             \
    if (smem_tile > 32 * 1024)                                                                  \
      DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                       (int)smem_tile));                                        \
    kern<<<grid, kBkThreads, smem_tile, st>>>(src, cur_seeds, cur_ub, cur_dev, k, key,          \
                                              (IdT *)ws.pad_col, cur, cap_mask, prev, prev_ub,  \
                                              prev_dev, prev_k);                                \
  } while (0)
      switch (mode) {
        case kUniform: DGS_TPICK(kUniform); break;
        case kUniformReplace: DGS_TPICK(kUniformReplace); break;
        case kBias: DGS_TPICK(kBias); break;
        default: DGS_TPICK(kBiasReplace); break;
      }
#undef DGS_TPICK
    } else {
      size_t smem = 0;
      if (k > 32 && mode == kUniform) smem = (size_t)kBkWarps * k * sizeof(int);
      if (k > 0 && mode == kBias) smem = (size_t)kBkWarps * k * 2 * sizeof(float);
      int gmem_scratch = 0;
      if (smem > 96 * 1024) {
        smem = 0;
        gmem_scratch = 1;
      }
      const int grid_pick = std::max(grid_for(cur_ub, kBkWarps, 8),
                                     grid_for(prev_ub * (1 + (int64_t)prev_k), kBkThreads * 4, 8));
#define DGS_FPICK(M)                                                                            \
  do {                                                                                          \
    auto kern = fused_pick_kernel<IdT, ET, M>;                                                  \
    if (smem > 32 * 1024)                                                                       \
      DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                       (int)smem));                                             \
    kern<<<grid_pick, kBkThreads, smem, st>>>(src, cur_seeds, cur_ub, cur_dev, k, key,          \
                                              (IdT *)ws.pad_col, cur, cap_mask, prev, prev_ub,  \
                                              prev_dev, prev_k, gmem_scratch);                  \
  } while (0)
      switch (mode) {
        case kUniform: DGS_FPICK(kUniform); break;
        case kUniformReplace: DGS_FPICK(kUniformReplace); break;
        case kBias: DGS_FPICK(kBias); break;
        default: DGS_FPICK(kBiasReplace); break;
      }
#undef DGS_FPICK
    }
    DGS_LAUNCH_CHECK();
    mark();
    const int grid_rank = grid_for(cur_ub, kBkTile, 8);
    fused_rank_kernel<<<grid_rank, kBkThreads, 0, st>>>(cur_ub, cur_dev, k, cur, ws, nnz_dev, nf_dev,
                                                        l > 0 ? 1 : 0);
    DGS_LAUNCH_CHECK();
    mark();
    const int grid_emit = grid_for(cur_ub + nnz_ub, kBkThreads * 2, 8);
    fused_emit_kernel<IdT><<<grid_emit, kBkThreads, 0, st>>>(
        cur_seeds, cur_ub, cur_dev, k, (const IdT *)ws.pad_col, cur, ws, (IdT *)out_frontier[l],
        (IdT *)out_row[l], (IdT *)out_col[l], l > 0 ? 1 : 0);
    DGS_LAUNCH_CHECK();
    mark();
    cur_seeds = (const IdT *)out_frontier[l];
    cur_dev = nf_dev;
  }
  if (timing) {
    cudaStreamSynchronize(st);
    fprintf(stderr, "[dgs blocks timing us]");
    for (int i = 1; i < nev; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, evs[i - 1], evs[i]);
      fprintf(stderr, " %s%.1f", (i - 1) % 3 == 0 ? "| " : "", ms * 1e3f);
    }
    fprintf(stderr, "\n");
    for (int i = 0; i < nev; ++i) cudaEventDestroy(evs[i]);
  }
  return 0;
}

}  // namespace dgsb

using namespace dgsb;

static void ws_register(const void *ws, int path, int itype, int B, int64_t S, int L,
                        const int64_t *fan_out, int64_t num_nodes, int64_t bytes) {
  WsInfo w;
  memset(&w, 0, sizeof(w));
  w.path = path;
  w.itype = itype;
  w.B = B;
  w.S = S;
  w.L = L;
  w.num_nodes = num_nodes;
  w.bytes = bytes;
  for (int l = 0; l < L && l < 16; ++l) w.fan[l] = fan_out[l];
  std::lock_guard<std::mutex> lk(g_ws_mu);
  g_ws[ws] = w;
}

// The plan a workspace was initialised for; checks the call against it.  exact_seeds: the legacy
// layout depends on the seed count itself; the multi layout only needs num_seeds <= the init value.
static int ws_lookup(const void *ws, int itype, int B, int64_t num_seeds, int L, const int64_t *fan_out,
                     int64_t num_nodes, int64_t ws_bytes, WsInfo *out) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws.find(ws);
  DGS_REQUIRE(it != g_ws.end(), "sample_blocks: workspace %p was never initialised with "
              "dgs_sample_blocks_ws_init / dgs_sample_blocks_multi_ws_init", ws);
  const WsInfo &w = it->second;
  bool same = w.itype == itype && w.L == L && w.num_nodes == num_nodes && B <= w.B && B >= 1;
  for (int l = 0; same && l < L; ++l) same = w.fan[l] == fan_out[l];
  same = same && (w.path == kPathMulti ? num_seeds <= w.S : num_seeds == w.S);
  DGS_REQUIRE(same, "sample_blocks: workspace was initialised for another configuration (itype %d, "
              "%d batch(es) of %lld seeds, %d layers, %lld nodes) - the call asks for itype %d, %d x "
              "%lld seeds, %d layers, %lld nodes or another fan-out", w.itype, w.B, (long long)w.S,
              w.L, (long long)w.num_nodes, itype, B, (long long)num_seeds, L, (long long)num_nodes);
  DGS_REQUIRE(ws_bytes >= w.bytes, "sample_blocks: workspace %lld < %lld bytes", (long long)ws_bytes,
              (long long)w.bytes);
  *out = w;
  return 0;
}

// Device alias of a pinned + mapped host buffer (null when the kernel cannot write it); the answer
// is remembered per workspace.
static long long *mapped_alias(const void *ws, int64_t *counts_host) {
  static const bool no_direct = getenv("DGS_BLOCKS_COUNTS_MEMCPY") != nullptr;
  if (no_direct || counts_host == nullptr) return nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto it = g_ws.find(ws);
    if (it != g_ws.end() && it->second.counts_host == counts_host) return it->second.counts_host_dev;
  }
  cudaPointerAttributes at;
  void *dp = nullptr;
  long long *alias = nullptr;
  if (cudaPointerGetAttributes(&at, counts_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
      cudaHostGetDevicePointer(&dp, counts_host, 0) == cudaSuccess)
    alias = (long long *)dp;
  cudaGetLastError();
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws.find(ws);
  if (it != g_ws.end()) {
    it->second.counts_host = counts_host;
    it->second.counts_host_dev = alias;
  }
  return alias;
}

extern "C" int64_t dgs_sample_blocks_multi_ws_bytes(int itype, int num_batches, int64_t num_seeds,
                                                    int num_layers, const int64_t *fan_out,
                                                    int64_t num_nodes) {
  if (!fan_out || !multi_path_ok(num_layers, fan_out, num_nodes, num_seeds)) {
    set_error("sample_blocks_multi: this configuration needs the single-batch path (unknown node "
              "count, fan-out 0 / too large for the tile phase, > 8 hops or no cooperative launch)");
    return -1;
  }
  MultiPlan p;
  if (multi_plan(itype, num_batches, num_seeds, num_layers, fan_out, num_nodes, &p, nullptr, nullptr))
    return -1;
  return p.bytes;
}

extern "C" int dgs_sample_blocks_multi_ws_init(void *ws, int64_t ws_bytes, int itype, int num_batches,
                                               int64_t num_seeds, int num_layers,
                                               const int64_t *fan_out, int64_t num_nodes,
                                               void *stream) {
  DGS_REQUIRE(ws && fan_out, "dgs_sample_blocks_multi_ws_init: null argument");
  DGS_REQUIRE(multi_path_ok(num_layers, fan_out, num_nodes, num_seeds),
              "dgs_sample_blocks_multi_ws_init: configuration not supported by the multi-batch kernel");
  MultiPlan p;
  BlocksWs w;
  if (multi_plan(itype, num_batches, num_seeds, num_layers, fan_out, num_nodes, &p, (char *)ws, &w))
    return 1;
  DGS_REQUIRE(ws_bytes >= p.bytes, "dgs_sample_blocks_multi_ws_init: workspace %lld < %lld bytes",
              (long long)ws_bytes, (long long)p.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < num_batches; ++b) {
    char *base = (char *)ws + (int64_t)b * p.stride;
    DGS_CUDA_OK(cudaMemsetAsync(base, 0, 256, st));
    DGS_CUDA_OK(cudaMemsetAsync((char *)w.hubs.done + (int64_t)b * p.stride, 0, kHubMaxRows * 4, st));
    DGS_CUDA_OK(cudaMemsetAsync(w.hop[0].table.base + (int64_t)b * p.stride, 0xFF, (size_t)p.table_bytes, st));
  }
  ws_register(ws, kPathMulti, itype, num_batches, num_seeds, num_layers, fan_out, num_nodes, p.bytes);
  return 0;
}

extern "C" int64_t dgs_sample_blocks_ws_bytes(int itype, int64_t num_seeds, int num_layers,
                                              const int64_t *fan_out, int64_t num_nodes) {
  if (!fan_out) return -1;
  if (multi_path_ok(num_layers, fan_out, num_nodes, num_seeds))
    return dgs_sample_blocks_multi_ws_bytes(itype, 1, num_seeds, num_layers, fan_out, num_nodes);
  BlocksPlan p;
  if (blocks_plan(itype, num_seeds, num_layers, fan_out, num_nodes, &p, nullptr, nullptr)) return -1;
  return p.bytes;
}

extern "C" int dgs_sample_blocks_ws_init(void *ws, int64_t ws_bytes, int itype, int64_t num_seeds,
                                         int num_layers, const int64_t *fan_out, int64_t num_nodes,
                                         void *stream) {
  DGS_REQUIRE(ws && fan_out, "dgs_sample_blocks_ws_init: null argument");
  if (multi_path_ok(num_layers, fan_out, num_nodes, num_seeds))
    return dgs_sample_blocks_multi_ws_init(ws, ws_bytes, itype, 1, num_seeds, num_layers, fan_out,
                                           num_nodes, stream);
  BlocksPlan p;
  BlocksWs w;
  if (blocks_plan(itype, num_seeds, num_layers, fan_out, num_nodes, &p, (char *)ws, &w)) return 1;
  DGS_REQUIRE(ws_bytes >= p.bytes, "dgs_sample_blocks_ws_init: workspace %lld < %lld bytes",
              (long long)ws_bytes, (long long)p.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  DGS_CUDA_OK(cudaMemsetAsync(w.done, 0, 256, st));
  for (int b = 0; b < 2; ++b) {
    blocks_ws_init_kernel<<<grid_for(p.table_bytes / 16, 256, 8), 256, 0, st>>>(
        (int4 *)w.hop[b].table.base, p.table_bytes / 16);
    DGS_LAUNCH_CHECK();
  }
  ws_register(ws, kPathLegacy, itype, 1, num_seeds, num_layers, fan_out, num_nodes, p.bytes);
  return 0;
}

// Wait until the hop sizes of the batch enqueued with counts_host have arrived there.
static int counts_wait(int64_t *counts_host, const int64_t *counts_dev, int64_t n_counts, bool by_kernel,
                       cudaStream_t st) {
  if (by_kernel) {
    volatile int64_t *flag = counts_host;
    unsigned int spins = 0;
    while (*flag == -1) {
      if ((++spins & 0x3ffu) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) break;
    }
    if (*flag != -1) return 0;
    // the stream drained (or failed) without the sizes arriving: take the copy path, which also
    // reports a kernel fault
  }
  DGS_CUDA_OK(cudaMemcpyAsync(counts_host, counts_dev, sizeof(int64_t) * n_counts,
                              cudaMemcpyDeviceToHost, st));
  DGS_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

static int sample_multi_impl(const dgs_graph_t *g, int B, const void *seeds, int64_t seeds_stride,
                             int64_t num_seeds, int num_layers, const int64_t *fan_out, int replace,
                             const uint64_t *rng_seeds, void *const *out_frontier,
                             void *const *out_row, void *const *out_col, int64_t out_stride,
                             const int64_t *cap_edges, const int64_t *cap_frontier,
                             int64_t *counts_dev, void *ws, int64_t ws_bytes, int64_t *counts_host,
                             cudaStream_t st, bool wait, const WsInfo &info) {
  MultiPlan p;
  BlocksWs w;
  // the layout is the one the workspace was initialised with (num_seeds may be smaller now)
  if (multi_plan(g->itype, info.B, info.S, num_layers, fan_out, g->num_nodes, &p, (char *)ws, &w)) return 1;
  GraphSrc src;
  if (build_graph_src(g, &src)) return 1;
  long long *host_dev = mapped_alias(ws, counts_host);
  if (counts_host && (host_dev || !wait)) *(volatile int64_t *)counts_host = -1;   // sentinel: sizes are >= 0
  // epoch tags: this call uses the table num_layers times; tags count DOWN from 0xFE so that a
  // newer entry beats every stale one under atomicMin (0xFF = the cleared state).  Out of tags:
  // clear the tables (a memset every 255 / L calls) and start over.
  unsigned int tag0;
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsInfo &reg = g_ws[ws];
    if (reg.uses + num_layers > 255) {
      for (int b = 0; b < info.B; ++b)
        DGS_CUDA_OK(cudaMemsetAsync(w.hop[0].table.base + (int64_t)b * p.stride, 0xFF,
                                    (size_t)p.table_bytes, st));
      reg.uses = 0;
    }
    tag0 = 0xFEu - (unsigned int)reg.uses;
    reg.uses += num_layers;
  }
  int rc = 0;
  DGS_ITYPE_SWITCH(g->itype, IdT, {
    DGS_ITYPE_SWITCH(g->etype, ET, {
      rc = launch_multi<IdT, ET>(src, B, (const IdT *)seeds, seeds_stride, num_seeds, num_layers, fan_out,
                                 replace, rng_seeds, out_frontier, out_row, out_col, out_stride,
                                 cap_edges, cap_frontier, counts_dev, w, p.stride, host_dev, tag0, st);
    });
  });
  if (rc) return rc;
  if (counts_host && wait)
    return counts_wait(counts_host, counts_dev, (int64_t)2 * num_layers * B, host_dev != nullptr, st);
  return 0;
}

static int sample_blocks_impl(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                              int num_layers, const int64_t *fan_out, int replace,
                              uint64_t rng_seed, void *const *out_frontier,
                              void *const *out_row, void *const *out_col,
                              const int64_t *cap_edges, const int64_t *cap_frontier,
                              int64_t *counts_dev, void *ws, int64_t ws_bytes, int64_t epoch,
                              int64_t *counts_host, void *stream, bool wait) {
  DGS_REQUIRE(g && fan_out && out_frontier && out_row && out_col && cap_edges && cap_frontier &&
                  counts_dev && ws,
              "dgs_sample_blocks: null argument");
  DGS_REQUIRE(num_seeds >= 0 && epoch >= 0, "dgs_sample_blocks: negative seed count / epoch");
  cudaStream_t st = (cudaStream_t)stream;
  if (num_seeds == 0) {
    DGS_REQUIRE(num_layers >= 1 && num_layers <= 16, "dgs_sample_blocks: 1..16 layers supported");
    DGS_CUDA_OK(cudaMemsetAsync(counts_dev, 0, sizeof(int64_t) * 2 * num_layers, st));
    if (counts_host) {
      memset(counts_host, 0, sizeof(int64_t) * 2 * num_layers);
      if (wait) DGS_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
  }
  DGS_REQUIRE(seeds != nullptr, "dgs_sample_blocks: null seeds");
  DGS_REQUIRE(num_layers >= 1 && num_layers <= 16, "dgs_sample_blocks: 1..16 layers supported");
  WsInfo info;
  if (ws_lookup(ws, g->itype, 1, num_seeds, num_layers, fan_out, g->num_nodes, ws_bytes, &info)) return 1;
  if (info.path == kPathMulti)
    return sample_multi_impl(g, 1, seeds, 0, num_seeds, num_layers, fan_out, replace, &rng_seed,
                             out_frontier, out_row, out_col, 0, cap_edges, cap_frontier, counts_dev, ws,
                             ws_bytes, counts_host, st, wait, info);
  BlocksPlan p;
  BlocksWs w;
  if (blocks_plan(g->itype, num_seeds, num_layers, fan_out, g->num_nodes, &p, (char *)ws, &w)) return 1;
  GraphSrc src;
  if (build_graph_src(g, &src)) return 1;
  // Can the kernel write the hop sizes into counts_host itself?  (pinned + mapped host memory)
  long long *host_dev = mapped_alias(ws, counts_host);
  if (counts_host && (host_dev || !wait)) *(volatile int64_t *)counts_host = -1;   // sentinel: hop sizes are >= 0
  int rc = 0;
  bool by_kernel = false;
  DGS_ITYPE_SWITCH(g->itype, IdT, {
    DGS_ITYPE_SWITCH(g->etype, ET, {
      rc = launch_blocks<IdT, ET>(src, (const IdT *)seeds, num_seeds, num_layers, fan_out, replace,
                                  rng_seed, epoch, out_frontier, out_row, out_col, cap_edges,
                                  cap_frontier, counts_dev, w, st, host_dev, &by_kernel);
    });
  });
  if (rc) return rc;
  if (counts_host) {
    // (not delivered by the kernel - unmapped memory or the multi-kernel path: the wait call
    // finds the sentinel untouched once the stream has drained and copies the sizes itself)
    if (!wait) return 0;
    // the one host round trip of the batch: hop sizes -> (pinned) host memory
    return counts_wait(counts_host, counts_dev, (int64_t)2 * num_layers, by_kernel, st);
  }
  return 0;
}

extern "C" int dgs_sample_blocks(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                                 int num_layers, const int64_t *fan_out, int replace,
                                 uint64_t rng_seed, void *const *out_frontier,
                                 void *const *out_row, void *const *out_col,
                                 const int64_t *cap_edges, const int64_t *cap_frontier,
                                 int64_t *counts_dev, void *ws, int64_t ws_bytes, int64_t epoch,
                                 int64_t *counts_host, void *stream) {
  return sample_blocks_impl(g, seeds, num_seeds, num_layers, fan_out, replace, rng_seed, out_frontier,
                            out_row, out_col, cap_edges, cap_frontier, counts_dev, ws, ws_bytes, epoch,
                            counts_host, stream, true);
}

// Same, but returns right after the launch: the kernel will deliver the hop sizes to counts_host
// (mapped pinned host memory, required) and dgs_sample_blocks_wait collects them - in between the
// caller can enqueue the work that follows the sampling (dgs.classes.BatchLoader: extract + labels).
extern "C" int dgs_sample_blocks_enqueue(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                                         int num_layers, const int64_t *fan_out, int replace,
                                         uint64_t rng_seed, void *const *out_frontier,
                                         void *const *out_row, void *const *out_col,
                                         const int64_t *cap_edges, const int64_t *cap_frontier,
                                         int64_t *counts_dev, void *ws, int64_t ws_bytes,
                                         int64_t epoch, int64_t *counts_host, void *stream) {
  DGS_REQUIRE(counts_host != nullptr, "dgs_sample_blocks_enqueue: counts_host is required");
  return sample_blocks_impl(g, seeds, num_seeds, num_layers, fan_out, replace, rng_seed, out_frontier,
                            out_row, out_col, cap_edges, cap_frontier, counts_dev, ws, ws_bytes, epoch,
                            counts_host, stream, false);
}

extern "C" int dgs_sample_blocks_wait(int64_t *counts_host, const int64_t *counts_dev, int num_layers,
                                      void *stream) {
  DGS_REQUIRE(counts_host && counts_dev && num_layers >= 1, "dgs_sample_blocks_wait: bad argument");
  return counts_wait(counts_host, counts_dev, (int64_t)2 * num_layers, true, (cudaStream_t)stream);
}

extern "C" int dgs_sample_blocks_multi(const dgs_graph_t *g, int num_batches, const void *seeds,
                                       int64_t seeds_stride_bytes, int64_t num_seeds, int num_layers,
                                       const int64_t *fan_out, int replace, const uint64_t *rng_seeds,
                                       void *const *out_frontier, void *const *out_row,
                                       void *const *out_col, int64_t out_stride_bytes,
                                       const int64_t *cap_edges, const int64_t *cap_frontier,
                                       int64_t *counts_dev, void *ws, int64_t ws_bytes,
                                       int64_t *counts_host, int wait, void *stream) {
  DGS_REQUIRE(g && seeds && fan_out && rng_seeds && out_frontier && out_row && out_col && cap_edges &&
                  cap_frontier && counts_dev && ws,
              "dgs_sample_blocks_multi: null argument");
  DGS_REQUIRE(num_batches >= 1 && num_batches <= DGS_MAX_BATCHES,
              "dgs_sample_blocks_multi: 1..%d batches per launch", DGS_MAX_BATCHES);
  DGS_REQUIRE(num_seeds >= 1, "dgs_sample_blocks_multi: empty batches");
  DGS_REQUIRE(num_layers >= 1 && num_layers <= 8, "dgs_sample_blocks_multi: 1..8 layers supported");
  DGS_REQUIRE(wait == 0 || counts_host != nullptr, "dgs_sample_blocks_multi: wait needs counts_host");
  WsInfo info;
  if (ws_lookup(ws, g->itype, num_batches, num_seeds, num_layers, fan_out, g->num_nodes, ws_bytes, &info))
    return 1;
  DGS_REQUIRE(info.path == kPathMulti, "dgs_sample_blocks_multi: workspace was initialised with "
              "dgs_sample_blocks_ws_init for the single-batch path");
  return sample_multi_impl(g, num_batches, seeds, seeds_stride_bytes, num_seeds, num_layers, fan_out,
                           replace, rng_seeds, out_frontier, out_row, out_col, out_stride_bytes, cap_edges,
                           cap_frontier, counts_dev, ws, ws_bytes, counts_host, (cudaStream_t)stream,
                           wait != 0, info);
}
