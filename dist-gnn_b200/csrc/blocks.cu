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
struct Tab {
  char *base;
  int direct;
  __device__ __forceinline__ unsigned int *first(uint64_t pos) const {
    return reinterpret_cast<unsigned int *>(base + (direct ? pos * 8 : pos * 16 + 8));
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
};

struct BlocksWs {
  unsigned int *done;
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
constexpr int kEmBatch = 4;    // padded slots per thread and pass in the emit phase (8 spills)
constexpr int kRkItems = 8;    // padded slots per thread and pass in the rank phase
constexpr int kFloydRegs = 16; // fan-outs up to this run Floyd's sampling in registers
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
template <typename IdT, typename ET, int MODE>
__device__ __forceinline__ void pick_tile_phase(const GraphSrc &g, const IdT *__restrict__ seeds,
                                                int64_t S_ub, int64_t S, int k, uint64_t rng_key,
                                                IdT *__restrict__ pad_col, const HopState &cur,
                                                uint64_t cap_mask,
                                                unsigned long long *fine = nullptr) {
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
  const int ts = pick_tile_seeds(S);
  const int64_t tiles = (S + ts - 1) / ts;
  const bool with_replace = (MODE == kUniformReplace || MODE == kBiasReplace);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
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
      if (MODE == kUniform && k <= kFloydRegs && deg > k) {
        // Floyd's subset sampling entirely in registers (fully unrolled, no shared-memory round
        // trips): draw t picks r in [0, deg-k+t], or deg-k+t itself when r was already picked
        unsigned int P[kFloydRegs];
#pragma unroll
        for (int q = 0; q < kFloydRegs / 4; ++q) {
          if (4 * q < k) {
            const uint4 r4 = Philox::gen(rng_key, (uint64_t)i, (uint64_t)q);
            P[4 * q] = r4.x; P[4 * q + 1] = r4.y; P[4 * q + 2] = r4.z; P[4 * q + 3] = r4.w;
          }
        }
#pragma unroll
        for (int t = 0; t < kFloydRegs; ++t) {
          if (t < k) {
            const unsigned int J = (unsigned int)(deg - k + t);
            const unsigned int r = rand_below(P[t], J + 1);
            bool dup = false;
#pragma unroll
            for (int q = 0; q < t; ++q) dup |= (P[q] == r);
            P[t] = dup ? J : r;
          }
        }
        unsigned int *dst = s_pick + (size_t)tid * k;
#pragma unroll
        for (int t = 0; t < kFloydRegs; ++t)
          if (t < k) dst[t] = P[t];
      }
    }
    if (tile == blockIdx.x) fstamp(1);
    if (MODE == kUniform && k > kFloydRegs) {
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
    if (MODE != kUniform || k > kFloydRegs) __syncthreads();
    if (MODE == kUniform && k > kFloydRegs) {
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
      if (hubs_shared) {
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
    if (tile == blockIdx.x) fstamp(2);
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
        item[u] = (unsigned int)(S_ub + i0 * k + el);
        if (el < slots) {
          const int si = el / k;
          const int j = el - si * k;
          if (j < s_cnt[si]) {
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
          cur.pos_col[e] = pos[u];
        }
      }
    }
    if (tid < ns) {
      while (seed_prev != (unsigned long long)kEmptyKey && seed_prev != (unsigned long long)seed_nid) {
        seed_pos = (seed_pos + 1) & cap_mask;
        seed_prev = atomicCAS(cur.table.key(seed_pos), (unsigned long long)kEmptyKey,
                              (unsigned long long)seed_nid);
      }
      atomicMin(cur.table.first(seed_pos), (unsigned int)(i0 + tid));
      cur.pos_seed[i0 + tid] = (unsigned int)seed_pos;
    }
    if (tile == blockIdx.x) fstamp(3);
  }
  fstamp(4);
}

// Rank phase: CTA per 64 seeds - flags first occurrences among the seeds (A) and the sampled
// neighbours (B), counts the edges (C), block scans, per-tile totals to prefA / prefB / prefC.
__device__ __forceinline__ void rank_tiles_phase(int64_t S_ub, int64_t S, int k, const HopState &cur,
                                                 const BlocksWs &ws, bool unique_seeds) {
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ int s_cnt[kBkTile];
  const int64_t tiles = (S + kBkTile - 1) / kBkTile;
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t i0 = tile * kBkTile;
    const int ns = (int)min((int64_t)kBkTile, S - i0);
    const int items = ns * k;
    const int64_t e0 = i0 * k;
    __syncthreads();
    // first wave of loads: seed slot + count, and the first pass of neighbour slots
    long long fa = 0, c = 0;
    unsigned int slot_a = 0;
    if (tid < ns) {
      slot_a = ldcg(cur.pos_seed + i0 + tid);
      c = ldcg(cur.cnt + i0 + tid);
      s_cnt[tid] = (int)c;
    }
    __syncthreads();
    // seeds that are a previous frontier are distinct: each is its own first occurrence
    if (tid < ns)
      fa = (unique_seeds || ldcg(cur.table.first(slot_a)) == (unsigned int)(i0 + tid)) ? 1 : 0;
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
          if (valid[u]) slot[u] = ldcg(cur.pos_col + e0 + el);
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
        if (fa) *cur.table.lrank(slot_a) = (unsigned int)(rac >> 32);
        if (tid < ns) ws.loff[i0 + tid] = (int)(rac & 0xffffffffll);
      }
      int mine = 0;
      bool fb[kRkItems];
#pragma unroll
      for (int u = 0; u < kRkItems; ++u) {
        fb[u] = valid[u] && first[u] == (unsigned int)(S_ub + e0 + el0 + u);
        mine += fb[u] ? 1 : 0;
      }
      long long r = carry + block_exclusive_scan<long long>((long long)mine, s_scan, &s_total);
#pragma unroll
      for (int u = 0; u < kRkItems; ++u)
        if (fb[u]) *cur.table.lrank(slot[u]) = (unsigned int)(r++);
      carry += s_total;
    }
    if (tid == 0) {
      ws.prefA[tile] = totA;
      ws.prefB[tile] = carry;
      ws.prefC[tile] = totC;
    }
  }
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
  pick_tile_phase<IdT, ET, MODE>(g, seeds, S_ub, S, k, rng_key, pad_col, cur, cap_mask);
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
                                   (IdT *)ws.pad_col, cur, a.cap_mask,
                                   a.trace ? a.trace + 256 + 8 * l : nullptr);
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

__global__ void blocks_ws_init_kernel(int4 *p, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = make_int4(-1, -1, -1, -1);
}

int build_graph_src(const dgs_graph_t *g, GraphSrc *out);  // sampling.cu

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

extern "C" int64_t dgs_sample_blocks_ws_bytes(int itype, int64_t num_seeds, int num_layers,
                                              const int64_t *fan_out, int64_t num_nodes) {
  BlocksPlan p;
  if (!fan_out || blocks_plan(itype, num_seeds, num_layers, fan_out, num_nodes, &p, nullptr, nullptr))
    return -1;
  return p.bytes;
}

extern "C" int dgs_sample_blocks_ws_init(void *ws, int64_t ws_bytes, int itype, int64_t num_seeds,
                                         int num_layers, const int64_t *fan_out, int64_t num_nodes,
                                         void *stream) {
  DGS_REQUIRE(ws && fan_out, "dgs_sample_blocks_ws_init: null argument");
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
  return 0;
}

// Wait until the hop sizes of the batch enqueued with counts_host have arrived there.
static int counts_wait(int64_t *counts_host, const int64_t *counts_dev, int num_layers, bool by_kernel,
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
  DGS_CUDA_OK(cudaMemcpyAsync(counts_host, counts_dev, sizeof(int64_t) * 2 * num_layers,
                              cudaMemcpyDeviceToHost, st));
  DGS_CUDA_OK(cudaStreamSynchronize(st));
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
  BlocksPlan p;
  BlocksWs w;
  if (blocks_plan(g->itype, num_seeds, num_layers, fan_out, g->num_nodes, &p, (char *)ws, &w)) return 1;
  DGS_REQUIRE(ws_bytes >= p.bytes, "dgs_sample_blocks: workspace %lld < %lld bytes",
              (long long)ws_bytes, (long long)p.bytes);
  GraphSrc src;
  if (build_graph_src(g, &src)) return 1;
  // Can the kernel write the hop sizes into counts_host itself?  (pinned + mapped host memory;
  // the answer is remembered for the last pointer asked about)
  long long *host_dev = nullptr;
  if (counts_host) {
    static const bool no_direct = getenv("DGS_BLOCKS_COUNTS_MEMCPY") != nullptr;
    static int64_t *last_host = nullptr;
    static long long *last_dev = nullptr;
    if (!no_direct) {
      if (counts_host != last_host) {
        cudaPointerAttributes at;
        void *dp = nullptr;
        if (cudaPointerGetAttributes(&at, counts_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
            cudaHostGetDevicePointer(&dp, counts_host, 0) == cudaSuccess)
          last_dev = (long long *)dp;
        else
          last_dev = nullptr;
        cudaGetLastError();
        last_host = counts_host;
      }
      host_dev = last_dev;
    }
    if (host_dev || !wait) *(volatile int64_t *)counts_host = -1;   // sentinel: hop sizes are >= 0
  }
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
    return counts_wait(counts_host, counts_dev, num_layers, by_kernel, st);
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
  return counts_wait(counts_host, counts_dev, num_layers, true, (cudaStream_t)stream);
}
