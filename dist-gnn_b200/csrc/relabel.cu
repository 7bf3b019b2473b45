// relabel.cu - order-preserving unique + id compaction of a sampled hop.
//
// Replaces TensorRelabelCUDA / Unique / Relabel / RelabelHashmap
// (src/sampling/cuda/tensor_relabel.cu:11-205): `unique` is the first-occurrence-order unique of
// the concatenated mapping tensors and every element of the tensors to relabel is replaced by its
// position in `unique` (-1 when absent).  The result is a deterministic function of the input and
// must be bit-exact with the reference.
//
// Differences in mechanism (not in result):
//   * no torch::cat: the mapping / relabel inputs are up to 4 parts addressed as one virtual array,
//     with optional device-side lengths so a whole multi-hop batch is enqueued without host syncs
//     (reference: 2 cats, 3 torch::full of the table size, 4 thrust passes, a cub scan and a
//     blocking D2H read of the unique count per hop, :82-159).
//   * one 16-byte slot {key, first occurrence, local rank} and a mixed hash instead of three
//     arrays with an identity hash (:63-69, clusters on consecutive ids).
//   * the table is persistent: the owners of the touched slots restore them to "empty" in the
//     last kernel, so there is no per-call fill of capacity-sized tensors.
// Kernels: insert -> rank (CTA scan + last-CTA tile prefixes) -> emit (unique + relabel) -> reset.
#include "dgs_common.cuh"

namespace dgsb {

constexpr int kRlThreads = 256;
constexpr int kRlItems = 8;                       // items per thread in the rank kernel
constexpr int kRlTile = kRlThreads * kRlItems;    // 2048
constexpr int kMaxParts = 4;

struct __align__(16) RlSlot {
  long long key;       // -1 empty
  unsigned int first;  // index of the first occurrence in the virtual mapping array
  unsigned int lrank;  // rank of that first occurrence inside its tile
};

template <typename IdT>
struct Parts {
  int n;
  const IdT *ptr[kMaxParts];
  int64_t count[kMaxParts];            // upper bound / exact
  const int64_t *count_dev[kMaxParts]; // optional live count
  IdT *out[kMaxParts];                 // relabel outputs (relabel side only)
  int alias[kMaxParts];                // relabel part p is mapping part alias[p] (or -1)
};

template <typename IdT>
struct PartView {
  int64_t start[kMaxParts + 1];
  __device__ __forceinline__ void init(const Parts<IdT> &p) {
    int64_t s = 0;
#pragma unroll
    for (int q = 0; q < kMaxParts; ++q) {
      start[q] = s;
      if (q < p.n) {
        int64_t c = p.count[q];
        if (p.count_dev[q]) c = min(c, *p.count_dev[q]);
        s += c;
      }
    }
    start[kMaxParts] = s;
  }
  __device__ __forceinline__ int64_t total() const { return start[kMaxParts]; }
  __device__ __forceinline__ int part_of(int64_t i) const {
    int q = 0;
#pragma unroll
    for (int t = 1; t < kMaxParts; ++t)
      if (i >= start[t]) q = t;
    return q;
  }
};

struct RelabelWs {
  unsigned int *done;      // [1]
  unsigned int *pos;       // [n] slot of item i
  unsigned int *lrank;     // [n] exclusive rank inside the tile (valid for first occurrences)
  long long *tile_prefix;  // [tiles + 1]
};

static inline int64_t rl_align(int64_t x) { return (x + 255) / 256 * 256; }
static int64_t rl_layout(int64_t n, char *base, RelabelWs *ws) {
  int64_t tiles = (n + kRlTile - 1) / kRlTile;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char *p = base ? base + off : nullptr;
    off += rl_align(bytes);
    return p;
  };
  char *done = take(256);
  char *pos = take(n * 4);
  char *lr = take(n * 4);
  char *tp = take((tiles + 1) * 8);
  if (ws) {
    ws->done = (unsigned int *)done;
    ws->pos = (unsigned int *)pos;
    ws->lrank = (unsigned int *)lr;
    ws->tile_prefix = (long long *)tp;
  }
  return off;
}

__device__ __forceinline__ uint64_t rl_hash(long long key, uint64_t cap_mask) {
  return mix64((uint64_t)key) & cap_mask;
}

// insert every mapping item, remember its slot, keep the minimum index per key
template <typename IdT>
__global__ void __launch_bounds__(kRlThreads)
rl_insert_kernel(Parts<IdT> map, RlSlot *table, uint64_t cap_mask, RelabelWs ws) {
  __shared__ PartView<IdT> pv;
  if (threadIdx.x == 0) pv.init(map);
  __syncthreads();
  const int64_t n = pv.total();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int q = pv.part_of(i);
    const long long key = (long long)map.ptr[q][i - pv.start[q]];
    uint64_t pos = rl_hash(key, cap_mask);
    while (true) {
      unsigned long long prev = atomicCAS((unsigned long long *)&table[pos].key,
                                          (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (prev == (unsigned long long)kEmptyKey || prev == (unsigned long long)key) break;
      pos = (pos + 1) & cap_mask;
    }
    atomicMin(&table[pos].first, (unsigned int)i);
    ws.pos[i] = (unsigned int)pos;
  }
}

// flag first occurrences, rank them inside 2048-item tiles, last CTA scans the tile totals
template <typename IdT>
__global__ void __launch_bounds__(kRlThreads)
rl_rank_kernel(Parts<IdT> map, RlSlot *table, RelabelWs ws, int64_t *num_unique) {
  __shared__ PartView<IdT> pv;
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ bool s_last;
  if (threadIdx.x == 0) pv.init(map);
  __syncthreads();
  const int64_t n = pv.total();
  const int64_t tiles = (n + kRlTile - 1) / kRlTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t base = tile * kRlTile + (int64_t)threadIdx.x * kRlItems;
    unsigned int slot[kRlItems];
    bool flag[kRlItems];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < kRlItems; ++u) {
      const int64_t i = base + u;
      flag[u] = false;
      if (i < n) {
        slot[u] = ws.pos[i];
        flag[u] = (table[slot[u]].first == (unsigned int)i);
      }
      mine += flag[u] ? 1 : 0;
    }
    long long excl = block_exclusive_scan<long long>((long long)mine, s_scan, &s_total);
    unsigned int r = (unsigned int)excl;
#pragma unroll
    for (int u = 0; u < kRlItems; ++u) {
      if (flag[u]) {
        table[slot[u]].lrank = r;
        ws.lrank[base + u] = r;
        ++r;
      }
    }
    if (threadIdx.x == 0) ws.tile_prefix[tile] = s_total;
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ws.done, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    long long carry = 0;
    volatile long long *tp = ws.tile_prefix;
    for (int64_t b = 0; b < tiles; b += kRlThreads) {
      const int64_t t = b + threadIdx.x;
      long long val = t < tiles ? tp[t] : 0;
      long long excl = block_exclusive_scan<long long>(val, s_scan, &s_total);
      if (t < tiles) tp[t] = carry + excl;
      carry += s_total;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      tp[tiles] = carry;
      *num_unique = carry;
      *ws.done = 0;
    }
  }
}

// write `unique` (first occurrences, in order) and the relabelled tensors
template <typename IdT>
__global__ void __launch_bounds__(kRlThreads)
rl_emit_kernel(Parts<IdT> map, Parts<IdT> rel, const RlSlot *__restrict__ table, uint64_t cap_mask,
               RelabelWs ws, IdT *__restrict__ unique) {
  __shared__ PartView<IdT> pv, rv;
  if (threadIdx.x == 0) {
    pv.init(map);
    rv.init(rel);
  }
  __syncthreads();
  const int64_t n = pv.total(), m = rv.total();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < n; i += stride) {
    const unsigned int p = ws.pos[i];
    const RlSlot s = table[p];
    if (s.first == (unsigned int)i)
      unique[ws.tile_prefix[i / kRlTile] + (long long)s.lrank] = (IdT)s.key;
  }
  for (int64_t j = tid; j < m; j += stride) {
    const int q = rv.part_of(j);
    const int64_t lj = j - rv.start[q];
    long long id = -1;
    if (rel.alias[q] >= 0) {
      const RlSlot s = table[ws.pos[pv.start[rel.alias[q]] + lj]];
      id = ws.tile_prefix[s.first / kRlTile] + (long long)s.lrank;
    } else {
      const long long key = (long long)rel.ptr[q][lj];
      uint64_t pos = rl_hash(key, cap_mask);
      while (true) {
        const int4 raw = ld_v4(&table[pos]);
        const long long k2 = ((long long)(uint32_t)raw.y << 32) | (uint32_t)raw.x;
        if (k2 == key) {
          id = ws.tile_prefix[(uint32_t)raw.z / kRlTile] + (long long)(uint32_t)raw.w;
          break;
        }
        if (k2 == kEmptyKey) break;
        pos = (pos + 1) & cap_mask;
      }
    }
    rel.out[q][lj] = (IdT)id;
  }
}

// owners restore their slots to empty so the table can be reused without a memset
template <typename IdT>
__global__ void __launch_bounds__(kRlThreads)
rl_reset_kernel(Parts<IdT> map, RlSlot *table, RelabelWs ws) {
  __shared__ PartView<IdT> pv;
  if (threadIdx.x == 0) pv.init(map);
  __syncthreads();
  const int64_t n = pv.total();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned int p = ws.pos[i];
    if (table[p].first == (unsigned int)i) {
      int4 e = make_int4(-1, -1, -1, -1);
      *reinterpret_cast<int4 *>(&table[p]) = e;
    }
  }
}

}  // namespace dgsb

using namespace dgsb;

extern "C" int64_t dgs_relabel_table_capacity(int64_t n) {
  // power of two >= 2 n (load factor <= 0.5); the reference uses 2^(floor(log2 n)+1) (:77-80)
  int64_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}
extern "C" int64_t dgs_relabel_table_bytes(int64_t n) {
  return dgs_relabel_table_capacity(n) * (int64_t)sizeof(RlSlot);
}
extern "C" int64_t dgs_relabel_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return rl_layout(n, nullptr, nullptr);
}

extern "C" int dgs_relabel(int itype, int n_map, const void *const *map_ptrs,
                           const int64_t *map_counts, const int64_t *const *map_counts_dev,
                           int n_rel, const void *const *rel_ptrs, const int64_t *rel_counts,
                           const int64_t *const *rel_counts_dev, void *const *rel_out,
                           void *unique_out, int64_t *num_unique_dev, void *table,
                           int64_t capacity, void *ws, void *stream) {
  DGS_REQUIRE(n_map >= 1 && n_map <= kMaxParts && n_rel >= 0 && n_rel <= kMaxParts,
              "dgs_relabel: 1..%d mapping parts and 0..%d relabel parts supported (got %d / %d)",
              kMaxParts, kMaxParts, n_map, n_rel);
  DGS_REQUIRE(map_ptrs && map_counts && num_unique_dev && table && ws,
              "dgs_relabel: null argument");
  DGS_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0,
              "dgs_relabel: capacity must be a power of two");
  int64_t n = 0, m = 0;
  for (int q = 0; q < n_map; ++q) {
    DGS_REQUIRE(map_counts[q] >= 0, "dgs_relabel: negative count");
    DGS_REQUIRE(map_counts[q] == 0 || map_ptrs[q], "dgs_relabel: null mapping part %d", q);
    n += map_counts[q];
  }
  for (int q = 0; q < n_rel; ++q) {
    DGS_REQUIRE(rel_counts[q] >= 0, "dgs_relabel: negative count");
    DGS_REQUIRE(rel_counts[q] == 0 || (rel_ptrs[q] && rel_out && rel_out[q]),
                "dgs_relabel: null relabel part %d", q);
    m += rel_counts[q];
  }
  DGS_REQUIRE(2 * n <= capacity, "dgs_relabel: table capacity %lld < 2 * %lld items",
              (long long)capacity, (long long)n);
  DGS_REQUIRE(n < (1ll << 32) - 1, "dgs_relabel: more than 2^32 mapping items");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    DGS_CUDA_OK(cudaMemsetAsync(num_unique_dev, 0, sizeof(int64_t), st));
    for (int q = 0; q < n_rel; ++q)
      if (rel_counts[q] > 0)
        DGS_CUDA_OK(cudaMemsetAsync(rel_out[q], 0xFF,
                                    (size_t)rel_counts[q] * (itype == DGS_I64 ? 8 : 4), st));
    return 0;
  }
  DGS_REQUIRE(unique_out != nullptr, "dgs_relabel: null unique output");
  RelabelWs w;
  rl_layout(n, (char *)ws, &w);
  DGS_ITYPE_SWITCH(itype, IdT, {
    Parts<IdT> mp, rp;
    memset(&mp, 0, sizeof(mp));
    memset(&rp, 0, sizeof(rp));
    mp.n = n_map;
    for (int q = 0; q < n_map; ++q) {
      mp.ptr[q] = (const IdT *)map_ptrs[q];
      mp.count[q] = map_counts[q];
      mp.count_dev[q] = map_counts_dev ? map_counts_dev[q] : nullptr;
    }
    rp.n = n_rel;
    for (int q = 0; q < n_rel; ++q) {
      rp.ptr[q] = (const IdT *)rel_ptrs[q];
      rp.count[q] = rel_counts[q];
      rp.count_dev[q] = rel_counts_dev ? rel_counts_dev[q] : nullptr;
      rp.out[q] = (IdT *)rel_out[q];
      rp.alias[q] = -1;
      for (int a = 0; a < n_map; ++a)
        if (rel_ptrs[q] == map_ptrs[a] && rel_counts[q] == map_counts[a] &&
            rp.count_dev[q] == mp.count_dev[a])
          rp.alias[q] = a;
    }
    const uint64_t mask = (uint64_t)capacity - 1;
    int grid_n = grid_for(n, kRlThreads, 8);
    rl_insert_kernel<IdT><<<grid_n, kRlThreads, 0, st>>>(mp, (RlSlot *)table, mask, w);
    DGS_LAUNCH_CHECK();
    int grid_t = grid_for(n, kRlTile, 4);
    rl_rank_kernel<IdT><<<grid_t, kRlThreads, 0, st>>>(mp, (RlSlot *)table, w, num_unique_dev);
    DGS_LAUNCH_CHECK();
    int grid_e = grid_for(n > m ? n : m, kRlThreads, 8);
    rl_emit_kernel<IdT><<<grid_e, kRlThreads, 0, st>>>(mp, rp, (const RlSlot *)table, mask, w,
                                                       (IdT *)unique_out);
    DGS_LAUNCH_CHECK();
    rl_reset_kernel<IdT><<<grid_n, kRlThreads, 0, st>>>(mp, (RlSlot *)table, w);
    DGS_LAUNCH_CHECK();
  });
  return 0;
}
