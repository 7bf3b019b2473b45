// sampling.cu - fan-out neighbour sampling over CSR (uniform / edge-weight biased, with or
// without replacement), plain and p2p-cached.
//
// Replaces the eight kernels + thrust/cub glue of
//   src/sampling/cuda/rowwise_sampling.cu:16-189        (uniform, plain)
//   src/sampling/cuda/rowwise_sampling_bias.cu:16-288   (biased, plain)
//   src/sampling/cuda/rowwise_sampling_p2p.cu:19-270    (uniform, p2p cache)
//   src/sampling/cuda/rowwise_sampling_bias_p2p.cu:18-385 (biased, p2p cache)
// by two kernels per hop that work for all variants:
//   plan_kernel : thread per seed - location-table probe, indptr pair from the owner (local HBM,
//                 NVLink peer shard or pinned host), output count, CTA-level exclusive scan; the
//                 last CTA to finish turns the per-tile totals into tile prefixes and publishes
//                 nnz on the device.  (reference: thrust::for_each + 2 cub launches + a blocking
//                 D2H read of nnz, rowwise_sampling_p2p.cu:181-228.)
//   pick_kernel : warp per seed - selection in registers / shared memory, then the picked
//                 neighbour ids are gathered and written as seed-major COO.
// Selection algorithms (all O(k) or O(deg) per seed with no global scratch):
//   uniform w/o replacement : Floyd's subset sampling held one element per lane (k <= 32) or in
//                             shared memory - O(k) work independent of the degree (the reference
//                             runs a deg-long reservoir with global atomicMax, :80-92)
//   uniform with replacement: one stateless Philox draw per output edge
//   biased w/o replacement  : A-Res keys log2(u)/w (same order as the reference's u^(1/w),
//                             rowwise_sampling_bias.cu:112-113) + warp-level "replace the minimum"
//                             reservoir of the k largest keys (any k, not only <= 32)
//   biased with replacement : two streaming passes over the weights (warp inclusive scan with
//                             carry, then inverse-CDF by ballot) - no global CDF temp
//                             (reference: rowwise_sampling_bias.cu:188-220)
// deg <= k (and num_picks < 0) is the copy path: neighbours in CSR order, bit-exact with the
// reference (rowwise_sampling.cu:71-77).  RNG is Philox4x32-10 keyed by the launch seed with
// counter (seed index, draw index): results do not depend on grid or block shape.
#include "dgs_common.cuh"
#include "p2p_server.h"

namespace dgsb {

constexpr int kPlanThreads = 256;  // = seeds per plan tile
constexpr int kPickWarps = 8;

struct GraphSrc {
  const void *indptr;
  const void *indices;
  const float *probs;
  PtrTable sh_indptr, sh_indices, sh_probs;
  const LocSlot *loc;
  uint64_t cap_mask;
};

struct SampleWs {
  long long *begin;       // [M] first edge of the seed's row inside its source array
  int *deg;               // [M]
  int *dev;               // [M] owner device, -1 = un-cached source
  long long *loff;        // [M] exclusive offset inside the seed's plan tile
  long long *tile_prefix; // [tiles + 1] exclusive tile prefixes (totals before the last CTA ran)
  unsigned int *done;     // [1] CTA completion counter (self-resetting)
};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static int64_t ws_layout(int64_t M, char *base, SampleWs *ws) {
  int64_t tiles = (M + kPlanThreads - 1) / kPlanThreads;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char *p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  char *done = take(256);
  char *begin = take(M * 8);
  char *deg = take(M * 4);
  char *dev = take(M * 4);
  char *loff = take(M * 8);
  char *tp = take((tiles + 1) * 8);
  if (ws) {
    ws->done = (unsigned int *)done;
    ws->begin = (long long *)begin;
    ws->deg = (int *)deg;
    ws->dev = (int *)dev;
    ws->loff = (long long *)loff;
    ws->tile_prefix = (long long *)tp;
  }
  return off;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ------------------------------------------------------------------------------------------
template <typename IdT, typename ET>
__global__ void __launch_bounds__(kPlanThreads)
plan_kernel(GraphSrc g, const IdT *__restrict__ seeds, int64_t num_seeds,
            const int64_t *__restrict__ num_seeds_dev, int64_t k, int replace, SampleWs ws,
            int64_t *__restrict__ out_nnz) {
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ bool s_last;
  const int64_t S = num_seeds_dev ? min(*num_seeds_dev, num_seeds) : num_seeds;
  const int64_t tiles = (S + kPlanThreads - 1) / kPlanThreads;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t i = tile * kPlanThreads + threadIdx.x;
    long long cnt = 0;
    if (i < S) {
      const long long nid = (long long)seeds[i];
      long long begin, end;
      int dev = -1;
      long long v = -1;
      if (g.loc != nullptr) v = loc_lookup(g.loc, g.cap_mask, nid);
      if (v >= 0) {
        dev = (int)((v >> kDevShift) & 0xff);
        const long long idx = v & kIdxMask;
        const ET *ip = reinterpret_cast<const ET *>(g.sh_indptr.p[dev]);
        begin = (long long)ip[idx];
        end = (long long)ip[idx + 1];
      } else {
        const ET *ip = reinterpret_cast<const ET *>(g.indptr);
        begin = (long long)ip[nid];
        end = (long long)ip[nid + 1];
      }
      const long long deg = end - begin;
      if (replace)
        cnt = (deg == 0 || k < 0) ? (k < 0 ? deg : 0) : k;
      else
        cnt = (k < 0 || deg < k) ? deg : k;
      ws.begin[i] = begin;
      ws.deg[i] = (int)deg;
      ws.dev[i] = dev;
    }
    long long excl = block_exclusive_scan<long long>(cnt, s_scan, &s_total);
    if (i < S) ws.loff[i] = excl;
    if (threadIdx.x == 0) ws.tile_prefix[tile] = s_total;
    __syncthreads();
  }
  // last CTA: exclusive scan of the tile totals, publish nnz.
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
    for (int64_t base = 0; base < tiles; base += kPlanThreads) {
      const int64_t t = base + threadIdx.x;
      long long val = t < tiles ? tp[t] : 0;
      long long excl = block_exclusive_scan<long long>(val, s_scan, &s_total);
      if (t < tiles) tp[t] = carry + excl;
      carry += s_total;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      tp[tiles] = carry;
      *out_nnz = carry;
      *ws.done = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
enum PickMode { kUniform = 0, kUniformReplace = 1, kBias = 2, kBiasReplace = 3 };

__device__ __forceinline__ uint32_t philox_u32(uint64_t key, uint64_t item, uint32_t draw) {
  uint4 r = Philox::gen(key, item, (uint64_t)(draw >> 2));
  const uint32_t c = draw & 3u;
  return c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
}

// Write the k picked neighbours.  w_idx may alias the first 4 k bytes of ocol (global scratch
// mode): chunks are processed from the top and every lane reads its position before any lane
// writes, so an 8-byte ocol[j] only overwrites positions >= j that were already consumed.
template <typename IdT>
__device__ __forceinline__ void emit_picks(const IdT *__restrict__ row, const int *w_idx, int k,
                                           int lane, IdT seed, IdT *orow, IdT *ocol) {
  for (int j0 = ((k - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
    const int j = j0 + lane;
    int p = 0;
    if (j < k) p = w_idx[j];
    __syncwarp();
    if (j < k) {
      ocol[j] = row[p];
      orow[j] = seed;
    }
    __syncwarp();
  }
}

template <typename IdT, typename ET, int MODE>
__global__ void __launch_bounds__(kPickWarps * 32)
pick_kernel(GraphSrc g, const IdT *__restrict__ seeds, int64_t num_seeds,
            const int64_t *__restrict__ num_seeds_dev, int64_t k64, uint64_t rng_key, SampleWs ws,
            IdT *out_row, IdT *out_col, int64_t capacity, int gmem_scratch) {
  extern __shared__ __align__(16) unsigned char pick_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int k = (int)k64;
  const int64_t S = num_seeds_dev ? min(*num_seeds_dev, num_seeds) : num_seeds;
  const int64_t warps_total = (int64_t)gridDim.x * kPickWarps;
  // per-warp scratch: k ints (uniform) or k floats + k ints (biased)
  int *w_idx = nullptr;
  float *w_key = nullptr;
  // Scratch for the selection state: shared memory, or - when num_picks is too large for it - the
  // seed's own k output slots (k * sizeof(IdT) >= k * 4 bytes each in out_row / out_col), which
  // are rewritten with the final (seed, neighbour) pairs afterwards.
  if (k > 0 && !gmem_scratch) {
    if (MODE == kUniform) {
      w_idx = reinterpret_cast<int *>(pick_smem) + (size_t)warp * k;
    } else if (MODE == kBias) {
      w_key = reinterpret_cast<float *>(pick_smem) + (size_t)warp * 2 * k;
      w_idx = reinterpret_cast<int *>(w_key + k);
    }
  }

  for (int64_t i = (int64_t)blockIdx.x * kPickWarps + warp; i < S; i += warps_total) {
    const int deg = ws.deg[i];
    if (deg == 0) continue;
    const int dev = ws.dev[i];
    const long long begin = ws.begin[i];
    const long long off = ws.tile_prefix[i / kPlanThreads] + ws.loff[i];
    const IdT seed = seeds[i];
    const IdT *row =
        reinterpret_cast<const IdT *>(dev < 0 ? g.indices : g.sh_indices.p[dev]) + begin;
    const bool with_replace = (MODE == kUniformReplace || MODE == kBiasReplace);
    const bool copy_path = (k < 0) || (!with_replace && deg <= k);
    const long long cnt = copy_path ? (long long)deg : (long long)k;
    if (off + cnt > capacity) continue;  // counting-only call (caller re-runs with a larger buffer)
    IdT *orow = out_row + off;
    IdT *ocol = out_col + off;
    if (gmem_scratch && !copy_path) {
      w_key = reinterpret_cast<float *>(orow);
      w_idx = reinterpret_cast<int *>(ocol);
    }

    if (copy_path) {
      // CSR-order copy, 4 independent loads in flight per lane
      int j = lane;
      for (; j + 96 < deg; j += 128) {
        IdT a = row[j], b = row[j + 32], c = row[j + 64], d = row[j + 96];
        ocol[j] = a; ocol[j + 32] = b; ocol[j + 64] = c; ocol[j + 96] = d;
        orow[j] = seed; orow[j + 32] = seed; orow[j + 64] = seed; orow[j + 96] = seed;
      }
      for (; j < deg; j += 32) {
        ocol[j] = row[j];
        orow[j] = seed;
      }
      continue;
    }

    if (MODE == kUniformReplace) {
      for (int j = lane; j < k; j += 32) {
        uint32_t p = rand_below(philox_u32(rng_key, (uint64_t)i, (uint32_t)j), (uint32_t)deg);
        ocol[j] = row[p];
        orow[j] = seed;
      }
    } else if (MODE == kUniform) {
      // Floyd: for t = 0..k-1, J = deg-k+t: r = U[0, J]; pick (r already chosen ? J : r)
      if (k <= 32) {
        uint32_t r_mine = 0;
        if (lane < k)
          r_mine = rand_below(philox_u32(rng_key, (uint64_t)i, (uint32_t)lane),
                              (uint32_t)(deg - k + lane + 1));
        uint32_t mine = 0xffffffffu;  // lane t holds the t-th pick
        for (int t = 0; t < k; ++t) {
          const uint32_t r = __shfl_sync(0xffffffffu, r_mine, t);
          const bool dup = __any_sync(0xffffffffu, lane < t && mine == r);
          if (lane == t) mine = dup ? (uint32_t)(deg - k + t) : r;
        }
        if (lane < k) {
          ocol[lane] = row[mine];
          orow[lane] = seed;
        }
      } else {
        for (int t0 = 0; t0 < k; t0 += 32) {
          const int t_mine = t0 + lane;
          uint32_t r_mine = 0;
          if (t_mine < k)
            r_mine = rand_below(philox_u32(rng_key, (uint64_t)i, (uint32_t)t_mine),
                                (uint32_t)(deg - k + t_mine + 1));
          const int lim = min(32, k - t0);
          for (int tt = 0; tt < lim; ++tt) {
            const int t = t0 + tt;
            const uint32_t r = __shfl_sync(0xffffffffu, r_mine, tt);
            bool found = false;
            for (int c = lane; c < t; c += 32) found |= ((uint32_t)w_idx[c] == r);
            const bool dup = __any_sync(0xffffffffu, found);
            if (lane == 0) w_idx[t] = dup ? (deg - k + t) : (int)r;
            __syncwarp();
          }
        }
        emit_picks(row, w_idx, k, lane, seed, orow, ocol);
      }
    } else if (MODE == kBias) {
      const float *wrow = (dev < 0 ? g.probs : reinterpret_cast<const float *>(g.sh_probs.p[dev])) + begin;
      // fill the reservoir with the first k items
      for (int t = lane; t < k; t += 32) {
        const float w = wrow[t];
        const float u = u32_to_unit(philox_u32(rng_key, (uint64_t)i, (uint32_t)t));
        w_key[t] = w > 0.f ? __log2f(u) / w : -INFINITY;
        w_idx[t] = t;
      }
      __syncwarp();
      // (min key, its slot) over the reservoir
      float lmin = INFINITY;
      int lslot = -1;
      for (int c = lane; c < k; c += 32) {
        const float v = w_key[c];
        if (v < lmin || lslot < 0) { lmin = v; lslot = c; }
      }
      float wmin = lmin;
      int wslot = lslot;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wmin, o);
        const int os = __shfl_xor_sync(0xffffffffu, wslot, o);
        if (os >= 0 && (wslot < 0 || ov < wmin || (ov == wmin && os < wslot))) { wmin = ov; wslot = os; }
      }
      for (int t0 = k; t0 < deg; t0 += 32) {
        const int t = t0 + lane;
        float key = -INFINITY;
        if (t < deg) {
          const float w = wrow[t];
          const float u = u32_to_unit(philox_u32(rng_key, (uint64_t)i, (uint32_t)t));
          key = w > 0.f ? __log2f(u) / w : -INFINITY;
        }
        unsigned mask = __ballot_sync(0xffffffffu, key > wmin);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const float ck = __shfl_sync(0xffffffffu, key, src);
          const int ci = t0 + src;
          if (ck > wmin) {  // warp-uniform
            if (lane == 0) { w_key[wslot] = ck; w_idx[wslot] = ci; }
            __syncwarp();
            lmin = INFINITY;
            lslot = -1;
            for (int c = lane; c < k; c += 32) {
              const float v = w_key[c];
              if (v < lmin || lslot < 0) { lmin = v; lslot = c; }
            }
            wmin = lmin;
            wslot = lslot;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, wmin, o);
              const int os = __shfl_xor_sync(0xffffffffu, wslot, o);
              if (os >= 0 && (wslot < 0 || ov < wmin || (ov == wmin && os < wslot))) { wmin = ov; wslot = os; }
            }
          }
        }
      }
      __syncwarp();
      emit_picks(row, w_idx, k, lane, seed, orow, ocol);
    } else {  // kBiasReplace
      const float *wrow = (dev < 0 ? g.probs : reinterpret_cast<const float *>(g.sh_probs.p[dev])) + begin;
      // pass 1: total weight, with exactly the arithmetic of pass 2
      float total = 0.f;
      for (int c = 0; c < deg; c += 32) {
        float w = (c + lane < deg) ? fmaxf(wrow[c + lane], 0.f) : 0.f;
        float inc = warp_inclusive_scan<float>(w, lane);
        total = total + __shfl_sync(0xffffffffu, inc, 31);
      }
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        const bool live = j < k;
        float thr = 0.f;
        if (live) thr = u32_to_unit(philox_u32(rng_key, (uint64_t)i, (uint32_t)j)) * total;
        int mypos = deg - 1;  // u == 1 / rounding: clamp like MIN(item, deg - 1), :212
        bool done = !live;
        float running = 0.f;
        for (int c = 0; c < deg; c += 32) {
          float w = (c + lane < deg) ? fmaxf(wrow[c + lane], 0.f) : 0.f;
          float inc = warp_inclusive_scan<float>(w, lane);
          const float cdf = running + inc;
          const float chunk_end = running + __shfl_sync(0xffffffffu, inc, 31);
          unsigned mask = __ballot_sync(0xffffffffu, !done && thr < chunk_end);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float r = __shfl_sync(0xffffffffu, thr, src);
            const unsigned b = __ballot_sync(0xffffffffu, cdf > r);
            if (lane == src) {
              mypos = c + __ffs(b) - 1;
              done = true;
            }
          }
          running = chunk_end;
          if (__all_sync(0xffffffffu, done)) break;
        }
        if (live) {
          ocol[j] = row[mypos];
          orow[j] = seed;
        }
      }
    }
  }
}

template <typename IdT, typename ET>
static int launch_sample(const GraphSrc &g, const IdT *seeds, int64_t num_seeds,
                         const int64_t *num_seeds_dev, int64_t k, int replace, uint64_t rng_seed,
                         IdT *out_row, IdT *out_col, int64_t capacity, int64_t *out_nnz,
                         const SampleWs &ws, cudaStream_t st) {
  const bool bias = g.probs != nullptr || g.sh_probs.p[0] != nullptr;
  {
    int grid = grid_for(num_seeds, kPlanThreads, 4);
    plan_kernel<IdT, ET><<<grid, kPlanThreads, 0, st>>>(g, seeds, num_seeds, num_seeds_dev, k,
                                                        replace, ws, out_nnz);
    DGS_LAUNCH_CHECK();
  }
  int mode = bias ? (replace ? kBiasReplace : kBias) : (replace ? kUniformReplace : kUniform);
  if (k < 0) mode = kUniform;  // pure copy
  size_t smem = 0;
  if (k > 32 && mode == kUniform) smem = (size_t)kPickWarps * k * sizeof(int);
  if (k > 0 && mode == kBias) smem = (size_t)kPickWarps * k * 2 * sizeof(float);
  int gmem_scratch = 0;
  if (smem > 96 * 1024) {  // huge num_picks: keep the selection state in the output slots instead
    smem = 0;
    gmem_scratch = 1;
  }
  int grid = grid_for(num_seeds, kPickWarps, 8);
#define DGS_PICK(M)                                                                             \
  do {                                                                                          \
    auto kern = pick_kernel<IdT, ET, M>;                                                        \
    if (smem > 48 * 1024)                                                                       \
      DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                       (int)smem));                                             \
    kern<<<grid, kPickWarps * 32, smem, st>>>(g, seeds, num_seeds, num_seeds_dev, k, rng_seed,  \
                                              ws, out_row, out_col, capacity, gmem_scratch);    \
  } while (0)
  switch (mode) {
    case kUniform: DGS_PICK(kUniform); break;
    case kUniformReplace: DGS_PICK(kUniformReplace); break;
    case kBias: DGS_PICK(kBias); break;
    default: DGS_PICK(kBiasReplace); break;
  }
#undef DGS_PICK
  DGS_LAUNCH_CHECK();
  return 0;
}

int build_graph_src(const dgs_graph_t *g, GraphSrc *out) {
  memset(out, 0, sizeof(*out));
  out->indptr = g->indptr;
  out->indices = g->indices;
  out->probs = g->probs;
  if (g->p2p_indptr) {
    DGS_REQUIRE(g->p2p_indices && g->loc_table, "graph: cached source needs indptr, indices and "
                "a location table");
    DGS_REQUIRE(g->loc_capacity > 0 && (g->loc_capacity & (g->loc_capacity - 1)) == 0,
                "graph: location-table capacity must be a power of two");
    for (int d = 0; d < g->p2p_indptr->world; ++d) {
      out->sh_indptr.p[d] = g->p2p_indptr->ptrs[d];
      out->sh_indices.p[d] = g->p2p_indices->ptrs[d];
      if (g->p2p_probs) out->sh_probs.p[d] = g->p2p_probs->ptrs[d];
    }
    out->loc = (const LocSlot *)g->loc_table;
    out->cap_mask = (uint64_t)g->loc_capacity - 1;
  } else {
    DGS_REQUIRE(g->indptr && g->indices, "graph: null indptr / indices");
  }
  return 0;
}

}  // namespace dgsb

using namespace dgsb;

extern "C" int64_t dgs_sample_ws_bytes(int64_t max_seeds) {
  if (max_seeds < 1) max_seeds = 1;
  return ws_layout(max_seeds, nullptr, nullptr);
}

extern "C" int dgs_sample_neighbors(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                                    const int64_t *num_seeds_dev, int64_t num_picks, int replace,
                                    uint64_t rng_seed, void *out_row, void *out_col,
                                    int64_t out_capacity, int64_t *out_nnz_dev, void *ws,
                                    void *stream) {
  DGS_REQUIRE(g != nullptr, "dgs_sample_neighbors: null graph");
  DGS_REQUIRE(num_seeds >= 0, "dgs_sample_neighbors: negative seed count");
  DGS_REQUIRE(out_nnz_dev && ws, "dgs_sample_neighbors: null nnz / workspace");
  DGS_REQUIRE(num_picks != 0 || true, "unreachable");
  DGS_REQUIRE(!(replace && num_picks < 0),
              "dgs_sample_neighbors: num_picks=-1 (all neighbours) cannot be combined with replace");
  cudaStream_t st = (cudaStream_t)stream;
  if (num_seeds == 0) {
    DGS_CUDA_OK(cudaMemsetAsync(out_nnz_dev, 0, sizeof(int64_t), st));
    return 0;
  }
  DGS_REQUIRE(seeds != nullptr, "dgs_sample_neighbors: null seeds");
  DGS_REQUIRE(out_capacity == 0 || (out_row && out_col), "dgs_sample_neighbors: null outputs");
  GraphSrc src;
  if (build_graph_src(g, &src)) return 1;
  SampleWs w;
  ws_layout(num_seeds, (char *)ws, &w);
  DGS_ITYPE_SWITCH(g->itype, IdT, {
    DGS_ITYPE_SWITCH(g->etype, ET, {
      return launch_sample<IdT, ET>(src, (const IdT *)seeds, num_seeds, num_seeds_dev, num_picks,
                                    replace, rng_seed, (IdT *)out_row, (IdT *)out_col,
                                    out_capacity, out_nnz_dev, w, st);
    });
  });
  return 0;
}

// ------------------------------------------------------------------------------------------
// Whole-batch driver: every hop's sample + relabel is enqueued back to back with device-side
// seed / edge counts, so a multi-hop batch costs no host round trip until the caller reads the
// 2*L counts.  Replaces the layer loops P2PCacheNodeClassificationSample{Uniform,Bias}
// (src/sampling/sampler.cc:14-62), which sync twice per hop.
extern "C" int dgs_sample_blocks(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                                 int num_layers, const int64_t *fan_out, int replace,
                                 uint64_t rng_seed, void *const *out_frontier,
                                 void *const *out_row, void *const *out_col,
                                 const int64_t *cap_edges, const int64_t *cap_frontier,
                                 int64_t *counts_dev, void *sample_ws, void *relabel_table,
                                 int64_t relabel_capacity, void *relabel_ws, void *stream) {
  DGS_REQUIRE(g && fan_out && out_frontier && out_row && out_col && cap_edges && cap_frontier &&
                  counts_dev && sample_ws && relabel_table && relabel_ws,
              "dgs_sample_blocks: null argument");
  DGS_REQUIRE(num_layers >= 1 && num_layers <= 16, "dgs_sample_blocks: 1..16 layers supported");
  DGS_REQUIRE(num_seeds >= 0, "dgs_sample_blocks: negative seed count");
  cudaStream_t st = (cudaStream_t)stream;
  if (num_seeds == 0) {
    DGS_CUDA_OK(cudaMemsetAsync(counts_dev, 0, sizeof(int64_t) * 2 * num_layers, st));
    return 0;
  }
  const void *cur_seeds = seeds;
  int64_t cur_ub = num_seeds;
  const int64_t *cur_count_dev = nullptr;
  for (int l = 0; l < num_layers; ++l) {
    // the fan-out list is walked from the back, like the reference (sampler.cc:20) and DGL
    const int64_t k = fan_out[num_layers - 1 - l];
    DGS_REQUIRE(k >= 0, "dgs_sample_blocks: fan_out must be >= 0 here (use the per-hop entry for "
                "-1 / full neighbourhoods)");
    const int64_t nnz_ub = cur_ub * k;
    DGS_REQUIRE(cap_edges[l] >= nnz_ub, "dgs_sample_blocks: layer %d edge capacity %lld < %lld", l,
                (long long)cap_edges[l], (long long)nnz_ub);
    DGS_REQUIRE(cap_frontier[l] >= cur_ub + nnz_ub,
                "dgs_sample_blocks: layer %d frontier capacity %lld < %lld", l,
                (long long)cap_frontier[l], (long long)(cur_ub + nnz_ub));
    DGS_REQUIRE(relabel_capacity >= 2 * (cur_ub + nnz_ub),
                "dgs_sample_blocks: relabel table too small for layer %d", l);
    int64_t *nnz_dev = counts_dev + 2 * l;
    int64_t *nfront_dev = counts_dev + 2 * l + 1;
    // distinct Philox key per hop
    const uint64_t key = rng_seed + 0x9E3779B97F4A7C15ull * (uint64_t)(l + 1);
    int rc = dgs_sample_neighbors(g, cur_seeds, cur_ub, cur_count_dev, k, replace, key, out_row[l],
                                  out_col[l], cap_edges[l], nnz_dev, sample_ws, stream);
    if (rc) return rc;
    const void *map_ptrs[2] = {cur_seeds, out_col[l]};
    const int64_t map_counts[2] = {cur_ub, nnz_ub};
    const int64_t *map_counts_dev[2] = {cur_count_dev, nnz_dev};
    const void *rel_ptrs[2] = {out_row[l], out_col[l]};
    const int64_t rel_counts[2] = {nnz_ub, nnz_ub};
    const int64_t *rel_counts_dev[2] = {nnz_dev, nnz_dev};
    void *rel_out[2] = {out_row[l], out_col[l]};  // relabelled in place
    rc = dgs_relabel(g->itype, 2, map_ptrs, map_counts, map_counts_dev, 2, rel_ptrs, rel_counts,
                     rel_counts_dev, rel_out, out_frontier[l], nfront_dev, relabel_table,
                     relabel_capacity, relabel_ws, stream);
    if (rc) return rc;
    cur_seeds = out_frontier[l];
    cur_ub = cur_ub + nnz_ub;
    cur_count_dev = nfront_dev;
  }
  return 0;
}
