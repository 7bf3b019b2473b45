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
#include "sampling_device.cuh"

namespace dgsb {

constexpr int kPlanThreads = 256;  // = seeds per plan tile
constexpr int kPickWarps = 8;

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
      long long begin, deg;
      int dev;
      resolve_seed<ET>(g, nid, &dev, &begin, &deg);
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
template <typename IdT>
struct CooEmit {
  IdT *orow, *ocol;
  IdT seed;
  __device__ __forceinline__ void operator()(int j, IdT v) {
    ocol[j] = v;
    orow[j] = seed;
  }
};

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

    const float *wrow = nullptr;
    if (MODE == kBias || MODE == kBiasReplace)
      wrow = (dev < 0 ? g.probs : reinterpret_cast<const float *>(g.sh_probs.p[dev])) + begin;
    CooEmit<IdT> emit{orow, ocol, seed};
    warp_select<IdT, MODE, CooEmit<IdT>>(row, wrow, deg, k, copy_path, rng_key, (uint64_t)i, lane,
                                        w_idx, w_key, emit);
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

// Test hook: the A-Res key of every edge t of one row, exactly as every biased selection path
// computes it (ares_key over Philox word t of (rng_key, item)).  The k largest keys (ties: smaller
// position first) ARE the weighted sample without replacement, so a test can check any selection
// path exactly with a top-k over this array.
__global__ void ares_keys_kernel(const float *__restrict__ w, int64_t deg, uint64_t rng_key,
                                 uint64_t item, float *__restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < deg;
       t += (int64_t)gridDim.x * blockDim.x)
    out[t] = ares_key(philox_u32(rng_key, item, (uint32_t)t), w[t]);
}

int build_graph_src(const dgs_graph_t *g, GraphSrc *out) {
  memset(out, 0, sizeof(*out));
  out->indptr = g->indptr;
  out->indices = g->indices;
  out->probs = g->probs;
  if (g->p2p_indptr) {
    DGS_REQUIRE(g->p2p_indices && (g->loc_table || g->loc_mod_world > 0),
                "graph: cached source needs indptr, indices and a location table (or modulo sharding)");
    DGS_REQUIRE(g->loc_mod_world > 0 ||
                    (g->loc_capacity > 0 && (g->loc_capacity & (g->loc_capacity - 1)) == 0),
                "graph: location-table capacity must be a power of two");
    DGS_REQUIRE(g->loc_mod_world == 0 || g->loc_mod_world == g->p2p_indptr->world,
                "graph: loc_mod_world must equal the p2p world size");
    out->mod_world = g->loc_mod_world;
    for (int d = 0; d < g->p2p_indptr->world; ++d) {
      out->sh_indptr.p[d] = g->p2p_indptr->ptrs[d];
      out->sh_indices.p[d] = g->p2p_indices->ptrs[d];
      if (g->p2p_probs) out->sh_probs.p[d] = g->p2p_probs->ptrs[d];
    }
    out->loc = g->loc_mod_world > 0 ? nullptr : (const LocSlot *)g->loc_table;
    out->cap_mask = g->loc_mod_world > 0 ? 0 : (uint64_t)g->loc_capacity - 1;
  } else {
    DGS_REQUIRE(g->indptr && g->indices, "graph: null indptr / indices");
  }
  return 0;
}

}  // namespace dgsb

using namespace dgsb;

extern "C" int dgs_debug_ares_keys(const float *weights, int64_t deg, uint64_t rng_key, uint64_t item,
                                   float *keys_out, void *stream) {
  DGS_REQUIRE(deg >= 0 && deg < (1ll << 32), "dgs_debug_ares_keys: bad degree");
  if (deg == 0) return 0;
  DGS_REQUIRE(weights && keys_out, "dgs_debug_ares_keys: null pointer");
  ares_keys_kernel<<<grid_for(deg, 256, 8), 256, 0, (cudaStream_t)stream>>>(weights, deg, rng_key, item,
                                                                          keys_out);
  DGS_LAUNCH_CHECK();
  return 0;
}

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

