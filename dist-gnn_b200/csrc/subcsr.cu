// subcsr.cu - build-time sub-CSR extraction for the cached node set, and the cache-policy
// heat propagation.
//
// Replaces
//   ExtractIndptr / ExtractEdgeData          src/sampling/cuda/utils.cu:12-101
//   ComputeFrontierHeat{,WithBias}           src/cache/cuda/preprocess_heat.cu:14-121
// The source CSR may live in pinned host memory (the reference reads it through UVA too).
#include "dgs_common.cuh"

namespace dgsb {

constexpr int kScThreads = 256;
constexpr int kScItems = 8;
constexpr int kScTile = kScThreads * kScItems;

struct ScanWs {
  unsigned int *done;
  long long *tile_prefix;
};
static int64_t sc_layout(int64_t n, char *base, ScanWs *ws) {
  int64_t tiles = (n + kScTile - 1) / kScTile;
  if (ws) {
    ws->done = (unsigned int *)base;
    ws->tile_prefix = (long long *)(base + 256);
  }
  return 256 + (tiles + 1) * 8;
}

template <typename IdT, typename ET>
__global__ void __launch_bounds__(kScThreads)
degree_scan_kernel(const IdT *__restrict__ nids, int64_t n, const ET *__restrict__ indptr,
                   ET *__restrict__ sub_indptr, ScanWs ws) {
  __shared__ long long s_scan[32];
  __shared__ long long s_total;
  __shared__ bool s_last;
  const int64_t tiles = (n + kScTile - 1) / kScTile;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t base = tile * kScTile + (int64_t)threadIdx.x * kScItems;
    long long d[kScItems];
    long long mine = 0;
#pragma unroll
    for (int u = 0; u < kScItems; ++u) {
      const int64_t i = base + u;
      d[u] = 0;
      if (i < n) {
        const long long nid = (long long)nids[i];
        d[u] = (long long)indptr[nid + 1] - (long long)indptr[nid];
      }
      mine += d[u];
    }
    long long excl = block_exclusive_scan<long long>(mine, s_scan, &s_total);
#pragma unroll
    for (int u = 0; u < kScItems; ++u) {
      const int64_t i = base + u;
      if (i < n) sub_indptr[i] = (ET)excl;
      excl += d[u];
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
    for (int64_t b = 0; b < tiles; b += kScThreads) {
      const int64_t t = b + threadIdx.x;
      long long val = t < tiles ? tp[t] : 0;
      long long excl = block_exclusive_scan<long long>(val, s_scan, &s_total);
      if (t < tiles) tp[t] = carry + excl;
      carry += s_total;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      tp[tiles] = carry;
      *ws.done = 0;
    }
  }
}

template <typename ET>
__global__ void add_tile_prefix_kernel(ET *__restrict__ sub_indptr, int64_t n, ScanWs ws) {
  const int64_t tiles = (n + kScTile - 1) / kScTile;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n)
      sub_indptr[i] = (ET)((long long)sub_indptr[i] + ws.tile_prefix[i / kScTile]);
    else
      sub_indptr[n] = (ET)ws.tile_prefix[tiles];
  }
}

// warp per cached node: copy its edge-data row into the compacted shard
template <typename IdT, typename ET, typename VT>
__global__ void __launch_bounds__(256)
extract_edge_data_kernel(const IdT *__restrict__ nids, int64_t n, const ET *__restrict__ indptr,
                         const ET *__restrict__ sub_indptr, const VT *__restrict__ edge_data,
                         VT *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += warps_total) {
    const long long nid = (long long)nids[i];
    const long long b = (long long)indptr[nid];
    const long long deg = (long long)indptr[nid + 1] - b;
    const VT *src = edge_data + b;
    VT *dst = out + (long long)sub_indptr[i];
    long long j = lane;
    for (; j + 96 < deg; j += 128) {
      VT a = src[j], c = src[j + 32], d = src[j + 64], e = src[j + 96];
      dst[j] = a; dst[j + 32] = c; dst[j + 64] = d; dst[j + 96] = e;
    }
    for (; j < deg; j += 32) dst[j] = src[j];
  }
}

// warp per seed heat propagation
template <typename IdT, typename ET>
__global__ void __launch_bounds__(256)
frontier_heat_kernel(const IdT *__restrict__ seeds, int64_t n, const ET *__restrict__ indptr,
                     const IdT *__restrict__ indices, const float *__restrict__ probs,
                     const float *__restrict__ seeds_heat, float *__restrict__ frontier_heat,
                     int64_t num_picks, int64_t indptr_diff) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += warps_total) {
    const long long row = (long long)seeds[i];
    const long long b = (long long)indptr[row] - indptr_diff;
    const long long deg = (long long)indptr[row + 1] - indptr_diff - b;
    if (deg <= 0) continue;
    const float h = seeds_heat[row];
    if (probs == nullptr) {
      // MIN(1, heat * num_picks / degree), preprocess_heat.cu:29 (evaluated in float)
      const float msg = fminf(1.f, h * (float)num_picks / (float)deg);
      for (long long j = lane; j < deg; j += 32) atomicAdd(frontier_heat + indices[b + j], msg);
    } else {
      float s = 0.f;
      for (long long j = lane; j < deg; j += 32) s += probs[b + j];
      s = warp_sum<float>(s);
      for (long long j = lane; j < deg; j += 32) {
        const float msg = fminf(1.f, h * (float)num_picks * (probs[b + j] / s));
        atomicAdd(frontier_heat + indices[b + j], msg);
      }
    }
  }
}

}  // namespace dgsb

using namespace dgsb;

extern "C" int64_t dgs_extract_indptr_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return sc_layout(n, nullptr, nullptr);
}

extern "C" int dgs_extract_indptr(int itype, int etype, const void *nids, int64_t n,
                                  const void *indptr, void *sub_indptr, void *scan_ws,
                                  void *stream) {
  DGS_REQUIRE(n >= 0 && sub_indptr && scan_ws, "dgs_extract_indptr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    DGS_CUDA_OK(cudaMemsetAsync(sub_indptr, 0, etype == DGS_I64 ? 8 : 4, st));
    return 0;
  }
  DGS_REQUIRE(nids && indptr, "dgs_extract_indptr: null input");
  ScanWs w;
  sc_layout(n, (char *)scan_ws, &w);
  int grid = grid_for(n, kScTile, 4);
  int grid2 = grid_for(n + 1, 256, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    DGS_ITYPE_SWITCH(etype, ET, {
      degree_scan_kernel<IdT, ET><<<grid, kScThreads, 0, st>>>((const IdT *)nids, n,
                                                               (const ET *)indptr,
                                                               (ET *)sub_indptr, w);
      DGS_LAUNCH_CHECK();
      add_tile_prefix_kernel<ET><<<grid2, 256, 0, st>>>((ET *)sub_indptr, n, w);
      DGS_LAUNCH_CHECK();
    });
  });
  return 0;
}

extern "C" int dgs_extract_edge_data(int itype, int etype, int elem_bytes, const void *nids,
                                     int64_t n, const void *indptr, const void *sub_indptr,
                                     const void *edge_data, void *sub_edge_data, void *stream) {
  DGS_REQUIRE(n >= 0, "dgs_extract_edge_data: negative n");
  if (n == 0) return 0;
  DGS_REQUIRE(nids && indptr && sub_indptr && edge_data, "dgs_extract_edge_data: null input");
  DGS_REQUIRE(elem_bytes == 4 || elem_bytes == 8,
              "dgs_extract_edge_data: element size must be 4 or 8 (int32/int64/float32), got %d",
              elem_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n, 8, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    DGS_ITYPE_SWITCH(etype, ET, {
      if (elem_bytes == 8)
        extract_edge_data_kernel<IdT, ET, long long><<<grid, 256, 0, st>>>(
            (const IdT *)nids, n, (const ET *)indptr, (const ET *)sub_indptr,
            (const long long *)edge_data, (long long *)sub_edge_data);
      else
        extract_edge_data_kernel<IdT, ET, int><<<grid, 256, 0, st>>>(
            (const IdT *)nids, n, (const ET *)indptr, (const ET *)sub_indptr,
            (const int *)edge_data, (int *)sub_edge_data);
      DGS_LAUNCH_CHECK();
    });
  });
  return 0;
}

extern "C" int dgs_frontier_heat(int itype, int etype, const void *seeds, int64_t n,
                                 const void *indptr, const void *indices, const float *probs,
                                 const float *seeds_heat, float *frontier_heat, int64_t num_picks,
                                 int64_t indptr_diff, void *stream) {
  DGS_REQUIRE(n >= 0, "dgs_frontier_heat: negative n");
  if (n == 0) return 0;
  DGS_REQUIRE(seeds && indptr && indices && seeds_heat && frontier_heat,
              "dgs_frontier_heat: null input");
  int grid = grid_for(n, 8, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    DGS_ITYPE_SWITCH(etype, ET, {
      frontier_heat_kernel<IdT, ET><<<grid, 256, 0, (cudaStream_t)stream>>>(
          (const IdT *)seeds, n, (const ET *)indptr, (const IdT *)indices, probs, seeds_heat,
          frontier_heat, num_picks, indptr_diff);
      DGS_LAUNCH_CHECK();
    });
  });
  return 0;
}

// ------------------------------------------------------------------ block construction
// The sampled COO of a hop leaves dgs_sample_blocks / dgs_sample_neighbors sorted by destination
// (coo_row ascending: seed i owns one contiguous run), so the CSC a message-passing layer wants -
// what dgl.create_block((coo_col, coo_row)) builds lazily in the caller
// (example/graphsage/node_classification.py:18-28) - is just the run boundaries:
// indptr[r] = first edge whose row is >= r.  Thread e closes the runs that end at edge e.
template <typename IdT>
__global__ void __launch_bounds__(256) row_runs_to_indptr_kernel(const IdT *__restrict__ row, int64_t nnz,
                                                                 int64_t num_rows,
                                                                 IdT *__restrict__ indptr,
                                                                 int *__restrict__ unsorted) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = e == 0 ? -1 : (int64_t)row[e - 1];
    int64_t hi = e == nnz ? num_rows : (int64_t)row[e];
    if (hi < lo || hi > num_rows || (e < nnz && hi == num_rows)) {
      if (unsorted) *unsorted = 1;
      continue;
    }
    for (int64_t r = lo + 1; r <= hi; ++r) indptr[r] = (IdT)e;
  }
}

extern "C" int dgs_coo_rows_to_indptr(int itype, const void *sorted_rows, int64_t nnz, int64_t num_rows,
                                      void *indptr, int *unsorted_flag_dev, void *stream) {
  DGS_REQUIRE(nnz >= 0 && num_rows >= 0 && indptr, "dgs_coo_rows_to_indptr: bad argument");
  DGS_REQUIRE(nnz == 0 || sorted_rows, "dgs_coo_rows_to_indptr: null rows");
  int grid = grid_for(nnz + 1, 256, 8);
  DGS_ITYPE_SWITCH(itype, IdT, {
    row_runs_to_indptr_kernel<IdT><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const IdT *)sorted_rows, nnz, num_rows, (IdT *)indptr, unsorted_flag_dev);
    DGS_LAUNCH_CHECK();
  });
  return 0;
}
