/*
 * dgs_oracle.c - CPU restatement of the reference's mini-batch data path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dist-gnn_b200/ may include, link or call this file;
 * it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs as the checker and the reported CPU baseline, never as the product path.
 *
 * Part 1 restates, function by function, the deterministic behaviour of the reference's CUDA
 * kernels (ids are int64, as in every reference test):
 *   orc_index_select        _IndexKernel / _IndexOneDimKernel   src/feature/cuda/feature_ops.cu:140-171
 *   orc_hashmap_*           Hashmap::{Update,SearchForPos,hash} src/hashmap/cuda/hashmap.h:18-84
 *                           CreateNidsP2PCacheHashMapCUDA       src/hashmap/cuda/hashmap.cu:15-77
 *   orc_extract_p2p         _IndexP2PCacheKernel                src/feature/cuda/feature_ops.cu:38-73
 *   orc_extract_indptr      ExtractIndptr                       src/sampling/cuda/utils.cu:12-42
 *   orc_extract_edge_data   _ExtractEdgeDataKernel              src/sampling/cuda/utils.cu:44-69
 *   orc_sample_copy_path    deg <= num_picks branch             src/sampling/cuda/rowwise_sampling.cu:71-77
 *                           + _GetSubIndptr                     src/sampling/cuda/rowwise_sampling.cu:16-45
 *   orc_relabel             Unique + Relabel                    src/sampling/cuda/tensor_relabel.cu:82-180
 *   orc_frontier_heat       _ComputeFrotierHeat{,WithBias}      src/cache/cuda/preprocess_heat.cu:14-98
 * Pinning: the reference's tests hold inputs but no expected outputs (they only print), so this
 * file is pinned (a) by the known answers derived from those inputs (tests/golden/kat_*.json,
 * SURVEY.md section 4) and (b) against the reference's own kernels compiled for sm_100a
 * (oracle/_ref, oracle/build_ref.sh) on the GPU box: tests/test_ref_differential.py, and the
 * fixtures that run wrote to tests/golden/ref_*.npz.
 *
 * Part 2 is the reported CPU baseline: a from-scratch restatement of DGL's CPU
 * sample_neighbors semantics (DGL >= 0.9.1 is the reference's only pin, README.md:9; it is not
 * installed here and not vendored, and the reference never calls dgl.sampling itself, so this
 * boundary is "parity unpinned"): per seed all neighbours if deg <= k, else k distinct uniform
 * picks (or weight-proportional without replacement), COO output, to_block-style relabel, and
 * an index_select of the feature rows; OpenMP over seeds / rows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

/* ------------------------------------------------------------------ gather */
void orc_index_select(const char *table, i64 row_bytes, const i64 *nids, i64 n, char *out) {
  for (i64 i = 0; i < n; ++i) memcpy(out + i * row_bytes, table + nids[i] * row_bytes, (size_t)row_bytes);
}

/* ------------------------------------------------------------------ reference hash map */
static uint32_t hash32shift(uint32_t k) { /* hashmap.h:51-58 */
  k ^= k >> 16;
  k *= 0x85ebca6bu;
  k ^= k >> 13;
  k *= 0xc2b2ae35u;
  k ^= k >> 16;
  return k;
}
static uint64_t hash64shift(uint64_t k) { /* hashmap.h:61-68 */
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
static uint32_t hash_key(i64 key, uint32_t cap) { return ((uint32_t)hash64shift((uint64_t)key)) & (cap - 1); }
static uint32_t hash_pos(uint32_t pos, uint32_t cap) { return hash32shift(pos) & (cap - 1); }

i64 orc_up_power(i64 key) { /* hashmap.h:92-95: 1 << (int)(log2(key) + 1) */
  return (i64)1 << (uint32_t)(log2((double)key) + 1);
}
i64 orc_hashmap_capacity(i64 n_unique) { return 2 * orc_up_power(n_unique); } /* hashmap.cu:20 */

static i64 hm_update(i64 *keys, i64 *vals, uint32_t cap, i64 key, i64 value) { /* hashmap.h:18-32 */
  uint32_t delta = 1;
  uint32_t pos = hash_key(key, cap);
  while (keys[pos] != key && keys[pos] != -1) {
    pos = hash_pos(pos + delta, cap);
    delta += 1;
  }
  keys[pos] = key;
  vals[pos] = value;
  return pos;
}
static i64 hm_search(const i64 *keys, uint32_t cap, i64 key) { /* hashmap.h:34-48 */
  uint32_t delta = 1;
  uint32_t pos = hash_key(key, cap);
  for (;;) {
    if (keys[pos] == key) return pos;
    if (keys[pos] == -1) return -1;
    pos = hash_pos(pos + delta, cap);
    delta += 1;
  }
}

/* hashmap.cu:37-72 - remote devices in the order (rank+1)%P .. (rank+P-1)%P, then the local one;
 * a later insert overwrites idx/devid.  Inside one device list the GPU threads race; lists must
 * hold unique ids, so sequential insertion gives the same lookups. */
void orc_hashmap_build(i64 *key, i64 *idx, i64 *devid, i64 cap, int world, int rank,
                       const i64 *const *dev_nids, const i64 *counts) {
  for (i64 i = 0; i < cap; ++i) key[i] = idx[i] = devid[i] = -1;
  for (int d = 0; d < world; ++d) {
    int index = (d + rank) % world;
    if (index == rank) continue;
    for (i64 i = 0; i < counts[index]; ++i) {
      i64 pos = hm_update(key, idx, (uint32_t)cap, dev_nids[index][i], i);
      devid[pos] = index;
    }
  }
  for (i64 i = 0; i < counts[rank]; ++i) {
    i64 pos = hm_update(key, idx, (uint32_t)cap, dev_nids[rank][i], i);
    devid[pos] = rank;
  }
}
void orc_hashmap_lookup(const i64 *key, const i64 *idx, const i64 *devid, i64 cap, const i64 *nids,
                        i64 n, i64 *out_dev, i64 *out_idx) {
  for (i64 i = 0; i < n; ++i) {
    i64 pos = hm_search(key, (uint32_t)cap, nids[i]);
    out_dev[i] = pos < 0 ? -1 : devid[pos];
    out_idx[i] = pos < 0 ? -1 : idx[pos];
  }
}

/* feature_ops.cu:38-73 with the lookup of :97-108 */
void orc_extract_p2p(const char *cpu_table, const char *const *shards, i64 row_bytes, const i64 *key,
                     const i64 *idx, const i64 *devid, i64 cap, const i64 *nids, i64 n, char *out) {
  for (i64 i = 0; i < n; ++i) {
    i64 pos = hm_search(key, (uint32_t)cap, nids[i]);
    const char *src = pos < 0 ? cpu_table + nids[i] * row_bytes : shards[devid[pos]] + idx[pos] * row_bytes;
    memcpy(out + i * row_bytes, src, (size_t)row_bytes);
  }
}

/* ------------------------------------------------------------------ sub-CSR */
void orc_extract_indptr(const i64 *nids, i64 n, const i64 *indptr, i64 *sub_indptr) {
  i64 acc = 0; /* degrees then exclusive sum over n + 1 entries, utils.cu:21-36 */
  for (i64 i = 0; i < n; ++i) {
    sub_indptr[i] = acc;
    acc += indptr[nids[i] + 1] - indptr[nids[i]];
  }
  sub_indptr[n] = acc;
}
void orc_extract_edge_data(const i64 *nids, i64 n, const i64 *indptr, const i64 *sub_indptr,
                           const char *edge_data, i64 elem_bytes, char *out) {
  for (i64 i = 0; i < n; ++i) { /* utils.cu:56-68 */
    i64 b = indptr[nids[i]], deg = indptr[nids[i] + 1] - b;
    memcpy(out + sub_indptr[i] * elem_bytes, edge_data + b * elem_bytes, (size_t)(deg * elem_bytes));
  }
}

/* ------------------------------------------------------------------ deterministic sampling path
 * All neighbours of every seed in CSR order, seed-major: what the reference produces when
 * num_picks >= every degree.  Returns nnz; with out_row == NULL only counts. */
i64 orc_sample_copy_path(const i64 *seeds, i64 n, const i64 *indptr, const i64 *indices, i64 *out_row,
                         i64 *out_col) {
  i64 nnz = 0;
  for (i64 i = 0; i < n; ++i) {
    i64 b = indptr[seeds[i]], e = indptr[seeds[i] + 1];
    if (out_row)
      for (i64 j = b; j < e; ++j) {
        out_row[nnz + (j - b)] = seeds[i];
        out_col[nnz + (j - b)] = indices[j];
      }
    nnz += e - b;
  }
  return nnz;
}

/* ------------------------------------------------------------------ relabel
 * tensor_relabel.cu:82-180: unique = ids of `mapping` in order of first occurrence; every id of
 * `to_relabel` -> its position in unique, -1 when absent.  Returns |unique|. */
typedef struct { i64 *keys; i64 *vals; i64 cap; } omap;
static void omap_init(omap *m, i64 n) {
  m->cap = 64;
  while (m->cap < 2 * n) m->cap <<= 1;
  m->keys = (i64 *)malloc(sizeof(i64) * (size_t)m->cap);
  m->vals = (i64 *)malloc(sizeof(i64) * (size_t)m->cap);
  for (i64 i = 0; i < m->cap; ++i) m->keys[i] = -1;
}
static i64 *omap_slot(omap *m, i64 key, int insert) {
  uint64_t pos = hash64shift((uint64_t)key) & (uint64_t)(m->cap - 1);
  for (;;) {
    if (m->keys[pos] == key) return &m->vals[pos];
    if (m->keys[pos] == -1) {
      if (!insert) return NULL;
      m->keys[pos] = key;
      m->vals[pos] = -1;
      return &m->vals[pos];
    }
    pos = (pos + 1) & (uint64_t)(m->cap - 1);
  }
}
i64 orc_relabel(const i64 *mapping, i64 n, const i64 *to_relabel, i64 m, i64 *unique, i64 *relabeled) {
  omap h;
  omap_init(&h, n);
  i64 u = 0;
  for (i64 i = 0; i < n; ++i) {
    i64 *v = omap_slot(&h, mapping[i], 1);
    if (*v < 0) {
      *v = u;
      unique[u++] = mapping[i];
    }
  }
  for (i64 j = 0; j < m; ++j) {
    i64 *v = omap_slot(&h, to_relabel[j], 0);
    relabeled[j] = v ? *v : -1;
  }
  free(h.keys);
  free(h.vals);
  return u;
}

/* ------------------------------------------------------------------ heat (float, serial order) */
void orc_frontier_heat(const i64 *seeds, i64 n, const i64 *indptr, const i64 *indices, const float *probs,
                       const float *seeds_heat, float *frontier_heat, i64 num_picks, i64 indptr_diff) {
  for (i64 i = 0; i < n; ++i) {
    i64 row = seeds[i], b = indptr[row] - indptr_diff, deg = indptr[row + 1] - indptr_diff - b;
    if (deg <= 0) continue;
    if (!probs) {
      float msg = fminf(1.f, seeds_heat[row] * (float)num_picks / (float)deg);
      for (i64 j = 0; j < deg; ++j) frontier_heat[indices[b + j]] += msg;
    } else {
      float s = 0.f;
      for (i64 j = 0; j < deg; ++j) s += probs[b + j];
      for (i64 j = 0; j < deg; ++j)
        frontier_heat[indices[b + j]] += fminf(1.f, seeds_heat[row] * (float)num_picks * (probs[b + j] / s));
    }
  }
}

/* ================================================================== Part 2: CPU baseline */
static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline uint64_t rnd_below(uint64_t *s, uint64_t n) { return (uint64_t)(((__uint128_t)splitmix64(s) * n) >> 64); }
static inline double rnd_unit(uint64_t *s) { return ((double)(splitmix64(s) >> 11) + 1.0) * (1.0 / 9007199254740993.0); }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arms of bench.py ask for all host cores
 * explicitly (n <= 0: the number of processors). */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
  (void)n;
#endif
}

/* DGL-semantics sample_neighbors (COO, seed-major).  probs == NULL: uniform.  replace = 0.
 * Two passes over the seeds: counts -> exclusive offsets -> fill.  Returns nnz; out arrays need
 * n * k entries (k >= 0) or the exact nnz when k < 0. */
i64 orc_cpu_sample_neighbors(const i64 *seeds, i64 n, const i64 *indptr, const i64 *indices, const float *probs,
                             i64 k, uint64_t rng_seed, i64 *offsets /* n + 1 */, i64 *out_row, i64 *out_col) {
  i64 i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) {
    i64 deg = indptr[seeds[i] + 1] - indptr[seeds[i]];
    offsets[i + 1] = (k < 0 || deg < k) ? deg : k;
  }
  offsets[0] = 0;
  for (i = 0; i < n; ++i) offsets[i + 1] += offsets[i];
  if (!out_row) return offsets[n];
#pragma omp parallel for schedule(dynamic, 64)
  for (i = 0; i < n; ++i) {
    const i64 row = seeds[i], b = indptr[row], deg = indptr[row + 1] - b, off = offsets[i];
    uint64_t st = rng_seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(i + 1));
    if (k < 0 || deg <= k) {
      for (i64 j = 0; j < deg; ++j) {
        out_row[off + j] = row;
        out_col[off + j] = indices[b + j];
      }
    } else if (!probs) {
      /* Floyd's subset sampling, k small */
      i64 pick[64];
      i64 *pk = k <= 64 ? pick : (i64 *)malloc(sizeof(i64) * (size_t)k);
      for (i64 t = 0; t < k; ++t) {
        i64 J = deg - k + t, r = (i64)rnd_below(&st, (uint64_t)(J + 1)), dup = 0;
        for (i64 q = 0; q < t; ++q) dup |= (pk[q] == r);
        pk[t] = dup ? J : r;
      }
      for (i64 t = 0; t < k; ++t) {
        out_row[off + t] = row;
        out_col[off + t] = indices[b + pk[t]];
      }
      if (pk != pick) free(pk);
    } else {
      /* A-Res: keep the k largest log(u)/w */
      double keyv[64];
      i64 idxv[64];
      double *kv = k <= 64 ? keyv : (double *)malloc(sizeof(double) * (size_t)k);
      i64 *iv = k <= 64 ? idxv : (i64 *)malloc(sizeof(i64) * (size_t)k);
      i64 mn = 0;
      for (i64 t = 0; t < deg; ++t) {
        double w = probs[b + t], key = w > 0 ? log(rnd_unit(&st)) / w : -INFINITY;
        if (t < k) {
          kv[t] = key;
          iv[t] = t;
          if (t == k - 1) {
            mn = 0;
            for (i64 q = 1; q < k; ++q)
              if (kv[q] < kv[mn]) mn = q;
          }
        } else if (key > kv[mn]) {
          kv[mn] = key;
          iv[mn] = t;
          mn = 0;
          for (i64 q = 1; q < k; ++q)
            if (kv[q] < kv[mn]) mn = q;
        }
      }
      for (i64 t = 0; t < k; ++t) {
        out_row[off + t] = row;
        out_col[off + t] = indices[b + iv[t]];
      }
      if (kv != keyv) { free(kv); free(iv); }
    }
  }
  return offsets[n];
}

/* to_block-style relabel with a direct-address map (scratch[num_nodes], all -1 on entry and exit). */
i64 orc_cpu_relabel(const i64 *seeds, i64 n, i64 *row, i64 *col, i64 nnz, i64 *frontier, i64 *scratch) {
  i64 u = 0;
  for (i64 i = 0; i < n; ++i)
    if (scratch[seeds[i]] < 0) { scratch[seeds[i]] = u; frontier[u++] = seeds[i]; }
  for (i64 j = 0; j < nnz; ++j)
    if (scratch[col[j]] < 0) { scratch[col[j]] = u; frontier[u++] = col[j]; }
  i64 j;
#pragma omp parallel for schedule(static)
  for (j = 0; j < nnz; ++j) {
    row[j] = scratch[row[j]];
    col[j] = scratch[col[j]];
  }
  for (i64 i = 0; i < u; ++i) scratch[frontier[i]] = -1;
  return u;
}

void orc_cpu_index_select(const char *table, i64 row_bytes, const i64 *nids, i64 n, char *out) {
  i64 i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) memcpy(out + i * row_bytes, table + nids[i] * row_bytes, (size_t)row_bytes);
}

/* One whole mini-batch on the CPU: L hops (fan_out walked from the back) + feature gather of the
 * input frontier.  Work buffers are caller-provided and sized for the worst case:
 *   buf_row/buf_col/buf_front : ub_L entries each, offsets: ub_{L-1} + 1, scratch: num_nodes (-1).
 * Returns total sampled edges; *out_rows = rows of features gathered. */
i64 orc_cpu_batch(const i64 *seeds, i64 n, const i64 *indptr, const i64 *indices, const float *probs,
                  const i64 *fan_out, int L, uint64_t rng_seed, const char *feat, i64 row_bytes, i64 *buf_row,
                  i64 *buf_col, i64 *buf_front_a, i64 *buf_front_b, i64 *offsets, i64 *scratch, char *feat_out,
                  i64 *out_rows) {
  const i64 *cur = seeds;
  i64 cur_n = n, edges = 0;
  i64 *fronts[2] = {buf_front_a, buf_front_b};
  for (int l = 0; l < L; ++l) {
    i64 k = fan_out[L - 1 - l];
    i64 nnz = orc_cpu_sample_neighbors(cur, cur_n, indptr, indices, probs, k, rng_seed + (uint64_t)l, offsets,
                                       buf_row, buf_col);
    edges += nnz;
    i64 *f = fronts[l & 1];
    i64 u = orc_cpu_relabel(cur, cur_n, buf_row, buf_col, nnz, f, scratch);
    cur = f;
    cur_n = u;
  }
  if (feat) orc_cpu_index_select(feat, row_bytes, cur, cur_n, feat_out);
  *out_rows = cur_n;
  return edges;
}
