/*
 * dgs_b200.h - C-ABI of the B200-native (sm_100a) mini-batch data path that replaces
 * CommediaJW/Dist-GNN's `dgs` plugin kernels.
 *
 * Boundary rules
 *   - extern "C", plain pointers + sizes, no torch / pybind types.
 *   - every entry returns 0 on success, non-zero on error; dgs_last_error() gives the text
 *     (argument errors are recoverable; the reference exit()s / abort()s on CUDA / NCCL errors,
 *     src/common/dgs_headers.h:11-34 - we report instead).
 *   - the library never allocates per-call memory: all inputs, outputs and workspaces are
 *     caller-owned device buffers (the Python host allocates them with torch), everything is
 *     enqueued on the caller's stream and nothing synchronises unless the entry says so.
 *   - ids are int32 or int64 (DGS_ID_TYPE_SWITCH, src/common/dgs_headers.h:47-58), indptr has
 *     its own int32/int64 type, edge weights are float32.
 *
 * Each block below cites the reference interface (file:line under the reference tree) it
 * replaces.  The Python binding that a reference maintainer would add is in INTEGRATION.md.
 */
#ifndef DGS_B200_H_
#define DGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGS_B200_ABI_VERSION 2
#define DGS_MAX_DEVICES 16
#define DGS_MAX_BATCHES 16 /* mini-batches per dgs_sample_blocks_multi launch */

typedef enum { DGS_I32 = 0, DGS_I64 = 1 } dgs_itype_t;

/* ------------------------------------------------------------------ misc / errors */
int dgs_abi_version(void);
const char *dgs_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t dgs_launch_count(void);
/* SM count of the current device (grid sizing), cached. */
int dgs_sm_count(void);

/* ------------------------------------------------------------------ random engine
 * replaces RandomEngine / ctx::randn_uint64  (src/context/context.h:7-20, context.cc:6).
 * dgs_seed() is an extension: the reference seeds from std::random_device and is not
 * reproducible. */
uint64_t dgs_randn_uint64(void);
void dgs_seed(uint64_t seed);

/* ------------------------------------------------------------------ pin memory
 * replaces TensorPinMemory / TensorUnpinMemory  (src/common/pin_memory.cc:7-19). */
int dgs_host_register(void *host_ptr, size_t nbytes);
int dgs_host_unregister(void *host_ptr);
/* cudaLimitMaxL2FetchGranularity of the current device (32 / 64 / 128 bytes): random row gathers
 * could over-fetch less with 32 (tuning knob; device-wide, never set implicitly; measured on B200:
 * no effect on the gather, tools/l2_granularity_probe.py). */
int dgs_set_l2_fetch_granularity(int bytes);
/* single-process multi-GPU use (tests / probes): let kernels on the current device dereference
 * memory of peer_device (cudaDeviceEnablePeerAccess). */
int dgs_enable_peer_access(int peer_device);

/* ------------------------------------------------------------------ NCCL context
 * replaces nccl::GetUniqueId / SetNCCL / NCCLContext::{Barrier_, NCCLTensorAllGather_}
 * (src/nccl/nccl_context.cc:13-112).  One process-global communicator, as in the reference. */
int dgs_nccl_get_unique_id(int64_t out_id[16]);
int dgs_nccl_set(int nranks, const int64_t id[16], int rank);
int dgs_nccl_rank(void);   /* 0 before dgs_nccl_set */
int dgs_nccl_world(void);  /* 1 before dgs_nccl_set */
int dgs_nccl_barrier(void);
/* all-gather of one int64 per rank -> host array out[world]. */
int dgs_nccl_allgather_i64(int64_t value, int64_t *out_host);
/* variable-size all-gather of raw bytes: rank r's send buffer lands in recv_dev[r] on every
 * rank (recv_dev[rank] may equal send_dev: skipped).  Blocking. */
int dgs_nccl_allgatherv(const void *send_dev, int64_t send_bytes, void *const *recv_dev,
                        const int64_t *recv_bytes);

/* ------------------------------------------------------------------ tensor p2p server
 * replaces cache::TensorP2PServer + tensor_p2p_server_wrapper::At
 * (src/cache/tensor_p2p_cache.h:11-73, tensor_p2p_cache.cc:11-132).
 * A shard is a copy of the caller's device buffer in a cuMemCreate (VMM) allocation whose POSIX fd
 * is handed to the other ranks of the box over unix sockets (SCM_RIGHTS) and cuMemMap'ped there;
 * every rank of the NCCL context gets a void*[world] table of mapped peer shards.  (Fallback when
 * VMM is unavailable on some rank, or DGS_P2P_LEGACY_IPC=1: raw cudaMalloc + CUDA-IPC handles, the
 * reference's mechanism - measured 8x slower for random row reads on B200.) */
typedef struct dgs_p2p_server dgs_p2p_server_t;
/* collective over the NCCL context (world 1: purely local).  dev_src == NULL allocates the shard
 * without copying (the caller fills dgs_p2p_server_ptr(s, rank) in place). */
int dgs_p2p_server_create(const void *dev_src, int64_t nbytes, dgs_p2p_server_t **out);
/* non-collective: adopt `world` pointers that already live in this process (emulated ranks on
 * one GPU for tests; shards are NOT owned / freed). */
int dgs_p2p_server_create_virtual(int world, int rank, void *const *ptrs, const int64_t *nbytes,
                                  dgs_p2p_server_t **out);
void *dgs_p2p_server_ptr(const dgs_p2p_server_t *s, int dev);
int64_t dgs_p2p_server_nbytes(const dgs_p2p_server_t *s, int dev);
int dgs_p2p_server_world(const dgs_p2p_server_t *s);
int dgs_p2p_server_rank(const dgs_p2p_server_t *s);
/* collective when barrier != 0 (the reference barriers in its destructor,
 * tensor_p2p_cache.cc:105-118). */
int dgs_p2p_server_destroy(dgs_p2p_server_t *s, int barrier);

/* ------------------------------------------------------------------ location table
 * replaces hashmap::cuda::Hashmap + CreateNidsP2PCacheHashMapCUDA
 * (src/hashmap/cuda/hashmap.h:12-95, hashmap.cu:15-77).
 * Layout: capacity x 16-byte slots {int64 key; int64 val}, key -1 = empty,
 * val = prio<<56 | dev<<48 | idx.  One 128-bit load per probe, linear probing. */
int64_t dgs_loc_table_capacity(int64_t n_unique); /* 2 * 2^(floor(log2 n)+1), hashmap.cu:20 */
int dgs_loc_table_build(void *table, int64_t capacity, int itype, int world, int rank,
                        const void *const *dev_nids, const int64_t *counts, void *stream);
/* per query: out_dev[i], out_idx[i] (same itype as ids), -1/-1 on miss. */
int dgs_loc_table_lookup(const void *table, int64_t capacity, int itype, const void *nids,
                         int64_t n, void *out_dev, void *out_idx, void *stream);
/* the three reference-visible tensors (key, idx, devid), hashmap.cu:21-35. */
int dgs_loc_table_unpack(const void *table, int64_t capacity, int itype, void *key, void *idx,
                         void *devid, void *stream);

/* ------------------------------------------------------------------ feature extract
 * replaces GetFeaturesCUDA / _IndexKernel (src/feature/cuda/feature_ops.cu:140-210) and
 * GetFeaturesP2PCacheCUDA / _IndexP2PCacheKernel (feature_ops.cu:38-138).
 * Rows are raw bytes (row_bytes = stride * element size), so any dtype works.
 * algo: 0 = auto, 1 = CTA-tile vectorised LDG/STG gather, 2 = TMA bulk (cp.async.bulk) staged
 * gather, 3 = warp-autonomous vectorised gather (auto: launches of >= 128 MB), 7 = row-aligned
 * warp gather (auto: peer shards whose rows are not multiples of 128 bytes); 4-6 = tuning variants
 * of 3 (tools/extract_probe.py). */
int dgs_index_select(const void *table, int64_t row_bytes, int itype, const void *nids, int64_t n,
                     void *out, int algo, void *stream);
/* Tuning knob (process-wide, default 8): resident CTAs per SM the gather kernels size their grid
 * for.  Lower it when a gather shares the GPU with another kernel on a second stream. */
int dgs_set_gather_ctas_per_sm(int ctas);
/* Tuning knob (process-wide): rows per CTA tile of the algo-1 gather, 1..64; 0 = default (64). */
int dgs_set_gather_tile_rows(int rows);
/* cached gather: row i comes from peer shard feat[dev][idx] when nids[i] hits the location
 * table, else from host_table[nids[i]] (pinned / registered host memory, may be NULL when every
 * id is cached). */
int dgs_extract_p2p(const dgs_p2p_server_t *feat, const void *host_table, int64_t row_bytes,
                    const void *loc_table, int64_t capacity, int itype, const void *nids,
                    int64_t n, void *out, int algo, void *stream);

/* modulo-sharded gather (extension): every row cached, row n = shard n % world, slot n / world. */
int dgs_extract_sharded(const dgs_p2p_server_t *feat, int64_t row_bytes, int itype,
                        const void *nids, int64_t n, void *out, int algo, void *stream);

/* request routing for the id-exchange variant of the sharded extract (extension; north star (4)
 * "NCCL used only for seed/ID exchange", SURVEY 8e "optional alternative to be measured against
 * pure peer loads"; closest reference code: the build-time all-gather-v of id lists,
 * src/nccl/nccl_context.cc:65-112).  Partitions n requested ids by owner (n mod world): send_idx =
 * the owners' slot numbers (n / world) grouped by owner, inv[i] = position of request i in that
 * grouped order, counts_dev[world] (int64, device) = requests per owner.  ws: dgs_route_ws_bytes. */
int64_t dgs_route_ws_bytes(int64_t n, int world);
int dgs_route_ids(int itype, const void *nids, int64_t n, int world, void *send_idx, void *inv,
                  int64_t *counts_dev, void *ws, int64_t ws_bytes, void *stream);

/* gather with a device-side row count (extension): n_dev (device int64) holds the live
 * number of rows, n_ub bounds the grid and the capacity of `out` - the extract of a mini-batch can
 * be enqueued right behind dgs_sample_blocks, before the host knows the frontier size.
 * feat == NULL: plain table; feat + mod_world > 0: modulo shards; feat + loc_table: hash cache
 * (table is then the pinned-host fallback). */
int dgs_extract_dyn(const void *table, const dgs_p2p_server_t *feat, const void *loc_table,
                    int64_t capacity, int mod_world, int64_t row_bytes, int itype, const void *nids,
                    int64_t n_ub, const int64_t *n_dev, void *out, int algo, void *stream);

/* ------------------------------------------------------------------ sub-CSR extraction
 * replaces ExtractIndptr / ExtractEdgeData (src/sampling/cuda/utils.cu:12-101).
 * indptr / edge_data may be device or mapped host memory. */
int64_t dgs_extract_indptr_ws_bytes(int64_t n); /* scan_ws size; first 256 bytes zero on entry */
int dgs_extract_indptr(int itype, int etype, const void *nids, int64_t n, const void *indptr,
                       void *sub_indptr /* n+1 */, void *scan_ws, void *stream);
int dgs_extract_edge_data(int itype, int etype, int elem_bytes, const void *nids, int64_t n,
                          const void *indptr, const void *sub_indptr, const void *edge_data,
                          void *sub_edge_data, void *stream);

/* ------------------------------------------------------------------ sampling
 * replaces RowWiseSampling{Uniform,Bias}CUDA (src/sampling/cuda/rowwise_sampling.cu:143-189,
 * rowwise_sampling_bias.cu:226-288) and the ...WithP2PCachingCUDA variants
 * (rowwise_sampling_p2p.cu:142-270, rowwise_sampling_bias_p2p.cu:227-385).
 *
 * Graph source description (where a seed's CSR row lives). */
typedef struct {
  int itype;                 /* ids: seeds / indices / outputs */
  int etype;                 /* indptr */
  /* un-cached source (plain ops, and the miss path of the cached sampler): device or mapped
   * host memory, indexed by global node id.  probs NULL => uniform. */
  const void *indptr;
  const void *indices;
  const float *probs;
  /* cached shards (NULL => plain op).  Shard-local CSR: indptr[local idx], indices, probs. */
  const dgs_p2p_server_t *p2p_indptr;
  const dgs_p2p_server_t *p2p_indices;
  const dgs_p2p_server_t *p2p_probs;
  const void *loc_table;
  int64_t loc_capacity;
  /* > 0 (= p2p world size): every node is cached and node n lives on device n % world at shard
   * slot n / world - the owner is computed, loc_table is not read (extension; the layout of the
   * sharded benchmarks). */
  int32_t loc_mod_world;
  /* number of nodes of the graph (ids are < num_nodes), 0 = unknown.  Only dgs_sample_blocks uses
   * it: with it the relabel tables are direct-addressed (8 bytes per node, one atomic per insert)
   * instead of hashed. */
  int64_t num_nodes;
} dgs_graph_t;

/* Workspace size (bytes) for a call over at most max_seeds seeds.  The first 256 bytes must be
 * zero on entry (they are left zero on exit). */
int64_t dgs_sample_ws_bytes(int64_t max_seeds);

/* One hop.  seeds: device ids; num_seeds_dev: optional device int64 holding the live seed count
 * (NULL => num_seeds is exact); num_seeds is then the upper bound used for grids / capacity.
 * num_picks < 0 => every neighbour (DGL convention; the reference has no such mode, it is
 * emulated there with num_picks >= max degree).
 * Outputs: out_row / out_col (capacity out_capacity entries, seed-major, CSR order on the copy
 * path, exactly the reference's layout), out_nnz_dev (device int64).  When out_capacity is
 * too small the call still counts (nnz is exact) and skips the seeds whose output would not fit
 * (the host re-runs with the exact size - the same result, RNG is counter based). */
int dgs_sample_neighbors(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                         const int64_t *num_seeds_dev, int64_t num_picks, int replace,
                         uint64_t rng_seed, void *out_row, void *out_col, int64_t out_capacity,
                         int64_t *out_nnz_dev, void *ws, void *stream);

/* Test hook (like the reference's _Test_* entries, src/pybind.cc:72-77): keys_out[t] = the A-Res
 * key log2(u_t) / w_t of edge t of a row of `deg` weights for (rng_key, item) - the k largest keys
 * (ties: smaller t first) are exactly what biased sampling without replacement returns for seed
 * index `item` under launch key rng_key (dgs_sample_neighbors: rng_key = rng_seed;
 * dgs_sample_blocks hop l: rng_seed + 0x9E3779B97F4A7C15 * (l + 1)). */
int dgs_debug_ares_keys(const float *weights, int64_t deg, uint64_t rng_key, uint64_t item,
                        float *keys_out, void *stream);

/* Whole mini-batch: num_layers hops (fan_out walked from the back, like the reference's
 * sampler.cc:20 and DGL), each hop sampled and relabelled: out_row / out_col are positions in
 * out_frontier[l], hop l+1's seeds are hop l's frontier.  One cooperative launch, all enqueued
 * with device-side counts; the caller reads counts_dev = {nnz_0, |frontier_0|, nnz_1, ...} (2 L
 * int64) once at the end.  Replaces P2PCacheNodeClassificationSample{Uniform,Bias}
 * (src/sampling/sampler.cc:14-62).  fan_out[i] >= 0 here.
 * Capacities: cap_edges[l] >= ub_l * k_l, cap_frontier[l] >= ub_l * (1 + k_l) with
 * ub_0 = num_seeds, ub_{l+1} = ub_l * (1 + k_l).
 * ws: dgs_sample_blocks_ws_bytes(...) bytes, initialised ONCE with dgs_sample_blocks_ws_init for
 * the same (itype, num_seeds, num_layers, fan_out, num_nodes = g->num_nodes) - the library
 * remembers what a workspace was initialised for and REFUSES a call that does not match (a
 * different layout would read the relabel tables at the wrong offsets); with direct-addressed
 * tables (num_nodes known) fewer seeds than at init are accepted.  One workspace serves one stream
 * at a time.  epoch = number of dgs_sample_blocks calls already made on this workspace since its
 * init (0, 1, 2, ...): only the hashed-table path (num_nodes = 0) uses it - the relabel table a call
 * leaves dirty there is wiped by the first kernel of the next call, which needs to know which of the
 * two alternating tables that is; the direct-table path tags its entries with an epoch it keeps
 * itself and ignores the argument.
 * One cooperative launch per batch (or 3 kernels per hop when the fan-out is 0 or too large for
 * the tile kernels), no memset, no trailing clean-up launch.
 * counts_host (optional, pinned host memory, 2 L int64): when non-NULL the call returns once the
 * counts have arrived there - the single host round trip of a batch.  If the memory is mapped
 * into the device's address space (cudaHostAlloc / cudaHostRegister under UVA) the cooperative
 * kernel writes the counts itself and the host polls for them (no copy engine, no stream
 * synchronisation; the kernel is still writing the last hop's outputs when the call returns - the
 * sizes are final before that phase - so later work on the same stream is ordered behind it as
 * usual and anything else must synchronise with the stream); otherwise they are copied and the
 * stream is synchronised. */
int64_t dgs_sample_blocks_ws_bytes(int itype, int64_t num_seeds, int num_layers,
                                   const int64_t *fan_out, int64_t num_nodes);
int dgs_sample_blocks_ws_init(void *ws, int64_t ws_bytes, int itype, int64_t num_seeds,
                              int num_layers, const int64_t *fan_out, int64_t num_nodes, void *stream);
int dgs_sample_blocks(const dgs_graph_t *g, const void *seeds, int64_t num_seeds, int num_layers,
                      const int64_t *fan_out, int replace, uint64_t rng_seed,
                      void *const *out_frontier, void *const *out_row, void *const *out_col,
                      const int64_t *cap_edges, const int64_t *cap_frontier, int64_t *counts_dev,
                      void *ws, int64_t ws_bytes, int64_t epoch, int64_t *counts_host, void *stream);

/* dgs_sample_blocks split in two (extension, used by dgs.classes.BatchLoader): _enqueue launches and
 * returns at once - the kernel delivers the hop sizes to counts_host (required; mapped pinned host
 * memory) - and _wait returns once they have arrived.  In between the caller enqueues whatever
 * follows the sampling on the same stream (extract, label gather). */
int dgs_sample_blocks_enqueue(const dgs_graph_t *g, const void *seeds, int64_t num_seeds,
                              int num_layers, const int64_t *fan_out, int replace, uint64_t rng_seed,
                              void *const *out_frontier, void *const *out_row, void *const *out_col,
                              const int64_t *cap_edges, const int64_t *cap_frontier,
                              int64_t *counts_dev, void *ws, int64_t ws_bytes, int64_t epoch,
                              int64_t *counts_host, void *stream);
int dgs_sample_blocks_wait(int64_t *counts_host, const int64_t *counts_dev, int num_layers,
                           void *stream);

/* B independent mini-batches of the same shape in ONE cooperative launch (extension):
 * a batch of ~1000 seeds cannot fill 148 SMs, so the hops of all B batches walk through the kernel's
 * phases together and share every grid barrier.  Bit-identical to B dgs_sample_blocks calls with
 * rng_seeds[b].  Batch b reads seeds + b * seeds_stride_bytes and writes out_*[l] + b *
 * out_stride_bytes (out_* / cap_* describe batch 0); counts_dev / counts_host hold [B][2 L] int64.
 * Needs direct-addressed relabel tables (g->num_nodes known), fan-outs in [1, ~120] and <= 8 hops:
 * dgs_sample_blocks_multi_ws_bytes returns -1 otherwise (use the single-batch entry).  ws: sized
 * and initialised ONCE for (itype, num_batches, num_seeds, num_layers, fan_out, num_nodes); later
 * calls may pass fewer batches / fewer seeds per batch.  The relabel table of a batch (8 bytes per
 * node) is never wiped: entries carry an 8-bit epoch tag that the library advances per hop (and
 * clears the table every 255 hops), so a workspace must be used by one stream at a time and the
 * items of a hop (seeds + padded slots) must stay below 2^24 (else -1 from ..._ws_bytes).
 * wait != 0: return once the sizes are in counts_host (pinned, required); wait == 0: enqueue only -
 * when counts_host is mapped pinned memory the kernel delivers the sizes there and
 * dgs_sample_blocks_wait(counts_host, counts_dev, num_layers * num_batches, stream) collects them. */
int64_t dgs_sample_blocks_multi_ws_bytes(int itype, int num_batches, int64_t num_seeds,
                                         int num_layers, const int64_t *fan_out, int64_t num_nodes);
int dgs_sample_blocks_multi_ws_init(void *ws, int64_t ws_bytes, int itype, int num_batches,
                                    int64_t num_seeds, int num_layers, const int64_t *fan_out,
                                    int64_t num_nodes, void *stream);
int dgs_sample_blocks_multi(const dgs_graph_t *g, int num_batches, const void *seeds,
                            int64_t seeds_stride_bytes, int64_t num_seeds, int num_layers,
                            const int64_t *fan_out, int replace, const uint64_t *rng_seeds,
                            void *const *out_frontier, void *const *out_row, void *const *out_col,
                            int64_t out_stride_bytes, const int64_t *cap_edges,
                            const int64_t *cap_frontier, int64_t *counts_dev, void *ws,
                            int64_t ws_bytes, int64_t *counts_host, int wait, void *stream);

/* ------------------------------------------------------------------ whole-batch loader
 * One call per mini-batch = the caller loop of example/graphsage/node_classification.py:219-230
 * (sampler._CAPI_sample_node_classifiction, feature_server._CAPI_get_feature, label index_select)
 * enqueued back to back with ONE host round trip: seeds H2D (when seeds_on_host: `seeds` is pinned
 * host memory, seeds_dev a device staging buffer of num_seeds ids) -> label gather into
 * labels_out_dev (+ D2H into pinned labels_out_host; the labels depend on the seeds only, so they go
 * first) -> dgs_sample_blocks_enqueue -> dgs_extract_dyn of the input frontier (size read on the
 * device; x_out holds x_rows_ub rows).  Outputs: `arena` (ids) holds hop l's frontier / row / col at
 * hop_offsets[3 l .. 3 l + 2] (elements) and the 2 L int64 hop sizes at counts_offset (elements).
 * Valid on the host when the call returns: the sizes in counts_host (pinned) and labels_out_host
 * (waited for through an event).  The call does NOT drain the stream: the last emit phase and the
 * extract may still be running, ordered on `stream` like the result of any CUDA op.  If the frontier
 * turned out larger than x_rows_ub (the caller's bound) only x_rows_ub rows were gathered. */
typedef struct {
  const void *table;            /* plain table (device / pinned host), or the pinned-host fallback of a
                                   cached source (may be NULL when every row is cached) */
  const dgs_p2p_server_t *feat; /* cached source: feature shards (NULL = plain table) */
  const void *loc_table;        /* location table of the cached source (NULL with mod_world != 0) */
  int64_t loc_capacity;
  int32_t mod_world;            /* > 0: modulo-sharded, 0: location table */
  int64_t row_bytes;
  const void *labels;           /* label table indexed by node id (device / pinned host) or NULL */
  int64_t label_bytes;          /* bytes per label row */
} dgs_features_t;
int dgs_load_batch(const dgs_graph_t *g, const dgs_features_t *f, const void *seeds, int seeds_on_host,
                   void *seeds_dev, int64_t num_seeds, int num_layers, const int64_t *fan_out,
                   int replace, uint64_t rng_seed, void *arena, const int64_t *hop_offsets,
                   const int64_t *cap_edges, const int64_t *cap_frontier, int64_t counts_offset,
                   void *ws, int64_t ws_bytes, int64_t epoch, int64_t *counts_host, void *x_out,
                   int64_t x_rows_ub, void *labels_out_dev, void *labels_out_host, int algo,
                   void *stream);

/* ------------------------------------------------------------------ relabel
 * replaces TensorRelabelCUDA (src/sampling/cuda/tensor_relabel.cu:182-205):
 * unique = first-occurrence-order unique of the concatenation of the mapping parts, every id of
 * the relabel parts is replaced by its position in `unique` (-1 if absent); rel_out[p] has the
 * shape of rel_ptrs[p].  Up to 4 parts on each side (no torch::cat needed); *_counts are exact
 * lengths or, when the matching *_counts_dev[p] (device int64) is non-NULL, upper bounds with the
 * live length on the device - this lets a whole multi-hop batch be enqueued with no host sync.
 * table: capacity x 16 B slots, capacity a power of two >= 2 * sum(map_counts), all-0xFF on entry
 * and restored to all-0xFF on exit (a persistent table never needs a memset).
 * ws: dgs_relabel_ws_bytes(sum(map_counts)) bytes whose first 256 bytes are zero on entry (they
 * are left zero on exit).  num_unique_dev: device int64. */
int64_t dgs_relabel_table_capacity(int64_t n);
int64_t dgs_relabel_table_bytes(int64_t n);
int64_t dgs_relabel_ws_bytes(int64_t n);
int dgs_relabel(int itype, int n_map, const void *const *map_ptrs, const int64_t *map_counts,
                const int64_t *const *map_counts_dev, int n_rel, const void *const *rel_ptrs,
                const int64_t *rel_counts, const int64_t *const *rel_counts_dev,
                void *const *rel_out, void *unique_out, int64_t *num_unique_dev, void *table,
                int64_t capacity, void *ws, void *stream);

/* ------------------------------------------------------------------ cache-policy heat
 * replaces ComputeFrontierHeat{,WithBias} (src/cache/cuda/preprocess_heat.cu:14-121). */
int dgs_frontier_heat(int itype, int etype, const void *seeds, int64_t n, const void *indptr,
                      const void *indices, const float *probs, const float *seeds_heat,
                      float *frontier_heat, int64_t num_picks, int64_t indptr_diff, void *stream);

/* ------------------------------------------------------------------ block construction
 * CSC row pointer of a sampled hop: replaces what dgl.create_block((coo_col, coo_row), ...) derives
 * in the caller (example/graphsage/node_classification.py:18-28).  `sorted_rows` = the hop's coo_row,
 * ascending.  Precondition: the sampling entry points emit ascending rows when the hop's seeds are
 * DISTINCT (always from the second hop on); a duplicate seed of hop 0 is relabelled to its first
 * occurrence and breaks the order - pass unsorted_flag_dev then.  indptr gets num_rows + 1 entries of
 * the id type.  *unsorted_flag_dev (optional, zeroed by the caller) is set to 1 if the rows are not
 * ascending or out of range (entries of indptr covered by the violating runs are then unwritten). */
int dgs_coo_rows_to_indptr(int itype, const void *sorted_rows, int64_t nnz, int64_t num_rows,
                           void *indptr, int *unsorted_flag_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DGS_B200_H_ */
