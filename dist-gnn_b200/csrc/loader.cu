// loader.cu - one C-ABI call per mini-batch (the whole-batch pipeline).
//
// The reference's training loop makes three plugin calls per batch with a host synchronisation
// after the sampling (example/graphsage/node_classification.py:219-230).  dgs_load_batch enqueues
// the same work back to back on the caller's stream -
//   seeds H2D -> label gather -> labels D2H -> multi-hop sample + relabel (dgs_sample_blocks_enqueue)
//   -> feature extract of the input frontier, whose size is read on the device (dgs_extract_dyn)
// - and then makes the batch's ONE host round trip (the hop sizes the sampling kernel writes into
// pinned memory).  Everything is built from the library's own public entries; the point of doing it
// in one call is the host side: one FFI crossing and no Python between the enqueues (measured:
// 0.192 -> 0.177 ms per end-to-end step at batch 1024, profiles/r02_results.md).
//
// What is valid on return: the hop sizes and the host copy of the labels (they only depend on the
// seeds, so they are gathered and copied back BEFORE the sampling kernel and waited for through an
// event - the call does not synchronise the stream); blocks, features and device labels are
// ordinary stream-ordered results: the extract may still be running, exactly like the output of any
// CUDA op, and the next batch's sampling queues behind it while the host is already building views.
//
// Why not a CUDA graph: the outputs of a batch are fresh caller-owned buffers and the RNG key
// changes, so every node's parameters differ from step to step - an exec-update per launch costs
// the CPU about what the plain launches cost.
#include <chrono>

#include "dgs_common.cuh"

using namespace dgsb;

// DGS_LOADER_TRACE=1: host-clock breakdown of the call (us, averaged over 100 calls) on stderr
namespace {
struct LoaderTrace {
  bool on = getenv("DGS_LOADER_TRACE") != nullptr;
  double acc[8] = {0};
  int n = 0;
  std::chrono::steady_clock::time_point t;
  void start() { if (on) t = std::chrono::steady_clock::now(); }
  void lap(int i) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::micro>(now - t).count();
    t = now;
  }
  void done() {
    if (!on || ++n < 100) return;
    fprintf(stderr, "[dgs loader trace us] h2d %.1f labels %.1f sample-enqueue %.1f extract-enqueue %.1f "
            "wait-counts %.1f labels-event %.1f\n", acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n,
            acc[5] / n);
    for (double &a : acc) a = 0;
    n = 0;
  }
};
LoaderTrace g_trace;
}  // namespace

extern "C" int dgs_load_batch(const dgs_graph_t *g, const dgs_features_t *f, const void *seeds,
                              int seeds_on_host, void *seeds_dev, int64_t num_seeds, int num_layers,
                              const int64_t *fan_out, int replace, uint64_t rng_seed, void *arena,
                              const int64_t *hop_offsets, const int64_t *cap_edges,
                              const int64_t *cap_frontier, int64_t counts_offset, void *ws,
                              int64_t ws_bytes, int64_t epoch, int64_t *counts_host, void *x_out,
                              int64_t x_rows_ub, void *labels_out_dev, void *labels_out_host, int algo,
                              void *stream) {
  DGS_REQUIRE(g && f && seeds && fan_out && arena && hop_offsets && cap_edges && cap_frontier && ws &&
                  counts_host && x_out,
              "dgs_load_batch: null argument");
  DGS_REQUIRE(num_seeds >= 1 && num_layers >= 1 && num_layers <= 16, "dgs_load_batch: bad sizes");
  DGS_REQUIRE(!seeds_on_host || seeds_dev, "dgs_load_batch: host seeds need a device staging buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t es = g->itype == DGS_I64 ? 8 : 4;
  const void *sd = seeds;
  g_trace.start();
  if (seeds_on_host) {
    DGS_CUDA_OK(cudaMemcpyAsync(seeds_dev, seeds, (size_t)(num_seeds * es), cudaMemcpyHostToDevice, st));
    sd = seeds_dev;
  }
  void *fr[16], *row[16], *col[16];
  char *base = (char *)arena;
  for (int l = 0; l < num_layers; ++l) {
    fr[l] = base + hop_offsets[3 * l] * es;
    row[l] = base + hop_offsets[3 * l + 1] * es;
    col[l] = base + hop_offsets[3 * l + 2] * es;
  }
  int64_t *counts_dev = (int64_t *)(base + counts_offset * es);
  g_trace.lap(0);
  int rc;
  // labels first: they depend on the seeds only, so their host copy is complete long before the
  // hop sizes arrive and the call never has to wait for the extract
  cudaEvent_t labels_ready = nullptr;
  if (f->labels != nullptr && labels_out_dev != nullptr) {
    rc = dgs_index_select(f->labels, f->label_bytes, g->itype, sd, num_seeds, labels_out_dev, 1, stream);
    if (rc) return rc;
    if (labels_out_host != nullptr) {
      DGS_CUDA_OK(cudaMemcpyAsync(labels_out_host, labels_out_dev, (size_t)(num_seeds * f->label_bytes),
                                  cudaMemcpyDeviceToHost, st));
      static thread_local cudaEvent_t ev[DGS_MAX_DEVICES] = {nullptr};   // one per device, reused
      int dev = 0;
      DGS_CUDA_OK(cudaGetDevice(&dev));
      DGS_REQUIRE(dev >= 0 && dev < DGS_MAX_DEVICES, "dgs_load_batch: device ordinal %d", dev);
      if (ev[dev] == nullptr) DGS_CUDA_OK(cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming));
      labels_ready = ev[dev];
      DGS_CUDA_OK(cudaEventRecord(labels_ready, st));
    }
  }
  g_trace.lap(1);
  rc = dgs_sample_blocks_enqueue(g, sd, num_seeds, num_layers, fan_out, replace, rng_seed, fr, row, col,
                                 cap_edges, cap_frontier, counts_dev, ws, ws_bytes, epoch, counts_host,
                                 stream);
  if (rc) return rc;
  g_trace.lap(2);
  // extract of the last hop's frontier: its live size is counts_dev[2 L - 1]
  rc = dgs_extract_dyn(f->table, f->feat, f->loc_table, f->loc_capacity, f->mod_world, f->row_bytes,
                       g->itype, fr[num_layers - 1], x_rows_ub, counts_dev + 2 * num_layers - 1, x_out, algo,
                       stream);
  if (rc) return rc;
  g_trace.lap(3);
  // the one host round trip: hop sizes (the extract enqueued above may still run)
  rc = dgs_sample_blocks_wait(counts_host, counts_dev, num_layers, stream);
  if (rc) return rc;
  g_trace.lap(4);
  if (labels_ready != nullptr) DGS_CUDA_OK(cudaEventSynchronize(labels_ready));
  g_trace.lap(5);
  g_trace.done();
  return 0;
}
