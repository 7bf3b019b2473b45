// sampling_device.cuh - warp-level neighbour selection shared by the per-hop sampler
// (sampling.cu) and the fused whole-batch pipeline (blocks.cu).
#pragma once
#include "dgs_common.cuh"

namespace dgsb {

struct GraphSrc {
  const void *indptr;
  const void *indices;
  const float *probs;
  PtrTable sh_indptr, sh_indices, sh_probs;
  const LocSlot *loc;
  uint64_t cap_mask;
  int mod_world;  // > 0: every node is cached, node n lives on device n % mod_world at slot n / mod_world
};

enum PickMode { kUniform = 0, kUniformReplace = 1, kBias = 2, kBiasReplace = 3 };

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ uint32_t philox_u32(uint64_t key, uint64_t item, uint32_t draw) {
  uint4 r = Philox::gen(key, item, (uint64_t)(draw >> 2));
  const uint32_t c = draw & 3u;
  return c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
}

// Resolve a seed: owner device (-1 = un-cached source), first edge, degree.
template <typename ET>
__device__ __forceinline__ void resolve_seed(const GraphSrc &g, long long nid, int *dev,
                                             long long *begin, long long *deg) {
  long long v = -1;
  if (g.mod_world > 0)
    v = ((nid % g.mod_world) << kDevShift) | (nid / g.mod_world);  // arithmetic owner, no table
  else if (g.loc != nullptr)
    v = loc_lookup(g.loc, g.cap_mask, nid);
  long long b, e;
  if (v >= 0) {
    *dev = (int)((v >> kDevShift) & 0xff);
    const long long idx = v & kIdxMask;
    const ET *ip = reinterpret_cast<const ET *>(g.sh_indptr.p[*dev]);
    b = (long long)ip[idx];
    e = (long long)ip[idx + 1];
  } else {
    *dev = -1;
    const ET *ip = reinterpret_cast<const ET *>(g.indptr);
    b = (long long)ip[nid];
    e = (long long)ip[nid + 1];
  }
  *begin = b;
  *deg = e - b;
}

// Emit the k selected neighbours.  w_idx may alias the memory the emit functor writes (global
// scratch mode: the first 4 k bytes of the seed's own output slots): chunks are processed from
// the top and every lane reads its position before any lane writes, so a write for output j only
// overwrites scratch entries >= j that were already consumed.
template <typename IdT, bool kPos>
__device__ __forceinline__ IdT pick_value(const IdT *__restrict__ row, long long p) {
  return kPos ? (IdT)p : row[p];
}

template <typename IdT, typename Emit, bool kPos = false>
__device__ __forceinline__ void emit_picks(const IdT *__restrict__ row, const int *w_idx, int k,
                                           int lane, Emit &emit) {
  for (int j0 = ((k - 1) / 32) * 32; j0 >= 0; j0 -= 32) {
    const int j = j0 + lane;
    int p = 0;
    if (j < k) p = w_idx[j];
    __syncwarp();
    if (j < k) emit(j, pick_value<IdT, kPos>(row, p));
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// A-Res for k <= 32 (weighted sampling without replacement = the k largest keys log2(u_t) / w_t).
// A reservoir that replaces its minimum item by item is a chain of ~k ln(deg / k) DEPENDENT warp
// operations per row (measured: ~10 us for a 512-weight row, 25 picks).  Instead:
//   * threshold: after the first 512 weights, tau = the k-th largest of the 32 per-lane maxima -
//     a lower bound of the k-th largest key, found with one round of shuffles;
//   * collect: every later key above tau is appended to a per-warp candidate list in shared
//     memory (ballot + popc, no dependence between candidates);
//   * tighten: when the list fills up (and at the end) every candidate counts how many others
//     beat it - all comparisons independent - and the k best move to the front IN ORDER, which
//     also yields the exact new tau.
// Keys depend only on (rng_key, item, t) and the order is total (key descending, ties by
// position), so neither the selected set nor the order of the picks depends on how a row is split
// over warps: a CTA can share a hub row (blocks.cu) and still agree with the one-warp kernels.
// The picks come out in descending key order, as in the reference (rowwise_sampling_bias.cu:127).
constexpr int kAresCap = 128;   // candidates per warp

struct AresBuf {
  float *key;
  int *idx;
};

// candidate list of warp w of this CTA (all callers run 8 warps per CTA)
__device__ __forceinline__ AresBuf ares_buf(int w) {
  __shared__ float s_ck[8][kAresCap];
  __shared__ int s_ci[8][kAresCap];
  return AresBuf{s_ck[w], s_ci[w]};
}

__device__ __forceinline__ float ares_key(uint32_t r, float w) {
  return w > 0.f ? __fdividef(__log2f(u32_to_unit(r)), w) : -INFINITY;
}

// candidate a is picked before candidate b
__device__ __forceinline__ bool ares_before(float ka, int ia, float kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

// Keep the k best of the M listed candidates at the front of the list, best first.
// Returns min(M, k); tau = the k-th key when there are at least k.
__device__ __forceinline__ int ares_tighten(const AresBuf &b, int M, int k, int lane, float &tau) {
  __syncwarp();
  float kk[kAresCap / 32];
  int ii[kAresCap / 32], rk[kAresCap / 32];
#pragma unroll
  for (int r = 0; r < kAresCap / 32; ++r) {
    const int e = lane + 32 * r;
    kk[r] = e < M ? b.key[e] : -INFINITY;
    ii[r] = e < M ? b.idx[e] : 0x7fffffff;
    rk[r] = 0;
  }
  for (int j = 0; j < M; ++j) {
    const float kj = b.key[j];   // broadcast reads
    const int ij = b.idx[j];
#pragma unroll
    for (int r = 0; r < kAresCap / 32; ++r) rk[r] += ares_before(kj, ij, kk[r], ii[r]) ? 1 : 0;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kAresCap / 32; ++r) {
    if (lane + 32 * r < M && rk[r] < k) {
      b.key[rk[r]] = kk[r];
      b.idx[rk[r]] = ii[r];
    }
  }
  __syncwarp();
  if (M >= k) tau = b.key[k - 1];
  return min(M, k);
}

// Scan the passes t_begin, t_begin + t_stride, ... (512 weights each; t_begin a multiple of 512)
// of one row and leave the (up to k) best candidates among them at the front of the warp's list,
// best first.  Every lane owns 4 quads of 4 consecutive elements per pass (quad q = elements
// t0 + 128 q + 4 lane ..+3 = exactly one Philox block: draw t is component t & 3 of block t >> 2,
// the mapping of philox_u32); all 16 weight loads of a lane are issued before any is used.
__device__ __forceinline__ int ares_collect(const float *__restrict__ wrow, int deg, int k,
                                            int t_begin, int t_stride, uint64_t rng_key,
                                            uint64_t item, int lane, const AresBuf &b) {
  int M = 0;
  float tau = -INFINITY;
  bool strict = false;   // false until tau is an exact k-th key: candidates equal to tau still count
  for (int t0 = t_begin; t0 < deg; t0 += t_stride) {
    float key[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int tb = t0 + 128 * q + 4 * lane;
#pragma unroll
      for (int c = 0; c < 4; ++c) key[q][c] = (tb + c < deg) ? wrow[tb + c] : 0.f;   // weights first
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int tb = t0 + 128 * q + 4 * lane;
      if (tb < deg) {
        const uint4 r4 = Philox::gen(rng_key, item, (uint64_t)(tb >> 2));
        key[q][0] = ares_key(r4.x, key[q][0]);
        key[q][1] = ares_key(r4.y, key[q][1]);
        key[q][2] = ares_key(r4.z, key[q][2]);
        key[q][3] = ares_key(r4.w, key[q][3]);
      } else {
        key[q][0] = key[q][1] = key[q][2] = key[q][3] = -INFINITY;
      }
    }
    if (t0 == t_begin) {
      // lower bound of the k-th largest key: the k-th largest of the 32 per-lane maxima
      float m = -INFINITY;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) m = fmaxf(m, key[q][c]);
      int rank = 0;
      for (int l = 0; l < 32; ++l) {
        const float o = __shfl_sync(0xffffffffu, m, l);
        rank += (o > m || (o == m && l < lane)) ? 1 : 0;
      }
      const unsigned has = __ballot_sync(0xffffffffu, rank == k - 1);
      tau = __shfl_sync(0xffffffffu, m, __ffs(has) - 1);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (t0 + 128 * q >= deg) break;   // warp-uniform
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int t = t0 + 128 * q + 4 * lane + c;
        const float kv = key[q][c];
        const bool pred = t < deg && (strict ? kv > tau : kv >= tau);
        const unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (mask) {
          if (pred) {
            const int pos = M + __popc(mask & ((1u << lane) - 1));
            b.key[pos] = kv;
            b.idx[pos] = t;
          }
          M += __popc(mask);
          if (M > kAresCap - 32) {   // no room for another column: keep the k best
            const int had = M;
            M = ares_tighten(b, M, k, lane, tau);
            strict = strict || had >= k;
          }
        }
      }
    }
    if (M >= k && (!strict || M > 2 * k)) {   // exact tau as soon as k candidates exist
      M = ares_tighten(b, M, k, lane, tau);
      strict = true;
    }
  }
  float unused;
  return ares_tighten(b, M, k, lane, unused);
}

// One warp selects the neighbours of one seed and hands (output slot j, neighbour id) pairs to
// `emit` (called by the lane that owns slot j).  copy_path: all `deg` neighbours in CSR order.
// w_idx / w_key: per-warp scratch of k ints (+ k floats for kBias) when k > 32 or MODE == kBias.
// kPos: hand the POSITION inside the row to `emit` instead of the neighbour id (no row loads here).
template <typename IdT, int MODE, typename Emit, bool kPos = false>
__device__ __forceinline__ void warp_select(const IdT *__restrict__ row,
                                            const float *__restrict__ wrow, int deg, int k,
                                            bool copy_path, uint64_t rng_key, uint64_t item,
                                            int lane, int *w_idx, float *w_key, Emit &emit) {
  if (copy_path) {
    // CSR-order copy, 4 independent loads in flight per lane
    int j = lane;
    for (; j + 96 < deg; j += 128) {
      IdT a = pick_value<IdT, kPos>(row, j), b = pick_value<IdT, kPos>(row, j + 32),
          c = pick_value<IdT, kPos>(row, j + 64), d = pick_value<IdT, kPos>(row, j + 96);
      emit(j, a); emit(j + 32, b); emit(j + 64, c); emit(j + 96, d);
    }
    for (; j < deg; j += 32) emit(j, pick_value<IdT, kPos>(row, j));
    return;
  }

  if (MODE == kUniformReplace) {
    for (int j = lane; j < k; j += 32) {
      uint32_t p = rand_below(philox_u32(rng_key, item, (uint32_t)j), (uint32_t)deg);
      emit(j, pick_value<IdT, kPos>(row, p));
    }
  } else if (MODE == kUniform) {
    // Floyd: for t = 0..k-1, J = deg-k+t: r = U[0, J]; pick (r already chosen ? J : r)
    if (k <= 32) {
      uint32_t r_mine = 0;
      if (lane < k)
        r_mine = rand_below(philox_u32(rng_key, item, (uint32_t)lane),
                            (uint32_t)(deg - k + lane + 1));
      uint32_t mine = 0xffffffffu;  // lane t holds the t-th pick
      for (int t = 0; t < k; ++t) {
        const uint32_t r = __shfl_sync(0xffffffffu, r_mine, t);
        const bool dup = __any_sync(0xffffffffu, lane < t && mine == r);
        if (lane == t) mine = dup ? (uint32_t)(deg - k + t) : r;
      }
      if (lane < k) emit(lane, pick_value<IdT, kPos>(row, mine));
    } else {
      for (int t0 = 0; t0 < k; t0 += 32) {
        const int t_mine = t0 + lane;
        uint32_t r_mine = 0;
        if (t_mine < k)
          r_mine = rand_below(philox_u32(rng_key, item, (uint32_t)t_mine),
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
      emit_picks<IdT, Emit, kPos>(row, w_idx, k, lane, emit);
    }
  } else if (MODE == kBias) {
    if (k <= 32) {
      const AresBuf b = ares_buf(threadIdx.x >> 5);
      const int M = ares_collect(wrow, deg, k, 0, 512, rng_key, item, lane, b);   // M == k (deg > k)
      if (lane < M) emit(lane, pick_value<IdT, kPos>(row, b.idx[lane]));
      __syncwarp();
    } else {
      // fill the reservoir with the first k items
      for (int t = lane; t < k; t += 32) {
        const float w = wrow[t];
        w_key[t] = ares_key(philox_u32(rng_key, item, (uint32_t)t), w);
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
      // stream the rest of the row, 512 weights per warp pass: every lane owns 4 quads of 4
      // consecutive elements (quad q = elements t0 + 128 q + 4 lane ..+3 = exactly one Philox block:
      // draw t is component t & 3 of block t >> 2, the same mapping as philox_u32, so the sample
      // does not depend on this blocking).  All 16 weight loads of a lane are issued before any is
      // used - a hub row (deg ~ 10^4) is a chain of ~deg/512 memory round trips instead of deg/32.
      for (int t0 = k & ~3; t0 < deg; t0 += 512) {
        float w4[4][4];
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tb = t0 + 128 * q + 4 * lane;
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int t = tb + c;
            w4[q][c] = (t >= k && t < deg) ? wrow[t] : 0.f;
          }
        }
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tq = t0 + 128 * q;
          if (tq >= deg) break;  // warp-uniform
          const int tb = tq + 4 * lane;
          float key4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (tb < deg) {
            const uint4 r4 = Philox::gen(rng_key, item, (uint64_t)(tb >> 2));
            const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
  #pragma unroll
            for (int c = 0; c < 4; ++c)
              if (tb + c >= k && tb + c < deg) key4[c] = ares_key(rr[c], w4[q][c]);
          }
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float key = key4[c];
            unsigned mask = __ballot_sync(0xffffffffu, key > wmin);
            while (mask) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const float ck = __shfl_sync(0xffffffffu, key, src);
              const int ci = tq + 4 * src + c;
              if (ck > wmin) {  // warp-uniform
                if (lane == 0) { w_key[wslot] = ck; w_idx[wslot] = ci; }
                __syncwarp();
                lmin = INFINITY;
                lslot = -1;
                for (int z = lane; z < k; z += 32) {
                  const float v = w_key[z];
                  if (v < lmin || lslot < 0) { lmin = v; lslot = z; }
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
        }
      }
      __syncwarp();
      emit_picks<IdT, Emit, kPos>(row, w_idx, k, lane, emit);
    }
  } else {  // kBiasReplace
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
      if (live) thr = u32_to_unit(philox_u32(rng_key, item, (uint32_t)j)) * total;
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
      if (live) emit(j, pick_value<IdT, kPos>(row, mypos));
    }
  }
}

}  // namespace dgsb
