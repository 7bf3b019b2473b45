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
// A-Res reservoir for k <= 32 (weighted sampling without replacement): the k largest keys
// log2(u_t) / w_t, one slot per lane in REGISTERS - an insertion is a broadcast + a 5-step arg-min
// butterfly, no shared-memory traffic.  Keys depend only on (rng_key, item, t), so the selected
// SET does not depend on how the row is split over warps; the picks are emitted in descending
// key order (ties: smaller position first), so neither does their order.  That lets a CTA share
// one hub row among its warps (blocks.cu) and still agree with the one-warp-per-seed kernels.
struct Res32 {
  float rkey;   // this lane's slot; the slots are kept SORTED ascending over lanes 0..k-1
                // (lanes >= k hold +inf), so the minimum is always lane 0 and an insertion is
                // one ballot + one shift by a lane instead of an arg-min butterfly
  int ridx;     // position inside the row, -1 = empty
  float wmin;   // key of lane 0
};

__device__ __forceinline__ float ares_key(uint32_t r, float w) {
  return w > 0.f ? __fdividef(__log2f(u32_to_unit(r)), w) : -INFINITY;
}

// empty reservoir (every real key replaces an empty slot)
__device__ __forceinline__ void res32_empty(Res32 &R, int k, int lane) {
  R.rkey = lane < k ? -INFINITY : INFINITY;
  R.ridx = -1;
  R.wmin = -INFINITY;
}

// reservoir holding the first k items of the row (bitonic sort of the 32 lanes)
__device__ __forceinline__ void res32_init(Res32 &R, const float *__restrict__ wrow, int k,
                                           uint64_t rng_key, uint64_t item, int lane) {
  R.rkey = INFINITY;
  R.ridx = 0;
  if (lane < k) {
    R.rkey = ares_key(philox_u32(rng_key, item, (uint32_t)lane), wrow[lane]);
    R.ridx = lane;
  }
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, R.rkey, stride);
      const int oi = __shfl_xor_sync(0xffffffffu, R.ridx, stride);
      const bool take_min = ((lane & stride) == 0) == ((lane & size) == 0);
      const bool other_less = ok < R.rkey || (ok == R.rkey && oi < R.ridx);
      const bool mine_less = R.rkey < ok || (R.rkey == ok && R.ridx < oi);
      if (take_min ? other_less : mine_less) { R.rkey = ok; R.ridx = oi; }
    }
  }
  R.wmin = __shfl_sync(0xffffffffu, R.rkey, 0);
}

// every lane offers one candidate (key, idx); those that beat the minimum enter, in lane order
__device__ __forceinline__ void res32_offer(Res32 &R, float key, int idx, int lane) {
  unsigned mask = __ballot_sync(0xffffffffu, key > R.wmin);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const float ck = __shfl_sync(0xffffffffu, key, src);
    const int ci = __shfl_sync(0xffffffffu, idx, src);
    if (ck > R.wmin) {  // warp-uniform
      // slots below the candidate's place move down one lane (the minimum falls out)
      const int pos = __popc(__ballot_sync(0xffffffffu, R.rkey < ck));   // >= 1
      const float nk = __shfl_down_sync(0xffffffffu, R.rkey, 1);
      const int ni = __shfl_down_sync(0xffffffffu, R.ridx, 1);
      if (lane < pos - 1) { R.rkey = nk; R.ridx = ni; }
      else if (lane == pos - 1) { R.rkey = ck; R.ridx = ci; }
      R.wmin = __shfl_sync(0xffffffffu, R.rkey, 0);
    }
  }
}

// One pass over 512 weights starting at t0 (a multiple of 4; items below k are skipped - they
// are in the initial reservoir): every lane owns 4 quads of 4 consecutive elements (quad q =
// elements t0 + 128 q + 4 lane ..+3 = exactly one Philox block: draw t is component t & 3 of
// block t >> 2, the same mapping as philox_u32).  All 16 weight loads of a lane are issued before
// any is used - a hub row (deg ~ 10^4) is a chain of ~deg/512 memory round trips, not deg/32.
__device__ __forceinline__ void res32_pass(Res32 &R, const float *__restrict__ wrow, int deg, int k,
                                           int t0, uint64_t rng_key, uint64_t item, int lane) {
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
      for (int c = 0; c < 4; ++c) key4[c] = ares_key(rr[c], w4[q][c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) res32_offer(R, key4[c], tb + c, lane);
  }
}

// output slot of this lane's pick: descending key, ties by ascending position
__device__ __forceinline__ int res32_rank(const Res32 &R, int k, int lane) {
  int rank = 0;
  for (int l = 0; l < k; ++l) {
    const float ok = __shfl_sync(0xffffffffu, R.rkey, l);
    const int oi = __shfl_sync(0xffffffffu, R.ridx, l);
    if (ok > R.rkey || (ok == R.rkey && oi < R.ridx)) ++rank;
  }
  return rank;
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
      Res32 R;
      res32_init(R, wrow, k, rng_key, item, lane);
      for (int t0 = k & ~3; t0 < deg; t0 += 512) res32_pass(R, wrow, deg, k, t0, rng_key, item, lane);
      const int j = res32_rank(R, k, lane);
      if (lane < k) emit(j, pick_value<IdT, kPos>(row, R.ridx));
    } else {
      // fill the reservoir with the first k items
      for (int t = lane; t < k; t += 32) {
        const float w = wrow[t];
        const float u = u32_to_unit(philox_u32(rng_key, item, (uint32_t)t));
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
              if (w4[q][c] > 0.f) key4[c] = __log2f(u32_to_unit(rr[c])) / w4[q][c];
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
