// Round-trip probe (B200): how long does one DEPENDENT global access take inside a grid of
// 296 x 256 threads, for the access kinds the fused batch kernel chains together?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rt_probe rt_probe.cu && ./rt_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// chain of `steps` dependent loads through a random permutation (footprint n entries of 4 B)
__global__ void chain_kernel(const unsigned int *perm, unsigned int n, int steps, int per_thread,
                             unsigned int *sink, unsigned long long *times, int mode) {
  const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int idx[8];
  for (int u = 0; u < per_thread; ++u) idx[u] = (gtid * 8u + u * 2654435761u) % n;
  __syncthreads();
  unsigned long long t0 = gtime();
  for (int s = 0; s < steps; ++s) {
    for (int u = 0; u < per_thread; ++u) idx[u] = mode == 0 ? __ldcg(perm + idx[u]) : __ldg(perm + idx[u]);
    if (mode == 2) __syncthreads();
  }
  unsigned long long t1 = gtime();
  unsigned int acc = 0;
  for (int u = 0; u < per_thread; ++u) acc += idx[u];
  sink[gtid] = acc;
  if (threadIdx.x == 0) times[blockIdx.x] = t1 - t0;
}

// write by one set of CTAs, grid barrier, read by others: latency of reading freshly written lines
__global__ void handoff_kernel(unsigned int *buf, unsigned int n, unsigned long long *times, unsigned int *sink) {
  cg::grid_group grid = cg::this_grid();
  const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int total = gridDim.x * blockDim.x;
  for (int round = 0; round < 4; ++round) {
    buf[(gtid + round) % n] = gtid + round;
    unsigned long long ta = gtime();
    grid.sync();
    unsigned long long tb = gtime();
    unsigned int v = __ldcg(buf + (gtid + total / 2 + round) % n);   // written by a far-away CTA
    unsigned int w = __ldcg(buf + (v * 7u + 13u) % total % n);       // dependent
    unsigned long long tc = gtime();
    sink[gtid] = w;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      times[round * 2] = tb - ta;
      times[round * 2 + 1] = tc - tb;
    }
    grid.sync();
  }
}

__global__ void atomic_kernel(unsigned long long *table, unsigned long long mask, int per_thread,
                              unsigned long long *times, int kind) {
  const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long t0 = gtime();
  unsigned long long acc = 0;
  for (int u = 0; u < per_thread; ++u) {
    unsigned long long h = (gtid * 8ull + u) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    unsigned long long pos = (h & mask) * 2;
    if (kind == 0) acc += atomicCAS(table + pos, ~0ull, h);
    else if (kind == 1) atomicMin((unsigned int *)(table + pos + 1), (unsigned int)gtid);
    else { acc += atomicCAS(table + pos, ~0ull, h); atomicMin((unsigned int *)(table + pos + 1), (unsigned int)gtid); }
  }
  if (acc == 12345) table[0] = acc;
  __threadfence();
  unsigned long long t1 = gtime();
  if (threadIdx.x == 0) times[blockIdx.x] = t1 - t0;
}

static double med(std::vector<unsigned long long> v) {
  std::sort(v.begin(), v.end());
  return v[v.size() / 2] * 1e-3;
}
#include <algorithm>

int main() {
  const int grid = 296, block = 256;
  unsigned long long *times;
  unsigned int *sink;
  cudaMalloc(&times, 4096 * 8);
  cudaMalloc(&sink, grid * block * 4);
  std::vector<unsigned long long> h(grid);
  for (unsigned int mb : {1u, 16u, 64u, 512u}) {
    unsigned int n = mb * 1024u * 1024u / 4;
    std::vector<unsigned int> p(n);
    for (unsigned int i = 0; i < n; ++i) p[i] = (unsigned int)(((unsigned long long)i * 2654435761ull + 12345) % n);
    unsigned int *d;
    cudaMalloc(&d, (size_t)n * 4);
    cudaMemcpy(d, p.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    for (int per : {1, 4, 8})
      for (int mode : {0, 2}) {
        for (int rep = 0; rep < 3; ++rep) chain_kernel<<<grid, block>>>(d, n, 8, per, sink, times, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), times, grid * 8, cudaMemcpyDeviceToHost);
        printf("chain footprint %4u MB  loads/thread/step %d  %s : %.2f us per dependent step (median CTA)\n", mb, per,
               mode == 0 ? "ld.cg        " : "ld.cg + bar  ", med(h) / 8);
      }
    cudaFree(d);
  }
  {
    unsigned int n = grid * block;
    unsigned int *buf;
    cudaMalloc(&buf, n * 4);
    void *args[] = {&buf, &n, &times, &sink};
    for (int rep = 0; rep < 3; ++rep) cudaLaunchCooperativeKernel((void *)handoff_kernel, dim3(grid), dim3(block), args, 0, 0);
    cudaDeviceSynchronize();
    unsigned long long t[8];
    cudaMemcpy(t, times, sizeof(t), cudaMemcpyDeviceToHost);
    for (int r = 0; r < 4; ++r) printf("handoff round %d: grid.sync %.2f us, 2 dependent reads of fresh lines %.2f us\n", r, t[2 * r] * 1e-3, t[2 * r + 1] * 1e-3);
  }
  for (unsigned int mb : {8u, 64u}) {
    unsigned long long slots = (unsigned long long)mb * 1024 * 1024 / 16;
    unsigned long long *tab;
    cudaMalloc(&tab, slots * 16);
    for (int kind : {0, 1, 2})
      for (int per : {1, 4, 8}) {
        cudaMemset(tab, 0xff, slots * 16);
        cudaDeviceSynchronize();
        atomic_kernel<<<grid, block>>>(tab, slots - 1, per, times, kind);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), times, grid * 8, cudaMemcpyDeviceToHost);
        printf("atomics table %3u MB  %d per thread  %s : %.2f us until fenced (median CTA), %.1f G ops/s\n", mb, per,
               kind == 0 ? "CAS64      " : kind == 1 ? "RED.min32  " : "CAS64+min32", med(h),
               (double)grid * block * per * (kind == 2 ? 2 : 1) / (med(h) * 1e-6) * 1e-9);
      }
    cudaFree(tab);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
