// Micro-benchmark (development tool, not part of the product): what bounds the fill pass of
// the tile lists (csrc/assign_tiles.cu, k_tile_lists<fill>)?  1.46e8 list entries claim a
// slot in one of 131072 lists (a counter each) and store a 32-byte record there.  Variants:
//   red        no-return atomic add on the counter only            (the count pass)
//   atom       returning atomic add, result kept alive, no store
//   atom+store returning atomic add, then the 32-byte store into the claimed slot (the fill pass)
//   store      no atomics: the 32-byte store into a slot derived from the entry index
//              (same scatter pattern: list = hash, slot = running index / lists)
//   atom x4    like atom+store with four independent entries in flight per thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fill_probe tools/fill_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { auto e = (x); if (e) { printf("fail %s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ void st32(double2 *p, size_t i, double v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" :: "l"(p + 2 * i), "d"(v) : "memory");
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) k_fill(uint32_t *cnt, double2 *lists, size_t n, uint32_t nlist, uint32_t cap,
    uint32_t *sink) {
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  uint32_t keep = 0;
  for (size_t i0 = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i0 < n; i0 += stride * UNROLL) {
    uint32_t list[UNROLL], pos[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t i = i0 + u * stride;
      list[u] = (uint32_t) (mix(i) % nlist);
      if (i >= n) continue;
      if (MODE == 0) { atomicAdd(cnt + list[u], 1u); pos[u] = 0; }
      else if (MODE == 3) pos[u] = (uint32_t) (i / nlist);
      else pos[u] = atomicAdd(cnt + list[u], 1u);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t i = i0 + u * stride;
      if (i >= n) continue;
      if (MODE == 1) keep += pos[u];
      if (MODE >= 2 && pos[u] < cap) st32(lists, (size_t) list[u] * cap + pos[u], (double) i);
    }
  }
  if (MODE == 1 && keep == 0xffffffffu) *sink = keep;
}

template <int MODE, int UNROLL>
static void run(const char *name, uint32_t *cnt, double2 *lists, size_t n, uint32_t nlist, uint32_t cap, uint32_t *sink) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemset(cnt, 0, nlist * 4));
    CK(cudaEventRecord(a));
    k_fill<MODE, UNROLL><<<148 * 16, 256>>>(cnt, lists, n, nlist, cap, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  printf("%-44s %7.3f ms  %6.1f G entries/s\n", name, best, n / best * 1e-6);
}

__global__ void __launch_bounds__(256) k_store_wrapped(double2 *lists, size_t n, uint32_t nlist, uint32_t w) {
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t list = (uint32_t) (mix(i) % nlist), pos = (uint32_t) (i / nlist) % w;
    st32(lists, (size_t) list * w + pos, (double) i);
  }
}

static void run_wrapped(const char *name, uint32_t *, double2 *lists, size_t n, uint32_t nlist, uint32_t w, uint32_t *) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a));
    k_store_wrapped<<<148 * 16, 256>>>(lists, n, nlist, w);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  printf("%-60s %7.3f ms  %6.1f G entries/s\n", name, best, n / best * 1e-6);
}

int main() {
  const size_t n = 146000000;
  const uint32_t nlist = 131072, cap = 1736;
  uint32_t *cnt, *sink;
  double2 *lists;
  CK(cudaMalloc(&cnt, nlist * 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMalloc(&lists, (size_t) nlist * cap * 32));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("%s; %zu entries into %u lists of %u slots (%.1f GB)\n", p.name, n, nlist, cap, (double) nlist * cap * 32 / 1e9);
  run<0, 1>("red (no return)", cnt, lists, n, nlist, cap, sink);
  run<1, 1>("atom (returning), no store", cnt, lists, n, nlist, cap, sink);
  run<1, 4>("atom (returning), no store, 4 in flight", cnt, lists, n, nlist, cap, sink);
  run<2, 1>("atom + 32-byte store", cnt, lists, n, nlist, cap, sink);
  run<2, 4>("atom + 32-byte store, 4 in flight", cnt, lists, n, nlist, cap, sink);
  run<3, 1>("32-byte store only (no atomics)", cnt, lists, n, nlist, cap, sink);
  run<3, 4>("32-byte store only, 4 in flight", cnt, lists, n, nlist, cap, sink);
  // the same number of stores into smaller footprints: slots wrap inside the first `w` of a
  // list (w * 4 MB in total): is it the L2 request rate, the DRAM write pattern or the TLB?
  for (uint32_t w : {8u, 64u, 512u}) {
    char name[96];
    snprintf(name, sizeof name, "store only, slots wrapped to %u per list (%.2f GB)", w, (double) nlist * w * 32 / 1e9);
    run_wrapped(name, cnt, lists, n, nlist, w, sink);
  }
  // fewer, longer lists (same 7.3 GB): 16384 lists of 13888 slots
  run<3, 1>("store only, 16384 lists x 13888 slots", cnt, lists, n, 16384, 13888, sink);
  run<3, 1>("store only, 1024 lists x 222208 slots", cnt, lists, n, 1024, 222208, sink);
  return 0;
}
