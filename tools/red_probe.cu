// Micro-benchmark (development tool, not part of the product): throughput of
// no-return fp64 / fp32 global reductions (RED.E.ADD) into an L2-resident window,
// in the access patterns of the mass-assignment scatter (csrc/assign.cu):
//   pattern 0: every lane its own random 32-byte sector              (1 element / sector)
//   pattern 1: groups of 3 adjacent lanes on 3 consecutive doubles at a random
//              8-byte aligned offset (the z-coalesced TSC stencil: 1-2 sectors / group)
//   pattern 2: groups of 4 lanes covering one aligned sector           (4 elements / sector)
//   pattern 3: like 1, but the window is 8.6 GB (DRAM resident, random)
// Prints elements/s and an estimate of sector requests/s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { auto e = (x); if (e) { printf("fail %s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

template <typename T, int PATTERN>
__global__ void __launch_bounds__(256) k_red(T *buf, size_t nelem, int iters) {
  const size_t tid = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int per = 32 / (int) sizeof(T);          // elements per sector
  for (int it = 0; it < iters; it++) {
    size_t idx;
    if (PATTERN == 0) idx = (mix(tid * 1315423911ull + it) % (nelem / per)) * per;
    else if (PATTERN == 1 || PATTERN == 3) {
      const int grp = lane / 3, sub = lane % 3;
      if (grp >= 10) continue;
      const size_t g = (tid / 32) * 10 + grp;
      idx = mix(g * 2654435761ull + it) % (nelem - 4) + sub;
    }
    else {
      const size_t g = tid / per;
      idx = (mix(g * 2654435761ull + it) % (nelem / per)) * per + (tid % per);
    }
    atomicAdd(buf + idx, (T) 1);     // result unused: compiles to RED
  }
}

// f32 vector reductions: every lane adds 4 consecutive floats (16-byte aligned) with one
// RED.E.ADD.F32x4; pairs = 1: random 16-byte groups, pairs = 2: a lane issues two
// adjacent groups (a 3-4 cell z stencil at an arbitrary offset)
template <int PAIRS>
__global__ void __launch_bounds__(256) k_red_v4(float *buf, size_t nelem, int iters) {
  const size_t tid = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; it++) {
    const size_t g = mix(tid * 1315423911ull + it) % (nelem / 4 - 2);
    for (int q = 0; q < PAIRS; q++)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(buf + 4 * (g + q)), "f"(1.f),
          "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
}

template <int PAIRS>
static void run_v4(const char *name, size_t bytes) {
  float *buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0, bytes));
  const size_t nelem = bytes / 4;
  const int blocks = 148 * 8, iters = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_red_v4<PAIRS><<<blocks, 256>>>(buf, nelem, 8);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  k_red_v4<PAIRS><<<blocks, 256>>>(buf, nelem, iters);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double) blocks * 256 * iters * PAIRS;
  printf("%-44s %7.3f ms  %8.2f G elem/s  %8.2f G vector-ops/s\n", name, ms, 4 * ops / ms / 1e6, ops / ms / 1e6);
  cudaFree(buf);
}

template <typename T, int PATTERN>
static void run(const char *name, size_t bytes) {
  T *buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0, bytes));
  const size_t nelem = bytes / sizeof(T);
  const int blocks = 148 * 8, iters = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_red<T, PATTERN><<<blocks, 256>>>(buf, nelem, 8);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  k_red<T, PATTERN><<<blocks, 256>>>(buf, nelem, iters);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double lanes = (PATTERN == 1 || PATTERN == 3) ? 30.0 / 32 : 1.0;
  const double elems = (double) blocks * 256 * iters * lanes;
  const double sect = PATTERN == 0 ? elems : (PATTERN == 2 ? elems / (32 / sizeof(T))
      : elems / 3 * (sizeof(T) == 8 ? 1.5 : 1.25));
  printf("%-44s %7.3f ms  %8.2f G elem/s  ~%7.2f G sector-req/s\n", name, ms, elems / ms / 1e6, sect / ms / 1e6);
  cudaFree(buf);
}

int main() {
  const size_t L2WIN = (size_t) 32 << 20, BIG = (size_t) 8 << 30;
  run<double, 0>("f64, 1 lane per random sector, 32 MB", L2WIN);
  run<double, 1>("f64, 3 adjacent lanes (TSC z), 32 MB", L2WIN);
  run<double, 2>("f64, 4 lanes per aligned sector, 32 MB", L2WIN);
  run<double, 3>("f64, 3 adjacent lanes, 8 GB window (DRAM)", BIG);
  run<float, 0>("f32, 1 lane per random sector, 32 MB", L2WIN);
  run<float, 1>("f32, 3 adjacent lanes, 32 MB", L2WIN);
  run<float, 2>("f32, 8 lanes per aligned sector, 32 MB", L2WIN);
  run_v4<1>("f32x4, 1 vector per lane, random, 32 MB", L2WIN);
  run_v4<2>("f32x4, 2 adjacent vectors per lane, 32 MB", L2WIN);
  return 0;
}
