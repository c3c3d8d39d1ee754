// Micro-benchmark (development tool, not part of the product): throughput of
// fp64 / fp32 atomic adds into a shared-memory mesh tile, in the access patterns a
// tile-accumulating mass-assignment kernel would have: a tile of
// 16 x 16 x 32 cells plus stencil reach = 20 x 20 x 36 reals, two blocks per SM,
// ~763 particles per tile and field (config 2), TSC stencil (27 points).
//   mode 0: one particle per lane — 27 sequential atomics per lane, the 32 lanes of
//           an instruction hit 32 unrelated cells (bank conflicts, rare CAS clashes)
//   mode 1: one particle per warp, one stencil point per lane (27 of 32 lanes active,
//           distinct cells: 3 contiguous in z x 3 rows x 3 planes)
//   mode 2 / 3: the same two patterns with plain (non-atomic) load-add-store; T = unsigned
//           uses the native ATOMS.ADD (fp64, fp32 and u64 adds are ATOMS.CAST.SPIN loops)
// Prints elements/s over the whole GPU and the time 5.4e9 updates (config 2) would take.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/smem_atomic_probe tools/smem_atomic_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { auto e = (x); if (e) { printf("fail %s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int EX = 20, EY = 20, EZ = 36, E = EX * EY * EZ;

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256, 2) k_tile(double *out, int ppb, unsigned seed) {
  extern __shared__ unsigned char raw[];
  T *sm = reinterpret_cast<T *>(raw);
  for (int i = threadIdx.x; i < E; i += blockDim.x) sm[i] = (T) 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (MODE == 0) {
    for (int p = threadIdx.x; p < ppb; p += blockDim.x) {
      const unsigned h = mix(seed + blockIdx.x * 1000003u + (unsigned) p);
      const int bx = 1 + (int) (h % 16u), by = 1 + (int) ((h >> 8) % 16u), bz = 1 + (int) ((h >> 16) % 32u);
      const T w = (T) (1.0 + (double) (h & 7u));
#pragma unroll
      for (int a = -1; a <= 1; a++)
#pragma unroll
        for (int b = -1; b <= 1; b++)
#pragma unroll
          for (int c = -1; c <= 1; c++)
            atomicAdd(&sm[((bx + a) * EY + (by + b)) * EZ + bz + c], w);
    }
  }
  else if (MODE == 2) {
    // plain read-modify-write, one particle per lane: NOT race-free (lanes and warps may
    // collide) — the rate a conflict-free colouring of the same accesses could reach
    for (int p = threadIdx.x; p < ppb; p += blockDim.x) {
      const unsigned h = mix(seed + blockIdx.x * 1000003u + (unsigned) p);
      const int bx = 1 + (int) (h % 16u), by = 1 + (int) ((h >> 8) % 16u), bz = 1 + (int) ((h >> 16) % 32u);
      const T w = (T) (1.0 + (double) (h & 7u));
#pragma unroll
      for (int a = -1; a <= 1; a++)
#pragma unroll
        for (int b = -1; b <= 1; b++)
#pragma unroll
          for (int c = -1; c <= 1; c++) {
            volatile T *q = &sm[((bx + a) * EY + (by + b)) * EZ + bz + c];
            *q = *q + w;
          }
    }
  }
  else if (MODE == 3) {
    // plain read-modify-write, one particle per warp and lane = stencil point (distinct
    // cells inside an instruction; warps of a block may still collide: rate probe only)
    const int a = lane / 9 - 1, b = (lane / 3) % 3 - 1, c = lane % 3 - 1;
    for (int p = warp; p < ppb; p += nw) {
      const unsigned h = mix(seed + blockIdx.x * 1000003u + (unsigned) p);
      const int bx = 1 + (int) (h % 16u), by = 1 + (int) ((h >> 8) % 16u), bz = 1 + (int) ((h >> 16) % 32u);
      const T w = (T) (1.0 + (double) (h & 7u));
      if (lane < 27) {
        volatile T *q = &sm[((bx + a) * EY + (by + b)) * EZ + bz + c];
        *q = *q + w;
      }
    }
  }
  else {
    const int a = lane / 9 - 1, b = (lane / 3) % 3 - 1, c = lane % 3 - 1;
    for (int p = warp; p < ppb; p += nw) {
      const unsigned h = mix(seed + blockIdx.x * 1000003u + (unsigned) p);
      const int bx = 1 + (int) (h % 16u), by = 1 + (int) ((h >> 8) % 16u), bz = 1 + (int) ((h >> 16) % 32u);
      const T w = (T) (1.0 + (double) (h & 7u));
      if (lane < 27) atomicAdd(&sm[((bx + a) * EY + (by + b)) * EZ + bz + c], w);
    }
  }
  __syncthreads();
  // checksum: ONE global atomic per block (one per thread — 1.2 million adds to a single
  // address — serialised at the L2 for ~1.9 ms and was what the first version of this
  // probe measured, whatever the shared-memory pattern)
  double s = 0.0;
  for (int i = threadIdx.x; i < E; i += blockDim.x) s += (double) sm[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if (lane == 0) part[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w2 = 0; w2 < nw; w2++) t += part[w2];
    atomicAdd(out, t);
  }
}

template <typename T, int MODE> void run(const char *what, int sms) {
  auto kern = k_tile<T, MODE>;
  const size_t smem = (size_t) E * sizeof(T);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  const int ppb = 763 * 8, blocks = sms * 2 * 16;
  double *out;
  CK(cudaMalloc(&out, 8));
  CK(cudaMemset(out, 0, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<blocks, 256, smem>>>(out, ppb, 1u);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    kern<<<blocks, 256, smem>>>(out, ppb, 7u + rep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  double sum; CK(cudaMemcpy(&sum, out, 8, cudaMemcpyDeviceToHost));
  const double elems = (double) blocks * ppb * 27.0;
  const double rate = elems / (best * 1e-3);
  printf("%-46s blocks/SM %d  %8.3f ms  %7.1f G elements/s  -> 5.4e9 updates in %6.2f ms  (checksum %.6g)\n",
      what, per_sm, best, rate * 1e-9, 5.4e9 / rate * 1e3, sum);
  CK(cudaFree(out));
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  printf("%s, %d SMs; tile %d x %d x %d, 256 threads per block, 763*8 particles per block and launch\n",
      pr.name, pr.multiProcessorCount, EX, EY, EZ);
  const int sms = pr.multiProcessorCount;
  run<double, 0>("f64, one particle per lane (27 atomics each)", sms);
  run<double, 1>("f64, one particle per warp (lane = stencil pt)", sms);
  run<float, 0>("f32, one particle per lane", sms);
  run<float, 1>("f32, one particle per warp", sms);
  run<unsigned, 0>("u32 native ATOMS.ADD, one particle per lane", sms);
  run<unsigned, 1>("u32 native ATOMS.ADD, one particle per warp", sms);
  run<double, 2>("f64 PLAIN rmw (racy), one particle per lane", sms);
  run<double, 3>("f64 PLAIN rmw, lane = stencil point", sms);
  run<float, 2>("f32 PLAIN rmw (racy), one particle per lane", sms);
  return 0;
}
