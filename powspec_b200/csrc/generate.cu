// Device-side synthetic catalogues for benchmarks and large-size tests
// (SURVEY.md §8d): counter-based Philox4x32-10, counter = particle index, so a
// catalogue is reproducible for a given seed whatever the launch shape, and the
// 10^8..10^10-particle inputs of the BASELINE configs never cross PCIe.
//   kind 0: uniform in [0, L)^3, w = 1
//   kind 1: clustered — N/1000 uniform centres, isotropic Gaussian offsets with
//           sigma = 2 (length units), periodic wrap; 20 % uniform background.
// The reference ships no generator (its inputs are ASCII catalogues).

#include "psb_internal.h"

namespace psb {

namespace {

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = {hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0};
    k0 += W0; k1 += W1;
  }
  return ctr;
}

// 53-bit uniform in [0, 1)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  const uint64_t m = ((uint64_t) (a >> 5) << 26) | (uint64_t) (b >> 6);
  return (double) m * 0x1p-53;
}

__device__ __forceinline__ double wrap_box(double x, double L) {
  if (x >= L) x -= L;
  if (x < 0) x += L;
  if (x >= L || x < 0) x = 0;   // value that rounds onto the face (quirk Q8)
  return x;
}

__global__ void __launch_bounds__(256) k_generate(double2 *__restrict__ out, size_t n,
    double L, int kind, uint32_t k0, uint32_t k1, size_t first) {
  for (size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x; j < n;
       j += (size_t) gridDim.x * blockDim.x) {
    const size_t i = first + j;         // global particle index = Philox counter
    const uint32_t ilo = (uint32_t) i, ihi = (uint32_t) (i >> 32);
    const U4 r0 = philox4x32_10({ilo, ihi, 0u, 0u}, k0, k1);
    const U4 r1 = philox4x32_10({ilo, ihi, 1u, 0u}, k0, k1);
    double x = u53(r0.x, r0.y) * L, y = u53(r0.z, r0.w) * L, z = u53(r1.x, r1.y) * L;
    if (kind == 1 && u53(r1.z, r1.w) >= 0.2) {
      const size_t c = i / 1000;
      const uint32_t clo = (uint32_t) c, chi = (uint32_t) (c >> 32);
      const U4 c0 = philox4x32_10({clo, chi, 2u, 0u}, k0, k1);
      const U4 c1 = philox4x32_10({clo, chi, 3u, 0u}, k0, k1);
      const U4 g0 = philox4x32_10({ilo, ihi, 4u, 0u}, k0, k1);
      const U4 g1 = philox4x32_10({ilo, ihi, 5u, 0u}, k0, k1);
      // Box-Muller: two pairs give four normals, three are used
      const double ra = sqrt(-2.0 * log(1.0 - u53(g0.x, g0.y)));
      const double rb = sqrt(-2.0 * log(1.0 - u53(g1.x, g1.y)));
      double sa, ca, sb, cb;
      sincospi(2.0 * u53(g0.z, g0.w), &sa, &ca);
      sincospi(2.0 * u53(g1.z, g1.w), &sb, &cb);
      const double sigma = 2.0;
      x = u53(c0.x, c0.y) * L + sigma * ra * ca;
      y = u53(c0.z, c0.w) * L + sigma * ra * sa;
      z = u53(c1.x, c1.y) * L + sigma * rb * cb;
      (void) sb;
    }
    out[2 * j] = make_double2(wrap_box(x, L), wrap_box(y, L));
    out[2 * j + 1] = make_double2(wrap_box(z, L), 1.0);
  }
}

}  // namespace

int launch_generate_at(double *out, size_t n, double boxsize, int kind, uint64_t seed,
    uint64_t first_index, cudaStream_t st) {
  if (!n) return 0;
  size_t b = (n + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  k_generate<<<(int) b, 256, 0, st>>>(reinterpret_cast<double2 *>(out), n, boxsize, kind,
      (uint32_t) seed, (uint32_t) (seed >> 32), (size_t) first_index);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_generate(double *out, size_t n, double boxsize, int kind, uint64_t seed,
    cudaStream_t st) {
  return launch_generate_at(out, n, boxsize, kind, seed, 0, st);
}

}  // namespace psb
