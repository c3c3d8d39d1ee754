// Hand-written strided 1024-point complex FFT pass for sm_100a (double).
//
// The 3-D r2c transform of the density mesh (the reference calls FFTW,
// src/multipole.c:444,459) is three 1-D passes.  The pass along z (contiguous
// rows, real input) stays with cuFFT; the passes along y and x are *strided*:
// consecutive points of one transform are a whole row / a whole plane apart.
// cuFFT's kernels for them run at ~3.2 TB/s on B200 (50 % of the measured HBM
// peak) and cannot know that columns beyond the last k-bin edge are never read.
// This kernel does one such pass in place:
//
//   * a tile = 4 consecutive k (one 64-byte segment per point) x all 1024 points
//     of the strided axis; 256 threads = 4 columns x 64 threads;
//   * 1024 = 16 x 16 x 4: two radix-16 passes held in registers (16 complex
//     doubles per thread) and one radix-4 pass, with two trips through shared
//     memory in between (layouts and column pitch chosen so that every 128-bit
//     access is bank-conflict free per quarter-warp — the first version had a
//     4-way conflict between the columns and was L1-bound at 76 % l1tex
//     throughput); the result goes from registers straight to HBM;
//   * twiddles: each thread derives the 15 powers it needs from one base root
//     (sincospi once per thread) by a depth-4 product tree;
//   * the pass along x can skip tiles whose smallest |k|^2 is already beyond the
//     last bin edge (21 % of the columns with KMAX at the Nyquist frequency).
//
// Forward transform, sign -1, unnormalised, natural order in and out: the same
// convention as the FFTW / cuFFT calls it replaces.

#include "psb_internal.h"

namespace psb {

namespace {

struct cd { double x, y; };

__device__ __forceinline__ cd cmul(cd a, cd b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
// multiply by -i and by +i
__device__ __forceinline__ cd mul_mi(cd a) { return {a.y, -a.x}; }

// forward 4-point DFT in place: (x0..x3) -> (X0..X3)
__device__ __forceinline__ void dft4(cd &x0, cd &x1, cd &x2, cd &x3) {
  const cd s02 = cadd(x0, x2), d02 = csub(x0, x2);
  const cd s13 = cadd(x1, x3), d13 = mul_mi(csub(x1, x3));     // -i (x1 - x3)
  x0 = cadd(s02, s13);
  x2 = csub(s02, s13);
  x1 = cadd(d02, d13);
  x3 = csub(d02, d13);
}

// forward 16-point DFT: a[0..15] -> natural-order result in a[]
__device__ __forceinline__ void dft16(cd (&a)[16]) {
  const double C1 = 0.92387953251128673848, S1 = 0.38268343236508977173;   // cos, sin(pi/8)
  const double H = 0.70710678118654752440;
  // stage 1: four 4-point DFTs over n2 (n = j + 4 n2); result r lands in a[j + 4 r]
#pragma unroll
  for (int j = 0; j < 4; j++) dft4(a[j], a[j + 4], a[j + 8], a[j + 12]);
  // twiddles w16^(j r), j, r = 1..3
  a[5] = cmul(a[5], cd{C1, -S1});       // j=1 r=1: w^1
  a[9] = cmul(a[9], cd{H, -H});         // j=1 r=2: w^2
  a[13] = cmul(a[13], cd{S1, -C1});     // j=1 r=3: w^3
  a[6] = cmul(a[6], cd{H, -H});         // j=2 r=1: w^2
  a[10] = mul_mi(a[10]);                // j=2 r=2: w^4 = -i
  a[14] = cmul(a[14], cd{-H, -H});      // j=2 r=3: w^6
  a[7] = cmul(a[7], cd{S1, -C1});       // j=3 r=1: w^3
  a[11] = cmul(a[11], cd{-H, -H});      // j=3 r=2: w^6
  a[15] = cmul(a[15], cd{-C1, S1});     // j=3 r=3: w^9
  // stage 2: for every r a 4-point DFT over j; result s is X[r + 4 s]
#pragma unroll
  for (int r = 0; r < 4; r++) dft4(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3]);
  // a[4 r + s] holds X[r + 4 s]: transpose the 4 x 4 index
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = r + 1; s < 4; s++) {
      const cd t = a[4 * r + s];
      a[4 * r + s] = a[4 * s + r];
      a[4 * s + r] = t;
    }
}

// a[p] *= w^p for p = 1..15, powers by a depth-4 product tree
__device__ __forceinline__ void twiddle_powers(cd (&a)[16], cd w1) {
  const cd w2 = cmul(w1, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4);
  const cd w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3); a[4] = cmul(a[4], w4);
  a[5] = cmul(a[5], w5); a[6] = cmul(a[6], w6); a[7] = cmul(a[7], w7); a[8] = cmul(a[8], w8);
  a[9] = cmul(a[9], cmul(w8, w1)); a[10] = cmul(a[10], cmul(w8, w2));
  a[11] = cmul(a[11], cmul(w8, w3)); a[12] = cmul(a[12], cmul(w8, w4));
  a[13] = cmul(a[13], cmul(w8, w5)); a[14] = cmul(a[14], cmul(w8, w6));
  a[15] = cmul(a[15], cmul(w8, w7));
}

constexpr int FN = 1024;        // transform length
// Complex elements per column region: 16 x 65 for the [p][t] layout, plus a shift
// of 8/TK elements.  A 128-bit shared-memory access is served per quarter-warp
// (8 consecutive lanes = TK columns x 8/TK threads of one column), and its 8 lanes
// must fall into 8 distinct 16-byte bank groups: within a column consecutive
// threads are one element apart, so columns are offset by 8/TK elements.
template <int TK> struct Pitch { static constexpr int value = 1040 + 8 / TK; };

// data:    base of the (Ng, Ng, ngk) complex array
// outer_n: number of values of the non-transformed slow index
// outer_stride / estride: element strides of that index / of the transformed axis
// k2a / k2b / k2max: optional skip test — a tile (o, k0) is skipped when
//   k2a[o] + k2b[k0] >= k2max (smallest |k|^2 of the tile beyond the last bin edge)
// TK = columns (consecutive k) per tile = 64-byte x TK/4 segments; 64 threads per column
template <int TK, int MINB>
__global__ void __launch_bounds__(64 * TK, MINB) k_fft1024_strided(double2 *__restrict__ data, int ngk,
    int outer_n, size_t outer_stride, size_t estride, const double *__restrict__ k2a,
    const double *__restrict__ k2b, double k2max) {
  extern __shared__ double2 sm[];
  const int c = threadIdx.x % TK, u = threadIdx.x / TK;       // column in tile, thread in column
  double2 *col = sm + (size_t) c * Pitch<TK>::value;
  // base roots of unity of this thread's roles
  cd w_t, w_t1;
  {
    double s, co;
    sincospi(-2.0 * u / 1024.0, &s, &co);     // w1024^u           (pass 1: t = u)
    w_t = {co, s};
    sincospi(-2.0 * (u >> 4) / 64.0, &s, &co);        // w64^t1    (pass 2: t1 = u / 16)
    w_t1 = {co, s};
  }
  const int ktiles = (ngk + TK - 1) / TK;
  const long ntile = (long) outer_n * ktiles;
  // next tile of this block that is not skipped (uniform per block), or ntile
  auto next_tile = [&](long t) {
    for (; t < ntile; t += gridDim.x) {
      if (!k2a) break;
      const int o = (int) (t / ktiles), k0 = (int) (t % ktiles) * TK;
      if (k2a[o] + k2b[k0] < k2max) break;
    }
    return t;
  };
  auto tile_ptr = [&](long t, bool &live) {
    const int o = (int) (t / ktiles), k0 = (int) (t % ktiles) * TK;
    live = (k0 + c) < ngk;
    return data + (size_t) o * outer_stride + k0 + c;
  };
  cd a[16];
  long tile = next_tile(blockIdx.x);
  bool live = false;
  double2 *g = nullptr;
  if (tile < ntile) {
    g = tile_ptr(tile, live);
    // ---- loads of pass 1: radix 16 over m, n = u + 64 m
#pragma unroll
    for (int m = 0; m < 16; m++) {
      double2 v = make_double2(0.0, 0.0);
      if (live) v = g[(size_t) (u + 64 * m) * estride];
      a[m] = {v.x, v.y};
    }
  }
  while (tile < ntile) {
    // ---- pass 1
    dft16(a);
    twiddle_powers(a, w_t);
#pragma unroll
    for (int p = 0; p < 16; p++) col[p * 65 + u] = make_double2(a[p].x, a[p].y);
    __syncthreads();
    // ---- pass 2: radix 16 over t2, t = t1 + 4 t2, for fixed (p, t1)
    const int p2 = u & 15, t1 = u >> 4;
#pragma unroll
    for (int t2 = 0; t2 < 16; t2++) {
      const double2 v = col[p2 * 65 + t1 + 4 * t2];
      a[t2] = {v.x, v.y};
    }
    __syncthreads();
    dft16(a);
    twiddle_powers(a, w_t1);
#pragma unroll
    for (int q1 = 0; q1 < 16; q1++) col[(q1 * 4 + t1) * 16 + p2] = make_double2(a[q1].x, a[q1].y);
    __syncthreads();
    // ---- prefetch: the registers of a[] are free during pass 3, so the next
    // tile's loads are put in flight now and overlap pass 3, its stores and the
    // barrier (one block per SM: nothing else would hide the load latency)
    const long ntl = next_tile(tile + gridDim.x);
    bool nlive = false;
    double2 *ng_ = nullptr;
    if (ntl < ntile) {
      ng_ = tile_ptr(ntl, nlive);
#pragma unroll
      for (int m = 0; m < 16; m++) {
        double2 v = make_double2(0.0, 0.0);
        if (nlive) v = ng_[(size_t) (u + 64 * m) * estride];
        a[m] = {v.x, v.y};
      }
    }
    // ---- pass 3: radix 4 over t1 for fixed (p, q1); X[p + 16 q1 + 256 q2]
    const int p3 = u & 15, gq = u >> 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int q1 = 4 * gq + i;
      cd d[4];
#pragma unroll
      for (int tt = 0; tt < 4; tt++) {
        const double2 v = col[(q1 * 4 + tt) * 16 + p3];
        d[tt] = {v.x, v.y};
      }
      dft4(d[0], d[1], d[2], d[3]);
      if (live) {
#pragma unroll
        for (int q2 = 0; q2 < 4; q2++)
          g[(size_t) (p3 + 16 * q1 + 256 * q2) * estride] = make_double2(d[q2].x, d[q2].y);
      }
    }
    __syncthreads();
    tile = ntl; g = ng_; live = nlive;
  }
}

}  // namespace

// In-place forward 1024-point transform along a strided axis of a
// (1024, 1024, ngk) complex double array.  axis 1: along y; axis 0: along x.
template <int TK, int MINB>
static int launch_variant(void *data, int ngk, int axis, const double *k2a, const double *k2b,
    double k2max, cudaStream_t st) {
  const size_t plane = (size_t) FN * ngk;
  const size_t smem = (size_t) TK * Pitch<TK>::value * sizeof(double2);
  auto kern = k_fft1024_strided<TK, MINB>;
  PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * MINB;
  if (axis == 1)        // along y: outer = x (stride one plane), points one row apart
    kern<<<grid, 64 * TK, smem, st>>>(static_cast<double2 *>(data), ngk, FN, plane, (size_t) ngk,
        nullptr, nullptr, 0.0);
  else                  // along x: outer = y (stride one row), points one plane apart
    kern<<<grid, 64 * TK, smem, st>>>(static_cast<double2 *>(data), ngk, FN, (size_t) ngk, plane,
        k2a, k2b, k2max);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_fft1024_strided(void *data, int ngk, int axis, const double *k2a, const double *k2b,
    double k2max, int variant, cudaStream_t st) {
  switch (variant) {
    case 1: return launch_variant<4, 3>(data, ngk, axis, k2a, k2b, k2max, st);
    case 3: return launch_variant<2, 4>(data, ngk, axis, k2a, k2b, k2max, st);
    case 4: return launch_variant<2, 6>(data, ngk, axis, k2a, k2b, k2max, st);
    case 0: return launch_variant<4, 2>(data, ngk, axis, k2a, k2b, k2max, st);
    default: return launch_variant<8, 1>(data, ngk, axis, k2a, k2b, k2max, st);
  }
}

}  // namespace psb
