// Hand-written strided complex FFT pass for sm_100a (double and single), N = 256 * R3
// with R3 in {2, 4, 6, 8}: N = 512, 1024, 1536, 2048.
//
// The 3-D r2c transform of the density mesh (the reference calls FFTW,
// src/multipole.c:444,459, src/mp_template.c:98) is three 1-D passes.  The pass
// along z (contiguous rows, real input) stays with cuFFT; the passes along y
// and x are *strided*: consecutive points of one transform are a whole row / a
// whole plane apart.  cuFFT's kernels for them run at ~3.2 TB/s on B200 for 1024
// (50 % of the measured HBM peak), need a 29 GB work area and run at 2.2 TB/s
// for 1536 = 2^9 * 3, and cannot know that columns beyond the last k-bin edge
// are never read.  This kernel does one such pass in place:
//
//   * a tile = TK consecutive k (one TK*16-byte segment per point) x all N
//     points of the strided axis; M = N/16 threads per column;
//   * N = 16 x 16 x R3: two radix-16 passes held in registers (16 complex
//     doubles per thread) and one radix-R3 pass, with two trips through shared
//     memory in between; layouts and column pitch are chosen so that every
//     128-bit access is bank-conflict free per quarter-warp (the first version
//     had a 4-way conflict between columns and was L1-bound at 76 % l1tex
//     throughput); the result goes from registers straight to HBM;
//   * twiddles: each thread derives the 15 powers it needs from one base root
//     (sincospi once per thread) by a depth-4 product tree — no table traffic;
//   * the registers of the 16-point array are free during the last pass, so the
//     next tile's loads are issued before it (one block per SM: nothing else
//     would hide the load latency);
//   * the pass along x can skip tiles whose smallest |k|^2 is already beyond the
//     last bin edge (22 % of the tiles with KMAX at the Nyquist frequency).
//
// Forward transform, sign -1, unnormalised, natural order in and out: the same
// convention as the FFTW / cuFFT calls it replaces.
//
// Single precision: a thread carries TWO columns (k and k + TK) packed in one
// 16-byte element, so the thread / shared-memory / register geometry is the same
// as in double (every shared access stays 128-bit, a[16] stays 64 registers) and a
// tile is 2 TK columns = the same number of bytes per point.  Twiddles are
// computed in double once per thread and rounded.
//
// Index algebra (n = input index, k = output index, M = 16 R3):
//   n = t + M m            (t < M, m < 16)        pass 1: radix 16 over m
//   k = p + 16 q           (p < 16, q < M)
//   X[p + 16 q] = sum_t w_N^{t p} w_M^{t q} sum_m x[t + M m] w_16^{m p}
//   t = t1 + R3 t2, q = q1 + 16 q2  (t1, q2 < R3; t2, q1 < 16)
//   sum_t ... = sum_t1 w_R3^{t1 q2} w_M^{t1 q1} sum_t2 B[t1 + R3 t2][p] w_16^{t2 q1}
//                     pass 3 (radix R3)  twiddle     pass 2 (radix 16)
// Shared memory holds B as [p][t] (row pitch M + 1); pass 2 works in place (slot
// p (M + 1) + t1 + R3 q1 holds C[p][q1][t1]), pass 3 reads R3 consecutive slots.

#include "psb_internal.h"

namespace psb {

namespace {

// launch-shape ablation (context option "fft_variant"): 0 = one wide block per SM,
// 1 = two narrower, independent blocks per SM (Ng = 1024)
int g_fft_variant = 0;

template <typename T> struct Cx { T x, y; };       // a twiddle factor

// the element one thread transforms: one complex double, or two complex floats
// (columns k and k + TK)
template <typename T> struct El;
template <> struct El<double> { double x, y; };
template <> struct El<float> { float x, y, z, w; };
using cd = El<double>;
using cf = El<float>;

__device__ __forceinline__ cd cmul(cd a, Cx<double> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
__device__ __forceinline__ cf cmul(cf a, Cx<float> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x, a.z * b.x - a.w * b.y, a.z * b.y + a.w * b.x};
}
template <typename T> __device__ __forceinline__ Cx<T> wmul(Cx<T> a, Cx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd mul_mi(cd a) { return {a.y, -a.x}; }      // times -i
__device__ __forceinline__ cd scal(cd a, double f) { return {a.x * f, a.y * f}; }
__device__ __forceinline__ cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
__device__ __forceinline__ cf mul_mi(cf a) { return {a.y, -a.x, a.w, -a.z}; }
__device__ __forceinline__ cf scal(cf a, float f) { return {a.x * f, a.y * f, a.z * f, a.w * f}; }

// shared-memory and global-memory access of an element
__device__ __forceinline__ void sm_put(double2 *p, cd v) { *p = make_double2(v.x, v.y); }
__device__ __forceinline__ void sm_put(float4 *p, cf v) { *p = make_float4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ cd sm_get(const double2 *p) { const double2 v = *p; return {v.x, v.y}; }
__device__ __forceinline__ cf sm_get(const float4 *p) { const float4 v = *p; return {v.x, v.y, v.z, v.w}; }
template <typename T> struct Mem;
template <> struct Mem<double> {
  using smem_t = double2;
  using gmem_t = double2;
  static constexpr int COLS = 1;
  static __device__ __forceinline__ cd load(const double2 *g, size_t off, int, bool la, bool) {
    double2 v = make_double2(0.0, 0.0);
    if (la) v = g[off];
    return {v.x, v.y};
  }
  static __device__ __forceinline__ cd from(double2 a, double2) { return {a.x, a.y}; }
  static __device__ __forceinline__ void store(double2 *g, size_t off, int, bool la, bool, cd v) {
    if (la) g[off] = make_double2(v.x, v.y);
  }
};
template <> struct Mem<float> {
  using smem_t = float4;
  using gmem_t = float2;
  static constexpr int COLS = 2;
  static __device__ __forceinline__ cf load(const float2 *g, size_t off, int tk, bool la, bool lb) {
    float2 a = make_float2(0.f, 0.f), b = a;
    if (la) a = g[off];
    if (lb) b = g[off + tk];
    return {a.x, a.y, b.x, b.y};
  }
  static __device__ __forceinline__ cf from(float2 a, float2 b) { return {a.x, a.y, b.x, b.y}; }
  static __device__ __forceinline__ void store(float2 *g, size_t off, int tk, bool la, bool lb, cf v) {
    if (la) g[off] = make_float2(v.x, v.y);
    if (lb) g[off + tk] = make_float2(v.z, v.w);
  }
};

// forward DFTs in place, natural order; E = El<T>
template <typename E> __device__ __forceinline__ void dft2(E &x0, E &x1) {
  const E t = csub(x0, x1);
  x0 = cadd(x0, x1);
  x1 = t;
}

template <typename E> __device__ __forceinline__ void dft4(E &x0, E &x1, E &x2, E &x3) {
  const E s02 = cadd(x0, x2), d02 = csub(x0, x2);
  const E s13 = cadd(x1, x3), d13 = mul_mi(csub(x1, x3));      // -i (x1 - x3)
  x0 = cadd(s02, s13);
  x2 = csub(s02, s13);
  x1 = cadd(d02, d13);
  x3 = csub(d02, d13);
}

template <typename T> __device__ __forceinline__ void dft3(El<T> &x0, El<T> &x1, El<T> &x2) {
  const T S = (T) 0.86602540378443864676;       // sin(pi/3)
  const El<T> t1 = cadd(x1, x2);
  const El<T> t2 = csub(x0, scal(t1, (T) 0.5));
  const El<T> t3 = mul_mi(scal(csub(x1, x2), S));  // -i sin(pi/3) (x1 - x2)
  x0 = cadd(x0, t1);
  x1 = cadd(t2, t3);
  x2 = csub(t2, t3);
}

template <typename T, int R> struct SmallDft;
template <typename T> struct SmallDft<T, 2> {
  static __device__ __forceinline__ void run(El<T> (&d)[2]) { dft2(d[0], d[1]); }
};
template <typename T> struct SmallDft<T, 4> {
  static __device__ __forceinline__ void run(El<T> (&d)[4]) { dft4(d[0], d[1], d[2], d[3]); }
};
template <typename T> struct SmallDft<T, 6> {
  // X[k] = E[k mod 3] + w6^k O[k mod 3], E / O = 3-point DFTs of the even / odd inputs
  static __device__ __forceinline__ void run(El<T> (&d)[6]) {
    const T S = (T) 0.86602540378443864676;
    El<T> e0 = d[0], e1 = d[2], e2 = d[4], o0 = d[1], o1 = d[3], o2 = d[5];
    dft3<T>(e0, e1, e2);
    dft3<T>(o0, o1, o2);
    const El<T> o1w = cmul(o1, Cx<T>{(T) 0.5, -S});       // w6^1
    const El<T> o2w = cmul(o2, Cx<T>{(T) -0.5, -S});      // w6^2
    d[0] = cadd(e0, o0); d[3] = csub(e0, o0);              // w6^3 = -1
    d[1] = cadd(e1, o1w); d[4] = csub(e1, o1w);            // w6^4 = -w6^1
    d[2] = cadd(e2, o2w); d[5] = csub(e2, o2w);            // w6^5 = -w6^2
  }
};
template <typename T> struct SmallDft<T, 8> {
  static __device__ __forceinline__ void run(El<T> (&d)[8]) {
    const T H = (T) 0.70710678118654752440;
    El<T> e0 = d[0], e1 = d[2], e2 = d[4], e3 = d[6], o0 = d[1], o1 = d[3], o2 = d[5], o3 = d[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const El<T> o1w = cmul(o1, Cx<T>{H, -H});             // w8^1
    const El<T> o2w = mul_mi(o2);                          // w8^2 = -i
    const El<T> o3w = cmul(o3, Cx<T>{-H, -H});            // w8^3
    d[0] = cadd(e0, o0); d[4] = csub(e0, o0);
    d[1] = cadd(e1, o1w); d[5] = csub(e1, o1w);
    d[2] = cadd(e2, o2w); d[6] = csub(e2, o2w);
    d[3] = cadd(e3, o3w); d[7] = csub(e3, o3w);
  }
};

// forward 16-point DFT: a[0..15] -> natural-order result in a[]
template <typename T> __device__ __forceinline__ void dft16(El<T> (&a)[16]) {
  const T C1 = (T) 0.92387953251128673848, S1 = (T) 0.38268343236508977173;   // cos, sin(pi/8)
  const T H = (T) 0.70710678118654752440;
  using W = Cx<T>;
  // stage 1: four 4-point DFTs over n2 (n = j + 4 n2); result r lands in a[j + 4 r]
#pragma unroll
  for (int j = 0; j < 4; j++) dft4(a[j], a[j + 4], a[j + 8], a[j + 12]);
  // twiddles w16^(j r), j, r = 1..3
  a[5] = cmul(a[5], W{C1, -S1});        // j=1 r=1: w^1
  a[9] = cmul(a[9], W{H, -H});          // j=1 r=2: w^2
  a[13] = cmul(a[13], W{S1, -C1});      // j=1 r=3: w^3
  a[6] = cmul(a[6], W{H, -H});          // j=2 r=1: w^2
  a[10] = mul_mi(a[10]);                // j=2 r=2: w^4 = -i
  a[14] = cmul(a[14], W{-H, -H});       // j=2 r=3: w^6
  a[7] = cmul(a[7], W{S1, -C1});        // j=3 r=1: w^3
  a[11] = cmul(a[11], W{-H, -H});       // j=3 r=2: w^6
  a[15] = cmul(a[15], W{-C1, S1});      // j=3 r=3: w^9
  // stage 2: for every r a 4-point DFT over j; result s is X[r + 4 s]
#pragma unroll
  for (int r = 0; r < 4; r++) dft4(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3]);
  // a[4 r + s] holds X[r + 4 s]: transpose the 4 x 4 index
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int s = r + 1; s < 4; s++) {
      const El<T> t = a[4 * r + s];
      a[4 * r + s] = a[4 * s + r];
      a[4 * s + r] = t;
    }
}

// a[p] *= w^p for p = 1..15, powers by a depth-4 product tree.  w1 is the correctly
// rounded double root and the squarings w2, w4, w8 always run in double: squaring
// doubles the phase error each time, and a float tree from a float root doubled the
// error of a 512^3 survey spectrum relative to cuFFT's float transform.  The
// remaining products (at most three deep, from correctly rounded factors) run in T.
template <typename T> __device__ __forceinline__ Cx<T> wcast(Cx<double> w) { return {(T) w.x, (T) w.y}; }
template <typename T> __device__ __forceinline__ void twiddle_powers(El<T> (&a)[16], Cx<double> wd) {
  const Cx<double> wd2 = wmul(wd, wd), wd4 = wmul(wd2, wd2), wd8 = wmul(wd4, wd4);
  const Cx<T> w1 = wcast<T>(wd), w2 = wcast<T>(wd2), w4 = wcast<T>(wd4), w8 = wcast<T>(wd8);
  const Cx<T> w3 = wmul(w2, w1), w5 = wmul(w4, w1), w6 = wmul(w4, w2), w7 = wmul(w4, w3);
  a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3); a[4] = cmul(a[4], w4);
  a[5] = cmul(a[5], w5); a[6] = cmul(a[6], w6); a[7] = cmul(a[7], w7); a[8] = cmul(a[8], w8);
  a[9] = cmul(a[9], wmul(w8, w1)); a[10] = cmul(a[10], wmul(w8, w2));
  a[11] = cmul(a[11], wmul(w8, w3)); a[12] = cmul(a[12], wmul(w8, w4));
  a[13] = cmul(a[13], wmul(w8, w5)); a[14] = cmul(a[14], wmul(w8, w6));
  a[15] = cmul(a[15], wmul(w8, w7));
}

// 16-byte elements per column region: 16 rows of M + 1 for the [p][t] layout of
// pass 1 (rounded up to a multiple of 8), plus a shift.  A 128-bit shared-memory
// access is served per quarter-warp: 8 consecutive lanes = min(TK, 8) columns x
// 8/TK threads of one column, and its 8 lanes must fall into 8 distinct 16-byte
// bank groups.  Within a column consecutive threads are one element apart, so
// columns are offset by 8/TK elements (1 element when TK >= 8), modulo 8.
template <int R3, int TK> struct Shape {
  static constexpr int N = 256 * R3;
  static constexpr int M = 16 * R3;             // threads per column
  static constexpr int ROW = M + 1;             // pitch of the [p][t] layout
  static constexpr int BASE = ((16 * ROW + 7) / 8) * 8;
  static constexpr int PITCH = BASE + (TK >= 8 ? 1 : 8 / TK);
  static constexpr int THREADS = M * TK;
};

// data:    base of the complex array
// outer_n: number of values of the non-transformed slow index
// outer_stride / estride: element strides of that index / of the transformed axis
// k2a / k2b / k2max: optional skip test — a tile (o, k0) is skipped when
//   k2a[o] + k2b[k0] >= k2max (smallest |k|^2 of the tile beyond the last bin edge)
template <typename T, int R3, int TK, int MINB>
__global__ void __launch_bounds__(Shape<R3, TK>::THREADS, MINB)
k_fft_strided(typename Mem<T>::gmem_t *__restrict__ data, int ngk, int outer_n, size_t outer_stride,
    size_t estride, const double *__restrict__ k2a, const double *__restrict__ k2b, double k2max,
    FftOut out, FftStoreSkip ss) {
  using S = Shape<R3, TK>;
  using E = El<T>;
  using MM = Mem<T>;
  using G = typename MM::gmem_t;
  constexpr int M = S::M, ROW = S::ROW;
  constexpr int WIDTH = TK * MM::COLS;                          // columns per tile
  extern __shared__ double2 sm_raw[];
  typename MM::smem_t *sm = reinterpret_cast<typename MM::smem_t *>(sm_raw);
  const int c = threadIdx.x % TK, u = threadIdx.x / TK;       // column in tile, thread in column
  typename MM::smem_t *col = sm + (size_t) c * S::PITCH;
  // roles: pass 1: t = u;  pass 2: (p2, t1) = (u % 16, u / 16)
  const int p2 = u & 15, t1 = u >> 4;
  Cx<double> w_t, w_t1;
  {
    double s, co;
    sincospi(-2.0 * u / (double) S::N, &s, &co);       // w_N^t
    w_t = {co, s};
    sincospi(-2.0 * t1 / (double) M, &s, &co);         // w_M^t1
    w_t1 = {co, s};
  }
  const int ktiles = (ngk + WIDTH - 1) / WIDTH;
  const long ntile = (long) outer_n * ktiles;
  // next tile of this block that is not skipped (uniform per block), or ntile
  auto next_tile = [&](long t) {
    for (; t < ntile; t += gridDim.x) {
      if (!k2a) break;
      const int o = (int) (t / ktiles), k0 = (int) (t % ktiles) * WIDTH;
      if (k2a[o] + k2b[k0] < k2max) break;
    }
    return t;
  };
  auto tile_ptr = [&](long t, bool &la, bool &lb) {
    const int o = (int) (t / ktiles), k0 = (int) (t % ktiles) * WIDTH;
    la = (k0 + c) < ngk;
    lb = (k0 + c + TK) < ngk;
    return data + (size_t) o * outer_stride + k0 + c;
  };
  // offset of this thread's column inside a block of the transposed output layout
  auto out_off = [&](long t) {
    const int o = (int) (t / ktiles), k0 = (int) (t % ktiles) * WIDTH;
    return (size_t) o * out.outer_stride + k0 + c;
  };
  E a[16];
  long tile = next_tile(blockIdx.x);
  bool la = false, lb = false;
  G *g = nullptr;
  if (tile < ntile) {
    g = tile_ptr(tile, la, lb);
#pragma unroll
    for (int m = 0; m < 16; m++) a[m] = MM::load(g, (size_t) (u + M * m) * estride, TK, la, lb);
  }
  while (tile < ntile) {
    // ---- pass 1: radix 16 over m
    dft16<T>(a);
    twiddle_powers<T>(a, w_t);
#pragma unroll
    for (int p = 0; p < 16; p++) sm_put(&col[p * ROW + u], a[p]);
    __syncthreads();
    // ---- pass 2: radix 16 over t2 (t = t1 + R3 t2) for fixed (p, t1)
#pragma unroll
    for (int t2 = 0; t2 < 16; t2++) a[t2] = sm_get(&col[p2 * ROW + t1 + R3 * t2]);
    dft16<T>(a);
    twiddle_powers<T>(a, w_t1);
    // in place: output q1 of thread (p, t1) takes the slot of its input t2 = q1, so no
    // barrier is needed between this pass's loads and stores
#pragma unroll
    for (int q1 = 0; q1 < 16; q1++) sm_put(&col[p2 * ROW + t1 + R3 * q1], a[q1]);
    __syncthreads();
    // ---- prefetch the next tile into the (now free) registers of a[]
    const long ntl = next_tile(tile + gridDim.x);
    bool nla = false, nlb = false;
    G *ng_ = nullptr;
    if (ntl < ntile) {
      ng_ = tile_ptr(ntl, nla, nlb);
#pragma unroll
      for (int m = 0; m < 16; m++) a[m] = MM::load(ng_, (size_t) (u + M * m) * estride, TK, nla, nlb);
    }
    // outputs beyond the last bin edge are not stored (FftStoreSkip): per tile, the part of
    // the test that does not depend on the output index
    double ska = 0.0, skb = 0.0, sko = 0.0;
    if (ss.k2t) {
      const int o = (int) (tile / ktiles), k0 = (int) (tile % ktiles) * WIDTH;
      sko = ss.k2o ? __ldg(ss.k2o + o) : 0.0;
      ska = __ldg(ss.k2k + (ss.per_column ? min(k0 + c, ngk - 1) : k0));
      skb = __ldg(ss.k2k + (ss.per_column ? min(k0 + c + TK, ngk - 1) : k0));
    }
    // ---- pass 3: radix R3 over t1 for the 256 (p, q1) pairs; X[p + 16 q1 + 256 q2]
    for (int pair = u; pair < 256; pair += M) {
      const int p3 = pair & 15, q1 = pair >> 4;
      E d[R3];
#pragma unroll
      for (int tt = 0; tt < R3; tt++) d[tt] = sm_get(&col[p3 * ROW + tt + R3 * q1]);
      SmallDft<T, R3>::run(d);
      if (out.ny == 0) {
#pragma unroll
        for (int q2 = 0; q2 < R3; q2++) {
          const int n = p3 + 16 * q1 + 256 * q2;
          bool ka = la, kb = lb;
          if (ss.k2t) {
            const double kto = __ldg(ss.k2t + n) + sko;
            ka = ka && (kto + ska < ss.k2max);
            kb = kb && (kto + skb < ss.k2max);
          }
          MM::store(g, (size_t) n * estride, TK, ka, kb, d[q2]);
        }
      }
      else {
        // transposed output for the slab decomposition: point y of the transform goes to
        // block y / ny (one per destination rank; the block may be peer memory), row y % ny
        const size_t oo = out_off(tile);
#pragma unroll
        for (int q2 = 0; q2 < R3; q2++) {
          const int y = p3 + 16 * q1 + 256 * q2, blk = y / out.ny;
          bool ka = la, kb = lb;
          if (ss.k2t) {
            const double kto = __ldg(ss.k2t + y) + sko;
            ka = ka && (kto + ska < ss.k2max);
            kb = kb && (kto + skb < ss.k2max);
          }
          MM::store(static_cast<G *>(out.base[blk]) + oo, (size_t) (y - blk * out.ny) * ngk, TK, ka, kb,
              d[q2]);
        }
      }
    }
    __syncthreads();
    tile = ntl; g = ng_; la = nla; lb = nlb;
  }
}

template <typename T, int R3, int TK, int MINB>
int launch_shape(void *data, int ng, int ngk, int axis, int outer_n, const double *k2a,
    const double *k2b, double k2max, const FftOut &out, const FftStoreSkip &ss, cudaStream_t st) {
  using S = Shape<R3, TK>;
  using G = typename Mem<T>::gmem_t;
  const size_t smem = (size_t) TK * S::PITCH * 16;
  auto kern = k_fft_strided<T, R3, TK, MINB>;
  PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int width = TK * Mem<T>::COLS;
  const long ntile = (long) outer_n * ((ngk + width - 1) / width);
  const int grid = (int) std::min<long>((long) sms * MINB, ntile);
  if (grid <= 0) return 0;
  const size_t row = (size_t) ngk;
  // axis 1: the array is (outer_n, ng, ngk); points one row apart, outer index = the
  //         plane of ng rows.
  // axis 0: the array is (ng, outer_n, ngk); points one plane of outer_n rows apart,
  //         outer index = the row inside the plane.
  if (axis == 1)
    kern<<<grid, S::THREADS, smem, st>>>(static_cast<G *>(data), ngk, outer_n,
        (size_t) ng * ngk, row, nullptr, nullptr, 0.0, out, ss);
  else
    kern<<<grid, S::THREADS, smem, st>>>(static_cast<G *>(data), ngk, outer_n, row,
        (size_t) outer_n * ngk, k2a, k2b, k2max, FftOut{}, ss);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// z pass: real-to-complex transform of contiguous rows (the first of the three
// 1-D passes of the r2c plan), same three in-register / shared-memory passes.
// One complex transform serves TWO real rows a, b packed as z = a + i b:
//   A[k] = (Z[k] + conj Z[N-k]) / 2,   B[k] = -i (Z[k] - conj Z[N-k]) / 2,
// un-mixed with one more read of shared memory (the last pass works in place there).
// Lanes run along the row (u fastest) so that global loads and stores are
// contiguous; a tile = TK elements = 2 TK rows (double) or 4 TK rows (float: an
// element carries two complex numbers = four rows).
// (Ablation: weighting the input by Y_lm(r_hat) on load, to save the weighted copy
// of src/mp_template.c:65-95, doubled the instruction count of this latency-bound
// kernel and was slower than the separate bandwidth-bound weighting pass.)
// ---------------------------------------------------------------------------
template <typename T> struct RowIO;
template <> struct RowIO<double> {
  static constexpr int ROWS = 2;
  static __device__ __forceinline__ cd load(const double *const (&r)[2], const bool (&lv)[2], int n) {
    cd v = {0.0, 0.0};
    if (lv[0]) v.x = r[0][n];
    if (lv[1]) v.y = r[1][n];
    return v;
  }
  // outputs of the rows from Z[k] and Z[N-k]
  static __device__ __forceinline__ void store(double2 *const (&o)[2], const bool (&lv)[2], int k, cd zk, cd zn) {
    if (lv[0]) o[0][k] = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    if (lv[1]) o[1][k] = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
  }
};
template <> struct RowIO<float> {
  static constexpr int ROWS = 4;
  static __device__ __forceinline__ cf load(const float *const (&r)[4], const bool (&lv)[4], int n) {
    cf v = {0.f, 0.f, 0.f, 0.f};
    if (lv[0]) v.x = r[0][n];
    if (lv[1]) v.y = r[1][n];
    if (lv[2]) v.z = r[2][n];
    if (lv[3]) v.w = r[3][n];
    return v;
  }
  static __device__ __forceinline__ void store(float2 *const (&o)[4], const bool (&lv)[4], int k, cf zk, cf zn) {
    if (lv[0]) o[0][k] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    if (lv[1]) o[1][k] = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
    if (lv[2]) o[2][k] = make_float2(0.5f * (zk.z + zn.z), 0.5f * (zk.w - zn.w));
    if (lv[3]) o[3][k] = make_float2(0.5f * (zk.w + zn.w), -0.5f * (zk.z - zn.z));
  }
};

// src: real rows, src_pitch reals apart; dst: complex rows, dst_pitch complex apart
// (in place: dst == src, src_pitch == 2 dst_pitch).  nrows rows in total.
template <typename T, int R3, int TK, int MINB>
__global__ void __launch_bounds__(Shape<R3, TK>::THREADS, MINB)
k_fft_rows(const T *__restrict__ src, typename Mem<T>::gmem_t *__restrict__ dst, long nrows,
    size_t src_pitch, size_t dst_pitch) {
  using S = Shape<R3, TK>;
  using E = El<T>;
  using IO = RowIO<T>;
  using G = typename Mem<T>::gmem_t;
  constexpr int M = S::M, ROW = S::ROW, N = S::N, RW = IO::ROWS;
  constexpr int RPT = TK * RW;                                 // rows per tile
  extern __shared__ double2 sm_raw[];
  typename Mem<T>::smem_t *sm = reinterpret_cast<typename Mem<T>::smem_t *>(sm_raw);
  const int u = threadIdx.x % M, c = threadIdx.x / M;         // lanes run along the row
  typename Mem<T>::smem_t *col = sm + (size_t) c * S::PITCH;
  const int p2 = u & 15, t1 = u >> 4;
  Cx<double> w_t, w_t1;
  {
    double s, co;
    sincospi(-2.0 * u / (double) N, &s, &co);
    w_t = {co, s};
    sincospi(-2.0 * t1 / (double) M, &s, &co);
    w_t1 = {co, s};
  }
  const long ntile = (nrows + RPT - 1) / RPT;
  E a[16];
  // loads the 16 points of this thread of tile t into a[]
  auto fetch = [&](long t) {
    const long r0 = t * RPT + (long) c * RW;
    const T *rp[RW];
    bool lv[RW];
#pragma unroll
    for (int j = 0; j < RW; j++) {
      lv[j] = r0 + j < nrows;
      rp[j] = src + (size_t) (lv[j] ? r0 + j : 0) * src_pitch + u;
    }
#pragma unroll
    for (int m = 0; m < 16; m++) a[m] = IO::load(rp, lv, M * m);
  };
  long tile = blockIdx.x;
  if (tile < ntile) fetch(tile);
  while (tile < ntile) {
    G *op[RW];
    bool olv[RW];
#pragma unroll
    for (int j = 0; j < RW; j++) {
      const long r = tile * RPT + (long) c * RW + j;
      olv[j] = r < nrows;
      op[j] = dst + (size_t) (olv[j] ? r : 0) * dst_pitch;
    }
    // ---- pass 1: radix 16 over m
    dft16<T>(a);
    twiddle_powers<T>(a, w_t);
#pragma unroll
    for (int p = 0; p < 16; p++) sm_put(&col[p * ROW + u], a[p]);
    __syncthreads();
    // ---- pass 2: radix 16 over t2
#pragma unroll
    for (int t2 = 0; t2 < 16; t2++) a[t2] = sm_get(&col[p2 * ROW + t1 + R3 * t2]);
    dft16<T>(a);
    twiddle_powers<T>(a, w_t1);
#pragma unroll
    for (int q1 = 0; q1 < 16; q1++) sm_put(&col[p2 * ROW + t1 + R3 * q1], a[q1]);
    __syncthreads();
    // ---- next tile's loads in flight during the last pass and the un-mixing
    const long ntl = tile + gridDim.x;
    if (ntl < ntile) fetch(ntl);
    // ---- pass 3: radix R3, in place in shared memory (a thread reads and writes
    // the same R3 slots): Z[p + 16 q1 + 256 q2] lands at p (M + 1) + R3 q1 + q2
    for (int pair = u; pair < 256; pair += M) {
      const int p3 = pair & 15, q1 = pair >> 4;
      E d[R3];
#pragma unroll
      for (int tt = 0; tt < R3; tt++) d[tt] = sm_get(&col[p3 * ROW + tt + R3 * q1]);
      SmallDft<T, R3>::run(d);
#pragma unroll
      for (int q2 = 0; q2 < R3; q2++) sm_put(&col[p3 * ROW + q2 + R3 * q1], d[q2]);
    }
    __syncthreads();
    // ---- un-mix the rows and store k = 0 .. N/2
    auto slot = [](int k) { return (k & 15) * ROW + (k >> 8) + R3 * ((k >> 4) & 15); };
#pragma unroll
    for (int it = 0; it <= N / 2 / M; it++) {
      const int k = u + M * it;
      if (k <= N / 2) {
        const E zk = sm_get(&col[slot(k)]), zn = sm_get(&col[slot(k ? N - k : 0)]);
        IO::store(op, olv, k, zk, zn);
      }
    }
    __syncthreads();
    tile = ntl;
  }
}

template <typename T, int R3, int TK, int MINB>
int launch_rows_shape(const void *src, void *dst, long nrows, size_t src_pitch, size_t dst_pitch,
    cudaStream_t st) {
  using S = Shape<R3, TK>;
  using G = typename Mem<T>::gmem_t;
  const size_t smem = (size_t) TK * S::PITCH * 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long ntile = (nrows + TK * RowIO<T>::ROWS - 1) / (TK * RowIO<T>::ROWS);
  const int grid = (int) std::min<long>((long) sms * MINB, ntile);
  if (grid <= 0) return 0;
  auto kern = k_fft_rows<T, R3, TK, MINB>;
  PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  kern<<<grid, S::THREADS, smem, st>>>(static_cast<const T *>(src), static_cast<G *>(dst), nrows,
      src_pitch, dst_pitch);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// z + y passes fused through the L2: ONE persistent kernel (one block per SM)
// walks a static list of work items
//     z(0) .. z(L-1),  [ z(p+L), y(p) ]  for p = 0 .. ng-1
// where z(p) = the r2c row tiles of plane p and y(p) = its strided column tiles.
// y(p) starts when a per-plane counter says all row tiles of the plane are done;
// the plane (8.4 MB at 1024^3) is then still in the 126 MB L2, so the y pass reads
// what the z pass wrote without going to HBM, and its in-place result overwrites
// lines that are still dirty: the two passes cost one HBM read and one write.
// (Doing the same with separate launches per plane group was a loss: 512 small
// launches cost more than the L2 hits save.)
//
// Items are assigned round-robin (item = block + i * grid), every block works
// through its items in increasing order and all blocks are resident, so the
// smallest unfinished item can always run: its dependencies have smaller indices.
// The next item's loads are issued before the current item's stores only when
// that keeps this property (its dependencies precede the current item).
// Both item types use the lane mapping of the strided pass (TK fastest), so the
// twiddles and the shared-memory layout are common.  Data written by other SMs
// inside the kernel is read with ld.global.cg (L1 is not coherent).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 ldcg2(const double2 *p) { return __ldcg(p); }
__device__ __forceinline__ float2 ldcg2(const float2 *p) { return __ldcg(p); }

template <typename T, int R3, int TK, int MINB>
__global__ void __launch_bounds__(Shape<R3, TK>::THREADS, MINB)
k_fft_zy(T *__restrict__ mesh, int ng, int ngk, int nplanes, int lag, int *__restrict__ done) {
  using S = Shape<R3, TK>;
  using E = El<T>;
  using MM = Mem<T>;
  using IO = RowIO<T>;
  using G = typename MM::gmem_t;
  constexpr int M = S::M, ROW = S::ROW, N = S::N, RW = IO::ROWS;
  constexpr int RPT = TK * RW;                                  // rows per z item
  constexpr int WIDTH = TK * MM::COLS;                          // columns per y item
  extern __shared__ double2 sm_raw[];
  typename MM::smem_t *sm = reinterpret_cast<typename MM::smem_t *>(sm_raw);
  const int c = threadIdx.x % TK, u = threadIdx.x / TK;
  typename MM::smem_t *col = sm + (size_t) c * S::PITCH;
  const int p2 = u & 15, t1 = u >> 4;
  Cx<double> w_t, w_t1;
  {
    double s, co;
    sincospi(-2.0 * u / (double) N, &s, &co);
    w_t = {co, s};
    sincospi(-2.0 * t1 / (double) M, &s, &co);
    w_t1 = {co, s};
  }
  const int nz = (ng + RPT - 1) / RPT, ny = (ngk + WIDTH - 1) / WIDTH, grp = nz + ny;
  const long nitems = (long) lag * nz + (long) nplanes * grp;
  const size_t plane_c = (size_t) ng * ngk;                     // complex elements per plane
  // item -> (kind, plane, tile); kind 0 = z, 1 = y, -1 = nothing (z beyond the last plane)
  auto decode = [&](long it, int &kind, int &pl, int &tl) {
    if (it < (long) lag * nz) { kind = 0; pl = (int) (it / nz); tl = (int) (it % nz); return; }
    const long j = it - (long) lag * nz;
    const int gq = (int) (j / grp), r = (int) (j % grp);
    if (r < nz) { pl = gq + lag; tl = r; kind = pl < nplanes ? 0 : -1; }
    else { kind = 1; pl = gq; tl = r - nz; }
  };
  // largest item index a y item of plane pl depends on
  auto dep_max = [&](int pl) {
    return pl < lag ? (long) lag * nz - 1 : (long) lag * nz + (long) (pl - lag) * grp + nz - 1;
  };
  auto next_item = [&](long it) {
    for (it += gridDim.x; it < nitems; it += gridDim.x) {
      int k_, p_, t_;
      decode(it, k_, p_, t_);
      if (k_ >= 0) break;
    }
    return it;
  };
  auto wait_plane = [&](int pl) {
    if (threadIdx.x == 0) {
      while (*reinterpret_cast<volatile int *>(done + pl) < nz) __nanosleep(200);
      __threadfence();
    }
    __syncthreads();
  };
  E a[16];
  auto fetch = [&](int kind, int pl, int tl) {
    if (kind == 0) {
      const int r0 = tl * RPT + c * RW;
      const T *rp[RW];
      bool lv[RW];
#pragma unroll
      for (int j = 0; j < RW; j++) {
        lv[j] = r0 + j < ng;
        rp[j] = mesh + 2 * (plane_c * pl + (size_t) (lv[j] ? r0 + j : 0) * ngk) + u;
      }
#pragma unroll
      for (int m = 0; m < 16; m++) a[m] = IO::load(rp, lv, M * m);
    }
    else {
      const int k0 = tl * WIDTH;
      const bool la = (k0 + c) < ngk, lb = (k0 + c + TK) < ngk;
      const G *g = reinterpret_cast<const G *>(mesh) + plane_c * pl + k0 + c;
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const size_t off = (size_t) (u + M * m) * ngk;
        if (MM::COLS == 1) {
          G v = {};
          if (la) v = ldcg2(g + off);
          a[m] = Mem<T>::from(v, v);
        }
        else {
          G va = {}, vb = {};
          if (la) va = ldcg2(g + off);
          if (lb) vb = ldcg2(g + off + TK);
          a[m] = Mem<T>::from(va, vb);
        }
      }
    }
  };
  long cur = blockIdx.x;
  int kind, pl, tl;
  if (cur < nitems) {
    decode(cur, kind, pl, tl);
    if (kind < 0) { cur = next_item(cur); if (cur < nitems) decode(cur, kind, pl, tl); }
  }
  if (cur < nitems) {
    if (kind == 1) wait_plane(pl);
    fetch(kind, pl, tl);
  }
  while (cur < nitems) {
    // ---- passes 1 and 2, common to both kinds
    dft16<T>(a);
    twiddle_powers<T>(a, w_t);
#pragma unroll
    for (int p = 0; p < 16; p++) sm_put(&col[p * ROW + u], a[p]);
    __syncthreads();
#pragma unroll
    for (int t2 = 0; t2 < 16; t2++) a[t2] = sm_get(&col[p2 * ROW + t1 + R3 * t2]);
    dft16<T>(a);
    twiddle_powers<T>(a, w_t1);
#pragma unroll
    for (int q1 = 0; q1 < 16; q1++) sm_put(&col[p2 * ROW + t1 + R3 * q1], a[q1]);
    __syncthreads();
    // ---- the next item: prefetch now if its dependencies precede this item
    const long nxt = next_item(cur);
    int nkind = -1, npl = 0, ntl = 0;
    bool fetched = false;
    if (nxt < nitems) {
      decode(nxt, nkind, npl, ntl);
      if (nkind == 0 || dep_max(npl) < cur) {
        if (nkind == 1) wait_plane(npl);
        fetch(nkind, npl, ntl);
        fetched = true;
      }
    }
    if (kind == 1) {
      // ---- y item: pass 3 straight to global memory
      const int k0 = tl * WIDTH;
      const bool la = (k0 + c) < ngk, lb = (k0 + c + TK) < ngk;
      G *g = reinterpret_cast<G *>(mesh) + plane_c * pl + k0 + c;
      for (int pair = u; pair < 256; pair += M) {
        const int p3 = pair & 15, q1 = pair >> 4;
        E d[R3];
#pragma unroll
        for (int tt = 0; tt < R3; tt++) d[tt] = sm_get(&col[p3 * ROW + tt + R3 * q1]);
        SmallDft<T, R3>::run(d);
#pragma unroll
        for (int q2 = 0; q2 < R3; q2++)
          MM::store(g, (size_t) (p3 + 16 * q1 + 256 * q2) * ngk, TK, la, lb, d[q2]);
      }
      __syncthreads();
    }
    else {
      // ---- z item: pass 3 in place in shared memory, then un-mix the packed rows
      for (int pair = u; pair < 256; pair += M) {
        const int p3 = pair & 15, q1 = pair >> 4;
        E d[R3];
#pragma unroll
        for (int tt = 0; tt < R3; tt++) d[tt] = sm_get(&col[p3 * ROW + tt + R3 * q1]);
        SmallDft<T, R3>::run(d);
#pragma unroll
        for (int q2 = 0; q2 < R3; q2++) sm_put(&col[p3 * ROW + q2 + R3 * q1], d[q2]);
      }
      __syncthreads();
      G *op[RW];
      bool olv[RW];
#pragma unroll
      for (int j = 0; j < RW; j++) {
        const int r = tl * RPT + c * RW + j;
        olv[j] = r < ng;
        op[j] = reinterpret_cast<G *>(mesh) + plane_c * pl + (size_t) (olv[j] ? r : 0) * ngk;
      }
      auto slot = [](int k) { return (k & 15) * ROW + (k >> 8) + R3 * ((k >> 4) & 15); };
#pragma unroll
      for (int it = 0; it <= N / 2 / M; it++) {
        const int k = u + M * it;
        if (k <= N / 2) {
          const E zk = sm_get(&col[slot(k)]), zn = sm_get(&col[slot(k ? N - k : 0)]);
          IO::store(op, olv, k, zk, zn);
        }
      }
      // publish: the rows of this tile are written (barrier, then one cumulative
      // fence by the thread that bumps the plane's counter — the grid-sync idiom)
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(done + pl, 1); }
    }
    if (nxt < nitems && !fetched) {
      if (nkind == 1) wait_plane(npl);
      fetch(nkind, npl, ntl);
    }
    cur = nxt; kind = nkind; pl = npl; tl = ntl;
  }
}

template <typename T, int R3, int TK, int MINB>
int launch_zy_shape(void *mesh, int ng, int ngk, int nplanes, int *done, cudaStream_t st) {
  using S = Shape<R3, TK>;
  const size_t smem = (size_t) TK * S::PITCH * 16;
  auto kern = k_fft_zy<T, R3, TK, MINB>;
  PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  int dev = 0, sms = 148, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, S::THREADS, smem));
  if (occ < 1) { set_error("fused z + y FFT kernel does not fit an SM\n"); return -1; }
  // every block must be resident (the items wait for each other)
  const int grid = sms * std::min(occ, MINB);
  const int nz = (ng + TK * RowIO<T>::ROWS - 1) / (TK * RowIO<T>::ROWS);
  const int ny = (ngk + TK * Mem<T>::COLS - 1) / (TK * Mem<T>::COLS);
  int lag = (grid + nz + ny) / (nz + ny) + 1;       // lag * (nz + ny) > grid
  if (lag > nplanes) lag = nplanes;
  PSB_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * nplanes, st));
  // cooperative launch: fails instead of deadlocking if the blocks cannot all be resident
  T *m = static_cast<T *>(mesh);
  void *args[] = {&m, &ng, &ngk, &nplanes, &lag, &done};
  PSB_CUDA(cudaLaunchCooperativeKernel((const void *) kern, dim3(grid), dim3(S::THREADS), args, smem, st));
  return 0;
}

template <typename T>
int launch_zy_any(void *mesh, int ng, int ngk, int nplanes, int *done, cudaStream_t st) {
  switch (ng) {
    case 512: return launch_zy_shape<T, 2, 16, 1>(mesh, ng, ngk, nplanes, done, st);
    case 1024:
      if (g_fft_variant == 1) return launch_zy_shape<T, 4, 4, 2>(mesh, ng, ngk, nplanes, done, st);
      return launch_zy_shape<T, 4, 8, 1>(mesh, ng, ngk, nplanes, done, st);
    case 1536: return launch_zy_shape<T, 6, 4, 1>(mesh, ng, ngk, nplanes, done, st);
    case 2048: return launch_zy_shape<T, 8, 4, 1>(mesh, ng, ngk, nplanes, done, st);
    default:
      set_error("no hand-written fused FFT for GRID_SIZE %d\n", ng);
      return -1;
  }
}

template <typename T>
int launch_rows_any(const void *src, void *dst, int ng, long nrows, size_t src_pitch,
    size_t dst_pitch, cudaStream_t st) {
  switch (ng) {
    case 512: return launch_rows_shape<T, 2, 16, 1>(src, dst, nrows, src_pitch, dst_pitch, st);
    case 1024:
      if (g_fft_variant == 1) return launch_rows_shape<T, 4, 4, 2>(src, dst, nrows, src_pitch, dst_pitch, st);
      return launch_rows_shape<T, 4, 8, 1>(src, dst, nrows, src_pitch, dst_pitch, st);
    case 1536: return launch_rows_shape<T, 6, 4, 1>(src, dst, nrows, src_pitch, dst_pitch, st);
    case 2048: return launch_rows_shape<T, 8, 4, 1>(src, dst, nrows, src_pitch, dst_pitch, st);
    default:
      set_error("no hand-written r2c FFT for GRID_SIZE %d\n", ng);
      return -1;
  }
}

template <typename T>
int launch_any(void *data, int ng, int ngk, int axis, int outer_n, const double *k2a,
    const double *k2b, double k2max, const FftOut &out, const FftStoreSkip &ss, cudaStream_t st) {
  switch (ng) {
    case 512: return launch_shape<T, 2, 16, 1>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, out, ss, st);
    case 1024:
      if (g_fft_variant == 1) return launch_shape<T, 4, 4, 2>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, out, ss, st);
      return launch_shape<T, 4, 8, 1>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, out, ss, st);
    case 1536: return launch_shape<T, 6, 4, 1>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, out, ss, st);
    case 2048: return launch_shape<T, 8, 4, 1>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, out, ss, st);
    default:
      set_error("no hand-written strided FFT for GRID_SIZE %d\n", ng);
      return -1;
  }
}

}  // namespace

void fft_set_variant(int v) { g_fft_variant = v; }

bool fft_strided_supported(int ng, int precision) {
  return (precision == 8 || precision == 4) && (ng == 512 || ng == 1024 || ng == 1536 || ng == 2048);
}

// In-place forward ng-point transform along a strided axis of a complex array
// (double for precision 8, float for 4) with rows of ngk elements.
// axis 1: along y of (outer_n, ng, ngk); axis 0: along x of (ng, outer_n, ngk)
// (outer_n = ng on a single GPU, the y-slab height in the slab-decomposed path).
int launch_fft_strided(void *data, int precision, int ng, int ngk, int axis, int outer_n,
    const double *k2a, const double *k2b, double k2max, cudaStream_t st, const FftStoreSkip *ss) {
  const FftStoreSkip none, &s = ss ? *ss : none;
  if (precision == 8)
    return launch_any<double>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, FftOut{}, s, st);
  return launch_any<float>(data, ng, ngk, axis, outer_n, k2a, k2b, k2max, FftOut{}, s, st);
}

// The y pass of the slab-decomposed transform with the transpose's packing (and, when
// the blocks are peer memory, the transpose itself) fused into its store: reads
// (outer_n, ng, ngk) from `data`, writes point y of plane o to
// out.base[y / out.ny] + (o * out.outer_stride + (y % out.ny) * ngk + k).
int launch_fft_strided_out(const void *data, int precision, int ng, int ngk, int outer_n,
    const FftOut &out, cudaStream_t st, const FftStoreSkip *ss) {
  const FftStoreSkip none, &s = ss ? *ss : none;
  if (out.ny <= 0 || ng % out.ny || ng / out.ny > FftOut::MAXB) {
    set_error("invalid transposed-output layout for the y pass\n");
    return -1;
  }
  void *d = const_cast<void *>(data);
  if (precision == 8)
    return launch_any<double>(d, ng, ngk, 1, outer_n, nullptr, nullptr, 0.0, out, s, st);
  return launch_any<float>(d, ng, ngk, 1, outer_n, nullptr, nullptr, 0.0, out, s, st);
}

// z and y passes of nplanes planes of a (nplanes, ng, 2 ngk) real mesh in place,
// fused through the L2 (persistent kernel).  done: nplanes ints of device scratch.
int launch_fft_zy(void *mesh, int precision, int ng, int ngk, int nplanes, int *done,
    cudaStream_t st) {
  if (precision == 8) return launch_zy_any<double>(mesh, ng, ngk, nplanes, done, st);
  return launch_zy_any<float>(mesh, ng, ngk, nplanes, done, st);
}

// Real-to-complex forward transform of nrows contiguous rows of ng reals (the z
// pass).  src rows are src_pitch reals apart, dst rows dst_pitch complex apart; in
// place when dst == src and src_pitch == 2 dst_pitch.
int launch_fft_rows(const void *src, void *dst, int precision, int ng, long nrows,
    size_t src_pitch, size_t dst_pitch, cudaStream_t st) {
  if (precision == 8)
    return launch_rows_any<double>(src, dst, ng, nrows, src_pitch, dst_pitch, st);
  return launch_rows_any<float>(src, dst, ng, nrows, src_pitch, dst_pitch, st);
}

}  // namespace psb
