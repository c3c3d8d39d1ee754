// Fourier-space stage for sm_100a: mode counting, the fused
// interlace-combine + window correction + mu / L_ell(mu) + per-bin reduction
// kernel, and the real-spherical-harmonic passes of the survey estimator.
//
// Replaces the reference's OpenMP loops (paths relative to cheng-zhao/powspec):
//   powspec_precomp          src/multipole.c:111-257   (k_spectrum<GEOM>)
//   interlace combine        src/multipole.c:462-487   (fused, or k_combine)
//   count_mode_sim_lin/_log  src/multipole.c:778-898, 911-1031 (k_spectrum<SIM>)
//   count_mode_lin/_log      src/multipole.c:579-660, 673-754  (k_spectrum<SURVEY>)
//   dens_k<ell>              src/mp_template.c:48-149  (k_ylm_weight_r, k_ylm_accum_k)
//   legpoly / YlmR_l<ell>    math/legpoly.h:47-61, math/spherical.h:54-260
//
// Design: the reference materialises an `alias` array of Ncmplx
// reals and reads it back in every pass; here the window factor is the product
// of three per-axis tables (computed on the host with libm, quirk Q1 included)
// and the "cell in range" flag is re-derived from the same IEEE operations, so a
// pass reads each complex cell exactly once and writes nothing but nl*nbin sums.
// One warp owns one (i,j) row of the half spectrum; along the row |k| is
// monotone, so lanes with equal bin index are contiguous and a segmented warp
// scan leaves one add per (bin, warp-row) into warp-private shared-memory bins —
// no atomics, deterministic order.  Block partials are reduced by a second tiny
// kernel in fixed order.

#include "psb_internal.h"
#include "psb_ylm.cuh"

#include <algorithm>

namespace psb {

namespace {

enum { MODE_GEOM = 0, MODE_SIM = 1, MODE_SURVEY = 2, MODE_SURVEY_YLM = 3 };

template <typename real> struct C2;
template <> struct C2<double> { using type = double2; };
template <> struct C2<float> { using type = float2; };

// math/legpoly.h:47-61
__device__ __forceinline__ double legendre(int ell, double x) {
  const double x2 = x * x;
  switch (ell) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return 1.5 * x2 - 0.5;
    case 3: return 2.5 * x * (x2 - 0.6);
    case 4: return 4.375 * x2 * x2 - 3.75 * x2 + 0.375;
    case 5: return 8.75 * x * (0.9 * x2 * x2 - x2 + 0x1.b6db6db6db6dbp-3);
    case 6: return 14.4375 * x2 * x2 * x2 - 19.6875 * x2 * x2 + 6.5625 * x2 - 0.3125;
    default: return 0.0;
  }
}

// Bin of a squared wavenumber.  The reference decides it with
//   kmod = sqrt(k2) [or 0.5*log10(k2)];  unused if kmod < kedge[0] || kmod >= kedge[nbin];
//   bin = (int) ((kmod - kedge[0]) / dk)              (src/multipole.c:145-159)
// which is a monotone step function of k2.  The host bisects, with exactly those
// operations (and glibc's log10 for log bins), the smallest double k2 that lands
// in each bin; the device then only needs a cheap guess and two comparisons
// against that table instead of an IEEE sqrt and division per cell — and the
// result is the reference's bin by construction.  `edges` = nbin+1 thresholds
// (shared memory); returns -1 for cells the reference marks unused (alias = 0).
__device__ __forceinline__ int bin_of(const BinGeom &g, const double *edges, double k2,
    double rk) {
  if (!(k2 >= edges[0]) || k2 >= edges[g.nbin]) return -1;
  // rk = 1/sqrt(k2) (or anything: the guess is corrected below)
  const double kc = g.logk ? 0.5 * log10(k2) : k2 * rk;
  int b = (int) ((kc - g.k0) * g.inv_dk);
  b = max(0, min(b, g.nbin - 1));
  while (k2 < edges[b]) b--;
  while (k2 >= edges[b + 1]) b++;
  return b;
}

// Segmented inclusive scan over contiguous equal keys, then the last lane of
// every segment adds its total into the warp's private bins.
template <int NV>
__device__ __forceinline__ void warp_bin_add(int key, double (&v)[NV], double *bins,
    int nbin, int lane) {
  // keys are non-decreasing along the used lanes, so once no lane finds its own key `off`
  // lanes below, every segment is shorter than `off` and the remaining steps would add
  // nothing: leave (warp-uniform).  Most rows have segments of 2-4 lanes.
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int kprev = __shfl_up_sync(0xffffffffu, key, off);
    const bool take = (lane >= off) && (kprev == key);
    if (__ballot_sync(0xffffffffu, take && key >= 0) == 0u) break;
#pragma unroll
    for (int q = 0; q < NV; q++) {
      const double t = __shfl_up_sync(0xffffffffu, v[q], off);
      if (take) v[q] += t;
    }
  }
  const int knext = __shfl_down_sync(0xffffffffu, key, 1);
  if (key >= 0 && (lane == 31 || knext != key)) {
#pragma unroll
    for (int q = 0; q < NV; q++) bins[q * nbin + key] += v[q];
  }
  __syncwarp();
}

template <typename real>
__device__ __forceinline__ double2 ld_c(const void *base, size_t idx) {
  using c2 = typename C2<real>::type;
  c2 v = __ldg(reinterpret_cast<const c2 *>(base) + idx);
  return make_double2((double) v.x, (double) v.y);
}

// MODE_GEOM : v = {cnt, km, lcnt[0..nl)}            NV = nl + 2
// MODE_SIM  : v = {pl[0..nl)}                       NV = nl
// MODE_SURVEY: v = {pl}                             NV = 1
// EVEN: the multipoles are 0, 2, 4, ... (poles[l] = 2 l, the usual request): ell is a
// compile-time constant of the unrolled loop, so the Legendre switch and the odd-ell test
// fold away (7 % of the instructions of this issue-bound kernel)
template <typename real, int NV, int MODE, bool INTERLACE, bool EVEN = false>
__global__ void __launch_bounds__(256) k_spectrum(BinGeom g, const void *__restrict__ Fa0,
    const void *__restrict__ Fa1, const void *__restrict__ Fb0,
    const void *__restrict__ Fb1, double *__restrict__ partials) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  const int nacc = NV * g.nbin;
  // layout: [nbin+1 bin thresholds][nwarp x nacc warp-private bins]
  double *edges = smem;
  double *bins = smem + (g.nbin + 1) + (size_t) warp * nacc;
  for (int q = threadIdx.x; q <= g.nbin; q += blockDim.x) edges[q] = g.k2edge[q];
  for (int q = lane; q < nacc; q += 32) bins[q] = 0.0;
  __syncthreads();

  const bool cross = (Fb0 != Fa0);
  // The mode-counting pass is pure geometry: when the line of sight has no x
  // (y) component the summands are even in n_x (n_y) — k^2 tables are exactly
  // symmetric — so only n >= 0 is visited and doubled (exact in binary FP).
  const int half = (g.ng >> 1) + 1;
  const int nI = (MODE == MODE_GEOM && g.symx) ? half : g.ng;
  const int nJ = (MODE == MODE_GEOM && g.symy) ? half : g.nj;
  const size_t nrows = (size_t) nI * nJ;
  const size_t wstride = (size_t) gridDim.x * nwarp;
  for (size_t row = (size_t) blockIdx.x * nwarp + warp; row < nrows; row += wstride) {
    const int i = (int) (row / nJ), j = g.j0 + (int) (row % nJ);
    double wsym = 1.0;
    if (MODE == MODE_GEOM) {
      const bool self_i = (i == 0) || (((g.ng & 1) == 0) && i == (g.ng >> 1));
      const bool self_j = (j == 0) || (((g.ng & 1) == 0) && j == (g.ng >> 1));
      if (g.symx && !self_i) wsym *= 2.0;
      if (g.symy && !self_j) wsym *= 2.0;
      // equal box sides and no line-of-sight component along x or y: (i, j) and (j, i) have
      // the same k^2, |k| and mu — only j <= i is visited
      if (g.symxy && g.j0 == 0 && g.nj == g.ng) {
        if (j > i) continue;
        if (j < i) wsym *= 2.0;
      }
    }
    const double ki = __ldg(g.kax[0] + i), kj = __ldg(g.kax[1] + j);
    const double k2ij = __dadd_rn(__ldg(g.kax2[0] + i), __ldg(g.kax2[1] + j));
    const double wij = __dmul_rn(__ldg(g.wax[0] + i), __ldg(g.wax[1] + j));
    const double muij = __dadd_rn(__dmul_rn(ki, g.los[0]), __dmul_rn(kj, g.los[1]));
    double yaz = 0.0, kxy2 = 0.0, kxy = 0.0;   // survey l > 0: azimuthal part of Y_lm(k_hat)
    if (MODE == MODE_SURVEY_YLM) {
      kxy2 = __dadd_rn(__dmul_rn(ki, ki), __dmul_rn(kj, kj));
      if (kxy2 != 0.0) {
        const double rxy = rsqrt(kxy2);
        kxy = kxy2 * rxy;
        yaz = g.ylm_nrm * ylm_azimuth(g.m, ki * rxy, kj * rxy);
      }
    }
    double pcij = 1.0, psij = 0.0;
    if (INTERLACE) {
      const double ci = __ldg(g.pc[0] + i), si = __ldg(g.ps[0] + i);
      const double cj = __ldg(g.pc[1] + j), sj = __ldg(g.ps[1] + j);
      pcij = ci * cj - si * sj;
      psij = si * cj + ci * sj;
    }
    const size_t rbase = ((size_t) i * g.nj + (j - g.j0)) * (size_t) g.ngk;
    // Cells beyond the last bin edge (with KMAX at the Nyquist frequency: the
    // corners of the cube, 48 % of all cells) are never read: along a row k^2
    // grows with k, so a row whose k = 0 cell is already out is skipped and a
    // row is left at the first 32-cell chunk that starts out of range.  The
    // test is the same `k2 < edges[nbin]` that decides the cells one by one.
    const double k2max = edges[g.nbin], k2min = edges[0];
    if (!(k2ij < k2max)) continue;
    for (int kb = 0; kb < g.ngk; kb += 32) {
      if (!(__dadd_rn(k2ij, __ldg(g.kax2[2] + kb)) < k2max)) break;
      if (__dadd_rn(k2ij, __ldg(g.kax2[2] + min(kb + 31, g.ngk - 1))) < k2min) continue;
      const int k = kb + lane;
      int key = -1;
      double v[NV];
#pragma unroll
      for (int q = 0; q < NV; q++) v[q] = 0.0;
      if (k < g.ngk) {
        double2 a0, a1, b0, b1;
        if (MODE != MODE_GEOM) {
          a0 = ld_c<real>(Fa0, rbase + k);
          if (INTERLACE) a1 = ld_c<real>(Fa1, rbase + k);
          if (cross) {
            b0 = ld_c<real>(Fb0, rbase + k);
            if (INTERLACE) b1 = ld_c<real>(Fb1, rbase + k);
          }
        }
        const double k2 = __dadd_rn(k2ij, __ldg(g.kax2[2] + k));
        const double rk = (k2 > 0.0) ? rsqrt(k2) : 0.0;
        key = bin_of(g, edges, k2, rk);
        // sims skip the DC mode in the sums (src/multipole.c:806,939) but not
        // in the counts (quirk Q3)
        if (MODE == MODE_SIM && k2 == 0.0) key = -1;
        if (key >= 0) {
          const bool edge = (k == 0) || (((g.ng & 1) == 0) && k == (g.ng >> 1));
          const double mult = (edge ? 1.0 : 2.0) * wsym;
          double p = 0.0;
          if (MODE != MODE_GEOM) {
            if (INTERLACE) {
              // delta = (F0 + e^{i s} F1) / 2, s = pi (n_i + n_j + k) / Ng
              const double ck = __ldg(g.pc[2] + k), sk = __ldg(g.ps[2] + k);
              const double c = pcij * ck - psij * sk, s = psij * ck + pcij * sk;
              a0.x = 0.5 * (a0.x + c * a1.x - s * a1.y);
              a0.y = 0.5 * (a0.y + s * a1.x + c * a1.y);
              if (cross) {
                b0.x = 0.5 * (b0.x + c * b1.x - s * b1.y);
                b0.y = 0.5 * (b0.y + s * b1.x + c * b1.y);
              }
            }
            if (!cross) b0 = a0;
            const double alias = __dmul_rn(wij, __ldg(g.wax[2] + k));
            p = (a0.x * b0.x + a0.y * b0.y) * alias * mult;
          }
          if (MODE == MODE_SURVEY) v[0] = p;
          else if (MODE == MODE_SURVEY_YLM) {
            // Re(Fk0 conj Fka_m) Y_lm(k_hat): by linearity this sums to the
            // reference's Re(Fk0 conj Fkl), Fkl = sum_m Fka_m Y_lm(k_hat)
            // (src/mp_template.c:101-137), without the Fkl field.  k_xy = 0 or
            // |k| = 0 enter unweighted (:121-124).
            double sh = 1.0;
            if (kxy2 != 0.0 && k2 != 0.0) {
              const double kz = __ldg(g.kax[2] + k);
              sh = yaz * ylm_polar(g.ell, g.m < 0 ? -g.m : g.m, kz * rk, kxy * rk);
            }
            v[0] = p * sh;
          }
          else {
            // mu = k.los / |k|  (src/multipole.c:812-813)
            const double mu = __dadd_rn(muij, __dmul_rn(__ldg(g.kax[2] + k), g.los[2])) * rk;
            constexpr int L0 = (MODE == MODE_GEOM) ? 2 : 0;
            if (MODE == MODE_GEOM) {
              v[0] = mult;
              v[1] = mult * (g.logk ? 0.0 : k2 * rk);   // |k|; overwritten for log bins (Q7)
            }
#pragma unroll
            for (int l = 0; l < NV - L0; l++) {
              const int ell = EVEN ? 2 * l : g.poles[l];
              // +mu / -mu half-planes cancel for odd ell (src/multipole.c:843-844):
              // only the k = 0 and Nyquist planes contribute (quirk Q6)
              if (!edge && (ell & 1)) continue;
              const double leg = legendre(ell, mu);
              // the DC mode has no direction: counted in cnt/km only
              if (MODE == MODE_GEOM) v[L0 + l] = (k2 == 0.0) ? 0.0 : mult * leg;
              else v[l] = p * leg;
            }
          }
        }
      }
      warp_bin_add<NV>(key, v, bins, g.nbin, lane);
    }
  }
  __syncthreads();
  // block partial: sum the warps' bins in fixed order
  double *out = partials + (size_t) blockIdx.x * nacc;
  for (int q = threadIdx.x; q < nacc; q += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarp; w++) s += smem[(g.nbin + 1) + (size_t) w * nacc + q];
    out[q] = s;
  }
}

// out[q] += sum over blocks, in block order
__global__ void k_reduce_partials(const double *__restrict__ partials, int nblk, int nacc,
    double *__restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nacc) return;
  double s = 0.0;
  for (int b = 0; b < nblk; b++) s += partials[(size_t) b * nacc + q];
  out[q] += s;
}

__global__ void k_unpack_geometry(const double *__restrict__ acc, int nbin, int nl,
    unsigned long long *__restrict__ cnt, double *__restrict__ km,
    double *__restrict__ lcnt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbin) return;
  cnt[b] = (unsigned long long) acc[b];         // exact: integer-valued sums < 2^53
  km[b] = acc[nbin + b];
  for (int l = 0; l < nl; l++) lcnt[l * nbin + b] = acc[(2 + l) * nbin + b];
}

int g_bin_threads = 256;        // block size of the binning kernels (option "bin_threads")

struct LaunchShape { int blocks, threads; size_t smem; };

template <typename K>
int shape_for(K kernel, int nacc, int nbin, LaunchShape &ls) {
  // bin thresholds + warp-private bins (nacc doubles per warp)
  int threads = g_bin_threads;
  auto bytes = [&](int t) { return ((size_t) nacc * (t / 32) + nbin + 1) * sizeof(double); };
  size_t smem = bytes(threads);
  while (smem > 200 * 1024 && threads > 32) {
    threads >>= 1;
    smem = bytes(threads);
  }
  if (smem > 200 * 1024) {
    set_error("too many (multipole x k-bin) accumulators for the binning kernel: %d\n", nacc);
    return -1;
  }
  PSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
      (int) smem));
  int per_sm = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ls.blocks = sms * per_sm;
  ls.threads = threads;
  ls.smem = smem;
  return 0;
}

constexpr int MAX_BLOCKS = 148 * 8;
// blocks per SM of the mode-counting pass (0: what fits; 1 measured best: list stage 5.21 vs
// 5.35 ms); it runs beside the tile-list
// pass on a side stream and competes with it for the SMs' registers (option "geom_blocks")
int g_geom_blocks = 1;

template <typename real, int NV, int MODE, bool IL, bool EVEN = false>
int run_spectrum(const BinGeom &g, const void *Fa0, const void *Fa1, const void *Fb0,
    const void *Fb1, double *out, double *scratch, size_t scratch_bytes, cudaStream_t st) {
  auto kern = k_spectrum<real, NV, MODE, IL, EVEN>;
  const int nacc = NV * g.nbin;
  LaunchShape ls;
  if (shape_for(kern, nacc, g.nbin, ls)) return -1;
  const int half = (g.ng >> 1) + 1;
  size_t rows = (size_t) ((MODE == MODE_GEOM && g.symx) ? half : g.ng)
      * ((MODE == MODE_GEOM && g.symy) ? half : g.nj);
  size_t need_blocks = (rows + ls.threads / 32 - 1) / (ls.threads / 32);
  if (MODE == MODE_GEOM && g_geom_blocks > 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ls.blocks = std::min(ls.blocks, sms * g_geom_blocks);
  }
  if ((size_t) ls.blocks > need_blocks) ls.blocks = (int) need_blocks;
  if (ls.blocks > MAX_BLOCKS) ls.blocks = MAX_BLOCKS;
  if ((size_t) ls.blocks * nacc * sizeof(double) > scratch_bytes) {
    set_error("internal: binning scratch too small\n");
    return -1;
  }
  kern<<<ls.blocks, ls.threads, ls.smem, st>>>(g, Fa0, Fa1, Fb0, Fb1, scratch);
  PSB_CUDA(cudaGetLastError());
  k_reduce_partials<<<(nacc + 127) / 128, 128, 0, st>>>(scratch, ls.blocks, nacc, out);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

void bin_set_geom_blocks(int n) { g_geom_blocks = n; }
void bin_set_threads(int n) { g_bin_threads = (n == 64 || n == 128 || n == 256) ? n : 256; }

size_t bin_scratch_bytes(const BinGeom &g) {
  // block partials plus one accumulator row for the geometry pass
  return ((size_t) MAX_BLOCKS + 1) * (size_t) (g.nl + 2) * g.nbin * sizeof(double);
}

int launch_geometry(const BinGeom &g, unsigned long long *cnt, double *km, double *lcnt,
    double *scratch, size_t scratch_bytes, cudaStream_t st) {
  const int nl = g.issim ? g.nl : 0;
  const int nacc = (nl + 2) * g.nbin;
  double *acc = scratch;        // first nacc doubles: the reduced accumulators
  double *part = scratch + nacc;
  PSB_CUDA(cudaMemsetAsync(acc, 0, nacc * sizeof(double), st));
  int rc = -1;
  const size_t pb = scratch_bytes - nacc * sizeof(double);
#define PSB_GEOM(NL)                                                            \
  case NL:                                                                      \
    rc = run_spectrum<double, NL + 2, MODE_GEOM, false>(g, nullptr, nullptr,    \
        nullptr, nullptr, acc, part, pb, st);                                   \
    break;
  switch (nl) {
    PSB_GEOM(0) PSB_GEOM(1) PSB_GEOM(2) PSB_GEOM(3) PSB_GEOM(4) PSB_GEOM(5)
    PSB_GEOM(6) PSB_GEOM(7)
    default: set_error("invalid number of multipoles: %d\n", nl); return -1;
  }
#undef PSB_GEOM
  if (rc) return rc;
  k_unpack_geometry<<<(g.nbin + 127) / 128, 128, 0, st>>>(acc, g.nbin, nl, cnt, km, lcnt);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

template <typename real>
static int launch_bin_t(const BinGeom &g, const void *Fa0, const void *Fa1, const void *Fb0,
    const void *Fb1, double *pl, double *scratch, size_t sb, cudaStream_t st) {
  const bool il = (Fa1 != nullptr);
  if (!g.issim && g.ell > 0)
    return run_spectrum<real, 1, MODE_SURVEY_YLM, false>(g, Fa0, nullptr, Fb0, nullptr, pl, scratch, sb, st);
  if (!g.issim) {
    return il ? run_spectrum<real, 1, MODE_SURVEY, true>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st)
              : run_spectrum<real, 1, MODE_SURVEY, false>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st);
  }
  bool even = g.nl <= 4;
  for (int l = 0; l < g.nl; l++) even = even && g.poles[l] == 2 * l;
#define PSB_SIM_EVEN(NL)                                                        \
  case NL:                                                                      \
    return il ? run_spectrum<real, NL, MODE_SIM, true, true>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st)  \
              : run_spectrum<real, NL, MODE_SIM, false, true>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st);
  if (even)
    switch (g.nl) { PSB_SIM_EVEN(1) PSB_SIM_EVEN(2) PSB_SIM_EVEN(3) PSB_SIM_EVEN(4) default: break; }
#undef PSB_SIM_EVEN
#define PSB_SIM(NL)                                                             \
  case NL:                                                                      \
    return il ? run_spectrum<real, NL, MODE_SIM, true>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st)  \
              : run_spectrum<real, NL, MODE_SIM, false>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, sb, st);
  switch (g.nl) {
    PSB_SIM(1) PSB_SIM(2) PSB_SIM(3) PSB_SIM(4) PSB_SIM(5) PSB_SIM(6) PSB_SIM(7)
    default: set_error("invalid number of multipoles: %d\n", g.nl); return -1;
  }
#undef PSB_SIM
}

int launch_bin(const BinGeom &g, int precision, const void *Fa0, const void *Fa1,
    const void *Fb0, const void *Fb1, double *pl, double *scratch, size_t scratch_bytes,
    cudaStream_t st) {
  return precision == 8
      ? launch_bin_t<double>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, scratch_bytes, st)
      : launch_bin_t<float>(g, Fa0, Fa1, Fb0, Fb1, pl, scratch, scratch_bytes, st);
}

// ---------------------------------------------------------------------------
// in-place interlace combination (surveys: the combined field is reused)
// ---------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) k_combine(BinGeom g, typename C2<real>::type *F0,
    const typename C2<real>::type *__restrict__ F1) {
  using c2 = typename C2<real>::type;
  const size_t nrows = (size_t) g.ng * g.nj;
  for (size_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int i = (int) (row / g.nj), j = g.j0 + (int) (row % g.nj);
    const double ci = g.pc[0][i], si = g.ps[0][i], cj = g.pc[1][j], sj = g.ps[1][j];
    const double cij = ci * cj - si * sj, sij = si * cj + ci * sj;
    for (int k = threadIdx.x; k < g.ngk; k += blockDim.x) {
      const double ck = g.pc[2][k], sk = g.ps[2][k];
      const double c = cij * ck - sij * sk, s = sij * ck + cij * sk;
      const size_t idx = row * g.ngk + k;
      c2 a = F0[idx], b = F1[idx];
      c2 r;
      r.x = (real) (0.5 * ((double) a.x + c * (double) b.x - s * (double) b.y));
      r.y = (real) (0.5 * ((double) a.y + s * (double) b.x + c * (double) b.y));
      F0[idx] = r;
    }
  }
}

int launch_combine(const BinGeom &g, int precision, void *F0, const void *F1,
    cudaStream_t st) {
  if (precision == 8)
    k_combine<double><<<148 * 8, 256, 0, st>>>(g, static_cast<double2 *>(F0),
        static_cast<const double2 *>(F1));
  else
    k_combine<float><<<148 * 8, 256, 0, st>>>(g, static_cast<float2 *>(F0),
        static_cast<const float2 *>(F1));
  PSB_CUDA(cudaGetLastError());
  return 0;
}

double ylm_norm(int l, int m) {
  const int am = m < 0 ? -m : m;
  double ratio = 1.0;
  for (int k = l - am + 1; k <= l + am; k++) ratio /= k;
  double nrm = sqrt((2 * l + 1) / (4 * 0x1.921fb54442d18p+1) * ratio);
  if (am) nrm *= 0x1.6a09e667f3bcdp+0;
  return nrm;
}

// out = Fr * Y_lm(r_hat), src/mp_template.c:65-95.  Both meshes padded (rowlen).
// One block per mesh row; a thread handles two adjacent cells per 16-byte (8-byte)
// access and two such pairs per iteration, so four loads are in flight per thread
// (one cell per iteration left the pass latency-bound at 3.4 TB/s).
template <typename real> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

// the cells of one row, l and |m| fixed at compile time
template <typename real, int L, int AM>
__device__ __forceinline__ void ylm_weight_row(const YlmGeom &g, const real *__restrict__ src,
    real *__restrict__ dst, double az, double r2, double rxy) {
  using pair = typename Pair<real>::type;
  const int npair = g.ng >> 1;                    // rowlen is even and rows are pair-aligned
  auto weight = [&](int k) {
    const double rk = (k + g.smin[2]) * g.bsize[2];
    const double ir3 = rsqrt(r2 + rk * rk);
    return az * ylm_polar_fixed<L, AM>(rk * ir3, rxy * ir3);
  };
  const pair *s2 = reinterpret_cast<const pair *>(src);
  pair *d2 = reinterpret_cast<pair *>(dst);
  for (int q = threadIdx.x; q < npair; q += 2 * blockDim.x) {
    const int q1 = q + blockDim.x;
    const bool two = q1 < npair;
    const pair a = s2[q];
    pair b = a;
    if (two) b = s2[q1];
    pair ra, rb;
    ra.x = (real) ((double) a.x * weight(2 * q));
    ra.y = (real) ((double) a.y * weight(2 * q + 1));
    d2[q] = ra;
    if (two) {
      rb.x = (real) ((double) b.x * weight(2 * q1));
      rb.y = (real) ((double) b.y * weight(2 * q1 + 1));
      d2[q1] = rb;
    }
  }
  if ((g.ng & 1) && threadIdx.x == 0) {                   // odd Ng: the last cell
    const int k = g.ng - 1;
    dst[k] = (real) ((double) src[k] * weight(k));
  }
}

template <typename real, int L>
__global__ void __launch_bounds__(256) k_ylm_weight_r(YlmGeom g, double nrm,
    const real *__restrict__ Fr, real *__restrict__ out) {
  const size_t nrows = (size_t) g.ng * g.ng;
  const int am = g.m < 0 ? -g.m : g.m;
  for (size_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int i = (int) (row / g.ng), j = (int) (row % g.ng);
    const double ri = (i + g.smin[0]) * g.bsize[0];
    const double rj = (j + g.smin[1]) * g.bsize[1];
    const double r2 = ri * ri + rj * rj;
    const double rxy = sqrt(r2);
    // azimuth (cos phi = ri / rxy, sin phi = rj / rxy) is constant along the row
    // (a row through the origin has no azimuth; its polar factor sin^|m| is 0 for m != 0)
    const double az = nrm * (rxy > 0.0 ? ylm_azimuth(g.m, ri / rxy, rj / rxy) : 1.0);
    const real *src = Fr + row * g.rowlen;
    real *dst = out + row * g.rowlen;
    if (row == 0) {                                         // quirk Q5 (:75-78)
      for (int k = threadIdx.x; k < g.ng; k += blockDim.x) dst[k] = src[k];
      continue;
    }
    switch (am) {                                           // uniform per launch
      case 0: ylm_weight_row<real, L, 0>(g, src, dst, az, r2, rxy); break;
      case 1: if (L >= 1) ylm_weight_row<real, L, (L >= 1 ? 1 : 0)>(g, src, dst, az, r2, rxy); break;
      case 2: if (L >= 2) ylm_weight_row<real, L, (L >= 2 ? 2 : 0)>(g, src, dst, az, r2, rxy); break;
      case 3: if (L >= 3) ylm_weight_row<real, L, (L >= 3 ? 3 : 0)>(g, src, dst, az, r2, rxy); break;
      case 4: if (L >= 4) ylm_weight_row<real, L, (L >= 4 ? 4 : 0)>(g, src, dst, az, r2, rxy); break;
      case 5: if (L >= 5) ylm_weight_row<real, L, (L >= 5 ? 5 : 0)>(g, src, dst, az, r2, rxy); break;
      default: if (L >= 6) ylm_weight_row<real, L, (L >= 6 ? 6 : 0)>(g, src, dst, az, r2, rxy); break;
    }
  }
}

template <typename real>
static int launch_ylm_weight_t(const YlmGeom &g, double nrm, const real *Fr, real *out, cudaStream_t st) {
  switch (g.ell) {
    case 1: k_ylm_weight_r<real, 1><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    case 2: k_ylm_weight_r<real, 2><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    case 3: k_ylm_weight_r<real, 3><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    case 4: k_ylm_weight_r<real, 4><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    case 5: k_ylm_weight_r<real, 5><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    case 6: k_ylm_weight_r<real, 6><<<148 * 8, 256, 0, st>>>(g, nrm, Fr, out); break;
    default: set_error("invalid multipole %d\n", g.ell); return -1;
  }
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_ylm_weight_r(const YlmGeom &g, int precision, const void *Fr, void *out,
    cudaStream_t st) {
  const double nrm = ylm_norm(g.ell, g.m);
  if (precision == 8)
    return launch_ylm_weight_t<double>(g, nrm, static_cast<const double *>(Fr), static_cast<double *>(out), st);
  return launch_ylm_weight_t<float>(g, nrm, static_cast<const float *>(Fr), static_cast<float *>(out), st);
}

// Fkl += Fka * Y_lm(k_hat) on cells the reference marks used, src/mp_template.c:101-137
template <typename real>
__global__ void __launch_bounds__(256) k_ylm_accum_k(YlmGeom g, BinGeom bg, double nrm,
    const typename C2<real>::type *__restrict__ Fka, typename C2<real>::type *Fkl) {
  using c2 = typename C2<real>::type;
  const size_t nrows = (size_t) g.ng * g.ng;
  const double f0 = 1.0 / g.bsize[0], f1 = 1.0 / g.bsize[1], f2 = 1.0 / g.bsize[2];
  for (size_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int i = (int) (row / g.ng), j = (int) (row % g.ng);
    const double ki = f0 * ((i <= (g.ng >> 1)) ? i : i - g.ng);
    const double kj = f1 * ((j <= (g.ng >> 1)) ? j : j - g.ng);
    const double k2 = ki * ki + kj * kj;
    const double k2ij = __dadd_rn(bg.kax2[0][i], bg.kax2[1][j]);
    for (int k = threadIdx.x; k < g.ngk; k += blockDim.x) {
      const double k2b = __dadd_rn(k2ij, bg.kax2[2][k]);
      if (!(k2b >= bg.k2edge[0]) || k2b >= bg.k2edge[bg.nbin]) continue;   // unused cell
      const double kk = f2 * k;
      const double k3 = sqrt(k2 + kk * kk);
      double sh = 1.0;
      if (k2 != 0.0 && k3 != 0.0) {
        const double kxy = sqrt(k2);
        sh = ylm_real(g.ell, g.m, nrm, kk / k3, kxy / k3, ki / kxy, kj / kxy);
      }
      const size_t idx = row * g.ngk + k;
      c2 a = Fka[idx], acc = Fkl[idx];
      acc.x = (real) ((double) acc.x + (double) a.x * sh);
      acc.y = (real) ((double) acc.y + (double) a.y * sh);
      Fkl[idx] = acc;
    }
  }
}

int launch_ylm_accum_k(const YlmGeom &g, const BinGeom &bg, int precision,
    const void *Fka, void *Fkl, cudaStream_t st) {
  const double nrm = ylm_norm(g.ell, g.m);
  if (precision == 8)
    k_ylm_accum_k<double><<<148 * 8, 256, 0, st>>>(g, bg, nrm,
        static_cast<const double2 *>(Fka), static_cast<double2 *>(Fkl));
  else
    k_ylm_accum_k<float><<<148 * 8, 256, 0, st>>>(g, bg, nrm,
        static_cast<const float2 *>(Fka), static_cast<float2 *>(Fkl));
  PSB_CUDA(cudaGetLastError());
  return 0;
}

template <typename real>
__global__ void k_scale(real *m, size_t n, double f) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x)
    m[i] = (real) ((double) m[i] * f);
}

int launch_scale(void *mesh, size_t n, double factor, int precision, cudaStream_t st) {
  if (precision == 8) k_scale<double><<<148 * 16, 256, 0, st>>>((double *) mesh, n, factor);
  else k_scale<float><<<148 * 16, 256, 0, st>>>((float *) mesh, n, factor);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb
