// Coordinate conversion of survey catalogues on the device: (RA [deg], Dec [deg],
// redshift) -> comoving Cartesian coordinates, in place on the 32-byte particle
// records.  Replaces cnvt_coord(), src/cnvt_coord.c:549-582:
// the reference runs this as a separate host pass over the arrays that are about
// to be uploaded; here it runs on the records once they are in HBM, before the
// bounds reduction.
//
//   * integration mode (src/cnvt_coord.c:295-343, 407-423): the comoving distance
//     is a Legendre-Gauss quadrature of c / (100 E(z)) over [0, z]; the order is
//     the smallest one (4..32) that converges to CONF.ecdst on 128 redshifts
//     spanning the catalogues (:356-396, :495-511).  The reference reads abscissas
//     and weights from a table (math/legauss.c); here they are computed on the host
//     by Newton iteration on P_n in long double (same numbers to the last bit or
//     one ulp) and passed in constant memory.
//   * interpolation mode (:102-137, :440-486): natural cubic spline through the
//     (z, d) samples of CONF.fcdst, bracket by bisection, evaluation as
//     math/cspline.c:94-108.
//
// The arithmetic keeps the reference's order of operations without FMA
// contraction (-std=c99, Makefile:2); sin/cos/pow are CUDA's (<= 2 ulp) where
// the reference has glibc's, so coordinates agree to a few ulp, not bit for bit.

#include "psb_internal.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <vector>

namespace psb {

namespace {

constexpr int LG_MIN = 4, LG_MAX = 32;          // math/legauss.h:40-41
constexpr double C_LIGHT = 299792.458;          // src/define.h:37
constexpr double DEG2RAD = 0x1.1df46a2529d39p-6;    // src/define.h:41

__constant__ double c_lgx[LG_MAX / 2 + 1];      // non-negative abscissas, largest first
__constant__ double c_lgw[LG_MAX / 2 + 1];      // weights; for odd orders the last is x = 0

// c / (100 E(z)), src/cnvt_coord.c:295-306
__host__ __device__ inline double integrand(double om, double ol, double ok, double widx, double z) {
#ifdef __CUDA_ARCH__
  const double z1 = __dadd_rn(z, 1.0);
  const double z2 = __dmul_rn(z1, z1);
  double d = __dmul_rn(__dmul_rn(om, z2), z1);
  if (ok != 0.0) d = __dadd_rn(d, __dmul_rn(ok, z2));
  if (widx != 0.0) d = __dadd_rn(d, __dmul_rn(ol, pow(z1, widx)));
  else d = __dadd_rn(d, ol);
  return __ddiv_rn(C_LIGHT * 0.01, __dsqrt_rn(d));
#else
  const double z1 = z + 1;
  const double z2 = z1 * z1;
  double d = om * z2 * z1;
  if (ok != 0.0) d += ok * z2;
  if (widx != 0.0) d += ol * std::pow(z1, widx);
  else d += ol;
  return C_LIGHT * 0.01 / std::sqrt(d);
#endif
}

// src/cnvt_coord.c:321-343 with the rule given as arrays
__host__ __device__ inline double legauss(const double *x, const double *w, int order, double om,
    double ol, double ok, double widx, double z) {
  const double zp = z * 0.5;
  double sum = 0;
  const int half = order >> 1;
  for (int i = 0; i < half; i++) {
#ifdef __CUDA_ARCH__
    const double za = __dmul_rn(zp, __dadd_rn(1.0, x[i])), zb = __dmul_rn(zp, __dsub_rn(1.0, x[i]));
    sum = __dadd_rn(sum, __dmul_rn(w[i], __dadd_rn(integrand(om, ol, ok, widx, za),
        integrand(om, ol, ok, widx, zb))));
#else
    const double za = zp * (1 + x[i]), zb = zp * (1 - x[i]);
    sum += w[i] * (integrand(om, ol, ok, widx, za) + integrand(om, ol, ok, widx, zb));
#endif
  }
#ifdef __CUDA_ARCH__
  if (order & 1) sum = __dadd_rn(sum, __dmul_rn(w[half], integrand(om, ol, ok, widx, zp)));
  return __dmul_rn(sum, zp);
#else
  if (order & 1) sum += w[half] * integrand(om, ol, ok, widx, zp);
  return sum * zp;
#endif
}

// spherical -> Cartesian, src/cnvt_coord.c:415-422
__device__ __forceinline__ void to_cartesian(double2 &a, double2 &b, double dist) {
  const double ra = __dmul_rn(a.x, DEG2RAD), dec = __dmul_rn(a.y, DEG2RAD);
  double sr, cr, sd, cdec;
  sincos(ra, &sr, &cr);
  sincos(dec, &sd, &cdec);
  const double dc = __dmul_rn(dist, cdec);
  a.x = __dmul_rn(dc, cr);
  a.y = __dmul_rn(dc, sr);
  b.x = __dmul_rn(dist, sd);
}

__global__ void __launch_bounds__(256) k_cnvt_integr(double2 *__restrict__ p, size_t n, int order,
    double om, double ol, double ok, double widx) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    double2 a = p[2 * i], b = p[2 * i + 1];
    const double dist = legauss(c_lgx, c_lgw, order, om, ol, ok, widx, b.x);
    to_cartesian(a, b, dist);
    p[2 * i] = a;
    p[2 * i + 1] = b;
  }
}

// src/cnvt_coord.c:102-137 (bracket) + math/cspline.c:94-108 (evaluation);
// HUGE_VAL outside the sampled range as in the reference
__global__ void __launch_bounds__(256) k_cnvt_interp(double2 *__restrict__ p, size_t n,
    const double *__restrict__ z, const double *__restrict__ d, const double *__restrict__ ypp,
    size_t nsp) {
  for (size_t q = blockIdx.x * (size_t) blockDim.x + threadIdx.x; q < n;
       q += (size_t) gridDim.x * blockDim.x) {
    double2 a = p[2 * q], b = p[2 * q + 1];
    const double zv = b.x;
    double dist = HUGE_VAL;
    if (!(zv < z[0] || zv >= z[nsp - 1])) {
      // the interval z[i] <= zv < z[i + 1] (samples ascending, src/cnvt_coord.c:102-137)
      size_t i = 0, above = nsp - 1;
      while (above - i > 1) {
        const size_t mid = i + ((above - i) >> 1);
        if (z[mid] <= zv) i = mid;
        else above = mid;
      }
      // cubic through the bracketing samples (z_lo, z_hi): linear part plus the two
      // second-derivative terms, operations in the order of math/cspline.c:94-108
      const double z_lo = z[i], z_hi = z[i + 1];
      const double width = __dsub_rn(z_hi, z_lo), up = __dsub_rn(zv, z_lo), down = __dsub_rn(z_hi, zv);
      const double width2 = __dmul_rn(width, width);
      const double chord = __dadd_rn(__dmul_rn(up, d[i + 1]), __dmul_rn(down, d[i]));
      const double bend_hi = __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(up, up), width2), up), ypp[i + 1]);
      const double bend_lo = __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(down, down), width2), down), ypp[i]);
      dist = __ddiv_rn(__dadd_rn(chord, __dmul_rn(0x1.5555555555555p-3, __dadd_rn(bend_hi, bend_lo))), width);
    }
    to_cartesian(a, b, dist);
    p[2 * q] = a;
    p[2 * q + 1] = b;
  }
}

}  // namespace

// Legendre-Gauss rule of the given order: the order/2 positive abscissas, largest
// first, with their weights; odd orders append the weight of x = 0.
void legauss_rule(int order, double *x, double *w) {
  const int half = order >> 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < half; i++) {
    long double t = cosl(pi * (i + 0.75L) / (order + 0.5L)), dp = 1;
    for (int it = 0; it < 100; it++) {
      long double p0 = 1, p1 = t;
      for (int k = 2; k <= order; k++) {
        const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
        p0 = p1; p1 = p2;
      }
      dp = order * (t * p1 - p0) / (t * t - 1);
      const long double dt = p1 / dp;
      t -= dt;
      if (fabsl(dt) < 1e-19L) break;
    }
    {
      long double p0 = 1, p1 = t;
      for (int k = 2; k <= order; k++) {
        const long double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
        p0 = p1; p1 = p2;
      }
      dp = order * (t * p1 - p0) / (t * t - 1);
    }
    x[i] = (double) t;
    w[i] = (double) (2 / ((1 - t * t) * dp * dp));
  }
  if (order & 1) {
    long double p0 = 1, p1 = 0;       // P_{n-1}(0) by recurrence
    for (int k = 2; k <= order; k++) {
      const long double p2 = -((k - 1) * p0) / k;
      p0 = p1; p1 = p2;
    }
    const long double dp = order * p0;  // P_n'(0) = n P_{n-1}(0)
    x[half] = 0;
    w[half] = (double) (2 / (dp * dp));
  }
}

// Smallest Legendre-Gauss order whose integral agrees with the previous order's to the
// relative tolerance `err` on every one of `num` equally spaced redshifts of
// [zmin, zmax] (the convergence test of src/cnvt_coord.c:356-396 on the samples of
// :277-278; the comparison starts from 0 below the lowest order); INT_MAX if even the
// highest order does not converge somewhere.
int legauss_order(double om, double ol, double ok, double widx, double err, double zmin,
    double zmax, int num) {
  constexpr int SLOT = LG_MAX / 2 + 1;
  std::vector<double> nodes((LG_MAX + 1) * SLOT), weights(nodes.size());
  for (int n = LG_MIN; n <= LG_MAX; n++) legauss_rule(n, &nodes[n * SLOT], &weights[n * SLOT]);
  int order = 0;
  for (int i = 0; i < num; i++) {
    const double z = zmin + i * (zmax - zmin) / (num - 1);
    int need = INT_MAX;
    double prev = 0;
    for (int n = LG_MIN; n <= LG_MAX; n++) {
      const double cur = legauss(&nodes[n * SLOT], &weights[n * SLOT], n, om, ol, ok, widx, z);
      if (!(fabs(cur - prev) > cur * err)) { need = n; break; }
      prev = cur;
    }
    order = std::max(order, need);
  }
  return order;
}

// Second derivatives m[] of the natural cubic spline through (x, y): the tridiagonal
// system  h[i-1] m[i-1] + 2 (h[i-1] + h[i]) m[i] + h[i] m[i+1] = 6 (s[i] - s[i-1])  with
// m[0] = m[n-1] = 0, solved by the Thomas algorithm.  The individual operations are the
// ones math/cspline.c:38-82 performs (interval widths and slopes first, pivot as a
// reciprocal, right-hand side times 6), so the table — and with it every interpolated
// distance — is the reference's to the last bit.
int cspline_second(const double *x, const double *y, size_t n, double *m) {
  if (n < 2) return -1;
  const size_t nseg = n - 1;
  std::vector<double> h(nseg), slope(nseg), upper(n, 0.0);
  for (size_t i = 0; i < nseg; i++) {
    h[i] = x[i + 1] - x[i];
    slope[i] = (y[i + 1] - y[i]) / h[i];
  }
  m[0] = m[n - 1] = 0;
  // forward sweep over the interior nodes: upper[] is the eliminated super-diagonal
  for (size_t i = 1; i < nseg; i++) {
    const double diag = (h[i] + h[i - 1]) * 2;
    const double pivot = 1 / (diag - h[i - 1] * upper[i - 1]);
    const double rhs = (slope[i] - slope[i - 1]) * 6;
    m[i] = (rhs - h[i - 1] * m[i - 1]) * pivot;
    upper[i] = h[i] * pivot;
  }
  // back substitution (upper[0] = 0: m[0] stays 0)
  for (size_t i = nseg; i-- > 0;) m[i] -= upper[i] * m[i + 1];
  return 0;
}

int launch_cnvt_integr(double *p, size_t n, int order, double om, double ol, double ok,
    double widx, cudaStream_t st) {
  if (order < LG_MIN || order > LG_MAX) { set_error("invalid Legendre-Gauss order %d\n", order); return -1; }
  double x[LG_MAX / 2 + 1] = {0}, w[LG_MAX / 2 + 1] = {0};
  legauss_rule(order, x, w);
  PSB_CUDA(cudaMemcpyToSymbolAsync(c_lgx, x, sizeof x, 0, cudaMemcpyHostToDevice, st));
  PSB_CUDA(cudaMemcpyToSymbolAsync(c_lgw, w, sizeof w, 0, cudaMemcpyHostToDevice, st));
  if (!n) return 0;
  const int blocks = (int) std::min<size_t>((n + 255) / 256, 148 * 16);
  k_cnvt_integr<<<blocks, 256, 0, st>>>(reinterpret_cast<double2 *>(p), n, order, om, ol, ok, widx);
  PSB_CUDA(cudaGetLastError());
  // x and w are on this frame: the copies must have been staged before returning
  PSB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int launch_cnvt_interp(double *p, size_t n, const double *z, const double *d, const double *ypp,
    size_t nsp, cudaStream_t st) {
  if (!n) return 0;
  const int blocks = (int) std::min<size_t>((n + 255) / 256, 148 * 16);
  k_cnvt_interp<<<blocks, 256, 0, st>>>(reinterpret_cast<double2 *>(p), n, z, d, ypp, nsp);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb
