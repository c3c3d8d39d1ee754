// Real spherical harmonics on the device, shared by binning.cu (k-space factor,
// r-space weighting pass) and fft_strided.cu (r-space weighting fused into the
// load of the z pass).  Internal header.
#pragma once

namespace psb {
namespace {

__device__ const double g_inv_int[8] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0, 1.0 / 7.0};

// ---------------------------------------------------------------------------
// real spherical harmonics (definition math/spherical.h:33-38), by recurrence:
//   Y_lm = N_lm P_l^|m|(cos t) {cos(m p) | 1 | sin(|m| p)},  P without the
//   Condon-Shortley phase, N_lm = sqrt((2l+1)/(4 pi) (l-|m|)!/(l+|m|)!) sqrt2^{m!=0}
// ---------------------------------------------------------------------------
// split in the part that depends on the azimuth only (constant along a mesh
// row, where x and y are fixed) and the polar part (varies along the row)
__device__ __forceinline__ double ylm_azimuth(int m, double cosp, double sinp) {
  const int am = m < 0 ? -m : m;
  double cm = 1.0, sm = 0.0;
  for (int k = 0; k < am; k++) {
    const double c2 = cm * cosp - sm * sinp;
    sm = sm * cosp + cm * sinp;
    cm = c2;
  }
  return m > 0 ? cm : (m < 0 ? sm : 1.0);
}

__device__ __forceinline__ double ylm_polar(int l, int am, double cost, double sint) {
  double pmm = 1.0;
  for (int k = 1; k <= am; k++) pmm *= (2 * k - 1) * sint;
  if (l == am) return pmm;
  double pm1 = pmm, cur = (2 * am + 1) * cost * pmm;
  for (int ll = am + 2; ll <= l; ll++) {
    const double nxt = ((2 * ll - 1) * cost * cur - (ll + am - 1) * pm1) * g_inv_int[ll - am];
    pm1 = cur; cur = nxt;
  }
  return cur;
}

// the same recurrence with l and |m| known at compile time: the loops unroll and
// the reciprocals fold into constants (identical sequence of operations, so the
// result is the runtime version's bit for bit)
template <int L, int AM>
__device__ __forceinline__ double ylm_polar_fixed(double cost, double sint) {
  double pmm = 1.0;
#pragma unroll
  for (int k = 1; k <= AM; k++) pmm *= (2 * k - 1) * sint;
  if (L == AM) return pmm;
  double pm1 = pmm, cur = (2 * AM + 1) * cost * pmm;
#pragma unroll
  for (int ll = AM + 2; ll <= L; ll++) {
    const double inv = ll - AM == 1 ? 1.0 : ll - AM == 2 ? 0.5 : ll - AM == 3 ? 1.0 / 3.0
        : ll - AM == 4 ? 0.25 : ll - AM == 5 ? 0.2 : ll - AM == 6 ? 1.0 / 6.0 : 1.0 / 7.0;
    const double nxt = ((2 * ll - 1) * cost * cur - (ll + AM - 1) * pm1) * inv;
    pm1 = cur; cur = nxt;
  }
  return cur;
}

__device__ __forceinline__ double ylm_real(int l, int m, double nrm, double cost,
    double sint, double cosp, double sinp) {
  return nrm * ylm_polar(l, m < 0 ? -m : m, cost, sint) * ylm_azimuth(m, cosp, sinp);
}

}  // namespace
}  // namespace psb
