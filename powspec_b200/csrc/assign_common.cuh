// Device helpers shared by the mass-assignment kernels (assign.cu: global-reduction
// scatter; assign_tiles.cu: owner-computes tiles): the grid coordinate in the reference's
// operation order and the per-axis stencils of src/genr_mesh.c.
#pragma once

#include "psb_internal.h"

namespace psb {

// A particle record {x, y, z, w} is 32 bytes = one DRAM / L2 sector: moved with ONE
// 256-bit access (LDG.E.256 / STG.E.256 on sm_100a), so that the scattered writes of the
// particle sorts reach the L2 as full-sector writes instead of two half-sector ones.
__device__ __forceinline__ void ld_record(const double2 *p, size_t i, double2 &a, double2 &b) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p + 2 * i));
}
__device__ __forceinline__ void st_record(double2 *p, size_t i, double2 a, double2 b) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
      :: "l"(p + 2 * i), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// ---------------------------------------------------------------------------
// grid coordinate and base cell, in the reference's operation order
// ---------------------------------------------------------------------------
__device__ __forceinline__ double grid_coord(double x, double org, double len, int ng) {
  // (x - org) * Ng / L : src/genr_mesh.c:57,91,152,256
  return __ddiv_rn(__dmul_rn(__dsub_rn(x, org), (double) ng), len);
}

__device__ __forceinline__ int base_cell(double t, int ng) {
  int c = (int) t;
  // quirk Q8: a coordinate that rounds to t == Ng indexes out
  // of bounds in the reference; wrap it instead
  if (c >= ng) c -= ng;
  // coordinates outside the box are rejected by def_box's checks (evaluated
  // after the scatter for simulation boxes): never index out of bounds meanwhile
  return min(max(c, 0), ng - 1);
}

// ---------------------------------------------------------------------------
// one axis of the assignment stencil: cells (periodic) and weights
// ---------------------------------------------------------------------------
template <int SCHEME> struct Stencil { static constexpr int N = SCHEME + 1; };

__device__ __forceinline__ int wrap_up(int c, int ng) { return (c == ng - 1) ? 0 : c + 1; }
__device__ __forceinline__ int wrap_dn(int c, int ng) { return (c == 0) ? ng - 1 : c - 1; }

template <int SCHEME>
__device__ __forceinline__ void axis_stencil(double t, int ng, int (&idx)[SCHEME + 1],
    double (&w)[SCHEME + 1]) {
  int c = (int) t;
  double d = t - (double) c;    // exact
  if (c >= ng) c -= ng;         // Q8 guard, see base_cell()
  c = min(max(c, 0), ng - 1);
  if constexpr (SCHEME == 0) {  // NGP, src/genr_mesh.c:60-66
    if (d >= 0.5) c = wrap_up(c, ng);
    idx[0] = c; w[0] = 1.0;
  }
  else if constexpr (SCHEME == 1) {     // CIC, src/genr_mesh.c:98-108
    idx[0] = c; idx[1] = wrap_up(c, ng);
    w[1] = d; w[0] = 1.0 - d;
  }
  else if constexpr (SCHEME == 2) {     // TSC, src/genr_mesh.c:157-173
    double h;
    if (d < 0.5) {
      idx[1] = c; idx[0] = wrap_dn(c, ng); idx[2] = wrap_up(c, ng);
      h = 0.5 - d;
    }
    else {
      idx[0] = c; idx[1] = wrap_up(c, ng); idx[2] = wrap_up(idx[1], ng);
      d = 1.0 - d;
      h = 0.5 + d;
    }
    w[0] = h * (h * 0.5);
    w[1] = 0.75 - d * d;
    w[2] = 1.0 - w[0] - w[1];
  }
  else {                                // PCS, src/genr_mesh.c:256-272 (units of 1/6)
    idx[1] = c; idx[0] = wrap_dn(c, ng); idx[2] = wrap_up(c, ng);
    idx[3] = wrap_up(idx[2], ng);
    double d2 = d * d;
    w[3] = d2 * d;
    w[2] = 1.0 + 3.0 * (d + d2 - w[3]);
    w[1] = 4.0 - 6.0 * d2 + 3.0 * w[3];
    w[0] = 6.0 - w[1] - w[2] - w[3];
  }
}

// ---------------------------------------------------------------------------
// The same (cell, fraction) without the XU pipe: the division by the (constant) box
// size as a Markstein reciprocal-FMA sequence, floor / int conversion by a magic-number
// add; re-done with the exact IEEE division whenever the fraction is within 1e-9 of a
// value that decides a cell (0, 1/2, 1), so the cell is always the reference's.
// ---------------------------------------------------------------------------
struct AxisXform { double org, ng, len, inv_len; };

__device__ __forceinline__ void split_floor(double t, int &c, double &d) {
  const double MAGIC = 6755399441055744.0;      // 1.5 * 2^52: integer part lands in the low word
  const double tm = __dadd_rn(t, MAGIC);
  c = __double2loint(tm);
  double r = __dsub_rn(tm, MAGIC);
  if (r > t) { r -= 1.0; c -= 1; }
  d = t - r;                                    // exact
}

__device__ __forceinline__ void grid_split(double x, const AxisXform &ax, int &c, double &d) {
  const double a = __dmul_rn(__dsub_rn(x, ax.org), ax.ng);
  // a / len, correctly rounded in all but pathological cases (Markstein)
  const double q0 = a * ax.inv_len;
  const double e = __fma_rn(-q0, ax.len, a);
  double t = __fma_rn(e, ax.inv_len, q0);
  split_floor(t, c, d);
  if (d < 1e-9 || d > 1.0 - 1e-9 || fabs(d - 0.5) < 1e-9) {
    t = __ddiv_rn(a, ax.len);                   // the reference's own arithmetic
    c = (int) t;
    d = t - (double) c;
  }
}

// stencil from (base cell, fraction); same formulas as axis_stencil()
template <int SCHEME>
__device__ __forceinline__ void stencil_from(int c, double d, int ng, int (&idx)[SCHEME + 1],
    double (&w)[SCHEME + 1]) {
  if (c >= ng) c -= ng;         // quirk Q8 guard
  c = min(max(c, 0), ng - 1);   // out-of-box input: stay in bounds, def_box rejects it later
  if constexpr (SCHEME == 0) {
    if (d >= 0.5) c = wrap_up(c, ng);
    idx[0] = c; w[0] = 1.0;
  }
  else if constexpr (SCHEME == 1) {
    idx[0] = c; idx[1] = wrap_up(c, ng);
    w[1] = d; w[0] = 1.0 - d;
  }
  else if constexpr (SCHEME == 2) {
    double h;
    if (d < 0.5) {
      idx[1] = c; idx[0] = wrap_dn(c, ng); idx[2] = wrap_up(c, ng);
      h = 0.5 - d;
    }
    else {
      idx[0] = c; idx[1] = wrap_up(c, ng); idx[2] = wrap_up(idx[1], ng);
      d = 1.0 - d;
      h = 0.5 + d;
    }
    w[0] = h * (h * 0.5);
    w[1] = 0.75 - d * d;
    w[2] = 1.0 - w[0] - w[1];
  }
  else {
    idx[1] = c; idx[0] = wrap_dn(c, ng); idx[2] = wrap_up(c, ng);
    idx[3] = wrap_up(idx[2], ng);
    double d2 = d * d;
    w[3] = d2 * d;
    w[2] = 1.0 + 3.0 * (d + d2 - w[3]);
    w[1] = 4.0 - 6.0 * d2 + 3.0 * w[3];
    w[0] = 6.0 - w[1] - w[2] - w[3];
  }
}

}  // namespace psb
