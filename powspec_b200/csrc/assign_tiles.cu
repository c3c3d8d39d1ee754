// Owner-computes mass assignment for sm_100a (the "cell-sorted batches in shared memory"
// formulation of the assignment loops of src/genr_mesh.c:50-412, 793-858):
//
//   1. k_tile_lists: every particle is appended to the list of every mesh TILE (16 x 16 x 32
//      cells) that the stencils of its fields reach — 1.46 entries per particle for TSC +
//      interlacing — either in ONE pass into fixed-capacity slots (with an overflow list for
//      the rare entry that does not fit, k_tile_overflow) or by count / scan / fill;
//   2. k_tile_accumulate: one block per (tile, field) keeps the tile in shared memory as
//      FIXED-POINT numbers in two 32-bit limbs and adds the in-tile part of each listed
//      particle's stencil with the native shared-memory ATOMS.ADD — the only shared atomic
//      add that is not a compare-and-swap loop on this part (2.4 T adds/s against 0.47 T/s
//      for fp64, tools/smem_atomic_probe.cu);
//   3. the tile is converted back IN PLACE (the two limbs of a cell are the 8 bytes of its
//      double) and written ONCE: double-precision meshes by one TMA tensor store per tile
//      (cp.async.bulk.tensor.3d, UTMASTG; the hardware clips partial tiles), the other
//      cases by 16-byte stores.  No memset of the mesh, no read-modify-write, no global
//      atomics.  DRAM traffic of the stage: particles + lists + one write of the mesh
//      (profiles/README.md).
//
// The lists only have to be SUPERSETS: count / fill use a cheap conservative cell range
// (one multiply per axis, 1e-6 of a cell of slack), the accumulation recomputes every
// stencil with the reference's arithmetic and drops what is outside its tile.
//
// Fixed point: a contribution v = w wx wy wz is scaled by 2^S / max|w| and rounded to an
// integer (S = 43 for tiles of fewer than 256 listed particles, one bit less for every
// doubling: the high limb must hold the sum below 2^30); low limb = 21 bits, so 2048 adds
// fit before the carries are folded (once per 2048 listed particles).  Rounding error per
// contribution <= 2^-(S+1) max|w| (4.5e-13 for a typical tile of BASELINE config 2): far
// inside the 1e-6 budget of P_ell(k), and — integer addition being associative — the
// mesh is bit-for-bit reproducible from run to run, which the global-reduction scatter
// is not.
//
// The cell a particle lands in and its weights come from the same routines as the
// scatter kernels (assign_common.cuh): the reference's arithmetic in the reference's
// order.

#include "assign_common.cuh"

#include <cuda.h>      // CUtensorMap (types only: the encoder is looked up at run time)

#include <algorithm>
#include <cfloat>
#include <cstring>

namespace psb {

namespace {

// tile shape / block shape of the accumulation (macros: ablation builds, profiles/README.md)
#ifndef PSB_TILE_TZ
#define PSB_TILE_TZ 32
#endif
#ifndef PSB_ACC_THREADS
#define PSB_ACC_THREADS 384
#endif
#ifndef PSB_ACC_BLOCKS
#define PSB_ACC_BLOCKS 3
#endif
constexpr int TX = 16, TY = 16, TZ = PSB_TILE_TZ;
constexpr int TCELLS = TX * TY * TZ;
constexpr int LOBITS = 21;
constexpr unsigned LOMASK = (1u << LOBITS) - 1;
constexpr unsigned BATCH = 1u << (32 - LOBITS);
constexpr int ACC_THREADS = PSB_ACC_THREADS, ACC_BLOCKS = PSB_ACC_BLOCKS;

struct TileDims { int ntx, nty, ntz; };

__host__ __device__ inline TileDims tile_dims(int ng) {
  return {(ng + TX - 1) / TX, (ng + TY - 1) / TY, (ng + TZ - 1) / TZ};
}

// A SUPERSET of the (periodic) cell range [lo, lo + cnt) that the stencils of all fields
// reach along one axis.  With t the grid coordinate and c = floor(t) the first cell is
// floor(t + A) - B (NGP: the nearest cell; CIC: c; TSC: the nearest cell - 1; PCS: c - 1,
// src/genr_mesh.c:60-66, 98-108, 157-167, 259-261) and the half-cell shifted field sees
// t + 1/2 (shift_cat, :590-602).  t is taken as (x - org) * (Ng / L), within 1e-11 of the
// reference's rounding of (x - org) * Ng / L; EPS covers that, so the exact stencil of the
// accumulation pass always lies inside the range.
template <int SCHEME, bool INTERLACE>
__device__ __forceinline__ void axis_reach(double x, double org, double scale, int ng, int &lo, int &cnt) {
  constexpr double A = (SCHEME == 0 || SCHEME == 2) ? 0.5 : 0.0;
  constexpr int B = (SCHEME >= 2) ? 1 : 0;
  constexpr double EPS = 1e-6;
  const double t = (x - org) * scale;
  int first = __double2int_rd(t + (A - EPS)) - B;
  const int last = __double2int_rd(t + (A + (INTERLACE ? 0.5 : 0.0) + EPS)) - B + SCHEME;
  cnt = min(max(last - first + 1, 1), SCHEME + 3);
  if (first < 0) first += ng;
  if (first >= ng) first -= ng;
  lo = min(max(first, 0), ng - 1);              // coordinates outside the box: stay in bounds
}

// tiles overlapped along one axis: the first cell's and, if different, the last cell's
// (cnt <= 6 <= cells of every tile, see tile_assign_supported)
__device__ __forceinline__ void axis_tiles(int lo, int cnt, int T, int ng, int &t0, int &t1) {
  int last = lo + cnt - 1;
  if (last >= ng) last -= ng;
  t0 = lo / T;
  t1 = last / T;
  if (t1 == t0) t1 = -1;
}

struct TileSet {
  uint32_t first;       // the tile of the lowest cells: every particle has it
  uint32_t step[3];     // distance to the second tile along x, y, z (0: none)
};

template <int SCHEME, bool INTERLACE>
__device__ __forceinline__ TileSet tile_set(double2 a, double2 b, const AssignGeom &g, const double (&scale)[3],
    const TileDims &td) {
  int lo[3], cnt[3], t0[3], t1[3];
  axis_reach<SCHEME, INTERLACE>(a.x, g.org[0], scale[0], g.ng, lo[0], cnt[0]);
  axis_reach<SCHEME, INTERLACE>(a.y, g.org[1], scale[1], g.ng, lo[1], cnt[1]);
  axis_reach<SCHEME, INTERLACE>(b.x, g.org[2], scale[2], g.ng, lo[2], cnt[2]);
  axis_tiles(lo[0], cnt[0], TX, g.ng, t0[0], t1[0]);
  axis_tiles(lo[1], cnt[1], TY, g.ng, t0[1], t1[1]);
  axis_tiles(lo[2], cnt[2], TZ, g.ng, t0[2], t1[2]);
  TileSet ts;
  ts.first = ((uint32_t) t0[0] * td.nty + t0[1]) * td.ntz + t0[2];
  // unsigned wrap-around differences: first + step is the neighbour's index
  ts.step[0] = t1[0] < 0 ? 0u : (uint32_t) (t1[0] - t0[0]) * (uint32_t) (td.nty * td.ntz);
  ts.step[1] = t1[1] < 0 ? 0u : (uint32_t) (t1[1] - t0[1]) * (uint32_t) td.ntz;
  ts.step[2] = t1[2] < 0 ? 0u : (uint32_t) (t1[2] - t0[2]);
  return ts;
}

// One-pass lists: every tile owns `cap` slots (no count pass, no scan); what does not fit
// goes to a global overflow list of (record, tile) that a small kernel adds to the mesh
// after the tiles have been stored.
struct OnePass {
  int index = 0;                // lists hold 4-byte particle indices instead of the 32-byte records
  uint32_t cap = 0;             // slots per tile; 0: exact lists (count + scan + fill)
  uint32_t ovcap = 0;
  uint32_t *ovcount = nullptr;
  double2 *ovrec = nullptr;
  uint32_t *ovtile = nullptr;
};

__device__ __forceinline__ void put_entry(const OnePass &op, double2 *__restrict__ out, uint32_t tile, uint32_t pos,
    double2 a, double2 b, uint32_t index) {
  if (op.cap == 0u || pos < op.cap) {
    const size_t slot = op.cap ? (size_t) tile * op.cap + pos : (size_t) pos;
    if (op.index) reinterpret_cast<uint32_t *>(out)[slot] = index;
    else st_record(out, slot, a, b);
    return;
  }
  const uint32_t q = atomicAdd(op.ovcount, 1u);
  if (q < op.ovcap) { st_record(op.ovrec, q, a, b); op.ovtile[q] = tile; }
}

// the up to seven further tiles of a particle that straddles tile faces
template <bool FILL>
__device__ __forceinline__ void tile_extra(const TileSet &ts, uint32_t *__restrict__ cnt_or_cursor,
    double2 *__restrict__ out, double2 a, double2 b, const OnePass &op, uint32_t index) {
#pragma unroll
  for (int m = 1; m < 8; m++) {
    if (((m & 1) && !ts.step[2]) || ((m & 2) && !ts.step[1]) || ((m & 4) && !ts.step[0])) continue;
    const uint32_t tile = ts.first + ((m & 1) ? ts.step[2] : 0u) + ((m & 2) ? ts.step[1] : 0u) +
        ((m & 4) ? ts.step[0] : 0u);
    const uint32_t pos = atomicAdd(cnt_or_cursor + tile, 1u);
    if (FILL) put_entry(op, out, tile, pos, a, b, index);
  }
}

// FILL = false: count the list lengths; FILL = true: write the records (exact lists: into
// the slots the scan of the counts assigned; one-pass lists: into the tile's own slots,
// counting as it goes).  The coordinate bounds and the largest |weight| of the block's
// particles are reduced by whichever pass comes first (STATS).  The fill pass waits for the
// returned list positions: UNROLL particles per thread are in flight.
template <int SCHEME, bool INTERLACE, bool FILL, int UNROLL, bool STATS>
__global__ void __launch_bounds__(256) k_tile_lists(const double2 *__restrict__ p, size_t n, AssignGeom g,
    uint32_t *__restrict__ cnt_or_cursor, double2 *__restrict__ out, double *__restrict__ partials,
    double *__restrict__ wmax_part, OnePass op) {
  const TileDims td = tile_dims(g.ng);
  const double scale[3] = {(double) g.ng / g.len[0], (double) g.ng / g.len[1], (double) g.ng / g.len[2]};
  double lo3[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi3[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, wm = 0.0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i0 = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i0 < n; i0 += stride * UNROLL) {
    double2 a[UNROLL], b[UNROLL];
    TileSet ts[UNROLL];
    uint32_t pos[UNROLL];
    bool more = false;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t i = i0 + u * stride;
      if (i < n) ld_record(p, i, a[u], b[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (i0 + u * stride >= n) continue;
      if (STATS) {
        lo3[0] = fmin(lo3[0], a[u].x); hi3[0] = fmax(hi3[0], a[u].x);
        lo3[1] = fmin(lo3[1], a[u].y); hi3[1] = fmax(hi3[1], a[u].y);
        lo3[2] = fmin(lo3[2], b[u].x); hi3[2] = fmax(hi3[2], b[u].x);
        wm = fmax(wm, fabs(b[u].y));
      }
      ts[u] = tile_set<SCHEME, INTERLACE>(a[u], b[u], g, scale, td);
      pos[u] = atomicAdd(cnt_or_cursor + ts[u].first, 1u);
      more |= (ts[u].step[0] | ts[u].step[1] | ts[u].step[2]) != 0u;
    }
    if (FILL) {
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        if (i0 + u * stride < n) put_entry(op, out, ts[u].first, pos[u], a[u], b[u], (uint32_t) (i0 + u * stride));
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        if (i0 + u * stride < n) tile_extra<FILL>(ts[u], cnt_or_cursor, out, a[u], b[u], op, (uint32_t) (i0 + u * stride));
    }
  }
  if (!STATS) return;
  __shared__ double s[7][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo3[a] = fmin(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
      hi3[a] = fmax(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
    }
    if (lane == 0) { s[a][warp] = lo3[a]; s[3 + a][warp] = hi3[a]; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
  if (lane == 0) s[6][warp] = wm;
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = s[threadIdx.x][0];
    for (int w = 1; w < 8; w++)
      v = (threadIdx.x < 3) ? fmin(v, s[threadIdx.x][w]) : fmax(v, s[threadIdx.x][w]);
    if (threadIdx.x < 6) { if (partials) partials[blockIdx.x * 6 + threadIdx.x] = v; }
    else wmax_part[blockIdx.x] = v;
  }
}

__global__ void k_wmax_reduce(const double *__restrict__ part, int nblk, double *__restrict__ wmax) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 32) v = fmax(v, part[i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (threadIdx.x == 0) *wmax = v;
}

// Shared-memory cell: {low limb, high limb}, 8 bytes — the place of the cell's double in
// the flushed tile.
//
// round(v) (|v| < 2^51) by the 1.5 * 2^52 magic add: the mantissa field is 2^51 + round(v),
// so the low limb is its low LOBITS bits and the high limb (the signed rest) is bits
// 21..52 of the pattern, whose two top bits (exponent LSB, the 2^51 offset) are cancelled
// by adding 2^30 modulo 2^32.
__device__ __forceinline__ void split_fixed(double v, uint32_t &lo, uint32_t &hi) {
  const double m = v + 6755399441055744.0;
  const uint32_t lw = (uint32_t) __double2loint(m), hw = (uint32_t) __double2hiint(m);
  lo = lw & LOMASK;
  hi = __funnelshift_r(lw, hw, LOBITS) + 0x40000000u;
}

// the inverse: hi * 2^LOBITS + lo as a double, |value| < 2^51, without the conversion pipe
__device__ __forceinline__ double join_fixed(uint32_t lo, uint32_t hi) {
  const long long x = (long long) (int) hi * (long long) (1u << LOBITS) + (long long) lo;
  return __longlong_as_double(x + 0x4338000000000000ll) - 6755399441055744.0;
}

// Which half of a cell's 8 bytes holds the low limb alternates with the row (y) parity:
// a row is 96 words = 0 modulo the 32 banks, so without the swap every low-limb add of a
// warp would fall on the 16 even banks and every high-limb add on the 16 odd ones (measured:
// 1.45x the bank conflicts of the layout with two separate limb arrays).
__device__ __forceinline__ uint32_t limb_swap(uint32_t row) { return row & 1u; }

// one row of the stencil, all z cells inside the tile (consecutive): no predicates
template <int OFF>
__device__ __forceinline__ void red_pair(uint32_t alo, uint32_t ahi, uint32_t lo, uint32_t hi) {
  asm volatile("red.shared.add.u32 [%0+%4], %2;\n\tred.shared.add.u32 [%1+%4], %3;"
      :: "r"(alo), "r"(ahi), "r"(lo), "r"(hi), "n"(OFF) : "memory");
}
template <int C, int NS>
__device__ __forceinline__ void row_inside(uint32_t alo, uint32_t ahi, double wxy, const double (&wz)[NS]) {
  if constexpr (C < NS) {
    uint32_t lo, hi;
    split_fixed(wxy * wz[C], lo, hi);
    red_pair<8 * C>(alo, ahi, lo, hi);
    row_inside<C + 1, NS>(alo, ahi, wxy, wz);
  }
}

// does the fraction decide nothing?  (grid_split's test, two comparisons folded into one)
__device__ __forceinline__ bool frac_is_safe(double d) {
  return fabs(fabs(d - 0.5) - 0.25) < 0.25 - 1e-9;
}

// grid_split (assign_common.cuh) with the three comparisons on the fraction folded
__device__ __forceinline__ void grid_split2(double x, const AxisXform &ax, int &c, double &d) {
  const double a = __dmul_rn(__dsub_rn(x, ax.org), ax.ng);
  const double q0 = a * ax.inv_len;
  const double e = __fma_rn(-q0, ax.len, a);
  double t = __fma_rn(e, ax.inv_len, q0);
  split_floor(t, c, d);
  if (!frac_is_safe(d)) {
    t = __ddiv_rn(a, ax.len);                   // the reference's own arithmetic
    c = (int) t;
    d = t - (double) c;
  }
}

// Adds the in-tile part of one listed particle's stencil.  PARTIAL = false: the common
// case, all z cells inside the tile, straight-line code; returns true (nothing added) if
// only some of them are, and the caller queues the particle for a PARTIAL = true pass —
// so that the predicated (branchy) code runs on full warps of such particles instead of
// being dragged through every warp by one lane.
template <int SCHEME, bool PARTIAL>
__device__ __forceinline__ bool tile_add(const double2 *__restrict__ parts, size_t idx, int f,
    const AssignGeom &g, double wnorm, int x0, int y0, int z0, uint32_t sm_lo) {
  constexpr int NS = SCHEME + 1;
  double2 a, b;
  ld_record(parts, idx, a, b);
  double x[3] = {a.x, a.y, b.x};
  const double *org = f ? g.sorg : g.org;
  if (f) {
    // shift_cat, src/genr_mesh.c:595-600: periodic wrap into the shifted box
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (x[k] >= __dadd_rn(g.sorg[k], g.len[k])) x[k] = __dsub_rn(x[k], g.len[k]);
  }
  int ix[NS], iy[NS], iz[NS], cc;
  double wx[NS], wy[NS], wz[NS], dd;
  const double ngd = (double) g.ng;
  grid_split2(x[2], AxisXform{org[2], ngd, g.len[2], g.inv_len[2]}, cc, dd);
  stencil_from<SCHEME>(cc, dd, g.ng, iz, wz);
  bool zin[NS];
  uint32_t lz[NS];
  int nin = 0;
#pragma unroll
  for (int c = 0; c < NS; c++) {
    lz[c] = (uint32_t) (iz[c] - z0);
    zin[c] = lz[c] < (uint32_t) TZ;
    nin += zin[c];
  }
  if (nin == 0) return false;
  // the straight-line path takes the z cells as consecutive: all inside the tile and no
  // periodic wrap among them (a tile that spans the whole axis holds both ends)
  if (!PARTIAL && (nin < NS || lz[NS - 1] != lz[0] + (uint32_t) (NS - 1))) return true;
  grid_split2(x[0], AxisXform{org[0], ngd, g.len[0], g.inv_len[0]}, cc, dd);
  stencil_from<SCHEME>(cc, dd, g.ng, ix, wx);
  grid_split2(x[1], AxisXform{org[1], ngd, g.len[1], g.inv_len[1]}, cc, dd);
  stencil_from<SCHEME>(cc, dd, g.ng, iy, wy);
  // the particle weight enters through the x weights (src/genr_mesh.c:110-111, 175-177);
  // wnorm = catalogue scale x fixed-point scale (x 1/216 for PCS, :274-278)
  const double pw = b.y * wnorm;
#pragma unroll
  for (int q = 0; q < NS; q++) wx[q] *= pw;
#pragma unroll
  for (int u = 0; u < NS; u++) {
    const uint32_t lx = (uint32_t) (ix[u] - x0);
    if (lx >= (uint32_t) TX) continue;
    const uint32_t xrow = sm_lo + lx * (uint32_t) (TY * TZ * 8) + (PARTIAL ? 0u : lz[0] * 8u);
#pragma unroll
    for (int v = 0; v < NS; v++) {
      const uint32_t ly = (uint32_t) (iy[v] - y0);
      if (ly >= (uint32_t) TY) continue;
      const double wxy = wx[u] * wy[v];
      const uint32_t swap = limb_swap(ly) * 4u;
      const uint32_t alo = xrow + ly * (uint32_t) (TZ * 8) + swap, ahi = alo + 4u - 2u * swap;
      if constexpr (!PARTIAL) row_inside<0, NS>(alo, ahi, wxy, wz);
      else {
#pragma unroll
        for (int c = 0; c < NS; c++) {
          if (!zin[c]) continue;
          uint32_t lo, hi;
          split_fixed(wxy * wz[c], lo, hi);
          red_pair<0>(alo + lz[c] * 8u, ahi + lz[c] * 8u, lo, hi);
        }
      }
    }
  }
  return false;
}

struct TileMaps { CUtensorMap m[2]; };

// MODE 0: the tile is stored (the mesh need not be initialised); 1: added to the mesh.
// TMA: the flush of a double-precision tile is one bulk tensor store.
template <int SCHEME, typename real, int NFIELD, int MODE, bool TMA>
__global__ void __launch_bounds__(ACC_THREADS, ACC_BLOCKS) k_tile_accumulate(const double2 *__restrict__ parts,
    const uint32_t *__restrict__ start, AssignGeom g, double wscale, const double *__restrict__ wmax_dev,
    real *__restrict__ mesh0, real *__restrict__ mesh1, const __grid_constant__ TileMaps maps, uint32_t cap,
    const uint32_t *__restrict__ ilist) {
  extern __shared__ __align__(1024) uint32_t sm[];      // cell c: words 2c, 2c + 1 = its limbs (limb_swap)
  __shared__ uint16_t queue[BATCH];             // listed particles (index in the batch) that straddle the tile's z faces
  __shared__ uint32_t nqueue;
  const uint32_t sm_lo = (uint32_t) __cvta_generic_to_shared(sm);
  const TileDims td = tile_dims(g.ng);
  const uint32_t ntile = (uint32_t) (td.ntx * td.nty * td.ntz);
  // |w wscale wx wy wz| <= wmax |wscale| for every scheme (PCS: weights in units of 1/6,
  // times 1/216)
  const double wbound = *wmax_dev * fabs(wscale);
  for (int i = threadIdx.x; i < TCELLS / 2; i += ACC_THREADS) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) nqueue = 0;
  __syncthreads();
  for (uint32_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int tz = tile % td.ntz, ty = (tile / td.ntz) % td.nty, tx = tile / (td.ntz * td.nty);
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    // exact lists: offsets from the scan; one-pass lists: `cap` slots per tile, start[] = lengths
    const uint32_t b0 = cap ? 0u : start[tile], np = cap ? min(start[tile], cap) : start[tile + 1] - b0;
    const size_t lbase = cap ? (size_t) tile * cap : (size_t) b0;
    // headroom: a cell receives at most one contribution per listed particle; the high
    // limb stays below 2^30 for 2^30 / 2^(S - LOBITS) contributions of the largest size
    int S = 43;
    for (uint32_t q = 1u << (30 - (43 - LOBITS)); q != 0u && q <= np; q <<= 1) S--;
    const double norm = wbound > 0.0 ? ldexp(1.0, S) / wbound : 0.0;
    const double inv = wbound > 0.0 ? wbound * ldexp(1.0, -S) : 0.0;
    const double wnorm = (SCHEME == 3 ? wscale * 0x1.2f684bda12f68p-8 : wscale) * norm;
#pragma unroll 1
    for (int f = 0; f < NFIELD; f++) {
      for (uint32_t base = 0; base < np; base += BATCH) {
        const uint32_t lim = min(np, base + BATCH);
        for (uint32_t j = base + threadIdx.x; j < lim; j += ACC_THREADS)
          if (tile_add<SCHEME, false>(parts, ilist ? (size_t) __ldg(ilist + lbase + j) : lbase + j, f, g, wnorm, x0, y0, z0, sm_lo))
            queue[atomicAdd(&nqueue, 1u)] = (uint16_t) (j - base);
        __syncthreads();
        const uint32_t nq = nqueue;
        for (uint32_t k = threadIdx.x; k < nq; k += ACC_THREADS)
          tile_add<SCHEME, true>(parts, ilist ? (size_t) __ldg(ilist + lbase + base + queue[k]) : lbase + base + queue[k], f, g, wnorm, x0, y0, z0, sm_lo);
        __syncthreads();
        if (threadIdx.x == 0) nqueue = 0;
        if (lim < np) {
          // fold the carries so that neither limb can wrap in the next batch
          for (int i = threadIdx.x; i < TCELLS; i += ACC_THREADS) {
            const uint32_t sw = limb_swap((uint32_t) (i / TZ));
            const uint32_t lo = sm[2 * i + sw];
            sm[2 * i + sw] = lo & LOMASK;
            sm[2 * i + 1 - sw] += lo >> LOBITS;
          }
          __syncthreads();
        }
      }
      // flush: every cell of the tile goes to the mesh once; the tile is left zeroed for
      // the next work item
      real *m = f ? mesh1 : mesh0;
      if constexpr (TMA) {
        for (int i = threadIdx.x; i < TCELLS / 2; i += ACC_THREADS) {
          const uint4 q = reinterpret_cast<const uint4 *>(sm)[i];
          const bool sw = limb_swap((uint32_t) (i / (TZ / 2))) != 0u;
          reinterpret_cast<double2 *>(sm)[i] = make_double2(join_fixed(sw ? q.y : q.x, sw ? q.x : q.y) * inv,
              join_fixed(sw ? q.w : q.z, sw ? q.z : q.w) * inv);
        }
        // make the generic-proxy writes visible to the TMA engine, then one thread stores
        // the box (z fastest); cells beyond the mesh are clipped by the tensor map
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
              :: "l"(&maps.m[f]), "r"(z0), "r"(y0), "r"(x0), "r"(sm_lo) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TCELLS / 2; i += ACC_THREADS)
          reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      else {
        // two z-neighbours per thread: (row, pair) advance by ACC_THREADS pairs per step
        constexpr int PAIRS = TZ / 2;
        int row = threadIdx.x / PAIRS, q = threadIdx.x % PAIRS;
        for (int i = threadIdx.x; i < TCELLS / 2; i += ACC_THREADS) {
          const uint4 c = reinterpret_cast<const uint4 *>(sm)[i];
          reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
          const int gx = x0 + (row / TY), gy = y0 + (row % TY), gz = z0 + 2 * q;
          if (gx < g.ng && gy < g.ng && gz < g.ng) {
            const bool sw = limb_swap((uint32_t) row) != 0u;
            const double v0 = join_fixed(sw ? c.y : c.x, sw ? c.x : c.y) * inv;
            const double v1 = join_fixed(sw ? c.w : c.z, sw ? c.z : c.w) * inv;
            real *cell = m + ((size_t) gx * g.ng + gy) * g.rowlen + gz;
            if (gz + 1 < g.ng) {
              // rowlen and gz are even: the pair is aligned
              if constexpr (sizeof(real) == 8) {
                double2 *c2 = reinterpret_cast<double2 *>(cell);
                if (MODE == 0) *c2 = make_double2(v0, v1);
                else { double2 o = *c2; o.x += v0; o.y += v1; *c2 = o; }
              }
              else {
                float2 *c2 = reinterpret_cast<float2 *>(cell);
                if (MODE == 0) *c2 = make_float2((float) v0, (float) v1);
                else { float2 o = *c2; o.x += (float) v0; o.y += (float) v1; *c2 = o; }
              }
            }
            else {
              if (MODE == 0) *cell = (real) v0;
              else *cell += (real) v0;
            }
          }
          q += ACC_THREADS % PAIRS;
          row += ACC_THREADS / PAIRS;
          if (q >= PAIRS) { q -= PAIRS; row++; }
        }
      }
      __syncthreads();
    }
  }
}

int g_fill_unroll = 2;  // particles in flight per thread of the fill pass (psb_set_option "tile_fill_unroll"): the
                        // one-pass fill with 4 needs 100 registers (2 blocks per SM): list stage 5.6 vs 4.6 ms with 2 or 1

// One-pass lists: the entries that did not fit their tile's slots.  One thread per
// (particle, tile) entry adds the in-tile part of the particle's stencils to the mesh with
// global atomics — after the accumulation has stored (or added) the tiles.
template <int SCHEME, typename real, int NFIELD>
__global__ void __launch_bounds__(256) k_tile_overflow(const double2 *__restrict__ rec,
    const uint32_t *__restrict__ tiles, uint32_t n, AssignGeom g, double wscale, real *__restrict__ mesh0,
    real *__restrict__ mesh1) {
  constexpr int NS = SCHEME + 1;
  const TileDims td = tile_dims(g.ng);
  const double ngd = (double) g.ng;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double2 a, b;
    ld_record(rec, i, a, b);
    const uint32_t tile = tiles[i];
    const int x0 = (int) (tile / (uint32_t) (td.ntz * td.nty)) * TX, y0 = (int) ((tile / td.ntz) % td.nty) * TY,
        z0 = (int) (tile % td.ntz) * TZ;
    double pw = b.y * wscale;
    if constexpr (SCHEME == 3) pw *= 0x1.2f684bda12f68p-8;
#pragma unroll 1
    for (int f = 0; f < NFIELD; f++) {
      double x[3] = {a.x, a.y, b.x};
      const double *org = f ? g.sorg : g.org;
      if (f) {
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (x[k] >= __dadd_rn(g.sorg[k], g.len[k])) x[k] = __dsub_rn(x[k], g.len[k]);
      }
      int ix[NS], iy[NS], iz[NS], cc;
      double wx[NS], wy[NS], wz[NS], dd;
      grid_split2(x[0], AxisXform{org[0], ngd, g.len[0], g.inv_len[0]}, cc, dd);
      stencil_from<SCHEME>(cc, dd, g.ng, ix, wx);
      grid_split2(x[1], AxisXform{org[1], ngd, g.len[1], g.inv_len[1]}, cc, dd);
      stencil_from<SCHEME>(cc, dd, g.ng, iy, wy);
      grid_split2(x[2], AxisXform{org[2], ngd, g.len[2], g.inv_len[2]}, cc, dd);
      stencil_from<SCHEME>(cc, dd, g.ng, iz, wz);
      real *m = f ? mesh1 : mesh0;
#pragma unroll
      for (int u = 0; u < NS; u++) {
        if ((uint32_t) (ix[u] - x0) >= (uint32_t) TX) continue;
        const double wxu = wx[u] * pw;
#pragma unroll
        for (int v = 0; v < NS; v++) {
          if ((uint32_t) (iy[v] - y0) >= (uint32_t) TY) continue;
          const double wxy = wxu * wy[v];
          real *row = m + ((size_t) ix[u] * g.ng + iy[v]) * g.rowlen;
#pragma unroll
          for (int c = 0; c < NS; c++)
            if ((uint32_t) (iz[c] - z0) < (uint32_t) TZ) atomicAdd(row + iz[c], (real) (wxy * wz[c]));
        }
      }
    }
  }
}

// mode 0: count (exact lists), 1: fill the slots the scan assigned, 2: one-pass fill
template <int SCHEME, bool INTERLACE>
int launch_lists(const double *p, size_t n, const AssignGeom &g, int mode, uint32_t *cnt, double *out,
    double *partials, double *wmax_part, const OnePass &op, cudaStream_t st) {
  const double2 *pp = reinterpret_cast<const double2 *>(p);
  const int nblk = row_keys_blocks(n);
  double2 *o = reinterpret_cast<double2 *>(out);
  if (mode == 2) {
    if (g_fill_unroll >= 4) k_tile_lists<SCHEME, INTERLACE, true, 4, true><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, partials, wmax_part, op);
    else if (g_fill_unroll >= 2) k_tile_lists<SCHEME, INTERLACE, true, 2, true><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, partials, wmax_part, op);
    else k_tile_lists<SCHEME, INTERLACE, true, 1, true><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, partials, wmax_part, op);
  }
  else if (mode == 1) {
    if (g_fill_unroll >= 4) k_tile_lists<SCHEME, INTERLACE, true, 4, false><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, nullptr, nullptr, op);
    else if (g_fill_unroll >= 2) k_tile_lists<SCHEME, INTERLACE, true, 2, false><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, nullptr, nullptr, op);
    else k_tile_lists<SCHEME, INTERLACE, true, 1, false><<<nblk, 256, 0, st>>>(pp, n, g, cnt, o, nullptr, nullptr, op);
  }
  else
    k_tile_lists<SCHEME, INTERLACE, false, 2, true><<<nblk, 256, 0, st>>>(pp, n, g, cnt, nullptr, partials, wmax_part, op);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda)
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled tensor_map_encoder() {
  static EncodeTiled fn = []() -> EncodeTiled {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiled>(p);
  }();
  return fn;
}

// the mesh as a (x, y, z) tensor of doubles whose z extent is the Ng cells of a row (not
// its padding), box = one tile
bool tile_tensor_map(CUtensorMap *tm, void *mesh, const AssignGeom &g) {
  EncodeTiled enc = tensor_map_encoder();
  if (!enc || !mesh || (reinterpret_cast<uintptr_t>(mesh) & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t) g.ng, (cuuint64_t) g.ng, (cuuint64_t) g.ng};
  const cuuint64_t strides[2] = {(cuuint64_t) g.rowlen * 8, (cuuint64_t) g.ng * g.rowlen * 8};
  const cuuint32_t box[3] = {TZ, TY, TX}, estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, mesh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_tile_tma = 1;     // ablation switch (psb_set_option "tile_tma")

template <int SCHEME, typename real>
int launch_accumulate(const double *parts, const uint32_t *start, const AssignGeom &g, double wscale,
    const double *wmax, bool add, void *m0, void *m1, uint32_t cap, const uint32_t *ilist, cudaStream_t st) {
  const size_t smem = (size_t) TCELLS * 8;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const double2 *pp = reinterpret_cast<const double2 *>(parts);
  real *a = static_cast<real *>(m0), *b = static_cast<real *>(m1);
  TileMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool tma = false;
  if constexpr (sizeof(real) == 8)
    tma = g_tile_tma && !add && tile_tensor_map(&maps.m[0], m0, g) && (!m1 || tile_tensor_map(&maps.m[1], m1, g));
#define PSB_ACC(NF, MODE, TMA)                                                                     \
  do {                                                                                            \
    auto kern = k_tile_accumulate<SCHEME, real, NF, MODE, TMA>;                                   \
    PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
    kern<<<ACC_BLOCKS * sms, ACC_THREADS, smem, st>>>(pp, start, g, wscale, wmax, a, b, maps, cap, ilist);              \
  } while (0)
  if constexpr (sizeof(real) == 8) {
    if (tma) { if (m1) PSB_ACC(2, 0, true); else PSB_ACC(1, 0, true); }
  }
  if (!tma) {
    if (m1) { if (add) PSB_ACC(2, 1, false); else PSB_ACC(2, 0, false); }
    else { if (add) PSB_ACC(1, 1, false); else PSB_ACC(1, 0, false); }
  }
#undef PSB_ACC
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

void tile_set_tma(int on) { g_tile_tma = on; }
void tile_set_fill_unroll(int u) { g_fill_unroll = u; }

// the whole mesh on this device, and tiles that a stencil range (<= 6 cells) can straddle
// in at most two pieces
bool tile_assign_supported(const AssignGeom &g) {
  if (g.nxloc != g.ng || g.nx != g.ng || g.x0 != 0 || g.xbase != 0 || g.ng < 16) return false;
  const int T[3] = {TX, TY, TZ};
  for (int a = 0; a < 3; a++) {
    const int r = g.ng % T[a];
    if (r != 0 && r < 6) return false;
  }
  return true;
}

// Slots per tile of the one-pass lists: 1.5 x the mean list length of a uniform catalogue
// (every axis: a stencil range of r cells straddles a tile face with probability (r-1)/T)
// + 64.  A Poisson tile of BASELINE config 2 (mean 1114) is 18 sigma below it; catalogues
// clustered on the tile scale overflow, and the caller falls back to the exact lists.
uint32_t tile_list_capacity(const AssignGeom &g, size_t n, int scheme, bool interlace) {
  const double r = scheme + 1 + (interlace ? 1 : 0);
  const TileDims td = tile_dims(g.ng);
  const double dup = (1 + (r - 1) / TX) * (1 + (r - 1) / TY) * (1 + (r - 1) / TZ);
  const double mean = (double) n * dup / ((double) td.ntx * td.nty * td.ntz);
  return (uint32_t) std::min(4.0e9, 1.5 * mean + 64.0) & ~3u;
}

size_t tile_list_count(const AssignGeom &g) {
  const TileDims td = tile_dims(g.ng);
  return (size_t) td.ntx * td.nty * td.ntz;
}

// cnt: tile_list_count + 1 counters (zeroed by the caller; the extra one stays 0 so that
// an exclusive scan leaves the total in it); partials: optional [row_keys_blocks(n)][6]
// coordinate bounds; wmax_part: [row_keys_blocks(n)]; wmax: the reduced max |w|
int launch_tile_count(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cnt, double *partials, double *wmax_part, double *wmax, cudaStream_t st) {
#define PSB_LISTS(S)                                                                              \
  case S:                                                                                         \
    if (interlace ? launch_lists<S, true>(p, n, g, 0, cnt, nullptr, partials, wmax_part, OnePass{}, st)  \
                  : launch_lists<S, false>(p, n, g, 0, cnt, nullptr, partials, wmax_part, OnePass{}, st)) \
      return -1;                                                                                  \
    break;
  switch (scheme) { PSB_LISTS(0) PSB_LISTS(1) PSB_LISTS(2) PSB_LISTS(3) default: return -1; }
#undef PSB_LISTS
  k_wmax_reduce<<<1, 32, 0, st>>>(wmax_part, row_keys_blocks(n), wmax);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// cursor: the exclusive scan of the counts (consumed); lists: records, grouped by tile
int launch_tile_fill(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cursor, void *lists, bool index, cudaStream_t st) {
  OnePass op;
  op.index = index;
  double *out = static_cast<double *>(lists);
#define PSB_LISTS(S)                                                                              \
  case S:                                                                                         \
    return interlace ? launch_lists<S, true>(p, n, g, 1, cursor, out, nullptr, nullptr, op, st)  \
                     : launch_lists<S, false>(p, n, g, 1, cursor, out, nullptr, nullptr, op, st);
  switch (scheme) { PSB_LISTS(0) PSB_LISTS(1) PSB_LISTS(2) PSB_LISTS(3) default: return -1; }
#undef PSB_LISTS
}

// One-pass lists (no count pass, no scan): cnt = tile_list_count zeroed counters (left holding
// the list lengths, possibly beyond the capacity), lists = tile_list_count * op.cap records;
// *op.ovcount (zeroed by the caller) counts the entries that did not fit.  Also reduces the
// coordinate bounds (optional) and max |w| like launch_tile_count.
int launch_tile_fill_onepass(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cnt, void *lists_, const TileOnePass &t, double *partials, double *wmax_part, double *wmax,
    cudaStream_t st) {
  OnePass op;
  double *lists = static_cast<double *>(lists_);
  op.index = t.index;
  op.cap = t.cap; op.ovcap = t.ovcap; op.ovcount = t.ovcount;
  op.ovrec = reinterpret_cast<double2 *>(t.ovrec); op.ovtile = t.ovtile;
#define PSB_LISTS(S)                                                                              \
  case S:                                                                                         \
    if (interlace ? launch_lists<S, true>(p, n, g, 2, cnt, lists, partials, wmax_part, op, st)    \
                  : launch_lists<S, false>(p, n, g, 2, cnt, lists, partials, wmax_part, op, st))  \
      return -1;                                                                                  \
    break;
  switch (scheme) { PSB_LISTS(0) PSB_LISTS(1) PSB_LISTS(2) PSB_LISTS(3) default: return -1; }
#undef PSB_LISTS
  k_wmax_reduce<<<1, 32, 0, st>>>(wmax_part, row_keys_blocks(n), wmax);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_tile_overflow(const double *ovrec, const uint32_t *ovtile, uint32_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *mesh0, void *mesh1, cudaStream_t st) {
  if (!n) return 0;
  const double2 *r = reinterpret_cast<const double2 *>(ovrec);
  const int grid = (int) std::min<uint32_t>((n + 255) / 256, 148 * 8);
#define PSB_OV(S, T)                                                                              \
  do {                                                                                            \
    if (mesh1) k_tile_overflow<S, T, 2><<<grid, 256, 0, st>>>(r, ovtile, n, g, wscale, (T *) mesh0, (T *) mesh1); \
    else k_tile_overflow<S, T, 1><<<grid, 256, 0, st>>>(r, ovtile, n, g, wscale, (T *) mesh0, (T *) nullptr);    \
  } while (0)
#define PSB_OVS(S) case S: if (precision == 8) PSB_OV(S, double); else PSB_OV(S, float); break;
  switch (scheme) { PSB_OVS(0) PSB_OVS(1) PSB_OVS(2) PSB_OVS(3) default: return -1; }
#undef PSB_OVS
#undef PSB_OV
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// start: tile_list_count + 1 offsets into lists; add: the meshes already hold data
int launch_tile_accumulate(const double *lists, const uint32_t *start, const AssignGeom &g, int scheme,
    int precision, double wscale, const double *wmax, bool add, void *mesh0, void *mesh1, cudaStream_t st,
    uint32_t cap, const uint32_t *ilist) {
#define PSB_DISPATCH(S)                                                                           \
  case S:                                                                                         \
    return precision == 8                                                                         \
        ? launch_accumulate<S, double>(lists, start, g, wscale, wmax, add, mesh0, mesh1, cap, ilist, st) \
        : launch_accumulate<S, float>(lists, start, g, wscale, wmax, add, mesh0, mesh1, cap, ilist, st);
  switch (scheme) {
    PSB_DISPATCH(0) PSB_DISPATCH(1) PSB_DISPATCH(2) PSB_DISPATCH(3)
    default: set_error("unrecognised particle assignment scheme: %d\n", scheme); return -1;
  }
#undef PSB_DISPATCH
}

}  // namespace psb
