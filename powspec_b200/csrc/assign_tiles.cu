// Owner-computes mass assignment for sm_100a (the "cell-sorted batches in shared memory"
// formulation of the assignment loops of src/genr_mesh.c:50-412, 793-858):
//
//   1. k_tile_lists<count> / scan / k_tile_lists<fill>: every particle is appended to the
//      list of every mesh TILE (16 x 16 x 48 cells) that the stencils of its fields reach
//      — 1.41 entries per particle for TSC + interlacing;
//   2. k_tile_accumulate: one block per (tile, field) keeps the tile in shared memory as
//      FIXED-POINT numbers in two 32-bit limbs and adds the in-tile part of each listed
//      particle's stencil with the native shared-memory ATOMS.ADD — the only shared atomic
//      add that is not a compare-and-swap loop on this part (2.4 T adds/s against 0.47 T/s
//      for fp64, tools/smem_atomic_probe.cu);
//   3. the tile is converted back and written ONCE with plain coalesced stores: no memset
//      of the mesh, no read-modify-write, no global atomics.  DRAM traffic of the stage:
//      particles + lists + one write of the mesh (profiles/README.md).
//
// Fixed point: a contribution v = w wx wy wz is scaled by 2^S / max|w| and rounded to an
// integer (S = 44 for tiles of up to 256 listed particles, one bit less for every
// doubling: the high limb must hold the sum); low limb = 21 bits, so 2048 adds fit before
// the carries are folded (once per 2048 listed particles).  Rounding error per
// contribution <= 2^-(S+1) max|w| (2.3e-13 for a typical tile of BASELINE config 2): far
// inside the 1e-6 budget of P_ell(k), and — integer addition being associative — the
// mesh is bit-for-bit reproducible from run to run, which the global-reduction scatter
// is not.
//
// The cell a particle lands in and its weights come from the same routines as the
// scatter kernels (assign_common.cuh): the reference's arithmetic in the reference's
// order.

#include "assign_common.cuh"

#include <cfloat>

namespace psb {

namespace {

constexpr int TX = 16, TY = 16, TZ = 48;
constexpr int TCELLS = TX * TY * TZ;
constexpr int LOBITS = 21;
constexpr unsigned LOMASK = (1u << LOBITS) - 1;
constexpr unsigned BATCH = 1u << (32 - LOBITS);
constexpr int ACC_THREADS = 512;

struct TileDims { int ntx, nty, ntz; };

__host__ __device__ inline TileDims tile_dims(int ng) {
  return {(ng + TX - 1) / TX, (ng + TY - 1) / TY, (ng + TZ - 1) / TZ};
}

// the (periodic) cell range [lo, lo + cnt) that the stencils of all fields reach along
// one axis
template <int SCHEME, bool INTERLACE>
__device__ __forceinline__ void axis_reach(double x, double org, double sorg, double len, double inv_len,
    int ng, int &lo, int &cnt) {
  constexpr int NS = SCHEME + 1;
  int i0[NS], c;
  double w0[NS], d;
  grid_split(x, AxisXform{org, (double) ng, len, inv_len}, c, d);
  stencil_from<SCHEME>(c, d, ng, i0, w0);
  lo = i0[0];
  cnt = NS;
  if constexpr (INTERLACE) {
    // shift_cat, src/genr_mesh.c:595-600
    if (x >= __dadd_rn(sorg, len)) x = __dsub_rn(x, len);
    int i1[NS];
    grid_split(x, AxisXform{sorg, (double) ng, len, inv_len}, c, d);
    stencil_from<SCHEME>(c, d, ng, i1, w0);
    int dd = i1[0] - lo;
    if (dd < 0) dd += ng;
    if (dd <= NS) cnt = dd + NS;                // the shifted stencil starts dd cells higher
    else { lo = i1[0]; cnt = (ng - dd) + NS; }  // ... or ng - dd cells lower
  }
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

// FILL = false: count the list lengths (and, optionally, the coordinate bounds and the
// largest |weight| of the block's particles); FILL = true: write the records
template <int SCHEME, bool INTERLACE, bool FILL>
__global__ void __launch_bounds__(256) k_tile_lists(const double2 *__restrict__ p, size_t n, AssignGeom g,
    uint32_t *__restrict__ cnt_or_cursor, double2 *__restrict__ out, double *__restrict__ partials,
    double *__restrict__ wmax_part) {
  const TileDims td = tile_dims(g.ng);
  double lo3[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi3[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, wm = 0.0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    double2 a, b;
    ld_record(p, i, a, b);
    if (!FILL) {
      lo3[0] = fmin(lo3[0], a.x); hi3[0] = fmax(hi3[0], a.x);
      lo3[1] = fmin(lo3[1], a.y); hi3[1] = fmax(hi3[1], a.y);
      lo3[2] = fmin(lo3[2], b.x); hi3[2] = fmax(hi3[2], b.x);
      wm = fmax(wm, fabs(b.y));
    }
    int lo[3], cnt[3], t0[3], t1[3];
    axis_reach<SCHEME, INTERLACE>(a.x, g.org[0], g.sorg[0], g.len[0], g.inv_len[0], g.ng, lo[0], cnt[0]);
    axis_reach<SCHEME, INTERLACE>(a.y, g.org[1], g.sorg[1], g.len[1], g.inv_len[1], g.ng, lo[1], cnt[1]);
    axis_reach<SCHEME, INTERLACE>(b.x, g.org[2], g.sorg[2], g.len[2], g.inv_len[2], g.ng, lo[2], cnt[2]);
    axis_tiles(lo[0], cnt[0], TX, g.ng, t0[0], t1[0]);
    axis_tiles(lo[1], cnt[1], TY, g.ng, t0[1], t1[1]);
    axis_tiles(lo[2], cnt[2], TZ, g.ng, t0[2], t1[2]);
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int tx = u ? t1[0] : t0[0];
      if (tx < 0) continue;
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const int ty = v ? t1[1] : t0[1];
        if (ty < 0) continue;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const int tz = s ? t1[2] : t0[2];
          if (tz < 0) continue;
          const uint32_t tile = ((uint32_t) tx * td.nty + ty) * td.ntz + tz;
          const uint32_t pos = atomicAdd(cnt_or_cursor + tile, 1u);
          if (FILL) st_record(out, pos, a, b);
        }
      }
    }
  }
  if (FILL) return;
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

// round(v) as a 52-bit two's complement integer by the 1.5 * 2^52 magic add (mantissa
// field = 2^51 + round(v)): low limb = bits 0..LOBITS-1 (>= 0), high limb = the signed rest
__device__ __forceinline__ void split_fixed(double v, uint32_t &lo, uint32_t &hi) {
  const double m = v + 6755399441055744.0;
  const uint32_t lw = (uint32_t) __double2loint(m), hw = (uint32_t) __double2hiint(m);
  lo = lw & LOMASK;
  hi = (uint32_t) ((int) (((hw ^ 0x80000u) << 12) | ((lw >> LOBITS) << (LOBITS - 20))) >> (LOBITS - 20));
}

// one row of the stencil, all z cells inside the tile (consecutive): no predicates
template <int OFF>
__device__ __forceinline__ void red_pair(uint32_t addr, uint32_t lo, uint32_t hi) {
  asm volatile("red.shared.add.u32 [%0+%3], %1;\n\tred.shared.add.u32 [%0+%4], %2;"
      :: "r"(addr), "r"(lo), "r"(hi), "n"(OFF), "n"(OFF + 4 * TCELLS) : "memory");
}
template <int C, int NS>
__device__ __forceinline__ void row_inside(uint32_t addr, double wxy, const double (&wz)[NS]) {
  if constexpr (C < NS) {
    uint32_t lo, hi;
    split_fixed(wxy * wz[C], lo, hi);
    red_pair<4 * C>(addr, lo, hi);
    row_inside<C + 1, NS>(addr, wxy, wz);
  }
}

// Adds the in-tile part of one listed particle's stencil.  PARTIAL = false: the common
// case, all z cells inside the tile, straight-line code; returns true (nothing added) if
// only some of them are, and the caller queues the particle for a PARTIAL = true pass —
// so that the predicated (branchy) code runs on full warps of such particles instead of
// being dragged through every warp by one lane.
template <int SCHEME, bool PARTIAL>
__device__ __forceinline__ bool tile_add(const double2 *__restrict__ parts, size_t idx, int f,
    const AssignGeom &g, double wscale, double norm, int x0, int y0, int z0, uint32_t sm_lo) {
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
  grid_split(x[2], AxisXform{org[2], ngd, g.len[2], g.inv_len[2]}, cc, dd);
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
  grid_split(x[0], AxisXform{org[0], ngd, g.len[0], g.inv_len[0]}, cc, dd);
  stencil_from<SCHEME>(cc, dd, g.ng, ix, wx);
  grid_split(x[1], AxisXform{org[1], ngd, g.len[1], g.inv_len[1]}, cc, dd);
  stencil_from<SCHEME>(cc, dd, g.ng, iy, wy);
  // the particle weight enters through the x weights (src/genr_mesh.c:110-111, 175-177);
  // PCS folds 1/216 into it (:274-278); then the fixed-point scale
  double pw = b.y * wscale;
  if constexpr (SCHEME == 3) pw *= 0x1.2f684bda12f68p-8;
#pragma unroll
  for (int q = 0; q < NS; q++) wx[q] = (wx[q] * pw) * norm;
#pragma unroll
  for (int u = 0; u < NS; u++) {
    const uint32_t lx = (uint32_t) (ix[u] - x0);
    if (lx >= (uint32_t) TX) continue;
#pragma unroll
    for (int v = 0; v < NS; v++) {
      const uint32_t ly = (uint32_t) (iy[v] - y0);
      if (ly >= (uint32_t) TY) continue;
      const double wxy = wx[u] * wy[v];
      const uint32_t row = sm_lo + ((lx * TY + ly) * TZ) * 4u;
      if constexpr (!PARTIAL) row_inside<0, NS>(row + lz[0] * 4u, wxy, wz);
      else {
#pragma unroll
        for (int c = 0; c < NS; c++) {
          if (!zin[c]) continue;
          uint32_t lo, hi;
          split_fixed(wxy * wz[c], lo, hi);
          red_pair<0>(row + lz[c] * 4u, lo, hi);
        }
      }
    }
  }
  return false;
}

// MODE 0: the tile is stored (the mesh need not be initialised); 1: added to the mesh
template <int SCHEME, typename real, int NFIELD, int MODE>
__global__ void __launch_bounds__(ACC_THREADS, 2) k_tile_accumulate(const double2 *__restrict__ parts,
    const uint32_t *__restrict__ start, AssignGeom g, double wscale, const double *__restrict__ wmax_dev,
    real *__restrict__ mesh0, real *__restrict__ mesh1) {
  extern __shared__ uint32_t sm[];              // lo[TCELLS] | hi[TCELLS]
  __shared__ uint32_t queue[BATCH];             // listed particles that straddle the tile's z faces
  __shared__ uint32_t nqueue;
  const uint32_t sm_lo = (uint32_t) __cvta_generic_to_shared(sm);
  const TileDims td = tile_dims(g.ng);
  const uint32_t nwork = (uint32_t) NFIELD * (uint32_t) (td.ntx * td.nty * td.ntz);
  // |w wscale wx wy wz| <= wmax |wscale| for every scheme (PCS: weights in units of 1/6,
  // times 1/216)
  const double wbound = *wmax_dev * fabs(wscale);
  for (int i = threadIdx.x; i < 2 * TCELLS; i += ACC_THREADS) sm[i] = 0;
  if (threadIdx.x == 0) nqueue = 0;
  __syncthreads();
  for (uint32_t work = blockIdx.x; work < nwork; work += gridDim.x) {
    const uint32_t tile = work / NFIELD;
    const int f = (int) (work % NFIELD);
    const int tz = tile % td.ntz, ty = (tile / td.ntz) % td.nty, tx = tile / (td.ntz * td.nty);
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const uint32_t b0 = start[tile], np = start[tile + 1] - b0;
    // headroom: a cell receives at most one contribution per listed particle; the high
    // limb holds 2^31 / 2^(S - LOBITS) contributions of the largest size
    int S = 44;
    for (uint32_t q = 1u << (31 - (44 - LOBITS)); q < np && S > 24; q <<= 1) S--;
    const double norm = wbound > 0.0 ? ldexp(1.0, S) / wbound : 0.0;
    for (uint32_t base = 0; base < np; base += BATCH) {
      const uint32_t lim = min(np, base + BATCH);
      for (uint32_t j = base + threadIdx.x; j < lim; j += ACC_THREADS)
        if (tile_add<SCHEME, false>(parts, (size_t) b0 + j, f, g, wscale, norm, x0, y0, z0, sm_lo))
          queue[atomicAdd(&nqueue, 1u)] = j;
      __syncthreads();
      const uint32_t nq = nqueue;
      for (uint32_t k = threadIdx.x; k < nq; k += ACC_THREADS)
        tile_add<SCHEME, true>(parts, (size_t) b0 + queue[k], f, g, wscale, norm, x0, y0, z0, sm_lo);
      __syncthreads();
      if (threadIdx.x == 0) nqueue = 0;
      if (lim < np) {
        // fold the carries so that neither limb can wrap in the next batch
        for (int i = threadIdx.x; i < TCELLS; i += ACC_THREADS) {
          const uint32_t lo = sm[i];
          sm[i] = lo & LOMASK;
          sm[TCELLS + i] += lo >> LOBITS;
        }
        __syncthreads();
      }
    }
    // flush: every cell of the tile goes to the mesh once, coalesced along z; the tile is
    // left zeroed for the next work item
    const double inv = wbound > 0.0 ? wbound * ldexp(1.0, -S) : 0.0;
    real *m = f ? mesh1 : mesh0;
    for (int c = threadIdx.x; c < TCELLS; c += ACC_THREADS) {
      const int lzc = c % TZ, lyc = (c / TZ) % TY, lxc = c / (TZ * TY);
      const uint32_t lo = sm[c], hi = sm[TCELLS + c];
      sm[c] = 0u; sm[TCELLS + c] = 0u;
      if (x0 + lxc < g.ng && y0 + lyc < g.ng && z0 + lzc < g.ng) {
        const double v = ((double) (int) hi * (double) (1u << LOBITS) + (double) lo) * inv;
        real *cell = m + ((size_t) (x0 + lxc) * g.ng + (y0 + lyc)) * g.rowlen + z0 + lzc;
        if (MODE == 0) *cell = (real) v;
        else *cell += (real) v;
      }
    }
    __syncthreads();
  }
}

template <int SCHEME, bool INTERLACE>
int launch_lists(const double *p, size_t n, const AssignGeom &g, bool fill, uint32_t *cnt, double *out,
    double *partials, double *wmax_part, cudaStream_t st) {
  const double2 *pp = reinterpret_cast<const double2 *>(p);
  const int nblk = row_keys_blocks(n);
  if (fill)
    k_tile_lists<SCHEME, INTERLACE, true><<<nblk, 256, 0, st>>>(pp, n, g, cnt,
        reinterpret_cast<double2 *>(out), nullptr, nullptr);
  else
    k_tile_lists<SCHEME, INTERLACE, false><<<nblk, 256, 0, st>>>(pp, n, g, cnt, nullptr, partials, wmax_part);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

template <int SCHEME, typename real>
int launch_accumulate(const double *parts, const uint32_t *start, const AssignGeom &g, double wscale,
    const double *wmax, bool add, void *m0, void *m1, cudaStream_t st) {
  const size_t smem = (size_t) 2 * TCELLS * 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const double2 *pp = reinterpret_cast<const double2 *>(parts);
  real *a = static_cast<real *>(m0), *b = static_cast<real *>(m1);
#define PSB_ACC(NF, MODE)                                                                         \
  do {                                                                                            \
    auto kern = k_tile_accumulate<SCHEME, real, NF, MODE>;                                        \
    PSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
    kern<<<2 * sms, ACC_THREADS, smem, st>>>(pp, start, g, wscale, wmax, a, b);                    \
  } while (0)
  if (m1) { if (add) PSB_ACC(2, 1); else PSB_ACC(2, 0); }
  else { if (add) PSB_ACC(1, 1); else PSB_ACC(1, 0); }
#undef PSB_ACC
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

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
    if (interlace ? launch_lists<S, true>(p, n, g, false, cnt, nullptr, partials, wmax_part, st)  \
                  : launch_lists<S, false>(p, n, g, false, cnt, nullptr, partials, wmax_part, st)) \
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
    uint32_t *cursor, double *lists, cudaStream_t st) {
#define PSB_LISTS(S)                                                                              \
  case S:                                                                                         \
    return interlace ? launch_lists<S, true>(p, n, g, true, cursor, lists, nullptr, nullptr, st)  \
                     : launch_lists<S, false>(p, n, g, true, cursor, lists, nullptr, nullptr, st);
  switch (scheme) { PSB_LISTS(0) PSB_LISTS(1) PSB_LISTS(2) PSB_LISTS(3) default: return -1; }
#undef PSB_LISTS
}

// start: tile_list_count + 1 offsets into lists; add: the meshes already hold data
int launch_tile_accumulate(const double *lists, const uint32_t *start, const AssignGeom &g, int scheme,
    int precision, double wscale, const double *wmax, bool add, void *mesh0, void *mesh1, cudaStream_t st) {
#define PSB_DISPATCH(S)                                                                           \
  case S:                                                                                         \
    return precision == 8                                                                         \
        ? launch_accumulate<S, double>(lists, start, g, wscale, wmax, add, mesh0, mesh1, st)      \
        : launch_accumulate<S, float>(lists, start, g, wscale, wmax, add, mesh0, mesh1, st);
  switch (scheme) {
    PSB_DISPATCH(0) PSB_DISPATCH(1) PSB_DISPATCH(2) PSB_DISPATCH(3)
    default: set_error("unrecognised particle assignment scheme: %d\n", scheme); return -1;
  }
#undef PSB_DISPATCH
}

}  // namespace psb
