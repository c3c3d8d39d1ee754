// Mass assignment for sm_100a: coordinate bounds, row-key counting sort and the
// NGP / CIC / TSC / PCS scatter, with the half-cell shifted (interlaced) field
// produced in the same pass over the particles.
//
// Replaces the reference's OpenMP loops (paths relative to cheng-zhao/powspec):
//   get_coord_bound   src/genr_mesh.c:427-492
//   ngp/cic/tsc/pcs   src/genr_mesh.c:50-71, 84-130, 143-234, 247-412
//   shift_cat         src/genr_mesh.c:590-602   (fused: per-particle wrap)
//   gen_dens          src/genr_mesh.c:793-858   (orchestrated in context.cu)
//
// Design: particles are 32-byte records read once with
// two 128-bit loads; a counting sort on the (x,y) row of the base cell makes
// every warp work on one mesh row neighbourhood, so the read-modify-write
// traffic of the fp64 reductions stays in L2 and each mesh sector goes to HBM
// once.  Accumulation uses no-return global reductions (RED.ADD.F64/F32), which
// on this part are faster than CAS-emulated fp64 shared-memory atomics.
//
// Arithmetic that decides an integer (the cell index) is spelled with
// round-to-nearest intrinsics in the reference's operation order
// ((x - origin) * Ng / L), so no FMA contraction can move a particle.

#include "assign_common.cuh"

#include <cfloat>

namespace psb {

static int grid_for(size_t n, int per_block, int max_blocks) {
  size_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > (size_t) max_blocks) b = max_blocks;
  return (int) b;
}

// ---------------------------------------------------------------------------
// bounds: per-block partial min/max of the three coordinates
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bounds(const double2 *__restrict__ p,
    size_t n, double *__restrict__ partials) {
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
  double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    double2 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    lo[0] = fmin(lo[0], a.x); hi[0] = fmax(hi[0], a.x);
    lo[1] = fmin(lo[1], a.y); hi[1] = fmax(hi[1], a.y);
    lo[2] = fmin(lo[2], b.x); hi[2] = fmax(hi[2], b.x);
  }
  __shared__ double s[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { s[a][warp] = lo[a]; s[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = s[threadIdx.x][0];
    for (int w = 1; w < 8; w++)
      v = (threadIdx.x < 3) ? fmin(v, s[threadIdx.x][w]) : fmax(v, s[threadIdx.x][w]);
    partials[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

int launch_bounds(const double *p, size_t n, double *partials, int nblk,
    cudaStream_t st) {
  k_bounds<<<nblk, 256, 0, st>>>(reinterpret_cast<const double2 *>(p), n, partials);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// counting sort by the (x,y) row of the base cell on the unshifted grid
// ---------------------------------------------------------------------------
// `partials` (optional, [gridDim.x][6]): per-block min/max of the coordinates,
// so that simulation boxes get def_box's bound checks without a separate pass
// over the catalogue (the record is one 32-byte sector: z comes for free).
__global__ void __launch_bounds__(256) k_row_keys(const double2 *__restrict__ p,
    size_t n, AssignGeom g, uint32_t *__restrict__ keys, uint32_t *__restrict__ hist,
    double *__restrict__ partials) {
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
  double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    double2 a = __ldg(p + 2 * i);
    if (partials) {
      const double z = __ldg(&p[2 * i + 1].x);
      lo[0] = fmin(lo[0], a.x); hi[0] = fmax(hi[0], a.x);
      lo[1] = fmin(lo[1], a.y); hi[1] = fmax(hi[1], a.y);
      lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
    }
    int cx = base_cell(grid_coord(a.x, g.org[0], g.len[0], g.ng), g.ng);
    int cy = base_cell(grid_coord(a.y, g.org[1], g.len[1], g.ng), g.ng);
    // strip-major row order: the sweep over x stays inside a strip of
    // g.strip rows, so the planes it touches fit in L2
    int cxl = cx - g.x0;                // plane index inside the owned x-slab
    if (cxl < 0) cxl += g.ng;
    if (cxl >= g.nx) cxl = g.nx - 1;    // not ours (caller's routing error): keep the key in range
    uint32_t key;
    if (g.xgroup > 0)           // coarse buckets: (strip, group of xgroup planes), unordered inside
      key = (uint32_t) (cy / g.strip) * (uint32_t) ((g.nx + g.xgroup - 1) / g.xgroup)
          + (uint32_t) (cxl / g.xgroup);
    else
      key = ((uint32_t) (cy / g.strip) * (uint32_t) g.nx + (uint32_t) cxl)
          * (uint32_t) g.strip + (uint32_t) (cy % g.strip);
    keys[i] = key;
    atomicAdd(hist + key, 1u);
  }
  if (!partials) return;
  __shared__ double s[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { s[a][warp] = lo[a]; s[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = s[threadIdx.x][0];
    for (int w = 1; w < 8; w++)
      v = (threadIdx.x < 3) ? fmin(v, s[threadIdx.x][w]) : fmax(v, s[threadIdx.x][w]);
    partials[blockIdx.x * 6 + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) k_row_scatter(const double2 *__restrict__ p,
    size_t n, const uint32_t *__restrict__ keys, uint32_t *__restrict__ cursor,
    double2 *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    double2 a, b;
    ld_record(p, i, a, b);
    uint32_t pos = atomicAdd(cursor + keys[i], 1u);
    st_record(out, pos, a, b);
  }
}

size_t row_key_count(const AssignGeom &g) {
  const size_t nstrip = ((size_t) g.ng + g.strip - 1) / g.strip;
  if (g.xgroup > 0) return nstrip * (size_t) ((g.nx + g.xgroup - 1) / g.xgroup);
  return nstrip * (size_t) g.nx * (size_t) g.strip;
}

int row_keys_blocks(size_t n) { return grid_for(n, 256, 148 * 16); }

int launch_row_keys(const double *p, size_t n, const AssignGeom &g, uint32_t *keys,
    uint32_t *hist, double *partials, cudaStream_t st) {
  k_row_keys<<<row_keys_blocks(n), 256, 0, st>>>(
      reinterpret_cast<const double2 *>(p), n, g, keys, hist, partials);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_row_scatter(const double *p, size_t n, const uint32_t *keys,
    uint32_t *cursor, double *sorted, cudaStream_t st) {
  k_row_scatter<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
      reinterpret_cast<const double2 *>(p), n, keys, cursor,
      reinterpret_cast<double2 *>(sorted));
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// x-plane of the global mesh -> plane of the local buffer (slab decomposition:
// the buffer holds planes xbase .. xbase+nxloc-1 of the periodic mesh, i.e. the
// owned slab plus its halo planes; single GPU: xbase = 0, nxloc = Ng)
__device__ __forceinline__ int local_plane(int ix, const AssignGeom &g) {
  int lp = ix - g.xbase;
  if (lp < 0) lp += g.ng;
  return lp;
}

template <typename real> __device__ __forceinline__ void red_add(real *addr, double v) {
  atomicAdd(addr, (real) v);    // result unused -> RED.E.ADD.F64 / .F32
}

template <int SCHEME, typename real>
__device__ __forceinline__ void scatter_one(const double x[3], double pw,
    const double org[3], const AssignGeom &g, real *__restrict__ mesh) {
  constexpr int NS = SCHEME + 1;
  int ix[NS], iy[NS], iz[NS];
  double wx[NS], wy[NS], wz[NS];
  axis_stencil<SCHEME>(grid_coord(x[0], org[0], g.len[0], g.ng), g.ng, ix, wx);
  axis_stencil<SCHEME>(grid_coord(x[1], org[1], g.len[1], g.ng), g.ng, iy, wy);
  axis_stencil<SCHEME>(grid_coord(x[2], org[2], g.len[2], g.ng), g.ng, iz, wz);
  // the particle weight enters through the x weights (src/genr_mesh.c:110-111,
  // 175-177); PCS folds 1/216 = 0x1.2f684bda12f68p-8 into it (:274-278)
  if constexpr (SCHEME == 3) pw *= 0x1.2f684bda12f68p-8;
#pragma unroll
  for (int a = 0; a < NS; a++) wx[a] *= pw;
#pragma unroll
  for (int a = 0; a < NS; a++) {
    const int lp = local_plane(ix[a], g);
    if (lp >= g.nxloc) continue;        // outside this slab's buffer (never for routed particles)
#pragma unroll
    for (int b = 0; b < NS; b++) {
      const double wxy = wx[a] * wy[b];
      real *row = mesh + ((size_t) lp * g.ng + iy[b]) * g.rowlen;
#pragma unroll
      for (int c = 0; c < NS; c++) red_add(row + iz[c], wxy * wz[c]);
    }
  }
}

template <int SCHEME, typename real, bool INTERLACE>
__global__ void __launch_bounds__(256) k_assign(const double2 *__restrict__ p, size_t n,
    AssignGeom g, double wscale, real *__restrict__ mesh0, real *__restrict__ mesh1) {
  // one particle per thread, blocks in particle order: with the strip-sorted
  // catalogue the set of mesh rows being updated moves monotonically, so every
  // mesh sector is fetched into L2 once and written back once
  const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (i < n) {
    double2 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    double x[3] = {a.x, a.y, b.x};
    const double pw = b.y * wscale;
    scatter_one<SCHEME, real>(x, pw, g.org, g, mesh0);
    if constexpr (INTERLACE) {
      // shift_cat, src/genr_mesh.c:595-600: periodic wrap into the shifted box
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (x[k] >= __dadd_rn(g.sorg[k], g.len[k])) x[k] = __dsub_rn(x[k], g.len[k]);
      scatter_one<SCHEME, real>(x, pw, g.sorg, g, mesh1);
    }
  }
}


// ---------------------------------------------------------------------------
// z-coalesced scatter.  The cells a particle touches along z are contiguous in
// memory (8..32 bytes), but with one particle per thread every lane of a RED
// instruction hits a different mesh row, i.e. 32 L2 sector requests per
// instruction — and the L2 tag/request rate, not HBM, bounds the kernel (ncu:
// lts__t_tag_requests ~72 % of peak, DRAM traffic at the algorithmic minimum).
// Here NZ = SCHEME+1 adjacent lanes share a particle, one z-cell each, so the
// L1 merges them into one or two sector requests: the L2 request count drops
// 1.6x (CIC), 2x (TSC), 2.3x (PCS).  The x/y stencils are recomputed by each of
// the NZ lanes; to keep that redundancy cheap the coordinate transform avoids
// the XU pipe: the division by the (constant) box size is a Markstein
// reciprocal-FMA sequence, floor/int conversion a magic-number add.  Both are
// re-done with the exact IEEE division whenever the fractional part is within
// 1e-9 of a value that decides a cell (0, 1/2, 1), so the cell a particle lands
// in is always the reference's.
// ---------------------------------------------------------------------------
template <int SCHEME, typename real>
__device__ __forceinline__ void scatter_coop(const double x[3], double pw, const double org[3],
    const AssignGeom &g, int zsel, real *__restrict__ mesh) {
  constexpr int NS = SCHEME + 1;
  int ix[NS], iy[NS], iz[NS];
  double wx[NS], wy[NS], wz[NS];
  int c; double d;
  const double ngd = (double) g.ng;
  grid_split(x[0], AxisXform{org[0], ngd, g.len[0], g.inv_len[0]}, c, d);
  stencil_from<SCHEME>(c, d, g.ng, ix, wx);
  grid_split(x[1], AxisXform{org[1], ngd, g.len[1], g.inv_len[1]}, c, d);
  stencil_from<SCHEME>(c, d, g.ng, iy, wy);
  grid_split(x[2], AxisXform{org[2], ngd, g.len[2], g.inv_len[2]}, c, d);
  stencil_from<SCHEME>(c, d, g.ng, iz, wz);
  // this lane's z-cell (compile-time unrolled select: no dynamic register indexing)
  int izc = iz[0];
  double wzc = wz[0];
#pragma unroll
  for (int q = 1; q < NS; q++)
    if (zsel == q) { izc = iz[q]; wzc = wz[q]; }
  if constexpr (SCHEME == 3) pw *= 0x1.2f684bda12f68p-8;
#pragma unroll
  for (int a = 0; a < NS; a++) wx[a] *= pw;
#pragma unroll
  for (int a = 0; a < NS; a++) {
    const int lp = local_plane(ix[a], g);
    if (lp >= g.nxloc) continue;        // outside this slab's buffer (never for routed particles)
#pragma unroll
    for (int b = 0; b < NS; b++) {
      real *row = mesh + ((size_t) lp * g.ng + iy[b]) * g.rowlen;
      red_add(row + izc, (wx[a] * wy[b]) * wzc);
    }
  }
}

// COOP_ITER particle groups per warp (loads issued up front), BLOCK threads per
// block, MINB blocks per SM for the register allocator.  Variants are selected
// at run time (option "coop_variant") for ablation; the default is the one
// measured fastest on B200 for TSC + interlacing.
template <int SCHEME, typename real, bool INTERLACE, int ITER, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_assign_coop(const double2 *__restrict__ p,
    size_t n, AssignGeom g, double wscale, real *__restrict__ mesh0, real *__restrict__ mesh1) {
  constexpr int NZ = SCHEME + 1;
  constexpr int PPW = 32 / NZ;                  // particles per warp and iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane / NZ, zsel = lane - sub * NZ;
  const size_t warp = blockIdx.x * (size_t) (BLOCK >> 5) + (threadIdx.x >> 5);
  const size_t first = warp * (PPW * ITER) + sub;
  // the NZ lanes of a particle load the same 32 bytes: one broadcast request
  double2 a[ITER], b[ITER];
#pragma unroll
  for (int it = 0; it < ITER; it++) {
    const size_t i = first + (size_t) it * PPW;
    if (sub < PPW && i < n) ld_record(p, i, a[it], b[it]);
  }
#pragma unroll
  for (int it = 0; it < ITER; it++) {
    const size_t i = first + (size_t) it * PPW;
    if (sub < PPW && i < n) {
      double x[3] = {a[it].x, a[it].y, b[it].x};
      const double pw = b[it].y * wscale;
      scatter_coop<SCHEME, real>(x, pw, g.org, g, zsel, mesh0);
      if constexpr (INTERLACE) {
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (x[k] >= __dadd_rn(g.sorg[k], g.len[k])) x[k] = __dsub_rn(x[k], g.len[k]);
        scatter_coop<SCHEME, real>(x, pw, g.sorg, g, zsel, mesh1);
      }
    }
  }
}

template <int SCHEME, typename real, int ITER, int BLOCK, int MINB>
static int launch_coop_v(const double2 *pp, size_t n, const AssignGeom &g, double wscale, void *m0,
    void *m1, cudaStream_t st) {
  constexpr int PPB = (BLOCK / 32) * ITER * (32 / (SCHEME + 1));   // particles per block
  const size_t nb = (n + PPB - 1) / PPB;
  if (nb > 0x7fffffffull) { set_error("too many particles in one assignment chunk\n"); return -1; }
  if (m1)
    k_assign_coop<SCHEME, real, true, ITER, BLOCK, MINB><<<(int) nb, BLOCK, 0, st>>>(pp, n, g, wscale,
        static_cast<real *>(m0), static_cast<real *>(m1));
  else
    k_assign_coop<SCHEME, real, false, ITER, BLOCK, MINB><<<(int) nb, BLOCK, 0, st>>>(pp, n, g, wscale,
        static_cast<real *>(m0), nullptr);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

template <int SCHEME, typename real>
static int launch_assign_t(const double *p, size_t n, const AssignGeom &g, double wscale,
    void *m0, void *m1, cudaStream_t st) {
  const double2 *pp = reinterpret_cast<const double2 *>(p);
  if (g.coop && SCHEME > 0) {
    switch (g.coop_variant) {
      case 1: return launch_coop_v<SCHEME, real, 1, 128, 8>(pp, n, g, wscale, m0, m1, st);
      case 2: return launch_coop_v<SCHEME, real, 2, 256, 4>(pp, n, g, wscale, m0, m1, st);
      case 3: return launch_coop_v<SCHEME, real, 1, 256, 6>(pp, n, g, wscale, m0, m1, st);
      case 4: return launch_coop_v<SCHEME, real, 2, 128, 8>(pp, n, g, wscale, m0, m1, st);
      default: return launch_coop_v<SCHEME, real, 1, 256, 1>(pp, n, g, wscale, m0, m1, st);
    }
  }
  const size_t nblk = (n + 255) / 256;
  if (nblk > 0x7fffffffull) { set_error("too many particles in one assignment chunk\n"); return -1; }
  const int grid = (int) nblk;
  if (m1)
    k_assign<SCHEME, real, true><<<grid, 256, 0, st>>>(pp, n, g, wscale,
        static_cast<real *>(m0), static_cast<real *>(m1));
  else
    k_assign<SCHEME, real, false><<<grid, 256, 0, st>>>(pp, n, g, wscale,
        static_cast<real *>(m0), nullptr);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

int launch_assign(const double *p, size_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *mesh0, void *mesh1, cudaStream_t st) {
  if (n == 0) return 0;
#define PSB_DISPATCH(S)                                                         \
  case S:                                                                       \
    return precision == 8                                                       \
        ? launch_assign_t<S, double>(p, n, g, wscale, mesh0, mesh1, st)         \
        : launch_assign_t<S, float>(p, n, g, wscale, mesh0, mesh1, st);
  switch (scheme) {
    PSB_DISPATCH(0)
    PSB_DISPATCH(1)
    PSB_DISPATCH(2)
    PSB_DISPATCH(3)
    default:
      set_error("unrecognised particle assignment scheme: %d\n", scheme);
      return -1;
  }
#undef PSB_DISPATCH
}

// ---------------------------------------------------------------------------
// slab decomposition helpers: owner of a particle = slab of its base x-cell
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_owner_keys(const double2 *__restrict__ p, size_t n,
    AssignGeom g, int nranks, uint32_t *__restrict__ keys, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = 0;
  __syncthreads();
  const int per = g.ng / nranks;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x) {
    double2 a = __ldg(p + 2 * i);
    int cx = base_cell(grid_coord(a.x, g.org[0], g.len[0], g.ng), g.ng);
    int r = cx / per;
    if (r >= nranks) r = nranks - 1;
    keys[i] = (uint32_t) r;
    atomicAdd(&sh[r], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nranks && sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}

int launch_owner_keys(const double *p, size_t n, const AssignGeom &g, int nranks, uint32_t *keys,
    uint32_t *hist, cudaStream_t st) {
  if (nranks > 64) { set_error("at most 64 slabs are supported\n"); return -1; }
  k_owner_keys<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(
      reinterpret_cast<const double2 *>(p), n, g, nranks, keys, hist);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// Block-aggregated partition by owner: with at most 64 destinations a plain
// per-particle atomic cursor serialises on 2..64 addresses.  Each block counts
// its particles per destination in shared memory, reserves one contiguous range
// per destination with a single global atomic, and fills it.
constexpr int OWN_PER_THREAD = 8;
__global__ void __launch_bounds__(256) k_owner_scatter(const double2 *__restrict__ p, size_t n,
    const uint32_t *__restrict__ keys, uint32_t *__restrict__ cursor, int nranks,
    double2 *__restrict__ out) {
  __shared__ uint32_t cnt[64], base[64];
  if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t tile = (size_t) blockIdx.x * (256 * OWN_PER_THREAD);
  uint32_t key[OWN_PER_THREAD], slot[OWN_PER_THREAD];
#pragma unroll
  for (int q = 0; q < OWN_PER_THREAD; q++) {
    const size_t i = tile + (size_t) q * 256 + threadIdx.x;
    key[q] = 0xffffffffu;
    if (i < n) { key[q] = keys[i]; slot[q] = atomicAdd(&cnt[key[q]], 1u); }
  }
  __syncthreads();
  if (threadIdx.x < nranks && cnt[threadIdx.x])
    base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, cnt[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < OWN_PER_THREAD; q++) {
    const size_t i = tile + (size_t) q * 256 + threadIdx.x;
    if (i < n) {
      const size_t pos = (size_t) base[key[q]] + slot[q];
      double2 a, b;
      ld_record(p, i, a, b);
      st_record(out, pos, a, b);
    }
  }
}

int launch_owner_scatter(const double *p, size_t n, const uint32_t *keys, uint32_t *cursor,
    int nranks, double *sorted, cudaStream_t st) {
  if (!n) return 0;
  const size_t nb = (n + 256 * OWN_PER_THREAD - 1) / (256 * OWN_PER_THREAD);
  k_owner_scatter<<<(unsigned) nb, 256, 0, st>>>(reinterpret_cast<const double2 *>(p), n, keys,
      cursor, nranks, reinterpret_cast<double2 *>(sorted));
  PSB_CUDA(cudaGetLastError());
  return 0;
}

template <typename real>
__global__ void k_add(real *__restrict__ dst, const real *__restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
       i += (size_t) gridDim.x * blockDim.x)
    dst[i] += src[i];
}

int launch_add(void *dst, const void *src, size_t n, int precision, cudaStream_t st) {
  if (!n) return 0;
  if (precision == 8)
    k_add<double><<<grid_for(n, 256, 148 * 16), 256, 0, st>>>((double *) dst, (const double *) src, n);
  else
    k_add<float><<<grid_for(n, 256, 148 * 16), 256, 0, st>>>((float *) dst, (const float *) src, n);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// padded (in-place FFT layout) -> dense copy, for tests
// ---------------------------------------------------------------------------
template <typename real>
__global__ void k_unpad(const real *__restrict__ src, real *__restrict__ dst, int ng,
    int rowlen) {
  const size_t nrow = (size_t) ng * ng;
  for (size_t r = blockIdx.x; r < nrow; r += gridDim.x)
    for (int k = threadIdx.x; k < ng; k += blockDim.x)
      dst[r * ng + k] = src[r * rowlen + k];
}

int launch_unpad_copy(const void *mesh, void *dst, int ng, int rowlen, int precision,
    cudaStream_t st) {
  if (precision == 8)
    k_unpad<double><<<148 * 8, 128, 0, st>>>(static_cast<const double *>(mesh),
        static_cast<double *>(dst), ng, rowlen);
  else
    k_unpad<float><<<148 * 8, 128, 0, st>>>(static_cast<const float *>(mesh),
        static_cast<float *>(dst), ng, rowlen);
  PSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb
