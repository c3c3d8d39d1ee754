// Prototype (development tool, not part of the product): OWNER-COMPUTES mass assignment.
//
//   1. partition: every particle is appended to the list of every mesh TILE its stencil
//      (both interlaced fields) reaches — counting pass, scan, fill (records duplicated
//      x1.5 for 16 x 16 x 48 tiles);
//   2. accumulate: one block per tile holds the tile of BOTH fields in shared memory as
//      fixed-point numbers in two 32-bit limbs and adds with the native ATOMS.ADD (the
//      only shared-memory atomic add that is not a CAS loop on sm_100a: 2.4 T adds/s vs
//      0.47 T/s for fp64, tools/smem_atomic_probe.cu); the in-tile part of each particle's
//      27-point stencil only;
//   3. flush: the tile is converted to double and written ONCE with plain coalesced
//      stores: no memset, no read-modify-write, no global atomics.
//
// Compared against the plain global-atomic scatter (one thread per particle) on the same
// particles: max |difference| per cell, and the time of every phase.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/tile_assign_probe tools/tile_assign_probe.cu
//   ./tools/tile_assign_probe [ng=1024] [npart=100000000] [clustered=0]
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>
#define CK(x) do { auto e_ = (x); if (e_ != cudaSuccess) { printf("fail %s: %s line %d\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int TX = 16, TY = 16, TZ = 48;
constexpr int TCELLS = TX * TY * TZ;
constexpr int LOBITS = 21;
constexpr unsigned LOMASK = (1u << LOBITS) - 1;

struct Geom { int ng, rowlen, ntx, nty, ntz; double scale; };   // scale = ng / L

__device__ __forceinline__ unsigned mixu(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ double u01(unsigned a, unsigned b) {
  return ((double) (a >> 6) * 67108864.0 + (double) (b >> 6)) * (1.0 / 4503599627370496.0);
}

__global__ void k_generate(double2 *p, size_t n, double L, int clustered) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const unsigned s = (unsigned) (i * 2654435761ull) ^ (unsigned) (i >> 32);
    double x[3];
    for (int a = 0; a < 3; a++) x[a] = u01(mixu(s + 11 * a + 1), mixu(s + 11 * a + 7)) * L;
    if (clustered && (i % 5)) {
      // blob centre from the particle's group of 1000, Gaussian-ish offset of ~2 cells
      const unsigned gsd = (unsigned) (i / 1000) * 7919u + 13u;
      for (int a = 0; a < 3; a++) {
        const double c = u01(mixu(gsd + 3 * a), mixu(gsd + 3 * a + 1)) * L;
        double g = 0;
        for (int q = 0; q < 4; q++) g += u01(mixu(s + 101 * a + q), mixu(s + 211 * a + q)) - 0.5;
        double v = c + g * 2.0 * 1.7320508 * (L / 512.0);
        v -= floor(v / L) * L;
        if (v >= L) v = 0;
        x[a] = v;
      }
    }
    p[2 * i] = make_double2(x[0], x[1]);
    p[2 * i + 1] = make_double2(x[2], 1.0);
  }
}

// TSC stencil of one axis: cells (periodic) and weights, src/genr_mesh.c:157-173
__device__ __forceinline__ void tsc(double t, int ng, int (&idx)[3], double (&w)[3]) {
  int c = (int) t;
  double d = t - (double) c;
  if (c >= ng) c -= ng;
  double h;
  if (d < 0.5) { idx[1] = c; idx[0] = c ? c - 1 : ng - 1; idx[2] = (c == ng - 1) ? 0 : c + 1; h = 0.5 - d; }
  else { idx[0] = c; idx[1] = (c == ng - 1) ? 0 : c + 1; idx[2] = (idx[1] == ng - 1) ? 0 : idx[1] + 1; d = 1.0 - d; h = 0.5 + d; }
  w[0] = h * (h * 0.5);
  w[1] = 0.75 - d * d;
  w[2] = 1.0 - w[0] - w[1];
}

// reference: plain global atomics, both fields
__global__ void k_ref(const double2 *__restrict__ p, size_t n, Geom g, double *m0, double *m1) {
  const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 a = p[2 * i], b = p[2 * i + 1];
  const double x[3] = {a.x, a.y, b.x};
  for (int f = 0; f < 2; f++) {
    int ix[3], iy[3], iz[3];
    double wx[3], wy[3], wz[3];
    double t[3];
    for (int k = 0; k < 3; k++) { t[k] = x[k] * g.scale + 0.5 * f; if (t[k] >= g.ng) t[k] -= g.ng; }
    tsc(t[0], g.ng, ix, wx); tsc(t[1], g.ng, iy, wy); tsc(t[2], g.ng, iz, wz);
    double *m = f ? m1 : m0;
    for (int q = 0; q < 3; q++) wx[q] *= b.y;
    for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) for (int s = 0; s < 3; s++)
      atomicAdd(m + ((size_t) ix[u] * g.ng + iy[v]) * g.rowlen + iz[s], (wx[u] * wy[v]) * wz[s]);
  }
}

// cells [lo, lo + len) (periodic) touched along one axis by both fields: base cell c of the
// unshifted grid, fraction d: unshifted TSC covers c-1..c+1 (d < 1/2) or c..c+2, the
// half-cell shifted grid always c..c+2 in its own indices (= array indices)
__device__ __forceinline__ void reach(double t, int ng, int &lo, int &len) {
  int c = (int) t;
  const double d = t - (double) c;
  if (c >= ng) c -= ng;
  if (d < 0.5) { lo = c ? c - 1 : ng - 1; len = 4; } else { lo = c; len = 3; }
}

// tiles along one axis overlapped by the (periodic) cells lo .. lo+len-1, len <= 4 <= cells
// of any tile: the first cell's tile and, if different, the last cell's
__device__ __forceinline__ void tiles_of(int lo, int len, int T, int ng, int &t0, int &t1) {
  int last = lo + len - 1;
  if (last >= ng) last -= ng;
  t0 = lo / T;
  t1 = last / T;
  if (t1 == t0) t1 = -1;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_tiles(const double2 *__restrict__ p, size_t n, Geom g,
    unsigned *__restrict__ cnt_or_cursor, double2 *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const double2 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    int lo[3], len[3], t0[3], t1[3];
    reach(a.x * g.scale, g.ng, lo[0], len[0]);
    reach(a.y * g.scale, g.ng, lo[1], len[1]);
    reach(b.x * g.scale, g.ng, lo[2], len[2]);
    tiles_of(lo[0], len[0], TX, g.ng, t0[0], t1[0]);
    tiles_of(lo[1], len[1], TY, g.ng, t0[1], t1[1]);
    tiles_of(lo[2], len[2], TZ, g.ng, t0[2], t1[2]);
    for (int u = 0; u < 2; u++) {
      const int tx = u ? t1[0] : t0[0];
      if (tx < 0) continue;
      for (int v = 0; v < 2; v++) {
        const int ty = v ? t1[1] : t0[1];
        if (ty < 0) continue;
        for (int s = 0; s < 2; s++) {
          const int tz = s ? t1[2] : t0[2];
          if (tz < 0) continue;
          const unsigned tile = ((unsigned) tx * g.nty + ty) * g.ntz + tz;
          const unsigned pos = atomicAdd(cnt_or_cursor + tile, 1u);
          if (FILL) asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(out + 2 * (size_t) pos), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
        }
      }
    }
  }
}

// fixed point: round(v) as a 52-bit two's complement integer by the 1.5 * 2^52 magic add
// (mantissa field = 2^51 + round(v)), split into a 22-bit low limb (>= 0) and the signed rest
__device__ __forceinline__ void add_fixed(unsigned *cell, double v) {
  const double m = v + 6755399441055744.0;
  const unsigned lw = (unsigned) __double2loint(m), hw = (unsigned) __double2hiint(m);
  atomicAdd(cell, lw & LOMASK);
  // bits LOBITS..51 of the 52-bit two's complement integer, sign-extended from bit 51
  atomicAdd(cell + 1, (unsigned) ((int) (((hw ^ 0x80000u) << 12) | ((lw >> LOBITS) << (LOBITS - 20))) >> (LOBITS - 20)));
}

// one block per (tile, field): 96 KB of shared memory, two blocks per SM, so that one
// block's flush (DRAM writes) overlaps the other's accumulation (shared-memory atomics)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_accumulate(const double2 *__restrict__ parts,
    const unsigned *__restrict__ start, Geom g, double *__restrict__ m0, double *__restrict__ m1) {
  extern __shared__ unsigned sm[];              // [cell][2 limbs]
  const unsigned nwork = 2u * (unsigned) (g.ntx * g.nty * g.ntz);
  for (int i = threadIdx.x; i < TCELLS * 2; i += THREADS) sm[i] = 0;
  __syncthreads();
  for (unsigned work = blockIdx.x; work < nwork; work += gridDim.x) {
    const unsigned tile = work >> 1;
    const int f = work & 1;
    const int tz = tile % g.ntz, ty = (tile / g.ntz) % g.nty, tx = tile / (g.ntz * g.nty);
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const unsigned b0 = start[tile], b1 = start[tile + 1];
    const unsigned np = b1 - b0;
    // headroom: a cell receives at most one contribution per listed particle; the low limb
    // takes 2^(32 - LOBITS) adds, the high limb 2^31 / 2^(S - LOBITS)
    int S = 44;
    for (unsigned q = 1u << (31 - (44 - LOBITS)); q < np && S > 24; q <<= 1) S--;
    const double sc = ldexp(1.0, S);
    const double shift = 0.5 * f;
    constexpr unsigned BATCH = 1u << (32 - LOBITS);
    for (unsigned base = 0; base < np; base += BATCH) {
      const unsigned lim = min(np, base + BATCH);
      for (unsigned j = base + threadIdx.x; j < lim; j += THREADS) {
        const double2 a = __ldg(parts + 2 * (size_t) (b0 + j)), b = __ldg(parts + 2 * (size_t) (b0 + j) + 1);
        int ix[3], iy[3], iz[3];
        double wx[3], wy[3], wz[3];
        double t[3] = {a.x * g.scale + shift, a.y * g.scale + shift, b.x * g.scale + shift};
#pragma unroll
        for (int k = 0; k < 3; k++) if (t[k] >= g.ng) t[k] -= g.ng;
        tsc(t[0], g.ng, ix, wx); tsc(t[1], g.ng, iy, wy); tsc(t[2], g.ng, iz, wz);
#pragma unroll
        for (int q = 0; q < 3; q++) wx[q] *= b.y * sc;
        unsigned lz[3];
#pragma unroll
        for (int s = 0; s < 3; s++) lz[s] = (unsigned) (iz[s] - z0);
        if (lz[0] >= (unsigned) TZ && lz[1] >= (unsigned) TZ && lz[2] >= (unsigned) TZ) continue;
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const unsigned lx = (unsigned) (ix[u] - x0);
          if (lx >= (unsigned) TX) continue;
#pragma unroll
          for (int v = 0; v < 3; v++) {
            const unsigned ly = (unsigned) (iy[v] - y0);
            if (ly >= (unsigned) TY) continue;
            const double wxy = wx[u] * wy[v];
            unsigned *row = sm + ((lx * TY + ly) * TZ) * 2;
#pragma unroll
            for (int s = 0; s < 3; s++)
              if (lz[s] < (unsigned) TZ) add_fixed(row + 2 * lz[s], wxy * wz[s]);
          }
        }
      }
      __syncthreads();
      if (lim < np) {
        // carry the low limbs into the high ones so that neither can wrap in the next batch
        for (int i = threadIdx.x; i < TCELLS; i += THREADS) {
          const unsigned lo = sm[2 * i];
          sm[2 * i] = lo & LOMASK;
          sm[2 * i + 1] += lo >> LOBITS;
        }
        __syncthreads();
      }
    }
    // flush: plain coalesced stores, every cell of the mesh is written exactly once; the
    // tile is left zeroed for the next work item
    const double inv = ldexp(1.0, -S);
    double *m = f ? m1 : m0;
    for (int c = threadIdx.x; c < TCELLS; c += THREADS) {
      const int lz = c % TZ, ly = (c / TZ) % TY, lx = c / (TZ * TY);
      const uint2 v2 = *reinterpret_cast<uint2 *>(sm + 2 * c);
      *reinterpret_cast<uint2 *>(sm + 2 * c) = make_uint2(0u, 0u);
      if (x0 + lx < g.ng && y0 + ly < g.ng && z0 + lz < g.ng)
        m[((size_t) (x0 + lx) * g.ng + (y0 + ly)) * g.rowlen + z0 + lz] =
            ((double) (int) v2.y * (double) (1 << LOBITS) + (double) v2.x) * inv;
    }
    __syncthreads();
  }
}

int main(int argc, char **argv) {
  const int ng = argc > 1 ? atoi(argv[1]) : 1024;
  const size_t n = argc > 2 ? (size_t) atof(argv[2]) : 100000000;
  const int clustered = argc > 3 ? atoi(argv[3]) : 0;
  const double L = 1000.0;
  Geom g;
  g.ng = ng; g.rowlen = 2 * (ng / 2 + 1); g.scale = ng / L;
  g.ntx = (ng + TX - 1) / TX; g.nty = (ng + TY - 1) / TY; g.ntz = (ng + TZ - 1) / TZ;
  const unsigned ntile = (unsigned) g.ntx * g.nty * g.ntz;
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  printf("%s, %d SMs; ng %d, %zu particles (%s), tiles %d x %d x %d cells: %u tiles\n", pr.name,
      pr.multiProcessorCount, ng, n, clustered ? "clustered" : "uniform", TX, TY, TZ, ntile);
  const size_t mesh_elems = (size_t) ng * ng * g.rowlen;
  double2 *p, *dup;
  double *r0, *r1, *m0, *m1;
  unsigned *cnt, *start;
  CK(cudaMalloc(&p, n * 32));
  CK(cudaMalloc(&r0, mesh_elems * 8)); CK(cudaMalloc(&r1, mesh_elems * 8));
  CK(cudaMalloc(&m0, mesh_elems * 8)); CK(cudaMalloc(&m1, mesh_elems * 8));
  CK(cudaMalloc(&cnt, (ntile + 1) * 4)); CK(cudaMalloc(&start, (ntile + 1) * 4));
  k_generate<<<148 * 8, 256>>>(p, n, L, clustered);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e[8];
  for (auto &x : e) CK(cudaEventCreate(&x));
  auto ms = [&](int a, int b) { float t; CK(cudaEventElapsedTime(&t, e[a], e[b])); return t; };

  // ---- reference
  CK(cudaMemset(r0, 0, mesh_elems * 8)); CK(cudaMemset(r1, 0, mesh_elems * 8));
  CK(cudaEventRecord(e[0]));
  k_ref<<<(unsigned) ((n + 255) / 256), 256>>>(p, n, g, r0, r1);
  CK(cudaEventRecord(e[1]));
  CK(cudaDeviceSynchronize());
  printf("reference (unsorted global atomics)      %8.3f ms\n", ms(0, 1));

  // ---- owner-computes
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, start, (int) ntile + 1);
  void *tmp;
  CK(cudaMalloc(&tmp, tmp_bytes));
  const size_t smem = (size_t) TCELLS * 2 * 4;
  auto kern = k_accumulate<512>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  CK(cudaMemset(m0, 0xff, mesh_elems * 8)); CK(cudaMemset(m1, 0xff, mesh_elems * 8));   // poison: no memset needed
  double2 *dupbuf = nullptr;
  size_t dupcap = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e[0]));
    CK(cudaMemsetAsync(cnt, 0, (ntile + 1) * 4));
    k_tiles<false><<<148 * 16, 256>>>(p, n, g, cnt, nullptr);
    CK(cudaEventRecord(e[1]));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, start, (int) ntile + 1);
    unsigned total;
    CK(cudaMemcpy(&total, start + ntile, 4, cudaMemcpyDeviceToHost));
    if (total > dupcap) { if (dupbuf) cudaFree(dupbuf); dupcap = total + total / 8; CK(cudaMalloc(&dupbuf, dupcap * 32)); }
    dup = dupbuf;
    CK(cudaEventRecord(e[2]));
    CK(cudaMemcpyAsync(cnt, start, (ntile + 1) * 4, cudaMemcpyDeviceToDevice));
    k_tiles<true><<<148 * 16, 256>>>(p, n, g, cnt, dup);
    CK(cudaEventRecord(e[3]));
    kern<<<2 * pr.multiProcessorCount, 512, smem>>>(dup, start, g, m0, m1);
    CK(cudaEventRecord(e[4]));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    printf("rep %d: count %.3f  scan+sync %.3f  fill %.3f (x%.3f records)  accumulate+flush %.3f   total %.3f ms\n",
        rep, ms(0, 1), ms(1, 2), ms(2, 3), (double) total / n, ms(3, 4), ms(0, 4));
  }
  // ---- compare
  std::vector<double> ha(mesh_elems), hb(mesh_elems);
  double worst = 0, sum = 0;
  for (int f = 0; f < 2; f++) {
    CK(cudaMemcpy(ha.data(), f ? r1 : r0, mesh_elems * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), f ? m1 : m0, mesh_elems * 8, cudaMemcpyDeviceToHost));
    for (int x = 0; x < ng; x++) for (int y = 0; y < ng; y++) for (int z = 0; z < ng; z++) {
      const size_t i = ((size_t) x * ng + y) * g.rowlen + z;
      worst = std::max(worst, fabs(ha[i] - hb[i]));
      sum += hb[i];
    }
  }
  printf("max |owner - reference| per cell: %.3e   (sum of both meshes %.6f, expected %.1f)\n", worst, sum, 2.0 * n);
  return 0;
}
