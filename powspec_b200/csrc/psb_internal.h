// Internal declarations shared by the translation units of libpowspec_b200.so.
// Not installed; the public ABI is include/powspec_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>

namespace psb {

// ---------------------------------------------------------------------------
// error handling: message in the reference's P_ERR style (src/define.h:102,129)
// ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();
void errors_quiet(bool on);      // record messages without printing them (this thread)

#define PSB_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      psb::set_error("CUDA failure %s at %s:%d: %s\n", #call, __FILE__,         \
          __LINE__, cudaGetErrorString(e_));                                    \
      return -1;                                                                \
    }                                                                           \
  } while (0)

// ---------------------------------------------------------------------------
// geometry of one assignment pass
// ---------------------------------------------------------------------------
struct AssignGeom {
  int ng;               // cells per side
  int rowlen;           // reals per z-row of the (padded, in-place FFT) mesh
  int strip;            // rows per strip of the sort order (row-key layout)
  int xgroup;           // > 0: coarse sort into (strip, xgroup planes) buckets only
  int coop;             // z-coalesced scatter (NZ lanes per particle)
  int coop_variant;     // launch shape of the z-coalesced kernel (ablation)
  int x0, nx;           // owned x-planes [x0, x0+nx) (single GPU: 0, Ng)
  int xbase, nxloc;     // planes held by the local buffer: xbase .. xbase+nxloc-1 (mod Ng)
  double org[3];        // lower box corner (MESH.min)
  double sorg[3];       // corner of the half-cell shifted box (MESH.smin)
  double len[3];        // box size
  double inv_len[3];    // RN(1 / len), for the division-free coordinate transform
};

// launch wrappers (assign.cu).  All asynchronous on `st`; return 0 / -1.
int launch_bounds(const double *p, size_t n, double *partials /*[nblk*6]*/,
    int nblk, cudaStream_t st);
// number of distinct sort keys for this geometry
size_t row_key_count(const AssignGeom &g);
int row_keys_blocks(size_t n);
// partials: optional [row_keys_blocks(n)][6] per-block coordinate bounds
int launch_row_keys(const double *p, size_t n, const AssignGeom &g,
    uint32_t *keys, uint32_t *hist, double *partials, cudaStream_t st);
int launch_row_scatter(const double *p, size_t n, const uint32_t *keys,
    uint32_t *cursor, double *sorted, cudaStream_t st);
// mesh1 may be null (no interlacing).  precision: 8 or 4.
int launch_assign(const double *p, size_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *mesh0, void *mesh1, cudaStream_t st);
// owner-computes assignment (assign_tiles.cu): per-tile particle lists, fixed-point
// accumulation in shared memory, one plain store per mesh cell
bool tile_assign_supported(const AssignGeom &g);
void tile_set_fill_unroll(int u);
void tile_set_tma(int on);      // 0: flush tiles with thread stores instead of TMA tensor stores (ablation)
size_t tile_list_count(const AssignGeom &g);
int launch_tile_count(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cnt, double *partials, double *wmax_part, double *wmax, cudaStream_t st);
// index: the lists hold 4-byte particle indices (into p) instead of copies of the records
int launch_tile_fill(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cursor, void *lists, bool index, cudaStream_t st);
// cap > 0: one-pass lists (cap slots per tile, start[] holds the list lengths); ilist: index
// lists (then `lists` is the particle array they point into)
int launch_tile_accumulate(const double *lists, const uint32_t *start, const AssignGeom &g, int scheme,
    int precision, double wscale, const double *wmax, bool add, void *mesh0, void *mesh1, cudaStream_t st,
    uint32_t cap = 0, const uint32_t *ilist = nullptr);
// one-pass lists: fixed capacity per tile, overflow list (assign_tiles.cu)
struct TileOnePass {
  int index = 0;                // index lists
  uint32_t cap = 0, ovcap = 0;
  uint32_t *ovcount = nullptr;  // device counter, zeroed by the caller
  double *ovrec = nullptr;      // ovcap records
  uint32_t *ovtile = nullptr;   // ovcap tile indices
};
uint32_t tile_list_capacity(const AssignGeom &g, size_t n, int scheme, bool interlace);
int launch_tile_fill_onepass(const double *p, size_t n, const AssignGeom &g, int scheme, bool interlace,
    uint32_t *cnt, void *lists, const TileOnePass &t, double *partials, double *wmax_part, double *wmax,
    cudaStream_t st);
int launch_tile_overflow(const double *ovrec, const uint32_t *ovtile, uint32_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *mesh0, void *mesh1, cudaStream_t st);
int launch_unpad_copy(const void *mesh, void *dst, int ng, int rowlen,
    int precision, cudaStream_t st);
int launch_owner_keys(const double *p, size_t n, const AssignGeom &g, int nranks,
    uint32_t *keys, uint32_t *hist, cudaStream_t st);
int launch_owner_scatter(const double *p, size_t n, const uint32_t *keys, uint32_t *cursor,
    int nranks, double *sorted, cudaStream_t st);
int launch_add(void *dst, const void *src, size_t n, int precision, cudaStream_t st);

// ---------------------------------------------------------------------------
// Fourier-space binning (binning.cu)
// ---------------------------------------------------------------------------
struct BinGeom {
  int ng, ngk;
  int nbin, nl;
  int poles[8];
  int issim, logk, intlace;
  int j0, nj;           // local range of the middle (y) index of the k-space array:
                        // element (i, j0+jl, k) lives at ((i*nj + jl)*ngk + k)
  int symx, symy;       // mode counting may fold n_x / n_y (los component is zero)
  int symxy;            // ... and swap them (both folded, equal box sides)
  double los[3];
  double k0;            // kedge[0]
  double k1;            // kedge[nbin]
  double dk;
  double inv_dk;
  // device tables, each of length ng (axis 2: ngk entries used)
  const double *kax[3];         // k_a(n)        src/multipole.c:130-141
  const double *kax2[3];        // k_a(n)^2
  const double *wax[3];         // window factor src/multipole.c:46-100
  const double *pc[3], *ps[3];  // cos/sin(pi n / Ng), interlace phase, :462-484
  const double *k2edge;         // [nbin+1] smallest k^2 landing in each bin (host-bisected)
  const double *ztab;           // z axis packed per k: {k^2, window, cos, sin, k, 0}
  double legc[8][4];            // L_ell(x) = [x] (c0 + c1 x^2 + c2 x^4 + c3 x^6) per multipole
  int legodd[8];
  int anyodd;
  int ell, m;                   // survey l > 0: weight the product by Y_lm(k_hat)
  double ylm_nrm;
};

// geometry-only pass: cnt (u64), km (sum of |k| or log k), lcnt[nl][nbin]
int launch_geometry(const BinGeom &g, unsigned long long *cnt, double *km,
    double *lcnt, double *scratch, size_t scratch_bytes, cudaStream_t st);
// data pass: pl[nl][nbin] += sum Re(Fa conj Fb) alias mult L_l(mu)   (sims)
//            pl[nbin]     += sum Re(Fa conj Fb) alias mult           (surveys)
// Fa0/Fb0: fields on the base grid; Fa1/Fb1: shifted-grid fields combined on
// the fly (null if not interlaced or already combined).
int launch_bin(const BinGeom &g, int precision, const void *Fa0, const void *Fa1,
    const void *Fb0, const void *Fb1, double *pl, double *scratch,
    size_t scratch_bytes, cudaStream_t st);
size_t bin_scratch_bytes(const BinGeom &g);
void bin_set_geom_blocks(int n);
void bin_set_threads(int n);
// in-place interlace combination F0 <- (F0 + phase F1)/2 (survey l>0 needs the field)
int launch_combine(const BinGeom &g, int precision, void *F0, const void *F1,
    cudaStream_t st);
// survey multipoles (src/mp_template.c): out = Fr * Y_lm(r) ; Fkl += Fka * Y_lm(k)
struct YlmGeom {
  int ng, ngk, rowlen, ell, m;
  double smin[3];       // grid coordinate of the box corner (src/genr_mesh.c:913-914)
  double bsize[3];
};
int launch_ylm_weight_r(const YlmGeom &g, int precision, const void *Fr, void *out,
    cudaStream_t st);
int launch_ylm_accum_k(const YlmGeom &g, const BinGeom &bg, int precision,
    const void *Fka, void *Fkl, cudaStream_t st);
int launch_scale(void *mesh, size_t n, double factor, int precision, cudaStream_t st);
double ylm_norm(int l, int m);

// fft_strided.cu: in-place forward ng-point pass along y (axis 1, array (outer_n, ng, ngk))
// or x (axis 0, array (ng, outer_n, ngk)) of a complex array; optional tile skipping
// for the x pass (k2a[outer] + k2b[k] >= k2max)
bool fft_strided_supported(int ng, int precision);
// transposed output of the y pass (slab decomposition): ng / ny blocks of (outer, ny, ngk)
struct FftOut {
  static constexpr int MAXB = 16;
  void *base[MAXB] = {};        // block q = rows [q ny, (q+1) ny) of the transformed axis
  int ny = 0;                   // 0: in place
  size_t outer_stride = 0;      // elements between outer indices inside a block (ny * ngk)
};
// Outputs that nobody reads need not be stored: the binning drops every cell whose
// (k_x^2 + k_y^2) + k_z^2 is not below the last bin edge (src/multipole.c:151-159), and the
// x pass skips whole tiles by the same test.  (k2t[n] + k2o[o]) + k2k[kk] < k2max keeps
// output n of the transformed axis (o = outer index; kk = the column, or the first column
// of the tile when per_column = 0 — the granularity of the x pass's tile test, for the y
// pass that feeds it).  Same additions in the same order as binning.cu, so no cell the
// binning uses is ever dropped.  k2t = nullptr: store everything.
struct FftStoreSkip {
  const double *k2t = nullptr, *k2o = nullptr, *k2k = nullptr;
  double k2max = 0;
  int per_column = 0;
};
int launch_fft_strided_out(const void *data, int precision, int ng, int ngk, int outer_n,
    const FftOut &out, cudaStream_t st, const FftStoreSkip *ss = nullptr);
void fft_set_variant(int v);
int launch_fft_strided(void *data, int precision, int ng, int ngk, int axis, int outer_n,
    const double *k2a, const double *k2b, double k2max, cudaStream_t st, const FftStoreSkip *ss = nullptr);

// z + y passes of a whole mesh in one persistent kernel (L2-resident hand-over)
int launch_fft_zy(void *mesh, int precision, int ng, int ngk, int nplanes, int *done,
    cudaStream_t st);
// r2c transform of contiguous rows (the z pass)
int launch_fft_rows(const void *src, void *dst, int precision, int ng, long nrows,
    size_t src_pitch, size_t dst_pitch, cudaStream_t st);

// cnvt.cu: coordinate conversion (src/cnvt_coord.c)
void legauss_rule(int order, double *x, double *w);
int legauss_order(double om, double ol, double ok, double widx, double err, double zmin,
    double zmax, int num);
int cspline_second(const double *x, const double *y, size_t n, double *ypp);
int launch_cnvt_integr(double *p, size_t n, int order, double om, double ol, double ok,
    double widx, cudaStream_t st);
int launch_cnvt_interp(double *p, size_t n, const double *z, const double *d, const double *ypp,
    size_t nsp, cudaStream_t st);

// ingest.cu: binary catalogue ingest
int npy_probe(const char *path, size_t *offset, size_t *nrow, int *ncol, int *elem);
int assemble_blocks();
int launch_assemble(const void *raw, int elem, size_t n, const int pos[3], int wcomp, int wfkp, int nz,
    int ncol, int issim, double *rec, double *partial, cudaStream_t st);

// generate.cu
int launch_generate(double *out, size_t n, double boxsize, int kind, uint64_t seed,
    cudaStream_t st);
int launch_generate_at(double *out, size_t n, double boxsize, int kind, uint64_t seed,
    uint64_t first_index, cudaStream_t st);

}  // namespace psb
