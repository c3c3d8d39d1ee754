// Host orchestration and C ABI (include/powspec_b200.h) of the B200 hot path:
// device buffers, cuFFT plans, the genr_mesh / powspec stage logic, result
// normalisation.  Everything numerical runs in the kernels of assign.cu and
// binning.cu or in cuFFT; what is computed on the host here is O(Ng) tables and
// O(nl * nbin) normalisation, as in the reference.
//
// Reference map (paths relative to cheng-zhao/powspec):
//   mesh_init / mesh_destroy   src/genr_mesh.c:650-747, 615-640 -> Context buffers
//   def_box                    src/genr_mesh.c:509-578          -> define_box()
//   gen_dens / genr_mesh       src/genr_mesh.c:793-858, 874-926 -> psb_mesh()
//   FFTW plans / execution     src/fftw_define.h:32-64, src/multipole.c:444,459,493
//                                                               -> cuFFT D2Z/Z2D in place
//   powspec_init               src/multipole.c:306-421          -> init_bins()
//   alias_corr                 src/multipole.c:46-100           -> window_axis() tables
//   dens_k0, powspec, count_mode  src/multipole.c:435-505, 1179-1278, 1044-1162
//                                                               -> psb_power()

#include "psb_context.h"

#include <unistd.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <fcntl.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace psb {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static thread_local bool g_quiet = false;       // record the message but do not print it

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  // the reference's P_ERR format, src/define.h:102,129
  if (!g_quiet) fprintf(stderr, "\n\x1B[31;1mError:\x1B[0m %s", g_err);
}
struct QuietErrors {
  QuietErrors() { g_quiet = true; }
  ~QuietErrors() { g_quiet = false; }
};
const char *get_error() { return g_err; }
void errors_quiet(bool on) { g_quiet = on; }

static const double PI = 0x1.921fb54442d18p+1;  // src/define.h:37

#define PSB_CUFFT(call)                                                         \
  do {                                                                          \
    cufftResult r_ = (call);                                                    \
    if (r_ != CUFFT_SUCCESS) {                                                  \
      psb::set_error("cuFFT failure %s at %s:%d: code %d\n", #call, __FILE__,   \
          __LINE__, (int) r_);                                                  \
      return -1;                                                                \
    }                                                                           \
  } while (0)

}  // namespace psb

using namespace psb;

namespace psb_host {


cudaEvent_t get_event(psb_context *c) {
  cudaEvent_t e;
  if (!c->evpool.empty()) { e = c->evpool.back(); c->evpool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}


void reset_timings(psb_context *c) {
  for (auto &iv : c->intervals) { c->evpool.push_back(iv.a); c->evpool.push_back(iv.b); }
  c->intervals.clear();
  for (double &m : c->ms) m = 0;
  c->host_h2d_ms = 0;
}

// PSB_TRACE=1: host wall-clock marks on stderr (ms since the first mark), to see
// where a run's time goes between the device stages (tools / debugging only)
void trace_mark(const char *what) {
  static const bool on = getenv("PSB_TRACE") != nullptr;
  if (!on) return;
  static struct timespec t0 = {0, 0};
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  if (!t0.tv_sec && !t0.tv_nsec) t0 = t;
  fprintf(stderr, "[psb-trace] %9.3f ms  %s\n",
      (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6, what);
}

void collect_timings(psb_context *c) {
  // PSB_TRACE: device-side timeline of every stage interval of the run (all streams),
  // in ms since the first one started
  static const bool trace = getenv("PSB_TRACE") != nullptr;
  if (trace && !c->intervals.empty()) {
    static const char *names[PSB_T_COUNT] = {"h2d", "bounds", "sort", "memset", "assign", "fft",
        "geom", "bin", "ylm", "fft_strided", "cnvt", "total"};
    cudaEvent_t origin = c->intervals[0].a;
    for (auto &iv : c->intervals) {
      float t0 = 0, t1 = 0;
      if (cudaEventElapsedTime(&t0, origin, iv.a) == cudaSuccess &&
          cudaEventElapsedTime(&t1, origin, iv.b) == cudaSuccess)
        fprintf(stderr, "[psb-trace]   device %-11s %9.3f -> %9.3f ms\n", names[iv.stage], t0, t1);
    }
    cudaGetLastError();
  }
  for (auto &iv : c->intervals) {
    float t = 0;
    if (cudaEventElapsedTime(&t, iv.a, iv.b) == cudaSuccess) c->ms[iv.stage] += t;
    c->evpool.push_back(iv.a);
    c->evpool.push_back(iv.b);
  }
  c->intervals.clear();
}

int check_params(const psb_params *p) {
  if (!p) { set_error("configuration parameters not loaded\n"); return -1; }
  if (p->ncat < 1 || p->ncat > 2) { set_error("invalid number of catalogs: %d\n", p->ncat); return -1; }
  if (p->assign < 0 || p->assign > 3) {
    set_error("unrecognised particle assignment scheme: %d\n", p->assign); return -1;
  }
  if (p->gsize < 2 || p->gsize > 65536) {       // POWSPEC_MAX_GSIZE, src/define.h:66
    set_error("invalid GRID_SIZE: %d\n", p->gsize); return -1;
  }
  if (p->npole < 1 || p->npole > 7) { set_error("invalid number of multipoles: %d\n", p->npole); return -1; }
  for (int i = 0; i < p->npole; i++)
    if (p->poles[i] < 0 || p->poles[i] > PSB_MAX_ELL || (i && p->poles[i] <= p->poles[i - 1])) {
      set_error("multipoles must be sorted, unique and <= %d\n", PSB_MAX_ELL); return -1;
    }
  if (p->precision != 8 && p->precision != 4) { set_error("precision must be 8 or 4\n"); return -1; }
  if (p->issim && !p->has_bsize) { set_error("BOX_SIZE is required for simulation boxes\n"); return -1; }
  if (!(p->kbin > 0)) { set_error("invalid BIN_SIZE\n"); return -1; }
  return 0;
}

// def_box, src/genr_mesh.c:509-578
int define_box(const psb_params *p, const double lo[3], const double hi[3], double bmin[3],
    double bsize[3]) {
  const char ax[3] = {'x', 'y', 'z'};
  for (int a = 0; a < 3; a++) {
    if (lo[a] > hi[a]) { set_error("invalid %c coordinate value in the catalogs\n", ax[a]); return -1; }
    if (p->issim) {
      if (lo[a] < 0) { set_error("%c coordinate below 0: %lf\n", ax[a], lo[a]); return -1; }
      if (hi[a] >= p->bsize[a]) {
        set_error("%c coordinate not smaller than BOX_SIZE: %lf\n", ax[a], hi[a]); return -1;
      }
    }
  }
  for (int a = 0; a < 3; a++) {
    if (p->issim) { bsize[a] = p->bsize[a]; bmin[a] = 0; }
    else if (p->has_bsize) {
      if (hi[a] - lo[a] > p->bsize[a]) {
        set_error("BOX_SIZE is too small for the %c coordinates, should be at least %.10lg\n",
            ax[a], hi[a] - lo[a]);
        return -1;
      }
      bmin[a] = (hi[a] + lo[a] - p->bsize[a]) * 0.5;
      bsize[a] = p->bsize[a];
    }
    else {
      bsize[a] = ceil((hi[a] - lo[a]) * (1 + p->bpad[a]) / 10) * 10;    // POWSPEC_BOX_CEIL
      bmin[a] = (hi[a] + lo[a] - bsize[a]) * 0.5;
    }
  }
  if (p->verbose && !p->issim)
    printf("  Box size: [%lg, %lg, %lg]\n"
        "  Box boundaries: [[%lg,%lg], [%lg,%lg], [%lg,%lg]]\n", bsize[0], bsize[1], bsize[2],
        bmin[0], bmin[0] + bsize[0], bmin[1], bmin[1] + bsize[1], bmin[2], bmin[2] + bsize[2]);
  return 0;
}

int h2d_async(psb_context *c, void *dst, const void *src, size_t bytes, bool pinned,
    cudaStream_t stream);
bool is_pinned(const void *p);

// host -> device copy of a whole particle array (surveys: the box depends on
// the bounds of all catalogues, so they are made resident first)
int upload(psb_context *c, const double *src, size_t n, DevBuf &dst) {
  const size_t bytes = n * 32;
  if (dst.reserve(bytes ? bytes : 32)) return -1;
  if (!n) return 0;
  StageScope sc(c, PSB_T_H2D, c->st);
  return h2d_async(c, dst.p, src, bytes, is_pinned(src), c->st);
}

int coordinate_bounds(psb_context *c, const double *dev, size_t n, double lo[3], double hi[3]) {
  if (!n) return 0;
  const int nblk = c->sms * 8;
  if (c->bounds_part.reserve(sizeof(double) * 6 * nblk)) return -1;
  {
    StageScope sc(c, PSB_T_BOUNDS, c->st);
    if (launch_bounds(dev, n, c->bounds_part.as<double>(), nblk, c->st)) return -1;
    c->launches++;
  }
  std::vector<double> h(6 * (size_t) nblk);
  PSB_CUDA(cudaMemcpyAsync(h.data(), c->bounds_part.p, h.size() * sizeof(double),
      cudaMemcpyDeviceToHost, c->st));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  for (int b = 0; b < nblk; b++)
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], h[6 * b + a]);
      hi[a] = std::max(hi[a], h[6 * b + 3 + a]);
    }
  return 0;
}

// cnvt_coord (src/cnvt_coord.c:440-582) on device-resident arrays, in place
int convert_arrays(psb_context *c, const psb_cosmo *cm, double *const *arr, const size_t *cnt,
    int narr, int *order_out) {
  if (order_out) *order_out = 0;
  if (cm->sample_z && cm->sample_d) {
    // interpolation of the tabulated distances (cnvt_coord_interp, :440-486)
    const size_t nsp = cm->nsample;
    std::vector<double> tab(3 * nsp);
    if (nsp < 2 || cspline_second(cm->sample_z, cm->sample_d, nsp, &tab[2 * nsp])) {
      set_error("failed to interpolate the sample points\n");
      return -1;
    }
    memcpy(&tab[0], cm->sample_z, nsp * sizeof(double));
    memcpy(&tab[nsp], cm->sample_d, nsp * sizeof(double));
    if (c->cnvt_tab.reserve(tab.size() * sizeof(double))) return -1;
    PSB_CUDA(cudaMemcpyAsync(c->cnvt_tab.p, tab.data(), tab.size() * sizeof(double),
        cudaMemcpyHostToDevice, c->st));
    PSB_CUDA(cudaStreamSynchronize(c->st));
    const double *t = c->cnvt_tab.as<double>();
    StageScope sc(c, PSB_T_CNVT, c->st);
    for (int i = 0; i < narr; i++) {
      if (launch_cnvt_interp(arr[i], cnt[i], t, t + nsp, t + 2 * nsp, nsp, c->st)) return -1;
      c->launches++;
    }
    return 0;
  }
  // Legendre-Gauss integration (cnvt_coord_integr, :495-537): redshift range of
  // all the catalogues (cnvt_z_sample, :160-280), order, conversion
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int i = 0; i < narr; i++)
    if (coordinate_bounds(c, arr[i], cnt[i], lo, hi)) return -1;
  if (lo[2] < 0) {
    set_error("invalid negative redshift in the catalogs: %g\n", lo[2]);
    return -1;
  }
  if (lo[2] > hi[2]) { set_error("invalid redshift value in the catalogs\n"); return -1; }
  const double widx = (cm->eos_w == -1) ? 0 : 3 * (1 + cm->eos_w);
  const int order = legauss_order(cm->omega_m, cm->omega_l, cm->omega_k, widx, cm->ecdst, lo[2],
      hi[2], 128 /* POWSPEC_INT_NUM_ZSP, src/define.h:93 */);
  if (order == INT_MAX) {
    set_error("failed to perform the convergency test for integrations\n");
    return -1;
  }
  if (order_out) *order_out = order;
  StageScope sc(c, PSB_T_CNVT, c->st);
  for (int i = 0; i < narr; i++) {
    if (launch_cnvt_integr(arr[i], cnt[i], order, cm->omega_m, cm->omega_l, cm->omega_k, widx, c->st))
      return -1;
    c->launches++;
  }
  return 0;
}

// counting sort by mesh row, then the scatter.  Chunked so that 32-bit offsets
// suffice and the scratch stays bounded.
// One device-resident chunk: counting sort by mesh row, then the scatter.
// `consumed` (optional) is recorded once the source buffer is no longer read.
// `bounds` (optional): the chunk's coordinate bounds are appended to
// c->bounds_part (6 doubles per block); bounds_finish() reduces them.
int zero_if_fresh(psb_context *c, const AssignGeom &g, int precision, void *m0, void *m1, bool *fresh) {
  if (!fresh || !*fresh) return 0;
  const size_t bytes = (size_t) g.nxloc * g.ng * g.rowlen * precision;
  for (void *m : {m0, m1}) {
    if (!m) continue;
    StageScope sc(c, PSB_T_MEMSET, c->st);
    PSB_CUDA(cudaMemsetAsync(m, 0, bytes, c->st));
  }
  *fresh = false;
  return 0;
}

// Owner-computes path (assign_tiles.cu): tile lists, then one block per tile accumulates in
// shared memory and writes the tile once.  The lists are built in ONE pass into
// fixed-capacity slots (no count pass, no scan) when the catalogue lets them: entries beyond
// a tile's capacity go to an overflow list that is added with global atomics afterwards (slow:
// ~0.25 ms per million entries); a catalogue that overflows more than 1/64 of its particles
// (clustered on the tile scale: measured 50.7 vs 46.0 ms per step on the clustered bench
// catalogue when everything that overflowed went through that list) is redone with the exact
// count + scan + fill, and the next 16 dense chunks go there directly.  Uniform catalogues:
// list stage 6.27 -> 5.73 ms.
int tile_lists_exact(psb_context *c, const double *src, size_t len, const AssignGeom &g, int scheme,
    bool interlace, double *partials, double *wmax, size_t ntile, bool index) {
  size_t tmp_bytes = c->tile_scan_bytes;
  uint32_t total = 0;
  PSB_CUDA(cudaMemsetAsync(c->tile_cnt.p, 0, (ntile + 1) * 4, c->st));
  if (launch_tile_count(src, len, g, scheme, interlace, c->tile_cnt.as<uint32_t>(), partials, wmax + 1, wmax, c->st))
    return -1;
  PSB_CUDA(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp_bytes, c->tile_cnt.as<uint32_t>(),
      c->tile_start.as<uint32_t>(), (int) ntile + 1, c->st));
  // the list buffer is sized from the total: one small host wait per chunk
  PSB_CUDA(cudaMemcpyAsync(&total, c->tile_start.as<uint32_t>() + ntile, 4, cudaMemcpyDeviceToHost, c->st));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  if (c->sorted.reserve((size_t) (total ? total : 1) * (index ? 4 : 32))) return -1;
  PSB_CUDA(cudaMemcpyAsync(c->tile_cnt.p, c->tile_start.p, (ntile + 1) * 4, cudaMemcpyDeviceToDevice, c->st));
  if (launch_tile_fill(src, len, g, scheme, interlace, c->tile_cnt.as<uint32_t>(), c->sorted.p, index, c->st))
    return -1;
  c->launches += 6;
  return 0;
}

int tile_assign_chunk(psb_context *c, const double *src, size_t len, const AssignGeom &g,
    int scheme, int precision, double wscale, void *m0, void *m1, cudaEvent_t consumed,
    double *partials, bool *fresh) {
  const size_t ntile = tile_list_count(g);
  const int nblk = row_keys_blocks(len);
  const bool interlace = m1 != nullptr;
  if (c->tile_cnt.reserve((ntile + 1) * 4) || c->tile_start.reserve((ntile + 1) * 4) ||
      c->wmax_buf.reserve(sizeof(double) * (nblk + 1)))
    return -1;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->tile_cnt.as<uint32_t>(), c->tile_start.as<uint32_t>(),
      (int) ntile + 1, c->st);
  if (c->cubtmp.reserve(tmp_bytes)) return -1;
  c->tile_scan_bytes = tmp_bytes;
  double *wmax = c->wmax_buf.as<double>();
  // one-pass lists: capacity per tile, overflow room for 1/64 of the particles
  uint32_t cap = c->opt_tile_cap > 0 ? (uint32_t) c->opt_tile_cap : tile_list_capacity(g, len, scheme, interlace);
  // Index lists (option tile_index, off): the scattered 32-byte record stores of the fill pass
  // run at 1.4 TB/s — every store opens another DRAM row (tools/fill_probe.cu: 3.3 ms for the
  // stores alone, 1.3 ms for the returning atomics alone) — while 4-byte indices merge in the
  // L2.  But then the accumulation has to gather the records, one random DRAM sector each,
  // behind a dependent load: measured list stage 5.6 -> 5.1 ms, accumulation 10.7 -> 15.4 ms.
  // The copies stay: the random access is paid once, as a store, where nothing waits for it.
  const bool index = c->opt_tile_index != 0;
  const size_t list_bytes = ntile * (size_t) cap * (index ? 4 : 32);
  bool onepass = c->opt_tile_onepass != 0 && c->tile_onepass_backoff == 0 && list_bytes <= ((size_t) 24 << 30);
  if (c->tile_onepass_backoff > 0) c->tile_onepass_backoff--;
  uint32_t nover = 0;
  TileOnePass op;
  {
    StageScope sc(c, PSB_T_SORT, c->st);
    if (onepass) {
      op.cap = cap;
      op.index = index;
      op.ovcap = c->opt_tile_ovcap > 0 ? (uint32_t) c->opt_tile_ovcap : (uint32_t) std::max<size_t>(4096, len / 64);
      if (c->sorted.reserve(list_bytes) || c->tile_ovrec.reserve((size_t) op.ovcap * 32) ||
          c->tile_ovtile.reserve((size_t) op.ovcap * 4 + 16))
        return -1;
      op.ovrec = c->tile_ovrec.as<double>();
      op.ovtile = c->tile_ovtile.as<uint32_t>();
      op.ovcount = op.ovtile + op.ovcap;        // the counter lives behind the tile indices
      PSB_CUDA(cudaMemsetAsync(c->tile_cnt.p, 0, (ntile + 1) * 4, c->st));
      PSB_CUDA(cudaMemsetAsync(op.ovcount, 0, 4, c->st));
      if (launch_tile_fill_onepass(src, len, g, scheme, interlace, c->tile_cnt.as<uint32_t>(), c->sorted.p,
            op, partials, wmax + 1, wmax, c->st))
        return -1;
      c->launches += 3;
      // how much overflowed: the same one small host wait the exact path spends on its total
      PSB_CUDA(cudaMemcpyAsync(&nover, op.ovcount, 4, cudaMemcpyDeviceToHost, c->st));
      PSB_CUDA(cudaStreamSynchronize(c->st));
      if (nover > op.ovcap) {                   // clustered on the tile scale: exact lists
        onepass = false;
        c->tile_onepass_backoff = 16;
      }
    }
    if (!onepass && tile_lists_exact(c, src, len, g, scheme, interlace, partials, wmax, ntile, index)) return -1;
  }
  // with index lists the accumulation still reads the source records
  if (consumed && !index) PSB_CUDA(cudaEventRecord(consumed, c->st));
  if (c->memset_pending) { PSB_CUDA(cudaStreamWaitEvent(c->st, c->memset_pending, 0)); c->memset_pending = nullptr; }
  const bool add = !(fresh && *fresh);
  if (fresh) *fresh = false;
  StageScope sc(c, PSB_T_ASSIGN, c->st);
  if (launch_tile_accumulate(index ? src : c->sorted.as<double>(),
        onepass ? c->tile_cnt.as<uint32_t>() : c->tile_start.as<uint32_t>(), g, scheme, precision, wscale, wmax, add,
        m0, m1, c->st, onepass ? cap : 0, index ? c->sorted.as<uint32_t>() : nullptr))
    return -1;
  c->launches++;
  if (consumed && index) PSB_CUDA(cudaEventRecord(consumed, c->st));
  if (onepass && nover) {
    if (launch_tile_overflow(op.ovrec, op.ovtile, nover, g, scheme, precision, wscale, m0, m1, c->st)) return -1;
    c->launches++;
  }
  c->assign_path = 1;
  c->tile_overflowed = nover;
  return 0;
}

int sort_assign_chunk(psb_context *c, const double *src, size_t len, const AssignGeom &g,
    int scheme, int precision, double wscale, void *m0, void *m1, cudaEvent_t consumed,
    bool bounds = false, bool *fresh = nullptr) {
  if (!len) return 0;
  const bool tiles = c->opt_owner != 0 && tile_assign_supported(g) && len < ((size_t) 1 << 29) &&
      (c->opt_owner > 0 || len >= (size_t) g.ng * g.ng * g.ng / 32);
  if (tiles) {
    double *partials = nullptr;
    if (bounds) {
      const int nblk = row_keys_blocks(len);
      if (c->bounds_part.reserve_keep(c->bounds_used + sizeof(double) * 6 * nblk, c->bounds_used, c->st)) return -1;
      partials = reinterpret_cast<double *>(c->bounds_part.as<char>() + c->bounds_used);
      c->bounds_used += sizeof(double) * 6 * nblk;
    }
    return tile_assign_chunk(c, src, len, g, scheme, precision, wscale, m0, m1, consumed, partials, fresh);
  }
  if (zero_if_fresh(c, g, precision, m0, m1, fresh)) return -1;
  c->assign_path = 0;
  const bool do_sort = c->opt_sort && len >= (size_t) c->opt_sort_min;
  double *partials = nullptr;
  if (bounds) {
    const int nblk = do_sort ? row_keys_blocks(len) : c->sms * 8;
    if (c->bounds_part.reserve_keep(c->bounds_used + sizeof(double) * 6 * nblk, c->bounds_used, c->st)) return -1;
    partials = reinterpret_cast<double *>(c->bounds_part.as<char>() + c->bounds_used);
    c->bounds_used += sizeof(double) * 6 * nblk;
    if (!do_sort) {
      StageScope sc(c, PSB_T_BOUNDS, c->st);
      if (launch_bounds(src, len, partials, nblk, c->st)) return -1;
      c->launches++;
    }
  }
  if (!do_sort) {
    if (c->memset_pending) { PSB_CUDA(cudaStreamWaitEvent(c->st, c->memset_pending, 0)); c->memset_pending = nullptr; }
    StageScope sc(c, PSB_T_ASSIGN, c->st);
    c->launches++;
    if (launch_assign(src, len, g, scheme, precision, wscale, m0, m1, c->st)) return -1;
    if (consumed) PSB_CUDA(cudaEventRecord(consumed, c->st));
    return 0;
  }
  const size_t nrow = row_key_count(g);
  if (nrow > 0x7fffffffull) { set_error("GRID_SIZE too large for the row sort\n"); return -1; }
  if (c->keys.reserve(len * 4) || c->sorted.reserve(len * 32) ||
      c->hist.reserve(nrow * 4) || c->cursor.reserve(nrow * 4))
    return -1;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->hist.as<uint32_t>(),
      c->cursor.as<uint32_t>(), (int) nrow, c->st);
  if (c->cubtmp.reserve(tmp_bytes)) return -1;
  {
    StageScope sc(c, PSB_T_SORT, c->st);
    PSB_CUDA(cudaMemsetAsync(c->hist.p, 0, nrow * 4, c->st));
    if (launch_row_keys(src, len, g, c->keys.as<uint32_t>(), c->hist.as<uint32_t>(), partials, c->st))
      return -1;
    PSB_CUDA(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp_bytes, c->hist.as<uint32_t>(),
        c->cursor.as<uint32_t>(), (int) nrow, c->st));
    if (launch_row_scatter(src, len, c->keys.as<uint32_t>(), c->cursor.as<uint32_t>(),
          c->sorted.as<double>(), c->st))
      return -1;
    c->launches += 5;
  }
  if (consumed) PSB_CUDA(cudaEventRecord(consumed, c->st));
  if (c->memset_pending) { PSB_CUDA(cudaStreamWaitEvent(c->st, c->memset_pending, 0)); c->memset_pending = nullptr; }
  StageScope sc(c, PSB_T_ASSIGN, c->st);
  if (launch_assign(c->sorted.as<double>(), len, g, scheme, precision, wscale, m0, m1, c->st))
    return -1;
  c->launches++;
  return 0;
}

// A catalogue that is already on the device.  Chunked so that 32-bit offsets
// suffice and the sort scratch stays bounded.
const size_t DEV_CHUNK = (size_t) 1 << 28;      // particles per chunk (8.6 GB of records)

int assign_catalog(psb_context *c, const double *dev, size_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *m0, void *m1, bool bounds, bool *fresh) {
  for (size_t off = 0; off < n; off += DEV_CHUNK)
    if (sort_assign_chunk(c, dev + 4 * off, std::min(DEV_CHUNK, n - off), g, scheme, precision,
          wscale, m0, m1, nullptr, bounds, fresh))
      return -1;
  return 0;
}

// room for the deferred bounds partials of `nchunk` chunks
int bounds_begin(psb_context *c, size_t nchunk) {
  c->bounds_used = 0;
  // sized from the grids the key pass really launches (row_keys_blocks does not depend on
  // the SM count); sort_assign_chunk grows the buffer, keeping its contents, if needed
  return c->bounds_part.reserve(sizeof(double) * 6 * (size_t) (row_keys_blocks((size_t) 1 << 40) + 64) * (nchunk + 1));
}

// def_box's checks for simulation boxes (src/genr_mesh.c:516-531), evaluated
// after the catalogue has been scattered
int bounds_finish(psb_context *c, const psb_params *par) {
  std::vector<double> h(c->bounds_used / sizeof(double));
  if (!h.empty()) {
    PSB_CUDA(cudaMemcpyAsync(h.data(), c->bounds_part.p, c->bounds_used, cudaMemcpyDeviceToHost, c->st));
    PSB_CUDA(cudaStreamSynchronize(c->st));
  }
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t q = 0; q + 5 < h.size(); q += 6)
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], h[q + a]);
      hi[a] = std::max(hi[a], h[q + 3 + a]);
    }
  for (int a = 0; a < 3; a++) c->bmax[a] = hi[a];
  double bmin_chk[3], bsize_chk[3];
  return define_box(par, lo, hi, bmin_chk, bsize_chk);
}

// Copy `bytes` from host memory into a device buffer on the copy stream.
// Pinned sources go straight to the copy engine; pageable ones are staged
// through a ring of pinned buffers filled by a persistent pool of host threads
// (hostcopy.cpp), so the PCIe transfer is not bound by one memcpy thread.
int h2d_async(psb_context *c, void *dst, const void *src, size_t bytes, bool pinned,
    cudaStream_t stream) {
  if (!bytes) return 0;
  if (pinned) {
    PSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  // staging ring: `nslot` pinned pieces of `CH` bytes (one allocation)
  const size_t CH = (size_t) std::min<long>(std::max<long>(c->opt_h2d_piece_mb, 1), 1024) << 20;
  const int nslot = (int) std::min<long>(std::max<long>(c->opt_h2d_slots, 2), PSB_STAGE_SLOTS);
  const bool wc = c->opt_h2d_wc != 0;
  if (c->pinned_base && (c->pinned_wc != wc || c->pinned_bytes != CH || c->pinned_slots != nslot)) {
    // an option changed: wait for the ring's last copies, then start over
    for (int i = 0; i < PSB_STAGE_SLOTS; i++)
      if (c->pinned_free[i]) PSB_CUDA(cudaEventSynchronize(c->pinned_free[i]));
    PSB_CUDA(cudaFreeHost(c->pinned_base));
    c->pinned_base = nullptr;
  }
  if (!c->pinned_base) {
    // (write-combined memory: only ever written by the cores and read by the DMA engine, which
    // then need not snoop the caches — measured equal on the B200 hosts, default off)
    PSB_CUDA(cudaHostAlloc(&c->pinned_base, CH * nslot, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    for (int i = 0; i < nslot; i++)
      if (!c->pinned_free[i]) PSB_CUDA(cudaEventCreateWithFlags(&c->pinned_free[i], cudaEventDisableTiming));
    c->pinned_bytes = CH;
    c->pinned_slots = nslot;
    c->pinned_wc = wc;
    c->pinned_next = 0;
  }
  unsigned hw = std::thread::hardware_concurrency();
  const int nthr = (int) std::max(1u, std::min((unsigned) std::max<long>(c->opt_h2d_threads, 1), hw ? hw : 1u));
  // the staging threads live as long as the context (round 1 created and joined them for
  // every 64 MB piece); the copy streams past the caches (hostcopy.cpp)
  if (!c->copy_pool || copy_pool_threads(c->copy_pool) != nthr) {
    copy_pool_destroy(c->copy_pool);
    c->copy_pool = copy_pool_create(nthr);
  }
  size_t off = 0;
  while (off < bytes) {
    const size_t len = std::min(CH, bytes - off);
    const int slot = c->pinned_next;
    c->pinned_next = (slot + 1) % nslot;
    PSB_CUDA(cudaEventSynchronize(c->pinned_free[slot]));
    char *stage = static_cast<char *>(c->pinned_base) + (size_t) slot * CH;
    const char *from = static_cast<const char *>(src) + off;
    copy_pool_run(c->copy_pool, stage, from, len);
    PSB_CUDA(cudaMemcpyAsync(static_cast<char *>(dst) + off, stage, len,
        cudaMemcpyHostToDevice, stream));
    PSB_CUDA(cudaEventRecord(c->pinned_free[slot], stream));
    off += len;
  }
  return 0;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  bool pinned = false;
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess)
    pinned = (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  return pinned;
}

// A catalogue in HOST memory, box known in advance (simulation boxes): chunks
// are uploaded on the copy stream while the previous chunk is being sorted and
// scattered, so PCIe and the SMs work concurrently and the catalogue never has
// to be resident as a whole.  Bounds partials of every chunk are left on the
// device (bounds_finish reduces them after the stream has drained).
//
// Chunk schedule.  Option stream_taper > 0 lets the last chunks halve down to that
// many particles, so that less is left to scatter once the last byte has arrived.
// Measured on B200 (config 2 from pinned memory): 94.9 ms per run against 92.6 ms with
// equal chunks — scattering a chunk costs one sweep of both meshes through DRAM
// (~6 ms) almost regardless of its size down to a few million particles, and below
// that ~1 ms per million (random read-modify-write of single sectors), slower than
// the 0.58 ms per million of the upload: small chunks queue up behind each other.
// Kept as an ablation, off by default.
std::vector<size_t> stream_schedule(const psb_context *c, size_t n) {
  const size_t CH = (size_t) std::max<long>(c->opt_stream_chunk, 1 << 16);
  const size_t tail = (size_t) std::max<long>(c->opt_stream_taper, 0);
  std::vector<size_t> len;
  for (size_t left = n; left; left -= len.back()) {
    size_t l = std::min(CH, left);
    if (tail && left > tail) l = std::min(l, std::max(tail, left / 2));
    len.push_back(l);
  }
  return len;
}

int stream_catalog(psb_context *c, const double *host, size_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *m0, void *m1, bool *fresh) {
  if (!n) return 0;
  const std::vector<size_t> sched = stream_schedule(c, n);
  const size_t chunk_max = *std::max_element(sched.begin(), sched.end());
  const size_t nchunk = sched.size();
  const bool pinned = is_pinned(host);
  for (int s = 0; s < 2; s++) {
    if (nchunk > (size_t) s && c->chunkbuf[s].reserve(chunk_max * 32)) return -1;
    if (!c->ev_filled[s]) {
      PSB_CUDA(cudaEventCreateWithFlags(&c->ev_filled[s], cudaEventDisableTiming));
      PSB_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[s], cudaEventDisableTiming));
    }
  }
  // The chunk buffers' last readers recorded ev_consumed[] (previous catalogue or
  // run; a never-recorded event is a no-op to wait on), so the first upload does
  // not have to wait for the mesh memsets queued on the compute stream.
  size_t off = 0;
  for (size_t k = 0; k < nchunk; off += sched[k], k++) {
    const size_t len = sched[k];
    const int s = (int) (k & 1);
    double *buf = c->chunkbuf[s].as<double>();
    PSB_CUDA(cudaStreamWaitEvent(c->st_copy, c->ev_consumed[s], 0));
    {
      StageScope sc(c, PSB_T_H2D, c->st_copy);
      if (h2d_async(c, buf, host + 4 * off, len * 32, pinned, c->st_copy)) return -1;
    }
    PSB_CUDA(cudaEventRecord(c->ev_filled[s], c->st_copy));
    PSB_CUDA(cudaStreamWaitEvent(c->st, c->ev_filled[s], 0));
    if (sort_assign_chunk(c, buf, len, g, scheme, precision, wscale, m0, m1, c->ev_consumed[s], true, fresh))
      return -1;
  }
  return 0;
}

// alias_corr, src/multipole.c:46-100 — evaluated on the host with libm for the
// 3 x Ng distinct arguments; the signed `x > 0.01` test reproduces quirk Q1.
double window_axis(int intlace, int scheme, double x) {
  if (intlace) {
    double f;
    if (x > 0.01) f = x / sin(x);
    else {
      const double t = x * x;
      f = 1 + 0x1.5555555555555p-3 * t + 0x1.3e93e93e93e94p-6 * t * t
        + 0x1.0cbb766210cbbp-9 * t * t * t;
    }
    f *= f;
    switch (scheme) {
      case 0: return f;
      case 1: return f * f;
      case 2: return f * f * f;
      default: f *= f; return f * f;
    }
  }
  double s;
  if (x > 0.01) s = sin(x);
  else {
    const double t = x * x;
    s = (1 - 0x1.5555555555555p-3 * t + 0x1.1111111111111p-7 * t * t
        - 0x1.a01a01a01a01ap-13 * t * t * t) * x;
  }
  s *= s;
  switch (scheme) {
    case 0: return 1;
    case 1: return 1 / (1 - 0x1.5555555555555p-1 * s);
    case 2: return 1 / (1 - s + 0x1.1111111111111p-3 * s * s);
    default: return 1 / (1 - 0x1.5555555555555p+0 * s + 0.4 * s * s
                 - 0x1.a01a01a01a01ap-7 * s * s * s);
  }
}

// the reference's bin decision, src/multipole.c:145-159 (linear: sqrt; log: glibc
// log10), evaluated on the host with exactly the reference's operations
int ref_bin(double k2, bool logk, double k0, double k1, double dk, int nbin) {
  const double kc = logk ? 0.5 * log10(k2) : sqrt(k2);
  if (kc < k0 || kc >= k1) return -1;
  const int b = (int) ((kc - k0) / dk);
  return (b < 0 || b >= nbin) ? -1 : b;
}

// smallest positive double x with pred(x) true, pred monotone false -> true
template <typename F> double bisect_first(F pred) {
  if (pred(0.0)) return 0.0;
  uint64_t lo = 1, hi = 0x7fefffffffffffffull;  // smallest subnormal .. DBL_MAX
  auto val = [](uint64_t u) { double d; memcpy(&d, &u, 8); return d; };
  if (!pred(val(hi))) return INFINITY;
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (pred(val(mid))) hi = mid; else lo = mid + 1;
  }
  return val(lo);
}

// Which library serves which pass is decided from the per-pass measurements on B200
// (tools/fft_pass_bench.py, profiles/README.md): the strided y pass is always ours;
// the r2c z pass is ours except where cuFFT's power-of-two kernels are faster (1024,
// and 512 in double); the x pass, whose points lie a whole plane apart, is ours in
// double up to 1536 and in single precision up to 1024 (where skipping the tiles
// beyond the last bin edge outweighs cuFFT's faster float kernel), cuFFT's above.
bool fft_own_z(const psb_context *c, int ng, int precision) {
  if (c->opt_fft_own_z >= 0) return c->opt_fft_own_z != 0;
  return !(ng == 1024 || (ng == 512 && precision == 8));
}

bool fft_own_x(const psb_context *c, int ng, int precision) {
  if (c->opt_fft_own_x >= 0) return c->opt_fft_own_x != 0;
  return precision == 8 ? ng <= 1536 : ng <= 1024;
}

// planes per group of the L2-blocked z + y passes: the largest divisor of ng whose
// planes (complex, in place) fit the L2 budget; ng when blocking is off
static int fft_group_planes(const psb_context *c, int ng, int precision) {
  if (c->opt_fft_l2_mb <= 0) return ng;
  const size_t plane = (size_t) ng * (ng / 2 + 1) * 2 * precision;
  int best = 1;
  for (int p = 1; p <= ng; p++)
    if (ng % p == 0 && p * plane <= ((size_t) c->opt_fft_l2_mb << 20)) best = p;
  return best;
}

int ensure_plans(psb_context *c, int ng, int precision, bool need_inv) {
  const bool own = c->opt_own_fft && fft_strided_supported(ng, precision);
  const int zp = own ? fft_group_planes(c, ng, precision) : 0;
  if (c->plan_ng != ng || c->plan_prec != precision || c->plan_zp != zp) {
    if (c->have_fwd) cufftDestroy(c->plan_fwd);
    if (c->have_inv) cufftDestroy(c->plan_inv);
    if (c->have_z) cufftDestroy(c->plan_z);
    if (c->have_z2) cufftDestroy(c->plan_z2);
    if (c->have_x) cufftDestroy(c->plan_x);
    c->have_fwd = c->have_inv = c->have_z = c->have_z2 = c->have_x = false;
    c->plan_ng = ng; c->plan_prec = precision; c->plan_zp = zp;
  }
  c->own_fft = own;
  const int ngk = ng / 2 + 1;
  long long n[3] = {ng, ng, ng};
  long long rembed[3] = {ng, ng, 2LL * ngk}, cembed[3] = {ng, ng, ngk};
  const long long rdist = (long long) ng * ng * 2 * ngk, cdist = (long long) ng * ng * ngk;
  size_t ws_f = 0, ws_i = 0, ws_z = 0;
  // cuFFT's 3-D plan only where the hand-written strided passes do not apply (its
  // work area is as large as the mesh for sizes that are not powers of two)
  if (!own && !c->have_fwd) {
    PSB_CUFFT(cufftCreate(&c->plan_fwd));
    PSB_CUFFT(cufftSetAutoAllocation(c->plan_fwd, 0));
    PSB_CUFFT(cufftMakePlanMany64(c->plan_fwd, 3, n, rembed, 1, rdist, cembed, 1, cdist,
        precision == 8 ? CUFFT_D2Z : CUFFT_R2C, 1, &ws_f));
    PSB_CUFFT(cufftSetStream(c->plan_fwd, c->st));
    c->have_fwd = true;
  }
  else if (c->have_fwd) PSB_CUFFT(cufftGetSize(c->plan_fwd, &ws_f));
  if (need_inv && !c->have_inv) {
    PSB_CUFFT(cufftCreate(&c->plan_inv));
    PSB_CUFFT(cufftSetAutoAllocation(c->plan_inv, 0));
    PSB_CUFFT(cufftMakePlanMany64(c->plan_inv, 3, n, cembed, 1, cdist, rembed, 1, rdist,
        precision == 8 ? CUFFT_Z2D : CUFFT_C2R, 1, &ws_i));
    PSB_CUFFT(cufftSetStream(c->plan_inv, c->st));
    c->have_inv = true;
  }
  else if (c->have_inv) PSB_CUFFT(cufftGetSize(c->plan_inv, &ws_i));
  // the hand-written strided passes need only a batched 1-D r2c along z from cuFFT
  if (own && !(c->opt_fft_fused && zp == ng) && !fft_own_z(c, ng, precision) && !c->have_z) {
    long long n1[1] = {ng}, re1[1] = {2LL * ngk}, ce1[1] = {ngk};
    PSB_CUFFT(cufftCreate(&c->plan_z));
    PSB_CUFFT(cufftSetAutoAllocation(c->plan_z, 0));
    PSB_CUFFT(cufftMakePlanMany64(c->plan_z, 1, n1, re1, 1, 2LL * ngk, ce1, 1, ngk,
        precision == 8 ? CUFFT_D2Z : CUFFT_R2C, (long long) zp * ng, &ws_z));
    PSB_CUFFT(cufftSetStream(c->plan_z, c->st));
    c->have_z = true;
  }
  else if (c->have_z) PSB_CUFFT(cufftGetSize(c->plan_z, &ws_z));
  // cuFFT's strided batched 1-D c2c for the x pass where it is the faster one
  size_t ws_x = 0;
  if (own && !fft_own_x(c, ng, precision) && !c->have_x) {
    long long n1[1] = {ng}, embed[1] = {ng};
    const long long lines = (long long) ng * ngk;
    PSB_CUFFT(cufftCreate(&c->plan_x));
    PSB_CUFFT(cufftSetAutoAllocation(c->plan_x, 0));
    PSB_CUFFT(cufftMakePlanMany64(c->plan_x, 1, n1, embed, lines, 1, embed, lines, 1,
        precision == 8 ? CUFFT_Z2Z : CUFFT_C2C, lines, &ws_x));
    PSB_CUFFT(cufftSetStream(c->plan_x, c->st));
    c->have_x = true;
  }
  else if (c->have_x) PSB_CUFFT(cufftGetSize(c->plan_x, &ws_x));
  const size_t ws = std::max(std::max(std::max(ws_f, ws_i), ws_z), ws_x);
  if (c->fftwork.reserve(ws ? ws : 256)) return -1;
  if (c->have_z) PSB_CUFFT(cufftSetWorkArea(c->plan_z, c->fftwork.p));
  if (c->have_x) PSB_CUFFT(cufftSetWorkArea(c->plan_x, c->fftwork.p));
  if (c->have_z && zp < ng && c->opt_fft_streams > 1 && !c->have_z2) {
    // a twin of the z plan on the side stream (own work area)
    long long n1[1] = {ng}, re1[1] = {2LL * ngk}, ce1[1] = {ngk};
    size_t ws2 = 0;
    PSB_CUFFT(cufftCreate(&c->plan_z2));
    PSB_CUFFT(cufftMakePlanMany64(c->plan_z2, 1, n1, re1, 1, 2LL * ngk, ce1, 1, ngk,
        precision == 8 ? CUFFT_D2Z : CUFFT_R2C, (long long) zp * ng, &ws2));
    PSB_CUFFT(cufftSetStream(c->plan_z2, c->st_aux));
    c->have_z2 = true;
  }
  if (c->have_fwd) PSB_CUFFT(cufftSetWorkArea(c->plan_fwd, c->fftwork.p));
  if (c->have_inv) PSB_CUFFT(cufftSetWorkArea(c->plan_inv, c->fftwork.p));
  return 0;
}

// skip_ok: only the binning reads the result, so the x pass may leave the
// columns beyond the last bin edge untransformed.
int fft_forward(psb_context *c, void *mesh, bool skip_ok) {
  StageScope sc(c, PSB_T_FFT, c->st);
  c->launches++;
  if (c->own_fft) {
    // z: hand-written r2c over contiguous rows (or cuFFT batched 1-D r2c); y, x:
    // hand-written strided passes (fft_strided.cu).  With an L2 budget the z and y
    // passes run group by group over a few planes (ablation: separate launches
    // per group cost more than the L2 hits save).
    const int ng = c->plan_ng, ngk = ng / 2 + 1, prec = c->plan_prec, zp = c->plan_zp;
    const size_t plane = (size_t) ng * ngk * 2 * prec;
    // z pass: cuFFT's power-of-two double kernel is a little faster than ours
    // (3.5 vs 4.1 ms at 1024^3); ours wins in single precision and for 1536 = 2^9 3
    // (cuFFT: 38 ms per 1536^3 pass)
    const bool own_z = fft_own_z(c, ng, prec);
    // cells beyond the last bin edge are never read when only the binning follows: the x
    // pass skips their tiles, and neither pass stores them (FftStoreSkip)
    const bool skip = skip_ok && c->opt_fft_skip && c->bins_ready;
    const bool sskip = skip && c->opt_fft_store_skip && fft_own_x(c, ng, prec);
    FftStoreSkip ss_y, ss_x;
    if (sskip) {
      ss_y.k2t = c->bg.kax2[1]; ss_y.k2k = c->bg.kax2[2]; ss_y.k2max = c->fft_k2max;
      ss_x.k2t = c->bg.kax2[0]; ss_x.k2o = c->bg.kax2[1]; ss_x.k2k = c->bg.kax2[2];
      ss_x.k2max = c->fft_k2max; ss_x.per_column = 1;
    }
    if (c->opt_fft_fused && zp == ng) {
      // z + y in one persistent kernel, handed over plane by plane through the L2
      if (c->fftdone.reserve(sizeof(int) * (size_t) ng)) return -1;
      if (launch_fft_zy(mesh, prec, ng, ngk, ng, c->fftdone.as<int>(), c->st)) return -1;
      c->launches += 2;
    }
    else {
      const bool two = c->opt_fft_streams > 1 && zp < ng && c->have_z2 && !own_z;
      if (two) {
        PSB_CUDA(cudaEventRecord(c->ev_aux_go, c->st));
        PSB_CUDA(cudaStreamWaitEvent(c->st_aux, c->ev_aux_go, 0));
      }
      int gi = 0;
      for (int x0 = 0; x0 < ng; x0 += zp, gi++) {
        char *grp = static_cast<char *>(mesh) + (size_t) x0 * plane;
        // with two streams the z pass of one group overlaps the y pass of the other
        const bool alt = two && (gi & 1);
        cudaStream_t sg = alt ? c->st_aux : c->st;
        if (own_z) {
          if (launch_fft_rows(grp, grp, prec, ng, (long) zp * ng, 2 * (size_t) ngk, ngk, sg)) return -1;
        }
        else {
          cufftHandle pz = alt ? c->plan_z2 : c->plan_z;
          if (prec == 8) PSB_CUFFT(cufftExecD2Z(pz, (cufftDoubleReal *) grp, (cufftDoubleComplex *) grp));
          else PSB_CUFFT(cufftExecR2C(pz, (cufftReal *) grp, (cufftComplex *) grp));
        }
        if (launch_fft_strided(grp, prec, ng, ngk, 1, zp, nullptr, nullptr, 0.0, sg, sskip ? &ss_y : nullptr)) return -1;
        c->launches += 2;
      }
      if (two) {
        PSB_CUDA(cudaEventRecord(c->ev_aux_done, c->st_aux));
        PSB_CUDA(cudaStreamWaitEvent(c->st, c->ev_aux_done, 0));
      }
    }
    StageScope own(c, PSB_T_FFT_STRIDED, c->st);     // the x pass
    c->launches++;
    if (!fft_own_x(c, ng, prec)) {
      if (prec == 8)
        PSB_CUFFT(cufftExecZ2Z(c->plan_x, (cufftDoubleComplex *) mesh, (cufftDoubleComplex *) mesh, CUFFT_FORWARD));
      else
        PSB_CUFFT(cufftExecC2C(c->plan_x, (cufftComplex *) mesh, (cufftComplex *) mesh, CUFFT_FORWARD));
      return 0;
    }
    if (launch_fft_strided(mesh, prec, ng, ngk, 0, ng, skip ? c->bg.kax2[1] : nullptr,
          skip ? c->bg.kax2[2] : nullptr, c->fft_k2max, c->st, sskip ? &ss_x : nullptr))
      return -1;
    return 0;
  }
  if (c->plan_prec == 8)
    PSB_CUFFT(cufftExecD2Z(c->plan_fwd, (cufftDoubleReal *) mesh, (cufftDoubleComplex *) mesh));
  else
    PSB_CUFFT(cufftExecR2C(c->plan_fwd, (cufftReal *) mesh, (cufftComplex *) mesh));
  return 0;
}

int fft_inverse(psb_context *c, void *mesh) {
  StageScope sc(c, PSB_T_FFT, c->st);
  c->launches++;
  if (c->plan_prec == 8)
    PSB_CUFFT(cufftExecZ2D(c->plan_inv, (cufftDoubleComplex *) mesh, (cufftDoubleReal *) mesh));
  else
    PSB_CUFFT(cufftExecC2R(c->plan_inv, (cufftComplex *) mesh, (cufftReal *) mesh));
  return 0;
}


// count_mode's normalisation and shot-noise subtraction (src/multipole.c:1047-1161)
// and the log-bin rescaling of powspec() (:1267-1274, quirk Q7); O(nl * nbin) on the host
void normalise(psb_result *res, const psb_params *par, bool issim, int nc, const double *shot,
    const double *norm) {
  const int nl = res->nl, nbin = res->nbin;
  if (issim) {
    for (int i = 0; i < nc; i++) {
      if (!res->has_pl[i]) continue;
      for (int l = 0; l < nl; l++)
        for (int b = 0; b < nbin; b++) {
          double &p = res->pl[i][(size_t) l * nbin + b];
          if (res->cnt[b]) {
            p /= norm[i] * res->cnt[b];
            p -= shot[i] * res->lcnt[(size_t) l * nbin + b] / res->cnt[b];
          }
          p *= 2 * par->poles[l] + 1;
        }
    }
    if (res->has_xpl)
      for (int l = 0; l < nl; l++)
        for (int b = 0; b < nbin; b++) {
          double &p = res->xpl[(size_t) l * nbin + b];
          if (res->cnt[b]) p /= sqrt(norm[0] * norm[1]) * res->cnt[b];
          p *= 2 * par->poles[l] + 1;
        }
  }
  else {
    for (int i = 0; i < nc; i++) {
      if (!res->has_pl[i]) continue;
      if (par->poles[0] == 0)
        for (int b = 0; b < nbin; b++)
          if (res->cnt[b]) {
            double &p = res->pl[i][b];
            p = p / res->cnt[b] - shot[i];
            p /= norm[i];
          }
      for (int n = 1; n < nl; n++)
        for (int b = 0; b < nbin; b++)
          if (res->cnt[b]) res->pl[i][(size_t) n * nbin + b] *= 4 * PI / (norm[i] * res->cnt[b]);
    }
    if (res->has_xpl) {
      if (par->poles[0] == 0)
        for (int b = 0; b < nbin; b++)
          if (res->cnt[b]) res->xpl[b] /= sqrt(norm[0] * norm[1]) * res->cnt[b];
      for (int n = 1; n < nl; n++)
        for (int b = 0; b < nbin; b++)
          if (res->cnt[b])
            res->xpl[(size_t) n * nbin + b] *= 2 * PI / (sqrt(norm[0] * norm[1]) * res->cnt[b]);
    }
  }
  if (par->logscale) {          // quirk Q7, src/multipole.c:1267-1274
    for (int i = 0; i < nbin; i++) {
      res->k[i] = pow(10, res->k[i]);
      res->km[i] = pow(10, res->k[i]);
      res->kedge[i] = pow(10, res->k[i]);
    }
    res->kedge[nbin] = pow(10, res->kedge[nbin]);
  }
}

// powspec_init (src/multipole.c:335-394) + the per-axis tables + powspec_precomp
// (src/multipole.c:111-257).  Everything here is independent of the particles,
// so psb_mesh launches it on the side stream before the scatter: the
// compute-bound mode counting then overlaps the bandwidth-bound upload, sort and
// memset instead of competing with the FFTs.
bool same_bins(const psb_context *c, const psb_params *p) {
  if (!c->bins_ready) return false;
  const psb_params &q = c->bins_par;
  if (p->gsize != q.gsize || p->logscale != q.logscale || p->npole != q.npole ||
      p->issim != q.issim || p->intlace != q.intlace || p->assign != q.assign ||
      p->kmin != q.kmin || p->kmax != q.kmax || p->kbin != q.kbin)
    return false;
  for (int i = 0; i < p->npole; i++) if (p->poles[i] != q.poles[i]) return false;
  for (int a = 0; a < 3; a++)
    if ((p->issim && p->los[a] != q.los[a]) || c->bins_box[a] != c->bsize[a]) return false;
  return true;
}

int prepare_bins(psb_context *c, const psb_params *par) {
  c->bins_ready = false;
  const int ng = par->gsize, ngk = ng / 2 + 1, nl = par->npole;
  const bool issim = par->issim, il = par->intlace;
  double bmax = std::max(c->bsize[0], std::max(c->bsize[1], c->bsize[2]));
  double kny = PI * ng / bmax;
  if (par->logscale) kny = log10(kny);
  double kmax = kny;
  if (par->kmax > 0 && kmax > par->kmax) kmax = par->kmax;
  const double nbf = round((kmax - par->kmin) / par->kbin);
  if (nbf >= INT_MAX) {
    set_error("too many wave number bins due to the small bin size: %.10lg\n", par->kbin);
    return -1;
  }
  int nbin = (int) nbf;
  if (par->kmin + par->kbin * nbin > kny) nbin -= 1;
  if (nbin < 1) {
    set_error("not enough k bins given the Nyquist frequency and the bin size: %.10lg\n", par->kbin);
    return -1;
  }
  c->nbin = nbin;
  c->kedge.resize(nbin + 1);
  for (int i = 0; i <= nbin; i++) c->kedge[i] = par->kmin + par->kbin * i;

  // per-axis tables: k, k^2, window, interlace phase (3 x 5 x ng) + k^2 edges
  const size_t tlen = (size_t) ng;
  std::vector<double> &T = c->host_tables;
  T.assign(15 * tlen + nbin + 1 + 6 * tlen, 0.0);
  const double fac = PI / ng;
  for (int a = 0; a < 3; a++) {
    const double vec = 2 * PI / c->bsize[a];    // src/multipole.c:113
    double *kax = &T[(0 + a) * tlen], *kax2 = &T[(3 + a) * tlen], *wax = &T[(6 + a) * tlen];
    double *pc = &T[(9 + a) * tlen], *ps = &T[(12 + a) * tlen];
    for (int i = 0; i < ng; i++) {
      const double n = (i <= (ng >> 1)) ? i : i - ng;
      wax[i] = window_axis(il, par->assign, n * fac);
      const double k = n * vec;
      kax[i] = k;
      kax2[i] = k * k;
      const double ph = fac * n;                // src/multipole.c:468-476
      pc[i] = cos(ph);
      ps[i] = sin(ph);
    }
  }
  {
    // smallest k^2 that the reference's arithmetic puts in bin >= b (or past the
    // last edge): a monotone predicate, bisected over the doubles
    const bool lg = par->logscale;
    const double k0 = c->kedge[0], k1 = c->kedge[nbin], dk = par->kbin;
    double *e = &T[15 * tlen];
    auto coord = [lg](double x) { return lg ? 0.5 * log10(x) : sqrt(x); };
    for (int b = 0; b < nbin; b++)
      e[b] = bisect_first([&](double x) {
        if (coord(x) >= k1) return true;
        return ref_bin(x, lg, k0, k1, dk, nbin) >= b;
      });
    e[nbin] = bisect_first([&](double x) { return coord(x) >= k1; });
  }
  {
    // z-axis tables packed per k for the binning kernel: one address, three 16-byte loads
    double *zt = &T[15 * tlen + nbin + 1];
    for (int k = 0; k < ng; k++) {
      zt[6 * k + 0] = T[(3 + 2) * tlen + k];     // k_z^2
      zt[6 * k + 1] = T[(6 + 2) * tlen + k];     // window
      zt[6 * k + 2] = T[(9 + 2) * tlen + k];     // cos(pi k / Ng)
      zt[6 * k + 3] = T[(12 + 2) * tlen + k];    // sin
      zt[6 * k + 4] = T[(0 + 2) * tlen + k];     // k_z
    }
  }
  if (c->tables.reserve(T.size() * sizeof(double) + 16)) return -1;
  BinGeom &bg = c->bg;
  memset(&bg, 0, sizeof bg);
  bg.ng = ng; bg.ngk = ngk; bg.nbin = nbin; bg.nl = nl;
  for (int i = 0; i < nl; i++) bg.poles[i] = par->poles[i];
  bg.issim = issim; bg.logk = par->logscale; bg.intlace = il;
  bg.j0 = 0; bg.nj = ng;
  bg.symx = c->opt_geom_sym && (!issim || par->los[0] == 0.0);
  bg.symy = c->opt_geom_sym && (!issim || par->los[1] == 0.0);
  bg.symxy = bg.symx && bg.symy && c->bsize[0] == c->bsize[1];
  for (int a = 0; a < 3; a++) {
    bg.los[a] = issim ? par->los[a] : 0.0;
    const double *base = c->tables.as<double>();
    bg.kax[a] = base + (0 + a) * tlen; bg.kax2[a] = base + (3 + a) * tlen;
    bg.wax[a] = base + (6 + a) * tlen;
    bg.pc[a] = base + (9 + a) * tlen; bg.ps[a] = base + (12 + a) * tlen;
  }
  bg.k2edge = c->tables.as<double>() + 15 * tlen;
  bg.ztab = c->tables.as<double>() + 15 * tlen + nbin + 1;
  if ((15 * tlen + nbin + 1) & 1) {
    // keep the packed table 16-byte aligned for the double2 loads
    T.insert(T.begin() + 15 * tlen + nbin + 1, 0.0);
    bg.ztab += 1;
  }
  {
    // math/legpoly.h:47-61 as coefficient rows: even: c0 + c1 x^2 + c2 x^4 + c3 x^6,
    // odd: x * (same form)
    static const double LC[7][4] = {
      {1, 0, 0, 0}, {1, 0, 0, 0}, {-0.5, 1.5, 0, 0}, {-1.5, 2.5, 0, 0},
      {0.375, -3.75, 4.375, 0}, {1.875, -8.75, 7.875, 0},
      {-0.3125, 6.5625, -19.6875, 14.4375}};
    bg.anyodd = 0;
    for (int i = 0; i < nl; i++) {
      const int ell = par->poles[i];
      for (int q = 0; q < 4; q++) bg.legc[i][q] = LC[ell][q];
      bg.legodd[i] = ell & 1;
      bg.anyodd |= ell & 1;
    }
  }
  bg.k0 = c->kedge[0]; bg.k1 = c->kedge[nbin]; bg.dk = par->kbin; bg.inv_dk = 1.0 / par->kbin;

  const size_t nacc = (size_t) nl * nbin;
  const size_t sb = bin_scratch_bytes(bg);
  c->bin_sb = sb;
  // device bins: cnt (u64) | km | lcnt[nl*nbin] | pl0 | pl1 | xpl
  const size_t bins_doubles = 2 * (size_t) nbin + 4 * nacc;
  if (c->binscratch.reserve(2 * sb) || c->bins.reserve(bins_doubles * sizeof(double))) return -1;
  PSB_CUDA(cudaMemcpyAsync(c->tables.p, T.data(), T.size() * sizeof(double),
      cudaMemcpyHostToDevice, c->st_geom));
  PSB_CUDA(cudaMemsetAsync(c->bins.p, 0, bins_doubles * sizeof(double), c->st_geom));
  unsigned long long *d_cnt = c->bins.as<unsigned long long>();
  double *d_km = c->bins.as<double>() + nbin;
  double *d_lcnt = d_km + nbin;
  {
    StageScope sc(c, PSB_T_GEOM, c->st_geom);
    if (launch_geometry(bg, d_cnt, d_km, d_lcnt, c->binscratch.as<double>(), sb, c->st_geom))
      return -1;
    c->launches += 3;
  }
  PSB_CUDA(cudaEventRecord(c->ev_geom, c->st_geom));
  c->bins_par = *par;
  for (int a = 0; a < 3; a++) c->bins_box[a] = c->bsize[a];
  c->bins_ready = true;
  return 0;
}

}  // namespace psb_host

using namespace psb_host;

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char *psb_last_error(void) { return get_error(); }

int psb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

psb_context *psb_create(int device) {
  int n = psb_device_count();
  if (n <= 0) {
    set_error("no CUDA device available: the powspec_b200 hot path has no CPU fallback\n");
    return nullptr;
  }
  if (device < 0 || device >= n) { set_error("invalid CUDA device %d (have %d)\n", device, n); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed\n", device); return nullptr; }
  psb_context *c = new psb_context();
  c->device = device;
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->st_geom, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->st_aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_aux_go, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_aux_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_memset[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_memset[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_geom, cudaEventDisableTiming) != cudaSuccess) {
    set_error("failed to create CUDA streams\n");
    delete c;
    return nullptr;
  }
  for (double &m : c->ms) m = 0;
  return c;
}

void psb_destroy(psb_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->have_fwd) cufftDestroy(c->plan_fwd);
  if (c->have_inv) cufftDestroy(c->plan_inv);
  if (c->have_z) cufftDestroy(c->plan_z);
  if (c->have_z2) cufftDestroy(c->plan_z2);
  if (c->have_x) cufftDestroy(c->plan_x);
  if (c->slab_have) { if (c->slab_has_yz) cufftDestroy(c->slab_yz); if (c->slab_has_x) cufftDestroy(c->slab_x); }
  for (int i = 0; i < 2; i++) {
    for (int j = 0; j < 2; j++) { c->part_in[i][j].release(); c->mesh[i][j].release(); }
    c->fkl[i].release(); c->fk0copy[i].release();
    c->chunkbuf[i].release();
    if (c->ev_filled[i]) cudaEventDestroy(c->ev_filled[i]);
    if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
  }
  if (c->st_copy) cudaStreamDestroy(c->st_copy);
  if (c->pinned_base) cudaFreeHost(c->pinned_base);
  for (int i = 0; i < PSB_STAGE_SLOTS; i++) if (c->pinned_free[i]) cudaEventDestroy(c->pinned_free[i]);
  copy_pool_destroy(c->copy_pool);
  c->copy_pool = nullptr;
  c->fka.release(); c->sorted.release(); c->keys.release(); c->hist.release();
  c->tile_cnt.release(); c->tile_start.release(); c->wmax_buf.release();
  c->tile_ovrec.release(); c->tile_ovtile.release();
  c->cursor.release(); c->cubtmp.release(); c->bounds_part.release(); c->fftwork.release(); c->fftdone.release(); c->cnvt_tab.release();
  c->tables.release(); c->binscratch.release(); c->bins.release();
  reset_timings(c);
  for (auto e : c->evpool) cudaEventDestroy(e);
  if (c->ev_geom) cudaEventDestroy(c->ev_geom);
  if (c->st) cudaStreamDestroy(c->st);
  if (c->st_geom) cudaStreamDestroy(c->st_geom);
  if (c->ev_aux_go) cudaEventDestroy(c->ev_aux_go);
  if (c->ev_aux_done) cudaEventDestroy(c->ev_aux_done);
  for (auto e : c->ev_memset) if (e) cudaEventDestroy(e);
  if (c->st_aux) cudaStreamDestroy(c->st_aux);
  delete c;
}

int psb_set_option(psb_context *c, const char *name, long value) {
  if (!c || !name) return -1;
  if (!strcmp(name, "sort")) { c->opt_sort = value; return 0; }
  if (!strcmp(name, "sort_min")) { c->opt_sort_min = value; return 0; }
  if (!strcmp(name, "strip")) { c->opt_strip = value; return 0; }
  if (!strcmp(name, "coop")) { c->opt_coop = value; return 0; }
  if (!strcmp(name, "owner")) { c->opt_owner = value; return 0; }
  if (!strcmp(name, "coop_variant")) { c->opt_coop_variant = value; return 0; }
  if (!strcmp(name, "xgroup")) { c->opt_xgroup = value; return 0; }
  if (!strcmp(name, "own_fft")) { c->opt_own_fft = value; return 0; }
  if (!strcmp(name, "fft_skip")) { c->opt_fft_skip = value; return 0; }
  if (!strcmp(name, "fft_l2_mb")) { c->opt_fft_l2_mb = value; return 0; }
  if (!strcmp(name, "fft_streams")) { c->opt_fft_streams = value; return 0; }
  if (!strcmp(name, "fft_own_z")) { c->opt_fft_own_z = value; return 0; }
  if (!strcmp(name, "fft_own_x")) { c->opt_fft_own_x = value; return 0; }
  if (!strcmp(name, "fft_fused")) { c->opt_fft_fused = value; return 0; }
  if (!strcmp(name, "fft_variant")) { fft_set_variant((int) value); return 0; }
  if (!strcmp(name, "fft_store_skip")) { c->opt_fft_store_skip = value; return 0; }
  if (!strcmp(name, "tile_fill_unroll")) { tile_set_fill_unroll((int) value); return 0; }
  if (!strcmp(name, "bin_threads")) { bin_set_threads((int) value); return 0; }
  if (!strcmp(name, "geom_blocks")) { bin_set_geom_blocks((int) value); return 0; }
  if (!strcmp(name, "tile_onepass")) { c->opt_tile_onepass = value; c->tile_onepass_backoff = 0; return 0; }
  if (!strcmp(name, "tile_cap")) { c->opt_tile_cap = value; return 0; }
  if (!strcmp(name, "tile_index")) { c->opt_tile_index = value; return 0; }
  if (!strcmp(name, "tile_ovcap")) { c->opt_tile_ovcap = value; return 0; }
  if (!strcmp(name, "tile_tma")) { tile_set_tma((int) value); return 0; }
  if (!strcmp(name, "l2_fetch")) {          // L2 fetch granularity hint in bytes (32 / 64 / 128)
    if (cudaSetDevice(c->device) != cudaSuccess ||
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) value) != cudaSuccess) {
      set_error("cudaLimitMaxL2FetchGranularity %ld rejected\n", value);
      return -1;
    }
    return 0;
  }
  if (!strcmp(name, "memset_overlap")) { c->opt_memset_overlap = value; return 0; }
  if (!strcmp(name, "survey_direct")) { c->opt_survey_direct = value; return 0; }
  if (!strcmp(name, "geom_sym")) { c->opt_geom_sym = value; return 0; }
  if (!strcmp(name, "stream")) { c->opt_stream = value; return 0; }
  if (!strcmp(name, "h2d_threads")) { c->opt_h2d_threads = value; return 0; }
  if (!strcmp(name, "h2d_wc")) { c->opt_h2d_wc = value; return 0; }
  if (!strcmp(name, "h2d_piece_mb")) { c->opt_h2d_piece_mb = value; return 0; }
  if (!strcmp(name, "h2d_slots")) { c->opt_h2d_slots = value; return 0; }
  if (!strcmp(name, "h2d_nt")) { copy_set_stream_stores((int) value); return 0; }
  if (!strcmp(name, "stream_chunk")) { c->opt_stream_chunk = value; return 0; }
  if (!strcmp(name, "stream_taper")) { c->opt_stream_taper = value; return 0; }
  set_error("unknown option: %s\n", name);
  return -1;
}

// genr_mesh, src/genr_mesh.c:874-926
int psb_mesh(psb_context *c, const psb_params *par, const psb_cats *cats) {
  if (!c) { set_error("no device context\n"); return -1; }
  if (check_params(par)) return -1;
  if (!cats) { set_error("catalogs not read\n"); return -1; }
  PSB_CUDA(cudaSetDevice(c->device));
  reset_timings(c);
  c->launches = 0;
  c->mesh_ready = false;
  c->bins_ready = false;
  c->par = *par;
  trace_mark("psb_mesh: enter");
  const int nc = par->ncat, ng = par->gsize, prec = par->precision;
  const int ngk = ng / 2 + 1, rowlen = 2 * ngk;
  const size_t mesh_bytes = (size_t) ng * ng * rowlen * prec;

  // Simulation boxes know their box before seeing a particle (min = 0, size =
  // BOX_SIZE), so a host catalogue can be streamed: upload, bounds, sort and
  // scatter chunk by chunk with PCIe and SMs overlapped; the bounds check of
  // def_box is evaluated once the stream has drained.  Surveys need the bounds
  // of every catalogue to define the box, so they are made resident first.
  const bool convert = cats->cnvt != nullptr;
  const bool streaming = par->issim && cats->memspace == PSB_MEM_HOST && c->opt_stream && !convert;
  // simulation boxes: box known in advance, bound checks deferred (computed by
  // the sort's key pass while it reads the catalogue anyway)
  const bool deferred = par->issim && (streaming || cats->memspace == PSB_MEM_DEVICE) && !convert;
  const double *dptr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  size_t cnt[2][2] = {{0, 0}, {0, 0}};
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  size_t nchunk_total = 0;
  for (int i = 0; i < nc; i++)
    for (int s = 0; s < (par->issim ? 1 : 2); s++) {
      const double *src = s ? cats->rand[i] : cats->data[i];
      const size_t n = s ? cats->nrand[i] : cats->ndata[i];
      cnt[i][s] = n;
      if (n && !src) { set_error("catalogs not read\n"); return -1; }
      nchunk_total += (streaming ? stream_schedule(c, n).size() : (n + DEV_CHUNK - 1) / DEV_CHUNK) + 1;
      if (streaming) continue;
      if (cats->memspace == PSB_MEM_DEVICE && !convert) dptr[i][s] = src;
      else if (cats->memspace == PSB_MEM_DEVICE) {
        // converted in place below: work on a copy, the caller's array stays as it is
        if (c->part_in[i][s].reserve(n ? n * 32 : 32)) return -1;
        PSB_CUDA(cudaMemcpyAsync(c->part_in[i][s].p, src, n * 32, cudaMemcpyDeviceToDevice, c->st));
        dptr[i][s] = c->part_in[i][s].as<double>();
      }
      else {
        if (upload(c, src, n, c->part_in[i][s])) return -1;
        dptr[i][s] = c->part_in[i][s].as<double>();
      }
      if (!deferred && !convert && coordinate_bounds(c, dptr[i][s], n, lo, hi)) return -1;
    }
  if (convert) {
    // cnvt_coord(), src/powspec.c:39, on the uploaded records
    double *arr[4];
    size_t an[4];
    int narr = 0;
    for (int i = 0; i < nc; i++)
      for (int s = 0; s < (par->issim ? 1 : 2); s++)
        if ((s ? cats->rcnvt[i] : cats->dcnvt[i]) && cnt[i][s]) {
          arr[narr] = const_cast<double *>(dptr[i][s]);
          an[narr++] = cnt[i][s];
        }
    int order = 0;
    if (narr && convert_arrays(c, cats->cnvt, arr, an, narr, &order)) return -1;
    if (par->verbose && order)
      printf("  Legendre-Gauss order %d chosen for the integration error %g\n", order, cats->cnvt->ecdst);
    for (int i = 0; i < nc; i++)
      for (int s = 0; s < (par->issim ? 1 : 2); s++)
        if (coordinate_bounds(c, dptr[i][s], cnt[i][s], lo, hi)) return -1;
  }
  if (deferred) {
    for (int a = 0; a < 3; a++) { c->bmin[a] = 0; c->bsize[a] = par->bsize[a]; }
    if (bounds_begin(c, nchunk_total)) return -1;
  }
  else {
    for (int a = 0; a < 3; a++) c->bmax[a] = hi[a];
    if (define_box(par, lo, hi, c->bmin, c->bsize)) return -1;
  }
  // the box is known: start the particle-independent part of powspec() now.  A
  // failure here (e.g. no k bin below the Nyquist frequency) belongs to powspec()
  // in the reference, so it is left for psb_power to report.
  {
    QuietErrors q;
    prepare_bins(c, par);
  }
  trace_mark("psb_mesh: bins prepared (host part), side stream launched");

  AssignGeom g;
  memset(&g, 0, sizeof g);
  g.ng = ng; g.rowlen = rowlen;
  g.strip = (int) std::min<long>(std::max<long>(c->opt_strip, 1), ng);
  g.xgroup = (int) c->opt_xgroup;
  g.coop = (int) c->opt_coop;
  g.coop_variant = (int) c->opt_coop_variant;
  g.x0 = 0; g.nx = ng; g.xbase = 0; g.nxloc = ng;
  for (int a = 0; a < 3; a++) {
    g.org[a] = c->bmin[a];
    g.len[a] = c->bsize[a];
    g.inv_len[a] = 1.0 / c->bsize[a];
    g.sorg[a] = c->bmin[a] - 0.5 * c->bsize[a] / ng;    // src/genr_mesh.c:817-818
  }

  // gen_dens, src/genr_mesh.c:793-858.  Randoms are scattered with weight
  // -alpha * w into the data mesh (the reference builds a second mesh and
  // subtracts, :802-806): one pass, no extra field.
  const int nf = par->intlace ? 2 : 1;
  for (int i = 0; i < nc; i++)
    for (int f = 0; f < nf; f++)
      if (c->mesh[i][f].reserve(mesh_bytes)) return -1;
  if (c->opt_memset_overlap) {
    // all memsets on a side stream: bandwidth-bound, they run under the
    // latency-bound particle sort; the first scatter into a catalogue's meshes
    // waits for them (sort_assign_chunk)
    PSB_CUDA(cudaEventRecord(c->ev_aux_go, c->st));
    PSB_CUDA(cudaStreamWaitEvent(c->st_aux, c->ev_aux_go, 0));
    for (int i = 0; i < nc; i++) {
      for (int f = 0; f < nf; f++) {
        StageScope sc(c, PSB_T_MEMSET, c->st_aux);
        PSB_CUDA(cudaMemsetAsync(c->mesh[i][f].p, 0, mesh_bytes, c->st_aux));
      }
      PSB_CUDA(cudaEventRecord(c->ev_memset[i], c->st_aux));
    }
  }
  // with the owner-computes assignment the first dense chunk STORES whole tiles, so the
  // meshes are not zeroed up front: whoever touches a mesh first takes care of it
  const bool lazy_zero = !c->opt_memset_overlap && c->opt_owner != 0 && tile_assign_supported(g);
  for (int i = 0; i < nc; i++) {
    bool fresh = lazy_zero;
    void *m0 = c->mesh[i][0].p, *m1 = par->intlace ? c->mesh[i][1].p : nullptr;
    if (c->opt_memset_overlap) c->memset_pending = c->ev_memset[i];
    else if (!lazy_zero)
      for (int f = 0; f < nf; f++) {
        StageScope sc(c, PSB_T_MEMSET, c->st);
        PSB_CUDA(cudaMemsetAsync(c->mesh[i][f].p, 0, mesh_bytes, c->st));
      }
    if (streaming) {
      if (stream_catalog(c, cats->data[i], cnt[i][0], g, par->assign, prec, 1.0, m0, m1, &fresh)) return -1;
    }
    else {
      if (assign_catalog(c, dptr[i][0], cnt[i][0], g, par->assign, prec, 1.0, m0, m1, deferred, &fresh)) return -1;
      if (!par->issim &&
          assign_catalog(c, dptr[i][1], cnt[i][1], g, par->assign, prec, -cats->alpha[i], m0, m1, false, &fresh))
        return -1;
    }
    if (zero_if_fresh(c, g, prec, m0, m1, &fresh)) return -1;      // empty catalogue
    // nothing was scattered (empty catalogue): later consumers still need the zeros
    if (c->memset_pending) { PSB_CUDA(cudaStreamWaitEvent(c->st, c->memset_pending, 0)); c->memset_pending = nullptr; }
    if (par->issim) {           // src/genr_mesh.c:904-909
      const double vol = c->bsize[0] * c->bsize[1] * c->bsize[2];
      c->shot[i] = vol / cats->wdata[i];
      c->norm[i] = cats->wdata[i] * cats->wdata[i] / vol;
    }
    else { c->shot[i] = cats->shot[i]; c->norm[i] = cats->norm[i]; }
    if (par->verbose) {
      static const char *names[] = {"NGP", "CIC", "TSC", "PCS"};
      if (nc == 2) printf("  Density field generated with %s for catalog %d\n", names[par->assign], i);
      else printf("  Density field generated with %s for the catalog\n", names[par->assign]);
    }
  }
  trace_mark("psb_mesh: all uploads and scatters enqueued");
  // The FFT plans belong to mesh generation in the reference too (mesh_init,
  // src/genr_mesh.c:738-743).  Creating them HERE, while the device is still busy with the
  // uploads and scatters queued above, takes the first call's ~60 ms of cuFFT
  // initialisation off the critical path of a one-shot caller (the reference's C host
  // calls genr_mesh / powspec once per process).
  if (ensure_plans(c, par->gsize, prec, false)) return -1;
  trace_mark("psb_mesh: FFT plans ready");
  if (deferred && bounds_finish(c, par)) return -1;
  trace_mark("psb_mesh: device drained, bounds checked");
  c->mesh_ready = true;
  return 0;
}

int psb_mesh_box(const psb_context *c, double bmin[3], double bsize[3], double bmax[3]) {
  if (!c || !c->mesh_ready) { set_error("meshes not generated\n"); return -1; }
  for (int a = 0; a < 3; a++) { bmin[a] = c->bmin[a]; bsize[a] = c->bsize[a]; bmax[a] = c->bmax[a]; }
  return 0;
}

int psb_copy_mesh(psb_context *c, int cat, int field, void *dst) {
  if (!c || !c->mesh_ready || cat < 0 || cat >= c->par.ncat || field < 0 || field > 1 ||
      (field == 1 && !c->par.intlace)) {
    set_error("no such mesh\n");
    return -1;
  }
  PSB_CUDA(cudaSetDevice(c->device));
  const int ng = c->par.gsize, prec = c->par.precision;
  const size_t bytes = (size_t) ng * ng * ng * prec;
  DevBuf tmp;
  if (tmp.reserve(bytes)) return -1;
  int rc = launch_unpad_copy(c->mesh[cat][field].p, tmp.p, ng, 2 * (ng / 2 + 1), prec, c->st);
  if (!rc && cudaMemcpyAsync(dst, tmp.p, bytes, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) rc = -1;
  if (cudaStreamSynchronize(c->st) != cudaSuccess) rc = -1;
  tmp.release();
  if (rc) set_error("failed to copy the mesh\n");
  return rc;
}

// powspec, src/multipole.c:1179-1278
psb_result *psb_power(psb_context *c, const psb_params *par) {
  if (!c) { set_error("no device context\n"); return nullptr; }
  if (check_params(par)) return nullptr;
  if (!c->mesh_ready) { set_error("meshes not generated\n"); return nullptr; }
  if (cudaSetDevice(c->device) != cudaSuccess) { set_error("cudaSetDevice failed\n"); return nullptr; }
  const int nc = c->par.ncat, ng = c->par.gsize, prec = c->par.precision;
  trace_mark("psb_power: enter");
  const int ngk = ng / 2 + 1;
  const bool issim = c->par.issim, il = c->par.intlace;
  const int nl = par->npole;
  const int lmax = par->poles[nl - 1];
  const bool need_ell = !issim && lmax > 0;
  const size_t mesh_bytes = (size_t) ng * ng * 2 * ngk * prec;
  const size_t ntot = (size_t) ng * ng * ng;

  // ---- k-bins, tables and mode counts: normally already in flight since psb_mesh
  if (!same_bins(c, par) && prepare_bins(c, par)) { c->mesh_ready = false; return nullptr; }
  const int nbin = c->nbin;
  const BinGeom &bg = c->bg;
  const size_t sb = c->bin_sb;
  psb_result *res = new psb_result();
  res->nbin = nbin; res->nl = nl;
  res->kedge = c->kedge; res->k.resize(nbin); res->km.assign(nbin, 0);
  res->cnt.assign(nbin, 0); res->lcnt.assign((size_t) nl * nbin, 0);
  for (int i = 0; i < nbin; i++) res->k[i] = (res->kedge[i] + res->kedge[i + 1]) * 0.5;
  auto fail = [&]() { delete res; c->mesh_ready = false; c->bins_ready = false; return (psb_result *) nullptr; };
  const size_t nacc = (size_t) nl * nbin;
  const size_t bins_doubles = 2 * (size_t) nbin + 4 * nacc;
  auto hard = [&](cudaError_t e) {
    if (e != cudaSuccess) { set_error("CUDA failure: %s\n", cudaGetErrorString(e)); return true; }
    return false;
  };
  double *d_lcnt = c->bins.as<double>() + 2 * nbin;
  double *d_pl[2] = {d_lcnt + nacc, d_lcnt + 2 * nacc};
  double *d_xpl = d_lcnt + 3 * nacc;
  double *scratch_bin = reinterpret_cast<double *>(c->binscratch.as<char>() + sb);
  // the bins are zeroed and the tables uploaded on the side stream
  if (hard(cudaStreamWaitEvent(c->st, c->ev_geom, 0))) return fail();
  // the x pass may skip the columns beyond the last bin edge unless the field is
  // transformed back afterwards (interlaced survey with l > 0)
  const bool skip_ok = !(need_ell && il);
  c->fft_k2max = c->host_tables[15 * (size_t) ng + nbin];

  // ---- dens_k0, src/multipole.c:435-505
  if (ensure_plans(c, ng, prec, need_ell && il)) return fail();
  if (par->verbose) printf("  Alias corrections and wave numbers are pre-computed\n");
  void *Fk0[2] = {nullptr, nullptr}, *Fk1[2] = {nullptr, nullptr}, *Fr[2] = {nullptr, nullptr};
  for (int i = 0; i < nc; i++) {
    void *A = c->mesh[i][0].p, *B = il ? c->mesh[i][1].p : nullptr;
    if (need_ell && !il) {
      // the real-space field is needed again for l > 0: transform a copy
      if (c->fk0copy[i].reserve(mesh_bytes)) return fail();
      if (hard(cudaMemcpyAsync(c->fk0copy[i].p, A, mesh_bytes, cudaMemcpyDeviceToDevice, c->st)))
        return fail();
      if (fft_forward(c, c->fk0copy[i].p, skip_ok)) return fail();
      Fk0[i] = c->fk0copy[i].p; Fr[i] = A;
    }
    else {
      if (fft_forward(c, A, skip_ok)) return fail();
      Fk0[i] = A;
    }
    if (par->verbose) {
      if (nc != 2) printf("  Done with computing 1 FFT for l = 0\n");
      else printf("  Done with computing 1 FFT for l = 0 with catalog %d\n", i + 1);
    }
    if (il) {
      if (fft_forward(c, B, skip_ok)) return fail();
      if (issim) Fk1[i] = B;    // combined on the fly inside the binning kernel
      else {
        StageScope sc(c, PSB_T_BIN, c->st);
        if (launch_combine(bg, prec, A, B, c->st)) return fail();
        c->launches++;
      }
      if (need_ell) {
        // back-transform of the combined field (src/multipole.c:489-500); for
        // sims nothing reads it (quirk Q4) so it is skipped there
        if (hard(cudaMemcpyAsync(B, A, mesh_bytes, cudaMemcpyDeviceToDevice, c->st))) return fail();
        if (fft_inverse(c, B)) return fail();
        if (launch_scale(B, mesh_bytes / prec, 1.0 / (double) ntot, prec, c->st)) return fail();
        c->launches++;
        Fr[i] = B;
      }
      if (par->verbose) printf("  Done with computing 2 FFTs for grid interlacing\n");
    }
  }

  // ---- mode counting
  if (issim) {
    for (int i = 0; i < nc; i++) {
      if (!par->isauto[i]) continue;
      StageScope sc(c, PSB_T_BIN, c->st);
      if (launch_bin(bg, prec, Fk0[i], Fk1[i], Fk0[i], Fk1[i], d_pl[i], scratch_bin, sb, c->st))
        return fail();
      c->launches += 2;
      res->has_pl[i] = true;
    }
    if (par->iscross && nc == 2) {
      StageScope sc(c, PSB_T_BIN, c->st);
      if (launch_bin(bg, prec, Fk0[0], Fk1[0], Fk0[1], Fk1[1], d_xpl, scratch_bin, sb, c->st))
        return fail();
      c->launches += 2;
      res->has_xpl = true;
    }
  }
  else {
    if (par->poles[0] == 0) {
      for (int i = 0; i < nc; i++) {
        if (!par->isauto[i]) continue;
        StageScope sc(c, PSB_T_BIN, c->st);
        if (launch_bin(bg, prec, Fk0[i], nullptr, Fk0[i], nullptr, d_pl[i], scratch_bin, sb, c->st))
          return fail();
        c->launches += 2;
      }
      if (par->iscross && nc == 2) {
        StageScope sc(c, PSB_T_BIN, c->st);
        if (launch_bin(bg, prec, Fk0[0], nullptr, Fk0[1], nullptr, d_xpl, scratch_bin, sb, c->st))
          return fail();
        c->launches += 2;
      }
    }
    for (int i = 0; i < nc; i++) res->has_pl[i] = par->isauto[i];
    res->has_xpl = par->iscross && nc == 2;
    // the reference's loop starts at the second multipole whatever the first is
    // (src/multipole.c:1229-1237)
    for (int n = 1; n < nl; n++) {
      const int ell = par->poles[n];
      if (c->fka.reserve(mesh_bytes)) return fail();
      const bool direct = c->opt_survey_direct != 0;
      for (int i = 0; i < nc; i++) {
        if (!direct) {
          if (c->fkl[i].reserve(mesh_bytes)) return fail();
          if (hard(cudaMemsetAsync(c->fkl[i].p, 0, mesh_bytes, c->st))) return fail();
        }
        YlmGeom yg;
        yg.ng = ng; yg.ngk = ngk; yg.rowlen = 2 * ngk; yg.ell = ell;
        for (int a = 0; a < 3; a++) {
          yg.smin[a] = c->bmin[a] * ng / c->bsize[a];   // src/genr_mesh.c:913-914
          yg.bsize[a] = c->bsize[a];
        }
        for (int m = -ell; m <= ell; m++) {
          yg.m = m;
          {
            StageScope sc(c, PSB_T_YLM, c->st);
            if (launch_ylm_weight_r(yg, prec, Fr[i], c->fka.p, c->st)) return fail();
            c->launches++;
          }
          if (fft_forward(c, c->fka.p, true)) return fail();
          if (!direct) {
            StageScope sc(c, PSB_T_YLM, c->st);
            if (launch_ylm_accum_k(yg, bg, prec, c->fka.p, c->fkl[i].p, c->st)) return fail();
            c->launches++;
            continue;
          }
          // Direct binning of Re(Fk0 conj Fka_m) Y_lm(k_hat): the sum over m is
          // the reference's Re(Fk0 conj Fkl) by linearity, without the Fkl field
          // and its read-modify-write pass per m.
          BinGeom bm = bg;
          bm.ell = ell; bm.m = m; bm.ylm_nrm = ylm_norm(ell, m);
          StageScope sc(c, PSB_T_BIN, c->st);
          if (par->isauto[i]) {
            if (launch_bin(bm, prec, Fk0[i], nullptr, c->fka.p, nullptr, d_pl[i] + (size_t) n * nbin,
                  scratch_bin, sb, c->st))
              return fail();
            c->launches += 2;
          }
          if (par->iscross && nc == 2) {
            // (Fk0[0], Fkl[1]) + (Fk0[1], Fkl[0]), src/multipole.c:1134-1135
            if (launch_bin(bm, prec, Fk0[1 - i], nullptr, c->fka.p, nullptr, d_xpl + (size_t) n * nbin,
                  scratch_bin, sb, c->st))
              return fail();
            c->launches += 2;
          }
        }
        if (par->verbose) {
          if (nc != 2) printf("  Done with computing %d FFTs for l = %d\n", 2 * ell + 1, ell);
          else printf("  Done with computing %d FFTs for l = %d with catalog %d\n", 2 * ell + 1, ell, i + 1);
        }
      }
      if (direct) continue;
      StageScope sc(c, PSB_T_BIN, c->st);
      for (int i = 0; i < nc; i++) {
        if (!par->isauto[i]) continue;
        if (launch_bin(bg, prec, Fk0[i], nullptr, c->fkl[i].p, nullptr, d_pl[i] + (size_t) n * nbin,
              scratch_bin, sb, c->st))
          return fail();
        c->launches += 2;
      }
      if (par->iscross && nc == 2) {
        if (launch_bin(bg, prec, Fk0[0], nullptr, c->fkl[1].p, nullptr, d_xpl + (size_t) n * nbin,
              scratch_bin, sb, c->st) ||
            launch_bin(bg, prec, Fk0[1], nullptr, c->fkl[0].p, nullptr, d_xpl + (size_t) n * nbin,
              scratch_bin, sb, c->st))
          return fail();
        c->launches += 4;
      }
    }
  }

  trace_mark("psb_power: all transforms and binning enqueued");
  // ---- results back: a few thousand doubles
  std::vector<double> hb(bins_doubles);
  if (hard(cudaMemcpyAsync(hb.data(), c->bins.p, bins_doubles * sizeof(double),
          cudaMemcpyDeviceToHost, c->st)) || hard(cudaStreamSynchronize(c->st)))
    return fail();
  trace_mark("psb_power: device drained, bins on the host");
  collect_timings(c);
  memcpy(res->cnt.data(), hb.data(), nbin * sizeof(double));
  memcpy(res->km.data(), hb.data() + nbin, nbin * sizeof(double));
  if (issim) memcpy(res->lcnt.data(), hb.data() + 2 * nbin, nacc * sizeof(double));
  for (int b = 0; b < nbin; b++) if (res->cnt[b]) res->km[b] /= res->cnt[b];
  for (int i = 0; i < 2; i++)
    if (res->has_pl[i]) res->pl[i].assign(hb.begin() + 2 * nbin + (1 + i) * nacc,
        hb.begin() + 2 * nbin + (2 + i) * nacc);
  if (res->has_xpl) res->xpl.assign(hb.begin() + 2 * nbin + 3 * nacc, hb.begin() + 2 * nbin + 4 * nacc);

  const double *shot = c->shot, *norm = c->norm;
  normalise(res, par, issim, nc, shot, norm);
  for (int i = 0; i < 2; i++) { res->shot[i] = i < nc ? shot[i] : 0; res->norm[i] = i < nc ? norm[i] : 0; }
  for (int a = 0; a < 3; a++) { res->bmin[a] = c->bmin[a]; res->bsize[a] = c->bsize[a]; res->bmax[a] = c->bmax[a]; }
  c->mesh_ready = false;        // the FFTs ran in place: the meshes are consumed
  c->bins_ready = false;
  c->ms[PSB_T_TOTAL] = 0;
  for (int s = 0; s < PSB_T_TOTAL; s++) if (s != PSB_T_FFT_STRIDED) c->ms[PSB_T_TOTAL] += c->ms[s];
  trace_mark("psb_power: leave");
  return res;
}

psb_result *psb_run(psb_context *c, const psb_params *par, const psb_cats *cats) {
  if (psb_mesh(c, par, cats)) return nullptr;
  return psb_power(c, par);
}

void psb_result_free(psb_result *r) { delete r; }
int psb_result_nbin(const psb_result *r) { return r ? r->nbin : -1; }
int psb_result_nl(const psb_result *r) { return r ? r->nl : -1; }

long psb_result_get(const psb_result *r, int what, int idx, void *dst) {
  if (!r || !dst) return -1;
  auto put = [&](const void *src, size_t n, size_t sz) { memcpy(dst, src, n * sz); return (long) n; };
  switch (what) {
    case PSB_GET_K: return put(r->k.data(), r->k.size(), 8);
    case PSB_GET_KEDGE: return put(r->kedge.data(), r->kedge.size(), 8);
    case PSB_GET_KM: return put(r->km.data(), r->km.size(), 8);
    case PSB_GET_CNT: return put(r->cnt.data(), r->cnt.size(), 8);
    case PSB_GET_LCNT: return put(r->lcnt.data(), r->lcnt.size(), 8);
    case PSB_GET_PL:
      if (idx < 0 || idx > 1 || !r->has_pl[idx]) return -1;
      return put(r->pl[idx].data(), r->pl[idx].size(), 8);
    case PSB_GET_XPL:
      if (!r->has_xpl) return -1;
      return put(r->xpl.data(), r->xpl.size(), 8);
    case PSB_GET_SHOT: return put(r->shot, 2, 8);
    case PSB_GET_NORM: return put(r->norm, 2, 8);
    case PSB_GET_BMIN: return put(r->bmin, 3, 8);
    case PSB_GET_BSIZE: return put(r->bsize, 3, 8);
    case PSB_GET_BMAX: return put(r->bmax, 3, 8);
    default: return -1;
  }
}

int psb_timings(const psb_context *c, double *ms, int n) {
  if (!c || !ms) return -1;
  for (int i = 0; i < n && i < PSB_T_COUNT; i++) ms[i] = c->ms[i];
  return std::min(n, (int) PSB_T_COUNT);
}

long psb_launch_count(const psb_context *c) { return c ? c->launches : -1; }

// the compute stream as a cudaStream_t (for hosts that want to record their own events on
// it or order their own work after the library's)
void *psb_stream(const psb_context *c) { return c ? (void *) c->st : nullptr; }

// which scatter the last chunk of the last psb_mesh used: 0 = global reductions
// (k_assign_coop), 1 = owner-computes tiles (k_tile_accumulate)
int psb_assign_path(const psb_context *c) { return c ? c->assign_path : -1; }
long psb_tile_overflow(const psb_context *c) { return c ? (long) c->tile_overflowed : -1; }


// ---------------------------------------------------------------------------
// Slab-decomposed mesh: building blocks for one rank.  The
// collectives between them (particle routing, halo planes, FFT transpose,
// allreduce of the bins) are issued by the host over NCCL
// (powspec_b200/distributed.py); buffers are caller-owned device memory.
// Simulation boxes only (BASELINE configs 4 and 5).
// ---------------------------------------------------------------------------
}  // extern "C"
namespace psb_host {
int slab_geom(psb_context *c, const psb_params *par, const psb_slab *sl, AssignGeom &g) {
  if (check_params(par)) return -1;
  if (!par->issim) { set_error("the slab-decomposed path handles simulation boxes only\n"); return -1; }
  if (!sl || sl->nranks < 1 || sl->rank < 0 || sl->rank >= sl->nranks ||
      par->gsize % sl->nranks || (sl->nranks > 1 && par->gsize / sl->nranks < PSB_HALO_HI)) {
    set_error("invalid slab decomposition: GRID_SIZE %d over %d ranks\n", par->gsize,
        sl ? sl->nranks : 0);
    return -1;
  }
  const int ng = par->gsize, nx = ng / sl->nranks;
  memset(&g, 0, sizeof g);
  g.ng = ng; g.rowlen = 2 * (ng / 2 + 1);
  g.strip = (int) std::min<long>(std::max<long>(c->opt_strip, 1), ng);
  g.xgroup = (int) c->opt_xgroup;
  g.coop = (int) c->opt_coop;
  g.coop_variant = (int) c->opt_coop_variant;
  g.x0 = sl->rank * nx; g.nx = nx;
  if (sl->nranks == 1) { g.xbase = 0; g.nxloc = ng; }
  else { g.xbase = (g.x0 - PSB_HALO_LO + ng) % ng; g.nxloc = nx + PSB_HALO_LO + PSB_HALO_HI; }
  for (int a = 0; a < 3; a++) {
    c->bmin[a] = 0; c->bsize[a] = par->bsize[a];
    g.org[a] = 0; g.len[a] = par->bsize[a]; g.inv_len[a] = 1.0 / par->bsize[a];
    g.sorg[a] = -0.5 * par->bsize[a] / ng;
  }
  return 0;
}
}  // namespace psb_host
extern "C" {

size_t psb_slab_mesh_elems(const psb_params *par, const psb_slab *sl) {
  if (!par || !sl || sl->nranks < 1 || par->gsize % sl->nranks) return 0;
  const size_t ng = par->gsize, nx = ng / sl->nranks;
  const size_t planes = sl->nranks == 1 ? ng : nx + PSB_HALO_LO + PSB_HALO_HI;
  return planes * ng * 2 * (ng / 2 + 1);
}

// group the particles by owning slab (owner = slab of the base x-cell on the
// unshifted grid); counts[r] particles for rank r, contiguous in `sorted`
int psb_slab_partition(psb_context *c, const psb_params *par, int nranks, const double *particles,
    size_t n, double *sorted, size_t *counts) {
  if (!c) { set_error("no device context\n"); return -1; }
  psb_slab sl = {nranks, 0};
  AssignGeom g;
  if (slab_geom(c, par, &sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  for (int r = 0; r < nranks; r++) counts[r] = 0;
  if (!n) return 0;
  if (n > 0xffffffffull) { set_error("too many particles in one partition call\n"); return -1; }
  if (c->keys.reserve(n * 4) || c->hist.reserve(64 * 4) || c->cursor.reserve(64 * 4)) return -1;
  PSB_CUDA(cudaMemsetAsync(c->hist.p, 0, 64 * 4, c->st));
  if (launch_owner_keys(particles, n, g, nranks, c->keys.as<uint32_t>(), c->hist.as<uint32_t>(), c->st))
    return -1;
  uint32_t h[64];
  PSB_CUDA(cudaMemcpyAsync(h, c->hist.p, 64 * 4, cudaMemcpyDeviceToHost, c->st));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  uint32_t cur[64], acc = 0;
  for (int r = 0; r < 64; r++) { cur[r] = acc; if (r < nranks) { counts[r] = h[r]; acc += h[r]; } }
  PSB_CUDA(cudaMemcpyAsync(c->cursor.p, cur, 64 * 4, cudaMemcpyHostToDevice, c->st));
  if (launch_owner_scatter(particles, n, c->keys.as<uint32_t>(), c->cursor.as<uint32_t>(), nranks, sorted,
        c->st))
    return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  c->launches += 2;
  return 0;
}

// scatter the rank's particles into its slab buffers (owned planes + halos);
// the buffers must have been zeroed by the caller; mesh1 NULL without interlacing
int psb_slab_assign(psb_context *c, const psb_params *par, const psb_slab *sl,
    const double *particles, size_t n, double wscale, void *mesh0, void *mesh1) {
  if (!c) { set_error("no device context\n"); return -1; }
  AssignGeom g;
  if (slab_geom(c, par, sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  if (par->intlace && !mesh1) { set_error("interlacing needs the second mesh\n"); return -1; }
  if (assign_catalog(c, particles, n, g, par->assign, par->precision, wscale, mesh0,
        par->intlace ? mesh1 : nullptr))
    return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

int psb_add(psb_context *c, void *dst, const void *src, size_t n, int precision) {
  if (!c) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  if (launch_add(dst, src, n, precision, c->st)) return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

}  // extern "C"
namespace psb_host {
int slab_plans(psb_context *c, int ng, int nx, int prec) {
  const bool own = c->opt_own_fft && fft_strided_supported(ng, prec);
  const bool own_x = own && fft_own_x(c, ng, prec), own_z = own && fft_own_z(c, ng, prec);
  if (c->slab_ng == ng && c->slab_nx == nx && c->slab_prec == prec && c->slab_own == own &&
      c->slab_own_x == own_x && c->slab_own_z == own_z)
    return 0;
  if (c->slab_have) {
    if (c->slab_has_yz) cufftDestroy(c->slab_yz);
    if (c->slab_has_x) cufftDestroy(c->slab_x);
    c->slab_have = c->slab_has_yz = c->slab_has_x = false;
  }
  const int ngk = ng / 2 + 1;
  size_t ws = 0;
  c->slab_zp = nx;
  if (own) {
    // y pass hand-written; z pass by cuFFT (1-D batched over groups of planes that fit
    // the L2 budget, if one is set) where that is the faster one
    int zp = 1;
    const int cap = std::min(nx, fft_group_planes(c, ng, prec));
    for (int p = 1; p <= cap; p++) if (nx % p == 0) zp = p;
    c->slab_zp = zp;
    if (!own_z) {
      long long n1[1] = {ng}, re1[1] = {2LL * ngk}, ce1[1] = {ngk};
      PSB_CUFFT(cufftCreate(&c->slab_yz));
      PSB_CUFFT(cufftMakePlanMany64(c->slab_yz, 1, n1, re1, 1, 2LL * ngk, ce1, 1, ngk,
          prec == 8 ? CUFFT_D2Z : CUFFT_R2C, (long long) zp * ng, &ws));
      PSB_CUFFT(cufftSetStream(c->slab_yz, c->st));
      c->slab_has_yz = true;
    }
  }
  else {
    long long n2[2] = {ng, ng}, rembed[2] = {ng, 2LL * ngk}, cembed[2] = {ng, ngk};
    PSB_CUFFT(cufftCreate(&c->slab_yz));
    PSB_CUFFT(cufftMakePlanMany64(c->slab_yz, 2, n2, rembed, 1, (long long) ng * 2 * ngk, cembed, 1,
        (long long) ng * ngk, prec == 8 ? CUFFT_D2Z : CUFFT_R2C, nx, &ws));
    PSB_CUFFT(cufftSetStream(c->slab_yz, c->st));
    c->slab_has_yz = true;
  }
  if (!own_x) {
    // after the transpose a rank holds (Ng_x, ny, Ngk): x has stride ny*Ngk
    long long n1[1] = {ng}, embed[1] = {ng};
    const long long lines = (long long) nx * ngk;       // ny == nx
    PSB_CUFFT(cufftCreate(&c->slab_x));
    PSB_CUFFT(cufftMakePlanMany64(c->slab_x, 1, n1, embed, lines, 1, embed, lines, 1,
        prec == 8 ? CUFFT_Z2Z : CUFFT_C2C, lines, &ws));
    PSB_CUFFT(cufftSetStream(c->slab_x, c->st));
    c->slab_has_x = true;
  }
  c->slab_ng = ng; c->slab_nx = nx; c->slab_prec = prec; c->slab_own = own;
  c->slab_own_x = own_x; c->slab_own_z = own_z; c->slab_have = true;
  return 0;
}
}  // namespace psb_host
extern "C" {

// 2-D r2c over (y, z) of every owned x-plane, in place; `owned` points at the
// first owned plane of the slab buffer
int psb_slab_fft_yz(psb_context *c, const psb_params *par, const psb_slab *sl, void *owned) {
  if (!c) return -1;
  AssignGeom g;
  if (slab_geom(c, par, sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  const int prec = par->precision;
  if (slab_plans(c, g.ng, g.nx, prec)) return -1;
  if (c->slab_own && c->opt_fft_fused && c->slab_zp == g.nx) {
    if (c->fftdone.reserve(sizeof(int) * (size_t) g.nx)) return -1;
    if (launch_fft_zy(owned, prec, g.ng, g.ng / 2 + 1, g.nx, c->fftdone.as<int>(), c->st)) return -1;
    c->launches += 2;
  }
  else if (c->slab_own) {
    const int ngk = g.ng / 2 + 1, zp = c->slab_zp;
    const size_t plane = (size_t) g.ng * ngk * 2 * prec;
    for (int x0 = 0; x0 < g.nx; x0 += zp) {
      char *grp = static_cast<char *>(owned) + (size_t) x0 * plane;
      if (c->slab_own_z) {
        if (launch_fft_rows(grp, grp, prec, g.ng, (long) zp * g.ng, 2 * (size_t) ngk, ngk, c->st))
          return -1;
      }
      else if (prec == 8) PSB_CUFFT(cufftExecD2Z(c->slab_yz, (cufftDoubleReal *) grp, (cufftDoubleComplex *) grp));
      else PSB_CUFFT(cufftExecR2C(c->slab_yz, (cufftReal *) grp, (cufftComplex *) grp));
      if (launch_fft_strided(grp, prec, g.ng, ngk, 1, zp, nullptr, nullptr, 0.0, c->st)) return -1;
      c->launches += 2;
    }
  }
  else if (prec == 8)
    PSB_CUFFT(cufftExecD2Z(c->slab_yz, (cufftDoubleReal *) owned, (cufftDoubleComplex *) owned));
  else
    PSB_CUFFT(cufftExecR2C(c->slab_yz, (cufftReal *) owned, (cufftComplex *) owned));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  c->launches++;
  return 0;
}

// pack the owned planes (nx, Ng, Ngk) complex into the all-to-all send layout
// [dest rank q][x local][y in slab q][k]
int psb_slab_pack(psb_context *c, const psb_params *par, const psb_slab *sl, const void *owned,
    void *sendbuf) {
  if (!c) return -1;
  AssignGeom g;
  if (slab_geom(c, par, sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  const size_t csz = 2 * (size_t) par->precision, ngk = g.ng / 2 + 1, ny = g.nx;
  const size_t width = ny * ngk * csz, spitch = (size_t) g.ng * ngk * csz;
  for (int q = 0; q < sl->nranks; q++)
    PSB_CUDA(cudaMemcpy2DAsync(static_cast<char *>(sendbuf) + (size_t) q * g.nx * width, width,
        static_cast<const char *>(owned) + (size_t) q * width, spitch, width, g.nx,
        cudaMemcpyDeviceToDevice, c->st));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

// 1-D c2c along x of the received (Ng_x, ny, Ngk) block, in place
int psb_slab_fft_x(psb_context *c, const psb_params *par, const psb_slab *sl, void *buf) {
  if (!c) return -1;
  AssignGeom g;
  if (slab_geom(c, par, sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  if (slab_plans(c, g.ng, g.nx, par->precision)) return -1;
  if (c->slab_own_x) {
    if (launch_fft_strided(buf, par->precision, g.ng, g.ng / 2 + 1, 0, g.nx, nullptr, nullptr, 0.0, c->st))
      return -1;
  }
  else if (par->precision == 8)
    PSB_CUFFT(cufftExecZ2Z(c->slab_x, (cufftDoubleComplex *) buf, (cufftDoubleComplex *) buf, CUFFT_FORWARD));
  else
    PSB_CUFFT(cufftExecC2C(c->slab_x, (cufftComplex *) buf, (cufftComplex *) buf, CUFFT_FORWARD));
  PSB_CUDA(cudaStreamSynchronize(c->st));
  c->launches++;
  return 0;
}

// raw multipole sums of the rank's y-slab, accumulated into pl_dev[nl*nbin]
// (device memory, so that the host can allreduce it); F*1 NULL without interlacing
int psb_slab_bin(psb_context *c, const psb_params *par, const psb_slab *sl, const void *Fa0,
    const void *Fa1, const void *Fb0, const void *Fb1, double *pl_dev) {
  if (!c) return -1;
  AssignGeom g;
  if (slab_geom(c, par, sl, g)) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  if (!same_bins(c, par) && prepare_bins(c, par)) return -1;
  BinGeom bg = c->bg;
  bg.j0 = g.x0; bg.nj = g.nx;         // y-slab after the transpose: same split as x
  PSB_CUDA(cudaStreamWaitEvent(c->st, c->ev_geom, 0));
  double *scratch_bin = reinterpret_cast<double *>(c->binscratch.as<char>() + c->bin_sb);
  if (launch_bin(bg, par->precision, Fa0, par->intlace ? Fa1 : nullptr, Fb0,
        par->intlace ? Fb1 : nullptr, pl_dev, scratch_bin, c->bin_sb, c->st))
    return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  c->launches += 2;
  return 0;
}

// mode counts (identical on every rank: pure geometry) + normalisation of the
// allreduced sums; pl0/pl1/xpl: host arrays of nl*nbin raw sums or NULL
psb_result *psb_slab_finish(psb_context *c, const psb_params *par, const double *pl0,
    const double *pl1, const double *xpl, const double wdata[2]) {
  if (!c) { set_error("no device context\n"); return nullptr; }
  psb_slab sl = {1, 0};
  AssignGeom g;
  if (slab_geom(c, par, &sl, g)) return nullptr;
  if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
  if (!same_bins(c, par) && prepare_bins(c, par)) return nullptr;
  const int nbin = c->nbin, nl = par->npole, nc = par->ncat;
  const size_t nacc = (size_t) nl * nbin;
  psb_result *res = new psb_result();
  res->nbin = nbin; res->nl = nl;
  res->kedge = c->kedge; res->k.resize(nbin); res->km.assign(nbin, 0);
  res->cnt.assign(nbin, 0); res->lcnt.assign(nacc, 0);
  for (int i = 0; i < nbin; i++) res->k[i] = (res->kedge[i] + res->kedge[i + 1]) * 0.5;
  std::vector<double> hb(2 * (size_t) nbin + nacc);
  if (cudaStreamWaitEvent(c->st, c->ev_geom, 0) != cudaSuccess ||
      cudaMemcpyAsync(hb.data(), c->bins.p, hb.size() * sizeof(double), cudaMemcpyDeviceToHost,
        c->st) != cudaSuccess || cudaStreamSynchronize(c->st) != cudaSuccess) {
    set_error("failed to read the mode counts back\n");
    delete res;
    return nullptr;
  }
  memcpy(res->cnt.data(), hb.data(), nbin * sizeof(double));
  memcpy(res->km.data(), hb.data() + nbin, nbin * sizeof(double));
  memcpy(res->lcnt.data(), hb.data() + 2 * nbin, nacc * sizeof(double));
  for (int b = 0; b < nbin; b++) if (res->cnt[b]) res->km[b] /= res->cnt[b];
  const double *src[2] = {pl0, pl1};
  for (int i = 0; i < 2; i++)
    if (src[i]) { res->has_pl[i] = true; res->pl[i].assign(src[i], src[i] + nacc); }
  if (xpl) { res->has_xpl = true; res->xpl.assign(xpl, xpl + nacc); }
  double shot[2] = {0, 0}, norm[2] = {0, 0};
  const double vol = c->bsize[0] * c->bsize[1] * c->bsize[2];
  for (int i = 0; i < nc; i++) { shot[i] = vol / wdata[i]; norm[i] = wdata[i] * wdata[i] / vol; }
  normalise(res, par, true, nc, shot, norm);
  for (int i = 0; i < 2; i++) { res->shot[i] = shot[i]; res->norm[i] = norm[i]; }
  for (int a = 0; a < 3; a++) { res->bmin[a] = 0; res->bsize[a] = c->bsize[a]; res->bmax[a] = c->bsize[a]; }
  c->bins_ready = false;
  collect_timings(c);
  return res;
}

double *psb_generate_catalog(psb_context *c, size_t n, double boxsize, int kind, uint64_t seed) {
  if (!c) { set_error("no device context\n"); return nullptr; }
  if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
  void *p = nullptr;
  if (cudaMalloc(&p, (n ? n : 1) * 32) != cudaSuccess) {
    set_error("failed to allocate the synthetic catalogue\n");
    cudaGetLastError();
    return nullptr;
  }
  if (launch_generate((double *) p, n, boxsize, kind, seed, c->st) ||
      cudaStreamSynchronize(c->st) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  return (double *) p;
}

int psb_generate_into(psb_context *c, double *dst_dev, size_t n, double boxsize, int kind, uint64_t seed,
    uint64_t first_index) {
  if (!c) { set_error("no device context\n"); return -1; }
  PSB_CUDA(cudaSetDevice(c->device));
  if (launch_generate_at(dst_dev, n, boxsize, kind, seed, first_index, c->st)) return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

void psb_device_free(psb_context *c, void *ptr) {
  if (c) cudaSetDevice(c->device);
  if (ptr) cudaFree(ptr);
}

// one in-place forward pass of the hand-written strided FFT on caller-owned device
// memory (building block of the slab path; tests)
int psb_fft_axis(psb_context *c, void *data_dev, int precision, int ng, int ngk, int axis,
    int outer_n) {
  if (!c) { set_error("no device context\n"); return -1; }
  if (!fft_strided_supported(ng, precision) || axis < 0 || axis > 3 || ngk < 1 || outer_n < 1 ||
      (axis >= 2 && ngk != ng / 2 + 1)) {
    set_error("psb_fft_axis: unsupported size %d / precision %d / axis %d\n", ng, precision, axis);
    return -1;
  }
  PSB_CUDA(cudaSetDevice(c->device));
  if (axis == 3) {
    if (c->fftdone.reserve(sizeof(int) * (size_t) outer_n)) return -1;
    if (launch_fft_zy(data_dev, precision, ng, ngk, outer_n, c->fftdone.as<int>(), c->st)) return -1;
  }
  else if (axis == 2) {
    if (launch_fft_rows(data_dev, data_dev, precision, ng, outer_n, 2 * (size_t) ngk, ngk, c->st))
      return -1;
  }
  else if (launch_fft_strided(data_dev, precision, ng, ngk, axis, outer_n, nullptr, nullptr, 0.0, c->st))
    return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

int psb_catalog_probe(const char *path, size_t *nrow, int *ncol, int *elem_bytes) {
  size_t off = 0, n = 0;
  int nc = 0, el = 0;
  if (!path || npy_probe(path, &off, &n, &nc, &el)) return -1;
  if (nrow) *nrow = n;
  if (ncol) *ncol = nc;
  if (elem_bytes) *elem_bytes = el;
  return 0;
}

int psb_catalog_load(psb_context *c, const char *path, const psb_columns *cols, int issim,
    double **records_dev, psb_catalog_sums *sums) {
  if (!c) { set_error("no device context\n"); return -1; }
  if (!path || !cols || !records_dev || !sums) { set_error("catalogs not read\n"); return -1; }
  size_t off = 0, n = 0;
  int ncol = 0, elem = 0;
  if (npy_probe(path, &off, &n, &ncol, &elem)) return -1;
  const int want[6] = {cols->pos[0], cols->pos[1], cols->pos[2], cols->wcomp, cols->wfkp, cols->nz};
  for (int q = 0; q < 6; q++)
    if (want[q] >= ncol || (q < 3 && want[q] < 0)) {
      set_error("not enough columns in `%s' (%d) for column %d\n", path, ncol, want[q]);
      return -1;
    }
  PSB_CUDA(cudaSetDevice(c->device));
  reset_timings(c);
  const size_t rowb = (size_t) ncol * elem, payload = n * rowb;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { set_error("cannot open file for reading: `%s'\n", path); return -1; }
  struct stat sb;
  if (fstat(fd, &sb) || (size_t) sb.st_size < off + payload) {
    close(fd);
    set_error("truncated .npy file: `%s'\n", path);
    return -1;
  }
  void *map = n ? mmap(nullptr, off + payload, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  close(fd);
  if (n && map == MAP_FAILED) { set_error("cannot map `%s'\n", path); return -1; }
  if (n) madvise(map, off + payload, MADV_SEQUENTIAL);
  const char *src = static_cast<const char *>(map) + off;
  double *rec = nullptr;
  DevBuf partial;
  auto fail = [&]() { partial.release(); if (rec) cudaFree(rec); if (n) munmap(map, off + payload); return -1; };
  if (cudaMalloc(&rec, n ? n * 32 : 32) != cudaSuccess) { set_error("out of device memory for the catalogue\n"); return fail(); }
  const size_t CH = (size_t) 1 << 22;                       // rows per chunk
  const size_t nchunk = (n + CH - 1) / CH;
  const int nblk = assemble_blocks();
  if (partial.reserve(sizeof(double) * 3 * nblk * (nchunk ? nchunk : 1))) return fail();
  for (int s = 0; s < 2; s++) {
    if (nchunk > (size_t) s && c->chunkbuf[s].reserve(std::min(n, CH) * rowb)) return fail();
    if (!c->ev_filled[s]) {
      if (cudaEventCreateWithFlags(&c->ev_filled[s], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_consumed[s], cudaEventDisableTiming) != cudaSuccess)
        return fail();
    }
  }
  size_t k = 0;
  for (size_t r0 = 0; r0 < n; r0 += CH, k++) {
    const size_t len = std::min(CH, n - r0);
    const int s = (int) (k & 1);
    if (cudaStreamWaitEvent(c->st_copy, c->ev_consumed[s], 0) != cudaSuccess) return fail();
    {
      StageScope sc(c, PSB_T_H2D, c->st_copy);
      if (h2d_async(c, c->chunkbuf[s].p, src + r0 * rowb, len * rowb, false, c->st_copy)) return fail();
    }
    if (cudaEventRecord(c->ev_filled[s], c->st_copy) != cudaSuccess ||
        cudaStreamWaitEvent(c->st, c->ev_filled[s], 0) != cudaSuccess)
      return fail();
    if (launch_assemble(c->chunkbuf[s].p, elem, len, cols->pos, cols->wcomp, cols->wfkp, cols->nz, ncol,
          issim, rec + 4 * r0, partial.as<double>() + 3 * (size_t) nblk * k, c->st))
      return fail();
    c->launches++;
    if (cudaEventRecord(c->ev_consumed[s], c->st) != cudaSuccess) return fail();
  }
  std::vector<double> h(3 * (size_t) nblk * nchunk);
  if (!h.empty() && cudaMemcpyAsync(h.data(), partial.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost,
        c->st) != cudaSuccess)
    return fail();
  if (cudaStreamSynchronize(c->st) != cudaSuccess) { set_error("catalogue ingest failed on the device\n"); return fail(); }
  partial.release();
  if (n) munmap(map, off + payload);
  collect_timings(c);
  sums->n = n; sums->sumw = sums->sumw2 = sums->sumw2n = 0;
  for (size_t q = 0; q + 2 < h.size(); q += 3) {
    sums->sumw += h[q]; sums->sumw2 += h[q + 1]; sums->sumw2n += h[q + 2];
  }
  *records_dev = rec;
  return 0;
}

// host-only part of the integration mode: the order cnvt_coord would choose
int psb_cnvt_order(const psb_cosmo *cm, double zmin, double zmax) {
  if (!cm || zmin < 0 || zmin > zmax) { set_error("invalid redshift value in the catalogs\n"); return -1; }
  const double widx = (cm->eos_w == -1) ? 0 : 3 * (1 + cm->eos_w);
  const int order = legauss_order(cm->omega_m, cm->omega_l, cm->omega_k, widx, cm->ecdst, zmin, zmax, 128);
  if (order == INT_MAX) { set_error("failed to perform the convergency test for integrations\n"); return -1; }
  return order;
}

int psb_cnvt_coord(psb_context *c, const psb_cosmo *cosmo, double *const *arrays_dev,
    const size_t *counts, int narrays, int *order) {
  if (!c) { set_error("no device context\n"); return -1; }
  if (!cosmo || !arrays_dev || !counts || narrays < 0) { set_error("catalogs not read\n"); return -1; }
  PSB_CUDA(cudaSetDevice(c->device));
  reset_timings(c);
  if (convert_arrays(c, cosmo, arrays_dev, counts, narrays, order)) return -1;
  PSB_CUDA(cudaStreamSynchronize(c->st));
  collect_timings(c);
  return 0;
}

int psb_copy_to_host(psb_context *c, void *dst, const void *src, size_t bytes) {
  if (!c) return -1;
  PSB_CUDA(cudaSetDevice(c->device));
  PSB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
