// Internal: the device context (streams, buffers, FFT plans, options) and the host-side
// helpers shared by the translation units that orchestrate kernels (context.cu: one
// GPU; dist.cu: one mesh over several GPUs).  Not installed; the public ABI is
// include/powspec_b200.h.
#pragma once

#include "psb_internal.h"
#include "../../include/powspec_b200.h"

#include <cufft.h>

#include <vector>

namespace psb {

// ---------------------------------------------------------------------------
// growable device buffer
// ---------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      set_error("failed to allocate %.3f GB of device memory: %s\n", bytes / 1e9,
          cudaGetErrorString(e));
      cudaGetLastError();
      return -1;
    }
    cap = bytes;
    return 0;
  }
  // grow, keeping the first `keep` bytes (copied on `st`; cudaFree waits for the device)
  int reserve_keep(size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= cap) return 0;
    void *np = nullptr;
    const size_t want = bytes + bytes / 2;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
      set_error("failed to allocate %.3f GB of device memory: %s\n", want / 1e9, cudaGetErrorString(e));
      cudaGetLastError();
      return -1;
    }
    if (p && keep) cudaMemcpyAsync(np, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, st);
    if (p) cudaFree(p);
    p = np; cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct Interval { int stage; cudaEvent_t a, b; };

}  // namespace psb

using psb::DevBuf;
using psb::Interval;
using psb::AssignGeom;
using psb::BinGeom;

#define PSB_STAGE_SLOTS 8

// staging copy pageable -> pinned memory by a persistent thread pool (hostcopy.cpp)
namespace psb_host {
struct CopyPool;
CopyPool *copy_pool_create(int nthreads);
void copy_pool_destroy(CopyPool *p);
int copy_pool_threads(const CopyPool *p);
void copy_pool_run(CopyPool *p, void *dst, const void *src, size_t bytes);
void copy_set_stream_stores(int on);
}  // namespace psb_host

// ---------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------
struct psb_context {
  int device = 0;
  int sms = 148;
  cudaStream_t st = nullptr;            // compute
  cudaStream_t st_geom = nullptr;       // data-independent mode counting
  cudaEvent_t ev_geom = nullptr;
  cudaStream_t st_aux = nullptr;        // mesh memsets, overlapped with the particle sort
  cudaEvent_t ev_aux_go = nullptr, ev_aux_done = nullptr, ev_memset[2] = {nullptr, nullptr};
  cudaEvent_t memset_pending = nullptr; // the scatter must wait for this memset first

  // particles
  DevBuf part_in[2][2];                 // [cat][data|rand] staged copies of host arrays
  DevBuf chunkbuf[2];                   // double-buffered device chunks of a streamed catalogue
  cudaStream_t st_copy = nullptr;       // H2D engine stream of the streaming path
  cudaEvent_t ev_filled[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  DevBuf sorted, keys, hist, cursor, cubtmp, bounds_part;
  DevBuf tile_cnt, tile_start, wmax_buf;  // owner-computes assignment: list counts / offsets, max |w|
  DevBuf tile_ovrec, tile_ovtile;         // one-pass lists: overflow entries (records, tile indices + counter)
  size_t tile_scan_bytes = 0;
  int tile_onepass_backoff = 0;           // dense chunks left that go straight to the exact lists
  uint32_t tile_overflowed = 0;           // overflow entries of the last dense chunk
  int assign_path = 0;                  // what the last scatter used: 0 global reductions, 1 owner-computes tiles
  size_t bounds_used = 0;               // bytes of bounds_part holding deferred bounds partials
  void *pinned_base = nullptr;          // staging ring for pageable sources: pinned_slots pieces of pinned_bytes
  size_t pinned_bytes = 0;
  int pinned_slots = 0, pinned_next = 0;
  bool pinned_wc = false;               // the ring is write-combined memory
  cudaEvent_t pinned_free[PSB_STAGE_SLOTS] = {};
  psb_host::CopyPool *copy_pool = nullptr;      // persistent staging threads (hostcopy.cpp)

  // meshes: [cat][field]; survey extras
  DevBuf mesh[2][2];
  DevBuf fkl[2], fka, fk0copy[2];

  // FFT
  cufftHandle plan_fwd = 0, plan_inv = 0, plan_z = 0, plan_z2 = 0, plan_x = 0;
  bool have_z = false, have_z2 = false, have_x = false;
  bool own_fft = false;                 // y/x passes by k_fft_strided (fft_strided.cu)
  double fft_k2max = 0;                 // last bin edge in k^2 (tile skipping of the x pass)
  int plan_ng = 0, plan_prec = 0, plan_zp = 0;
  bool have_fwd = false, have_inv = false;
  DevBuf fftwork, fftdone, cnvt_tab;

  // slab-decomposed FFT plans
  cufftHandle slab_yz = 0, slab_x = 0;
  int slab_ng = 0, slab_nx = 0, slab_prec = 0, slab_zp = 1;
  bool slab_have = false, slab_own = false, slab_own_x = false, slab_own_z = false;
  bool slab_has_yz = false, slab_has_x = false;

  // tables and bins
  DevBuf tables, binscratch, bins;
  std::vector<double> host_tables;

  // options
  long opt_sort = 1;
  long opt_sort_min = 1 << 16;
  long opt_geom_sym = 1;                // fold +-n_x, +-n_y in the mode-counting pass
  long opt_tile_onepass = 1;            // tile lists in one pass (fixed capacity + overflow list)
  long opt_tile_index = 0;              // 1: tile lists of 4-byte particle indices instead of record copies (ablation: slower)
  long opt_tile_cap = 0;                // > 0: slots per tile (tests: forces overflow)
  long opt_tile_ovcap = 0;              // > 0: room of the overflow list (tests: forces the fallback)
  long opt_owner = -1;                  // owner-computes tile assignment: 1 / 0 / -1 = when the chunk is dense
                                        // enough to pay for writing every mesh cell (>= Ntot / 32 particles)
  long opt_coop = 1;                    // z-coalesced scatter kernel
  long opt_coop_variant = 0;
  long opt_survey_direct = 1;           // survey l > 0: bin Fk0 x Fka_m directly (no Fkl field)
  long opt_fft_skip = 1;                // x pass skips the columns beyond the last bin edge
  long opt_fft_store_skip = 1;          // the y and x passes do not store cells beyond the last bin edge
  long opt_fft_l2_mb = 0;               // L2 budget of a z + y plane group (0: whole mesh at once)
  long opt_fft_streams = 1;             // 2: alternate the plane groups between two streams
  long opt_fft_fused = 0;               // z + y passes in one persistent kernel (L2 hand-over)
  long opt_fft_own_z = -1;              // hand-written r2c z pass: 1 / 0 (cuFFT batched 1-D) / -1 auto
  long opt_fft_own_x = -1;              // hand-written x pass: 1 / 0 (cuFFT strided batched 1-D) / -1 auto
  long opt_memset_overlap = 0;          // mesh memsets on a side stream, under the particle sort
  long opt_own_fft = 1;                 // hand-written strided FFT passes where available
  long opt_xgroup = 0;                  // > 0: coarse bucket sort (planes per bucket)
  long opt_strip = 64;                  // rows per strip of the sort order
  long opt_h2d_threads = 16;            // host threads staging pageable memory into pinned buffers
                                        // (capped at the hardware concurrency; 8 -> 16 on the 16-core
                                        // B200 host: 36 -> 44 GB/s, config 2 from malloc'd memory 121 -> 107 ms;
                                        // since round 2 a persistent pool with streaming stores, hostcopy.cpp)
  long opt_h2d_wc = 0;                  // staging ring in write-combined pinned memory
  long opt_h2d_piece_mb = 16;           // bytes per piece of the staging ring (MiB) and
  long opt_h2d_slots = 3;               // pieces in the ring (2 .. PSB_STAGE_SLOTS): 48 MB stay in the host's share
                                        // of the last-level cache.  ms per config-2 step from malloc'd memory
                                        // (profiles/r2_v1?_pageable_ring*.jsonl; pinned source: 90.6): 3 x 16 MB
                                        // 94.5 / 94.9, 4 x 12: 94.8 / 95.4, 4 x 8: 95.2 / 95.7, 4 x 16: 94.8 ... 107
                                        // (borderline), 6 x 16: 111.7, 2 x 16: 117 (too shallow), 8 x 2: 110.6;
                                        // streaming stores into 2 x 64 MB: 98.7 ... 100.1
  long opt_stream = 1;                  // overlap H2D with assignment for host catalogues (sims)
  long opt_stream_chunk = 12500000;     // particles per streamed chunk (400 MB): the smallest whose
                                        // sort + scatter (~6 ms, one sweep of the meshes) still keeps
                                        // up with its upload (7.2 ms); config 2 e2e 93.9 / 92.5 / 91.6 ms
                                        // with 20 M / 16.8 M / 12.5 M
  long opt_stream_taper = 0;            // > 0: the last chunks halve down to this many particles (measured: slower)

  // state carried from psb_mesh to psb_power
  bool mesh_ready = false;
  psb_params par;
  double bmin[3], bsize[3], bmax[3];
  double shot[2], norm[2];

  // k-bins, per-axis tables and mode counts prepared ahead of the FFTs
  bool bins_ready = false;
  psb_params bins_par;
  double bins_box[3];
  int nbin = 0;
  std::vector<double> kedge;
  BinGeom bg;
  size_t bin_sb = 0;

  // timings
  std::vector<Interval> intervals;
  std::vector<cudaEvent_t> evpool;
  double ms[PSB_T_COUNT];
  double host_h2d_ms = 0;
  long launches = 0;
};

struct psb_result {
  int nbin = 0, nl = 0;
  std::vector<double> k, kedge, km, lcnt, pl[2], xpl;
  std::vector<unsigned long long> cnt;
  bool has_pl[2] = {false, false}, has_xpl = false;
  double shot[2] = {0, 0}, norm[2] = {0, 0};
  double bmin[3], bsize[3], bmax[3];
};

namespace psb_host {

cudaEvent_t get_event(psb_context *c);

// CUDA-event interval of one stage on stream s, accumulated into c->ms[stage]
struct StageScope {
  psb_context *c; int stage; cudaStream_t s; cudaEvent_t a;
  StageScope(psb_context *c_, int stage_, cudaStream_t s_) : c(c_), stage(stage_), s(s_) {
    a = get_event(c);
    cudaEventRecord(a, s);
  }
  ~StageScope() {
    cudaEvent_t b = get_event(c);
    cudaEventRecord(b, s);
    c->intervals.push_back({stage, a, b});
  }
};

void reset_timings(psb_context *c);
void collect_timings(psb_context *c);
int check_params(const psb_params *p);
int h2d_async(psb_context *c, void *dst, const void *src, size_t bytes, bool pinned,
    cudaStream_t stream);
bool is_pinned(const void *p);
// fresh (optional): *fresh says the meshes have not been initialised yet; the first
// scatter then either stores whole tiles (owner-computes path) or zeroes them first
int assign_catalog(psb_context *c, const double *dev, size_t n, const AssignGeom &g, int scheme,
    int precision, double wscale, void *m0, void *m1, bool bounds = false, bool *fresh = nullptr);
void normalise(psb_result *res, const psb_params *par, bool issim, int nc, const double *shot,
    const double *norm);
bool same_bins(const psb_context *c, const psb_params *p);
int prepare_bins(psb_context *c, const psb_params *par);
int slab_geom(psb_context *c, const psb_params *par, const psb_slab *sl, AssignGeom &g);
int slab_plans(psb_context *c, int ng, int nx, int prec);
bool fft_own_z(const psb_context *c, int ng, int precision);
bool fft_own_x(const psb_context *c, int ng, int precision);

}  // namespace psb_host
