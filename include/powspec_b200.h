/*
 * powspec_b200.h — C ABI of the B200-native replacement for powspec's hot path
 * (mass assignment -> r2c FFT -> window-deconvolved multipole binning).
 *
 * Plain C: no CUDA / torch types in any signature.  Every entry point cites the
 * reference interface it replaces (paths relative to cheng-zhao/powspec).
 *
 * Two layers are exported by libpowspec_b200.so:
 *
 *  1. the "psb_" API below: the same stages with plain structs, used by the
 *     Python host mirror (powspec_b200/api.py), the tests and bench.py;
 *  2. the five symbols of the reference's own seam (include/powspec_refabi.h):
 *     genr_mesh, mesh_destroy, powspec, powspec_destroy, powspec_assign_names —
 *     what the reference's unchanged C host links against instead of
 *     genr_mesh.o / multipole.o.
 *
 * Error convention: functions returning pointers return NULL on
 * failure, functions returning int return non-zero; a message in the reference's
 * P_ERR style ("\n\x1B[31;1mError:\x1B[0m ...", src/define.h:102,129) is written
 * to stderr and also kept for psb_last_error().  There is NO CPU fallback: if no
 * CUDA device is usable every compute entry fails loudly.
 */
#ifndef POWSPEC_B200_H
#define POWSPEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_VERSION 1
#define PSB_MAX_ELL 6           /* POWSPEC_MAX_ELL, src/define.h:94 */
#define PSB_MAX_POLES 8

/* particle assignment schemes: powspec_assign_t, src/genr_mesh.h:38-43 */
enum { PSB_NGP = 0, PSB_CIC = 1, PSB_TSC = 2, PSB_PCS = 3 };

/* A particle record is the reference's DATA {double x[3]; double w;}
 * (src/read_cata.h:42-45): 4 doubles, 32 bytes, array of structures. */

/* The members of CONF (src/load_conf.h:40-93) that genr_mesh() and powspec()
 * read (src/genr_mesh.c:654-669,683,896-897; src/multipole.c:310-319,341-344),
 * plus the two choices that are compile-time in the reference
 * (-DSINGLE_PREC, Makefile:14) or absent (device). */
typedef struct {
  int ncat;             /* CONF.ndata: number of catalogues, 1 or 2            */
  int issim;            /* CONF.issim   (CUBIC_SIM)                            */
  int intlace;          /* CONF.intlace (GRID_INTERLACE)                       */
  int assign;           /* CONF.assign  (PARTICLE_ASSIGN) PSB_NGP..PSB_PCS     */
  int gsize;            /* CONF.gsize   (GRID_SIZE)                            */
  int logscale;         /* CONF.logscale: kmin/kmax/kbin are then in log10     */
  int verbose;          /* CONF.verbose: print the reference's progress lines  */
  int npole;            /* CONF.npole                                          */
  int poles[PSB_MAX_POLES];     /* CONF.poles: sorted, unique, 0..6            */
  int has_bsize;        /* CONF.bsize != NULL                                  */
  int isauto[2];        /* CONF.isauto                                         */
  int iscross;          /* CONF.iscross                                        */
  double los[3];        /* CONF.los   (LINE_OF_SIGHT, sims)                    */
  double bsize[3];      /* CONF.bsize (BOX_SIZE)                               */
  double bpad[3];       /* CONF.bpad  (BOX_PAD, surveys without BOX_SIZE)      */
  double kmin, kmax, kbin;      /* CONF.kmin/kmax/kbin; kmax <= 0: unset       */
  int precision;        /* sizeof mesh real: 8 (default build) or 4 (SINGLE_PREC) */
  int device;           /* CUDA device ordinal                                  */
} psb_params;

/* Where the particle arrays live. */
enum {
  PSB_MEM_HOST = 0,     /* host memory (pageable or pinned; detected)          */
  PSB_MEM_DEVICE = 1    /* already resident on `device` (bench "value" leg)    */
};

/* Coordinate conversion (RA, Dec, redshift) -> comoving Cartesian, the members of
 * CONF that cnvt_coord() reads (src/cnvt_coord.c:440-582, src/load_conf.h:66-73).
 * sample_z != NULL: cubic-spline interpolation of the (z, distance) table that
 * CONF.fcdst names (the host reads the file); otherwise Legendre-Gauss
 * integration with the order chosen for the error ecdst. */
typedef struct {
  double omega_m, omega_l, omega_k, eos_w;      /* CONF.omega_m/omega_l/omega_k/eos_w */
  double ecdst;                                 /* CONF.ecdst                         */
  const double *sample_z, *sample_d;            /* CONF.fcdst contents (host), or NULL */
  size_t nsample;
} psb_cosmo;

/* The members of CATA (src/read_cata.h:47-59) the path reads. */
typedef struct {
  const double *data[2];        /* CATA.data[i]: ndata[i] x {x,y,z,w}          */
  const double *rand[2];        /* CATA.rand[i] (surveys)                      */
  size_t ndata[2], nrand[2];
  double wdata[2], wrand[2];    /* CATA.wdata / wrand                          */
  double alpha[2];              /* CATA.alpha (surveys, from read_cata)        */
  double shot[2], norm[2];      /* surveys: inputs; sims: computed by the path
                                   (src/genr_mesh.c:904-909)                   */
  int memspace;                 /* PSB_MEM_HOST / PSB_MEM_DEVICE               */
  /* optional: convert the coordinates on the device before anything else
   * (replaces the host pass cnvt_coord(), src/powspec.c:39); dcnvt / rcnvt are
   * CONF.dcnvt / CONF.rcnvt (DATA_CONVERT / RAND_CONVERT) per catalogue.  The
   * caller's arrays are not modified. */
  const psb_cosmo *cnvt;
  int dcnvt[2], rcnvt[2];
} psb_cats;

typedef struct psb_context psb_context;         /* device, stream, buffers, FFT plans */
typedef struct psb_result psb_result;           /* what PK + MESH metadata carry       */

/* Context: owns the CUDA stream, mesh buffers, cuFFT plans and scratch for one
 * device; buffers are grown on demand and reused across runs.  Replaces
 * mesh_init / mesh_destroy (src/genr_mesh.c:650-747, 615-640) and the FFTW plan
 * handling of src/fftw_define.h:32-64.  Returns NULL if no CUDA device. */
psb_context *psb_create(int device);
void psb_destroy(psb_context *ctx);

/* The whole replaced span: genr_mesh() (src/genr_mesh.c:874-926) followed by
 * powspec() (src/multipole.c:1179-1278).  The particle arrays are only read. */
psb_result *psb_run(psb_context *ctx, const psb_params *par, const psb_cats *cats);

/* Stage-level entry points (the same work in two calls, as the reference's
 * main() makes them, src/powspec.c:47,55).  psb_mesh leaves the density meshes
 * on the device inside ctx; psb_power consumes them. */
int psb_mesh(psb_context *ctx, const psb_params *par, const psb_cats *cats);
psb_result *psb_power(psb_context *ctx, const psb_params *par);

void psb_result_free(psb_result *res);

/* Result accessors (PK members, src/multipole.h:38-61; MESH metadata,
 * src/genr_mesh.h:48-70; CATA.shot/norm). */
enum {
  PSB_GET_K = 0,        /* double[nbin]      PK.k                              */
  PSB_GET_KEDGE,        /* double[nbin+1]    PK.kedge                          */
  PSB_GET_KM,           /* double[nbin]      PK.km                             */
  PSB_GET_CNT,          /* uint64[nbin]      PK.cnt                            */
  PSB_GET_LCNT,         /* double[nl*nbin]   PK.lcnt (sims)                    */
  PSB_GET_PL,           /* double[nl*nbin]   PK.pl[idx]                        */
  PSB_GET_XPL,          /* double[nl*nbin]   PK.xpl                            */
  PSB_GET_SHOT,         /* double[2]         CATA.shot                         */
  PSB_GET_NORM,         /* double[2]         CATA.norm                         */
  PSB_GET_BMIN,         /* double[3]         MESH.min                          */
  PSB_GET_BSIZE,        /* double[3]         MESH.bsize                        */
  PSB_GET_BMAX          /* double[3]         MESH.max                          */
};
int psb_result_nbin(const psb_result *res);
int psb_result_nl(const psb_result *res);
/* copies into dst, returns the element count or -1 if absent */
long psb_result_get(const psb_result *res, int what, int idx, void *dst);

/* Copy a density mesh out of the context (tests): field 0 = Fr, 1 = Frl (the
 * half-cell shifted one), catalogue `cat`; dst holds gsize^3 reals of the
 * run's precision, unpadded, z fastest (IDX, src/define.h:134).  Only valid
 * between psb_mesh and psb_power. */
int psb_copy_mesh(psb_context *ctx, int cat, int field, void *dst);

/* Box defined by the last psb_mesh (def_box, src/genr_mesh.c:509-578): MESH.min,
 * MESH.bsize, MESH.max. */
int psb_mesh_box(const psb_context *ctx, double bmin[3], double bsize[3], double bmax[3]);

/* Device timings of the last run, in milliseconds, measured with CUDA events on
 * the context's stream.  Index with PSB_T_*.  PSB_T_FFT covers the whole transforms;
 * PSB_T_FFT_STRIDED is the part of it spent in the hand-written strided passes. */
enum {
  PSB_T_H2D = 0, PSB_T_BOUNDS, PSB_T_SORT, PSB_T_MEMSET, PSB_T_ASSIGN,
  PSB_T_FFT, PSB_T_GEOM, PSB_T_BIN, PSB_T_YLM, PSB_T_FFT_STRIDED, PSB_T_CNVT, PSB_T_TOTAL,
  PSB_T_COUNT
};
int psb_timings(const psb_context *ctx, double *ms, int n);
/* number of kernel launches (ours + cuFFT calls counted as 1) of the last run */
long psb_launch_count(const psb_context *ctx);
/* the context's compute stream (a cudaStream_t), e.g. to record timing events on it */
void *psb_stream(const psb_context *ctx);
/* the scatter the last psb_mesh used: 0 = global reductions, 1 = owner-computes tiles */
int psb_assign_path(const psb_context *ctx);
/* owner-computes path: list entries of the last dense chunk that did not fit their tile's
 * slots and were added through the overflow list (0 for catalogues near uniform on the tile
 * scale; -1 without a context) */
long psb_tile_overflow(const psb_context *ctx);

/* Tunables (tests / ablations; the list is in psb_set_option, csrc/context.cu):
 * "sort", "strip", "coop", "owner", "own_fft", "fft_fused", "stream", "stream_chunk",
 * "stream_taper", "h2d_threads", "h2d_nt" (0: plain memcpy in the staging pool), "h2d_wc" (1: write-combined staging buffers), "h2d_piece_mb", "h2d_slots" (the staging ring), "survey_direct", "geom_sym", and for the owner-computes
 * assignment "tile_onepass", "tile_cap", "tile_ovcap", "tile_index", "tile_tma",
 * "tile_fill_unroll", for the FFT passes "fft_skip", "fft_store_skip", "fft_variant",
 * "geom_blocks" ...; returns non-zero for an unknown name */
int psb_set_option(psb_context *ctx, const char *name, long value);

const char *psb_last_error(void);
int psb_device_count(void);

/* ---------------------------------------------------------------------------
 * Slab-decomposed mesh for meshes that exceed or strain one GPU.
 * The reference has no distributed path (one address space, src/genr_mesh.c:650-747);
 * this is new design.  Rank r of nranks owns the x-planes [r*Ng/nranks,
 * (r+1)*Ng/nranks) of every real field.  These entry points are the per-rank
 * building blocks; the exchanges between them — particle routing (all-to-all-v),
 * halo planes (neighbour send/recv + psb_add), FFT transpose (all-to-all),
 * bin sums (allreduce) — are issued by the host over NCCL
 * (powspec_b200/distributed.py).  All buffers are caller-owned device memory.
 * Simulation boxes only. */
typedef struct { int nranks, rank; } psb_slab;
#define PSB_HALO_LO 1   /* halo planes below the slab (TSC/PCS reach i0-1)               */
#define PSB_HALO_HI 3   /* above: i0+2, +1 for the half-cell shifted (interlaced) field */
/* reals in one slab buffer: (nx + halos) * Ng * 2(Ng/2+1); nranks == 1: no halos */
size_t psb_slab_mesh_elems(const psb_params *par, const psb_slab *slab);
int psb_slab_partition(psb_context *ctx, const psb_params *par, int nranks,
    const double *particles_dev, size_t n, double *sorted_dev, size_t *counts);
int psb_slab_assign(psb_context *ctx, const psb_params *par, const psb_slab *slab,
    const double *particles_dev, size_t n, double wscale, void *mesh0, void *mesh1);
int psb_add(psb_context *ctx, void *dst, const void *src, size_t n, int precision);
int psb_slab_fft_yz(psb_context *ctx, const psb_params *par, const psb_slab *slab, void *owned);
int psb_slab_pack(psb_context *ctx, const psb_params *par, const psb_slab *slab,
    const void *owned, void *sendbuf);
int psb_slab_fft_x(psb_context *ctx, const psb_params *par, const psb_slab *slab, void *buf);
int psb_slab_bin(psb_context *ctx, const psb_params *par, const psb_slab *slab,
    const void *Fa0, const void *Fa1, const void *Fb0, const void *Fb1, double *pl_dev);
psb_result *psb_slab_finish(psb_context *ctx, const psb_params *par, const double *pl0,
    const double *pl1, const double *xpl, const double wdata[2]);

/* ---------------------------------------------------------------------------
 * The slab decomposition driven from inside the library (csrc/dist.cu): the exchanges
 * are issued by the library itself, stream-ordered, over one of two transports.
 *
 *  - psb_dist: ONE rank.  One process per GPU (bench.py under torchrun): rank 0 calls
 *    psb_dist_unique_id, the host distributes the 128 bytes, every rank calls
 *    psb_dist_create_nccl (libnccl.so.2 is dlopen'ed: the copy the host process has
 *    already loaded, else POWSPEC_B200_NCCL, else the loader path).
 *  - psb_group: ALL ranks in one process, one host thread per rank, exchanges by peer
 *    copies (no NCCL).  This is what the reference's single-process C host gets through
 *    genr_mesh()/powspec() with POWSPEC_B200_DEVICES=0,1,...; a device may be listed
 *    more than once (virtual ranks, tests).
 *
 * begin / add / finish are collective: every rank makes the same calls.  Simulation
 * boxes only.  Where the ranks can map each other's memory (same process, or CUDA IPC),
 * the y pass of the FFT stores its result straight into the destination ranks' buffers
 * (FFT + transpose in one kernel, no all-to-all); POWSPEC_B200_P2P=0 or option "p2p"
 * forces the all-to-all path. */
typedef struct psb_dist psb_dist;
typedef struct psb_group psb_group;
int psb_dist_unique_id(void *id128);
psb_dist *psb_dist_create_nccl(psb_context *ctx, int nranks, int rank, const void *id128);
void psb_dist_destroy(psb_dist *d);
int psb_dist_set_option(psb_dist *d, const char *name, long value);
int psb_dist_begin(psb_dist *d, const psb_params *par);
/* one chunk of THIS rank's share of catalogue `cat`: device memory, any distribution */
int psb_dist_add(psb_dist *d, int cat, const double *particles_dev, size_t n);
/* the same from host memory, in nchunks pieces (same number on every rank) whose uploads
 * overlap the routing and scatter of the previous piece */
int psb_dist_add_host(psb_dist *d, int cat, const double *particles_host, size_t n, int nchunks);
/* wdata: global sum of weights per catalogue; every rank gets the same result */
psb_result *psb_dist_finish(psb_dist *d, const double wdata[2]);
/* CUDA-event stage times of the last run on this rank, ms */
enum {
  PSB_D_ROUTE = 0, PSB_D_ASSIGN, PSB_D_HALO, PSB_D_FFT_ZY, PSB_D_TRANSPOSE, PSB_D_FFT_X, PSB_D_BIN,
  PSB_D_REDUCE, PSB_D_COUNT
};
int psb_dist_timings(const psb_dist *d, double *ms, int n);
/* [0] transpose bytes this rank sent, [1] particle bytes it routed away, [2] 1 if the
 * transposes were peer stores fused into the y pass */
int psb_dist_traffic(const psb_dist *d, double *out, int n);
const char *psb_dist_transport(const psb_dist *d);
psb_context *psb_dist_context(psb_dist *d);

psb_group *psb_group_create(const int *devices, int nranks);
void psb_group_destroy(psb_group *g);
int psb_group_size(const psb_group *g);
psb_dist *psb_group_rank(psb_group *g, int r);
int psb_group_set_option(psb_group *g, const char *name, long value);
/* genr_mesh() / powspec() over the group: rank r takes the r-th contiguous share of
 * every catalogue (host memory, or device memory of one GPU) */
int psb_group_mesh(psb_group *g, const psb_params *par, const psb_cats *cats);
psb_result *psb_group_power(psb_group *g, const psb_params *par);
psb_result *psb_group_run(psb_group *g, const psb_params *par, const psb_cats *cats);

/* Binary catalogue ingest (the step before the boundary): a
 * NumPy .npy file — 2-D, C order, little-endian float64 or float32, shape
 * (N, ncols) — is streamed from the page cache to the device and turned into the
 * particle records and catalogue sums that read_ascii_data() (io/read_ascii.c:868-902)
 * produces from the same numbers: {x, y, z, w = wcomp * wfkp} (sims: w = wcomp),
 * sumw = sum wcomp (CATA.wdata / wrand), sumw2 = sum w^2, sumw2n = sum wcomp wfkp^2 n(z).
 * Columns are 0-based; -1 = absent (wcomp = 1, wfkp = 1, n(z) = 0).  The records are
 * device memory owned by the caller (psb_device_free) and go to psb_mesh with
 * memspace PSB_MEM_DEVICE. */
typedef struct { int pos[3]; int wcomp, wfkp, nz; } psb_columns;
typedef struct { size_t n; double sumw, sumw2, sumw2n; } psb_catalog_sums;
/* header only (no device needed): rows, columns, element size in bytes */
int psb_catalog_probe(const char *path, size_t *nrow, int *ncol, int *elem_bytes);
int psb_catalog_load(psb_context *ctx, const char *path, const psb_columns *cols, int issim,
    double **records_dev, psb_catalog_sums *sums);

/* cnvt_coord() (src/cnvt_coord.c:549-582) on DEVICE-resident particle arrays, in
 * place: arrays_dev[i] holds counts[i] records {RA deg, Dec deg, z, w}.  The
 * Legendre-Gauss order is chosen from the redshift range of all the arrays
 * (:495-511) and returned in *order (0 in interpolation mode).  Errors as the
 * reference: negative redshift, no convergence up to order 32. */
int psb_cnvt_coord(psb_context *ctx, const psb_cosmo *cosmo, double *const *arrays_dev,
    const size_t *counts, int narrays, int *order);
/* the Legendre-Gauss order that conversion chooses for redshifts in [zmin, zmax]
 * (src/cnvt_coord.c:356-396 on 128 samples, :277-278); host only; -1 if none of 4..32
 * converges to cosmo->ecdst */
int psb_cnvt_order(const psb_cosmo *cosmo, double zmin, double zmax);

/* One in-place forward pass (sign -1, unnormalised: the convention of the FFTW
 * r2c plan of src/genr_mesh.c:738-743 along one axis) of the hand-written
 * strided FFT over caller-owned device memory: complex double (precision 8) or
 * float (4), rows of ngk elements; axis 1: along y of (outer_n, ng, ngk);
 * axis 0: along x of (ng, outer_n, ngk); axis 2: real-to-complex along z of
 * outer_n contiguous rows of 2 ngk reals (ng used, ngk = ng/2 + 1), in place;
 * axis 3: the z and y passes of outer_n planes (outer_n, ng, 2 ngk) of reals in
 * one persistent kernel (planes handed from the z to the y pass through the L2).
 * ng in {512, 1024, 1536, 2048}. */
int psb_fft_axis(psb_context *ctx, void *data_dev, int precision, int ng, int ngk, int axis,
    int outer_n);

/* Device-side synthetic catalogue generator for benchmarks:
 * fills n x {x,y,z,w} on the device; kind 0 = uniform in [0,L)^3, 1 = clustered.
 * Returns a device pointer to be released with psb_device_free. */
double *psb_generate_catalog(psb_context *ctx, size_t n, double boxsize, int kind,
    uint64_t seed);
/* same, into caller-owned device memory; particle j of the call is particle
 * first_index + j of the (seed, kind) catalogue, so a catalogue can be produced
 * in chunks and by several ranks */
int psb_generate_into(psb_context *ctx, double *dst_dev, size_t n, double boxsize, int kind,
    uint64_t seed, uint64_t first_index);
void psb_device_free(psb_context *ctx, void *ptr);
/* copy device catalogue to host (tests) */
int psb_copy_to_host(psb_context *ctx, void *dst, const void *src_dev, size_t bytes);
/* test hook, no GPU needed: the staging copy that carries pageable host catalogues (the
 * reference's malloc'd DATA arrays, src/read_cata.c:86-189) into the pinned upload buffers —
 * a persistent pool of `nthreads` host threads (csrc/hostcopy.cpp; plain or, with option
 * "h2d_nt", non-temporal stores).
 * Copies src to dst `repeats` times; returns the number of threads the pool was asked for */
int psb_test_host_copy(void *dst, const void *src, size_t bytes, int nthreads, int repeats);
/* the same, as the upload path uses it: `bytes` of src streamed through two alternating staging
 * buffers of `piece` bytes by ONE pool; returns the seconds taken (tools/staging_bench.py) */
double psb_test_host_stage(const void *src, size_t bytes, size_t piece, int nthreads, int stream_stores);

#ifdef __cplusplus
}
#endif
#endif
