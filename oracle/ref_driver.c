/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 *
 * In-memory driver around the UNMODIFIED reference hot path.  It is compiled
 * (only where /root/reference exists) together with the reference's own
 * src/genr_mesh.c and src/multipole.c into oracle/_ref/libpowspec_ref.so, and
 * does what the reference's main() does between read_cata and save_res
 * (src/powspec.c:47-55): fill CONF (only the fields the two stages read,
 * SURVEY.md §8b) and CATA from caller arrays, call genr_mesh() and powspec(),
 * and hand the PK contents back as flat arrays.  No ASCII parsing is involved,
 * so it doubles as the timed CPU baseline ("kind": "reference").
 *
 * Struct definitions come from the reference headers via the include path; no
 * reference source is copied here.
 */
#define _GNU_SOURCE
#include "load_conf.h"
#include "read_cata.h"
#include "genr_mesh.h"
#include "multipole.h"
#include "oracle_abi.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef OMP
#include <omp.h>
#endif

struct oracle_result_s {
  int nbin, nl, ncat, issim;
  int poles[8];
  double *k, *kedge, *km, *lcnt;
  unsigned long long *cnt;
  double *pl[2], *xpl;
  double shot[2], norm[2];
  double bmin[3], bsize[3];
  double t_mesh, t_pk;
  /* optional copies of the real-space meshes right after genr_mesh */
  size_t ntot;
  int fft_real_size;
  void *Fr[2], *Frl[2];
};

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

const char *oracle_backend(void) {
#ifdef SINGLE_PREC
  return "reference (unmodified genr_mesh.c+multipole.c, -DSINGLE_PREC) + fftcpu-shim";
#else
  return "reference (unmodified genr_mesh.c+multipole.c) + fftcpu-shim";
#endif
}

int oracle_real_size(void) { return (int) sizeof(FFT_REAL); }

/* same tolerance of NULL members as cata_destroy (src/read_cata.c:199-212) */
static void free_cata(CATA *cat) {
  if (!cat) return;
  for (int i = 0; i < cat->num; i++) {
    if (cat->data && cat->data[i]) free(cat->data[i]);
    if (cat->rand && cat->rand[i]) free(cat->rand[i]);
  }
  free(cat->data); free(cat->rand); free(cat->ndata); free(cat->nrand);
  free(cat->wdata); free(cat->wrand); free(cat->alpha); free(cat->shot);
  free(cat->norm); free(cat);
}

static DATA *copy_particles(const double *xyzw, size_t n) {
  DATA *d = malloc((n ? n : 1) * sizeof(DATA));
  if (!d) return NULL;
  if (n) memcpy(d, xyzw, n * sizeof(DATA));     /* DATA is {x[3], w}: 4 doubles */
  return d;
}

oracle_result *oracle_run(const oracle_params *par, const oracle_cats *in,
    int keep_mesh) {
  if (sizeof(DATA) != 4 * sizeof(double)) return NULL;
  CONF conf;
  memset(&conf, 0, sizeof conf);
  double los[3] = {par->los[0], par->los[1], par->los[2]};
  double bsize[3] = {par->bsize[0], par->bsize[1], par->bsize[2]};
  double bpad[3] = {par->bpad[0], par->bpad[1], par->bpad[2]};
  int poles[8];
  for (int i = 0; i < par->npole; i++) poles[i] = par->poles[i];
  conf.ndata = par->ncat;
  conf.issim = par->issim;
  conf.los = los;
  conf.bsize = par->has_bsize ? bsize : NULL;
  conf.bpad = bpad;
  conf.gsize = par->gsize;
  conf.assign = par->assign;
  conf.intlace = par->intlace;
  conf.poles = poles;
  conf.npole = par->npole;
  conf.kmin = par->kmin;
  conf.kmax = par->kmax;
  conf.logscale = par->logscale;
  conf.kbin = par->kbin;
  conf.isauto[0] = par->isauto[0];
  conf.isauto[1] = par->isauto[1];
  conf.iscross = par->iscross;
  conf.verbose = par->verbose;

  /* CATA laid out as cata_init (src/read_cata.c:34-72) does, with plain malloc
     because genr_mesh free()s the particle arrays (src/genr_mesh.c:917-922). */
  CATA *cat = calloc(1, sizeof *cat);
  if (!cat) return NULL;
  const int nc = par->ncat;
  cat->num = nc;
  cat->data = calloc(nc, sizeof(DATA *));
  cat->rand = calloc(nc, sizeof(DATA *));
  cat->ndata = calloc(nc, sizeof(size_t));
  cat->nrand = calloc(nc, sizeof(size_t));
  cat->wdata = calloc(nc, sizeof(double));
  cat->wrand = calloc(nc, sizeof(double));
  cat->alpha = calloc(nc, sizeof(double));
  cat->shot = calloc(nc, sizeof(double));
  cat->norm = calloc(nc, sizeof(double));
  for (int i = 0; i < nc; i++) {
    cat->data[i] = copy_particles(in->data[i], in->ndata[i]);
    cat->ndata[i] = in->ndata[i];
    cat->wdata[i] = in->wdata[i];
    if (!par->issim) {
      cat->rand[i] = copy_particles(in->rand[i], in->nrand[i]);
      cat->nrand[i] = in->nrand[i];
      cat->wrand[i] = in->wrand[i];
      cat->alpha[i] = in->alpha[i];
      cat->shot[i] = in->shot[i];
      cat->norm[i] = in->norm[i];
    }
  }

  oracle_result *res = calloc(1, sizeof *res);
  if (!res) { free_cata(cat); return NULL; }

  /* the reference prints progress on stdout; keep it out of test logs */
  FILE *saved = NULL;
  int quiet = !par->verbose;
  if (quiet) { fflush(stdout); saved = stdout; stdout = fopen("/dev/null", "w"); }

  double t0 = now_s();
  MESH *mesh = genr_mesh(&conf, cat);
  double t1 = now_s();
  const double t_mesh = t1 - t0;
  PK *pk = NULL;
  if (mesh) {
    if (keep_mesh) {
      res->ntot = mesh->Ntot;
      res->fft_real_size = (int) sizeof(FFT_REAL);
      for (int i = 0; i < nc; i++) {
        res->Fr[i] = malloc(mesh->Ntot * sizeof(FFT_REAL));
        if (res->Fr[i]) memcpy(res->Fr[i], mesh->Fr[i], mesh->Ntot * sizeof(FFT_REAL));
        if (mesh->Frl && mesh->intlace) {
          res->Frl[i] = malloc(mesh->Ntot * sizeof(FFT_REAL));
          if (res->Frl[i])
            memcpy(res->Frl[i], mesh->Frl[i], mesh->Ntot * sizeof(FFT_REAL));
        }
      }
    }
    t1 = now_s();
    pk = powspec(&conf, cat, mesh);
  }
  double t2 = now_s();
  if (quiet) { fclose(stdout); stdout = saved; }

  if (!mesh || !pk) {
    if (mesh) mesh_destroy(mesh);
    free_cata(cat);
    oracle_free(res);
    return NULL;
  }
  res->t_mesh = t_mesh;
  res->t_pk = t2 - t1;
  res->nbin = pk->nbin;
  res->nl = pk->nl;
  res->ncat = nc;
  res->issim = par->issim;
  for (int i = 0; i < pk->nl; i++) res->poles[i] = pk->poles[i];
  const size_t nb = pk->nbin, nlb = (size_t) pk->nl * pk->nbin;
  res->k = malloc(nb * sizeof(double));
  res->km = malloc(nb * sizeof(double));
  res->kedge = malloc((nb + 1) * sizeof(double));
  res->cnt = malloc(nb * sizeof(unsigned long long));
  res->lcnt = calloc(nlb, sizeof(double));
  memcpy(res->k, pk->k, nb * sizeof(double));
  memcpy(res->km, pk->km, nb * sizeof(double));
  memcpy(res->kedge, pk->kedge, (nb + 1) * sizeof(double));
  for (size_t i = 0; i < nb; i++) res->cnt[i] = pk->cnt[i];
  if (pk->lcnt) memcpy(res->lcnt, pk->lcnt, nlb * sizeof(double));
  for (int c = 0; c < 2; c++) {
    if (c < nc && pk->pl[c]) {
      res->pl[c] = malloc(nlb * sizeof(double));
      for (int l = 0; l < pk->nl; l++)
        memcpy(res->pl[c] + (size_t) l * nb, pk->pl[c][l], nb * sizeof(double));
    }
  }
  if (pk->xpl) {
    res->xpl = malloc(nlb * sizeof(double));
    for (int l = 0; l < pk->nl; l++)
      memcpy(res->xpl + (size_t) l * nb, pk->xpl[l], nb * sizeof(double));
  }
  for (int i = 0; i < nc; i++) {
    res->shot[i] = cat->shot[i];
    res->norm[i] = cat->norm[i];
  }
  for (int i = 0; i < 3; i++) {
    res->bmin[i] = mesh->min[i];
    res->bsize[i] = mesh->bsize[i];
  }
  mesh_destroy(mesh);
  powspec_destroy(pk);
  free_cata(cat);
  return res;
}

void oracle_free(oracle_result *r) {
  if (!r) return;
  free(r->k); free(r->kedge); free(r->km); free(r->lcnt); free(r->cnt);
  free(r->pl[0]); free(r->pl[1]); free(r->xpl);
  for (int i = 0; i < 2; i++) { free(r->Fr[i]); free(r->Frl[i]); }
  free(r);
}

int oracle_nbin(const oracle_result *r) { return r->nbin; }
int oracle_nl(const oracle_result *r) { return r->nl; }
size_t oracle_ntot(const oracle_result *r) { return r->ntot; }
double oracle_time(const oracle_result *r, int which) {
  return which == 0 ? r->t_mesh : r->t_pk;
}

/* what: see ORACLE_GET_* in oracle_abi.h.  Returns the number of elements
   copied into dst (doubles unless stated), or -1 when absent. */
long oracle_get(const oracle_result *r, int what, int idx, void *dst) {
  const size_t nb = r->nbin, nlb = (size_t) r->nl * r->nbin;
  switch (what) {
    case ORACLE_GET_K: memcpy(dst, r->k, nb * 8); return (long) nb;
    case ORACLE_GET_KEDGE: memcpy(dst, r->kedge, (nb + 1) * 8); return (long) nb + 1;
    case ORACLE_GET_KM: memcpy(dst, r->km, nb * 8); return (long) nb;
    case ORACLE_GET_CNT: memcpy(dst, r->cnt, nb * 8); return (long) nb;
    case ORACLE_GET_LCNT: memcpy(dst, r->lcnt, nlb * 8); return (long) nlb;
    case ORACLE_GET_PL:
      if (idx < 0 || idx > 1 || !r->pl[idx]) return -1;
      memcpy(dst, r->pl[idx], nlb * 8); return (long) nlb;
    case ORACLE_GET_XPL:
      if (!r->xpl) return -1;
      memcpy(dst, r->xpl, nlb * 8); return (long) nlb;
    case ORACLE_GET_SHOT: memcpy(dst, r->shot, 16); return 2;
    case ORACLE_GET_NORM: memcpy(dst, r->norm, 16); return 2;
    case ORACLE_GET_BMIN: memcpy(dst, r->bmin, 24); return 3;
    case ORACLE_GET_BSIZE: memcpy(dst, r->bsize, 24); return 3;
    case ORACLE_GET_FR:         /* FFT_REAL elements */
      if (idx < 0 || idx > 1 || !r->Fr[idx]) return -1;
      memcpy(dst, r->Fr[idx], r->ntot * r->fft_real_size); return (long) r->ntot;
    case ORACLE_GET_FRL:
      if (idx < 0 || idx > 1 || !r->Frl[idx]) return -1;
      memcpy(dst, r->Frl[idx], r->ntot * r->fft_real_size); return (long) r->ntot;
    default: return -1;
  }
}
