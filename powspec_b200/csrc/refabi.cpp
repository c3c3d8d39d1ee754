// The reference's seam: genr_mesh / mesh_destroy / powspec / powspec_destroy /
// powspec_assign_names with the reference's signatures and struct layouts
// (include/powspec_refabi.h), implemented on the psb_* device path.  This is
// what the reference's unchanged C host (powspec.c, load_conf.c, read_cata.c,
// cnvt_coord.c, save_res.c) links against instead of genr_mesh.o + multipole.o.
//
// Behaviour kept from the reference: progress text on stdout
// ("Generating meshes for FFT ..." / FMT_DONE, src/genr_mesh.c:875,924;
// "Evaluating power spectra ...", src/multipole.c:1180,1276), NULL + P_ERR-style
// message on failure, genr_mesh takes ownership of (frees) the particle arrays
// (src/genr_mesh.c:917-922) and writes cat->shot/norm for simulation boxes
// (:904-909), powspec leaves MESH metadata alive for save_res, PK arrays are
// plain malloc memory laid out as powspec_init does (src/multipole.c:306-421).
//
// Runtime knobs the reference fixes at compile time:
//   POWSPEC_B200_PRECISION = 8 | 4   (the reference's -DSINGLE_PREC, Makefile:14)
//   POWSPEC_B200_DEVICE    = CUDA device ordinal (default 0)
//   POWSPEC_B200_DEVICES   = "0,1,2,3" / "0-7": ONE mesh x-slab-decomposed over these
//                            devices (psb_group: one host thread per device inside the
//                            library, peer copies / peer stores, no NCCL) — how this
//                            single-process C host runs meshes that exceed one GPU
//                            (BASELINE configs 4 and 5).  Simulation boxes; a survey or a
//                            pending coordinate conversion falls back to the first device.
//   POWSPEC_B200_TIMING    = path: write a JSON line with the device stage timings
//                            (CUDA events) of the run; never touches the output file

#include "../../include/powspec_b200.h"
#include "../../include/powspec_refabi.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define FMT_DONE "\r\x1B[70C[\x1B[32;1mDONE\x1B[0m]\n"  /* src/define.h:104 */
#define P_ERR(...) fprintf(stderr, "\n\x1B[31;1mError:\x1B[0m " __VA_ARGS__)

static psb_context *g_ctx = nullptr;
static psb_group *g_group = nullptr;      // POWSPEC_B200_DEVICES: the slab-decomposed path

// "0,1,2" and ranges "0-3" (also mixed); a device may be listed more than once
static int parse_devices(const char *s, int *dev, int maxn) {
  int n = 0;
  while (s && *s && n < maxn) {
    char *end = nullptr;
    const long a = strtol(s, &end, 10);
    if (end == s || a < 0) return -1;
    long b = a;
    s = end;
    if (*s == '-') {
      b = strtol(s + 1, &end, 10);
      if (end == s + 1 || b < a) return -1;
      s = end;
    }
    for (long d = a; d <= b && n < maxn; d++) dev[n++] = (int) d;
    if (*s == ',') s++;
    else if (*s) return -1;
  }
  return n;
}

// Coordinate conversion requested through cnvt_coord() below: carried out on the
// device by the next genr_mesh(), on the records it uploads anyway.
static struct {
  bool pending = false;
  psb_cosmo cosmo;
  int dcnvt[2] = {0, 0}, rcnvt[2] = {0, 0};
  double *z = nullptr, *d = nullptr;
} g_cnvt;

// first two columns of an ASCII file, '#' comments and blank lines skipped
// (what read_ascii_simple does for the reference, io/read_ascii.c:392-470)
static int read_two_columns(const char *fname, double **x, double **y, size_t *num) {
  FILE *fp = fopen(fname, "r");
  if (!fp) { P_ERR("cannot open file for reading: `%s'\n", fname); return -1; }
  size_t cap = 1024, n = 0;
  double *a = static_cast<double *>(malloc(cap * sizeof(double)));
  double *b = static_cast<double *>(malloc(cap * sizeof(double)));
  char *line = nullptr;
  size_t len = 0;
  int rc = (a && b) ? 0 : -1;
  while (!rc && getline(&line, &len, fp) != -1) {
    const char *p = line;
    while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\v' || *p == '\f') ++p;
    if (*p == '#' || *p == '\0') continue;
    if (n == cap) {
      cap *= 2;
      double *na = static_cast<double *>(realloc(a, cap * sizeof(double)));
      double *nb = static_cast<double *>(realloc(b, cap * sizeof(double)));
      if (na) a = na;
      if (nb) b = nb;
      if (!na || !nb) { rc = -1; break; }
    }
    if (sscanf(p, "%lf %lf", a + n, b + n) != 2) { P_ERR("failed to read line: %s\n", p); rc = -1; break; }
    n++;
  }
  free(line);
  fclose(fp);
  if (rc) { free(a); free(b); return rc; }
  *x = a; *y = b; *num = n;
  return 0;
}

static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static void fill_params(const psb_ref_CONF *conf, psb_params *p) {
  memset(p, 0, sizeof *p);
  p->ncat = conf->ndata;
  p->issim = conf->issim;
  p->intlace = conf->intlace;
  p->assign = conf->assign;
  p->gsize = conf->gsize;
  p->logscale = conf->logscale;
  p->verbose = conf->verbose;
  p->npole = conf->npole;
  for (int i = 0; i < conf->npole && i < PSB_MAX_POLES; i++) p->poles[i] = conf->poles[i];
  p->has_bsize = conf->bsize != nullptr;
  p->isauto[0] = conf->isauto[0];
  p->isauto[1] = conf->isauto[1];
  p->iscross = conf->iscross;
  for (int a = 0; a < 3; a++) {
    p->los[a] = conf->los ? conf->los[a] : (a == 2 ? 1.0 : 0.0);
    p->bsize[a] = conf->bsize ? conf->bsize[a] : 0.0;
    p->bpad[a] = conf->bpad ? conf->bpad[a] : 0.02;     /* DEFAULT_BOX_PAD */
  }
  p->kmin = conf->kmin;
  p->kmax = conf->kmax;
  p->kbin = conf->kbin;
  p->precision = env_int("POWSPEC_B200_PRECISION", 8);
  p->device = env_int("POWSPEC_B200_DEVICE", 0);
}

extern "C" {

const char *powspec_assign_names[] = {"NGP", "CIC", "TSC", "PCS"};

psb_ref_MESH *genr_mesh(const psb_ref_CONF *conf, psb_ref_CATA *cat) {
  printf("Generating meshes for FFT ...");
  if (!conf) { P_ERR("configuration parameters not loaded\n"); return nullptr; }
  if (conf->verbose) printf("\n");
  fflush(stdout);
  if (!cat) { P_ERR("catalogs not read\n"); return nullptr; }

  psb_params par;
  fill_params(conf, &par);
  int devs[16], ndev = 0;
  if (const char *list = getenv("POWSPEC_B200_DEVICES")) {
    ndev = parse_devices(list, devs, 16);
    if (ndev < 0) { P_ERR("invalid POWSPEC_B200_DEVICES: `%s'\n", list); return nullptr; }
    if (ndev >= 1) par.device = devs[0];
  }
  const bool slabs = ndev >= 2 && conf->issim && !g_cnvt.pending;
  if (ndev >= 2 && !slabs && conf->verbose)
    printf("  POWSPEC_B200_DEVICES: surveys run on one device (%d)\n", par.device);
  if (slabs) {
    if (!g_group) g_group = psb_group_create(devs, ndev);
    if (!g_group) return nullptr;
  }
  else {
    if (!g_ctx) g_ctx = psb_create(par.device);
    if (!g_ctx) return nullptr;
  }

  psb_cats in;
  memset(&in, 0, sizeof in);
  in.memspace = PSB_MEM_HOST;
  for (int i = 0; i < cat->num && i < 2; i++) {
    in.data[i] = cat->data ? reinterpret_cast<const double *>(cat->data[i]) : nullptr;
    in.rand[i] = cat->rand ? reinterpret_cast<const double *>(cat->rand[i]) : nullptr;
    in.ndata[i] = cat->ndata[i];
    in.nrand[i] = cat->nrand[i];
    in.wdata[i] = cat->wdata[i];
    in.wrand[i] = cat->wrand[i];
    in.alpha[i] = cat->alpha[i];
    in.shot[i] = cat->shot[i];
    in.norm[i] = cat->norm[i];
  }
  if (g_cnvt.pending) {
    in.cnvt = &g_cnvt.cosmo;
    for (int i = 0; i < 2; i++) { in.dcnvt[i] = g_cnvt.dcnvt[i]; in.rcnvt[i] = g_cnvt.rcnvt[i]; }
  }
  const int mesh_rc = slabs ? psb_group_mesh(g_group, &par, &in) : psb_mesh(g_ctx, &par, &in);
  if (g_cnvt.pending) {
    free(g_cnvt.z); free(g_cnvt.d);
    g_cnvt.z = g_cnvt.d = nullptr;
    g_cnvt.pending = false;
  }
  if (mesh_rc) return nullptr;

  psb_ref_MESH *mesh = static_cast<psb_ref_MESH *>(calloc(1, sizeof *mesh));
  if (!mesh) { P_ERR("failed to initalise the meshes\n"); return nullptr; }
  mesh->num = conf->ndata;
  mesh->Ng = conf->gsize;
  mesh->Ngk = (mesh->Ng >> 1) + 1;
  mesh->Ntot = (size_t) mesh->Ng * mesh->Ng * mesh->Ng;
  mesh->Ncmplx = (size_t) mesh->Ng * mesh->Ng * mesh->Ngk;
  mesh->issim = conf->issim;
  mesh->intlace = conf->intlace;
  mesh->assign = conf->assign;
  mesh->fft_init = false;
  /* box metadata as def_box leaves it, read by save_res (src/save_res.c:70-82) */
  if (slabs) {                  /* def_box for simulation boxes, src/genr_mesh.c:516-522 */
    for (int a = 0; a < 3; a++) { mesh->min[a] = 0; mesh->bsize[a] = mesh->max[a] = conf->bsize[a]; }
  }
  else psb_mesh_box(g_ctx, mesh->min, mesh->bsize, mesh->max);
  if (conf->issim) {            /* src/genr_mesh.c:904-909 */
    for (int i = 0; i < cat->num; i++) {
      const double vol = mesh->bsize[0] * mesh->bsize[1] * mesh->bsize[2];
      cat->shot[i] = vol / cat->wdata[i];
      cat->norm[i] = cat->wdata[i] * cat->wdata[i] / vol;
    }
  }
  for (int a = 0; a < 3; a++) mesh->smin[a] = mesh->min[a] * mesh->Ng / mesh->bsize[a];

  /* ownership of the particles ends here, as in the reference */
  for (int i = 0; i < cat->num; i++) {
    if (cat->data && cat->data[i]) free(cat->data[i]);
    if (cat->rand && cat->rand[i]) free(cat->rand[i]);
  }
  free(cat->data); cat->data = nullptr;
  free(cat->rand); cat->rand = nullptr;

  printf(FMT_DONE);
  return mesh;
}

// cnvt_coord(), src/cnvt_coord.c:549-582.  Optional part of the seam: when the
// host is linked WITHOUT its own cnvt_coord.o this one is used.  It validates
// the request, reads the distance table if one is named, and leaves the
// conversion itself to genr_mesh(), which performs it on the device on the
// records it uploads (the host arrays keep RA / Dec / z; genr_mesh frees them).
// What can FAIL is checked here, on the host, so that the reference's error contract
// holds (message + POWSPEC_ERR_CNVT from cnvt_coord(), src/cnvt_coord.c:549-582, before any
// "DONE" is printed): negative / invalid redshifts (:185-268), the convergence test of
// the Legendre-Gauss order on the catalogues' redshift range (:356-396, :495-511), a table
// that cannot be interpolated (:453-456).
int cnvt_coord(const psb_ref_CONF *conf, psb_ref_CATA *cat) {
  if (!conf) { P_ERR("configuration parameters not loaded\n"); return -10; /* POWSPEC_ERR_CONF */ }
  if (!conf->cnvt) return 0;
  printf("Converting coordinates ...");
  if (conf->verbose) printf("\n");
  fflush(stdout);
  if (!cat) { P_ERR("catalogs not read\n"); return -11; /* POWSPEC_ERR_CATA */ }
  memset(&g_cnvt.cosmo, 0, sizeof g_cnvt.cosmo);
  g_cnvt.cosmo.omega_m = conf->omega_m;
  g_cnvt.cosmo.omega_l = conf->omega_l;
  g_cnvt.cosmo.omega_k = conf->omega_k;
  g_cnvt.cosmo.eos_w = conf->eos_w;
  g_cnvt.cosmo.ecdst = conf->ecdst;
  if (conf->fcdst) {
    size_t n = 0;
    if (conf->verbose) printf("\n  Reading samples from file: %s\n", conf->fcdst);
    if (read_two_columns(conf->fcdst, &g_cnvt.z, &g_cnvt.d, &n)) return -12; /* POWSPEC_ERR_CNVT */
    g_cnvt.cosmo.sample_z = g_cnvt.z;
    g_cnvt.cosmo.sample_d = g_cnvt.d;
    g_cnvt.cosmo.nsample = n;
  }
  for (int i = 0; i < cat->num && i < 2; i++) {
    g_cnvt.dcnvt[i] = conf->dcnvt ? conf->dcnvt[i] : 0;     /* DEFAULT_CONVERT = false */
    g_cnvt.rcnvt[i] = conf->rcnvt ? conf->rcnvt[i] : 0;
  }
  auto drop_table = [&]() { free(g_cnvt.z); free(g_cnvt.d); g_cnvt.z = g_cnvt.d = nullptr; };
  if (conf->fcdst) {
    if (g_cnvt.cosmo.nsample < 2) {
      P_ERR("failed to interpolate the sample points\n");
      drop_table();
      return -12;
    }
  }
  else {
    // redshift range of everything that will be converted (cnvt_z_sample)
    double zmin = 1.7976931348623157e308, zmax = -1.7976931348623157e308;
    for (int i = 0; i < cat->num && i < 2; i++)
      for (int r = 0; r < 2; r++) {
        const psb_ref_DATA *p = r ? (cat->rand ? cat->rand[i] : nullptr) : (cat->data ? cat->data[i] : nullptr);
        const size_t n = r ? (cat->nrand ? cat->nrand[i] : 0) : cat->ndata[i];
        if (!(r ? g_cnvt.rcnvt[i] : g_cnvt.dcnvt[i]) || !p) continue;
        for (size_t k = 0; k < n; k++) {
          const double z = p[k].x[2];
          if (z < 0) {
            P_ERR("invalid negative redshift in the %s catalog:\n(%g, %g, %g)\n", r ? "random" : "data",
                p[k].x[0], p[k].x[1], z);
            return -12;
          }
          if (zmax < z) zmax = z;
          if (zmin > z) zmin = z;
        }
      }
    if (zmin > zmax) { P_ERR("invalid redshift value in the catalogs\n"); return -12; }
    if (psb_cnvt_order(&g_cnvt.cosmo, zmin, zmax) < 0) {
      P_ERR("failed to perform the convergency test for integrations\n");
      return -12;
    }
  }
  g_cnvt.pending = true;
  if (conf->verbose) printf("  Conversion scheduled on the device (with the mesh generation)\n");
  printf(FMT_DONE);
  return 0;
}

void mesh_destroy(psb_ref_MESH *mesh) {
  /* device buffers and cuFFT plans live in the context */
  if (g_ctx) { psb_destroy(g_ctx); g_ctx = nullptr; }
  if (g_group) { psb_group_destroy(g_group); g_group = nullptr; }
  free(mesh);
}

psb_ref_PK *powspec(const psb_ref_CONF *conf, const psb_ref_CATA *cat, psb_ref_MESH *mesh) {
  printf("Evaluating power spectra ...");
  if (!conf) { P_ERR("configuration parameters not loaded\n"); return nullptr; }
  if (conf->verbose) printf("\n");
  fflush(stdout);
  if (!cat) { P_ERR("catalogs not read\n"); return nullptr; }
  if (!mesh || !(g_ctx || g_group)) { P_ERR("meshes not generated\n"); return nullptr; }

  psb_params par;
  fill_params(conf, &par);
  psb_result *res = g_group ? psb_group_power(g_group, &par) : psb_power(g_ctx, &par);
  if (!res) return nullptr;

  psb_ref_PK *pk = static_cast<psb_ref_PK *>(calloc(1, sizeof *pk));
  const int nl = psb_result_nl(res), nb = psb_result_nbin(res);
  bool ok = pk != nullptr;
  if (ok) {
    pk->issim = conf->issim; pk->log = conf->logscale;
    if (pk->issim) for (int a = 0; a < 3; a++) pk->los[a] = conf->los[a];
    pk->isauto[0] = conf->isauto[0]; pk->isauto[1] = conf->isauto[1];
    pk->iscross = conf->iscross;
    pk->nl = nl; pk->nbin = nb; pk->dk = conf->kbin; pk->nomp = 1;
    pk->poles = static_cast<int *>(malloc(nl * sizeof(int)));
    pk->kedge = static_cast<double *>(malloc((nb + 1) * sizeof(double)));
    pk->k = static_cast<double *>(malloc(nb * sizeof(double)));
    pk->km = static_cast<double *>(malloc(nb * sizeof(double)));
    pk->cnt = static_cast<size_t *>(malloc(nb * sizeof(size_t)));
    ok = pk->poles && pk->kedge && pk->k && pk->km && pk->cnt;
    if (ok && pk->issim) ok = (pk->lcnt = static_cast<double *>(malloc(sizeof(double) * nl * nb))) != nullptr;
  }
  double *flat = ok ? static_cast<double *>(malloc(sizeof(double) * nl * nb)) : nullptr;
  ok = ok && flat;
  if (ok) {
    memcpy(pk->poles, conf->poles, nl * sizeof(int));
    psb_result_get(res, PSB_GET_KEDGE, 0, pk->kedge);
    psb_result_get(res, PSB_GET_K, 0, pk->k);
    psb_result_get(res, PSB_GET_KM, 0, pk->km);
    psb_result_get(res, PSB_GET_CNT, 0, pk->cnt);      /* size_t == uint64 on LP64 */
    if (pk->issim) psb_result_get(res, PSB_GET_LCNT, 0, pk->lcnt);
    for (int c = 0; c <= 2 && ok; c++) {
      const bool cross = c == 2;
      if (cross ? !conf->iscross : !conf->isauto[c]) continue;
      double **rows = static_cast<double **>(calloc(nl, sizeof(double *)));
      if (!rows) { ok = false; break; }
      if (cross) pk->xpl = rows; else pk->pl[c] = rows;
      const long got = cross ? psb_result_get(res, PSB_GET_XPL, 0, flat)
                             : psb_result_get(res, PSB_GET_PL, c, flat);
      for (int l = 0; l < nl; l++) {
        rows[l] = static_cast<double *>(calloc(nb, sizeof(double)));
        if (!rows[l]) { ok = false; break; }
        if (got > 0) memcpy(rows[l], flat + (size_t) l * nb, nb * sizeof(double));
      }
    }
  }
  free(flat);
  psb_result_free(res);
  if (const char *tpath = getenv("POWSPEC_B200_TIMING")) {
    static const char *names[] = {"h2d", "bounds", "sort", "memset", "assign", "fft", "geom", "bin", "ylm",
      "fft_strided", "cnvt"};
    double ms[PSB_T_COUNT];
    if (*tpath && g_group) {
      // slab-decomposed run: the stage times of rank 0 (CUDA events), ms
      static const char *dn[] = {"route", "assign", "halo", "fft_zy", "transpose", "fft_x", "bin", "reduce"};
      double dm[PSB_D_COUNT];
      if (psb_dist_timings(psb_group_rank(g_group, 0), dm, PSB_D_COUNT) >= 0) {
        if (FILE *f = fopen(tpath, "a")) {
          fprintf(f, "{\"grid\": %d, \"devices\": %d, \"stages_ms\": {", conf->gsize, psb_group_size(g_group));
          for (int i = 0; i < PSB_D_COUNT; i++) fprintf(f, "%s\"%s\": %.4f", i ? ", " : "", dn[i], dm[i]);
          fprintf(f, "}}\n");
          fclose(f);
        }
      }
    }
    else if (*tpath && psb_timings(g_ctx, ms, PSB_T_COUNT) > 0) {
      if (FILE *f = fopen(tpath, "a")) {
        fprintf(f, "{\"grid\": %d, \"launches\": %ld, \"stages_ms\": {", conf->gsize, psb_launch_count(g_ctx));
        for (int i = 0; i < 11; i++) fprintf(f, "%s\"%s\": %.4f", i ? ", " : "", names[i], ms[i]);
        fprintf(f, "}}\n");
        fclose(f);
      }
    }
  }
  if (!ok) {
    P_ERR("failed to initialise the power spectra\n");
    powspec_destroy(pk);
    return nullptr;
  }
  printf(FMT_DONE);
  return pk;
}

void powspec_destroy(psb_ref_PK *pk) {
  if (!pk) return;
  free(pk->poles); free(pk->kedge); free(pk->k); free(pk->km); free(pk->cnt); free(pk->lcnt);
  for (int i = 0; i < 2; i++)
    if (pk->pl[i]) {
      for (int j = 0; j < pk->nl; j++) free(pk->pl[i][j]);
      free(pk->pl[i]);
    }
  if (pk->xpl) {
    for (int j = 0; j < pk->nl; j++) free(pk->xpl[j]);
    free(pk->xpl);
  }
  free(pk->pcnt); free(pk->plcnt);
  free(pk);
}

}  // extern "C"
