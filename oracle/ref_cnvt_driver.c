/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 *
 * In-memory driver of the UNMODIFIED reference cnvt_coord() (src/cnvt_coord.c,
 * compiled where it lies under /root/reference together with math/legauss.c,
 * math/cspline.c and the io/ + lib/ files read_ascii_simple needs): fills the
 * members of CONF / CATA that cnvt_coord reads (src/cnvt_coord.c:440-582) and
 * converts the caller's {RA, Dec, z, w} records in place.
 */
#define _GNU_SOURCE
#include "load_conf.h"
#include "read_cata.h"
#include "cnvt_coord.h"
#include <stdlib.h>
#include <string.h>

/* fcdst: name of a (z, distance) table or NULL for Legendre-Gauss integration.
 * data[i] / rand[i]: n x 4 doubles, converted in place when dcnvt[i] / rcnvt[i]. */
int oracle_cnvt(double omega_m, double omega_l, double omega_k, double eos_w, double ecdst,
    const char *fcdst, int ncat, double **data, const size_t *ndata, double **rand,
    const size_t *nrand, const int *dcnvt, const int *rcnvt) {
  CONF conf;
  CATA cat;
  memset(&conf, 0, sizeof conf);
  memset(&cat, 0, sizeof cat);
  bool dc[2] = {false, false}, rc[2] = {false, false};
  DATA *dp[2] = {NULL, NULL}, *rp[2] = {NULL, NULL};
  size_t nd[2] = {0, 0}, nr[2] = {0, 0};
  for (int i = 0; i < ncat && i < 2; i++) {
    dc[i] = dcnvt[i]; rc[i] = rcnvt[i];
    dp[i] = (DATA *) data[i]; rp[i] = rand ? (DATA *) rand[i] : NULL;
    nd[i] = ndata[i]; nr[i] = nrand ? nrand[i] : 0;
  }
  conf.ndata = ncat;
  conf.cnvt = true;
  conf.dcnvt = dc; conf.rcnvt = rc;
  conf.omega_m = omega_m; conf.omega_l = omega_l; conf.omega_k = omega_k;
  conf.eos_w = eos_w; conf.ecdst = ecdst;
  conf.fcdst = (char *) fcdst;
  conf.verbose = false;
  cat.num = ncat;
  cat.data = dp; cat.rand = rp; cat.ndata = nd; cat.nrand = nr;
  return cnvt_coord(&conf, &cat);
}
