/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 *
 * Plain-C ABI shared by the two oracle libraries:
 *   oracle/_ref/libpowspec_ref.so   unmodified reference + ref_driver.c
 *   oracle/libpowspec_port.so       CPU restatement (oracle/pspec_port.c)
 * Both export the same entry points so that tests and bench.py's cpu_baseline
 * leg can swap them.  Field meanings follow the reference's CONF / CATA
 * members (src/load_conf.h:40-93, src/read_cata.h:42-59).
 */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int ncat;             /* CONF.ndata: 1 or 2 catalogues                      */
  int issim;            /* CUBIC_SIM                                          */
  int intlace;          /* GRID_INTERLACE                                     */
  int assign;           /* PARTICLE_ASSIGN: 0 NGP, 1 CIC, 2 TSC, 3 PCS        */
  int gsize;            /* GRID_SIZE                                          */
  int logscale;         /* LOG_SCALE (kmin/kmax/kbin then in log10)           */
  int verbose;
  int npole;            /* number of multipoles                               */
  int poles[8];         /* sorted, unique, <= 6                               */
  int has_bsize;        /* BOX_SIZE given (mandatory for sims)                */
  int isauto[2];
  int iscross;
  double los[3];        /* LINE_OF_SIGHT (sims)                               */
  double bsize[3];
  double bpad[3];       /* BOX_PAD (surveys without BOX_SIZE)                 */
  double kmin, kmax, kbin;      /* kmax <= 0 means unset                      */
} oracle_params;

typedef struct {
  const double *data[2];        /* ndata x {x, y, z, w}                       */
  const double *rand[2];        /* nrand x {x, y, z, w} (surveys)             */
  size_t ndata[2], nrand[2];
  double wdata[2], wrand[2];    /* sum of completeness weights                */
  double alpha[2], shot[2], norm[2];    /* surveys: as read_cata computes     */
} oracle_cats;

typedef struct oracle_result_s oracle_result;

enum {
  ORACLE_GET_K = 0, ORACLE_GET_KEDGE, ORACLE_GET_KM, ORACLE_GET_CNT,
  ORACLE_GET_LCNT, ORACLE_GET_PL, ORACLE_GET_XPL, ORACLE_GET_SHOT,
  ORACLE_GET_NORM, ORACLE_GET_BMIN, ORACLE_GET_BSIZE, ORACLE_GET_FR,
  ORACLE_GET_FRL
};

const char *oracle_backend(void);
int oracle_real_size(void);
/* keep_mesh != 0 also stores copies of the real-space meshes (Fr, and Frl when
   interlaced) as they are right after mesh generation. Returns NULL on error. */
oracle_result *oracle_run(const oracle_params *par, const oracle_cats *in,
    int keep_mesh);
void oracle_free(oracle_result *r);
int oracle_nbin(const oracle_result *r);
int oracle_nl(const oracle_result *r);
size_t oracle_ntot(const oracle_result *r);
double oracle_time(const oracle_result *r, int which);  /* 0 mesh, 1 pk */
long oracle_get(const oracle_result *r, int what, int idx, void *dst);

#ifdef __cplusplus
}
#endif
#endif
