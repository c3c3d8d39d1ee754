/*
 * powspec_refabi.h — the reference's own seam, as exported by
 * libpowspec_b200.so.
 *
 * The reference program is a six-stage pipeline (src/powspec.c:23-75); this
 * library replaces the two translation units genr_mesh.o and multipole.o by
 * link-time substitution.  They define exactly five external symbols
 * (SURVEY.md §8b), declared here with the reference's signatures:
 *
 *   MESH *genr_mesh(const CONF *, CATA *)            src/genr_mesh.h:86
 *   void  mesh_destroy(MESH *)                       src/genr_mesh.h:94
 *   PK   *powspec(const CONF *, const CATA *, MESH *) src/multipole.h:78
 *   void  powspec_destroy(PK *)                      src/multipole.h:86
 *   const char *powspec_assign_names[]               src/genr_mesh.h:45
 *
 * The struct definitions below are ABI mirrors (same member order and types, so
 * the same layout) of the reference's CONF (src/load_conf.h:40-93), DATA / CATA
 * (src/read_cata.h:42-59), MESH (src/genr_mesh.h:48-70) and PK
 * (src/multipole.h:38-61, built with -DOMP as the reference Makefile does,
 * Makefile:17).  They are prefixed psb_ref_ so that this header can be included
 * next to the reference's own headers;
 * tests/test_library_cpu.py::test_refabi_struct_layout_matches_reference_headers compiles a
 * translation unit that includes both and static-asserts sizeof and every offset.
 * FFT_PLAN / FFT_REAL* / FFT_CMPLX* members are pointers in both precisions
 * (src/fftw_define.h:32-48), so one layout serves -DSINGLE_PREC as well.
 */
#ifndef POWSPEC_REFABI_H
#define POWSPEC_REFABI_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  char *fconf; int ndata;
  char **dfname, **rfname; int *dftype, *rftype; long *dskip, *rskip;
  char *dcmt, *rcmt; char **dfmtr, **rfmtr, **dpos, **rpos, **dwcomp, **rwcomp;
  char **dwfkp, **rwfkp, **dnz, **rnz, **dsel, **rsel;
  bool has_asc[2]; bool *dcnvt, *rcnvt; bool cnvt;
  double omega_m, omega_l, omega_k, eos_w, ecdst; char *fcdst;
  bool issim; double *los, *bsize, *bpad; int gsize, assign; bool intlace;
  int *poles; int npole; double kmin, kmax; bool logscale; double kbin;
  char **oauto; char *ocross; bool isauto[2]; bool iscross; bool oheader;
  int ovwrite; bool verbose;
} psb_ref_CONF;

typedef struct { double x[3]; double w; } psb_ref_DATA;

typedef struct {
  int num; psb_ref_DATA **data, **rand; size_t *ndata, *nrand;
  double *wdata, *wrand, *alpha, *shot, *norm;
} psb_ref_CATA;

typedef struct {
  int num, Ng, Ngk; size_t Ntot, Ncmplx;
  double min[3], max[3], smin[3], bsize[3];
  bool issim, intlace, fft_init; int assign;
  void *r2c, *c2r;              /* FFT_PLAN: unused here (cuFFT plans live in the context) */
  void **Fr, **Frl; void *alias; void **Fk0, **Fkl; void *Fka;   /* device-resident: NULL */
} psb_ref_MESH;

typedef struct {
  bool issim, log, isauto[2], iscross;
  int nl, nbin, nmu; int *poles; double los[3]; double dk;
  double *kedge, *k, *km; size_t *cnt; double *lcnt; double **pl[2]; double **xpl;
  int nomp; double *pcnt, *plcnt;       /* -DOMP members, left NULL */
} psb_ref_PK;

#ifndef PSB_REFABI_NO_PROTOTYPES
/* In a translation unit that also includes the reference headers, define
 * PSB_REFABI_NO_PROTOTYPES: the prototypes there are the same symbols. */
extern const char *powspec_assign_names[];
psb_ref_MESH *genr_mesh(const psb_ref_CONF *conf, psb_ref_CATA *cat);
void mesh_destroy(psb_ref_MESH *mesh);
psb_ref_PK *powspec(const psb_ref_CONF *conf, const psb_ref_CATA *cat, psb_ref_MESH *mesh);
void powspec_destroy(psb_ref_PK *pk);
/* optional sixth symbol: replaces cnvt_coord.o (src/cnvt_coord.h:50) when the host
 * is linked without it; the conversion then runs on the device inside genr_mesh */
int cnvt_coord(const psb_ref_CONF *conf, psb_ref_CATA *cat);
#endif

#ifdef __cplusplus
}
#endif
#endif
