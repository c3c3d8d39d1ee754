/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 * FFTW3-API symbols needed by the unmodified reference (see fftw3.h here),
 * forwarded to the oracle's own CPU FFT engine (oracle/fft_cpu.c).
 */
#define _GNU_SOURCE
#include "fftw3.h"
#include "../fft_cpu.h"
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct oracle_fftw_plan_s { fftcpu_d_plan3d *pl; };
struct oracle_fftwf_plan_s { fftcpu_f_plan3d *pl; };

static void *shim_alloc(size_t n) {
  void *p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64)) return NULL;
  return p;
}

static void shim_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void) n;
#endif
}

/* ---- double ---- */
void *fftw_malloc(size_t n) { return shim_alloc(n); }
void fftw_free(void *p) { free(p); }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int n) { shim_threads(n); }
void fftw_cleanup_threads(void) {}
void fftw_cleanup(void) {}

static fftw_plan plan_d(int n0, int n1, int n2) {
  fftw_plan p = malloc(sizeof *p);
  if (!p) return NULL;
  p->pl = fftcpu_d_plan3d_create(n0, n1, n2);
  if (!p->pl) { free(p); return NULL; }
  return p;
}
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in,
    fftw_complex *out, unsigned flags) {
  (void) in; (void) out; (void) flags;
  return plan_d(n0, n1, n2);
}
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in,
    double *out, unsigned flags) {
  (void) in; (void) out; (void) flags;
  return plan_d(n0, n1, n2);
}
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out) {
  fftcpu_d_r2c_3d(p->pl, in, (double *) out);
}
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out) {
  fftcpu_d_c2r_3d(p->pl, (double *) in, out);
}
void fftw_destroy_plan(fftw_plan p) {
  if (!p) return;
  fftcpu_d_plan3d_destroy(p->pl);
  free(p);
}

/* ---- float ---- */
void *fftwf_malloc(size_t n) { return shim_alloc(n); }
void fftwf_free(void *p) { free(p); }
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int n) { shim_threads(n); }
void fftwf_cleanup_threads(void) {}
void fftwf_cleanup(void) {}

static fftwf_plan plan_f(int n0, int n1, int n2) {
  fftwf_plan p = malloc(sizeof *p);
  if (!p) return NULL;
  p->pl = fftcpu_f_plan3d_create(n0, n1, n2);
  if (!p->pl) { free(p); return NULL; }
  return p;
}
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in,
    fftwf_complex *out, unsigned flags) {
  (void) in; (void) out; (void) flags;
  return plan_f(n0, n1, n2);
}
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in,
    float *out, unsigned flags) {
  (void) in; (void) out; (void) flags;
  return plan_f(n0, n1, n2);
}
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out) {
  fftcpu_f_r2c_3d(p->pl, in, (float *) out);
}
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out) {
  fftcpu_f_c2r_3d(p->pl, (float *) in, out);
}
void fftwf_destroy_plan(fftwf_plan p) {
  if (!p) return;
  fftcpu_f_plan3d_destroy(p->pl);
  free(p);
}
