/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 *
 * Minimal FFTW3-API header for building the UNMODIFIED reference sources
 * (/root/reference/src/genr_mesh.c, multipole.c) into oracle/_ref/ in an image
 * that has no FFTW.  It declares exactly the symbols the reference binds
 * through src/fftw_define.h:32-62 (double and -DSINGLE_PREC twins); they are
 * implemented by fftw_shim.c on top of oracle/fft_cpu.c.
 *
 * This is NOT FFTW and is not API-complete; it exists so that the reference's
 * own arithmetic can run as the parity oracle.  FFT backend name reported in
 * benchmarks: "fftcpu-shim (Stockham mixed radix, OpenMP)".
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct oracle_fftw_plan_s *fftw_plan;
typedef struct oracle_fftwf_plan_s *fftwf_plan;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);
void fftw_cleanup(void);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in,
    fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in,
    double *out, unsigned flags);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out);
void fftw_destroy_plan(fftw_plan p);

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int nthreads);
void fftwf_cleanup_threads(void);
void fftwf_cleanup(void);
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in,
    fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in,
    float *out, unsigned flags);
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out);
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out);
void fftwf_destroy_plan(fftwf_plan p);

#ifdef __cplusplus
}
#endif
#endif
