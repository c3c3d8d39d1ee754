/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 * Public interface of the oracle's CPU FFT engine (see fft_cpu_impl.h).
 * Layouts follow the FFTW conventions the reference relies on
 * (src/genr_mesh.c:738-743, src/multipole.c:444,459,493): row-major n0 x n1 x n2
 * reals <-> n0 x n1 x (n2/2+1) interleaved complex, forward sign -1, backward
 * sign +1, both unnormalised.
 */
#ifndef ORACLE_FFT_CPU_H
#define ORACLE_FFT_CPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftcpu_d_plan3d_s fftcpu_d_plan3d;
typedef struct fftcpu_f_plan3d_s fftcpu_f_plan3d;

fftcpu_d_plan3d *fftcpu_d_plan3d_create(int n0, int n1, int n2);
void fftcpu_d_plan3d_destroy(fftcpu_d_plan3d *pl);
void fftcpu_d_r2c_3d(const fftcpu_d_plan3d *pl, const double *in, double *out);
/* destroys `in` */
void fftcpu_d_c2r_3d(const fftcpu_d_plan3d *pl, double *in, double *out);

fftcpu_f_plan3d *fftcpu_f_plan3d_create(int n0, int n1, int n2);
void fftcpu_f_plan3d_destroy(fftcpu_f_plan3d *pl);
void fftcpu_f_r2c_3d(const fftcpu_f_plan3d *pl, const float *in, float *out);
void fftcpu_f_c2r_3d(const fftcpu_f_plan3d *pl, float *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
