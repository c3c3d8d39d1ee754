/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 *
 * Self-contained CPU FFT engine used by the oracle only: it stands in for the
 * FFTW3 dependency of the reference (src/fftw_define.h:32-62 binds fftw_* /
 * fftwf_*; FFTW itself is not vendored in /root/reference and is not installed
 * in this image).  The DFT is mathematically defined (forward sign -1, backward
 * sign +1, both unnormalised, half spectrum on the last axis), so any correct
 * double FFT agrees with FFTW to ~1e-15 relative.
 *
 * Algorithm: Stockham autosort, mixed radix {4,2,3,5,generic}, with the batch of
 * transforms as the innermost (contiguous) index so that the butterfly loops
 * vectorise without shuffles.  Real transforms pair two real lines into one
 * complex line.  OpenMP over blocks of lines.
 *
 * This file is a "template": include it with FC_REAL and FC_NAME(x) defined.
 */

#ifndef FC_REAL
#error "define FC_REAL and FC_NAME before including fft_cpu_impl.h"
#endif

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FC_MAXSTAGE 64

typedef struct {
  int n;                        /* transform length                       */
  int nstage;                   /* number of Stockham stages              */
  int radix[FC_MAXSTAGE];       /* radix of every stage                   */
  FC_REAL *tw[FC_MAXSTAGE];     /* twiddles per stage: [m][r-1][2]        */
  FC_REAL *gen;                 /* roots of unity of the largest generic radix */
  int genr;
} FC_NAME(plan1d);

struct FC_NAME(plan3d_s) {
  int n0, n1, n2, n2c;          /* n2c = n2/2 + 1                         */
  int batch;                    /* lines transformed together             */
  FC_NAME(plan1d) p0, p1, p2;
};

/* ------------------------------------------------------------------------- */
/* 1-D plans                                                                 */
/* ------------------------------------------------------------------------- */

static int FC_NAME(plan1d_init)(FC_NAME(plan1d) *p, int n) {
  memset(p, 0, sizeof *p);
  p->n = n;
  int rem = n, ns = 0;
  /* prefer radix 4, then 2, 3, 5, then whatever prime is left */
  while (rem % 4 == 0) { p->radix[ns++] = 4; rem /= 4; }
  while (rem % 2 == 0) { p->radix[ns++] = 2; rem /= 2; }
  while (rem % 3 == 0) { p->radix[ns++] = 3; rem /= 3; }
  while (rem % 5 == 0) { p->radix[ns++] = 5; rem /= 5; }
  for (int f = 7; rem > 1; f += 2) {
    while (rem % f == 0) {
      if (ns >= FC_MAXSTAGE) return -1;
      p->radix[ns++] = f; rem /= f;
      if (f > p->genr) p->genr = f;
    }
  }
  p->nstage = ns;
  /* twiddles: stage with current length nc = r*m needs w_nc^{p*j}, p<m, 0<j<r */
  int nc = n;
  for (int s = 0; s < ns; s++) {
    int r = p->radix[s], m = nc / r;
    p->tw[s] = malloc(sizeof(FC_REAL) * 2 * (size_t) m * (r - 1) + 16);
    if (!p->tw[s]) return -1;
    for (int q = 0; q < m; q++)
      for (int j = 1; j < r; j++) {
        /* reduce the angle index exactly in integers before calling libm */
        long long idx = ((long long) q * j) % nc;
        double ang = -2.0 * M_PI * (double) idx / (double) nc;
        p->tw[s][2 * ((size_t) q * (r - 1) + (j - 1)) + 0] = (FC_REAL) cos(ang);
        p->tw[s][2 * ((size_t) q * (r - 1) + (j - 1)) + 1] = (FC_REAL) sin(ang);
      }
    nc = m;
  }
  if (p->genr) {
    /* table of exp(-2 pi i k / R) for the largest generic radix R; smaller
       generic radices recompute on the fly (rare path)                      */
    p->gen = malloc(sizeof(FC_REAL) * 2 * p->genr);
    if (!p->gen) return -1;
    for (int k = 0; k < p->genr; k++) {
      double ang = -2.0 * M_PI * k / p->genr;
      p->gen[2 * k] = (FC_REAL) cos(ang);
      p->gen[2 * k + 1] = (FC_REAL) sin(ang);
    }
  }
  return 0;
}

static void FC_NAME(plan1d_free)(FC_NAME(plan1d) *p) {
  for (int s = 0; s < p->nstage; s++) free(p->tw[s]);
  free(p->gen);
  memset(p, 0, sizeof *p);
}

/* ------------------------------------------------------------------------- */
/* Stockham stages.  Arrays are split re/im, element (t, q) at t*S + q where */
/* S (= stride, a multiple of the batch) is contiguous.                      */
/* ------------------------------------------------------------------------- */

static inline void FC_NAME(stage2)(int m, size_t S, const FC_REAL *tw,
    const FC_REAL *restrict xr, const FC_REAL *restrict xi,
    FC_REAL *restrict yr, FC_REAL *restrict yi) {
  for (int p = 0; p < m; p++) {
    const FC_REAL wr = tw[2 * p], wi = tw[2 * p + 1];
    const FC_REAL *ar = xr + S * p, *ai = xi + S * p;
    const FC_REAL *br = xr + S * (p + m), *bi = xi + S * (p + m);
    FC_REAL *o0r = yr + S * (2 * p), *o0i = yi + S * (2 * p);
    FC_REAL *o1r = o0r + S, *o1i = o0i + S;
    for (size_t q = 0; q < S; q++) {
      FC_REAL tr = ar[q] - br[q], ti = ai[q] - bi[q];
      o0r[q] = ar[q] + br[q];
      o0i[q] = ai[q] + bi[q];
      o1r[q] = tr * wr - ti * wi;
      o1i[q] = tr * wi + ti * wr;
    }
  }
}

static inline void FC_NAME(stage4)(int m, size_t S, const FC_REAL *tw,
    const FC_REAL *restrict xr, const FC_REAL *restrict xi,
    FC_REAL *restrict yr, FC_REAL *restrict yi) {
  for (int p = 0; p < m; p++) {
    const FC_REAL w1r = tw[6 * p], w1i = tw[6 * p + 1];
    const FC_REAL w2r = tw[6 * p + 2], w2i = tw[6 * p + 3];
    const FC_REAL w3r = tw[6 * p + 4], w3i = tw[6 * p + 5];
    const FC_REAL *a0r = xr + S * p, *a0i = xi + S * p;
    const FC_REAL *a1r = a0r + S * m, *a1i = a0i + S * m;
    const FC_REAL *a2r = a1r + S * m, *a2i = a1i + S * m;
    const FC_REAL *a3r = a2r + S * m, *a3i = a2i + S * m;
    FC_REAL *o0r = yr + S * (4 * (size_t) p), *o0i = yi + S * (4 * (size_t) p);
    FC_REAL *o1r = o0r + S, *o1i = o0i + S;
    FC_REAL *o2r = o1r + S, *o2i = o1i + S;
    FC_REAL *o3r = o2r + S, *o3i = o2i + S;
    for (size_t q = 0; q < S; q++) {
      /* forward DFT-4: uses -i for the odd differences */
      FC_REAL s02r = a0r[q] + a2r[q], s02i = a0i[q] + a2i[q];
      FC_REAL d02r = a0r[q] - a2r[q], d02i = a0i[q] - a2i[q];
      FC_REAL s13r = a1r[q] + a3r[q], s13i = a1i[q] + a3i[q];
      FC_REAL d13r = a1r[q] - a3r[q], d13i = a1i[q] - a3i[q];
      /* -i * d13 = (d13i, -d13r) */
      FC_REAL t1r = d02r + d13i, t1i = d02i - d13r;
      FC_REAL t2r = s02r - s13r, t2i = s02i - s13i;
      FC_REAL t3r = d02r - d13i, t3i = d02i + d13r;
      o0r[q] = s02r + s13r;
      o0i[q] = s02i + s13i;
      o1r[q] = t1r * w1r - t1i * w1i;
      o1i[q] = t1r * w1i + t1i * w1r;
      o2r[q] = t2r * w2r - t2i * w2i;
      o2i[q] = t2r * w2i + t2i * w2r;
      o3r[q] = t3r * w3r - t3i * w3i;
      o3i[q] = t3r * w3i + t3i * w3r;
    }
  }
}

/* generic radix (3, 5 and any other prime): O(r^2) butterfly */
static void FC_NAME(stageg)(int r, int m, size_t S, const FC_REAL *tw,
    const FC_REAL *restrict xr, const FC_REAL *restrict xi,
    FC_REAL *restrict yr, FC_REAL *restrict yi) {
  FC_REAL cr[r], ci[r];         /* exp(-2 pi i k / r), in double then cast */
  for (int k = 0; k < r; k++) {
    double ang = -2.0 * M_PI * k / r;
    cr[k] = (FC_REAL) cos(ang);
    ci[k] = (FC_REAL) sin(ang);
  }
  for (int p = 0; p < m; p++) {
    for (int j = 0; j < r; j++) {
      FC_REAL *outr = yr + S * ((size_t) r * p + j);
      FC_REAL *outi = yi + S * ((size_t) r * p + j);
      FC_REAL wr = 1, wi = 0;
      if (j) {
        wr = tw[2 * ((size_t) p * (r - 1) + (j - 1))];
        wi = tw[2 * ((size_t) p * (r - 1) + (j - 1)) + 1];
      }
      for (size_t q = 0; q < S; q++) {
        FC_REAL accr = 0, acci = 0;
        for (int k = 0; k < r; k++) {
          int e = (j * k) % r;
          FC_REAL ar = xr[S * ((size_t) p + (size_t) k * m) + q];
          FC_REAL ai = xi[S * ((size_t) p + (size_t) k * m) + q];
          accr += ar * cr[e] - ai * ci[e];
          acci += ar * ci[e] + ai * cr[e];
        }
        outr[q] = accr * wr - acci * wi;
        outi[q] = accr * wi + acci * wr;
      }
    }
  }
}

/* Forward transform of `B` interleaved lines of length p->n.
 * In/out: (re, im); scratch: (wr, wi), all of size n*B.  The result is left in
 * (re, im).  A backward (sign +1) transform is obtained by swapping the roles
 * of re and im at the call site. */
static void FC_NAME(fft_lines)(const FC_NAME(plan1d) *p, size_t B,
    FC_REAL *re, FC_REAL *im, FC_REAL *wr, FC_REAL *wi) {
  FC_REAL *xr = re, *xi = im, *yr = wr, *yi = wi;
  int nc = p->n;
  size_t S = B;
  for (int s = 0; s < p->nstage; s++) {
    int r = p->radix[s], m = nc / r;
    if (r == 4) FC_NAME(stage4)(m, S, p->tw[s], xr, xi, yr, yi);
    else if (r == 2) FC_NAME(stage2)(m, S, p->tw[s], xr, xi, yr, yi);
    else FC_NAME(stageg)(r, m, S, p->tw[s], xr, xi, yr, yi);
    FC_REAL *t;
    t = xr; xr = yr; yr = t;
    t = xi; xi = yi; yi = t;
    nc = m;
    S *= r;
  }
  if (xr != re) {
    memcpy(re, xr, sizeof(FC_REAL) * p->n * B);
    memcpy(im, xi, sizeof(FC_REAL) * p->n * B);
  }
}

/* ------------------------------------------------------------------------- */
/* 3-D real <-> half-complex                                                 */
/* ------------------------------------------------------------------------- */

FC_NAME(plan3d) *FC_NAME(plan3d_create)(int n0, int n1, int n2) {
  if (n0 < 1 || n1 < 1 || n2 < 1) return NULL;
  FC_NAME(plan3d) *pl = calloc(1, sizeof *pl);
  if (!pl) return NULL;
  pl->n0 = n0; pl->n1 = n1; pl->n2 = n2; pl->n2c = n2 / 2 + 1;
  pl->batch = (int) (64 / sizeof(FC_REAL));
  if (FC_NAME(plan1d_init)(&pl->p0, n0) || FC_NAME(plan1d_init)(&pl->p1, n1)
      || FC_NAME(plan1d_init)(&pl->p2, n2)) {
    FC_NAME(plan3d_destroy)(pl);
    return NULL;
  }
  return pl;
}

void FC_NAME(plan3d_destroy)(FC_NAME(plan3d) *pl) {
  if (!pl) return;
  FC_NAME(plan1d_free)(&pl->p0);
  FC_NAME(plan1d_free)(&pl->p1);
  FC_NAME(plan1d_free)(&pl->p2);
  free(pl);
}

/* complex pass along one strided axis of the half-complex array `c`
 * (interleaved re,im).  `n` points with stride `str` (in complex elements);
 * `nouter` x `ninner` independent lines: line (o, k) starts at o*ostr + k.
 * backward != 0 swaps re/im to conjugate the transform. */
static void FC_NAME(pass_strided)(const FC_NAME(plan1d) *p, FC_REAL *c,
    size_t str, size_t nouter, size_t ostr, size_t ninner, int B,
    int backward) {
  const int n = p->n;
  const size_t nblk = (ninner + B - 1) / B;
#pragma omp parallel
  {
    FC_REAL *buf = malloc(sizeof(FC_REAL) * 4 * (size_t) n * B);
    FC_REAL *re = buf, *im = buf + (size_t) n * B;
    FC_REAL *w0 = im + (size_t) n * B, *w1 = w0 + (size_t) n * B;
#pragma omp for collapse(2) schedule(static)
    for (size_t o = 0; o < nouter; o++) {
      for (size_t blk = 0; blk < nblk; blk++) {
        size_t k0 = blk * B;
        int nb = (int) ((ninner - k0 < (size_t) B) ? ninner - k0 : (size_t) B);
        FC_REAL *base = c + 2 * (o * ostr + k0);
        for (int t = 0; t < n; t++) {
          const FC_REAL *src = base + 2 * str * t;
          FC_REAL *dr = (backward ? im : re) + (size_t) t * B;
          FC_REAL *di = (backward ? re : im) + (size_t) t * B;
          for (int b = 0; b < nb; b++) { dr[b] = src[2 * b]; di[b] = src[2 * b + 1]; }
          for (int b = nb; b < B; b++) { dr[b] = 0; di[b] = 0; }
        }
        FC_NAME(fft_lines)(p, B, re, im, w0, w1);
        for (int t = 0; t < n; t++) {
          FC_REAL *dst = base + 2 * str * t;
          const FC_REAL *sr = (backward ? im : re) + (size_t) t * B;
          const FC_REAL *si = (backward ? re : im) + (size_t) t * B;
          for (int b = 0; b < nb; b++) { dst[2 * b] = sr[b]; dst[2 * b + 1] = si[b]; }
        }
      }
    }
    free(buf);
  }
}

void FC_NAME(r2c_3d)(const FC_NAME(plan3d) *pl, const FC_REAL *in, FC_REAL *out) {
  const int n0 = pl->n0, n1 = pl->n1, n2 = pl->n2, n2c = pl->n2c, B = pl->batch;
  const size_t nrow = (size_t) n0 * n1;
  const size_t nblk = (nrow + 2 * B - 1) / (2 * (size_t) B);

  /* pass along the last (contiguous) axis: two real rows per complex line */
#pragma omp parallel
  {
    FC_REAL *buf = malloc(sizeof(FC_REAL) * 4 * (size_t) n2 * B);
    FC_REAL *re = buf, *im = buf + (size_t) n2 * B;
    FC_REAL *w0 = im + (size_t) n2 * B, *w1 = w0 + (size_t) n2 * B;
#pragma omp for schedule(static)
    for (size_t blk = 0; blk < nblk; blk++) {
      size_t r0 = blk * 2 * B;
      for (int b = 0; b < B; b++) {
        size_t ra = r0 + 2 * b, rb = ra + 1;
        const FC_REAL *pa = (ra < nrow) ? in + ra * n2 : NULL;
        const FC_REAL *pb = (rb < nrow) ? in + rb * n2 : NULL;
        for (int t = 0; t < n2; t++) {
          re[(size_t) t * B + b] = pa ? pa[t] : 0;
          im[(size_t) t * B + b] = pb ? pb[t] : 0;
        }
      }
      FC_NAME(fft_lines)(&pl->p2, B, re, im, w0, w1);
      /* separate: Xa[k] = (Z[k] + conj Z[n-k]) / 2, Xb[k] = (Z[k] - conj Z[n-k]) / 2i */
      for (int b = 0; b < B; b++) {
        size_t ra = r0 + 2 * b, rb = ra + 1;
        if (ra >= nrow) break;
        FC_REAL *oa = out + 2 * ra * n2c;
        FC_REAL *ob = (rb < nrow) ? out + 2 * rb * n2c : NULL;
        for (int k = 0; k < n2c; k++) {
          int kn = (k == 0) ? 0 : n2 - k;
          FC_REAL zr = re[(size_t) k * B + b], zi = im[(size_t) k * B + b];
          FC_REAL yr = re[(size_t) kn * B + b], yi = -im[(size_t) kn * B + b];
          oa[2 * k] = (FC_REAL) 0.5 * (zr + yr);
          oa[2 * k + 1] = (FC_REAL) 0.5 * (zi + yi);
          if (ob) {
            /* (Z - conjZ') / (2i) = ((zi - yi) - i (zr - yr)) / 2 */
            ob[2 * k] = (FC_REAL) 0.5 * (zi - yi);
            ob[2 * k + 1] = (FC_REAL) -0.5 * (zr - yr);
          }
        }
      }
    }
    free(buf);
  }
  /* pass along axis 1: lines (i, k), stride n2c */
  FC_NAME(pass_strided)(&pl->p1, out, (size_t) n2c, (size_t) n0,
      (size_t) n1 * n2c, (size_t) n2c, B, 0);
  /* pass along axis 0: lines (j, k) = flat index, stride n1*n2c */
  FC_NAME(pass_strided)(&pl->p0, out, (size_t) n1 * n2c, 1, 0,
      (size_t) n1 * n2c, B, 0);
}

void FC_NAME(c2r_3d)(const FC_NAME(plan3d) *pl, FC_REAL *in, FC_REAL *out) {
  const int n0 = pl->n0, n1 = pl->n1, n2 = pl->n2, n2c = pl->n2c, B = pl->batch;
  const size_t nrow = (size_t) n0 * n1;
  const size_t nblk = (nrow + 2 * B - 1) / (2 * (size_t) B);

  FC_NAME(pass_strided)(&pl->p0, in, (size_t) n1 * n2c, 1, 0,
      (size_t) n1 * n2c, B, 1);
  FC_NAME(pass_strided)(&pl->p1, in, (size_t) n2c, (size_t) n0,
      (size_t) n1 * n2c, (size_t) n2c, B, 1);

#pragma omp parallel
  {
    FC_REAL *buf = malloc(sizeof(FC_REAL) * 4 * (size_t) n2 * B);
    FC_REAL *re = buf, *im = buf + (size_t) n2 * B;
    FC_REAL *w0 = im + (size_t) n2 * B, *w1 = w0 + (size_t) n2 * B;
#pragma omp for schedule(static)
    for (size_t blk = 0; blk < nblk; blk++) {
      size_t r0 = blk * 2 * B;
      /* Z[k] = Xa[k] + i Xb[k]; Z[n-k] = conj Xa[k] + i conj Xb[k] */
      for (int b = 0; b < B; b++) {
        size_t ra = r0 + 2 * b, rb = ra + 1;
        const FC_REAL *pa = (ra < nrow) ? in + 2 * ra * n2c : NULL;
        const FC_REAL *pb = (rb < nrow) ? in + 2 * rb * n2c : NULL;
        for (int k = 0; k < n2c; k++) {
          FC_REAL ar = pa ? pa[2 * k] : 0, ai = pa ? pa[2 * k + 1] : 0;
          FC_REAL br = pb ? pb[2 * k] : 0, bi = pb ? pb[2 * k + 1] : 0;
          /* the imaginary parts of the self-conjugate bins are ignored, as a
             half-complex inverse does */
          if (k == 0 || 2 * k == n2) { ai = 0; bi = 0; }
          int kn = (k == 0) ? 0 : n2 - k;
          /* swapped storage (im, re) so that fft_lines performs sign +1 */
          im[(size_t) k * B + b] = ar - bi;     /* Re Z[k]  */
          re[(size_t) k * B + b] = ai + br;     /* Im Z[k]  */
          if (kn != k) {
            im[(size_t) kn * B + b] = ar + bi;  /* Re Z[n-k] */
            re[(size_t) kn * B + b] = br - ai;  /* Im Z[n-k] */
          }
        }
      }
      FC_NAME(fft_lines)(&pl->p2, B, re, im, w0, w1);
      for (int b = 0; b < B; b++) {
        size_t ra = r0 + 2 * b, rb = ra + 1;
        if (ra >= nrow) break;
        FC_REAL *oa = out + ra * n2;
        FC_REAL *ob = (rb < nrow) ? out + rb * n2 : NULL;
        for (int t = 0; t < n2; t++) {
          oa[t] = im[(size_t) t * B + b];
          if (ob) ob[t] = re[(size_t) t * B + b];
        }
      }
    }
    free(buf);
  }
}

#undef FC_MAXSTAGE
