/*
 * TEST INFRASTRUCTURE (oracle) — not part of the shipped product path.
 * Instantiates the CPU FFT engine for double and float.
 */
#define _GNU_SOURCE
#include "fft_cpu.h"

#define FC_REAL double
#define FC_NAME(x) fftcpu_d_##x
#include "fft_cpu_impl.h"
#undef FC_REAL
#undef FC_NAME

#define FC_REAL float
#define FC_NAME(x) fftcpu_f_##x
#include "fft_cpu_impl.h"
#undef FC_REAL
#undef FC_NAME
