// Micro-benchmark (development tool, not part of the product): cuFFT D2Z 3-D layouts.
#include <cufft.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { auto e = (x); if (e) { printf("fail %s: %d line %d\n", #x, (int) e, __LINE__); exit(1); } } while (0)
static float run(cufftHandle p, double *in, cufftDoubleComplex *out, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; i++) CK(cufftExecD2Z(p, in, out));
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) CK(cufftExecD2Z(p, in, out));
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main(int argc, char **argv) {
  int ng = argc > 1 ? atoi(argv[1]) : 1024;
  long long n[3] = {ng, ng, ng};
  for (int pad = 0; pad < 3; pad++) {
    int ngk = ng / 2 + 1;
    int ngkp = pad == 0 ? ngk : (pad == 1 ? ((ngk + 7) & ~7) : ((ngk + 15) & ~15) + 8);
    long long rembed[3] = {ng, ng, 2LL * ngkp}, cembed[3] = {ng, ng, ngkp};
    size_t bytes = (size_t) ng * ng * ngkp * 16;
    double *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    cufftHandle p; size_t ws;
    CK(cufftCreate(&p));
    CK(cufftMakePlanMany64(p, 3, n, rembed, 1, (long long) ng * ng * 2 * ngkp, cembed, 1, (long long) ng * ng * ngkp, CUFFT_D2Z, 1, &ws));
    float ms = run(p, buf, (cufftDoubleComplex *) buf, 5);
    printf("in-place ngk_pad=%d: %.3f ms  (work %.2f GB)\n", ngkp, ms, ws / 1e9);
    cufftDestroy(p);
    if (pad == 0) {
      double *rin; CK(cudaMalloc(&rin, (size_t) ng * ng * ng * 8)); CK(cudaMemset(rin, 0, (size_t) ng * ng * ng * 8));
      CK(cufftCreate(&p));
      CK(cufftMakePlanMany64(p, 3, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 1, &ws));
      ms = run(p, rin, (cufftDoubleComplex *) buf, 5);
      printf("out-of-place dense: %.3f ms (work %.2f GB)\n", ms, ws / 1e9);
      cufftDestroy(p); cudaFree(rin);
    }
    // decomposed: batched 2-D (y,z) r2c + strided 1-D along x
    {
      long long n2[2] = {ng, ng}, re2[2] = {ng, 2LL * ngkp}, ce2[2] = {ng, ngkp};
      cufftHandle p2, p1;
      CK(cufftCreate(&p2));
      CK(cufftMakePlanMany64(p2, 2, n2, re2, 1, (long long) ng * 2 * ngkp, ce2, 1, (long long) ng * ngkp, CUFFT_D2Z, ng, &ws));
      long long n1[1] = {ng}, e1[1] = {ng}; long long lines = (long long) ng * ngkp;
      CK(cufftCreate(&p1));
      CK(cufftMakePlanMany64(p1, 1, n1, e1, lines, 1, e1, lines, 1, CUFFT_Z2Z, lines, &ws));
      cudaEvent_t a, b, c2; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c2);
      for (int w = 0; w < 2; w++) { CK(cufftExecD2Z(p2, buf, (cufftDoubleComplex *) buf)); CK(cufftExecZ2Z(p1, (cufftDoubleComplex *) buf, (cufftDoubleComplex *) buf, CUFFT_FORWARD)); }
      float t2 = 0, t1 = 0;
      for (int r = 0; r < 5; r++) {
        cudaEventRecord(a); CK(cufftExecD2Z(p2, buf, (cufftDoubleComplex *) buf)); cudaEventRecord(b);
        CK(cufftExecZ2Z(p1, (cufftDoubleComplex *) buf, (cufftDoubleComplex *) buf, CUFFT_FORWARD)); cudaEventRecord(c2);
        cudaEventSynchronize(c2); float x, y; cudaEventElapsedTime(&x, a, b); cudaEventElapsedTime(&y, b, c2); t2 += x; t1 += y;
      }
      printf("  decomposed ngk_pad=%d: 2-D batch %.3f ms + 1-D strided %.3f ms\n", ngkp, t2 / 5, t1 / 5);
      cufftDestroy(p2); cufftDestroy(p1);
    }
    cudaFree(buf);
  }
  return 0;
}
