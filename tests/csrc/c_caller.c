/* Compiled C caller of the drop-in boundary (plain C, no Python in the process): calls eigb200_zhegvdx / eigb200_dsygvdx
 * with exactly the buffers and lwork formulas the reference's test driver uses for its "CUSTOM" case
 * (test_driver/test_zhegvdx.F90:266-293: lwork = N, lrwork = 1+5N+2N*N, liwork = 3+5N on the host;
 * lwork_d = 2*64*64 + 65*N, lrwork_d = N on the device; test_dsygvdx.F90:292-314 for the real case), as a Fortran/C
 * application would through the ISO_C_BINDING shim.  Verifies the residual ||A x - lambda B x|| on the host.
 *
 *   c_caller z|d N M   -> prints "OK ..." and exits 0, or "FAIL ..." and exits 1
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eigb200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL cuda %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

static double lcg(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53);
}

int main(int argc, char** argv) {
  if (argc < 4) { printf("usage: c_caller z|d N M\n"); return 2; }
  const int cplx = argv[1][0] == 'z';
  const int n = atoi(argv[2]), m = atoi(argv[3]);
  const int es = cplx ? 2 : 1;          /* doubles per element */
  const size_t nn = (size_t)n * n;
  double* A = (double*)calloc(nn * es, sizeof(double));
  double* B = (double*)calloc(nn * es, sizeof(double));
  unsigned long long seed = 12345;
  /* Hermitian A with O(1) entries, B = I*4 + small Hermitian perturbation (well conditioned, positive definite) */
  for (int c = 0; c < n; ++c)
    for (int r = 0; r <= c; ++r) {
      double ar = lcg(&seed) - 0.5, ai = cplx && r != c ? lcg(&seed) - 0.5 : 0.0;
      double br = 0.5 * (lcg(&seed) - 0.5) / sqrt((double)n), bi = cplx && r != c ? 0.5 * (lcg(&seed) - 0.5) / sqrt((double)n) : 0.0;
      if (r == c) br += 4.0;
      A[((size_t)c * n + r) * es] = ar; B[((size_t)c * n + r) * es] = br;
      A[((size_t)r * n + c) * es] = ar; B[((size_t)r * n + c) * es] = br;
      if (cplx) {
        A[((size_t)c * n + r) * es + 1] = ai; B[((size_t)c * n + r) * es + 1] = bi;
        A[((size_t)r * n + c) * es + 1] = -ai; B[((size_t)r * n + c) * es + 1] = -bi;
      }
    }
  if (eigb200_init() != 0) { printf("FAIL init: %s\n", eigb200_last_error()); return 1; }
  /* device buffers, sized as the reference's driver does */
  const int lwork_d = 2 * 64 * 64 + (cplx ? 65 : 66) * n, lrwork_d = n;
  double *A_d, *B_d, *Z_d, *w_d, *work_d, *rwork_d = NULL;
  CK(cudaMalloc((void**)&A_d, nn * es * sizeof(double)));
  CK(cudaMalloc((void**)&B_d, nn * es * sizeof(double)));
  CK(cudaMalloc((void**)&Z_d, nn * es * sizeof(double)));
  CK(cudaMalloc((void**)&w_d, (size_t)n * sizeof(double)));
  CK(cudaMalloc((void**)&work_d, (size_t)lwork_d * es * sizeof(double)));
  if (cplx) CK(cudaMalloc((void**)&rwork_d, (size_t)lrwork_d * sizeof(double)));
  CK(cudaMemcpy(A_d, A, nn * es * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B_d, B, nn * es * sizeof(double), cudaMemcpyHostToDevice));
  /* host workspaces: the reference's formulas (pinned in the reference; plain or NULL is accepted here) */
  const int lwork_h = cplx ? n : 1 + 6 * n + 2 * n * n, lrwork_h = 1 + 5 * n + 2 * n * n, liwork_h = 3 + 5 * n;
  double* work_h = (double*)malloc((size_t)lwork_h * es * sizeof(double));
  double* rwork_h = cplx ? (double*)malloc((size_t)lrwork_h * sizeof(double)) : NULL;
  int* iwork_h = (int*)malloc((size_t)liwork_h * sizeof(int));
  double *Z_h, *w_h;
  CK(cudaMallocHost((void**)&Z_h, nn * es * sizeof(double)));
  CK(cudaMallocHost((void**)&w_h, (size_t)n * sizeof(double)));
  int info = 7;
  if (cplx)
    eigb200_zhegvdx(n, A_d, n, B_d, n, Z_d, n, 1, m, w_d, work_d, lwork_d, rwork_d, lrwork_d, work_h, lwork_h, rwork_h, lrwork_h,
                    iwork_h, liwork_h, Z_h, n, w_h, &info, 0);
  else
    eigb200_dsygvdx(n, A_d, n, B_d, n, Z_d, n, 1, m, w_d, work_d, lwork_d, work_h, lwork_h, iwork_h, liwork_h, Z_h, n, w_h, &info, 0);
  if (info != 0) { printf("FAIL info=%d: %s\n", info, eigb200_last_error()); return 1; }
  /* host residual of every returned pair, in units of n eps ||A||_1 ||x|| */
  double anorm = 0.0;
  for (int c = 0; c < n; ++c) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += hypot(A[((size_t)c * n + r) * es], cplx ? A[((size_t)c * n + r) * es + 1] : 0.0);
    if (s > anorm) anorm = s;
  }
  double worst = 0.0;
  double* y = (double*)malloc((size_t)n * 2 * sizeof(double));
  for (int j = 0; j < m; ++j) {
    const double lam = w_h[j];
    double xn = 0.0, rn = 0.0;
    memset(y, 0, (size_t)n * 2 * sizeof(double));
    for (int c = 0; c < n; ++c) {
      const double xr = Z_h[((size_t)j * n + c) * es], xi = cplx ? Z_h[((size_t)j * n + c) * es + 1] : 0.0;
      xn += xr * xr + xi * xi;
      for (int r = 0; r < n; ++r) {
        const size_t o = ((size_t)c * n + r) * es;
        const double mr = A[o] - lam * B[o], mi = cplx ? A[o + 1] - lam * B[o + 1] : 0.0;
        y[2 * r] += mr * xr - mi * xi;
        y[2 * r + 1] += mr * xi + mi * xr;
      }
    }
    for (int r = 0; r < n; ++r) rn += y[2 * r] * y[2 * r] + y[2 * r + 1] * y[2 * r + 1];
    const double rel = sqrt(rn) / (n * 2.220446049250313e-16 * anorm * sqrt(xn));
    if (rel > worst) worst = rel;
    if (j > 0 && w_h[j] < w_h[j - 1]) { printf("FAIL eigenvalues not ascending at %d\n", j); return 1; }
  }
  if (!(worst < 30.0)) { printf("FAIL residual %.3g (gate 30)\n", worst); return 1; }
  printf("OK %s n=%d m=%d residual_max=%.3g w[0]=%.12g w[m-1]=%.12g\n", cplx ? "zhegvdx" : "dsygvdx", n, m, worst, w_h[0], w_h[m - 1]);
  eigb200_finalize();
  return 0;
}
