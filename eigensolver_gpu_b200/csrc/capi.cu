// eigb200 -- extern "C" boundary (include/eigb200.h).
#include "eigb200.h"
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"
#include <string.h>

namespace eigb200 {
const char* last_error();
}
using namespace eigb200;

#define API_BEGIN() \
  if (ctx_init() != 0) return -1;

extern "C" {

int eigb200_init(void) { return ctx_init(); }
int eigb200_finalize(void) {
  Context& c = ctx();
  if (c.scratch) { cudaFree(c.scratch); c.scratch = nullptr; c.scratch_bytes = 0; }
  return 0;
}
const char* eigb200_last_error(void) { return last_error(); }
int eigb200_set_stream(void* s) { ctx().stream = (cudaStream_t)s; return 0; }
int eigb200_version(void) { return 100; }

int eigb200_set_option(const char* name, int value) { return set_option(name, value); }
int eigb200_get_option(const char* name) { return get_option(name); }

int eigb200_dgemm(char ta, char tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B,
                  int ldb, double beta, double* C, int ldc) {
  API_BEGIN();
  return gemm<double>(ctx().stream, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0);
}
int eigb200_zgemm(char ta, char tb, int m, int n, int k, double alpha, const void* A, int lda, const void* B,
                  int ldb, double beta, void* C, int ldc) {
  API_BEGIN();
  return gemm<double2>(ctx().stream, ta, tb, m, n, k, alpha, (const double2*)A, lda, (const double2*)B, ldb, beta,
                       (double2*)C, ldc, 0);
}
int eigb200_dsyr2k(int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                   double* C, int ldc) {
  API_BEGIN();
  return her2k_upper<double>(ctx().stream, 'N', n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
int eigb200_zher2k(int n, int k, double alpha, const void* A, int lda, const void* B, int ldb, double beta, void* C,
                   int ldc) {
  API_BEGIN();
  return her2k_upper<double2>(ctx().stream, 'N', n, k, alpha, (const double2*)A, lda, (const double2*)B, ldb, beta,
                              (double2*)C, ldc);
}

int eigb200_dsymv(int n, const double* A, int lda, const double* x, double* y) {
  API_BEGIN();
  return hemv_upper<double>(ctx().stream, n, A, lda, x, y);
}
int eigb200_zhemv(int n, const void* A, int lda, const void* x, void* y) {
  API_BEGIN();
  return hemv_upper<double2>(ctx().stream, n, (const double2*)A, lda, (const double2*)x, (double2*)y);
}
int eigb200_dsytrd(int n, double* A, int lda, double* d, double* e, double* tau) {
  API_BEGIN();
  return hetrd_upper<double>(ctx().stream, n, A, lda, d, e, tau);
}
int eigb200_zhetrd(int n, void* A, int lda, double* d, double* e, void* tau) {
  API_BEGIN();
  return hetrd_upper<double2>(ctx().stream, n, (double2*)A, lda, d, e, (double2*)tau);
}

int eigb200_dstedc(int n, double* d, double* e, double* Q, int ldq) {
  API_BEGIN();
  size_t need = stedc_scratch_bytes(n);
  void* scr = ctx_scratch(need);
  if (!scr) return -1;
  return stedc_device(ctx().stream, n, d, e, Q, ldq, scr, ctx().scratch_bytes);
}

}  // extern "C"
