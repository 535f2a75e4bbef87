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
int eigb200_set_a_ready_event(void* ev) { ctx().a_ready = (cudaEvent_t)ev; return 0; }
int eigb200_version(void) { return 100; }

// ---- multi-GPU plumbing for the distributed tridiagonalization -------------------------------------------------
int eigb200_mg_alloc(long long bytes, void** dptr, char* handle64) {
  API_BEGIN();
  void* p = nullptr;
  EIGB_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  EIGB_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  EIGB_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *dptr = p;
  return 0;
}
int eigb200_mg_open(const char* handle64, void** dptr) {
  API_BEGIN();
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  EIGB_CUDA_CHECK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int eigb200_mg_flag_bytes(int n, int world) { return world * mg_flag_stride(n) * (int)sizeof(unsigned long long); }
int eigb200_mg_config(int rank, int world, void** wbufs, void** flags, long long wbuf_bytes, void* panel_hook) {
  API_BEGIN();
  if (world < 1 || world > 8 || rank < 0 || rank >= world) { set_last_error("eigb200_mg_config: bad rank/world"); return -1; }
  MgConfig& M = mg();
  M.rank = rank; M.P = world; M.wbuf_bytes = wbuf_bytes;
  // exchange buffers are sized for complex elements: wbuf_bytes = world * 2 * (n + 64) * 16
  M.flag_stride = world > 0 ? mg_flag_stride((int)(wbuf_bytes / ((long long)world * 32)) - 64) : 0;
  for (int q = 0; q < world; ++q) { M.wbuf[q] = wbufs ? wbufs[q] : nullptr; M.flags[q] = flags ? (unsigned long long*)flags[q] : nullptr; }
  M.hook = (panel_hook_t)panel_hook;
  M.active = world > 1;
  // every word of this rank's exchange buffer starts "unset" (sytrd.cu, multi-GPU exchange); the caller synchronises the
  // ranks before the first solve
  if (world > 1 && M.wbuf[rank]) EIGB_CUDA_CHECK(cudaMemset(M.wbuf[rank], 0xFF, (size_t)wbuf_bytes));
  return 0;
}
int eigb200_mg_unique_id(char* id128) {
  API_BEGIN();
  return mg_unique_id(id128);
}
int eigb200_mg_init(int rank, int world, const char* id128) {
  API_BEGIN();
  return mg_init(rank, world, id128);
}
int eigb200_mg_finalize(void) { return mg_finalize(); }
int eigb200_mg_allgather_columns(void* M_d, int ld, int ncols, int elem_bytes) {
  API_BEGIN();
  return mg_allgather_columns(ctx().stream, M_d, ld, ncols, elem_bytes);
}
int eigb200_mg_column_range(int ncols, int world, int rank, int* c0, int* c1) {
  if (world < 1 || rank < 0 || rank >= world) return -1;
  mg_column_range(ncols, world, rank, *c0, *c1);
  return 0;
}

int eigb200_prof_enable(int on) { prof_enable(on); return 0; }
int eigb200_prof_reset(void) { prof_reset(); return 0; }
int eigb200_prof_collect(double* ms, int* cnt, long long* launches) { prof_collect(ms, cnt, launches); return 0; }

long long eigb200_trace_read(unsigned long long* out, long long max_count) {
  auto& v = trace_store();
  long long n = (long long)v.size() < max_count ? (long long)v.size() : max_count;
  for (long long i = 0; i < n; ++i) out[i] = v[i];
  return n;
}

int eigb200_probe_peaks(double* out4) {
  API_BEGIN();
  return probe_peaks(ctx().stream, out4);
}

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
  if (stedc_device(ctx().stream, n, d, e, Q, ldq, scr, ctx().scratch_bytes) != 0) return -1;
  if (status_fetch(ctx().stream) != 0) return -1;
  if (ctx().h_status[ST_STEDC] != 0) { set_last_error("dstedc: device status %d (1 QL, 2 secular, 3 non-finite)", ctx().h_status[ST_STEDC]); return -1; }
  return 0;
}

int eigb200_dstedc_range(int n, double* d, double* e, double* Q, int ldq, int c_lo, int c_hi) {
  API_BEGIN();
  void* scr = ctx_scratch(stedc_scratch_bytes(n));
  if (!scr) return -1;
  if (stedc_device(ctx().stream, n, d, e, Q, ldq, scr, ctx().scratch_bytes, c_lo, c_hi) != 0) return -1;
  if (status_fetch(ctx().stream) != 0) return -1;
  if (ctx().h_status[ST_STEDC] != 0) { set_last_error("dstedc: device status %d (1 QL, 2 secular, 3 non-finite)", ctx().h_status[ST_STEDC]); return -1; }
  return 0;
}
int eigb200_dpotrf(int n, double* B, int ldb, int* info_h) {
  API_BEGIN();
  return potrf_upper<double>(ctx().stream, n, B, ldb, info_h);
}
int eigb200_zpotrf(int n, void* B, int ldb, int* info_h) {
  API_BEGIN();
  return potrf_upper<double2>(ctx().stream, n, (double2*)B, ldb, info_h);
}
int eigb200_dsygst(int n, double* A, int lda, const double* U, int ldu) {
  API_BEGIN();
  return hegst_upper<double>(ctx().stream, n, A, lda, U, ldu, (double*)nullptr, 0);
}
int eigb200_zhegst(int n, void* A, int lda, const void* U, int ldu) {
  API_BEGIN();
  return hegst_upper<double2>(ctx().stream, n, (double2*)A, lda, (const double2*)U, ldu, (double2*)nullptr, 0);
}
int eigb200_dtrsm(char side, char trans, int m, int n, const double* U, int ldu, double* B, int ldb) {
  API_BEGIN();
  return trsm_upper<double>(ctx().stream, side, trans == 'T' ? 'C' : trans, m, n, U, ldu, B, ldb);
}
int eigb200_ztrsm(char side, char trans, int m, int n, const void* U, int ldu, void* B, int ldb) {
  API_BEGIN();
  return trsm_upper<double2>(ctx().stream, side, trans, m, n, (const double2*)U, ldu, (double2*)B, ldb);
}
int eigb200_dormtr(int n, int m, const double* A, int lda, const double* tau, double* Z, int ldz) {
  API_BEGIN();
  void* scr = ctx_scratch(ormtr_scratch_bytes(n, m, sizeof(double)));
  if (!scr) return -1;
  return ormtr_upper<double>(ctx().stream, n, m, A, lda, tau, Z, ldz, scr, ctx().scratch_bytes);
}
int eigb200_zunmtr(int n, int m, const void* A, int lda, const void* tau, void* Z, int ldz) {
  API_BEGIN();
  void* scr = ctx_scratch(ormtr_scratch_bytes(n, m, sizeof(double2)));
  if (!scr) return -1;
  return ormtr_upper<double2>(ctx().stream, n, m, (const double2*)A, lda, (const double2*)tau, (double2*)Z, ldz, scr,
                              ctx().scratch_bytes);
}
int64_t eigb200_scratch_bytes(int n, int is_complex) {
  size_t es = is_complex ? 16 : 8;
  size_t a = (size_t)n * n * 8 + 256 + stedc_scratch_bytes(n);
  size_t b = ormtr_scratch_bytes(n, n, (int)es);
  size_t c = ((size_t)n * n / 16 + (size_t)n * 200) * es + (1 << 20);
  size_t r = a > b ? a : b;
  return (int64_t)(r > c ? r : c);
}

int eigb200_dsygvdx(int n, double* A, int lda, double* B, int ldb, double* Z, int ldz, int il, int iu, double* w,
                    double* work, int lwork, double* work_h, int lwork_h, int* iwork_h, int liwork_h, double* Z_h,
                    int ldz_h, double* w_h, int* info, int skip_host_copy) {
  (void)work_h; (void)iwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return hegvdx_driver<double>(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, nullptr, 0, lwork_h, 0, liwork_h, Z_h,
                               ldz_h, w_h, info, skip_host_copy);
}
int eigb200_zhegvdx(int n, void* A, int lda, void* B, int ldb, void* Z, int ldz, int il, int iu, double* w, void* work,
                    int lwork, double* rwork, int lrwork, void* work_h, int lwork_h, double* rwork_h, int lrwork_h,
                    int* iwork_h, int liwork_h, void* Z_h, int ldz_h, double* w_h, int* info, int skip_host_copy) {
  (void)work_h; (void)rwork_h; (void)iwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return hegvdx_driver<double2>(n, (double2*)A, lda, (double2*)B, ldb, (double2*)Z, ldz, il, iu, w, (double2*)work,
                                lwork, rwork, lrwork, lwork_h, lrwork_h, liwork_h, (double2*)Z_h, ldz_h, w_h, info,
                                skip_host_copy);
}
int eigb200_dsygvdx_mg(int n, double* A, int lda, double* B, int ldb, double* Z, int ldz, int il, int iu, double* w,
                       double* work, int lwork, double* work_h, int lwork_h, int* iwork_h, int liwork_h, double* Z_h,
                       int ldz_h, double* w_h, int* info, int skip_host_copy) {
  (void)work_h; (void)iwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return hegvdx_mg_driver<double>(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, nullptr, 0, lwork_h, 0, liwork_h, Z_h,
                                  ldz_h, w_h, info, skip_host_copy);
}
int eigb200_zhegvdx_mg(int n, void* A, int lda, void* B, int ldb, void* Z, int ldz, int il, int iu, double* w, void* work,
                       int lwork, double* rwork, int lrwork, void* work_h, int lwork_h, double* rwork_h, int lrwork_h,
                       int* iwork_h, int liwork_h, void* Z_h, int ldz_h, double* w_h, int* info, int skip_host_copy) {
  (void)work_h; (void)rwork_h; (void)iwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return hegvdx_mg_driver<double2>(n, (double2*)A, lda, (double2*)B, ldb, (double2*)Z, ldz, il, iu, w, (double2*)work,
                                   lwork, rwork, lrwork, lwork_h, lrwork_h, liwork_h, (double2*)Z_h, ldz_h, w_h, info,
                                   skip_host_copy);
}
int eigb200_dsyevd(int il, int iu, int n, double* A, int lda, double* Z, int ldz, double* w, double* work, int lwork,
                   double* work_h, int lwork_h, int* iwork_h, int liwork_h, double* Z_h, int ldz_h, double* w_h,
                   int* info) {
  (void)work_h; (void)lwork_h; (void)iwork_h; (void)liwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return heevd_driver<double>(il, iu, n, A, lda, Z, ldz, w, work, lwork, nullptr, 0, Z_h, ldz_h, w_h, info);
}
int eigb200_zheevd(int il, int iu, int n, void* A, int lda, void* Z, int ldz, double* w, void* work, int lwork,
                   double* rwork, int lrwork, void* work_h, int lwork_h, double* rwork_h, int lrwork_h, int* iwork_h,
                   int liwork_h, void* Z_h, int ldz_h, double* w_h, int* info) {
  (void)work_h; (void)lwork_h; (void)rwork_h; (void)lrwork_h; (void)iwork_h; (void)liwork_h;
  int dummy = 0;
  if (!info) info = &dummy;
  if (ctx_init() != 0) { *info = -1; return -1; }
  return heevd_driver<double2>(il, iu, n, (double2*)A, lda, (double2*)Z, ldz, w, (double2*)work, lwork, rwork, lrwork,
                               (double2*)Z_h, ldz_h, w_h, info);
}

}  // extern "C"
