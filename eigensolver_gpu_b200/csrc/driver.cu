// eigb200 -- generalized and standard drivers behind the C ABI.
//
// dsygvdx_gpu / zhegvdx_gpu (dsygvdx_gpu.F90:71-168, zhegvdx_gpu.F90:75-182) and dsyevd_gpu / zheevd_gpu
// (dsyevd_gpu.F90:32-132, zheevd_gpu.F90:32-134) re-built on the sm_100a stages of this library:
//   potrf(B) -> [save tril(A) in Z] -> hegst -> hetrd -> [restore tril(A)] -> stedc ON DEVICE -> select il..iu
//   -> back-transform -> trsm with U -> copies to the host buffers.
// Same argument lists, workspace checks and info convention as the reference; the host workspaces are
// accepted (and size-checked, so that a caller sized for the reference keeps working) but unused because the
// divide and conquer no longer runs on the CPU.
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"

namespace eigb200 {

namespace {

template <typename T>
__global__ void select_columns_kernel(const double* __restrict__ Q, int64_t ldq, int n, int c0, int m, T* Z,
                                      int64_t ldz) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r < n && c < m) Z[r + (int64_t)c * ldz] = from_real<T>(Q[r + (int64_t)(c0 + c) * ldq]);
}

}  // namespace

// Standard problem on device data: A (upper) -> w(all n, ascending), Z(:, 0:m) eigenvectors il..iu (1-based).
// d_e, d_tau: device workspaces of n doubles / n T.
template <typename T>
int heevd_core(cudaStream_t s, int n, int il, int iu, T* A, int64_t lda, T* Z, int64_t ldz, double* w, double* d_e,
               T* d_tau, const T* restore_from, int64_t ld_restore) {
  const int m = iu - il + 1;
  if (hetrd_upper<T>(s, n, A, lda, w, d_e, d_tau, /*sync_status=*/false) != 0) return -1;
  if (restore_from) {
    if (restore_lower<T>(s, n, A, lda, restore_from, ld_restore) != 0) return -1;
  }
  size_t nn = (size_t)n * n * sizeof(double);
  size_t need_dc = nn + 256 + stedc_scratch_bytes(n);
  size_t need_bt = ormtr_scratch_bytes(n, m, sizeof(T));
  size_t need = need_dc > need_bt ? need_dc : need_bt;
  char* scr = (char*)ctx_scratch(need);
  if (!scr) return -1;
  double* Qt = (double*)scr;
  if (stedc_device(s, n, w, d_e, Qt, n, scr + ((nn + 255) & ~size_t(255)), ctx().scratch_bytes - ((nn + 255) & ~size_t(255)),
                   il - 1, iu) != 0)
    return -1;
  select_columns_kernel<T><<<dim3(cdiv(n, 256), m), 256, 0, s>>>(Qt, n, n, il - 1, m, Z, ldz);
  EIGB_LAUNCH_CHECK();
  if (ormtr_upper<T>(s, n, m, A, lda, d_tau, Z, ldz, scr, ctx().scratch_bytes) != 0) return -1;
  return 0;
}

template <typename T>
int copy_results_to_host(cudaStream_t s, int n, int m, const T* Z, int64_t ldz, const double* w, T* Z_h, int64_t ldz_h,
                         double* w_h, bool skip_z) {
  if (w_h) EIGB_CUDA_CHECK(cudaMemcpyAsync(w_h, w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (!skip_z && Z_h) {
    EIGB_CUDA_CHECK(cudaMemcpy2DAsync(Z_h, (size_t)ldz_h * sizeof(T), Z, (size_t)ldz * sizeof(T), (size_t)n * sizeof(T),
                                      m, cudaMemcpyDeviceToHost, s));
  }
  return status_fetch(s);      // (stream synchronisation inside)
}

// every early return of a driver reports through *info as well (the reference's only error channel)
#define EIGB_DRV_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      *info = -1;                                                                              \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

// workspace / argument checks shared by the single- and multi-GPU drivers: zhegvdx_gpu.F90:106-127 /
// dsygvdx_gpu.F90:100-113 (64-bit arithmetic: the reference's 1+5N+2N*N overflows default integers at N >= 32767)
template <typename T>
int hegvdx_check_args(int n, int lda, int ldb, int ldz, int il, int iu, int lwork, int lrwork, int lwork_h, int lrwork_h,
                      int liwork_h, int* info) {
  const bool cplx = is_cplx<T>::value;
  const char* name = cplx ? "zhegvdx_gpu" : "dsygvdx_gpu";
  const int64_t N = n;
  const char* msg = nullptr;
  if (cplx) {
    if (lwork < 2 * 64 * 64 + 65 * N) msg = "lwork must be at least 2*64*64 + 65*N";
    else if (lrwork < N) msg = "lrwork must be at least N";
    else if (lwork_h < N) msg = "lwork_h must be at least N";
    // (when the formula exceeds a default integer the caller cannot state it -- N >= 32767 -- and the host
    // workspace is unused here anyway: accept whatever was passed, including an overflowed negative value)
    else if (1 + 5 * N + 2 * N * N <= 2147483647LL && lrwork_h < 1 + 5 * N + 2 * N * N)
      msg = "lrwork_h must be at least 1 + 5*N + 2*N*N";
    else if (liwork_h < N) msg = "liwork_h must be at least 3 + 5*N";
  } else {
    if (lwork < 2 * 64 * 64 + 66 * N) msg = "lwork must be at least 2*64*64 + 66*N";
    else if (1 + 6 * N + 2 * N * N <= 2147483647LL && lwork_h < 1 + 6 * N + 2 * N * N)
      msg = "lwork_h must be at least 1 + 6*N + 2*N*N";
    else if (liwork_h < N) msg = "liwork_h must be at least 3 + 5*N";
  }
  if (!msg && (n < 0 || lda < n || ldb < n || ldz < n)) msg = "N, lda, ldb, ldz inconsistent";
  if (!msg && n > 0 && (il < 1 || iu > n || il > iu)) msg = "need 1 <= il <= iu <= N";
  if (msg) {
    printf(" %s error: %s\n", name, msg);
    set_last_error("%s error: %s", name, msg);
    *info = -1;
    return -1;
  }
  return 0;
}
template int hegvdx_check_args<double>(int, int, int, int, int, int, int, int, int, int, int, int*);
template int hegvdx_check_args<double2>(int, int, int, int, int, int, int, int, int, int, int, int*);

template <typename T>
int hegvdx_driver(int n, T* A, int lda, T* B, int ldb, T* Z, int ldz, int il, int iu, double* w, T* work, int lwork,
                  double* rwork, int lrwork, int lwork_h, int lrwork_h, int liwork_h, T* Z_h, int ldz_h, double* w_h,
                  int* info, int skip_host_copy) {
  const bool cplx = is_cplx<T>::value;
  *info = 0;
  const char* name = cplx ? "zhegvdx_gpu" : "dsygvdx_gpu";
  // the "A is ready" event is one-shot: consumed by this call whatever its outcome
  cudaEvent_t a_ready = ctx().a_ready;
  ctx().a_ready = nullptr;
  if (hegvdx_check_args<T>(n, lda, ldb, ldz, il, iu, lwork, lrwork, lwork_h, lrwork_h, liwork_h, info) != 0) return -1;
  if (n == 0) return 0;
  cudaStream_t s = ctx().stream;
  const int m = iu - il + 1;
  // The pivot status of the factorization stays on the device and is read back with the final synchronisation of the
  // call (the reference blocks on devInfo right here, zhegvdx_gpu.F90:136-141): nothing below can hang on a failed
  // factorization -- every device loop is bounded -- and the host keeps enqueueing work meanwhile.
  prof_begin(PROF_POTRF, s);
  int prc = potrf_upper<T>(s, n, B, ldb, nullptr, /*sync_status=*/false);
  prof_end(PROF_POTRF, s);
  if (prc != 0) {
    printf(" %s error: potrf failed!\n", name);
    *info = -1;
    return -1;
  }
  // a caller that uploads A asynchronously on another stream while B is being factored hands over the event of
  // that copy (eigb200_set_a_ready_event, one-shot): wait for it before A is touched
  if (a_ready != nullptr) EIGB_DRV_CHECK(cudaStreamWaitEvent(s, a_ready, 0));
  // tril(A) -> Z, A <- U^-H A U^-1 (zhegvdx_gpu.F90:145-158)
  prof_begin(PROF_HEGST, s);
  int hrc = hegst_upper<T>(s, n, A, lda, B, ldb, Z, ldz);
  prof_end(PROF_HEGST, s);
  if (hrc != 0) { *info = -1; return -1; }
  double* d_e = cplx ? rwork : reinterpret_cast<double*>(work);
  T* d_tau = cplx ? work : work + n;
  // the strict lower triangle of A is restored from Z before Z is overwritten (zheevd_gpu.F90:89-96); Z is
  // needed as the save area until then, so the restore happens inside heevd_core right after hetrd.
  if (heevd_core<T>(s, n, il, iu, A, lda, Z, ldz, w, d_e, d_tau, Z, ldz) != 0) {
    printf(" %s error: eigensolver stage failed: %s\n", name, "see eigb200_last_error()");
    *info = -1;
    return -1;
  }
  // eigenvectors of the generalized problem: Z <- U^-1 Z (zhegvdx_gpu.F90:169).  With a host copy requested the
  // solve runs by column blocks and the D2H of a finished block overlaps with the solve of the next one on a side
  // stream (the reference does one blocking cudaMemcpy2D after the trsm, zhegvdx_gpu.F90:172-180).
  const bool want_z = skip_host_copy == 0 && Z_h != nullptr;
  Context& c = ctx();
  prof_begin(PROF_TRSM, s);
  int trc = 0;
  if (want_z && c.stream2 != nullptr && c.ev1 != nullptr && m >= 1024) {
    const int nblk = m >= 4096 ? 4 : 2;
    const int cb = (((m + nblk - 1) / nblk) + 63) & ~63;
    cudaEvent_t evs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ne = 0;
    for (int c0 = 0; c0 < m && trc == 0; c0 += cb, ++ne) {
      const int mc = m - c0 < cb ? m - c0 : cb;
      trc = trsm_upper<T>(s, 'L', 'N', n, mc, B, ldb, Z + (int64_t)c0 * ldz, ldz);
      if (trc != 0) break;
      if (cudaEventCreateWithFlags(&evs[ne], cudaEventDisableTiming) != cudaSuccess) { trc = -1; break; }
      if (cudaEventRecord(evs[ne], s) != cudaSuccess || cudaStreamWaitEvent(c.stream2, evs[ne], 0) != cudaSuccess) { trc = -1; break; }
      if (cudaMemcpy2DAsync(Z_h + (int64_t)c0 * ldz_h, (size_t)ldz_h * sizeof(T), Z + (int64_t)c0 * ldz,
                            (size_t)ldz * sizeof(T), (size_t)n * sizeof(T), mc, cudaMemcpyDeviceToHost, c.stream2)
          != cudaSuccess) trc = -1;
    }
    prof_end(PROF_TRSM, s);
    if (trc == 0 && w_h) {
      if (cudaMemcpyAsync(w_h, w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) trc = -1;
    }
    cudaError_t e1 = cudaStreamSynchronize(c.stream2);
    const int e2 = status_fetch(s);
    for (int i = 0; i < 8; ++i) if (evs[i]) cudaEventDestroy(evs[i]);
    if (trc != 0 || e1 != cudaSuccess || e2 != 0) {
      printf(" %s error: solve with U / copy to host failed!\n", name);
      *info = -1;
      return -1;
    }
    if (status_check(name) != 0) { printf(" %s error: %s\n", name, "see eigb200_last_error()"); *info = -1; return -1; }
    return 0;
  }
  trc = trsm_upper<T>(s, 'L', 'N', n, m, B, ldb, Z, ldz);
  prof_end(PROF_TRSM, s);
  if (trc != 0) { *info = -1; return -1; }
  if (copy_results_to_host<T>(s, n, m, Z, ldz, w, Z_h, ldz_h, w_h, skip_host_copy != 0) != 0) {
    printf(" %s error: copy to host failed!\n", name);
    *info = -1;
    return -1;
  }
  if (status_check(name) != 0) { printf(" %s error: %s\n", name, "see eigb200_last_error()"); *info = -1; return -1; }
  return 0;
}

template <typename T>
int heevd_driver(int il, int iu, int n, T* A, int lda, T* Z, int ldz, double* w, T* work, int lwork, double* rwork,
                 int lrwork, T* Z_h, int ldz_h, double* w_h, int* info) {
  const bool cplx = is_cplx<T>::value;
  *info = 0;
  ctx().a_ready = nullptr;       // one-shot event of the generalized driver: never left armed
  const int64_t N = n;
  const char* name = cplx ? "zheevd_gpu" : "dsyevd_gpu";
  const char* msg = nullptr;
  if (cplx) {
    if (lwork < 2 * 64 * 64 + 65 * N) msg = "lwork must be at least 2*64*64 + 65*N";
    else if (lrwork < N) msg = "lrwork must be at least N";
  } else {
    if (lwork < 2 * 64 * 64 + 66 * N) msg = "lwork must be at least 2*64*64 + 66*N";
  }
  if (!msg && (n < 0 || lda < n || ldz < n)) msg = "N, lda, ldz inconsistent";
  if (!msg && n > 0 && (il < 1 || iu > n || il > iu)) msg = "need 1 <= il <= iu <= N";
  if (msg) {
    printf(" %s error: %s\n", name, msg);
    set_last_error("%s error: %s", name, msg);
    *info = -1;
    return -1;
  }
  if (n == 0) return 0;
  cudaStream_t s = ctx().stream;
  double* d_e = cplx ? rwork : reinterpret_cast<double*>(work);
  T* d_tau = cplx ? work : work + n;
  if (heevd_core<T>(s, n, il, iu, A, lda, Z, ldz, w, d_e, d_tau, (const T*)nullptr, 0) != 0) { *info = -1; return -1; }
  EIGB_DRV_CHECK(cudaMemsetAsync(ctx().d_info + ST_POTRF, 0, sizeof(int), s));
  if (copy_results_to_host<T>(s, n, iu - il + 1, Z, ldz, w, Z_h, ldz_h, w_h, false) != 0) { *info = -1; return -1; }
  if (status_check(name) != 0) { printf(" %s error: %s\n", name, "see eigb200_last_error()"); *info = -1; return -1; }
  return 0;
}

template int hegvdx_driver<double>(int, double*, int, double*, int, double*, int, int, int, double*, double*, int,
                                   double*, int, int, int, int, double*, int, double*, int*, int);
template int hegvdx_driver<double2>(int, double2*, int, double2*, int, double2*, int, int, int, double*, double2*, int,
                                    double*, int, int, int, int, double2*, int, double*, int*, int);
template int heevd_driver<double>(int, int, int, double*, int, double*, int, double*, double*, int, double*, int,
                                  double*, int, double*, int*);
template int heevd_driver<double2>(int, int, int, double2*, int, double2*, int, double*, double2*, int, double*, int,
                                   double2*, int, double*, int*);

}  // namespace eigb200
