/* eigb200 -- C ABI of the B200-native generalized symmetric/Hermitian-definite eigensolver.
 *
 * Drop-in boundary for the reference's Fortran entry points (NVIDIA/Eigensolver_gpu):
 *   dsygvdx_gpu  lib_eigsolve/dsygvdx_gpu.F90:71-168      zhegvdx_gpu  lib_eigsolve/zhegvdx_gpu.F90:75-182
 *   dsyevd_gpu   lib_eigsolve/dsyevd_gpu.F90:32-132       zheevd_gpu   lib_eigsolve/zheevd_gpu.F90:32-134
 *   init_eigsolve_gpu  lib_eigsolve/eigsolve_vars.F90:39-59
 * The thin ISO_C_BINDING shim modules in fortran/ forward the reference argument lists to these functions
 * (c_devloc of the device arrays); INTEGRATION.md shows the binding.
 *
 * Conventions: column-major, 1-based il/iu, UPLO='U', ITYPE=1, JOBZ='V', RANGE='I'.  Pointers with the
 * suffix _d are device pointers on the current device, _h are host pointers.  Complex elements are
 * interleaved (re, im) doubles.  All functions return 0 on success and -1 on error (the reference's info
 * convention, zhegvdx_gpu.F90:106-127); eigb200_last_error() returns the message.
 * No function here has a CPU fallback: without a CUDA sm_100 device every compute entry fails with -1.
 */
#ifndef EIGB200_H
#define EIGB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- context: replaces module eigsolve_vars (eigsolve_vars.F90:25-61) ------------------------------ */
int eigb200_init(void);                         /* idempotent; init_eigsolve_gpu (eigsolve_vars.F90:39-59) */
int eigb200_finalize(void);
const char* eigb200_last_error(void);
int eigb200_set_stream(void* cuda_stream);      /* stream the hot path is issued on (default: legacy stream 0) */
/* one-shot: the next eigb200_{d,z}*gvdx call waits for this cudaEvent_t before it touches A (B is factored first), so
 * that a caller can upload A on a second stream while the Cholesky factorization of B runs; no reference analogue
 * (the reference's callers own the H2D copies, test_driver/test_zhegvdx.F90:266-293) */
int eigb200_set_a_ready_event(void* cuda_event);
int eigb200_version(void);
/* device scratch (bytes) the library allocates internally for order n (it never asks the caller for more
 * device workspace than the reference minima) */
int64_t eigb200_scratch_bytes(int n, int is_complex);
/* tunables (hard-coded in the reference: zhegvdx_gpu.F90:156, zheevd_gpu.F90:63-64); name/value pairs:
 * "trd_nb", "bt_nb", "symv_tma", "trd_coop", "verbose" */
int eigb200_set_option(const char* name, int value);
int eigb200_get_option(const char* name);

/* ---- multi-GPU (one process per GPU): 1-D block-cyclic distribution of the trailing matrix inside ?sytrd/?hetrd ----
 * Caller-owned communicator (the Python test harness): mg_alloc: cudaMalloc `bytes` and export a 64-byte CUDA IPC handle;
 * mg_open: map a peer's allocation; mg_config: rank/world (world <= 8), the exchange buffers of ALL ranks (own entry = the
 * local pointer; this rank's buffer is filled with 0xFF: every exchanged word starts "unset" and validates itself, see
 * csrc/sytrd.cu), the size of one exchange buffer (>= world*2*(n+2) complex elements) and a callback
 * `void hook(int i0, int nbp, int owner)` that must broadcast columns [i0, i0+nbp) of A from rank `owner` to all ranks on
 * the library's stream.  The flag arrays are no longer used (kept in the signature; may be NULL).  The caller synchronises
 * the ranks between mg_config and the first solve.  world == 1 switches the mode off. */
int eigb200_mg_alloc(long long bytes, void** dptr, char* handle64);
int eigb200_mg_open(const char* handle64, void** dptr);
int eigb200_mg_config(int rank, int world, void** wbufs, void** flags, long long wbuf_bytes, void* panel_hook);
/* size of one rank's (unused, see above) flag array for order n; exchange buffers are always sized for complex elements:
 * wbuf_bytes = world * 2 * (n + 64) * 16 */
int eigb200_mg_flag_bytes(int n, int world);
/* Library-owned communicator (what a Fortran/MPI caller uses; the functions above remain for callers that own the
 * communicator themselves).  Rendezvous: rank 0 calls eigb200_mg_unique_id (ncclGetUniqueId), the caller distributes the
 * 128 bytes (MPI_Bcast, torch.distributed, a file ...), then EVERY rank calls eigb200_mg_init with the device it will
 * compute on current.  NCCL is resolved at run time (dlopen of libnccl.so.2); world <= 8 ranks on one NVLink domain.
 * mg_allgather_columns: every rank owns the contiguous column block eigb200_mg_column_range(ncols, world, rank) (64-column
 * aligned) of the ld x ncols device matrix; afterwards all ranks hold all blocks -- e.g. to upload 1/world of A per rank
 * over PCIe and assemble over NVLink. */
int eigb200_mg_unique_id(char* id128);
int eigb200_mg_init(int rank, int world, const char* id128);
int eigb200_mg_finalize(void);
int eigb200_mg_allgather_columns(void* M_d, int ld, int ncols, int elem_bytes);
int eigb200_mg_column_range(int ncols, int world, int rank, int* c0, int* c1);

/* optional profiling: CUDA-event timing per stage category and a count of the kernels this library launched.
 * categories (index into ms[8], cnt[8]): 0 potrf, 1 hegst, 2 hetrd panel kernel, 3 hetrd rank-2k update,
 * 4 stedc, 5 back-transformation, 6 final trsm, 7 other.  collect() synchronises the device. */
int eigb200_prof_enable(int on);
int eigb200_prof_reset(void);
int eigb200_prof_collect(double* ms, int* cnt, long long* launches);

/* live peak probes for the roofline denominators (about 0.3 s): out4[0] FP64 tensor (DMMA m8n8k4) TFLOP/s, out4[1] FP64 FMA
 * TFLOP/s, out4[2] HBM read-only GB/s, out4[3] HBM copy GB/s (read + write bytes); no reference analogue */
int eigb200_probe_peaks(double* out4);

/* profiling aid: with option "trd_trace"=1 the tridiagonalization records 16 globaltimer stamps (ns) per column
 * (slots 0-4: phase A start/end, after barrier, phase B end, after barrier; 5-13: finer steps inside the phases,
 * see tools/trace_hetrd.py); read them back here. Returns the count. */
long long eigb200_trace_read(unsigned long long* out, long long max_count);

/* ---- generalized drivers (the drop-in entry points) ------------------------------------------------ */
/* dsygvdx_gpu(N,A,lda,B,ldb,Z,ldz,il,iu,w,work,lwork,work_h,lwork_h,iwork_h,liwork_h,Z_h,ldz_h,w_h,info,
 *             _skip_host_copy)                                       dsygvdx_gpu.F90:71-72 */
int eigb200_dsygvdx(int n, double* A_d, int lda, double* B_d, int ldb, double* Z_d, int ldz, int il, int iu,
                    double* w_d, double* work_d, int lwork, double* work_h, int lwork_h, int* iwork_h,
                    int liwork_h, double* Z_h, int ldz_h, double* w_h, int* info, int skip_host_copy);
/* zhegvdx_gpu(N,A,lda,B,ldb,Z,ldz,il,iu,w,work,lwork,rwork,lrwork,work_h,lwork_h,rwork_h,lrwork_h,iwork_h,
 *             liwork_h,Z_h,ldz_h,w_h,info,_skip_host_copy)           zhegvdx_gpu.F90:75-76 */
int eigb200_zhegvdx(int n, void* A_d, int lda, void* B_d, int ldb, void* Z_d, int ldz, int il, int iu,
                    double* w_d, void* work_d, int lwork, double* rwork_d, int lrwork, void* work_h,
                    int lwork_h, double* rwork_h, int lrwork_h, int* iwork_h, int liwork_h, void* Z_h,
                    int ldz_h, double* w_h, int* info, int skip_host_copy);

/* ---- multi-GPU generalized drivers: ONE problem over the ranks of eigb200_mg_init (no reference analogue: the reference
 * is single-GPU).  Same argument lists and checks as above; collective: every rank calls with the SAME A, B
 * (replicated device inputs) and its own buffers.  On exit on every rank: B <- U, w(1:N), Z(:,1:m) (column blocks gathered
 * unless option "mg_gather_z" = 0), host copies as requested; A is destroyed in BOTH triangles (Z serves as workspace).
 * Partition: hegst / back-transform / final solve by right-hand-side columns, hetrd trailing matrix 1-D block-cyclic with the
 * per-column exchange inside the panel kernel over NVLink (peer-mapped buffers, no NCCL call per column), the root merge of
 * the divide & conquer by eigenvector columns, potrf by block columns from order 20000 on (replicated below, like the lower
 * levels of the divide & conquer).  world == 1: the single-GPU driver. */
int eigb200_dsygvdx_mg(int n, double* A_d, int lda, double* B_d, int ldb, double* Z_d, int ldz, int il, int iu,
                       double* w_d, double* work_d, int lwork, double* work_h, int lwork_h, int* iwork_h,
                       int liwork_h, double* Z_h, int ldz_h, double* w_h, int* info, int skip_host_copy);
int eigb200_zhegvdx_mg(int n, void* A_d, int lda, void* B_d, int ldb, void* Z_d, int ldz, int il, int iu,
                       double* w_d, void* work_d, int lwork, double* rwork_d, int lrwork, void* work_h,
                       int lwork_h, double* rwork_h, int lrwork_h, int* iwork_h, int liwork_h, void* Z_h,
                       int ldz_h, double* w_h, int* info, int skip_host_copy);

/* ---- standard drivers: dsyevd_gpu.F90:32-33 / zheevd_gpu.F90:32-33 (jobz='V', uplo='U') ---------------- */
int eigb200_dsyevd(int il, int iu, int n, double* A_d, int lda, double* Z_d, int ldz, double* w_d,
                   double* work_d, int lwork, double* work_h, int lwork_h, int* iwork_h, int liwork_h,
                   double* Z_h, int ldz_h, double* w_h, int* info);
int eigb200_zheevd(int il, int iu, int n, void* A_d, int lda, void* Z_d, int ldz, double* w_d, void* work_d,
                   int lwork, double* rwork_d, int lrwork, void* work_h, int lwork_h, double* rwork_h,
                   int lrwork_h, int* iwork_h, int liwork_h, void* Z_h, int ldz_h, double* w_h, int* info);

/* ---- stage entry points (device pointers; used by the parity tests and by multi-GPU orchestration) -- */
/* Cholesky B = U^H U, upper (replaces cusolverDn?potrf, zhegvdx_gpu.F90:135).  info_h: 0 or index of the
 * first non-positive pivot (host int). */
int eigb200_dpotrf(int n, double* B_d, int ldb, int* info_h);
int eigb200_zpotrf(int n, void* B_d, int ldb, int* info_h);
/* A <- U^-H A U^-1, upper (dsygst_gpu.F90:31-98 / zhegst_gpu.F90:31-109) */
int eigb200_dsygst(int n, double* A_d, int lda, const double* U_d, int ldu);
int eigb200_zhegst(int n, void* A_d, int lda, const void* U_d, int ldu);
/* tridiagonalization, upper (dsytrd_gpu.F90:30-95 / zhetrd_gpu.F90:30-96): d(n), e(n-1), tau(n-1) on device */
int eigb200_dsytrd(int n, double* A_d, int lda, double* d_d, double* e_d, double* tau_d);
int eigb200_zhetrd(int n, void* A_d, int lda, double* d_d, double* e_d, void* tau_d);
/* y = A x, A symmetric/Hermitian, upper triangle read (dsymv_gpu.F90:33-150 / zhemv_gpu.F90:33-193) */
int eigb200_dsymv(int n, const double* A_d, int lda, const double* x_d, double* y_d);
int eigb200_zhemv(int n, const void* A_d, int lda, const void* x_d, void* y_d);
/* C(upper) -= A B^H + B A^H  (cublasDsyr2k / cublasZher2k call sites dsytrd_gpu.F90:66 / zhetrd_gpu.F90:67) */
int eigb200_dsyr2k(int n, int k, double alpha, const double* A_d, int lda, const double* B_d, int ldb,
                   double beta, double* C_d, int ldc);
int eigb200_zher2k(int n, int k, double alpha, const void* A_d, int lda, const void* B_d, int ldb, double beta,
                   void* C_d, int ldc);
/* C = alpha op(A) op(B) + beta C; transa/transb in 'N','T','C' */
int eigb200_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double* A_d, int lda,
                  const double* B_d, int ldb, double beta, double* C_d, int ldc);
int eigb200_zgemm(char transa, char transb, int m, int n, int k, double alpha, const void* A_d, int lda,
                  const void* B_d, int ldb, double beta, void* C_d, int ldc);
/* all eigenpairs of the symmetric tridiagonal (d,e) on the device (replaces host ?stedc('I'),
 * dsyevd_gpu.F90:99 / zheevd_gpu.F90:101): w ascending in d_d, Q (n x n, real) in Q_d */
int eigb200_dstedc(int n, double* d_d, double* e_d, double* Q_d, int ldq);
/* same, eigenVECTORS only for the sorted columns [c_lo, c_hi) (0-based, half open; all n eigenvalues are still returned):
 * the root merge of the divide & conquer skips the other columns, which are undefined on exit -- what the drivers use for
 * il..iu subsets and what splits the merge over the ranks in the multi-GPU drivers */
int eigb200_dstedc_range(int n, double* d_d, double* e_d, double* Q_d, int ldq, int c_lo, int c_hi);
/* Z <- Q Z with Q from ?sytrd/?hetrd (back-transformation, dsyevd_gpu.F90:117-128 / zheevd_gpu.F90:119-130) */
int eigb200_dormtr(int n, int m, const double* A_d, int lda, const double* tau_d, double* Z_d, int ldz);
int eigb200_zunmtr(int n, int m, const void* A_d, int lda, const void* tau_d, void* Z_d, int ldz);
/* triangular solves with upper U: side 'L' trans 'N': X = U^-1 B (cublas?trsm, zhegvdx_gpu.F90:169);
 * side 'L' trans 'C': X = U^-H B;  side 'R' trans 'N': X = B U^-1.  B is m x n, overwritten. */
int eigb200_dtrsm(char side, char trans, int m, int n, const double* U_d, int ldu, double* B_d, int ldb);
int eigb200_ztrsm(char side, char trans, int m, int n, const void* U_d, int ldu, void* B_d, int ldb);

#ifdef __cplusplus
}
#endif
#endif
