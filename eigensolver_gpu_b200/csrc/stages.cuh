// eigb200 -- internal stage interfaces (host-callable, all work issued on ctx().stream).
#pragma once
#include "common.cuh"

namespace eigb200 {

void prof_enable(int on);
void prof_reset();
void prof_collect(double* ms, int* cnt, long long* launches);
int set_option(const char* name, int value);
int get_option(const char* name);

struct Options {
  int trd_nb = 64;      // tridiagonalization panel width (reference: 32, zheevd_gpu.F90:63)
  int bt_nb = 128;      // back-transformation block (reference: 64, zheevd_gpu.F90:64)
  int symv_tma = 1;     // stage symv/hemv tiles through TMA (cp.async.bulk.tensor) when alignment allows
  int trd_coop = 1;     // persistent cooperative panel kernel (0: one launch per phase)
  int mg_switch_n = -1; // multi-GPU hetrd: below this trailing order all ranks continue replicated (-1: 3072 on 2 ranks, else 2048)
  int mg_dist_min_n = -1; // multi-GPU driver: distribute the tridiagonalization from this order on (-1: 6144 for 2 ranks, else 4096)
  int mg_potrf_min_n = 20000; // multi-GPU driver: distribute the Cholesky factorization from this order on (-1: always replicated;
                              // below, the look-ahead single-GPU factorization replicated on every rank is faster)
  int mg_gather_z = 1;    // multi-GPU driver: gather the eigenvector column blocks so that every rank holds Z(:, 1:m)
  int trd_upc = 3;      // tile engine: target number of tile units per CTA (strip length heuristic)
  int trd_ctab = 1;     // tile engine: strip length from the host-built table (replay of the unit queue); 0: closed-form heuristic
  int trd_prefetch = 0; // tiles per CTA prefetched into L2 during phase A (-1: 256 KB worth, 0: off -- no gain measured)
  int hegst_hb = 0;     // block size of the reduction to standard form (0: 2048 for n >= 4096, else 1024)
  int gemm_tma = 1;     // host-parameter GEMMs on the TMA-fed kernel (0: cp.async kernel everywhere)
  int gemm_tma_dbg = 0; // debugging aid for the TMA-fed kernel (bit 0: no early stage refill, bit 1: proxy fence before a stage release)
  int potrf_pb = 0;     // Cholesky: block-row height of the look-ahead variant (0: 192 complex / 256 real; < 0: plain recursion)
  int trsm_leaf256 = 1; // solves with a finished factor: 256x256 inverted diagonal blocks applied by the GEMM kernel
  int trd_l2keep_mb = 32; // tile engine: MB of the trailing matrix (its top tile rows) kept in L2 with evict_last; 0: no hints
  int nvtx = 0;         // 1: NVTX3 range per stage; 2: with a stream synchronisation at both ends (toolbox.F90:71-97)
  int trd_trace = 0;    // 1: per-column stamps; k > 1: also per-CTA begin/end stamps of phase B for the product of order k    // record per-column globaltimer stamps of the panel kernel (profiling aid)
};
Options& opts();

// multi-GPU configuration of the tridiagonalization (set through eigb200_mg_config)
typedef void (*panel_hook_t)(int i0, int nbp, int owner);
struct MgConfig {
  int rank = 0, P = 1;
  void* wbuf[8] = {nullptr};              // exchange buffers of all ranks (peer-mapped), wbuf[rank] is local
  unsigned long long* flags[8] = {nullptr};
  int64_t wbuf_bytes = 0;                 // size of ONE rank's exchange buffer
  int flag_stride = 0;                    // flags per source rank in a flag array: one per 32-row group + 1 (v^H A v)
  unsigned long long seq = 0;             // monotonic column sequence number (flags never reset)
  panel_hook_t hook = nullptr;            // legacy (caller-owned communicator): called before every panel to broadcast its columns
  void* comm = nullptr;                   // ncclComm_t created by eigb200_mg_init (library-owned communicator)
  bool own_exchange = false;              // exchange buffers allocated/mapped by mg_ensure_exchange (not handed in by mg_config)
  bool active = false;                    // distribute the NEXT tridiagonalization (set by the multi-GPU driver / mg_config)
};
MgConfig& mg();
// multi-GPU plumbing (mg.cu)
void mg_column_range(int ncols, int world, int rank, int& c0, int& c1);
int mg_flag_stride(int n);
int mg_unique_id(char* id128);
int mg_init(int rank, int world, const char* id128);
int mg_finalize();
int mg_ensure_exchange(cudaStream_t s, int n);
int mg_bcast_columns(cudaStream_t s, void* A, int64_t ld, int c0, int nc, int owner, int elem_bytes);
int mg_allgather_columns(cudaStream_t s, void* A, int64_t ld, int ncols, int elem_bytes);
int mg_bcast(cudaStream_t s, void* ptr, size_t bytes, int root);
int mg_group(bool start);
int mg_allreduce_max_int(cudaStream_t s, int* dptr);
}
#include <vector>
namespace eigb200 {
std::vector<unsigned long long>& trace_store();

// y = A x (A Hermitian, upper triangle read), deterministic tile reduction
template <typename T> int hemv_upper(cudaStream_t s, int n, const T* A, int64_t lda, const T* x, T* y);
// blocked tridiagonalization, UPLO='U'
// sync_status = false: the watchdog status word stays in ctx().d_info[ST_HETRD] for the caller's final status_fetch
template <typename T> int hetrd_upper(cudaStream_t s, int n, T* A, int64_t lda, double* d, double* e, T* tau, bool sync_status = true);
// copies the device status words to the pinned host mirror and synchronises the stream; then status_check() turns
// them into the reference's info convention (0 ok, -1 + message)
int status_fetch(cudaStream_t s);
// live peak probes (probe.cu): out[0] DMMA TFLOP/s, out[1] DFMA TFLOP/s, out[2] HBM read GB/s, out[3] HBM copy GB/s
int probe_peaks(cudaStream_t s, double* out);
int status_check(const char* who, int* potrf_pivot = nullptr);

// tridiagonal divide & conquer on the device
size_t stedc_scratch_bytes(int n);
int stedc_device(cudaStream_t s, int n, double* d, double* e, double* Q, int64_t ldq, void* scratch,
                 size_t scratch_bytes, int c_lo = 0, int c_hi = 2147483647);

// Cholesky / triangular solves / reduction to standard form (trsm.cu)
template <typename T> int symmetrize_from_upper(cudaStream_t s, int n, T* A, int64_t lda, T* save, int64_t lds);
template <typename T> int restore_lower(cudaStream_t s, int n, T* A, int64_t lda, const T* save, int64_t lds);
// sync_status = false: *info_h is not written, the pivot index stays in ctx().d_info[ST_POTRF]
template <typename T> int potrf_upper(cudaStream_t s, int n, T* B, int64_t ldb, int* info_h, bool sync_status = true);
// distributed variant (collective over the ranks of eigb200_mg_init): B replicated on entry, U replicated on exit, pivot
// status max-reduced into ctx().d_info[ST_POTRF] on every rank
template <typename T> int potrf_upper_mg(cudaStream_t s, int n, T* B, int64_t ldb);
template <typename T>
int trsm_upper(cudaStream_t s, char side, char trans, int m, int n, const T* U, int64_t ldu, T* B, int64_t ldb);
template <typename T>
int hegst_upper(cudaStream_t s, int n, T* A, int64_t lda, const T* U, int64_t ldu, T* save, int64_t lds);
// back-transformation (ormtr.cu)
size_t ormtr_scratch_bytes(int n, int m, int esize);
template <typename T>
int ormtr_upper(cudaStream_t s, int n, int m, const T* A, int64_t lda, const T* tau, T* Z, int64_t ldz, void* scratch,
                size_t scratch_bytes);
// drivers (driver.cu)
template <typename T>
int hegvdx_driver(int n, T* A, int lda, T* B, int ldb, T* Z, int ldz, int il, int iu, double* w, T* work, int lwork,
                  double* rwork, int lrwork, int lwork_h, int lrwork_h, int liwork_h, T* Z_h, int ldz_h, double* w_h,
                  int* info, int skip_host_copy);
template <typename T>
int hegvdx_check_args(int n, int lda, int ldb, int ldz, int il, int iu, int lwork, int lrwork, int lwork_h, int lrwork_h,
                      int liwork_h, int* info);
template <typename T>
int hegvdx_mg_driver(int n, T* A, int lda, T* B, int ldb, T* Z, int ldz, int il, int iu, double* w, T* work, int lwork,
                     double* rwork, int lrwork, int lwork_h, int lrwork_h, int liwork_h, T* Z_h, int ldz_h, double* w_h,
                     int* info, int skip_host_copy);
template <typename T>
int heevd_driver(int il, int iu, int n, T* A, int lda, T* Z, int ldz, double* w, T* work, int lwork, double* rwork,
                 int lrwork, T* Z_h, int ldz_h, double* w_h, int* info);

}  // namespace eigb200
