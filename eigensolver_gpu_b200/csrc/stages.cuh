// eigb200 -- internal stage interfaces (host-callable, all work issued on ctx().stream).
#pragma once
#include "common.cuh"

namespace eigb200 {

int set_option(const char* name, int value);
int get_option(const char* name);

struct Options {
  int trd_nb = 64;      // tridiagonalization panel width (reference: 32, zheevd_gpu.F90:63)
  int bt_nb = 128;      // back-transformation block (reference: 64, zheevd_gpu.F90:64)
  int symv_tma = 1;     // stage symv/hemv tiles through TMA (cp.async.bulk.tensor) when alignment allows
  int trd_coop = 1;     // persistent cooperative panel kernel (0: one launch per phase)
};
Options& opts();

// y = A x (A Hermitian, upper triangle read), deterministic tile reduction
template <typename T> int hemv_upper(cudaStream_t s, int n, const T* A, int64_t lda, const T* x, T* y);
// blocked tridiagonalization, UPLO='U'
template <typename T> int hetrd_upper(cudaStream_t s, int n, T* A, int64_t lda, double* d, double* e, T* tau);

// tridiagonal divide & conquer on the device
size_t stedc_scratch_bytes(int n);
int stedc_device(cudaStream_t s, int n, double* d, double* e, double* Q, int64_t ldq, void* scratch,
                 size_t scratch_bytes);

}  // namespace eigb200
