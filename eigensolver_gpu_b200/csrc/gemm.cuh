// eigb200 -- FP64 / complex-FP64 tensor-core GEMM family for sm_100a.
//
// One templated kernel serves every dense contraction on the hot path: syr2k/her2k trailing updates
// (reference: cublasZher2k/Dsyr2k call sites zhetrd_gpu.F90:67,82 / dsytrd_gpu.F90:66,81 and
// zhegst_gpu.F90:95-96), the GEMMs of hegst/potrf/trsm, the back-transformation GEMMs
// (zheevd_gpu.F90:193,201) and the divide-and-conquer merges.  FP64 tensor cores on sm_100a are reached
// through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no f64 kind), operands staged in shared memory by
// cp.async in a 3-stage ring.  Complex products are 4 real DMMAs on the (re, im) planes of one
// shared-memory fragment load.
#pragma once
#include "common.cuh"

namespace eigb200 {

template <typename T>
struct GemmParams {
  int M, N;
  int nseg;                 // 1, or 2 for rank-2k style  C = alpha*(A0*B0 + A1*B1) + beta*C
  const T* A[2]; int64_t lda[2];
  const T* B[2]; int64_t ldb[2];
  int K[2];
  double sa[2], sb[2];      // +1, or -1 to conjugate the operand (complex only)
  T* C; int64_t ldc;
  double alpha, beta;
  int mode;                 // 0: full C, 1: upper triangle of C only (row <= col)
  int real_diag;            // complex + mode 1: force Im C(i,i) = 0
  const int* colmap;        // optional: output column gn is stored at column colmap[gn] of C
  int diag_off;             // mode 1: element (gm, gn) is written iff gm <= gn + diag_off (C is a column block)
};

// op(A)(m,k): AK=false -> A[m + k*lda] ('N');  AK=true -> A[k + m*lda] ('T'/'C', conj via sa=-1)
// op(B)(k,n): BK=true  -> B[k + n*ldb] ('N');  BK=false -> B[n + k*ldb] ('T'/'C', conj via sb=-1)
// dev_params != nullptr: grid.z CTAs read their parameter block from device memory (sizes decided on device).
template <typename T>
int gemm_launch(cudaStream_t s, bool AK, bool BK, const GemmParams<T>& p, const GemmParams<T>* dev_params = nullptr,
                int batch = 1, int maxM = 0, int maxN = 0);

// TMA-fed variant for host-side parameters (gemm_tma.cu): 0 launched, 1 not applicable (use the cp.async kernel), -1 error
template <typename T>
int gemm_launch_tma(cudaStream_t s, bool AK, bool BK, const GemmParams<T>& p);

// CTA tile of the kernel for element type T (callers that need to know how many CTAs a shape produces)
template <typename T> void gemm_tile_dims(int& bm, int& bn);

// BLAS-like convenience: C = alpha*op(A)*op(B) + beta*C
template <typename T>
int gemm(cudaStream_t s, char ta, char tb, int M, int N, int K, double alpha, const T* A, int64_t lda, const T* B,
         int64_t ldb, double beta, T* C, int64_t ldc, int mode = 0);

// C(upper) = alpha*(A*B^H + B*A^H) + beta*C     trans='N': A,B are n x k
// C(upper) = alpha*(A^H*B + B^H*A) + beta*C     trans='C': A,B are k x n
template <typename T>
int her2k_upper(cudaStream_t s, char trans, int n, int k, double alpha, const T* A, int64_t lda, const T* B,
                int64_t ldb, double beta, T* C, int64_t ldc);

// C(upper) = alpha*A^H*A + beta*C (trans='C', A is k x n)  or alpha*A*A^H + beta*C (trans='N', A is n x k)
template <typename T>
int herk_upper(cudaStream_t s, char trans, int n, int k, double alpha, const T* A, int64_t lda, double beta, T* C,
               int64_t ldc);

}  // namespace eigb200
