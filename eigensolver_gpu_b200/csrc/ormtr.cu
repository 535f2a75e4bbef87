// eigb200 -- eigenvector back-transformation Z <- Q Z, Q = H(n-1)...H(1) from the tridiagonalization, sm_100a.
//
// Replaces the ?ormtr/?unmtr loop of the reference: zheevd_gpu.F90:119-130 with zlarft_gpu (:136-176,
// finish_T_block_kernel :215-279) and zlarfb_gpu (:178-213); real: dsyevd_gpu.F90:117-128, 134-174, 176-210,
// 212-276.  Same compact-WY mathematics (Z <- (I - V T V^H) Z per block of reflectors, T lower triangular,
// LAPACK ?larft('Backward','Columnwise')), different schedule:
//   * block width bt_nb (default 128) instead of 64, so that the two big GEMMs run at K = 128 on the DMMA kernel;
//   * all unit-lower-trapezoidal V panels are materialised once in a packed workspace (the reference patches A
//     in place and restores it, K22/K28) and ALL block-reflector factors T are built up front by one batched
//     GEMM (V^H V) + one batched triangular-recurrence kernel -- no per-block host synchronisation
//     (the reference has a cudaStreamSynchronize per block, zheevd_gpu.F90:194).
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"
#include <vector>
#include <string.h>

namespace eigb200 {

namespace {

constexpr int BTMAX = 256;

// VW(:, j) for reflector j (0-based, j = 0..n-2): rows 0..j-1... stored exactly as LAPACK expects for the block:
// v_j has j+1 entries: A(0:j, j+1) with the unit element at row j; rows below are zero up to the block height.
template <typename T>
__global__ void bt_prepare_v_kernel(const T* __restrict__ A, int64_t lda, int n, int ib, T* VW, int64_t ldv) {
  const int j = blockIdx.y;                  // reflector index, 0..n-2
  if (j >= n - 1) return;
  const int blk = j / ib;
  const int jend = min((blk + 1) * ib, n - 1);   // one past the last reflector of this block
  const int mi = jend;                       // rows of the block = index of last reflector + 1
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= mi) return;
  T v;
  if (r < j) v = A[r + (int64_t)(j + 1) * lda];
  else if (r == j) v = from_real<T>(1.0);
  else v = zero_<T>();
  VW[r + (int64_t)j * ldv] = v;
}

// T = larft('B','C') from T0 = V^H V (full ib x ib block in Tm, ld = BTMAX) and tau.  One CTA per block; the
// block stays in global memory (L1/L2 resident), the column being formed is staged in shared memory.
template <typename T>
__global__ void __launch_bounds__(256) bt_finish_t_kernel(T* Tm_all, const T* __restrict__ tau, int n, int ib) {
  __shared__ T colbuf[BTMAX];
  const int blk = blockIdx.x;
  const int j0 = blk * ib;
  const int ibb = min(ib, n - 1 - j0);
  if (ibb <= 0) return;
  T* t = Tm_all + (int64_t)blk * BTMAX * BTMAX;
  const int ld = BTMAX;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < ibb * ibb; idx += blockDim.x) {
    const int r = idx % ibb, c = idx / ibb;
    T v = zero_<T>();
    if (r == c) v = tau[j0 + c];
    else if (r > c) v = neg_(mul_(tau[j0 + c], t[r + c * ld]));
    t[r + c * ld] = v;
  }
  __syncthreads();
  // backward recurrence: T(c+1:, c) <- T(c+1:, c+1:) * T(c+1:, c)   (lower triangular matvec)
  for (int c = ibb - 2; c >= 0; --c) {
    const int r = c + 1 + tid;
    if (r < ibb) colbuf[r] = t[r + c * ld];
    __syncthreads();
    if (r < ibb) {
      T acc = zero_<T>();
      for (int l = c + 1; l <= r; ++l) fma_(acc, t[r + l * ld], colbuf[l]);
      t[r + c * ld] = acc;
    }
    __syncthreads();
  }
}

// X(0:rows, 0:cols) = sum over the S slices (fixed order: deterministic) of a split-K product
template <typename T>
__global__ void bt_sum_slices_kernel(const T* __restrict__ Xs, int S, int64_t slice, T* X, int rows, int cols, int ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r >= rows || c >= cols) return;
  const int64_t o = r + (int64_t)c * ld;
  T acc = Xs[o];
  for (int q = 1; q < S; ++q) acc = add_(acc, Xs[o + q * slice]);
  X[o] = acc;
}

// V^H Z has only ib rows: its tiles seldom fill the 2 x 148 CTA slots of the GPU evenly (m = 8192, ib = 128: 256 tiles on
// 296 slots = one wave at 86 %; few eigenvector columns: a fraction of a wave).  Its K range (the block height) is
// therefore split over S CTA groups whose partial products are summed afterwards; S minimises the number of waves per
// unit of work, ceil(ntile S / slots) / S, with a small penalty per slice for the extra pass over the partial products.
int bt_splits(int m, int ib, int esize, int cap = 16) {
  int bm, bn;
  if (esize == 16) gemm_tile_dims<double2>(bm, bn); else gemm_tile_dims<double>(bm, bn);
  const int ntile = ((ib + bm - 1) / bm) * ((m + bn - 1) / bn);
  if (ntile <= 0) return 1;
  const int slots = 2 * ctx().num_sms;
  int best = 1;
  double best_cost = 1e30;
  for (int S = 1; S <= 16 && S <= cap; ++S) {
    const double waves = (double)(((long long)ntile * S + slots - 1) / slots) / S;
    const double cost = waves * (1.0 + 0.004 * (S - 1));
    if (cost < best_cost - 1e-12) { best_cost = cost; best = S; }
  }
  return best;
}

}  // namespace

size_t ormtr_scratch_bytes(int n, int m, int esize) {
  int ib = opts().bt_nb < BTMAX ? opts().bt_nb : BTMAX;
  size_t nblk = (size_t)(n > 1 ? (n - 2) / ib + 1 : 1);
  const size_t S = (size_t)bt_splits(m, ib, esize);
  return ((size_t)n * n + nblk * BTMAX * BTMAX + (2 + (S > 1 ? S : 0)) * (size_t)BTMAX * m) * esize +
         (nblk * (S + 1) + 8) * 256;
}

// Z(0:n, 0:m) <- Q Z.  A holds the reflectors (v_j in A(0:j, j+1), unit element explicit or not -- it is
// regenerated), tau(n-1).  `scratch` must provide ormtr_scratch_bytes(n, m, sizeof(T)).
template <typename T>
int ormtr_upper(cudaStream_t s, int n, int m, const T* A, int64_t lda, const T* tau, T* Z, int64_t ldz, void* scratch,
                size_t scratch_bytes) {
  if (n <= 1 || m <= 0) return 0;
  const int ib = opts().bt_nb < BTMAX ? opts().bt_nb : BTMAX;
  const int nref = n - 1;
  const int nblk = (nref + ib - 1) / ib;
  if (scratch_bytes < ormtr_scratch_bytes(n, m, sizeof(T))) { set_last_error("ormtr: scratch too small"); return -1; }
  ProfScope ps(PROF_ORMTR, s);
  Arena ar(scratch, scratch_bytes);
  const int64_t ldv = n;
  T* VW = ar.take<T>((size_t)n * n);
  T* Tm = ar.take<T>((size_t)nblk * BTMAX * BTMAX);
  T* X1 = ar.take<T>((size_t)BTMAX * m);
  T* X2 = ar.take<T>((size_t)BTMAX * m);
  const int S = bt_splits(m, ib, (int)sizeof(T));
  T* X1s = S > 1 ? ar.take<T>((size_t)S * BTMAX * m) : nullptr;
  GemmParams<T>* GP = ar.take<GemmParams<T>>(nblk);
  GemmParams<T>* GPS = S > 1 ? ar.take<GemmParams<T>>((size_t)nblk * S) : nullptr;
  if (!GP || (S > 1 && (!X1s || !GPS))) { set_last_error("ormtr: scratch arena exhausted"); return -1; }
  bt_prepare_v_kernel<T><<<dim3(cdiv(n, 256), nref), 256, 0, s>>>(A, lda, n, ib, VW, ldv);
  EIGB_LAUNCH_CHECK();
  // T0 = V^H V for every block (small outputs, long K): ONE batched launch, parameter blocks read from device memory
  {
    std::vector<GemmParams<T>> hp(nblk);
    for (int b = 0; b < nblk; ++b) {
      const int j0 = b * ib, ibb = (nref - j0 < ib) ? nref - j0 : ib, mi = j0 + ibb;
      GemmParams<T>& q = hp[b];
      memset(&q, 0, sizeof(q));
      q.M = ibb; q.N = ibb; q.nseg = 1;
      q.A[0] = VW + (int64_t)j0 * ldv; q.lda[0] = ldv; q.B[0] = q.A[0]; q.ldb[0] = ldv; q.K[0] = mi;
      q.A[1] = q.A[0]; q.B[1] = q.B[0]; q.lda[1] = ldv; q.ldb[1] = ldv; q.K[1] = 0;
      q.sa[0] = q.sa[1] = -1.0; q.sb[0] = q.sb[1] = 1.0;       // op(A) = V^H
      q.C = Tm + (int64_t)b * BTMAX * BTMAX; q.ldc = BTMAX;
      q.alpha = 1.0; q.beta = 0.0; q.mode = 0; q.real_diag = 0; q.colmap = nullptr;
    }
    EIGB_CUDA_CHECK(cudaMemcpyAsync(GP, hp.data(), sizeof(GemmParams<T>) * nblk, cudaMemcpyHostToDevice, s));
    GemmParams<T> dummy{};
    if (gemm_launch<T>(s, true, true, dummy, GP, nblk, ib, ib) != 0) return -1;
  }
  bt_finish_t_kernel<T><<<nblk, 256, 0, s>>>(Tm, tau, n, ib);
  EIGB_LAUNCH_CHECK();
  std::vector<int> nsplit(nblk, 1);
  if (S > 1) {
    // split-K parameter blocks of X1 = V^H Z for every reflector block, uploaded once
    std::vector<GemmParams<T>> hp((size_t)nblk * S);
    for (int b = 0; b < nblk; ++b) {
      const int j0 = b * ib, ibb = (nref - j0 < ib) ? nref - j0 : ib, mi = j0 + ibb;
      int Sb = bt_splits(m, ibb, (int)sizeof(T), mi / 256 < 1 ? 1 : mi / 256);      // slices of at least 256 rows
      if (Sb > S) Sb = S;
      const int kc = (((mi + Sb - 1) / Sb) + 15) & ~15;
      Sb = (mi + kc - 1) / kc;
      nsplit[b] = Sb;
      for (int q = 0; q < Sb; ++q) {
        GemmParams<T>& g = hp[(size_t)b * S + q];
        memset(&g, 0, sizeof(g));
        const int kbeg = q * kc, kk = (mi - kbeg < kc) ? mi - kbeg : kc;
        g.M = ibb; g.N = m; g.nseg = 1;
        g.A[0] = VW + (int64_t)j0 * ldv + kbeg; g.lda[0] = ldv; g.B[0] = Z + kbeg; g.ldb[0] = ldz; g.K[0] = kk;
        g.A[1] = g.A[0]; g.B[1] = g.B[0]; g.lda[1] = ldv; g.ldb[1] = ldz; g.K[1] = 0;
        g.sa[0] = g.sa[1] = -1.0; g.sb[0] = g.sb[1] = 1.0;     // op(A) = V^H
        g.C = X1s + (int64_t)q * BTMAX * m; g.ldc = BTMAX;
        g.alpha = 1.0; g.beta = 0.0; g.mode = 0; g.real_diag = 0; g.colmap = nullptr;
      }
    }
    EIGB_CUDA_CHECK(cudaMemcpyAsync(GPS, hp.data(), sizeof(GemmParams<T>) * hp.size(), cudaMemcpyHostToDevice, s));
  }
  // apply the blocks in ascending order: Z <- Z - V (T (V^H Z))
  for (int b = 0; b < nblk; ++b) {
    const int j0 = b * ib, ibb = (nref - j0 < ib) ? nref - j0 : ib, mi = j0 + ibb;
    const T* V = VW + (int64_t)j0 * ldv;
    if (nsplit[b] > 1) {
      GemmParams<T> dummy{};
      if (gemm_launch<T>(s, true, true, dummy, GPS + (size_t)b * S, nsplit[b], ibb, m) != 0) return -1;
      bt_sum_slices_kernel<T><<<dim3(cdiv(ibb, 128), m), 128, 0, s>>>(X1s, nsplit[b], (int64_t)BTMAX * m, X1, ibb, m, BTMAX);
      EIGB_LAUNCH_CHECK();
    } else if (gemm<T>(s, 'C', 'N', ibb, m, mi, 1.0, V, ldv, Z, ldz, 0.0, X1, BTMAX) != 0) return -1;
    if (gemm<T>(s, 'N', 'N', ibb, m, ibb, 1.0, Tm + (int64_t)b * BTMAX * BTMAX, BTMAX, X1, BTMAX, 0.0, X2, BTMAX) != 0)
      return -1;
    if (gemm<T>(s, 'N', 'N', mi, m, ibb, -1.0, V, ldv, X2, BTMAX, 1.0, Z, ldz) != 0) return -1;
  }
  return 0;
}

template int ormtr_upper<double>(cudaStream_t, int, int, const double*, int64_t, const double*, double*, int64_t, void*,
                                 size_t);
template int ormtr_upper<double2>(cudaStream_t, int, int, const double2*, int64_t, const double2*, double2*, int64_t,
                                  void*, size_t);

}  // namespace eigb200
