// eigb200 -- Cholesky factorization, triangular solves and the reduction to standard form, sm_100a.
//
// Replaces: cusolverDn?potrf (zhegvdx_gpu.F90:135 / dsygvdx_gpu.F90:121), the cuBLAS trsm/gemm/her2k chain
// of zhegst_gpu / dsygst_gpu (zhegst_gpu.F90:51-107 / dsygst_gpu.F90:51-96) and the final cublas?trsm
// (zhegvdx_gpu.F90:169 / dsygvdx_gpu.F90:155).
//
// Everything is built from two pieces: (1) 64x64 diagonal blocks are factored / inverted inside one CTA in
// shared memory, (2) all O(n^3) work is rank-64 updates on the DMMA GEMM kernel (gemm.cu).  A triangular
// solve with many right-hand sides is then "multiply by the inverted diagonal block, update the rest" --
// backward stable in the same sense as the blocked cuBLAS TRSM it replaces.
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"

namespace eigb200 {

namespace {

constexpr int NB = 64;     // diagonal block size

// inverse of an upper-triangular 64x64 block held in shared memory s (column-major, ld = NB+1), result in
// inv (same layout, strictly lower part zero).  One column per thread (threads 0..nb-1).
template <typename T>
__device__ __forceinline__ T recip_(T a);
template <> __device__ __forceinline__ double recip_<double>(double a) { return 1.0 / a; }
template <> __device__ __forceinline__ double2 recip_<double2>(double2 a) {
  double r, den;
  if (fabs(a.y) <= fabs(a.x)) { r = a.y / a.x; den = a.x + a.y * r; return mkz(1.0 / den, -r / den); }
  r = a.x / a.y; den = a.y + a.x * r; return mkz(r / den, -1.0 / den);
}

template <typename T>
__device__ void tri_inverse_smem(const T* s, T* inv, int nb) {
  const int j = threadIdx.x;
  if (j < nb) {
    for (int i = nb - 1; i > j; --i) inv[i + j * (NB + 1)] = zero_<T>();
    inv[j + j * (NB + 1)] = recip_<T>(s[j + j * (NB + 1)]);
    for (int i = j - 1; i >= 0; --i) {
      T acc = zero_<T>();
      for (int l = i + 1; l <= j; ++l) fma_(acc, s[i + l * (NB + 1)], inv[l + j * (NB + 1)]);
      inv[i + j * (NB + 1)] = neg_(mul_(acc, recip_<T>(s[i + i * (NB + 1)])));
    }
  }
}

// Dinv[b] = inverse of the b-th 64x64 diagonal block of the upper-triangular U (batched over blocks)
template <typename T>
__global__ void __launch_bounds__(NB) trtri_blocks_kernel(const T* __restrict__ U, int64_t ldu, int n, T* Dinv) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* s = reinterpret_cast<T*>(dyn_smem);
  T* inv = s + NB * (NB + 1);
  const int b = blockIdx.x, r0 = b * NB, nb = min(NB, n - r0);
  for (int c = 0; c < nb; ++c) {
    const int r = threadIdx.x;
    if (r < nb) s[r + c * (NB + 1)] = (r <= c) ? U[(r0 + r) + (int64_t)(r0 + c) * ldu] : zero_<T>();
  }
  __syncthreads();
  tri_inverse_smem<T>(s, inv, nb);
  __syncthreads();
  T* out = Dinv + (int64_t)b * NB * NB;
  for (int c = 0; c < NB; ++c) {
    const int r = threadIdx.x;
    out[r + c * NB] = (r < nb && c < nb) ? inv[r + c * (NB + 1)] : zero_<T>();
  }
}

// Cholesky of one 64x64 diagonal block (upper: A = U^H U) + its inverse.  info: first failing pivot (1-based,
// global index) via atomicMin-style CAS on *info (0 = ok).
template <typename T>
__global__ void __launch_bounds__(256) potf2_block_kernel(T* A, int64_t lda, int r0, int nb, T* Dinv, int* info) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* s = reinterpret_cast<T*>(dyn_smem);
  T* inv = s + NB * (NB + 1);
  __shared__ int bad;
  const int tid = threadIdx.x;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int r = idx % nb, c = idx / nb;
    s[r + c * (NB + 1)] = (r <= c) ? A[(r0 + r) + (int64_t)(r0 + c) * lda] : zero_<T>();
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double piv = real_(s[j + j * (NB + 1)]);
    if (!(piv > 0.0)) { if (tid == 0 && bad == 0) bad = r0 + j + 1; }
    const double rp = sqrt(piv), irp = 1.0 / rp;
    __syncthreads();
    // scale row j
    for (int c = j + tid; c < nb; c += blockDim.x)
      s[j + c * (NB + 1)] = (c == j) ? from_real<T>(rp) : scale_(s[j + c * (NB + 1)], irp);
    __syncthreads();
    // trailing update a(i,l) -= conj(u(j,i)) u(j,l), j < i <= l
    const int m = nb - j - 1;
    for (int idx = tid; idx < m * m; idx += blockDim.x) {
      const int i = j + 1 + idx % m, l = j + 1 + idx / m;
      if (i <= l) {
        T t = zero_<T>();
        fmac_(t, s[j + i * (NB + 1)], s[j + l * (NB + 1)]);
        s[i + l * (NB + 1)] = sub_(s[i + l * (NB + 1)], t);
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int r = idx % nb, c = idx / nb;
    if (r <= c) {
      T v = s[r + c * (NB + 1)];
      if (r == c) v = from_real<T>(real_(v));
      A[(r0 + r) + (int64_t)(r0 + c) * lda] = v;
    }
  }
  tri_inverse_smem<T>(s, inv, nb);
  __syncthreads();
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx % NB, c = idx / NB;
    Dinv[r + c * NB] = (r < nb && c < nb) ? inv[r + c * (NB + 1)] : zero_<T>();
  }
  if (tid == 0 && bad != 0) atomicCAS(info, 0, bad);
}

// In-place multiply of a 64-row (left) or 64-column (right) panel by a 64x64 block M (or M^H):
//  LEFT : B(r0:r0+nb, :)  <- op(M) * B(r0:r0+nb, :)      grid.x over column strips of 64
//  RIGHT: B(:, c0:c0+nb)  <- B(:, c0:c0+nb) * op(M)      grid.x over row strips of 64
template <typename T, bool LEFT, bool CONJT>
__global__ void __launch_bounds__(256) diag_mult_kernel(const T* __restrict__ M, T* B, int64_t ldb, int off, int nb,
                                                        int other_beg, int other_end) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* sm = reinterpret_cast<T*>(dyn_smem);
  T* sb = sm + NB * (NB + 1);
  const int tid = threadIdx.x;
  const int o0 = other_beg + blockIdx.x * NB;
  const int on = min(NB, other_end - o0);
  if (on <= 0) return;
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx % NB, c = idx / NB;
    // sm holds op(M): element (r, c)
    T v = CONJT ? conj_(M[c + r * NB]) : M[r + c * NB];
    sm[r + c * (NB + 1)] = v;
  }
  // sb(r, c): LEFT: rows r = panel row, c = strip column; RIGHT: r = strip row, c = panel column
  for (int idx = tid; idx < NB * NB; idx += blockDim.x) {
    const int r = idx % NB, c = idx / NB;
    T v = zero_<T>();
    if (LEFT) { if (r < nb && c < on) v = B[(off + r) + (int64_t)(o0 + c) * ldb]; }
    else      { if (r < on && c < nb) v = B[(o0 + r) + (int64_t)(off + c) * ldb]; }
    sb[r + c * (NB + 1)] = v;
  }
  __syncthreads();
  // each thread computes 16 outputs: rows r = tid % 64, columns c = tid/64 + 4*q
  const int r = tid % NB;
  T acc[NB / 4];
#pragma unroll
  for (int q = 0; q < NB / 4; ++q) acc[q] = zero_<T>();
  for (int l = 0; l < NB; ++l) {
    if (LEFT) {
      const T a = sm[r + l * (NB + 1)];
#pragma unroll
      for (int q = 0; q < NB / 4; ++q) fma_(acc[q], a, sb[l + (tid / NB + 4 * q) * (NB + 1)]);
    } else {
      const T a = sb[r + l * (NB + 1)];
#pragma unroll
      for (int q = 0; q < NB / 4; ++q) fma_(acc[q], a, sm[l + (tid / NB + 4 * q) * (NB + 1)]);
    }
  }
#pragma unroll
  for (int q = 0; q < NB / 4; ++q) {
    const int c = tid / NB + 4 * q;
    if (LEFT) { if (r < nb && c < on) B[(off + r) + (int64_t)(o0 + c) * ldb] = acc[q]; }
    else      { if (r < on && c < nb) B[(o0 + r) + (int64_t)(off + c) * ldb] = acc[q]; }
  }
}

template <typename T>
__global__ void symmetrize_kernel(T* A, int64_t lda, int n, T* save, int64_t lds) {
  // save strict lower triangle of A into `save` (if not null), then A(i,j) = conj(A(j,i)) for i > j
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n && i > j) {
    if (save) save[i + (int64_t)j * lds] = A[i + (int64_t)j * lda];
    A[i + (int64_t)j * lda] = conj_(A[j + (int64_t)i * lda]);
  } else if (i == j && i < n) {
    A[i + (int64_t)i * lda] = from_real<T>(real_(A[i + (int64_t)i * lda]));
  }
}
template <typename T>
__global__ void restore_lower_kernel(T* A, int64_t lda, int n, const T* save, int64_t lds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n && i > j) A[i + (int64_t)j * lda] = save[i + (int64_t)j * lds];
}

template <typename T> constexpr size_t blk_smem() { return 2 * (size_t)NB * (NB + 1) * sizeof(T); }

template <typename K>
int enable_smem(K kern, size_t bytes) {
  EIGB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
template <typename T>
int enable_all_smem() {
  static bool done = false;
  if (done) return 0;
  if (enable_smem(trtri_blocks_kernel<T>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(potf2_block_kernel<T>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, true>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, false>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, false, false>, blk_smem<T>()) != 0) return -1;
  done = true;
  return 0;
}

}  // namespace

template <typename T>
int symmetrize_from_upper(cudaStream_t s, int n, T* A, int64_t lda, T* save, int64_t lds) {
  if (n <= 0) return 0;
  symmetrize_kernel<T><<<dim3(cdiv(n, 256), n), 256, 0, s>>>(A, lda, n, save, lds);
  EIGB_LAUNCH_CHECK();
  return 0;
}
template <typename T>
int restore_lower(cudaStream_t s, int n, T* A, int64_t lda, const T* save, int64_t lds) {
  if (n <= 0) return 0;
  restore_lower_kernel<T><<<dim3(cdiv(n, 256), n), 256, 0, s>>>(A, lda, n, save, lds);
  EIGB_LAUNCH_CHECK();
  return 0;
}

// Cholesky B = U^H U (upper, in place).  *info_h = 0 or the 1-based index of the first non-positive pivot.
template <typename T>
int potrf_upper(cudaStream_t s, int n, T* B, int64_t ldb, int* info_h) {
  *info_h = 0;
  if (n <= 0) return 0;
  if (enable_all_smem<T>() != 0) return -1;
  Context& c = ctx();
  void* scr = ctx_scratch((size_t)NB * NB * sizeof(T) + 256);
  if (!scr) return -1;
  T* Dinv = (T*)scr;
  int* dinfo = c.d_info + 1;
  EIGB_CUDA_CHECK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  for (int k = 0; k < n; k += NB) {
    const int nb = n - k < NB ? n - k : NB;
    potf2_block_kernel<T><<<1, 256, blk_smem<T>(), s>>>(B, ldb, k, nb, Dinv, dinfo);
    EIGB_LAUNCH_CHECK();
    const int rest = n - k - nb;
    if (rest > 0) {
      // row panel: U(k, k+nb:) = U_kk^-H * A(k, k+nb:)
      diag_mult_kernel<T, true, true><<<cdiv(rest, NB), 256, blk_smem<T>(), s>>>(Dinv, B, ldb, k, nb, k + nb, n);
      EIGB_LAUNCH_CHECK();
      // trailing update A22 -= U12^H U12 (upper)
      T* U12 = B + k + (int64_t)(k + nb) * ldb;
      if (herk_upper<T>(s, 'C', rest, nb, -1.0, U12, ldb, 1.0, B + (k + nb) + (int64_t)(k + nb) * ldb, ldb) != 0)
        return -1;
    }
  }
  EIGB_CUDA_CHECK(cudaMemcpyAsync(info_h, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  return 0;
}

// Triangular solves with the upper-triangular U (n_u x n_u):
//   side 'L', trans 'N':  B (n_u x ncols) <- U^-1  B
//   side 'L', trans 'C':  B (n_u x ncols) <- U^-H  B
//   side 'R', trans 'N':  B (nrows x n_u) <- B U^-1
template <typename T>
int trsm_upper(cudaStream_t s, char side, char trans, int m, int n, const T* U, int64_t ldu, T* B, int64_t ldb) {
  if (m <= 0 || n <= 0) return 0;
  if (enable_all_smem<T>() != 0) return -1;
  const int nu = (side == 'L') ? m : n;
  const int nblk = cdiv(nu, NB);
  void* scr = ctx_scratch((size_t)nblk * NB * NB * sizeof(T) + 256);
  if (!scr) return -1;
  T* Dinv = (T*)scr;
  trtri_blocks_kernel<T><<<nblk, NB, blk_smem<T>(), s>>>(U, ldu, nu, Dinv);
  EIGB_LAUNCH_CHECK();
  if (side == 'L' && trans == 'N') {
    for (int b = nblk - 1; b >= 0; --b) {
      const int r0 = b * NB, nb = nu - r0 < NB ? nu - r0 : NB;
      diag_mult_kernel<T, true, false><<<cdiv(n, NB), 256, blk_smem<T>(), s>>>(Dinv + (int64_t)b * NB * NB, B, ldb, r0, nb, 0, n);
      EIGB_LAUNCH_CHECK();
      if (r0 > 0) {   // B(0:r0, :) -= U(0:r0, r0:r0+nb) * X_b
        if (gemm<T>(s, 'N', 'N', r0, n, nb, -1.0, U + (int64_t)r0 * ldu, ldu, B + r0, ldb, 1.0, B, ldb) != 0) return -1;
      }
    }
  } else if (side == 'L') {   // U^-H: forward
    for (int b = 0; b < nblk; ++b) {
      const int r0 = b * NB, nb = nu - r0 < NB ? nu - r0 : NB;
      diag_mult_kernel<T, true, true><<<cdiv(n, NB), 256, blk_smem<T>(), s>>>(Dinv + (int64_t)b * NB * NB, B, ldb, r0, nb, 0, n);
      EIGB_LAUNCH_CHECK();
      const int rest = nu - r0 - nb;
      if (rest > 0) {   // B(r0+nb:, :) -= U(r0:r0+nb, r0+nb:)^H * X_b
        if (gemm<T>(s, 'C', 'N', rest, n, nb, -1.0, U + r0 + (int64_t)(r0 + nb) * ldu, ldu, B + r0, ldb, 1.0,
                    B + r0 + nb, ldb) != 0) return -1;
      }
    }
  } else {                    // right, no-trans: forward over column blocks
    for (int b = 0; b < nblk; ++b) {
      const int c0 = b * NB, nb = nu - c0 < NB ? nu - c0 : NB;
      diag_mult_kernel<T, false, false><<<cdiv(m, NB), 256, blk_smem<T>(), s>>>(Dinv + (int64_t)b * NB * NB, B, ldb, c0, nb, 0, m);
      EIGB_LAUNCH_CHECK();
      const int rest = nu - c0 - nb;
      if (rest > 0) {   // B(:, c0+nb:) -= X_b * U(c0:c0+nb, c0+nb:)
        if (gemm<T>(s, 'N', 'N', m, rest, nb, -1.0, B + (int64_t)c0 * ldb, ldb, U + c0 + (int64_t)(c0 + nb) * ldu, ldu,
                    1.0, B + (int64_t)(c0 + nb) * ldb, ldb) != 0) return -1;
      }
    }
  }
  return 0;
}

// Reduction to standard form A <- U^-H A U^-1 (zhegst_gpu.F90:31-109 / dsygst_gpu.F90:31-98).
// A's upper triangle is read; on exit the full Hermitian result is stored (both triangles).  The caller's
// strict lower triangle of A is saved into `save` first when save != nullptr (the reference keeps it in Z,
// zhegvdx_gpu.F90:145-152).
template <typename T>
int hegst_upper(cudaStream_t s, int n, T* A, int64_t lda, const T* U, int64_t ldu, T* save, int64_t lds) {
  if (n <= 0) return 0;
  if (symmetrize_from_upper<T>(s, n, A, lda, save, lds) != 0) return -1;
  if (trsm_upper<T>(s, 'L', 'C', n, n, U, ldu, A, lda) != 0) return -1;
  if (trsm_upper<T>(s, 'R', 'N', n, n, U, ldu, A, lda) != 0) return -1;
  return 0;
}

#define EIGB_INST(T)                                                                                     \
  template int symmetrize_from_upper<T>(cudaStream_t, int, T*, int64_t, T*, int64_t);                    \
  template int restore_lower<T>(cudaStream_t, int, T*, int64_t, const T*, int64_t);                      \
  template int potrf_upper<T>(cudaStream_t, int, T*, int64_t, int*);                                     \
  template int trsm_upper<T>(cudaStream_t, char, char, int, int, const T*, int64_t, T*, int64_t);        \
  template int hegst_upper<T>(cudaStream_t, int, T*, int64_t, const T*, int64_t, T*, int64_t);
EIGB_INST(double)
EIGB_INST(double2)
#undef EIGB_INST

}  // namespace eigb200
