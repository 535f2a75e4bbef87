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
#include <vector>
#include <string.h>

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

// Dinv[b] = inverse of the b-th 64x64 diagonal block of the upper-triangular U.  grid (blocks, 8), 256 threads:
// one warp per column j of the inverse, lanes = rows (lane, lane+32), column-oriented back substitution
// (x_l final -> rows i < l get  acc_i -= u(i,l) x_l): 64 dependent steps of one shuffle + two FMAs per lane
// instead of one thread walking an O(j^2) dependent chain.  b0 = first block handled by blockIdx.x == 0.
template <typename T>
__global__ void __launch_bounds__(256) trtri_blocks_kernel(const T* __restrict__ U, int64_t ldu, int n, T* Dinv, int b0) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* s = reinterpret_cast<T*>(dyn_smem);            // U block, column-major, ld = NB+1
  T* rinv = s + NB * (NB + 1);                      // 1 / u_ll
  const int b = b0 + blockIdx.x, r0 = b * NB, nb = min(NB, n - r0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx & (NB - 1), c = idx >> 6;
    T v = zero_<T>();
    if (r < nb && c < nb) { if (r <= c) v = U[(r0 + r) + (int64_t)(r0 + c) * ldu]; }
    else if (r == c) v = from_real<T>(1.0);         // identity padding of a partial last block
    s[r + c * (NB + 1)] = v;
  }
  __syncthreads();
  if (tid < NB) rinv[tid] = recip_<T>(s[tid + tid * (NB + 1)]);
  __syncthreads();
  const int j = blockIdx.y * 8 + warp;              // column of the inverse
  T acc0 = zero_<T>(), acc1 = zero_<T>(), x0 = zero_<T>(), x1 = zero_<T>();
  if (j == lane) acc0 = from_real<T>(1.0);
  if (j == lane + 32) acc1 = from_real<T>(1.0);
  for (int l = j; l >= 0; --l) {
    const T cand = (l >> 5) ? acc1 : acc0;
    T xl;
    if constexpr (is_cplx<T>::value) {
      xl = mkz(__shfl_sync(0xffffffffu, cand.x, l & 31), __shfl_sync(0xffffffffu, cand.y, l & 31));
    } else {
      xl = __shfl_sync(0xffffffffu, cand, l & 31);
    }
    xl = mul_(xl, rinv[l]);
    if (lane == (l & 31)) { if (l >> 5) x1 = xl; else x0 = xl; }
    const T* col = s + l * (NB + 1);
    if (lane < l) acc0 = sub_(acc0, mul_(col[lane], xl));
    if (lane + 32 < l) acc1 = sub_(acc1, mul_(col[lane + 32], xl));
  }
  T* out = Dinv + (int64_t)b * NB * NB + (int64_t)j * NB;
  out[lane] = (lane < nb && j < nb) ? x0 : zero_<T>();
  out[lane + 32] = (lane + 32 < nb && j < nb) ? x1 : zero_<T>();
}

// Cholesky of one 64x64 diagonal block (upper: A = U^H U).  info: first failing pivot (1-based, global index)
// via CAS on *info (0 = ok).  256 threads: thread (l = tid % 64, q = tid / 64) owns rows q, q+4, ... of column l.
template <typename T>
__global__ void __launch_bounds__(256) potf2_block_kernel(T* A, int64_t lda, int r0, int nb, int* info) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* s = reinterpret_cast<T*>(dyn_smem);
  __shared__ int bad;
  const int tid = threadIdx.x;
  const int l = tid & (NB - 1), q = tid >> 6;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx & (NB - 1), c = idx >> 6;
    s[r + c * (NB + 1)] = (r <= c && c < nb) ? A[(r0 + r) + (int64_t)(r0 + c) * lda] : zero_<T>();
  }
  __syncthreads();
  // outer-product Cholesky with ONE barrier per step: row j is used unscaled (a_il -= conj(a_ji) a_jl / a_jj) and the
  // 1/sqrt(pivot) scaling of the rows is applied in a single pass at the end (a_jj is final after step j)
  for (int j = 0; j < nb; ++j) {
    const double piv = real_(s[j + j * (NB + 1)]);
    if (!(piv > 0.0)) { if (tid == 0 && bad == 0) bad = r0 + j + 1; }
    if (l > j && l < nb) {
      const T f = scale_(s[j + l * (NB + 1)], 1.0 / piv);
      for (int i = j + 1 + q; i <= l; i += 4) {
        T t = zero_<T>();
        fmac_(t, s[j + i * (NB + 1)], f);
        s[i + l * (NB + 1)] = sub_(s[i + l * (NB + 1)], t);
      }
    }
    __syncthreads();
  }
  // row r scaled by 1/sqrt(a_rr); diagonal = sqrt(a_rr)
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx & (NB - 1), c = idx >> 6;
    if (r <= c && c < nb) {
      const double drr = real_(s[r + r * (NB + 1)]);
      const T v = (r == c) ? from_real<T>(sqrt(drr)) : scale_(s[r + c * (NB + 1)], 1.0 / sqrt(drr));
      A[(r0 + r) + (int64_t)(r0 + c) * lda] = v;
    }
  }
  if (tid == 0 && bad != 0) atomicCAS(info, 0, bad);
}

// In-place multiply of a 64-row (left) or 64-column (right) panel by a 64x64 block M (or M^H), one CTA per strip
// of SW columns (left) / rows (right) of the panel:
//  LEFT : B(off:off+nb, :)  <- op(M) * B(off:off+nb, :)
//  RIGHT: B(:, off:off+nb)  <- B(:, off:off+nb) * op(M)
// SW = 64 when the panel is long enough to fill the GPU, 16 otherwise (4x the CTAs: one SM needs ~8 us for a
// 64^3 complex product whatever the instruction mix, FP64 FMA and DMMA peak rates being equal on this part).
template <typename T, bool LEFT, bool CONJT, int SW>
__global__ void __launch_bounds__(256) diag_mult_kernel(const T* __restrict__ M, T* B, int64_t ldb, int off, int nb,
                                                        int other_beg, int other_end) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T* sm = reinterpret_cast<T*>(dyn_smem);
  T* sb = sm + NB * (NB + 1);
  constexpr int OPT = SW / 4;                       // outputs per thread
  constexpr int LDB_S = (LEFT ? NB : SW) + 1;       // sb: LEFT 64 x SW, RIGHT SW x 64 (column-major)
  const int tid = threadIdx.x;
  const int o0 = other_beg + blockIdx.x * SW;
  const int on = min(SW, other_end - o0);
  if (on <= 0) return;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int r = idx & (NB - 1), c = idx >> 6;
    sm[r + c * (NB + 1)] = CONJT ? conj_(M[c + r * NB]) : M[r + c * NB];     // element (r, c) of op(M)
  }
  if (LEFT) {
    for (int idx = tid; idx < NB * SW; idx += 256) {
      const int r = idx & (NB - 1), c = idx >> 6;
      sb[r + c * LDB_S] = (r < nb && c < on) ? B[(off + r) + (int64_t)(o0 + c) * ldb] : zero_<T>();
    }
  } else {
    for (int idx = tid; idx < NB * SW; idx += 256) {
      const int r = idx % SW, c = idx / SW;
      sb[r + c * LDB_S] = (r < on && c < nb) ? B[(o0 + r) + (int64_t)(off + c) * ldb] : zero_<T>();
    }
  }
  __syncthreads();
  T acc[OPT];
#pragma unroll
  for (int q = 0; q < OPT; ++q) acc[q] = zero_<T>();
  if (LEFT) {
    const int r = tid & (NB - 1), cg = tid >> 6;    // columns cg + 4q
    for (int l = 0; l < NB; ++l) {
      const T a = sm[r + l * (NB + 1)];
#pragma unroll
      for (int q = 0; q < OPT; ++q) fma_(acc[q], a, sb[l + (cg + 4 * q) * LDB_S]);
    }
#pragma unroll
    for (int q = 0; q < OPT; ++q) {
      const int c = cg + 4 * q;
      if (r < nb && c < on) B[(off + r) + (int64_t)(o0 + c) * ldb] = acc[q];
    }
  } else {
    constexpr int CG = 256 / SW;                    // column groups; columns cg + CG*q
    const int r = tid % SW, cg = tid / SW;
    for (int l = 0; l < NB; ++l) {
      const T a = sb[r + l * LDB_S];
#pragma unroll
      for (int q = 0; q < OPT; ++q) fma_(acc[q], a, sm[l + (cg + CG * q) * (NB + 1)]);
    }
#pragma unroll
    for (int q = 0; q < OPT; ++q) {
      const int c = cg + CG * q;
      if (r < on && c < nb) B[(o0 + r) + (int64_t)(off + c) * ldb] = acc[q];
    }
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
template <typename T> constexpr size_t tri_smem() { return ((size_t)NB * (NB + 1) + NB) * sizeof(T); }

template <typename K>
int enable_smem(K kern, size_t bytes) {
  EIGB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
template <typename T>
int enable_all_smem() {
  static OncePerDevice once;
  if (!once.need()) return 0;
  if (enable_smem(trtri_blocks_kernel<T>, tri_smem<T>()) != 0) return -1;
  if (enable_smem(potf2_block_kernel<T>, tri_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, true, 64>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, false, 64>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, false, false, 64>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, true, 16>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, true, false, 16>, blk_smem<T>()) != 0) return -1;
  if (enable_smem(diag_mult_kernel<T, false, false, 16>, blk_smem<T>()) != 0) return -1;
  once.done();
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

// ---- recursive building blocks --------------------------------------------------------------------------
// Splitting [lo, hi) near the middle turns almost all flops into GEMMs with a large inner dimension (half of
// them at K = n/2, a quarter at K = n/4, ...), which is where the DMMA kernel is efficient; only the leaves use
// inverted diagonal blocks: 64 x 64 ones (diag_mult_kernel) while a factorization is still in progress, 256 x 256
// ones applied by the GEMM kernel (out of place + copy back) once U is final -- a quarter of the leaves, and the
// launches below 256 (the latency-bound part of a solve) disappear.
constexpr int LB = 256;    // large leaf
template <typename T>
struct TriInv {
  const T* d64 = nullptr;      // [ceil(n/64)][64*64]
  const T* d256 = nullptr;     // [ceil(n/256)][256*256] or nullptr
  T* tmp = nullptr;            // out-of-place result of a 256-leaf
  size_t tmp_elems = 0;
};
template <typename T>
static inline int split_point(int lo, int hi, const TriInv<T>& ti) {
  if (ti.d256 && (lo % LB) == 0 && hi - lo > LB) {
    const int nb = (hi - lo + LB - 1) / LB;
    return lo + (nb / 2) * LB;
  }
  const int nb = (hi - lo + NB - 1) / NB;
  return lo + (nb / 2) * NB;
}

// Triangular solve restricted to the diagonal range [lo, hi) of U.  B is addressed with GLOBAL indices:
//   side 'L': rows [lo, hi) of B, `other` columns;   side 'R': columns [lo, hi) of B, `other` rows.
template <typename T>
int trsm_rec(cudaStream_t s, char side, char trans, int lo, int hi, int other, const T* U, int64_t ldu, T* B,
             int64_t ldb, const TriInv<T>& ti) {
  if (ti.d256 && (lo % LB) == 0 && hi - lo <= LB && hi - lo > NB && ti.tmp_elems >= (size_t)LB * 64) {
    const T* M = ti.d256 + (int64_t)(lo / LB) * LB * LB;
    const int nb = hi - lo;
    const int chunk = (int)((ti.tmp_elems / LB) < (size_t)other ? (ti.tmp_elems / LB) : (size_t)other);
    for (int o0 = 0; o0 < other; o0 += chunk) {
      const int oc = other - o0 < chunk ? other - o0 : chunk;
      if (side == 'L') {
        T* Bp = B + lo + (int64_t)o0 * ldb;                  // nb x oc
        if (gemm<T>(s, trans == 'N' ? 'N' : 'C', 'N', nb, oc, nb, 1.0, M, LB, Bp, ldb, 0.0, ti.tmp, LB) != 0) return -1;
        EIGB_CUDA_CHECK(cudaMemcpy2DAsync(Bp, (size_t)ldb * sizeof(T), ti.tmp, (size_t)LB * sizeof(T), (size_t)nb * sizeof(T),
                                          oc, cudaMemcpyDeviceToDevice, s));
      } else {
        T* Bp = B + o0 + (int64_t)lo * ldb;                  // oc x nb
        if (gemm<T>(s, 'N', 'N', oc, nb, nb, 1.0, Bp, ldb, M, LB, 0.0, ti.tmp, oc) != 0) return -1;
        EIGB_CUDA_CHECK(cudaMemcpy2DAsync(Bp, (size_t)ldb * sizeof(T), ti.tmp, (size_t)oc * sizeof(T), (size_t)oc * sizeof(T),
                                          nb, cudaMemcpyDeviceToDevice, s));
      }
    }
    return 0;
  }
  if (hi - lo <= NB) {
    const T* M = ti.d64 + (int64_t)(lo / NB) * NB * NB;
    const int nb = hi - lo;
    if (other >= 120 * NB) {
      if (side == 'L' && trans == 'N')
        diag_mult_kernel<T, true, false, 64><<<cdiv(other, 64), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
      else if (side == 'L')
        diag_mult_kernel<T, true, true, 64><<<cdiv(other, 64), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
      else
        diag_mult_kernel<T, false, false, 64><<<cdiv(other, 64), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
    } else {
      if (side == 'L' && trans == 'N')
        diag_mult_kernel<T, true, false, 16><<<cdiv(other, 16), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
      else if (side == 'L')
        diag_mult_kernel<T, true, true, 16><<<cdiv(other, 16), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
      else
        diag_mult_kernel<T, false, false, 16><<<cdiv(other, 16), 256, blk_smem<T>(), s>>>(M, B, ldb, lo, nb, 0, other);
    }
    EIGB_LAUNCH_CHECK();
    return 0;
  }
  const int mid = split_point<T>(lo, hi, ti);
  const int n1 = mid - lo, n2 = hi - mid;
  const T* U12 = U + lo + (int64_t)mid * ldu;          // n1 x n2
  if (side == 'L' && trans == 'N') {                   // X2 = U22^-1 B2 ; B1 -= U12 X2 ; X1 = U11^-1 B1
    if (trsm_rec<T>(s, side, trans, mid, hi, other, U, ldu, B, ldb, ti) != 0) return -1;
    if (gemm<T>(s, 'N', 'N', n1, other, n2, -1.0, U12, ldu, B + mid, ldb, 1.0, B + lo, ldb) != 0) return -1;
    return trsm_rec<T>(s, side, trans, lo, mid, other, U, ldu, B, ldb, ti);
  } else if (side == 'L') {                            // X1 = U11^-H B1 ; B2 -= U12^H X1 ; X2 = U22^-H B2
    if (trsm_rec<T>(s, side, trans, lo, mid, other, U, ldu, B, ldb, ti) != 0) return -1;
    if (gemm<T>(s, 'C', 'N', n2, other, n1, -1.0, U12, ldu, B + lo, ldb, 1.0, B + mid, ldb) != 0) return -1;
    return trsm_rec<T>(s, side, trans, mid, hi, other, U, ldu, B, ldb, ti);
  } else {                                             // X1 = B1 U11^-1 ; B2 -= X1 U12 ; X2 = B2 U22^-1
    if (trsm_rec<T>(s, side, trans, lo, mid, other, U, ldu, B, ldb, ti) != 0) return -1;
    if (gemm<T>(s, 'N', 'N', other, n2, n1, -1.0, B + (int64_t)lo * ldb, ldb, U12, ldu, 1.0, B + (int64_t)mid * ldb, ldb)
        != 0) return -1;
    return trsm_rec<T>(s, side, trans, mid, hi, other, U, ldu, B, ldb, ti);
  }
}

// 64x64 inverse blocks -> the block diagonal of the 256x256 inverse blocks (everything else zero)
template <typename T>
__global__ void dinv_place_kernel(const T* __restrict__ d64, int nblk64, T* d256) {
  const int b = blockIdx.x;                         // 64-block
  if (b >= nblk64) return;
  T* dst = d256 + (int64_t)(b / 4) * LB * LB + (int64_t)(b % 4) * NB * (LB + 1);
  const T* src = d64 + (int64_t)b * NB * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) dst[(idx & (NB - 1)) + (int64_t)(idx >> 6) * LB] = src[idx];
}

// Scratch (in elements of T unless noted) of a solve with a finished factor of order nu and `other` right-hand sides
template <typename T>
size_t triinv_scratch_bytes(int nu, int other) {
  const size_t nb64 = (size_t)cdiv(nu, NB), nb256 = (size_t)cdiv(nu, LB);
  size_t tmp = (size_t)LB * (size_t)(other < 64 ? 64 : other);
  const size_t cap = (size_t)LB * 16384;             // 256 x 16384 elements at most, larger panels go in chunks
  if (tmp > cap) tmp = cap;
  return (nb64 * NB * NB + nb256 * LB * LB + nb256 * 2 * 128 * 128 + tmp) * sizeof(T) + (4 * nb256 + 8) * sizeof(GemmParams<T>) + 4096;
}

// Inverses of the diagonal blocks of the finished factor U: 64x64 by substitution (trtri_blocks_kernel), then two
// batched merge levels  inv([A B; 0 C]) = [A^-1, -A^-1 B C^-1; 0, C^-1]  (4 device-parameter GEMM launches for the
// whole matrix) up to 256x256.  Carves everything from the context scratch.
template <typename T>
int build_triinv(cudaStream_t s, int nu, int other, const T* U, int64_t ldu, TriInv<T>& ti) {
  const int nb64 = cdiv(nu, NB), nb256 = cdiv(nu, LB);
  const bool big = opts().trsm_leaf256 != 0 && nu > LB;
  void* scr = ctx_scratch(big ? triinv_scratch_bytes<T>(nu, other) : (size_t)nb64 * NB * NB * sizeof(T) + 256);
  if (!scr) return -1;
  Arena ar(scr, ctx().scratch_bytes);
  T* d64 = ar.take<T>((size_t)nb64 * NB * NB);
  trtri_blocks_kernel<T><<<dim3(nb64, 8), 256, tri_smem<T>(), s>>>(U, ldu, nu, d64, 0);
  EIGB_LAUNCH_CHECK();
  ti.d64 = d64;
  if (!big) return 0;
  T* d256 = ar.take<T>((size_t)nb256 * LB * LB);
  T* t1 = ar.take<T>((size_t)nb256 * 2 * 128 * 128);
  size_t tmp = (size_t)LB * (size_t)(other < 64 ? 64 : other);
  if (tmp > (size_t)LB * 16384) tmp = (size_t)LB * 16384;
  T* tmpb = ar.take<T>(tmp);
  GemmParams<T>* GP = ar.take<GemmParams<T>>((size_t)4 * nb256 + 8);
  if (!GP) { set_last_error("trsm: scratch arena too small"); return -1; }
  EIGB_CUDA_CHECK(cudaMemsetAsync(d256, 0, (size_t)nb256 * LB * LB * sizeof(T), s));
  dinv_place_kernel<T><<<nb64, 256, 0, s>>>(d64, nb64, d256);
  EIGB_LAUNCH_CHECK();
  // parameter blocks: [0, n1) level-1 first products, [n1, 2 n1) level-1 second products, then level 2 likewise
  std::vector<GemmParams<T>> hp;
  auto mk = [&](int M, int N, int K, const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, double alpha) {
    GemmParams<T> g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.nseg = 1;
    g.A[0] = A; g.lda[0] = lda; g.B[0] = B; g.ldb[0] = ldb; g.K[0] = K;
    g.A[1] = A; g.lda[1] = lda; g.B[1] = B; g.ldb[1] = ldb; g.K[1] = 0;
    g.sa[0] = g.sa[1] = 1.0; g.sb[0] = g.sb[1] = 1.0;
    g.C = C; g.ldc = ldc; g.alpha = alpha; g.beta = 0.0; g.mode = 0; g.real_diag = 0; g.colmap = nullptr;
    hp.push_back(g);
  };
  // level 1: 64-block pairs (2p, 2p+1) inside a 256 block
  int cnt1 = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int p2 = 0; 2 * p2 + 1 < nb64; ++p2) {
      const int b0 = 2 * p2, b1 = b0 + 1;
      const int w = (nu - b1 * NB) < NB ? nu - b1 * NB : NB;          // width of the second block
      T* blk = d256 + (int64_t)(b0 / 4) * LB * LB;
      const int o0 = (b0 % 4) * NB, o1 = o0 + NB;                      // offsets inside the 256 block
      T* t = t1 + (int64_t)p2 * NB * NB;
      if (pass == 0) mk(NB, w, NB, blk + o0 + (int64_t)o0 * LB, LB, U + b0 * NB + (int64_t)b1 * NB * ldu, ldu, t, NB, 1.0);
      else           mk(NB, w, w, t, NB, blk + o1 + (int64_t)o1 * LB, LB, blk + o0 + (int64_t)o1 * LB, LB, -1.0);
      if (pass == 0) ++cnt1;
    }
  }
  // level 2: halves of a 256 block
  int cnt2 = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int q = 0; q < nb256; ++q) {
      const int r0 = q * LB, c0 = r0 + 128;
      if (c0 >= nu) continue;
      const int w = (nu - c0) < 128 ? nu - c0 : 128;
      T* blk = d256 + (int64_t)q * LB * LB;
      T* t = t1 + (int64_t)q * 128 * 128;
      if (pass == 0) mk(128, w, 128, blk, LB, U + r0 + (int64_t)c0 * ldu, ldu, t, 128, 1.0);
      else           mk(128, w, w, t, 128, blk + 128 + (int64_t)128 * LB, LB, blk + (int64_t)128 * LB, LB, -1.0);
      if (pass == 0) ++cnt2;
    }
  }
  if (!hp.empty()) {
    EIGB_CUDA_CHECK(cudaMemcpyAsync(GP, hp.data(), sizeof(GemmParams<T>) * hp.size(), cudaMemcpyHostToDevice, s));
    GemmParams<T> dummy{};
    if (cnt1 > 0) {
      if (gemm_launch<T>(s, false, true, dummy, GP, cnt1, NB, NB) != 0) return -1;
      if (gemm_launch<T>(s, false, true, dummy, GP + cnt1, cnt1, NB, NB) != 0) return -1;
    }
    if (cnt2 > 0) {
      if (gemm_launch<T>(s, false, true, dummy, GP + 2 * cnt1, cnt2, 128, 128) != 0) return -1;
      if (gemm_launch<T>(s, false, true, dummy, GP + 2 * cnt1 + cnt2, cnt2, 128, 128) != 0) return -1;
    }
  }
  ti.d256 = d256; ti.tmp = tmpb; ti.tmp_elems = tmp;
  return 0;
}

template <typename T>
int potrf_rec(cudaStream_t s, int lo, int hi, T* A, int64_t lda, T* Dinv, int* dinfo) {
  TriInv<T> ti; ti.d64 = Dinv;
  if (hi - lo <= NB) {
    potf2_block_kernel<T><<<1, 256, tri_smem<T>(), s>>>(A, lda, lo, hi - lo, dinfo);
    EIGB_LAUNCH_CHECK();
    trtri_blocks_kernel<T><<<dim3(1, 8), 256, tri_smem<T>(), s>>>(A, lda, hi, Dinv, lo / NB);
    EIGB_LAUNCH_CHECK();
    return 0;
  }
  const int mid = split_point<T>(lo, hi, ti);
  if (potrf_rec<T>(s, lo, mid, A, lda, Dinv, dinfo) != 0) return -1;
  // U12 = U11^-H A12 : rows [lo, mid) of the column block [mid, hi)
  T* A12cols = A + (int64_t)mid * lda;
  if (trsm_rec<T>(s, 'L', 'C', lo, mid, hi - mid, A, lda, A12cols, lda, ti) != 0) return -1;
  // A22 -= U12^H U12 (upper)
  if (herk_upper<T>(s, 'C', hi - mid, mid - lo, -1.0, A + lo + (int64_t)mid * lda, lda, 1.0,
                    A + mid + (int64_t)mid * lda, lda) != 0) return -1;
  return potrf_rec<T>(s, mid, hi, A, lda, Dinv, dinfo);
}

// Cholesky B = U^H U (upper, in place).  *info_h = 0 or the 1-based index of the first non-positive pivot.
//
// Right-looking by PB-wide block rows with ONE block of look-ahead on a high-priority side stream: the latency-bound
// part of step k+1 (diagonal block: recursive, tiny grids; row panel: solves with 64x64 inverted blocks) runs next to the
// bulk rank-PB update of step k, which fills the GPU on the caller's stream.  Below 2 PB the plain recursion is used.
//   hi stream : potrf(D_k) -> U_k* = D_k^-H A_k* -> update of the NEXT block row  A_(k+1)* -= U_k,(k+1)^H U_k*
//   main      : A_** -= U_k*^H U_k*  on the remaining trailing matrix (upper)
// (the reference calls cusolverDn?potrf here, zhegvdx_gpu.F90:135)
template <typename T>
int potrf_upper(cudaStream_t s, int n, T* B, int64_t ldb, int* info_h, bool sync_status) {
  if (sync_status) *info_h = 0;
  if (n <= 0) return 0;
  if (enable_all_smem<T>() != 0) return -1;
  Context& c = ctx();
  const int nblk = cdiv(n, NB);
  void* scr = ctx_scratch((size_t)nblk * NB * NB * sizeof(T) + 256);
  if (!scr) return -1;
  T* Dinv = (T*)scr;
  int* dinfo = c.d_info + ST_POTRF;
  EIGB_CUDA_CHECK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  const int PB = opts().potrf_pb > 0 ? ((opts().potrf_pb + NB - 1) / NB) * NB : (is_cplx<T>::value ? 192 : 256);
  if (n < 2 * PB || c.stream_hi == nullptr || opts().potrf_pb < 0) {
    if (potrf_rec<T>(s, 0, n, B, ldb, Dinv, dinfo) != 0) return -1;
  } else {
    cudaStream_t sh = c.stream_hi;
    TriInv<T> ti; ti.d64 = Dinv;
    std::vector<cudaEvent_t> evs;
    auto new_event = [&]() { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); return e; };
    int rc = 0;
    cudaEvent_t e0 = new_event();
    if (cudaEventRecord(e0, s) != cudaSuccess || cudaStreamWaitEvent(sh, e0, 0) != cudaSuccess) rc = -1;
    cudaEvent_t ev_bulk_prev = nullptr;
    for (int k0 = 0; k0 < n && rc == 0; k0 += PB) {
      const int kb = n - k0 < PB ? n - k0 : PB;
      const int r = n - k0 - kb;
      // diagonal block and row panel on the high-priority stream (their inputs were updated by the look-ahead step)
      if (potrf_rec<T>(sh, k0, k0 + kb, B, ldb, Dinv, dinfo) != 0) { rc = -1; break; }
      if (r <= 0) break;
      T* U12 = B + k0 + (int64_t)(k0 + kb) * ldb;                  // kb x r
      if (trsm_rec<T>(sh, 'L', 'C', k0, k0 + kb, r, B, ldb, B + (int64_t)(k0 + kb) * ldb, ldb, ti) != 0) { rc = -1; break; }
      cudaEvent_t ev_panel = new_event();
      if (cudaEventRecord(ev_panel, sh) != cudaSuccess) { rc = -1; break; }
      const int nb2 = r < PB ? r : PB;
      T* A22 = B + (k0 + kb) + (int64_t)(k0 + kb) * ldb;           // r x r trailing matrix (upper)
      // look-ahead: the next block row  A22(0:nb2, 0:r) -= U12(:, 0:nb2)^H U12(:, 0:r)  (upper part of its diagonal block).
      // Rows of the next block were last written by the bulk update of the previous step: wait for it.
      if (ev_bulk_prev && cudaStreamWaitEvent(sh, ev_bulk_prev, 0) != cudaSuccess) { rc = -1; break; }
      {
        GemmParams<T> p{};
        p.M = nb2; p.N = r; p.nseg = 1;
        p.A[0] = U12; p.lda[0] = ldb; p.B[0] = U12; p.ldb[0] = ldb; p.K[0] = kb;
        p.A[1] = U12; p.lda[1] = ldb; p.B[1] = U12; p.ldb[1] = ldb; p.K[1] = 0;
        p.sa[0] = p.sa[1] = -1.0; p.sb[0] = p.sb[1] = 1.0;
        p.C = A22; p.ldc = ldb; p.alpha = -1.0; p.beta = 1.0; p.mode = 1; p.real_diag = 1; p.colmap = nullptr; p.diag_off = 0;
        if (gemm_launch<T>(sh, true, true, p) != 0) { rc = -1; break; }
      }
      // bulk: the rest of the trailing matrix on the caller's stream, overlapping with the next diagonal block / panel
      if (r > nb2) {
        if (cudaStreamWaitEvent(s, ev_panel, 0) != cudaSuccess) { rc = -1; break; }
        if (herk_upper<T>(s, 'C', r - nb2, kb, -1.0, U12 + (int64_t)nb2 * ldb, ldb, 1.0, A22 + nb2 + (int64_t)nb2 * ldb, ldb) != 0) {
          rc = -1; break;
        }
        ev_bulk_prev = new_event();
        if (cudaEventRecord(ev_bulk_prev, s) != cudaSuccess) { rc = -1; break; }
      } else {
        ev_bulk_prev = nullptr;
      }
    }
    // join: everything issued on the side stream is ordered before what the caller's stream does next
    cudaEvent_t ej = new_event();
    if (cudaEventRecord(ej, sh) != cudaSuccess || cudaStreamWaitEvent(s, ej, 0) != cudaSuccess) rc = -1;
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (rc != 0) { if (cudaGetLastError() != cudaSuccess) set_last_error("potrf: stream/event call failed"); return -1; }
  }
  if (sync_status) {
    EIGB_CUDA_CHECK(cudaMemcpyAsync(info_h, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
    EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return 0;
}

// pack / unpack of a row block of a column-major matrix:  P(0:kb, c) <-> A(r0:r0+kb, c0 + c),  c in [0, nc)
template <typename T>
__global__ void rowblock_copy_kernel(T* A, int64_t lda, int r0, int kb, int c0, int nc, T* P, int64_t ldp, int to_matrix) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= kb || c >= nc) return;
  T* a = A + (r0 + r) + (int64_t)(c0 + c) * lda;
  T* q = P + r + (int64_t)c * ldp;
  if (to_matrix) *a = *q; else *q = *a;
}

// Distributed Cholesky (multi-GPU drivers): right-looking by MB-wide block columns dealt cyclically over the ranks.
//   step k: owner factors the diagonal block and broadcasts block column k; every rank inverts the 64x64 diagonal
//           blocks it needs, solves the row panel U_kj = U_kk^-H A_kj on ITS block columns j > k, the panel pieces are
//           packed, all-gathered (one NCCL group) and unpacked, then every rank updates ITS block columns
//           A(k+1:j, j) -= U_k,(k+1:j)^H U_kj  (one rank-MB update per owned block column, upper part only).
// U ends up complete on every rank (the later stages need all of it).  The pivot status is max-reduced.
template <typename T>
int potrf_upper_mg(cudaStream_t s, int n, T* B, int64_t ldb) {
  MgConfig& M = mg();
  if (M.comm == nullptr || M.P <= 1) return potrf_upper<T>(s, n, B, ldb, nullptr, false);
  if (n <= 0) return 0;
  if (enable_all_smem<T>() != 0) return -1;
  Context& c = ctx();
  const int P = M.P, rank = M.rank;
  const int MB = 512;
  const int nblk64 = cdiv(n, NB), nbk = cdiv(n, MB);
  const size_t dinv_elems = (size_t)nblk64 * NB * NB;
  void* scr = ctx_scratch((dinv_elems + (size_t)MB * n) * sizeof(T) + 1024);
  if (!scr) return -1;
  Arena ar(scr, c.scratch_bytes);
  T* Dinv = ar.take<T>(dinv_elems);
  T* Pk = ar.take<T>((size_t)MB * n);                 // packed row panel U_k* (kb x r, ld = MB)
  if (!Pk) { set_last_error("potrf (multi-GPU): scratch arena too small"); return -1; }
  int* dinfo = c.d_info + ST_POTRF;
  EIGB_CUDA_CHECK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  TriInv<T> ti; ti.d64 = Dinv;
  const int es = (int)sizeof(T);
  for (int k = 0; k < nbk; ++k) {
    const int k0 = k * MB, kb = n - k0 < MB ? n - k0 : MB, own = k % P;
    if (rank == own) { if (potrf_rec<T>(s, k0, k0 + kb, B, ldb, Dinv, dinfo) != 0) return -1; }
    if (mg_bcast_columns(s, B, ldb, k0, kb, own, es) != 0) return -1;
    const int r0 = k0 + kb, r = n - r0;
    if (r <= 0) break;
    if (rank != own) {      // inverted 64x64 diagonal blocks of U_kk (the owner has them from its factorization)
      trtri_blocks_kernel<T><<<dim3(cdiv(kb, NB), 8), 256, tri_smem<T>(), s>>>(B, ldb, k0 + kb, Dinv, k0 / NB);
      EIGB_LAUNCH_CHECK();
    }
    // row panel on this rank's block columns, packed into Pk
    for (int j = k + 1; j < nbk; ++j) {
      if (j % P != rank) continue;
      const int j0 = j * MB, wj = n - j0 < MB ? n - j0 : MB;
      if (trsm_rec<T>(s, 'L', 'C', k0, k0 + kb, wj, B, ldb, B + (int64_t)j0 * ldb, ldb, ti) != 0) return -1;
      rowblock_copy_kernel<T><<<dim3(cdiv(kb, 64), cdiv(wj, 4)), dim3(64, 4), 0, s>>>(B, ldb, k0, kb, j0, wj, Pk + (int64_t)(j0 - r0) * MB, MB, 0);
      EIGB_LAUNCH_CHECK();
    }
    if (mg_group(true) != 0) return -1;
    for (int j = k + 1; j < nbk; ++j) {
      const int j0 = j * MB, wj = n - j0 < MB ? n - j0 : MB;
      if (mg_bcast(s, Pk + (int64_t)(j0 - r0) * MB, (size_t)wj * MB * es, j % P) != 0) { mg_group(false); return -1; }
    }
    if (mg_group(false) != 0) return -1;
    // the complete row panel goes back into B (U must be complete everywhere); pieces this rank solved are already there
    rowblock_copy_kernel<T><<<dim3(cdiv(kb, 64), cdiv(r, 4)), dim3(64, 4), 0, s>>>(B, ldb, k0, kb, r0, r, Pk, MB, 1);
    EIGB_LAUNCH_CHECK();
    // trailing update of this rank's block columns: rows r0 .. j0+wj (upper part of the diagonal block)
    for (int j = k + 1; j < nbk; ++j) {
      if (j % P != rank) continue;
      const int j0 = j * MB, wj = n - j0 < MB ? n - j0 : MB;
      GemmParams<T> p{};
      p.M = j0 + wj - r0; p.N = wj; p.nseg = 1;
      p.A[0] = Pk; p.lda[0] = MB; p.B[0] = Pk + (int64_t)(j0 - r0) * MB; p.ldb[0] = MB; p.K[0] = kb;
      p.A[1] = p.A[0]; p.lda[1] = MB; p.B[1] = p.B[0]; p.ldb[1] = MB; p.K[1] = 0;
      p.sa[0] = p.sa[1] = -1.0; p.sb[0] = p.sb[1] = 1.0;
      p.C = B + r0 + (int64_t)j0 * ldb; p.ldc = ldb; p.alpha = -1.0; p.beta = 1.0;
      p.mode = 1; p.real_diag = 1; p.colmap = nullptr; p.diag_off = j0 - r0;
      if (gemm_launch<T>(s, true, true, p) != 0) return -1;
    }
  }
  return mg_allreduce_max_int(s, dinfo);
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
  const int other = (side == 'L') ? n : m;
  TriInv<T> ti;
  if (build_triinv<T>(s, nu, other, U, ldu, ti) != 0) return -1;
  return trsm_rec<T>(s, side, trans, 0, nu, other, U, ldu, B, ldb, ti);
}

// Reduction to standard form A <- U^-H A U^-1 (zhegst_gpu.F90:31-109 / dsygst_gpu.F90:31-98).
// A's upper triangle is read; on exit the upper triangle holds the result (diagonal blocks are full Hermitian,
// as in the reference).  The caller's strict lower triangle of A is saved into `save` first when save != nullptr
// (the reference keeps it in Z, zhegvdx_gpu.F90:145-152).
//
// Same blocked algorithm as the reference / LAPACK ?hegst(1,'U') -- k N^3 flops, half of the "two full TRSM"
// formulation -- with block size 1024-2048 and every solve done by the recursive TRSM above:
//   A_kk <- U_kk^-H herm(A_kk) U_kk^-1 ;  A_k* <- U_kk^-H A_k* ;  A_k* -= 1/2 A_kk U_k* ;
//   A_** -= A_k*^H U_k* + U_k*^H A_k* ;   A_k* -= 1/2 A_kk U_k* ;  A_k* <- A_k* U_**^-1
template <typename T>
int hegst_upper(cudaStream_t s, int n, T* A, int64_t lda, const T* U, int64_t ldu, T* save, int64_t lds) {
  if (n <= 0) return 0;
  if (enable_all_smem<T>() != 0) return -1;
  if (save) {
    // save tril(A) (the diagonal blocks get overwritten below); symmetrize_kernel with a save area also mirrors,
    // which is harmless: only the upper triangle is read afterwards
    if (symmetrize_from_upper<T>(s, n, A, lda, save, lds) != 0) return -1;
  }
  TriInv<T> Dinv;
  if (build_triinv<T>(s, n, n, U, ldu, Dinv) != 0) return -1;
  const int HB = opts().hegst_hb > 0 ? ((opts().hegst_hb + LB - 1) / LB) * LB : ((n >= 4096) ? 2048 : 1024);
  for (int k = 0; k < n; k += HB) {
    const int kb = n - k < HB ? n - k : HB;
    const int r = n - k - kb;
    T* Akk = A + k + (int64_t)k * lda;
    // diagonal block: complete it, then two-sided solve restricted to [k, k+kb)
    symmetrize_kernel<T><<<dim3(cdiv(kb, 256), kb), 256, 0, s>>>(Akk, lda, kb, (T*)nullptr, 0);
    EIGB_LAUNCH_CHECK();
    if (trsm_rec<T>(s, 'L', 'C', k, k + kb, kb, U, ldu, A + (int64_t)k * lda, lda, Dinv) != 0) return -1;
    if (trsm_rec<T>(s, 'R', 'N', k, k + kb, kb, U, ldu, A + k, lda, Dinv) != 0) return -1;
    if (r > 0) {
      T* Akr = A + k + (int64_t)(k + kb) * lda;            // A_k*  (kb x r)
      const T* Ukr = U + k + (int64_t)(k + kb) * ldu;      // U_k*
      T* Arr = A + (k + kb) + (int64_t)(k + kb) * lda;     // A_**
      if (trsm_rec<T>(s, 'L', 'C', k, k + kb, r, U, ldu, A + (int64_t)(k + kb) * lda, lda, Dinv) != 0) return -1;
      if (gemm<T>(s, 'N', 'N', kb, r, kb, -0.5, Akk, lda, Ukr, ldu, 1.0, Akr, lda) != 0) return -1;
      if (her2k_upper<T>(s, 'C', r, kb, -1.0, Akr, lda, Ukr, ldu, 1.0, Arr, lda) != 0) return -1;
      if (gemm<T>(s, 'N', 'N', kb, r, kb, -0.5, Akk, lda, Ukr, ldu, 1.0, Akr, lda) != 0) return -1;
      if (trsm_rec<T>(s, 'R', 'N', k + kb, n, kb, U, ldu, A + k, lda, Dinv) != 0) return -1;
    }
  }
  return 0;
}

#define EIGB_INST(T)                                                                                     \
  template int symmetrize_from_upper<T>(cudaStream_t, int, T*, int64_t, T*, int64_t);                    \
  template int restore_lower<T>(cudaStream_t, int, T*, int64_t, const T*, int64_t);                      \
  template int potrf_upper<T>(cudaStream_t, int, T*, int64_t, int*, bool);                                  \
  template int potrf_upper_mg<T>(cudaStream_t, int, T*, int64_t);                                  \
  template int trsm_upper<T>(cudaStream_t, char, char, int, int, const T*, int64_t, T*, int64_t);        \
  template int hegst_upper<T>(cudaStream_t, int, T*, int64_t, const T*, int64_t, T*, int64_t);
EIGB_INST(double)
EIGB_INST(double2)
#undef EIGB_INST

}  // namespace eigb200
