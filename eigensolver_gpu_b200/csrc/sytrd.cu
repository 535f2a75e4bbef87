// eigb200 -- blocked Householder tridiagonalization (UPLO='U') and the symv/hemv tile engine, sm_100a.
//
// Replaces zhetrd_gpu/zlatrd_gpu (zhetrd_gpu.F90:30-165), dsytrd_gpu/dlatrd_gpu (dsytrd_gpu.F90:30-164), the
// per-column kernels K10-K16 (zhetrd_gpu.F90:211-879), zhemv_gpu/dsymv_gpu (zhemv_gpu.F90:33-193,
// dsymv_gpu.F90:33-150) and the final-block kernel zhetd2_gpu/dsytd2_gpu.
//
// Design (B200-first, not the reference's 4 launches per column with FP64 atomics):
//  * one persistent cooperative kernel per panel, 2 grid barriers per column:
//      phase A (row-parallel "vector" phase): finish W(:,c+1) from the previous column's partial sums,
//              bring column c up to date with the panel's reflectors, partial norms;
//      phase B (tile phase): every CTA derives the Householder scalars (larfg) redundantly, then streams its
//              share of the 64x64 tiles of the upper triangle ONCE, using each tile for A_IJ x_J and
//              A_IJ^H x_I (warp-shuffle reductions), plus the V^H v / W^H v partial dots and v^H A v.
//  * all cross-CTA reductions go through per-tile partial buffers summed in a fixed order: deterministic,
//    no FP64 atomics (the reference's results depend on atomicAdd ordering).
//  * w^H v is obtained algebraically (v^H A v - 2 Re(z1^H z2)), which removes a third barrier per column.
//  * the same panel code runs down to column 1, so no separate unblocked 32x32 kernel is needed.
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"
#include <cooperative_groups.h>

namespace eigb200 {

namespace {

constexpr int TB = 64;        // symv/hemv tile edge
constexpr int NT = 256;       // threads per CTA
constexpr int NW = NT / 32;   // warps
constexpr int CPW = TB / NW;  // tile columns per warp (8)
constexpr int NBMAX = 128;    // max panel width
constexpr int MAXZU = 64;     // max number of row chunks for the V^H v / W^H v partial dots

template <typename T>
struct TrdP {
  T* A; int64_t lda;
  int i0, nbp;                 // panel = columns [i0, i0+nbp)
  T* W; int64_t ldw;           // W(:, c) <-> global column i0 + c
  double* d; double* e; T* tau;
  T* xbuf;                     // unscaled updated column
  T* Pd; T* Pt; int64_t ldp;   // hemv partials: Pd[J*ldp + r] (direct), Pt[I*ldp + r] (transposed)
  T* zpart;                    // [MAXZU][2][NBMAX]
  double* npart;               // [G] partial sums of squares
  double* vavpart;             // [G] partial v^H A v
  T* alpha_slot;               // a(j-1, j) before scaling
  unsigned* barrier;
  int* status;
  int vec_ok;                  // 16-byte loads allowed (real: A 16B aligned and lda even)
};

template <typename T>
struct PanelSmem {
  T z1[NBMAX], z2[NBMAX], rowV[NBMAX], rowW[NBMAX];
  T red[NW * TB];
  T yt[TB];
  T xI[TB], xJ[TB];
  T scal[4];        // generic scalar broadcast slots
  double dscal[8];
};

__device__ __forceinline__ double ldcg_(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg_(const double2* p) { return __ldcg(p); }

// ---- grid barrier (monotonic counter, watchdog-protected) --------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target, int* status) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned long long spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (++spins > (1ull << 26)) { atomicExch(status, 77); break; }   // watchdog: never hang the GPU
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red /* >= NW entries */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = zero_<T>();
#pragma unroll
  for (int w = 0; w < NW; ++w) s = add_(s, red[w]);
  return s;
}

__host__ __device__ __forceinline__ int zchunk_rows(int n) {
  int ch = 128;
  while ((n + ch - 1) / ch > MAXZU) ch *= 2;
  return ch;
}

// ---- Householder scalars: LAPACK ?larfg without the safmin loop (zhetrd_gpu.F90:277-331, dsytrd_gpu.F90:408-443)
__device__ __forceinline__ void larfg_scalars(double alpha, double xnorm2, double& beta, double& tau, double& scale) {
  if (xnorm2 == 0.0) { beta = alpha; tau = 0.0; scale = 1.0; return; }
  double xnorm = sqrt(xnorm2);
  beta = -copysign(hypot(alpha, xnorm), alpha);
  tau = (beta - alpha) / beta;
  scale = 1.0 / (alpha - beta);
}
__device__ __forceinline__ void larfg_scalars(double2 alpha, double xnorm2, double& beta, double2& tau, double2& scale) {
  if (xnorm2 == 0.0 && alpha.y == 0.0) { beta = alpha.x; tau = mkz(0, 0); scale = mkz(1, 0); return; }
  double xnorm = sqrt(xnorm2);
  double sc = fmax(fmax(fabs(alpha.x), fabs(alpha.y)), xnorm);
  double a = alpha.x / sc, b = alpha.y / sc, c = xnorm / sc;
  double nrm = sc * sqrt(a * a + b * b + c * c);
  beta = -copysign(nrm, alpha.x);
  tau = mkz((beta - alpha.x) / beta, -alpha.y / beta);
  // scale = 1 / (alpha - beta), Smith's division
  double xr = alpha.x - beta, xi = alpha.y;
  if (fabs(xi) <= fabs(xr)) {
    double r = xi / xr, den = xr + xi * r;
    scale = mkz(1.0 / den, -r / den);
  } else {
    double r = xr / xi, den = xi + xr * r;
    scale = mkz(r / den, -1.0 / den);
  }
}

// =====================================================================================================
// Tile engine: one 64x64 tile (I,J), I<=J, of the upper triangle of the leading n x n block of A.
// Outputs: direct partial  Pd[J*ldp + I*TB + r] = sum_c A(r,c) x(c)         (off-diagonal and diagonal)
//          transposed      Pt[I*ldp + J*TB + c] = sum_r conj(A(r,c)) x(r)    (off-diagonal only)
// Returns this thread's contribution to Re(x^H A x).
// Thread mapping: warp w owns tile columns [8w, 8w+8); lane l owns rows {2l,2l+1} (real, one 128-bit load)
// or {l, l+32} (complex, two 128-bit loads): every global load is a coalesced 128-bit access.
// =====================================================================================================
template <typename T>
__device__ __forceinline__ void tile_rows(int lane, int& r0, int& r1) {
  if (is_cplx<T>::value) { r0 = lane; r1 = lane + 32; } else { r0 = 2 * lane; r1 = 2 * lane + 1; }
}

template <typename T>
__device__ __forceinline__ void load_tile_regs(const T* __restrict__ A, int64_t lda, int n, int I, int J, int vec_ok,
                                               T (&a)[CPW][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int r0, r1; tile_rows<T>(lane, r0, r1);
  const int gr0 = I * TB + r0, gr1 = I * TB + r1;
#pragma unroll
  for (int q = 0; q < CPW; ++q) {
    const int gc = J * TB + warp * CPW + q;
    const T* col = A + (int64_t)gc * lda;
    if (gc < n) {
      if constexpr (!is_cplx<T>::value) {
        if (vec_ok && gr1 < n) {
          double2 v = __ldg(reinterpret_cast<const double2*>(col + gr0));
          a[q][0] = v.x; a[q][1] = v.y;
        } else {
          a[q][0] = (gr0 < n) ? __ldg(col + gr0) : 0.0;
          a[q][1] = (gr1 < n) ? __ldg(col + gr1) : 0.0;
        }
      } else {
        a[q][0] = (gr0 < n) ? __ldg(col + gr0) : zero_<T>();
        a[q][1] = (gr1 < n) ? __ldg(col + gr1) : zero_<T>();
      }
    } else {
      a[q][0] = zero_<T>(); a[q][1] = zero_<T>();
    }
  }
}

// xI/xJ (tile slices of x) must already be in shared memory.  All threads call; contains __syncthreads.
template <typename T>
__device__ __forceinline__ double tile_compute(const T (&a)[CPW][2], int n, int I, int J, const T* xI, const T* xJ,
                                               T* red, T* yt, T* Pd, T* Pt, int64_t ldp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int r0, r1; tile_rows<T>(lane, r0, r1);
  const int rr[2] = {r0, r1};
  T accd[2] = {zero_<T>(), zero_<T>()};
  T acct[CPW];
  const bool diag = (I == J);
  const T xr[2] = {xI[r0], xI[r1]};
#pragma unroll
  for (int q = 0; q < CPW; ++q) {
    acct[q] = zero_<T>();
    const int c = warp * CPW + q;
    const T xc = xJ[c];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!diag) {
        fma_(accd[h], a[q][h], xc);
        fmac_(acct[q], a[q][h], xr[h]);
      } else {
        if (rr[h] < c) {
          fma_(accd[h], a[q][h], xc);
          fmac_(acct[q], a[q][h], xr[h]);
        } else if (rr[h] == c) {
          fma_(accd[h], from_real<T>(real_(a[q][h])), xc);   // Hermitian: diagonal is real
        }
      }
    }
  }
  // transposed partials: reduce over the 32 lanes (rows)
#pragma unroll
  for (int q = 0; q < CPW; ++q) {
    T s = warp_sum(acct[q]);
    if (lane == 0) yt[warp * CPW + q] = s;
  }
  // direct partials: reduce over the 8 warps (column groups)
  red[warp * TB + r0] = accd[0];
  red[warp * TB + r1] = accd[1];
  __syncthreads();
  double vav = 0.0;
  if (tid < TB) {
    T s = zero_<T>();
#pragma unroll
    for (int w = 0; w < NW; ++w) s = add_(s, red[w * TB + tid]);
    const int gr = I * TB + tid;
    if (diag) {
      s = add_(s, yt[tid]);
      if (gr < n) {
        Pd[(int64_t)J * ldp + gr] = s;
        T t = zero_<T>(); fmac_(t, xI[tid], s);
        vav = real_(t);
      }
    } else {
      Pd[(int64_t)J * ldp + gr] = s;                 // off-diagonal tile: all 64 rows are < n
      T t = zero_<T>(); fmac_(t, xI[tid], s);
      vav = 2.0 * real_(t);
      const int gc = J * TB + tid;
      if (gc < n) Pt[(int64_t)I * ldp + gc] = yt[tid];
    }
  }
  __syncthreads();
  return vav;
}

__device__ __forceinline__ void tile_from_index(int idx, int& I, int& J) {
  // idx = J(J+1)/2 + I, 0 <= I <= J
  int j = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
  while ((j + 1) * (j + 2) / 2 <= idx) ++j;
  while (j * (j + 1) / 2 > idx) --j;
  J = j; I = idx - j * (j + 1) / 2;
}

// sum of the partials belonging to row r of an order-n product (T_n = ceil(n/TB) tiles per side)
template <typename T>
__device__ __forceinline__ T gather_partials(const T* Pd, const T* Pt, int64_t ldp, int n, int r) {
  const int Tn = (n + TB - 1) / TB, I = r / TB;
  T s0 = zero_<T>(), s1 = zero_<T>(), s2 = zero_<T>(), s3 = zero_<T>();
  int J = I;
  for (; J + 3 < Tn; J += 4) {
    T a = ldcg_(Pd + (int64_t)J * ldp + r), b = ldcg_(Pd + (int64_t)(J + 1) * ldp + r);
    T c = ldcg_(Pd + (int64_t)(J + 2) * ldp + r), d = ldcg_(Pd + (int64_t)(J + 3) * ldp + r);
    s0 = add_(s0, a); s1 = add_(s1, b); s2 = add_(s2, c); s3 = add_(s3, d);
  }
  for (; J < Tn; ++J) s0 = add_(s0, ldcg_(Pd + (int64_t)J * ldp + r));
  int K = 0;
  for (; K + 3 < I; K += 4) {
    T a = ldcg_(Pt + (int64_t)K * ldp + r), b = ldcg_(Pt + (int64_t)(K + 1) * ldp + r);
    T c = ldcg_(Pt + (int64_t)(K + 2) * ldp + r), d = ldcg_(Pt + (int64_t)(K + 3) * ldp + r);
    s0 = add_(s0, a); s1 = add_(s1, b); s2 = add_(s2, c); s3 = add_(s3, d);
  }
  for (; K < I; ++K) s1 = add_(s1, ldcg_(Pt + (int64_t)K * ldp + r));
  return add_(add_(s0, s1), add_(s2, s3));
}

// ---- stand-alone symv/hemv (eigb200_dsymv / eigb200_zhemv) ----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT, 2) hemv_tiles_kernel(const T* __restrict__ A, int64_t lda, int n,
                                                           const T* __restrict__ x, T* Pd, T* Pt, int64_t ldp,
                                                           int vec_ok) {
  __shared__ T red[NW * TB];
  __shared__ T yt[TB];
  __shared__ T xI[TB], xJ[TB];
  const int Tn = (n + TB - 1) / TB, ntile = Tn * (Tn + 1) / 2;
  for (int idx = blockIdx.x; idx < ntile; idx += gridDim.x) {
    int I, J; tile_from_index(idx, I, J);
    T a[CPW][2];
    load_tile_regs<T>(A, lda, n, I, J, vec_ok, a);
    if (threadIdx.x < TB) {
      int gi = I * TB + threadIdx.x, gj = J * TB + threadIdx.x;
      xI[threadIdx.x] = gi < n ? x[gi] : zero_<T>();
      xJ[threadIdx.x] = gj < n ? x[gj] : zero_<T>();
    }
    __syncthreads();
    tile_compute<T>(a, n, I, J, xI, xJ, red, yt, Pd, Pt, ldp);
  }
}
template <typename T>
__global__ void hemv_reduce_kernel(const T* Pd, const T* Pt, int64_t ldp, int n, T* y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) y[r] = gather_partials<T>(Pd, Pt, ldp, n, r);
}

// =====================================================================================================
// Panel phases
// =====================================================================================================
template <typename T>
__device__ void phase_a(const TrdP<T>& p, int c, PanelSmem<T>& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int nbp = p.nbp;
  const int j = p.i0 + c;                       // column brought up to date in this phase (c may be -1)
  const int jp = j + 1;                         // order of the product done for column c+1
  const bool have_prev = (c + 1 <= nbp - 1) && (jp >= 1) && !(c == -1 && p.i0 == 0);
  const int cprev = c + 1;
  T tau_p = zero_<T>();
  double alpha_p = 0.0;

  if (have_prev) {
    // -- z1 = V^H v, z2 = W^H v from the row-chunk partials; v^H A v from the per-CTA partials
    const int zch = zchunk_rows(jp);
    const int nzu = (jp + zch - 1) / zch;
    const int nf = nbp - 1 - cprev;             // finished columns cc in (cprev, nbp)
    for (int q = tid; q < 2 * nf; q += NT) {
      const int which = q / nf, cc = cprev + 1 + (q % nf);
      T s0 = zero_<T>(), s1 = zero_<T>();
      int u = 0;
      for (; u + 1 < nzu; u += 2) {
        T a = ldcg_(p.zpart + ((int64_t)u * 2 + which) * NBMAX + cc);
        T b = ldcg_(p.zpart + ((int64_t)(u + 1) * 2 + which) * NBMAX + cc);
        s0 = add_(s0, a); s1 = add_(s1, b);
      }
      if (u < nzu) s0 = add_(s0, ldcg_(p.zpart + ((int64_t)u * 2 + which) * NBMAX + cc));
      (which ? sm.z2 : sm.z1)[cc] = add_(s0, s1);
    }
    // row j of V and W (finished columns); rowW[cprev] is filled below
    for (int cc = cprev + 1 + tid; cc < nbp; cc += NT) {
      sm.rowV[cc] = ldcg_(p.A + j + (int64_t)(p.i0 + cc) * p.lda);
      sm.rowW[cc] = ldcg_(p.W + j + (int64_t)cc * p.ldw);
    }
    if (tid == 0) sm.rowV[cprev] = from_real<T>(1.0);   // unit element of v_{c+1} sits in row j
    double vv = 0.0;
    for (int g = tid; g < G; g += NT) vv += __ldcg(p.vavpart + g);
    vv = block_sum<double>(vv, sm.dscal);    // (contains __syncthreads: z1/z2/rowV/rowW visible after it)
    // rho = v^H A v - 2 Re(z1^H z2)
    double zz = 0.0;
    for (int cc = cprev + 1 + tid; cc < nbp; cc += NT) {
      T t = zero_<T>(); fmac_(t, sm.z1[cc], sm.z2[cc]);
      zz += real_(t);
    }
    __syncthreads();
    zz = block_sum<double>(zz, sm.dscal);
    tau_p = ldcg_(p.tau + j);                   // written by CTA 0 in the previous phase B (before a barrier)
    const double rho = vv - 2.0 * zz;
    alpha_p = -0.5 * abs2_(tau_p) * rho;
    // row j of the new W column: u_j = wraw_j - sum_cc (W(j,cc) z1(cc) + V(j,cc) z2(cc))
    T part = zero_<T>();
    for (int cc = cprev + 1 + tid; cc < nbp; cc += NT) {
      fma_(part, sm.rowW[cc], sm.z1[cc]);
      fma_(part, sm.rowV[cc], sm.z2[cc]);
    }
    __syncthreads();
    part = block_sum<T>(part, sm.red);
    if (warp == 0) {
      // wraw_j: lane-strided gather of the partials of row j, fixed order
      const int Tn = (jp + TB - 1) / TB, I = j / TB;
      T wr = zero_<T>();
      for (int J = I + lane; J < Tn; J += 32) wr = add_(wr, ldcg_(p.Pd + (int64_t)J * p.ldp + j));
      for (int K = lane; K < I; K += 32) wr = add_(wr, ldcg_(p.Pt + (int64_t)K * p.ldp + j));
      wr = warp_sum(wr);
      if (lane == 0) {
        T w = mul_(tau_p, sub_(wr, part));
        sm.rowW[cprev] = add_(w, from_real<T>(alpha_p));    // + alpha' * v(j), v(j) = 1
      }
    }
    __syncthreads();
  }

  // -- row-parallel part
  double nrm = 0.0;
  const int nrows = (c >= 0) ? (j + 1) : jp;     // c == -1: only finish W(:, 0), rows [0, i0)
  for (int r = cta * NT + tid; r < nrows; r += G * NT) {
    T t1 = zero_<T>(), t2 = zero_<T>();
    for (int cc = cprev + 1; cc < nbp; ++cc) {
      const T vv = ldcg_(p.A + r + (int64_t)(p.i0 + cc) * p.lda);
      const T ww = ldcg_(p.W + r + (int64_t)cc * p.ldw);
      if (have_prev) { fma_(t1, ww, sm.z1[cc]); fma_(t1, vv, sm.z2[cc]); }
      if (c >= 0) { fma_(t2, vv, conj_(sm.rowW[cc])); fma_(t2, ww, conj_(sm.rowV[cc])); }
    }
    if (have_prev) {
      const T wraw = gather_partials<T>(p.Pd, p.Pt, p.ldp, jp, r);
      const T vnew = ldcg_(p.A + r + (int64_t)(j + 1) * p.lda);
      T wnew = mul_(tau_p, sub_(wraw, t1));
      wnew = add_(wnew, scale_(vnew, alpha_p));
      p.W[r + (int64_t)cprev * p.ldw] = wnew;
      if (c >= 0) { fma_(t2, vnew, conj_(sm.rowW[cprev])); fma_(t2, wnew, conj_(sm.rowV[cprev])); }
    }
    if (c >= 0) {
      T a = sub_(ldcg_(p.A + r + (int64_t)j * p.lda), t2);
      if (r == j) {
        a = from_real<T>(real_(a));
        p.d[j] = real_(a);
        p.A[j + (int64_t)j * p.lda] = a;
      } else {
        p.xbuf[r] = a;
        if (r < j - 1) nrm += abs2_(a);
        if (r == j - 1) *p.alpha_slot = a;
      }
    }
  }
  if (c >= 0) {
    __syncthreads();
    nrm = block_sum<double>(nrm, sm.dscal);
    if (tid == 0) p.npart[cta] = nrm;
  }
}

template <typename T>
__device__ void phase_b(const TrdP<T>& p, int c, PanelSmem<T>& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int j = p.i0 + c;            // order of the product; reflector index j-1
  if (j <= 0) return;
  // -- Householder scalars (every CTA, same summation order => identical values everywhere)
  double x2 = 0.0;
  for (int g = tid; g < G; g += NT) x2 += __ldcg(p.npart + g);
  x2 = block_sum<double>(x2, sm.dscal);
  const T alpha = ldcg_(p.alpha_slot);
  double beta; T tau, scale;
  larfg_scalars(alpha, x2, beta, tau, scale);
  if (cta == 0 && tid == 0) { p.e[j - 1] = beta; p.tau[j - 1] = tau; }
  // -- store v in A(:, j) (nobody reads column j in this phase)
  for (int r = cta * NT + tid; r < j; r += G * NT) {
    T v = (r == j - 1) ? from_real<T>(1.0) : mul_(scale, ldcg_(p.xbuf + r));
    p.A[r + (int64_t)j * p.lda] = v;
  }
  auto xval = [&](int r) -> T {
    if (r >= j) return zero_<T>();
    if (r == j - 1) return from_real<T>(1.0);
    return mul_(scale, ldcg_(p.xbuf + r));
  };
  const int nf = p.nbp - 1 - c;      // finished columns cc in (c, nbp)
  const int zch = zchunk_rows(j);
  const int nzu = (nf > 0) ? (j + zch - 1) / zch : 0;
  const int Tn = (j + TB - 1) / TB, ntile = Tn * (Tn + 1) / 2;
  double vav = 0.0;
  for (int unit = cta; unit < nzu + ntile; unit += G) {
    if (unit < nzu) {
      // ---- partial dots z1 = V^H v, z2 = W^H v over rows [unit*zch, ...)
      const int rbeg = unit * zch, rend = min(j, rbeg + zch);
      for (int q = warp; q < 2 * nf; q += NW) {
        const int which = q / nf, cc = c + 1 + (q % nf);
        const T* col = which ? (p.W + (int64_t)cc * p.ldw) : (p.A + (int64_t)(p.i0 + cc) * p.lda);
        T s = zero_<T>();
        for (int r = rbeg + lane; r < rend; r += 32) fmac_(s, ldcg_(col + r), xval(r));
        s = warp_sum(s);
        if (lane == 0) p.zpart[((int64_t)unit * 2 + which) * NBMAX + cc] = s;
      }
    } else {
      int I, J; tile_from_index(unit - nzu, I, J);
      T a[CPW][2];
      load_tile_regs<T>(p.A, p.lda, j, I, J, p.vec_ok, a);
      __syncthreads();
      if (tid < TB) sm.xI[tid] = xval(I * TB + tid);
      else if (tid < 2 * TB) sm.xJ[tid - TB] = xval(J * TB + tid - TB);
      __syncthreads();
      vav += tile_compute<T>(a, j, I, J, sm.xI, sm.xJ, sm.red, sm.yt, p.Pd, p.Pt, p.ldp);
    }
  }
  __syncthreads();
  vav = block_sum<double>(vav, sm.dscal);
  if (tid == 0) p.vavpart[cta] = vav;
}

template <typename T>
__global__ void __launch_bounds__(NT, 2) panel_coop_kernel(TrdP<T> p) {
  __shared__ PanelSmem<T> sm;
  unsigned target = 0;
  for (int c = p.nbp - 1; c >= -1; --c) {
    phase_a<T>(p, c, sm);
    if (c < 0) break;
    target += gridDim.x;
    grid_barrier(p.barrier, target, p.status);
    phase_b<T>(p, c, sm);
    target += gridDim.x;
    grid_barrier(p.barrier, target, p.status);
  }
}
template <typename T>
__global__ void __launch_bounds__(NT, 2) phase_a_kernel(TrdP<T> p, int c) {
  __shared__ PanelSmem<T> sm;
  phase_a<T>(p, c, sm);
}
template <typename T>
__global__ void __launch_bounds__(NT, 2) phase_b_kernel(TrdP<T> p, int c) {
  __shared__ PanelSmem<T> sm;
  phase_b<T>(p, c, sm);
}

template <typename T>
int panel_grid(int& grid) {
  int per_sm = 0;
  EIGB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, panel_coop_kernel<T>, NT, 0));
  if (per_sm < 1) { set_last_error("panel kernel does not fit on an SM"); return -1; }
  if (per_sm > 2) per_sm = 2;
  grid = per_sm * ctx().num_sms;
  return 0;
}

}  // namespace

// scratch layout helper shared by hemv() and hetrd()
template <typename T>
static size_t partial_elems(int n, int64_t& ldp) {
  ldp = ((int64_t)n + 63) & ~int64_t(63);
  int Tn = (n + TB - 1) / TB;
  return (size_t)ldp * (size_t)(Tn > 0 ? Tn : 1);
}

template <typename T>
int hemv_upper(cudaStream_t s, int n, const T* A, int64_t lda, const T* x, T* y) {
  if (n <= 0) return 0;
  int64_t ldp;
  size_t pe = partial_elems<T>(n, ldp);
  void* scr = ctx_scratch(2 * pe * sizeof(T) + 512);
  if (!scr) return -1;
  Arena ar(scr, ctx().scratch_bytes);
  T* Pd = ar.take<T>(pe);
  T* Pt = ar.take<T>(pe);
  int vec_ok = is_cplx<T>::value ? 1 : ((((uintptr_t)A & 15) == 0 && (lda & 1) == 0) ? 1 : 0);
  int Tn = (n + TB - 1) / TB, ntile = Tn * (Tn + 1) / 2;
  int grid = 2 * ctx().num_sms;
  if (grid > ntile) grid = ntile;
  hemv_tiles_kernel<T><<<grid, NT, 0, s>>>(A, lda, n, x, Pd, Pt, ldp, vec_ok);
  EIGB_LAUNCH_CHECK();
  hemv_reduce_kernel<T><<<cdiv(n, 256), 256, 0, s>>>(Pd, Pt, ldp, n, y);
  EIGB_LAUNCH_CHECK();
  return 0;
}

// Tridiagonalization driver.  A: n x n (upper read/written), outputs d(n), e(n-1), tau(n-1) on device.
// On exit the reflectors v_j (j = 1..n-1, 1-based) are in A(1:j-1, j+1) with the unit element stored
// explicitly, exactly as the reference leaves them (zhetrd_gpu.F90:92).
template <typename T>
int hetrd_upper(cudaStream_t s, int n, T* A, int64_t lda, double* d, double* e, T* tau) {
  if (n <= 0) return 0;
  Context& c = ctx();
  const int nb = opts().trd_nb < NBMAX ? opts().trd_nb : NBMAX;
  int grid = 0;
  if (panel_grid<T>(grid) != 0) return -1;
  int64_t ldp;
  size_t pe = partial_elems<T>(n, ldp);
  size_t bytes = (2 * pe + (size_t)n * nb + (size_t)n + (size_t)MAXZU * 2 * NBMAX + 64) * sizeof(T) +
                 (size_t)(2 * grid + 64) * sizeof(double) + 4096 + 16 * 256;
  void* scr = ctx_scratch(bytes);
  if (!scr) return -1;
  Arena ar(scr, c.scratch_bytes);
  TrdP<T> p{};
  p.A = A; p.lda = lda; p.d = d; p.e = e; p.tau = tau;
  p.Pd = ar.take<T>(pe); p.Pt = ar.take<T>(pe); p.ldp = ldp;
  p.W = ar.take<T>((size_t)n * nb); p.ldw = n;
  p.xbuf = ar.take<T>(n);
  p.zpart = ar.take<T>((size_t)MAXZU * 2 * NBMAX);
  p.alpha_slot = ar.take<T>(16);
  p.npart = ar.take<double>(grid);
  p.vavpart = ar.take<double>(grid);
  p.barrier = ar.take<unsigned>(64);
  p.status = c.d_info;
  if (!p.barrier) { set_last_error("hetrd: scratch arena too small"); return -1; }
  p.vec_ok = is_cplx<T>::value ? 1 : ((((uintptr_t)A & 15) == 0 && (lda & 1) == 0) ? 1 : 0);
  EIGB_CUDA_CHECK(cudaMemsetAsync(p.status, 0, sizeof(int), s));
  const bool coop = opts().trd_coop != 0;

  int hi = n;                         // columns [0, hi) still to reduce
  while (hi > 0) {
    int nbp = hi < nb ? hi : nb;
    // keep panels aligned so that the last (leftmost) panel absorbs the remainder
    if (hi > nb && (hi % nb) != 0) nbp = hi % nb;
    p.i0 = hi - nbp; p.nbp = nbp;
    prof_begin(PROF_PANEL, s);
    if (coop) {
      EIGB_CUDA_CHECK(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), s));
      void* args[] = {&p};
      EIGB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)panel_coop_kernel<T>, dim3(grid), dim3(NT), args, 0, s));
      count_launch(1);
    } else {
      for (int cc = nbp - 1; cc >= -1; --cc) {
        phase_a_kernel<T><<<grid, NT, 0, s>>>(p, cc);
        if (cc < 0) break;
        phase_b_kernel<T><<<grid, NT, 0, s>>>(p, cc);
      }
      count_launch(2 * nbp + 1);
      EIGB_CUDA_CHECK(cudaGetLastError());
    }
    prof_end(PROF_PANEL, s);
    // trailing update A(0:i0, 0:i0) -= V W^H + W V^H   (zhetrd_gpu.F90:67 / dsytrd_gpu.F90:66)
    if (p.i0 > 0) {
      ProfScope ps(PROF_HER2K, s);
      if (her2k_upper<T>(s, 'N', p.i0, nbp, -1.0, A + (int64_t)p.i0 * lda, lda, p.W, p.ldw, 1.0, A, lda) != 0)
        return -1;
    }
    hi = p.i0;
  }
  int st = 0;
  EIGB_CUDA_CHECK(cudaMemcpyAsync(&st, p.status, sizeof(int), cudaMemcpyDeviceToHost, s));
  EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  if (st != 0) { set_last_error("hetrd: device status %d (grid barrier watchdog)", st); return -1; }
  return 0;
}

template int hemv_upper<double>(cudaStream_t, int, const double*, int64_t, const double*, double*);
template int hemv_upper<double2>(cudaStream_t, int, const double2*, int64_t, const double2*, double2*);
template int hetrd_upper<double>(cudaStream_t, int, double*, int64_t, double*, double*, double*);
template int hetrd_upper<double2>(cudaStream_t, int, double2*, int64_t, double*, double*, double2*);

}  // namespace eigb200
