// eigb200 -- on-device symmetric tridiagonal eigensolver (divide and conquer), sm_100a.
//
// Replaces the reference's HOST call  zstedc('I') / dstedc('I')  (zheevd_gpu.F90:99-107, dsyevd_gpu.F90:97-105)
// and the D2H/H2D traffic around it (zheevd_gpu.F90:85-86,110-111).  The mathematics is the published
// Cuppen / Gu-Eisenstat divide and conquer that LAPACK ?stedc implements (rank-one tearing, deflation,
// secular equation, Loewner re-computation of z, eigenvector update by GEMM); see SURVEY.md Appendix B.
// Nothing runs on the host: the tree shape depends only on n, every data-dependent size (number of
// non-deflated roots k, column-type counts) stays in device memory and the batched DMMA GEMMs read their
// parameter blocks from there.
//
// Per level, all merges are processed by one launch of each kernel (blockIdx.y = merge):
//   sort -> deflate (1 CTA per merge, sequential scan in shared memory) -> apply Givens rotations ->
//   gather columns by type -> secular roots (one warp per root) -> final ranks -> z-hat (one warp per row)
//   -> eigenvectors of the rank-one system (one warp per column) -> batched GEMM (top / bottom halves,
//   type-structured) writing straight into the sorted column positions -> scatter deflated columns.
#include "common.cuh"
#include "gemm.cuh"
#include "secular.cuh"
#include "stages.cuh"

namespace eigb200 {

namespace {

constexpr int LEAF = 32;
constexpr double QL_EPS = 2.220446049250313e-16;    // leaf QL iteration: EISPACK tql2's machine epsilon
constexpr double DC_EPS = 1.1102230246251565e-16;   // deflation: LAPACK dlamch('Epsilon'), as dlaed2

__host__ __device__ __forceinline__ int leaf_bound(int i, int n, int L) { return (int)(((long long)i * n) / L); }

struct DcWork {
  int n, L, levels;
  double* D;            // eigenvalues (in/out), length n
  double* E;            // off-diagonal, length n
  double* Q; int64_t ldq;
  double* WS; int64_t ldw;   // gathered columns
  double* X; int64_t ldx;    // delta / rank-one eigenvectors
  double *Dsort, *zsort, *dlam, *wz, *zhat, *Dnew, *Ddef, *rotc, *rots, *scale;
  int *idx, *typ, *srccol, *posg, *cmap, *dstcol, *rotp, *rotn, *permnd, *permdf;
  int *K;               // per merge: k, k1, k2, k3, nrot  (5 ints per merge)
  GemmParams<double>* gp;
  int* sel;             // root merge restricted to the wanted columns: [j_lo, j_hi) of the non-deflated roots
  int* status;          // device status word: 1 leaf QL did not converge, 2 secular equation, 3 non-finite eigenvalue
};

struct MergeGeom { int lo, n1, n2, n; };
__device__ __forceinline__ MergeGeom merge_geom(const DcWork& w, int level, int m) {
  const int span = 1 << level;
  MergeGeom g;
  g.lo = leaf_bound(m * span, w.n, w.L);
  const int mid = leaf_bound(m * span + span / 2, w.n, w.L);
  const int hi = leaf_bound((m + 1) * span, w.n, w.L);
  g.n1 = mid - g.lo; g.n2 = hi - mid; g.n = hi - g.lo;
  return g;
}

// ---- scaling ---------------------------------------------------------------------------------------
__global__ void dc_scale_kernel(DcWork w) {
  __shared__ double red[32];
  double mx = 0.0;
  bool bad = false;
  for (int i = threadIdx.x; i < w.n; i += blockDim.x) {
    // NaN/Inf in T: every comparison-based step below (ranks, deflation, column maps) would go wrong in silence or index
    // out of bounds.  Report through the status word and continue on a sanitised copy so that nothing can fault.
    if (!isfinite(w.D[i])) { bad = true; w.D[i] = 0.0; }
    mx = fmax(mx, fabs(w.D[i]));
    if (i < w.n - 1) {
      if (!isfinite(w.E[i])) { bad = true; w.E[i] = 0.0; }
      mx = fmax(mx, fabs(w.E[i]));
    }
  }
  if (bad) atomicMax(w.status, 3);
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, red[i]);
  const double sc = (mx > 0.0 && isfinite(mx)) ? mx : 1.0;
  if (threadIdx.x == 0) w.scale[0] = sc;
  const double inv = 1.0 / sc;
  for (int i = threadIdx.x; i < w.n; i += blockDim.x) {
    w.D[i] *= inv;
    if (i < w.n - 1) w.E[i] *= inv;
  }
}
__global__ void dc_unscale_kernel(DcWork w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.n) {
    const double v = w.D[i] * w.scale[0];
    w.D[i] = v;
    if (!isfinite(v)) atomicMax(w.status, 3);
  }
}

// ---- leaves: implicit QL with eigenvectors (EISPACK tql2 recurrence), one warp per leaf -----------------
__global__ void __launch_bounds__(128) dc_leaf_kernel(DcWork w) {
  __shared__ double V[4][LEAF][LEAF + 1];
  __shared__ double sd[4][LEAF], se[4][LEAF];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = blockIdx.x * 4 + wid;
  if (leaf >= w.L) return;
  const int lo = leaf_bound(leaf, w.n, w.L), hi = leaf_bound(leaf + 1, w.n, w.L);
  const int n = hi - lo;
  double* d = sd[wid];
  double* e = se[wid];
  if (lane < n) {
    double dv = w.D[lo + lane];
    // rank-one tearing at both leaf boundaries: d -= |e| (every leaf boundary is a tear of some merge)
    if (lane == 0 && lo > 0) dv -= fabs(w.E[lo - 1]);
    if (lane == n - 1 && hi < w.n) dv -= fabs(w.E[hi - 1]);
    d[lane] = dv;
    e[lane] = (lane < n - 1) ? w.E[lo + lane] : 0.0;
  }
  for (int c = 0; c < n; ++c) V[wid][lane][c] = (lane == c) ? 1.0 : 0.0;
  __syncwarp();
  double f = 0.0, tst1 = 0.0;
  for (int l = 0; l < n; ++l) {
    tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
    int m = l;
    while (m < n - 1) { if (fabs(e[m]) <= QL_EPS * tst1) break; ++m; }
    if (m > l) {
      int iter = 0;
      bool again;
      do {
        ++iter;
        // every lane runs the same scalar recurrence; shared d/e are read, then (after a warp barrier)
        // overwritten by all lanes with identical values
        double g = d[l];
        const double el = e[l];
        double p = (d[l + 1] - g) / (2.0 * el);
        double r = hypot(p, 1.0);
        if (p < 0) r = -r;
        const double dl = el / (p + r), dl1 = el * (p + r);
        const double h = g - dl;
        const double el1 = e[l + 1];
        __syncwarp();
        d[l] = dl; d[l + 1] = dl1;
        if (lane >= l + 2 && lane < n) d[lane] -= h;
        __syncwarp();
        f += h;
        p = d[m];
        double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          const double ei = e[i], di = d[i];
          __syncwarp();
          c3 = c2; c2 = c; s2 = s;
          g = c * ei;
          const double hh = c * p;
          r = hypot(p, ei);
          s = ei / r; c = p / r;
          p = c * di - s * g;
          e[i + 1] = s2 * r;               // uses the previous rotation's sine
          d[i + 1] = hh + s * (c * g + s * di);
          const double vh = V[wid][lane][i + 1], vi = V[wid][lane][i];
          V[wid][lane][i + 1] = s * vi + c * vh;
          V[wid][lane][i] = c * vi - s * vh;
        }
        __syncwarp();
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        __syncwarp();
        e[l] = s * p; d[l] = c * p;
        __syncwarp();
        again = fabs(e[l]) > QL_EPS * tst1;
        if (again && iter >= 60) { again = false; if (lane == 0) atomicMax(w.status, 1); }   // ?steql's 30*n limit
        __syncwarp();
      } while (again);
    }
    const double dfin = d[l] + f;
    __syncwarp();
    d[l] = dfin; e[l] = 0.0;
    __syncwarp();
  }
  // ascending order: rank of each eigenvalue (ties by index), then scatter columns
  int rank = 0;
  double mine = (lane < n) ? d[lane] : 0.0;
  for (int y = 0; y < n; ++y) {
    const double v = d[y];
    rank += (v < mine || (v == mine && y < lane)) ? 1 : 0;
  }
  if (lane < n) {
    w.D[lo + rank] = mine;
    for (int r = 0; r < n; ++r) w.Q[(lo + r) + (int64_t)(lo + rank) * w.ldq] = V[wid][r][lane];
  }
}

// ---- per-merge kernels -------------------------------------------------------------------------------
// sort the union of the two children's (sorted) eigenvalues; form z
__global__ void dc_sort_kernel(DcWork w, int level) {
  const MergeGeom g = merge_geom(w, level, blockIdx.y);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const double* D = w.D + g.lo;
  const double di = D[i];
  const double rho_e = w.E[g.lo + g.n1 - 1];
  int rank;
  double z;
  const double rs2 = 0.70710678118654752440;
  if (i < g.n1) {
    // number of child-2 entries strictly below di
    int a = 0, b = g.n2;
    while (a < b) { int mid = (a + b) >> 1; if (D[g.n1 + mid] < di) a = mid + 1; else b = mid; }
    rank = i + a;
    z = w.Q[(g.lo + g.n1 - 1) + (int64_t)(g.lo + i) * w.ldq] * rs2;
  } else {
    // number of child-1 entries <= di
    int a = 0, b = g.n1;
    while (a < b) { int mid = (a + b) >> 1; if (D[mid] <= di) a = mid + 1; else b = mid; }
    rank = (i - g.n1) + a;
    z = w.Q[(g.lo + g.n1) + (int64_t)(g.lo + i) * w.ldq] * rs2;
    if (rho_e < 0.0) z = -z;
  }
  w.Dsort[g.lo + rank] = di;
  w.zsort[g.lo + rank] = z;
  w.idx[g.lo + rank] = i;
  w.typ[g.lo + rank] = (i < g.n1) ? 1 : 3;
}

// deflation (the dlaed2 scan) -- one CTA per merge, thread 0 scans, data staged in shared memory when it fits
constexpr int DEFL_SMEM_ELEMS = 12288;   // 12288 * 16 B = 192 KB
__global__ void __launch_bounds__(256) dc_deflate_kernel(DcWork w, int level, int use_smem) {
  extern __shared__ double sh[];
  __shared__ double red[16];
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int n = g.n, lo = g.lo;
  double* Ds = w.Dsort + lo;
  double* zs = w.zsort + lo;
  if (use_smem) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sh[i] = Ds[i]; sh[n + i] = zs[i]; }
    Ds = sh; zs = sh + n;
  }
  double dmax = 0.0, zmax = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { dmax = fmax(dmax, fabs(Ds[i])); zmax = fmax(zmax, fabs(zs[i])); }
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = dmax; red[8 + (threadIdx.x >> 5)] = zmax; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  dmax = 0.0; zmax = 0.0;
  for (int i = 0; i < 8; ++i) { dmax = fmax(dmax, red[i]); zmax = fmax(zmax, red[8 + i]); }

  const double rho = 2.0 * fabs(w.E[lo + g.n1 - 1]);
  const double tol = 8.0 * DC_EPS * fmax(dmax, zmax);
  int* typ = w.typ + lo;
  const int* idx = w.idx + lo;
  double* dlam = w.dlam + lo; double* wz = w.wz + lo; double* Ddef = w.Ddef + lo;
  int* permnd = w.permnd + lo; int* permdf = w.permdf + lo;
  int* rotp = w.rotp + lo; int* rotn = w.rotn + lo; double* rotc = w.rotc + lo; double* rots = w.rots + lo;
  int k = 0, nd = 0, nrot = 0;
  if (rho * zmax <= tol) {
    for (int p = 0; p < n; ++p) { Ddef[nd] = Ds[p]; permdf[nd++] = p; }
  } else {
    int pj = -1;
    for (int p = 0; p < n; ++p) {
      const double zp = zs[p];
      if (rho * fabs(zp) <= tol) { Ddef[nd] = Ds[p]; permdf[nd++] = p; continue; }
      if (pj < 0) { pj = p; continue; }
      double s = zs[pj], c = zp;
      const double tau = hypot(c, s);
      const double t = Ds[p] - Ds[pj];
      c /= tau; s = -s / tau;
      if (fabs(t * c * s) <= tol) {
        zs[p] = tau; zs[pj] = 0.0;
        rotp[nrot] = idx[pj]; rotn[nrot] = idx[p]; rotc[nrot] = c; rots[nrot] = s; ++nrot;
        const double dp = Ds[pj], dn = Ds[p];
        Ds[pj] = dp * c * c + dn * s * s;
        Ds[p] = dp * s * s + dn * c * c;
        if (typ[pj] != typ[p]) typ[p] = 2;
        Ddef[nd] = Ds[pj]; permdf[nd++] = pj;
        pj = p;
      } else {
        dlam[k] = Ds[pj]; wz[k] = zs[pj]; permnd[k] = pj; ++k;
        pj = p;
      }
    }
    if (pj >= 0) { dlam[k] = Ds[pj]; wz[k] = zs[pj]; permnd[k] = pj; ++k; }
  }
  // group the non-deflated columns by type: [type 1 | type 2 | type 3]
  int cnt[4] = {0, 0, 0, 0};
  for (int i = 0; i < k; ++i) cnt[typ[permnd[i]]]++;
  int off[4]; off[1] = 0; off[2] = cnt[1]; off[3] = cnt[1] + cnt[2];
  int* posg = w.posg + lo; int* srccol = w.srccol + lo;
  for (int i = 0; i < k; ++i) {
    const int t = typ[permnd[i]];
    const int gpos = off[t]++;
    posg[i] = gpos;
    srccol[gpos] = idx[permnd[i]];
  }
  for (int e = 0; e < nd; ++e) srccol[k + e] = idx[permdf[e]];
  int* K = w.K + 5 * (int64_t)m;
  K[0] = k; K[1] = cnt[1]; K[2] = cnt[2]; K[3] = cnt[3]; K[4] = nrot;
  // GEMM parameter blocks: top rows use types 1+2, bottom rows types 2+3
  const int k1 = cnt[1], k2 = cnt[2], k3 = cnt[3];
  for (int half = 0; half < 2; ++half) {
    GemmParams<double> p;
    p.M = half ? g.n2 : g.n1; p.N = k; p.nseg = 1;
    const int kk = half ? (k2 + k3) : (k1 + k2);
    const int rowoff = half ? g.n1 : 0, koff = half ? k1 : 0;
    p.A[0] = w.WS + (lo + rowoff) + (int64_t)(lo + koff) * w.ldw; p.lda[0] = w.ldw;
    p.B[0] = w.X + (lo + koff) + (int64_t)lo * w.ldx; p.ldb[0] = w.ldx;
    p.K[0] = kk; p.A[1] = p.A[0]; p.B[1] = p.B[0]; p.lda[1] = p.lda[0]; p.ldb[1] = p.ldb[0]; p.K[1] = 0;
    p.sa[0] = p.sa[1] = p.sb[0] = p.sb[1] = 1.0;
    p.C = w.Q + (lo + rowoff) + (int64_t)lo * w.ldq; p.ldc = w.ldq;
    p.alpha = 1.0; p.beta = 0.0; p.mode = 0; p.real_diag = 0;
    p.colmap = w.cmap + lo;
    w.gp[2 * m + half] = p;
  }
}

__global__ void dc_rotate_kernel(DcWork w, int level) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int nrot = w.K[5 * m + 4];
  if (nrot == 0) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g.n) return;
  double* Qb = w.Q + (g.lo + r) + (int64_t)g.lo * w.ldq;
  for (int t = 0; t < nrot; ++t) {
    const int cp = w.rotp[g.lo + t], cn = w.rotn[g.lo + t];
    const double c = w.rotc[g.lo + t], s = w.rots[g.lo + t];
    const double x = Qb[(int64_t)cp * w.ldq], y = Qb[(int64_t)cn * w.ldq];
    Qb[(int64_t)cp * w.ldq] = c * x + s * y;
    Qb[(int64_t)cn * w.ldq] = c * y - s * x;
  }
}

__global__ void dc_gather_kernel(DcWork w, int level) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int col = blockIdx.x;
  if (col >= g.n) return;
  const int src = w.srccol[g.lo + col];
  const double* s = w.Q + g.lo + (int64_t)(g.lo + src) * w.ldq;
  double* d = w.WS + g.lo + (int64_t)(g.lo + col) * w.ldw;
  for (int r = threadIdx.x; r < g.n; r += blockDim.x) d[r] = s[r];
}

struct WarpSecularEval {
  int k, j, lane; const double* d; const double* z;
  __device__ __forceinline__ SecularSums operator()(int K, double tau) const {
    double psi = 0, phi = 0, dpsi = 0, dphi = 0, tj = 0, tj1 = 0;
    const double dK = d[K];
    for (int i = lane; i < k; i += 32) {
      const double zi = z[i];
      const double delta = (d[i] - dK) - tau;
      const double t = zi / delta;
      const double term = zi * t;
      if (i <= j) { psi += term; dpsi += t * t; } else { phi += term; dphi += t * t; }
      if (i == j) tj = term;
      if (i == j + 1) tj1 = term;
    }
    SecularSums s;
    s.psi = warp_sum(psi); s.phi = warp_sum(phi); s.dpsi = warp_sum(dpsi); s.dphi = warp_sum(dphi);
    s.tj = warp_sum(tj); s.tj1 = warp_sum(tj1);
    return s;
  }
};

__global__ void __launch_bounds__(256) dc_secular_kernel(DcWork w, int level) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int k = w.K[5 * m];
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= k) return;
  const double* d = w.dlam + g.lo;
  const double* z = w.wz + g.lo;
  const double rho = 2.0 * fabs(w.E[g.lo + g.n1 - 1]);
  double zn2 = 0.0;
  for (int i = lane; i < k; i += 32) zn2 += z[i] * z[i];
  zn2 = warp_sum(zn2);
  WarpSecularEval ev{k, j, lane, d, z};
  int Ko, it; double tau;
  secular_root(k, j, d, z, rho, zn2, ev, Ko, tau, it);
  if (it > 80 && lane == 0) atomicMax(w.status, 2);
  const double dK = d[Ko];
  if (lane == 0) w.Dnew[g.lo + j] = dK + tau;
  const int* posg = w.posg + g.lo;
  double* Xc = w.X + g.lo + (int64_t)(g.lo + j) * w.ldx;
  for (int i = lane; i < k; i += 32) Xc[posg[i]] = (d[i] - dK) - tau;
}

// final ascending order of [roots (ascending) ; deflated (any order)]
__global__ void dc_rank_kernel(DcWork w, int level) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int k = w.K[5 * m];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.n) return;
  const double* R = w.Dnew + g.lo;
  const double* F = w.Ddef + g.lo;
  const double mine = (e < k) ? R[e] : F[e - k];
  int rank = 0;
  for (int y = 0; y < k; ++y) { const double v = R[y]; rank += (v < mine || (v == mine && y < e)) ? 1 : 0; }
  for (int y = k; y < g.n; ++y) { const double v = F[y - k]; rank += (v < mine || (v == mine && y < e)) ? 1 : 0; }
  w.D[g.lo + rank] = mine;
  if (e < k) w.cmap[g.lo + e] = rank; else w.dstcol[g.lo + e - k] = rank;
}

// z-hat_i = sign(z_i) sqrt( - prod_j (d_i - lambda_j) / prod_{j != i} (d_i - d_j) )   (Gu-Eisenstat)
__global__ void __launch_bounds__(256) dc_zhat_kernel(DcWork w, int level) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int k = w.K[5 * m];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= k) return;
  const double* d = w.dlam + g.lo;
  const int gi = w.posg[g.lo + i];
  const double* Xr = w.X + (g.lo + gi) + (int64_t)g.lo * w.ldx;
  const double di = d[i];
  double p = 1.0;
  for (int j = lane; j < k; j += 32) {
    const double x = Xr[(int64_t)j * w.ldx];
    p *= (j == i) ? x : x / (di - d[j]);
  }
  for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) w.zhat[g.lo + gi] = copysign(sqrt(fabs(p)), w.wz[g.lo + i]);
}

__global__ void __launch_bounds__(256) dc_formu_kernel(DcWork w, int level, int restricted) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int k = w.K[5 * m];
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= k) return;
  if (restricted && (j < w.sel[0] || j >= w.sel[1])) return;     // column not wanted (root merge only)
  double* Xc = w.X + g.lo + (int64_t)(g.lo + j) * w.ldx;
  const double* zh = w.zhat + g.lo;
  if (k == 1) { if (lane == 0) Xc[0] = 1.0; return; }
  double s = 0.0;
  for (int q = lane; q < k; q += 32) { const double v = zh[q] / Xc[q]; s += v * v; }
  s = warp_sum(s);
  const double inv = 1.0 / sqrt(s);
  for (int q = lane; q < k; q += 32) Xc[q] = (zh[q] / Xc[q]) * inv;
}

__global__ void dc_scatter_kernel(DcWork w, int level, int c_lo, int c_hi) {
  const int m = blockIdx.y;
  const MergeGeom g = merge_geom(w, level, m);
  const int k = w.K[5 * m];
  const int e = blockIdx.x;
  if (e >= g.n - k) return;
  const int dst = w.dstcol[g.lo + e];
  if (dst < c_lo || dst >= c_hi) return;                         // (root merge with a restricted column range)
  const double* s = w.WS + g.lo + (int64_t)(g.lo + k + e) * w.ldw;
  double* d = w.Q + g.lo + (int64_t)(g.lo + dst) * w.ldq;
  for (int r = threadIdx.x; r < g.n; r += blockDim.x) d[r] = s[r];
}

// Root merge, eigenvectors wanted only for the final (sorted) columns [c_lo, c_hi): the roots are ascending, so their
// final positions cmap[0..k) are increasing and the wanted ones form one contiguous range [j_lo, j_hi) of the
// non-deflated columns.  Restrict the two GEMMs of the merge (their parameter blocks live in device memory) to it.
__global__ void dc_restrict_kernel(DcWork w, int c_lo, int c_hi) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int k = w.K[0];
  const int* cmap = w.cmap;           // root merge: lo = 0
  int a = 0, b = k;
  while (a < b) { const int mid = (a + b) >> 1; if (cmap[mid] < c_lo) a = mid + 1; else b = mid; }
  const int j_lo = a;
  b = k;
  while (a < b) { const int mid = (a + b) >> 1; if (cmap[mid] < c_hi) a = mid + 1; else b = mid; }
  const int j_hi = a;
  w.sel[0] = j_lo; w.sel[1] = j_hi;
  for (int half = 0; half < 2; ++half) {
    GemmParams<double>& p = w.gp[half];
    p.N = j_hi - j_lo;
    p.B[0] += (int64_t)j_lo * w.ldx; p.B[1] = p.B[0];
    p.colmap += j_lo;
  }
}

__global__ void dc_identity_kernel(double* Q, int64_t ldq, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Q[i + (int64_t)i * ldq] = 1.0;
}

}  // namespace

size_t stedc_scratch_bytes(int n) {
  size_t nn = (size_t)n * n;
  return 2 * nn * sizeof(double) + (size_t)n * (10 * sizeof(double) + 10 * sizeof(int)) + 4096 +
         (size_t)(n / 2 + 2) * (5 * sizeof(int) + 2 * sizeof(GemmParams<double>)) + 64 * 256;
}

// All eigenpairs of the symmetric tridiagonal (d, e): d <- eigenvalues ascending, Q <- eigenvectors.
// e is destroyed.  `scratch` must provide stedc_scratch_bytes(n).
// c_lo, c_hi (0-based, half open; 0, n = all): the eigenVECTOR columns the caller will read.  All n eigenvalues are always
// computed; columns of Q outside the range are undefined on exit (the root merge skips them: with m << n requested
// eigenpairs -- or a rank's share of them in the multi-GPU drivers -- the dominant GEMM shrinks by n / (c_hi - c_lo)).
int stedc_device(cudaStream_t s, int n, double* d, double* e, double* Q, int64_t ldq, void* scratch,
                 size_t scratch_bytes, int c_lo, int c_hi) {
  if (c_lo < 0) c_lo = 0;
  if (c_hi > n) c_hi = n;
  if (c_hi < c_lo) c_hi = c_lo;
  if (n <= 0) return 0;
  EIGB_CUDA_CHECK(cudaMemsetAsync(ctx().d_info + ST_STEDC, 0, sizeof(int), s));
  EIGB_CUDA_CHECK(cudaMemset2DAsync(Q, ldq * sizeof(double), 0, (size_t)n * sizeof(double), n, s));
  if (n == 1) {
    dc_identity_kernel<<<1, 32, 0, s>>>(Q, ldq, 1);
    EIGB_LAUNCH_CHECK();
    return 0;
  }
  if (scratch_bytes < stedc_scratch_bytes(n)) { set_last_error("stedc: scratch too small"); return -1; }
  DcWork w{};
  w.n = n;
  int levels = 0;
  while (((n + (1 << levels) - 1) >> levels) > LEAF) ++levels;
  w.levels = levels; w.L = 1 << levels;
  w.D = d; w.E = e; w.Q = Q; w.ldq = ldq;
  Arena ar(scratch, scratch_bytes);
  w.ldw = n; w.ldx = n;
  w.WS = ar.take<double>((size_t)n * n);
  w.X = ar.take<double>((size_t)n * n);
  w.Dsort = ar.take<double>(n); w.zsort = ar.take<double>(n); w.dlam = ar.take<double>(n); w.wz = ar.take<double>(n);
  w.zhat = ar.take<double>(n); w.Dnew = ar.take<double>(n); w.Ddef = ar.take<double>(n);
  w.rotc = ar.take<double>(n); w.rots = ar.take<double>(n); w.scale = ar.take<double>(8);
  w.idx = ar.take<int>(n); w.typ = ar.take<int>(n); w.srccol = ar.take<int>(n); w.posg = ar.take<int>(n);
  w.cmap = ar.take<int>(n); w.dstcol = ar.take<int>(n); w.rotp = ar.take<int>(n); w.rotn = ar.take<int>(n);
  w.permnd = ar.take<int>(n); w.permdf = ar.take<int>(n);
  w.K = ar.take<int>(5 * (size_t)(n / 2 + 2));
  w.sel = ar.take<int>(8);
  w.gp = ar.take<GemmParams<double>>(2 * (size_t)(n / 2 + 2));
  if (!w.gp) { set_last_error("stedc: scratch arena exhausted"); return -1; }
  w.status = ctx().d_info + ST_STEDC;

  ProfScope ps(PROF_STEDC, s);
  dc_scale_kernel<<<1, 1024, 0, s>>>(w);
  count_launch(1);
  dc_leaf_kernel<<<cdiv(w.L, 4), 128, 0, s>>>(w);
  EIGB_LAUNCH_CHECK();
  static OncePerDevice once;
  if (once.need()) {
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(dc_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         DEFL_SMEM_ELEMS * 16));
    once.done();
  }
  for (int level = 1; level <= levels; ++level) {
    const int nmerge = w.L >> level;
    int maxn = (int)(((long long)n * (1 << level) + w.L - 1) / w.L) + 1;
    if (maxn > n) maxn = n;
    const int use_smem = maxn <= DEFL_SMEM_ELEMS ? 1 : 0;
    dim3 gthr(cdiv(maxn, 256), nmerge), gwarp(cdiv(maxn, 8), nmerge), gcol(maxn, nmerge);
    dc_sort_kernel<<<gthr, 256, 0, s>>>(w, level);
    dc_deflate_kernel<<<dim3(1, nmerge), 256, use_smem ? (size_t)maxn * 16 : 0, s>>>(w, level, use_smem);
    dc_rotate_kernel<<<gthr, 256, 0, s>>>(w, level);
    dc_gather_kernel<<<gcol, 256, 0, s>>>(w, level);
    dc_secular_kernel<<<gwarp, 256, 0, s>>>(w, level);
    dc_rank_kernel<<<gthr, 256, 0, s>>>(w, level);
    const bool restricted = level == levels && (c_lo > 0 || c_hi < n);
    if (restricted) { dc_restrict_kernel<<<1, 32, 0, s>>>(w, c_lo, c_hi); count_launch(1); }
    dc_zhat_kernel<<<gwarp, 256, 0, s>>>(w, level);
    dc_formu_kernel<<<gwarp, 256, 0, s>>>(w, level, restricted ? 1 : 0);
    count_launch(7);
    EIGB_LAUNCH_CHECK();
    GemmParams<double> dummy{};
    if (gemm_launch<double>(s, false, true, dummy, w.gp, 2 * nmerge, maxn, maxn) != 0) return -1;
    dc_scatter_kernel<<<gcol, 256, 0, s>>>(w, level, restricted ? c_lo : 0, restricted ? c_hi : n);
    EIGB_LAUNCH_CHECK();
  }
  dc_unscale_kernel<<<cdiv(n, 256), 256, 0, s>>>(w);
  EIGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace eigb200
