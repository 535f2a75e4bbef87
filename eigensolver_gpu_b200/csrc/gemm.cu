// eigb200 -- DMMA (FP64 tensor core) GEMM kernel, sm_100a.  See gemm.cuh for the role of this file.
#include "gemm.cuh"

namespace eigb200 {

namespace {

constexpr int KPAD = 4;      // K-major smem row = BKT + 4 elements (conflict-free fragment reads)
// Two 8-warp CTAs per SM (<= 128 registers per thread): while one CTA waits at its k-tile barrier, runs its
// prologue or its epilogue, the other one keeps the DMMA pipe busy.  The cp.async ring gets as many stages as fit
// in half of the SM's shared memory (2..4).
constexpr int SMEM_PER_CTA = 115712;

template <typename T> struct Cfg;
template <> struct Cfg<double>  { static constexpr int BM = 128, BN = 64, WGM = 4, WGN = 2, PAD = 4, BKT = 16; };
template <> struct Cfg<double2> { static constexpr int BM = 64,  BN = 64, WGM = 2, WGN = 4, PAD = 2, BKT = 16; };
template <typename T, bool AK, bool BK> __host__ __device__ constexpr int stage_elems_() {
  using C_ = Cfg<T>;
  return (AK ? C_::BM * (C_::BKT + KPAD) : C_::BKT * (C_::BM + C_::PAD)) + (BK ? C_::BN * (C_::BKT + KPAD) : C_::BKT * (C_::BN + C_::PAD));
}
template <typename T, bool AK, bool BK> __host__ __device__ constexpr int num_stages_() {
  constexpr int n = SMEM_PER_CTA / (stage_elems_<T, AK, BK>() * (int)sizeof(T));
  return n > 4 ? 4 : (n < 2 ? 2 : n);
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
  }
}
// ---- mbarrier-tracked cp.async ring (no CTA-wide barrier per k-tile) -------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_addr(bar)) : "memory");
}
// this thread's outstanding cp.async operations arrive on the barrier when they complete (count pre-accounted: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_test(bar, parity)) { if (++spins > (1u << 26)) __trap(); }      // watchdog: fail, never hang
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// Loads one operand tile (EXT x BKT, EXT = BM or BN) into shared memory.
//  KMAJ=false: global element (x, k) at G[x + k*ld]  -> smem S[k*(EXT+PAD) + x]
//  KMAJ=true : global element (x, k) at G[k + x*ld]  -> smem S[x*(BKT+KPAD) + k]
// CH = elements per cp.async chunk (16 bytes when CH*sizeof(T)==16, else 8).
template <typename T, bool KMAJ, int EXT, int PAD, int CH, int NT, int BKT>
__device__ __forceinline__ void load_tile(T* S, const T* __restrict__ G, int64_t ld, int x0, int k0, int X, int K,
                                          int tid) {
  constexpr int CB = CH * (int)sizeof(T);
  if constexpr (!KMAJ) {
    constexpr int CPR = EXT / CH;             // chunks per k-row
    constexpr int TOTAL = CPR * BKT;
#pragma unroll
    for (int id = tid; id < TOTAL; id += NT) {
      int k = id / CPR, x = (id % CPR) * CH;
      int gx = x0 + x, gk = k0 + k;
      int valid = (gk < K) ? min(max(X - gx, 0), CH) : 0;
      const T* src = valid ? (G + gx + (int64_t)gk * ld) : G;
      cp_async<CB>(S + k * (EXT + PAD) + x, src, valid * (int)sizeof(T));
    }
  } else {
    constexpr int CPR = BKT / CH;             // chunks per x-row
    constexpr int TOTAL = CPR * EXT;
#pragma unroll
    for (int id = tid; id < TOTAL; id += NT) {
      int x = id / CPR, k = (id % CPR) * CH;
      int gx = x0 + x, gk = k0 + k;
      int valid = (gx < X) ? min(max(K - gk, 0), CH) : 0;
      const T* src = valid ? (G + gk + (int64_t)gx * ld) : G;
      cp_async<CB>(S + x * (BKT + KPAD) + k, src, valid * (int)sizeof(T));
    }
  }
}

template <typename T, bool AK, bool BK, int CH>
__global__ void __launch_bounds__(Cfg<T>::WGM * Cfg<T>::WGN * 32, 2) gemm_kernel(GemmParams<T> pv, const GemmParams<T>* __restrict__ dev) {
  using C_ = Cfg<T>;
  constexpr int BM = C_::BM, BN = C_::BN, WGM = C_::WGM, WGN = C_::WGN, PAD = C_::PAD, BKT = C_::BKT;
  constexpr int NT = WGM * WGN * 32;
  constexpr int WTM = BM / WGM, WTN = BN / WGN, MI = WTM / 8, NI = WTN / 8;
  constexpr int A_ELEMS = AK ? BM * (BKT + KPAD) : BKT * (BM + PAD);
  constexpr int B_ELEMS = BK ? BN * (BKT + KPAD) : BKT * (BN + PAD);
  constexpr int LDA_S = AK ? (BKT + KPAD) : (BM + PAD);
  constexpr int LDB_S = BK ? (BKT + KPAD) : (BN + PAD);
  constexpr bool CPLX = is_cplx<T>::value;
  constexpr int STAGES = num_stages_<T, AK, BK>();

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  // ring barriers: full[s] completes when every thread's cp.async writes into stage s have landed, empty[s] when all
  // warps have read it.  The warps of a CTA are never aligned by a CTA-wide barrier inside the k loop: a warp that is
  // ahead computes its next k-tile instead of waiting, which keeps the tensor pipe fed across k-tile boundaries.
  __shared__ __align__(8) uint64_t ring_full[4], ring_empty[4];

  GemmParams<T> p = dev ? dev[blockIdx.z] : pv;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= p.M || n0 >= p.N) return;
  if (p.mode == 1 && m0 > n0 + BN - 1 + p.diag_off) return;   // tile strictly below the diagonal

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) { mbar_init(&ring_full[st], NT); mbar_init(&ring_empty[st], NT / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int wm0 = (warp % WGM) * WTM, wn0 = (warp / WGM) * WTN;
  const int g = lane >> 2, t = lane & 3;

  const int KT0 = (p.K[0] + BKT - 1) / BKT;
  const int KT1 = (p.nseg > 1) ? (p.K[1] + BKT - 1) / BKT : 0;
  const int KT = KT0 + KT1;

  auto issue = [&](int kt) {
    if (kt < KT) {
      int seg = kt >= KT0;
      int k0 = (kt - (seg ? KT0 : 0)) * BKT;
      T* As = smem + (kt % STAGES) * (A_ELEMS + B_ELEMS);
      T* Bs = As + A_ELEMS;
      load_tile<T, AK, BM, PAD, CH, NT, BKT>(As, p.A[seg], p.lda[seg], m0, k0, p.M, p.K[seg], tid);
      load_tile<T, BK, BN, PAD, CH, NT, BKT>(Bs, p.B[seg], p.ldb[seg], n0, k0, p.N, p.K[seg], tid);
      cp_async_mbar_arrive(&ring_full[kt % STAGES]);
    }
  };
  // loads of k-tile j (j >= STAGES) reuse the stage of k-tile j - STAGES: all warps must have released it
  auto stage_free = [&](int j, bool block) -> bool {
    if (j >= KT) return true;                       // nothing to load
    uint64_t* bar = &ring_empty[j % STAGES];
    const unsigned par = (unsigned)((j / STAGES) - 1) & 1u;
    if (block) { mbar_wait(bar, par); return true; }
    unsigned ok = 0;
    if (lane == 0) ok = mbar_test(bar, par) ? 1u : 0u;
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
  };

  double acc[MI][NI][CPLX ? 4 : 2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int r = 0; r < (CPLX ? 4 : 2); ++r) acc[i][j][r] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  if (p.beta != 0.0 && p.colmap == nullptr) {
    // the epilogue reads the C tile: pull it into L2 now (128-byte lines, column by column)
    constexpr int LPC = BM * (int)sizeof(T) / 128;          // lines per tile column
    for (int id = tid; id < LPC * BN; id += NT) {
      const int cn = n0 + id / LPC, gm = m0 + (id % LPC) * (128 / (int)sizeof(T));
      if (cn < p.N && gm < p.M && !(p.mode == 1 && gm > cn + p.diag_off))
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p.C + gm + (int64_t)cn * p.ldc));
    }
  }

  for (int kt = 0; kt < KT; ++kt) {
    // refill the stage released one k-tile ago now if every warp is through with it, else after this k-tile's math
    const int jn = kt + STAGES - 1;
    const bool early = jn < STAGES || stage_free(jn, false);
    if (early) issue(jn);
    mbar_wait(&ring_full[kt % STAGES], (unsigned)(kt / STAGES) & 1u);
    const T* As = smem + (kt % STAGES) * (A_ELEMS + B_ELEMS);
    const T* Bs = As + A_ELEMS;
    double sa = 1.0, sb = 1.0;
    if constexpr (CPLX) { int seg = kt >= KT0; sa = p.sa[seg]; sb = p.sb[seg]; }
#pragma unroll
    for (int k4 = 0; k4 < BKT / 4; ++k4) {
      const int kk = k4 * 4 + t;
      if constexpr (!CPLX) {
        double a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          int m = wm0 + i * 8 + g;
          a[i] = AK ? As[m * LDA_S + kk] : As[kk * LDA_S + m];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          int n = wn0 + j * 8 + g;
          b[j] = BK ? Bs[n * LDB_S + kk] : Bs[kk * LDB_S + n];
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      } else {
        double are[MI], aim[MI], nim[MI], bre[NI], bim[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          int m = wm0 + i * 8 + g;
          double2 v = AK ? As[m * LDA_S + kk] : As[kk * LDA_S + m];
          are[i] = v.x; aim[i] = v.y * sa; nim[i] = -aim[i];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          int n = wn0 + j * 8 + g;
          double2 v = BK ? Bs[n * LDB_S + kk] : Bs[kk * LDB_S + n];
          bre[j] = v.x; bim[j] = v.y * sb;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            dmma(acc[i][j][0], acc[i][j][1], are[i], bre[j]);
            dmma(acc[i][j][0], acc[i][j][1], nim[i], bim[j]);
            dmma(acc[i][j][2], acc[i][j][3], are[i], bim[j]);
            dmma(acc[i][j][2], acc[i][j][3], aim[i], bre[j]);
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring_empty[kt % STAGES]);
    if (!early) { stage_free(jn, true); issue(jn); }
  }

  // epilogue: C = alpha*acc + beta*C.  All loads of a row group are issued before the first store: a store
  // followed by a load of another C element cannot be reordered by the compiler, and one DRAM round trip per
  // element (32 of them in a row) used to cost more than the main loop of a K = 64 update.
  const double alpha = p.alpha, beta = p.beta;
  const bool rd = beta != 0.0;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int gm = m0 + wm0 + i * 8 + g;
    T old[NI][2];
    T* cps[NI][2];
    bool ok[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gn = n0 + wn0 + j * 8 + 2 * t + r;
        ok[j][r] = gm < p.M && gn < p.N && !(p.mode == 1 && gm > gn + p.diag_off);
        const int cn = (ok[j][r] && p.colmap) ? p.colmap[gn] : gn;
        cps[j][r] = p.C + gm + (int64_t)cn * p.ldc;
        old[j][r] = (ok[j][r] && rd) ? *cps[j][r] : zero_<T>();
      }
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!ok[j][r]) continue;
        if constexpr (!CPLX) {
          double v = alpha * acc[i][j][r];
          if (rd) v += beta * old[j][r];
          *cps[j][r] = v;
        } else {
          double2 v = mkz(alpha * acc[i][j][r], alpha * acc[i][j][2 + r]);
          if (rd) { v.x += beta * old[j][r].x; v.y += beta * old[j][r].y; }
          const int gn = n0 + wn0 + j * 8 + 2 * t + r;
          if (p.real_diag && gm == gn + p.diag_off) v.y = 0.0;
          *cps[j][r] = v;
        }
      }
    }
  }
}

template <typename T, bool AK, bool BK, int CH>
int launch_one(cudaStream_t s, const GemmParams<T>& p, const GemmParams<T>* dev, int batch, int maxM, int maxN) {
  using C_ = Cfg<T>;
  constexpr int BKT = C_::BKT;
  constexpr int A_ELEMS = AK ? C_::BM * (BKT + KPAD) : BKT * (C_::BM + C_::PAD);
  constexpr int B_ELEMS = BK ? C_::BN * (BKT + KPAD) : BKT * (C_::BN + C_::PAD);
  constexpr int SMEM = num_stages_<T, AK, BK>() * (A_ELEMS + B_ELEMS) * (int)sizeof(T);
  static OncePerDevice once;
  auto kern = gemm_kernel<T, AK, BK, CH>;
  if (once.need()) {
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    once.done();
  }
  int M = dev ? maxM : p.M, N = dev ? maxN : p.N;
  if (M <= 0 || N <= 0) return 0;
  dim3 grid(cdiv(M, C_::BM), cdiv(N, C_::BN), batch);
  kern<<<grid, C_::WGM * C_::WGN * 32, SMEM, s>>>(p, dev);
  EIGB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
bool aligned16(const GemmParams<T>& p) {
  if (sizeof(T) == 16) return true;
  for (int s = 0; s < p.nseg; ++s) {
    if (((uintptr_t)p.A[s] & 15) || ((uintptr_t)p.B[s] & 15) || (p.lda[s] & 1) || (p.ldb[s] & 1)) return false;
  }
  return true;
}

}  // namespace

template <typename T>
int gemm_launch(cudaStream_t s, bool AK, bool BK, const GemmParams<T>& p, const GemmParams<T>* dev, int batch,
                int maxM, int maxN) {
  if (!dev) {
    if (p.M <= 0 || p.N <= 0) return 0;
    // host-side parameters: the TMA-fed kernel (gemm_tma.cu) unless alignment forbids tensor maps or the option is off
    const int r = gemm_launch_tma<T>(s, AK, BK, p);
    if (r <= 0) return r;
  }
  // device-side parameter blocks cannot be inspected here: they must be 16-byte aligned by construction for
  // complex; for real the 8-byte chunk path is used unless the caller-side template says otherwise.
  constexpr bool CP = is_cplx<T>::value;
  const bool vec = CP ? true : (dev ? false : aligned16(p));
#define EIGB_DISPATCH(AKv, BKv)                                                            \
  if (AK == AKv && BK == BKv) {                                                            \
    if (CP) return launch_one<T, AKv, BKv, 1>(s, p, dev, batch, maxM, maxN);               \
    if (vec) return launch_one<T, AKv, BKv, CP ? 1 : 2>(s, p, dev, batch, maxM, maxN);      \
    return launch_one<T, AKv, BKv, 1>(s, p, dev, batch, maxM, maxN);                       \
  }
  EIGB_DISPATCH(false, false)
  EIGB_DISPATCH(false, true)
  EIGB_DISPATCH(true, false)
  EIGB_DISPATCH(true, true)
#undef EIGB_DISPATCH
  return -1;
}

template <typename T> void gemm_tile_dims(int& bm, int& bn) { bm = Cfg<T>::BM; bn = Cfg<T>::BN; }
template void gemm_tile_dims<double>(int&, int&);
template void gemm_tile_dims<double2>(int&, int&);

template <typename T>
int gemm(cudaStream_t s, char ta, char tb, int M, int N, int K, double alpha, const T* A, int64_t lda, const T* B,
         int64_t ldb, double beta, T* C, int64_t ldc, int mode) {
  GemmParams<T> p{};
  p.M = M; p.N = N; p.nseg = 1;
  p.A[0] = A; p.lda[0] = lda; p.B[0] = B; p.ldb[0] = ldb; p.K[0] = K;
  p.A[1] = A; p.lda[1] = lda; p.B[1] = B; p.ldb[1] = ldb; p.K[1] = 0;
  p.sa[0] = (ta == 'C') ? -1.0 : 1.0; p.sb[0] = (tb == 'C') ? -1.0 : 1.0;
  p.sa[1] = p.sa[0]; p.sb[1] = p.sb[0];
  p.C = C; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.mode = mode; p.real_diag = 0;
  return gemm_launch<T>(s, ta != 'N', tb == 'N', p);
}

template <typename T>
int her2k_upper(cudaStream_t s, char trans, int n, int k, double alpha, const T* A, int64_t lda, const T* B,
                int64_t ldb, double beta, T* C, int64_t ldc) {
  GemmParams<T> p{};
  p.M = n; p.N = n; p.nseg = 2;
  p.A[0] = A; p.lda[0] = lda; p.B[0] = B; p.ldb[0] = ldb; p.K[0] = k;
  p.A[1] = B; p.lda[1] = ldb; p.B[1] = A; p.ldb[1] = lda; p.K[1] = k;
  if (trans == 'N') {       // A*B^H + B*A^H : op(A) plain, op(B) conj-transposed
    p.sa[0] = 1; p.sb[0] = -1; p.sa[1] = 1; p.sb[1] = -1;
  } else {                  // A^H*B + B^H*A
    p.sa[0] = -1; p.sb[0] = 1; p.sa[1] = -1; p.sb[1] = 1;
  }
  p.C = C; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.mode = 1; p.real_diag = 1;
  return gemm_launch<T>(s, trans != 'N', trans != 'N', p);
}

template <typename T>
int herk_upper(cudaStream_t s, char trans, int n, int k, double alpha, const T* A, int64_t lda, double beta, T* C,
               int64_t ldc) {
  GemmParams<T> p{};
  p.M = n; p.N = n; p.nseg = 1;
  p.A[0] = A; p.lda[0] = lda; p.B[0] = A; p.ldb[0] = lda; p.K[0] = k;
  p.A[1] = A; p.lda[1] = lda; p.B[1] = A; p.ldb[1] = lda; p.K[1] = 0;
  if (trans == 'N') { p.sa[0] = 1; p.sb[0] = -1; } else { p.sa[0] = -1; p.sb[0] = 1; }
  p.sa[1] = p.sa[0]; p.sb[1] = p.sb[0];
  p.C = C; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.mode = 1; p.real_diag = 1;
  return gemm_launch<T>(s, trans != 'N', trans != 'N', p);
}

#define EIGB_INST(T)                                                                                              \
  template int gemm_launch<T>(cudaStream_t, bool, bool, const GemmParams<T>&, const GemmParams<T>*, int, int, int); \
  template int gemm<T>(cudaStream_t, char, char, int, int, int, double, const T*, int64_t, const T*, int64_t,      \
                       double, T*, int64_t, int);                                                                 \
  template int her2k_upper<T>(cudaStream_t, char, int, int, double, const T*, int64_t, const T*, int64_t, double,  \
                              T*, int64_t);                                                                       \
  template int herk_upper<T>(cudaStream_t, char, int, int, double, const T*, int64_t, double, T*, int64_t);
EIGB_INST(double)
EIGB_INST(double2)
#undef EIGB_INST

}  // namespace eigb200
