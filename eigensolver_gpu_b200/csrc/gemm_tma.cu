// eigb200 -- TMA-fed DMMA GEMM (sm_100a): the host-parameter launches of the GEMM family (rank-2k trailing updates of
// hetrd/hegst/potrf, the recursive TRSM updates, the back-transformation GEMMs -- reference call sites zhetrd_gpu.F90:67,82,
// zhegst_gpu.F90:84-104, zheevd_gpu.F90:193-201, zhegvdx_gpu.F90:169).
//
//  * operands are staged with cp.async.bulk.tensor.2d (SASS UTMALDG) into a ring of shared-memory stages, every box
//    128 bytes wide (16 doubles) in the SWIZZLE_128B layout; out-of-bounds parts of a box read as zero, which replaces
//    all edge predication of the cp.async kernel (gemm.cu).  Issuing a stage is a dozen instructions of ONE warp (warp 0,
//    one lane per box) -- a dedicated producer warp would cost the 8 math warps their 128 registers (9 warps x 2 CTAs
//    leave 96 per thread: spills), so warp 0 refills the ring between its own k-tiles, early when the stage is already
//    free and after its math otherwise;
//  * 8 warps issue mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no f64 kind).  Which matrix row a fragment
//    lane holds is free as long as the epilogue uses the same map, so the rows are permuted per layout such that every
//    fragment load from the swizzled tile is bank-conflict free (derivation in DESIGN.md, "TMA-fed GEMM");
//  * full/empty mbarriers per stage, no CTA-wide barrier in the k loop; two CTAs per SM so that one CTA's epilogue
//    overlaps with the other's main loop.
// Device-parameter (batched, sizes decided on the device) and unaligned launches stay on the cp.async kernel.
#include "gemm.cuh"
#include "stages.cuh"
#include "tmap.cuh"
#include <string.h>

namespace eigb200 {

namespace {

constexpr int BKT = 16;                 // k-tile
constexpr int NCW = 8;                  // math warps (warp 0 also issues the TMA loads)
constexpr int NTH = NCW * 32;

template <typename T> struct TCfg;
template <> struct TCfg<double>  { static constexpr int BM = 128, BN = 64, WGM = 4, WGN = 2, DPE = 1, STAGES = 4; };
template <> struct TCfg<double2> { static constexpr int BM = 64,  BN = 64, WGM = 2, WGN = 4, DPE = 2, STAGES = 3; };


template <typename T>
struct TmaArgs {
  int M, N, KT0, KT1;
  double sa[2], sb[2];
  T* C; int64_t ldc;
  double alpha, beta;
  int mode, real_diag, diag_off;
  const int* colmap;
  int dbg;                 // debugging aid (option gemm_tma_dbg): 1 no early refill, 2 proxy fence before releasing a stage
};

__device__ __forceinline__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(s_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(s_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0, spins = 0;
  const unsigned a = s_addr(bar);
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();       // watchdog: a broken pipeline must fail, never hang the GPU
  } while (!ok);
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               :: "r"(s_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Local index (inside the warp tile) of the matrix row/column that fragment f, fragment lane-row g (0..7) stands for.
// Chosen per operand layout so that the 16 (real, 8-byte loads: per half-warp) / 8 (complex, 16-byte loads: per
// quarter-warp) lanes that the hardware serves together hit 32 distinct banks of the 128B-swizzled tile.
template <bool KMAJ, bool CPLX>
__device__ __forceinline__ int frag_x(int g, int f) {
  if constexpr (CPLX) return 8 * f + (g >> 1) + 4 * (g & 1);
  else if constexpr (!KMAJ) return 16 * (f >> 1) + 4 * (f & 1) + (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2);
  else return 16 * (f >> 1) + 2 * g + (f & 1);
}
// byte offset of element (x, k) of an operand tile (EXT x BKT) in its stage:
//   x-contiguous operand (KMAJ = false): boxes of {16 doubles of x} x {BKT k-lines},   line = k, 16-byte chunk = x
//   k-contiguous operand (KMAJ = true) : boxes of {16 doubles of k} x {EXT  x-lines},  line = x, 16-byte chunk = k
// SWIZZLE_128B stores chunk c of line l at chunk position c ^ (l & 7).
template <bool KMAJ, int EXT, int DPE>
__device__ __forceinline__ int elem_off(int x, int k) {
  if constexpr (!KMAJ) {
    const int xd = x * DPE;
    return (xd >> 4) * (BKT * 128) + k * 128 + ((((xd & 15) >> 1) ^ (k & 7)) << 4) + ((xd & 1) << 3);
  } else {
    const int kd = k * DPE;
    return (kd >> 4) * (EXT * 128) + x * 128 + ((((kd & 15) >> 1) ^ (x & 7)) << 4) + ((kd & 1) << 3);
  }
}

template <typename T, bool AK, bool BK>
__global__ void __launch_bounds__(NTH, 2) gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
                                                          const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                                                          const __grid_constant__ TmaArgs<T> p) {
  // Every tensor map is a kernel parameter of its own and every parameter is __grid_constant__ (no struct of maps, no
  // run-time index into one): a build that passed the maps as one struct next to a plain by-value argument block gave
  // wrong products now and then on long solves (n >= 10240), reproducibly per build and gone with this signature --
  // see DESIGN.md ("TMA-fed GEMM": what went wrong).
  using C_ = TCfg<T>;
  constexpr int BM = C_::BM, BN = C_::BN, WGM = C_::WGM, WGN = C_::WGN, DPE = C_::DPE, STAGES = C_::STAGES;
  constexpr bool CPLX = is_cplx<T>::value;
  constexpr int WTM = BM / WGM, WTN = BN / WGN, MI = WTM / 8, NI = WTN / 8;
  constexpr int A_BYTES = BM * BKT * (int)sizeof(T), B_BYTES = BN * BKT * (int)sizeof(T), STAGE_BYTES = A_BYTES + B_BYTES;
  // op(B)(k, n): BK = true -> B[k + n*ldb] (k-contiguous); BK = false -> B[n + k*ldb] (n-contiguous)
  constexpr bool A_KMAJ = AK, B_KMAJ = BK;
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));

  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (p.mode == 1 && m0 > n0 + BN - 1 + p.diag_off) return;   // tile strictly below the diagonal (CTA-uniform)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KT = p.KT0 + p.KT1;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();

  // ---- stage refill (warp 0): one lane per 128-byte-wide box
  constexpr int NBA = A_KMAJ ? (BKT * DPE) / 16 : (BM * DPE) / 16;
  constexpr int NBB = B_KMAJ ? (BKT * DPE) / 16 : (BN * DPE) / 16;
  constexpr int BOXA = A_KMAJ ? BM * 128 : BKT * 128;       // bytes per box
  constexpr int BOXB = B_KMAJ ? BN * 128 : BKT * 128;
  static_assert(NBA + NBB <= 32, "one lane per box");
  auto issue = [&](int kt) {                                 // warp 0 only; the stage is known to be free
    const int s = kt % STAGES;
    const int seg = kt >= p.KT0 ? 1 : 0;
    const int k0 = (kt - (seg ? p.KT0 : 0)) * BKT;
    unsigned char* As = smem + (size_t)s * STAGE_BYTES;
    unsigned char* Bs = As + A_BYTES;
    if (lane == 0) mb_expect_tx(&full[s], (unsigned)STAGE_BYTES);
    __syncwarp();
    const CUtensorMap* ma = seg ? &map_a1 : &map_a0;
    const CUtensorMap* mb = seg ? &map_b1 : &map_b0;
    if (lane < NBA) {
      if constexpr (A_KMAJ) tma_load(As + lane * BOXA, ma, k0 * DPE + 16 * lane, m0, &full[s]);
      else                  tma_load(As + lane * BOXA, ma, m0 * DPE + 16 * lane, k0, &full[s]);
    } else if (lane < NBA + NBB) {
      const int b = lane - NBA;
      if constexpr (B_KMAJ) tma_load(Bs + b * BOXB, mb, k0 * DPE + 16 * b, n0, &full[s]);
      else                  tma_load(Bs + b * BOXB, mb, n0 * DPE + 16 * b, k0, &full[s]);
    }
  };
  // k-tile j >= STAGES reuses the stage of k-tile j - STAGES: every warp must have released it
  auto stage_free = [&](int j, bool block) -> bool {
    uint64_t* bar = &empty[j % STAGES];
    const unsigned par = (unsigned)((j / STAGES) - 1) & 1u;
    if (block) { mb_wait(bar, par); return true; }
    unsigned ok = 0;
    if (lane == 0) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(s_addr(bar)), "r"(par) : "memory");
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
  };
  if (warp == 0) {
    for (int kt = 0; kt < STAGES - 1 && kt < KT; ++kt) issue(kt);
  }

  // ===================== consumer warps =====================
  const int wm0 = (warp % WGM) * WTM, wn0 = (warp / WGM) * WTN;
  const int g = lane >> 2, t = lane & 3;
  int xa[MI], xb[NI];                                  // rows / columns (inside the CTA tile) of this lane's fragments
#pragma unroll
  for (int i = 0; i < MI; ++i) xa[i] = wm0 + frag_x<A_KMAJ, CPLX>(g, i);
#pragma unroll
  for (int j = 0; j < NI; ++j) xb[j] = wn0 + frag_x<B_KMAJ, CPLX>(g, j);

  double acc[MI][NI][CPLX ? 4 : 2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int r = 0; r < (CPLX ? 4 : 2); ++r) acc[i][j][r] = 0.0;

  if (p.beta != 0.0 && p.colmap == nullptr) {
    // the epilogue reads the C tile: pull it into L2 now (128-byte lines, column by column)
    constexpr int LPC = BM * (int)sizeof(T) / 128;          // lines per tile column
    for (int id = tid; id < LPC * BN; id += NTH) {
      const int cn = n0 + id / LPC, gm = m0 + (id % LPC) * (128 / (int)sizeof(T));
      if (cn < p.N && gm < p.M && !(p.mode == 1 && gm > cn + p.diag_off))
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p.C + gm + (int64_t)cn * p.ldc));
    }
  }

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % STAGES;
    const int jn = kt + STAGES - 1;
    bool pending = false;
    if (warp == 0 && jn < KT) {
      if (jn < STAGES || (!(p.dbg & 1) && stage_free(jn, false))) issue(jn); else pending = true;
    }
    mb_wait(&full[s], (unsigned)(kt / STAGES) & 1u);
    const unsigned char* As = smem + (size_t)s * STAGE_BYTES;
    const unsigned char* Bs = As + A_BYTES;
    double sa = 1.0, sb = 1.0;
    if constexpr (CPLX) { const int seg = kt >= p.KT0; sa = p.sa[seg]; sb = p.sb[seg]; }
#pragma unroll
    for (int k4 = 0; k4 < BKT / 4; ++k4) {
      const int kk = k4 * 4 + t;
      if constexpr (!CPLX) {
        double a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(As + elem_off<A_KMAJ, BM, DPE>(xa[i], kk));
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = *reinterpret_cast<const double*>(Bs + elem_off<B_KMAJ, BN, DPE>(xb[j], kk));
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      } else {
        double are[MI], aim[MI], nim[MI], bre[NI], bim[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(As + elem_off<A_KMAJ, BM, DPE>(xa[i], kk));
          are[i] = v.x; aim[i] = v.y * sa; nim[i] = -aim[i];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const double2 v = *reinterpret_cast<const double2*>(Bs + elem_off<B_KMAJ, BN, DPE>(xb[j], kk));
          bre[j] = v.x; bim[j] = v.y * sb;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            dmma884(acc[i][j][0], acc[i][j][1], are[i], bre[j]);
            dmma884(acc[i][j][0], acc[i][j][1], nim[i], bim[j]);
            dmma884(acc[i][j][2], acc[i][j][3], are[i], bim[j]);
            dmma884(acc[i][j][2], acc[i][j][3], aim[i], bre[j]);
          }
      }
    }
    if (p.dbg & 2) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) mb_arrive(&empty[s]);              // this warp is done with the stage
    if (pending) { stage_free(jn, true); issue(jn); }
  }

  // epilogue: C = alpha*acc + beta*C.  The fragment lane (g, t) holds rows xa[i] and the columns that B's fragment rows
  // 2t, 2t+1 stand for.  All loads of a row group are issued before the first store (see gemm.cu).
  const double alpha = p.alpha, beta = p.beta;
  const bool rd = beta != 0.0;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int gm = m0 + xa[i];
    T old[NI][2];
    T* cps[NI][2];
    bool ok[NI][2];
    int gns[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gn = n0 + wn0 + frag_x<B_KMAJ, CPLX>(2 * t + r, j);
        gns[j][r] = gn;
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
          if (p.real_diag && gm == gns[j][r] + p.diag_off) v.y = 0.0;
          *cps[j][r] = v;
        }
      }
    }
  }
}

// tensor map over one operand: dim0 = the contiguous index in doubles, dim1 = the strided index; box = 16 doubles x rows
inline int make_operand_map(CUtensorMap* out, const void* base, uint64_t dim0_d, uint64_t dim1, uint64_t ld_bytes, uint32_t box1) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)ptr;
  }
  if (!fn) return -1;
  if (((uintptr_t)base & 15) || (ld_bytes & 15) || dim0_d == 0 || dim1 == 0) return -1;
  cuuint64_t gdim[2] = {dim0_d, dim1};
  cuuint64_t gstr[1] = {ld_bytes};
  cuuint32_t box[2] = {16, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

template <typename T, bool AK, bool BK>
int launch_tma(cudaStream_t s, const GemmParams<T>& p) {
  using C_ = TCfg<T>;
  constexpr int DPE = C_::DPE;
  constexpr int STAGE_BYTES = (C_::BM + C_::BN) * BKT * (int)sizeof(T);
  constexpr int SMEM = C_::STAGES * STAGE_BYTES + 1024;
  static OncePerDevice once;
  auto kern = gemm_tma_kernel<T, AK, BK>;
  if (once.need()) {
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    once.done();
  }
  struct { CUtensorMap a[2], b[2]; } maps;
  memset(&maps, 0, sizeof(maps));
  for (int seg = 0; seg < p.nseg; ++seg) {
    const int K = p.K[seg];
    if (K <= 0) continue;
    // op(A)(m, k): AK -> A[k + m*lda] (k-contiguous), else A[m + k*lda];  op(B)(k, n): BK -> B[k + n*ldb], else B[n + k*ldb]
    int rc;
    if (AK) rc = make_operand_map(&maps.a[seg], p.A[seg], (uint64_t)K * DPE, (uint64_t)p.M, (uint64_t)p.lda[seg] * sizeof(T), C_::BM);
    else    rc = make_operand_map(&maps.a[seg], p.A[seg], (uint64_t)p.M * DPE, (uint64_t)K, (uint64_t)p.lda[seg] * sizeof(T), BKT);
    if (rc != 0) return 1;
    if (BK) rc = make_operand_map(&maps.b[seg], p.B[seg], (uint64_t)K * DPE, (uint64_t)p.N, (uint64_t)p.ldb[seg] * sizeof(T), C_::BN);
    else    rc = make_operand_map(&maps.b[seg], p.B[seg], (uint64_t)p.N * DPE, (uint64_t)K, (uint64_t)p.ldb[seg] * sizeof(T), BKT);
    if (rc != 0) return 1;
  }
  TmaArgs<T> a;
  a.M = p.M; a.N = p.N;
  a.KT0 = p.K[0] > 0 ? (p.K[0] + BKT - 1) / BKT : 0;
  a.KT1 = (p.nseg > 1 && p.K[1] > 0) ? (p.K[1] + BKT - 1) / BKT : 0;
  a.sa[0] = p.sa[0]; a.sa[1] = p.sa[1]; a.sb[0] = p.sb[0]; a.sb[1] = p.sb[1];
  a.C = p.C; a.ldc = p.ldc; a.alpha = p.alpha; a.beta = p.beta;
  a.mode = p.mode; a.real_diag = p.real_diag; a.diag_off = p.diag_off; a.colmap = p.colmap;
  a.dbg = opts().gemm_tma_dbg;
  if (a.KT0 == 0 && a.KT1 > 0) {              // keep segment 0 non-empty (the kernel maps k-tiles [0, KT0) to segment 0)
    maps.a[0] = maps.a[1]; maps.b[0] = maps.b[1];
    a.KT0 = a.KT1; a.KT1 = 0; a.sa[0] = a.sa[1]; a.sb[0] = a.sb[1];
  }
  dim3 grid(cdiv(p.M, C_::BM), cdiv(p.N, C_::BN), 1);
  kern<<<grid, NTH, SMEM, s>>>(maps.a[0], maps.b[0], maps.a[1], maps.b[1], a);
  EIGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Returns 0 if the TMA kernel was launched, 1 if this product must use the cp.async kernel (alignment, option off),
// -1 on error.
template <typename T>
int gemm_launch_tma(cudaStream_t s, bool AK, bool BK, const GemmParams<T>& p) {
  if (!opts().gemm_tma) return 1;
  if (p.M <= 0 || p.N <= 0) return 0;
  if (p.K[0] <= 0 && !(p.nseg > 1 && p.K[1] > 0)) return 1;     // no k range: the plain kernel handles beta-only updates
  if (AK && BK) return launch_tma<T, true, true>(s, p);
  if (AK && !BK) return launch_tma<T, true, false>(s, p);
  if (!AK && BK) return launch_tma<T, false, true>(s, p);
  return launch_tma<T, false, false>(s, p);
}
template int gemm_launch_tma<double>(cudaStream_t, bool, bool, const GemmParams<double>&);
template int gemm_launch_tma<double2>(cudaStream_t, bool, bool, const GemmParams<double2>&);

}  // namespace eigb200
