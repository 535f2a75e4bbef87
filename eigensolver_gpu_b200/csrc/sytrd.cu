// eigb200 -- blocked Householder tridiagonalization (UPLO='U') and the symv/hemv tile engine, sm_100a.
//
// Replaces zhetrd_gpu/zlatrd_gpu (zhetrd_gpu.F90:30-165), dsytrd_gpu/dlatrd_gpu (dsytrd_gpu.F90:30-164), the
// per-column kernels K10-K16 (zhetrd_gpu.F90:211-879), zhemv_gpu/dsymv_gpu (zhemv_gpu.F90:33-193,
// dsymv_gpu.F90:33-150) and the final-block kernel zhetd2_gpu/dsytd2_gpu.
//
// Design (B200-first, not the reference's 4 launches per column with FP64 atomics):
//  * one persistent cooperative kernel per panel (1 CTA of 16 worker warps + 1 TMA producer warp per SM), 2 grid
//    barriers per column:
//      phase A (row-parallel): finish W(:,c+1) from the previous column's partial sums, store the reflector,
//              bring column c up to date with the panel's reflectors, partial norms.  The worker warps own a block
//              of rows; the producer warp (idle here) gathers and derives everything common to all rows.
//      phase B (tile phase): warp 0 derives the Householder scalars while the producer warp already streams 64x64
//              tiles of the upper triangle ONCE through a shared-memory ring filled by cp.async.bulk.tensor (TMA,
//              SASS UTMALDG, SWIZZLE_128B, L2 eviction hints); each tile serves A_IJ x_J and A_IJ^H x_I.  Tiles are
//              walked in column strips (transposed sums stay in registers for a strip, no CTA barrier per tile) and
//              handed out through an atomic queue in longest-processing-time order.  The z-dot units (V^H x, W^H x)
//              are complete per-CTA dot products that ride in front of the queue.
//  * the product runs on the unscaled column (no multiplication by the Householder scale per element); everything
//    that needs integer divisions is derived one column ahead by an otherwise idle thread (ColDesc).
//  * all cross-CTA reductions go through slots determined by the work unit and summed in a fixed order:
//    deterministic whatever CTA ran a unit, no FP64 atomics (the reference's results depend on atomicAdd ordering).
//  * w^H v is obtained algebraically (v^H A v - 2 Re(z1^H z2)), which removes a third barrier per column.
//  * the same panel code runs down to column 1, so no separate unblocked 32x32 kernel is needed.
//  * MG variant: tile columns distributed 1-D block-cyclically over P GPUs; phase A of the next column reduces the local
//    partial sums of its rows, delivers them to the same CTA of every peer (stores into peer-mapped buffers over NVLink,
//    one release flag per 32-row group) and adds the peers' shares in rank order: no extra phase, no third barrier.
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"
#include "tmap.cuh"
#include "unitmap.cuh"
#include <vector>
#include <string.h>

namespace eigb200 {

std::vector<unsigned long long>& trace_store() { static std::vector<unsigned long long> v; return v; }

namespace {

using tile::TB; using tile::MAXBANDS; using tile::ColDesc; using tile::UnitMap; using tile::strip_len; using tile::compute_desc;
constexpr int NT = 512;       // consumer/worker threads per CTA (16 warps: 4 tile rows each)
constexpr int NW = NT / 32;
constexpr int NTT = NT + 32;  // + one TMA producer warp
constexpr int NWT = NTT / 32;
constexpr int NBMAX = 128;    // max panel width
constexpr int TRSLOTS = 16;    // globaltimer stamps per column when tracing

// Ring of tile stages filled by TMA.  One stage = a 64x64 tile + the x slices of the tile's rows and columns.
// A tile arrives as NBOX boxes of 16 doubles x 64 columns (8 KB each, one cp.async.bulk.tensor.2d per box) in the
// 128-byte swizzled layout, which makes the "one lane per tile column" reads bank-conflict free.
template <typename T> struct RingCfg;
template <> struct RingCfg<double>  { static constexpr int STAGES = 6, NBOX = 4, DPE = 1; };
template <> struct RingCfg<double2> { static constexpr int STAGES = 3, NBOX = 8, DPE = 2; };
constexpr int BOX_BYTES = 128 * TB;
template <typename T> __host__ __device__ constexpr int stage_elems() { return TB * TB + 2 * TB; }
template <typename T> __host__ __device__ constexpr int ring_bytes() {
  return RingCfg<T>::STAGES * stage_elems<T>() * (int)sizeof(T) + 1024;    // + slack for 1024-byte alignment
}

template <typename T>
struct TrdP {
  T* A; int64_t lda;
  int i0, nbp;                 // panel = columns [i0, i0+nbp)
  T* W; int64_t ldw;           // W(:, c) <-> global column i0 + c
  double* d; double* e; T* tau;
  T* xbuf;                     // unscaled updated column
  T* Pd; T* Pt; int64_t ldp;   // partials: Pd[J*ldp + r] (direct, from tile (I(r),J), J > I), Pt[k*ldp + r] (band k)
  T* zfin;                     // [2][NBMAX]: z1 = V^H v, z2 = W^H v of the latest reflector
  double* npart;               // [G] partial sums of squares
  double* vavunit;             // [units] partial v^H A v, one slot per tile unit (deterministic whatever CTA ran it)
  T* alpha_slot;               // a(j-1, j) before scaling
  T* scale_slot;               // 1 / (alpha - beta) of the latest reflector
  double* beta_slot;           // [0] beta of the latest reflector, [1] the stored diagonal element A(j-1, j-1) its product saw
  unsigned* barrier;
  unsigned* qctr;              // [NBMAX] tile-unit queue heads, one per panel column (zeroed per panel)
  int* status;
  int use_tma;                 // tiles staged through the TMA ring (needs 16-byte aligned columns)
  int upc;                     // target number of tile units per CTA (strip length heuristic)
  const unsigned char* ctab;   // strip length per number of tile rows (host-built, tile::build_strip_table); nullptr: heuristic
  int keepI;                   // L2 residency: tile rows < keepI are loaded with evict_last (0: no cache hints)
  int npf;                     // tiles each CTA prefetches into L2 during phase A (0: off)
  int etrace_j; int64_t etrace_off;   // (profiling aid) per-CTA begin/end stamps of phase B for the product of order etrace_j
  unsigned long long* trace;   // optional: TRSLOTS globaltimer stamps per column (CTA 0), profiling aid
  // multi-GPU (1-D block-cyclic distribution of the 64-wide tile columns of the trailing matrix over P ranks)
  int rank, P;                 // P == 1: single GPU
  T* peer_w[8];                // peer_w[q] = rank q's exchange buffer [P][2][wstride] (peer-mapped, NVLink)
  unsigned long long* peer_flag[8];   // peer_flag[q] = rank q's arrival flags [P sources][fstride]: one per (CTA, 32-row group) + [fstride-1] for v^H A v
  int fstride;                 // flags per source rank
  int64_t wstride;             // elements per exchange slot (>= n + 2)
  unsigned long long seq_base; // sequence number of this panel's first column (flags are monotonic)
};

// per-thread pipeline state of the tile ring; persists across columns inside the cooperative kernel.
// Consumers: par bit s = parity to wait for on full[s] (starts 0).  Producer: par bit s = parity to wait for on
// empty[s] (starts 1: a fresh barrier counts as "previous phase complete").
struct RingState {
  int stage;
  unsigned par;
  unsigned strips;   // consumers: strip ends seen so far (phase of the single-buffered strip scratch barrier)
};

// what the producer warp tells the consumers about the tile in a ring stage
struct __align__(16) TileMeta { int I, J, unit, flags; };
constexpr int MF_FIRST = 1, MF_LAST = 2, MF_DIAG = 4, MF_END = 8;

// real: the strip scratch is double buffered (parity of the strip), which saves the second consumer barrier of a
// strip end; complex keeps one buffer (shared memory is taken by the ring) and two barriers
template <typename T> struct StripBuf { static constexpr int N = is_cplx<T>::value ? 1 : 2; };
template <typename T>
struct EngineSmem {
  T yt[StripBuf<T>::N][NW][TB];   // strip end: transposed partial sums of the 16 row-group warps
  T ydiag[StripBuf<T>::N][TB];    // D units: product of the diagonal tile with x_J (both triangles)
  double vred[StripBuf<T>::N][NW];   // strip end: per-warp v^H A v of the unit
  int bstart[MAXBANDS + 2];   // multi-GPU: prefix counts of the F units per band
};
template <typename T>
struct PhaseASmem {
  T z1[NBMAX], z2[NBMAX], rowV[NBMAX], rowW[NBMAX];
  T ared[NW * 32 * 2];   // per-warp slices of (wraw - t1, t2) for the rows in flight
  T s_tau, s_scale, s_wj;   // scalars of the previous reflector, from the scalar warp
  double s_alpha, s_beta;   // alpha' and beta of the previous reflector
};
// phase A and the tile engine never run at the same time: their scratch is overlaid
template <typename T>
struct PanelSmem {
  union U {
    PhaseASmem<T> a;
    EngineSmem<T> e;
    __device__ U() {}
  } u;
  T tred[NWT];
  double dscal[NWT];
  double nred[NWT];
  uint64_t full[8], empty[8];   // ring mbarriers
  TileMeta meta[8];
  ColDesc cd[2];                // descriptors of the current and the next product (parity of the panel column)
};

__device__ __forceinline__ double ldcg_(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg_(const double2* p) { return __ldcg(p); }

// ---- mbarrier / bulk-copy (TMA) wrappers ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  const unsigned a = smem_u32(bar);
  unsigned spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();     // watchdog: a broken pipeline must fail, never hang the GPU
  } while (!ok);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// same with an L2 eviction-priority policy (createpolicy)
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n"
               :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n"
               :: "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;\n" :: "n"(NT) : "memory"); }

// ---- grid barrier (monotonic counter, watchdog-protected) --------------------------------------------
template <class Hook>
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target, int* status, bool sys, Hook hook) {
  __syncthreads();
  if (threadIdx.x == 0) {
    hook();                      // last per-CTA global stores of the phase (e.g. the partial norm)
    // arrive: a release reduction (no return value to wait for; cumulative over the CTA's stores through the bar.sync above)
    if (sys) { __threadfence_system(); atomicAdd(bar, 1u); }
    else asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" :: "l"(bar) : "memory");
    unsigned long long spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (++spins > (1ull << 25)) { atomicExch(status, 77); __trap(); }   // watchdog: never hang the GPU
    }
  }
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red /* >= NWT entries */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = zero_<T>();
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) s = add_(s, red[w]);
  return s;
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
// Tile engine.  y = A x for the upper-stored Hermitian leading n x n block, as partial sums:
//   units:  F(k, J)  band k (tile rows [kC, kC+C)), tile column J >= (k+1)C : C off-diagonal tiles
//           D(J)     tile column J: off-diagonal tiles of the partial band floor(J/C) + the diagonal tile
//   outputs: Pd[J*ldp + r]  = (A_IJ x_J)(r)            for every off-diagonal tile (I, J), r in block I
//            Pt[k*ldp + c]  = sum over the unit's tiles of (A_IJ^H x_I)(c)  (+ the full diagonal-tile product for
//                             D units), c in block J, k = the unit's band
//            vavunit[u]     = the unit's share of Re(x^H A x)
//   so  y(r) = sum_{J > I(r)} Pd[J][r] + sum_{k <= floor(I(r)/C)} Pt[k][r].
// Every output slot is determined by the unit, not by the CTA that ran it, so the units can be handed out
// dynamically: unit `cta` is static, the following ones come from an atomic queue head (F units first, then the
// D units in decreasing size -- longest-processing-time order -- so that all CTAs finish within about one tile).
// One producer warp decodes units, announces every tile through a small descriptor in shared memory and (ring
// on) fetches it with TMA; 16 consumer warps follow the descriptors: warp w owns tile rows [4w, 4w+4), lane l owns
// tile columns l and l+32.
// =====================================================================================================
__device__ __forceinline__ void ring_init(uint64_t* full, uint64_t* empty, int stages, RingState& rs) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    mbar_init(&full[7], 3);       // "strip scratch free": the three reducing warps of a strip end arrive
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  rs.stage = 0;
  rs.par = (threadIdx.x >= NT) ? 0xffffffffu : 0u;
  rs.strips = 0;
}

// sum of 4 per-lane values over the 32 lanes of a warp with 6 shuffles per scalar: on return lane 8*h holds the
// total of v[h] (h = 0..3) in v[0]
__device__ __forceinline__ void warp_reduce4(double (&v)[4], int lane) {
  {
    const bool up = lane & 16;
    const double s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
    const double r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
    v[0] = (up ? v[2] : v[0]) + r0;
    v[1] = (up ? v[3] : v[1]) + r1;
  }
  {
    const bool up = lane & 8;
    const double s = up ? v[0] : v[1];
    const double r = __shfl_xor_sync(0xffffffffu, s, 8);
    v[0] = (up ? v[1] : v[0]) + r;
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ void warp_reduce4(double2 (&v)[4], int lane) {
  double re[4] = {v[0].x, v[1].x, v[2].x, v[3].x}, im[4] = {v[0].y, v[1].y, v[2].y, v[3].y};
  warp_reduce4(re, lane); warp_reduce4(im, lane);
  v[0] = mkz(re[0], im[0]);
}

// All NTT threads must call.  P > 1: tabulates the per-band unit counts of this rank's ownership pattern.
template <typename T>
__device__ __forceinline__ UnitMap engine_prepare(int n, int C, int rank, int P, EngineSmem<T>& es) {
  if (P > 1) {
    const int Tn = (n + TB - 1) / TB;
    const int KB = UnitMap::num_bands(Tn, C);
    UnitMap tmp; tmp.Tn = Tn; tmp.C = C; tmp.rank = rank; tmp.P = P;
    tmp.TnO = rank < Tn ? (Tn - rank + P - 1) / P : 0;
    for (int k = threadIdx.x; k < KB; k += blockDim.x) es.bstart[k + 1] = tmp.band_count(k);
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; es.bstart[0] = 0; for (int k = 1; k <= KB; ++k) { acc += es.bstart[k]; es.bstart[k] = acc; } }
    __syncthreads();
  }
  UnitMap um;
  um.init(n, C, rank, P, es.bstart);
  return um;
}

template <typename T>
struct ConsumerState {
  T acct[2];        // transposed sums of the strip (this lane's two tile columns, this warp's rows)
  double vav;       // the unit's share of Re(x^H A x)
  int par;          // strip parity (selects the strip scratch buffer when it is double buffered)
};

// reduction of NV per-lane values over the 32 lanes of a warp: on return lane (32/NV)*v holds the total of
// value v in v[0] (NV = 4: 6 shuffles per scalar, NV = 8: 9)
template <int NV>
__device__ __forceinline__ void warp_reduce_n(double (&v)[NV], int lane) {
  if constexpr (NV == 8) {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double s = up ? v[k] : v[k + 4];
      const double r = __shfl_xor_sync(0xffffffffu, s, 16);
      v[k] = (up ? v[k + 4] : v[k]) + r;
    }
    {
      const bool u8 = lane & 8;
      const double s0 = u8 ? v[0] : v[2], s1 = u8 ? v[1] : v[3];
      const double r0 = __shfl_xor_sync(0xffffffffu, s0, 8), r1 = __shfl_xor_sync(0xffffffffu, s1, 8);
      v[0] = (u8 ? v[2] : v[0]) + r0;
      v[1] = (u8 ? v[3] : v[1]) + r1;
    }
    {
      const bool u4 = lane & 4;
      const double s = u4 ? v[0] : v[1];
      const double r = __shfl_xor_sync(0xffffffffu, s, 4);
      v[0] = (u4 ? v[1] : v[0]) + r;
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  } else {
    double w[4] = {v[0], v[1], v[2], v[3]};
    warp_reduce4(w, lane);
    v[0] = w[0];
  }
}

// One pass of a consumer warp over NP (1 or 2) tiles of the same tile column.  m[t] = {I, J, unit, flags} of
// tile t, stg[t] its ring stage.  Warp w owns tile rows [4w, 4w+4); lane l owns tile columns l and l+32.
template <typename T, int NP, class XR>
__device__ __forceinline__ void process_tiles(ConsumerState<T>& cs, const int4 (&m)[NP], const int (&stg)[NP],
                                              const T* __restrict__ A, int64_t lda, int n, const T* __restrict__ xsrc,
                                              XR xfix, T* Pd, T* Pt, int64_t ldp, double* vavunit, const UnitMap& um,
                                              bool tma, T* ring, uint64_t* empty, EngineSmem<T>& es, uint64_t* ytfree,
                                              RingState& rs) {
  constexpr int DPE = RingCfg<T>::DPE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r4 = 4 * warp;
  const int J = m[0].y;
  constexpr int NBUF = StripBuf<T>::N;
  const int pb = NBUF > 1 ? cs.par : 0;
  if (m[0].w & MF_FIRST) {
    cs.acct[0] = zero_<T>(); cs.acct[1] = zero_<T>(); cs.vav = 0.0;
    // single-buffered strip scratch (ydiag, yt, vred): the readers of the previous strip end must be done; they
    // arrived long ago in practice, so this costs one already-complete try_wait instead of a second CTA-wide barrier
    if (NBUF == 1) mbar_wait(ytfree, (rs.strips & 1u) ^ 1u);
  }
  T a[NP][4][2], xr[NP][4], xc[2];
  if (tma) {
#pragma unroll
    for (int t = 0; t < NP; ++t) {
      const T* tile = ring + (size_t)stg[t] * stage_elems<T>();
      // this warp's 4 rows live in box (4w*DPE)/16, 16-byte chunks k0.. of each 128-byte column line; the
      // SWIZZLE_128B layout stores chunk k of line c at chunk position k ^ (c & 7)
      const char* box = reinterpret_cast<const char*>(tile) + ((r4 * DPE) / 16) * BOX_BYTES;
      const int k0 = ((r4 * DPE) % 16) / 2;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
        const char* line = box + c * 128;
        if constexpr (!is_cplx<T>::value) {
          const double2 v0 = *reinterpret_cast<const double2*>(line + (((k0) ^ (c & 7)) << 4));
          const double2 v1 = *reinterpret_cast<const double2*>(line + (((k0 + 1) ^ (c & 7)) << 4));
          a[t][0][q] = v0.x; a[t][1][q] = v0.y; a[t][2][q] = v1.x; a[t][3][q] = v1.y;
        } else {
#pragma unroll
          for (int h = 0; h < 4; ++h) a[t][h][q] = *reinterpret_cast<const double2*>(line + (((k0 + h) ^ (c & 7)) << 4));
        }
        if (t == 0) xc[q] = xfix(J * TB + c, tile[TB * TB + TB + c]);
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) xr[t][h] = xfix(m[t].x * TB + r4 + h, tile[TB * TB + r4 + h]);
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < NP; ++t) mbar_arrive(&empty[stg[t]]);          // this warp is done with the stage(s)
    }
#pragma unroll
    for (int t = 0; t < NP; ++t) {
      if (m[t].w & MF_DIAG) {
        // the box also brought the (ignored) lower triangle: keep r < c, make the diagonal real
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int r = r4 + h, c = lane + 32 * q;
            if (r > c) a[t][h][q] = zero_<T>();
            else if (r == c) a[t][h][q] = from_real<T>(real_(a[t][h][q]));
          }
      }
    }
  } else {
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < NP; ++t) mbar_arrive(&empty[stg[t]]);          // the stage carried only the descriptor
    }
    // masked loads straight from global memory
#pragma unroll
    for (int t = 0; t < NP; ++t) {
      const int I = m[t].x;
      const bool diag = (m[t].w & MF_DIAG) != 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q, gc = J * TB + c;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int r = r4 + h, gr = I * TB + r;
          const bool ok = gr < n && gc < n && (!diag || r <= c);
          a[t][h][q] = ok ? __ldg(A + gr + (int64_t)gc * lda) : zero_<T>();
          if (diag && r == c) a[t][h][q] = from_real<T>(real_(a[t][h][q]));    // Hermitian: real diagonal
        }
        if (t == 0) xc[q] = xfix(gc, gc < n ? xsrc[gc] : zero_<T>());
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) { const int gr = I * TB + r4 + h; xr[t][h] = xfix(gr, gr < n ? xsrc[gr] : zero_<T>()); }
    }
  }
  // ---- products: direct (rows) and transposed (columns)
  T accd[NP][4];
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const bool diag = (m[t].w & MF_DIAG) != 0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      accd[t][h] = zero_<T>();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        fma_(accd[t][h], a[t][h][q], xc[q]);
        if (!diag || (r4 + h) < (lane + 32 * q)) fmac_(cs.acct[q], a[t][h][q], xr[t][h]);
      }
      // v^H A v: conj(x_r) (A x)_r; an off-diagonal tile also stands for its mirror image (done after the lane
      // reduction below, on the row totals), and so does the strictly upper part of a diagonal tile
      if (diag) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          T u = zero_<T>(); fma_(u, a[t][h][q], xc[q]);
          T tt = zero_<T>(); fmac_(tt, xr[t][h], u);
          cs.vav += ((r4 + h) < (lane + 32 * q) ? 2.0 : 1.0) * real_(tt);
        }
      }
    }
  }
  // ---- direct sums: reduce over the lanes, store (off-diagonal tile: all 64 rows are < n)
  if constexpr (NP == 2) {
    double v[8];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int h = 0; h < 4; ++h) v[t * 4 + h] = real_(accd[t][h]);
    warp_reduce_n<8>(v, lane);
    if ((lane & 3) == 0) {
      const int t = lane >> 4, h = (lane >> 2) & 3;
      const int4 mt = t ? m[NP - 1] : m[0];
      if (!(mt.w & MF_DIAG)) {
        Pd[(int64_t)J * ldp + mt.x * TB + r4 + h] = from_real<T>(v[0]);
        T xh = xr[0][0];
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) if (tt == t && hh == h) xh = xr[tt][hh];
        cs.vav += 2.0 * real_(xh) * v[0];
      } else {
        es.ydiag[pb][r4 + h] = from_real<T>(v[0]);
      }
    }
  } else {
    warp_reduce4(accd[0], lane);
    if ((lane & 7) == 0) {
      const int h = lane >> 3;
      if (!(m[0].w & MF_DIAG)) {
        Pd[(int64_t)J * ldp + m[0].x * TB + r4 + h] = accd[0][0];
        T xh = xr[0][0];
#pragma unroll
        for (int hh = 1; hh < 4; ++hh) if (hh == h) xh = xr[0][hh];
        T tt = zero_<T>(); fmac_(tt, xh, accd[0][0]);
        cs.vav += 2.0 * real_(tt);
      } else {
        es.ydiag[pb][r4 + h] = accd[0][0];
      }
    }
  }
  const int4 ml = m[NP - 1];
  if (ml.w & MF_LAST) {
    // strip end: combine the transposed sums of the 16 warps (+ the diagonal tile's direct part) -> Pt[band]
    const bool diag = (ml.w & MF_DIAG) != 0;
    es.yt[pb][warp][lane] = cs.acct[0];
    es.yt[pb][warp][lane + 32] = cs.acct[1];
    const double vs = warp_sum(cs.vav);
    if (lane == 0) es.vred[pb][warp] = vs;
    consumer_barrier();
    if (tid < TB) {
      T s = diag ? es.ydiag[pb][tid] : zero_<T>();
#pragma unroll
      for (int w = 0; w < NW; ++w) s = add_(s, es.yt[pb][w][tid]);
      const int band = ((diag ? J : ml.x) * um.rcpC) >> 16;
      if (J * TB + tid < n) Pt[(int64_t)band * ldp + J * TB + tid] = s;
    } else if (tid == TB) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += es.vred[pb][w];
      vavunit[ml.z] = s;
    }
    // single buffer: the readers signal the "scratch free" barrier, checked at the start of the next unit.  Double
    // buffer: this buffer is written again two strip ends from now, i.e. after the readers have gone through the
    // next strip's barrier.
    if (NBUF == 1 && tid < TB + 32) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(ytfree);
    }
    rs.strips += 1;
    cs.par ^= 1;
  }
}

// XR: (global index r, raw x value) -> x(r) (must return 0 for r >= n); xsrc: x stored with at least
// roundup64(n) readable entries.  All NTT threads must call; ends with a CTA barrier.
// PH: called once by the producer warp (all 32 lanes) after its first ring-full of tile fetches is in flight (or at
// the end if it had fewer): work that must happen during the tile phase without holding up anybody's tiles.
template <typename T, class XR, class PH>
__device__ void engine_run(const UnitMap& um, const T* __restrict__ A, int64_t lda, int n, const T* __restrict__ xsrc,
                           XR xfix, T* Pd, T* Pt, int64_t ldp, double* vavunit, unsigned* qctr, int cta, int G, bool tma,
                           T* ring, uint64_t* full, uint64_t* empty, TileMeta* meta, RingState& rs, EngineSmem<T>& es,
                           const CUtensorMap* tmap, ColDesc* next_cd, int jnext, int Pdesc, int upc, const unsigned char* ctab,
                           PH producer_hook) {
  constexpr int S = RingCfg<T>::STAGES, NBOX = RingCfg<T>::NBOX, DPE = RingCfg<T>::DPE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int st = rs.stage;

  if (warp >= NW) {
    // ===================== producer warp =====================
    // L2 residency: the tile stream of one column (up to 0.5 GB) would evict everything else -- the partial sums,
    // V and W that phase A reads back -- and itself before the next column comes round.  Tiles of the top tile rows
    // (a fixed subset that fits in L2) are kept with evict_last, the rest streams through with evict_first.
    uint64_t pol_keep = 0, pol_stream = 0;
    if (um.keepI > 0) {
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol_keep));
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol_stream));
    }
    int unit = cta;                       // first unit static, the following ones from the queue
    int issued = 0;
    for (;;) {
      const bool end = unit >= um.total;
      int J = 0, I0 = 0, I1 = 0; bool hd = false;
      int nu = 0;
      if (!end) {
        // take the next unit id now: the round trip of the atomic overlaps with this unit's tiles
        if (lane == 0) nu = G + (int)atomicAdd(qctr, 1u);
        um.decode(unit, J, I0, I1, hd);
      }
      const int ntile = end ? 1 : (I1 - I0) + (hd ? 1 : 0);
      for (int t = 0; t < ntile; ++t) {
        const bool diag = !end && hd && (t == ntile - 1);
        const int I = diag ? J : I0 + t;
        mbar_wait(&empty[st], (rs.par >> st) & 1u);
        rs.par ^= (1u << st);
        T* dst = ring + (size_t)st * stage_elems<T>();
        if (lane == 0) {
          TileMeta m;
          m.I = I; m.J = J; m.unit = unit;
          m.flags = end ? MF_END : ((t == 0 ? MF_FIRST : 0) | (t == ntile - 1 ? MF_LAST : 0) | (diag ? MF_DIAG : 0));
          meta[st] = m;
          if (end || !tma) mbar_arrive(&full[st]);
          else mbar_expect_tx(&full[st], (unsigned)(stage_elems<T>() * sizeof(T)));
        }
        __syncwarp();
        if (!end && tma) {
          if (lane < NBOX) {
            if (um.keepI > 0)
              tma_load_2d_hint(reinterpret_cast<char*>(dst) + lane * BOX_BYTES, tmap, (I * TB) * DPE + lane * 16, J * TB, &full[st],
                               I < um.keepI ? pol_keep : pol_stream);
            else
              tma_load_2d(reinterpret_cast<char*>(dst) + lane * BOX_BYTES, tmap, (I * TB) * DPE + lane * 16, J * TB, &full[st]);
          }
          else if (lane == NBOX) bulk_copy_g2s(dst + TB * TB, xsrc + I * TB, (unsigned)(TB * sizeof(T)), &full[st]);
          else if (lane == NBOX + 1) bulk_copy_g2s(dst + TB * TB + TB, xsrc + J * TB, (unsigned)(TB * sizeof(T)), &full[st]);
        }
        st = (st + 1) % S;
        if (!end && ++issued == S) producer_hook();
      }
      if (end) break;
      unit = __shfl_sync(0xffffffffu, nu, 0);
    }
    if (issued < S) producer_hook();
    // idle from here on: derive the next product's descriptor while the consumers drain the ring
    if (lane == 0 && next_cd != nullptr) compute_desc(*next_cd, jnext, G, Pdesc, upc, ctab);
  } else {
    // ===================== consumer warps =====================
    ConsumerState<T> cs;
    cs.acct[0] = zero_<T>(); cs.acct[1] = zero_<T>(); cs.vav = 0.0; cs.par = 0;
    for (;;) {
      mbar_wait(&full[st], (rs.par >> st) & 1u);
      rs.par ^= (1u << st);
      const int4 m0 = *reinterpret_cast<const int4*>(&meta[st]);
      const int st0 = st;
      st = (st + 1) % S;
      if (m0.w & MF_END) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st0]);
        break;
      }
      if constexpr (!is_cplx<T>::value) {
        if (tma && !(m0.w & MF_LAST)) {
          // real tiles are half the size of complex ones and the per-tile costs (descriptor, shuffle reduction,
          // stores) would dominate: take the next tile of the unit (same tile column) in the same pass
          mbar_wait(&full[st], (rs.par >> st) & 1u);
          rs.par ^= (1u << st);
          const int4 m1 = *reinterpret_cast<const int4*>(&meta[st]);
          const int st1 = st;
          st = (st + 1) % S;
          const int4 mm[2] = {m0, m1};
          const int ss[2] = {st0, st1};
          process_tiles<T, 2>(cs, mm, ss, A, lda, n, xsrc, xfix, Pd, Pt, ldp, vavunit, um, tma, ring, empty, es, &full[7], rs);
          continue;
        }
      }
      const int4 mm[1] = {m0};
      const int ss[1] = {st0};
      process_tiles<T, 1>(cs, mm, ss, A, lda, n, xsrc, xfix, Pd, Pt, ldp, vavunit, um, tma, ring, empty, es, &full[7], rs);
    }
  }
  rs.stage = st;
  __syncthreads();
}

// sum of the partials belonging to row r of an order-n product computed with strip length C
template <typename T>
__device__ __forceinline__ T gather_partials(const T* Pd, const T* Pt, int64_t ldp, int n, int C, int r) {
  const int Tn = (n + TB - 1) / TB, I = r / TB;
  T s0 = zero_<T>(), s1 = zero_<T>();
  for (int J = I + 1; J < Tn; ++J) s0 = add_(s0, ldcg_(Pd + (int64_t)J * ldp + r));
  for (int k = 0; k <= I / C; ++k) s1 = add_(s1, ldcg_(Pt + (int64_t)k * ldp + r));
  return add_(s0, s1);
}

// ---- multi-GPU exchange helpers ---------------------------------------------------------------------------
// The exchanged words validate themselves: a slot holds the all-ones bit pattern (a NaN no computation produces) until the
// sender's plain 8-byte stores land; the reader polls the data words (relaxed system-scope loads served by L2) and puts
// the pattern back after use.  No flags and no system-scope fences (measured at ~4 us each on NVLink) on the critical path.
__device__ __forceinline__ bool unset_(double v) { return __double_as_longlong(v) == -1ll; }
__device__ __forceinline__ bool unset_(double2 v) { return unset_(v.x) || unset_(v.y); }
template <typename T> __device__ __forceinline__ T unset_value();
template <> __device__ __forceinline__ double unset_value<double>() { return __longlong_as_double(-1ll); }
template <> __device__ __forceinline__ double2 unset_value<double2>() { return mkz(__longlong_as_double(-1ll), __longlong_as_double(-1ll)); }
__device__ __forceinline__ double ld_poll(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_poll(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double shfl_(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double2 shfl_(double2 v, int src) {
  return mkz(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// sequence number of the exchange that follows the product of panel column c
template <typename T>
__device__ __forceinline__ unsigned long long col_seq(const TrdP<T>& p, int c) {
  return p.seq_base + (unsigned long long)(p.nbp - 1 - c) + 1ull;
}
// this rank's slot for source rank q and parity par in the exchange buffer of rank `dst`
template <typename T>
__device__ __forceinline__ T* ex_slot(const TrdP<T>& p, int dst, int q, unsigned par) {
  return p.peer_w[dst] + ((int64_t)q * 2 + par) * p.wstride;
}

// profiling aid: CTA 0 / thread 0 records a globaltimer stamp in slot k of panel column c
template <typename T>
__device__ __forceinline__ void tstamp(const TrdP<T>& p, int c, int k) {
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && c >= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(size_t)(p.i0 + c) * TRSLOTS + k] = t;
  }
}

// ---- stand-alone symv/hemv (eigb200_dsymv / eigb200_zhemv): the same engine + a gather kernel -------------
template <typename T>
__global__ void __launch_bounds__(NTT, 1) hemv_tiles_kernel(const __grid_constant__ CUtensorMap tmap,
                                                            const T* __restrict__ A, int64_t lda, int n,
                                                            const T* __restrict__ xpad, T* Pd, T* Pt, int64_t ldp, int C,
                                                            int tma, double* vavunit, unsigned* qctr) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ EngineSmem<T> es;
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ TileMeta meta[8];
  T* ring = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
  RingState rs;
  ring_init(full, empty, RingCfg<T>::STAGES, rs);
  auto xfix = [n](int r, T raw) -> T { return r < n ? raw : zero_<T>(); };
  const UnitMap um = engine_prepare<T>(n, C, 0, 1, es);
  engine_run<T>(um, A, lda, n, xpad, xfix, Pd, Pt, ldp, vavunit, qctr, blockIdx.x, gridDim.x, tma != 0, ring, full, empty,
                meta, rs, es, &tmap, nullptr, 0, 1, 6, nullptr, []() {});
}
template <typename T>
__global__ void hemv_reduce_kernel(const T* Pd, const T* Pt, int64_t ldp, int n, int C, T* y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) y[r] = gather_partials<T>(Pd, Pt, ldp, n, C, r);
}
template <typename T>
__global__ void pad_copy_kernel(const T* x, int n, T* xpad, int npad) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < npad) xpad[r] = r < n ? x[r] : zero_<T>();
}

// =====================================================================================================
// Panel phases
// =====================================================================================================
// Phase A (row-parallel): with the partial sums of the previous column's product (order jp = j+1) complete,
//   * finish  W(:, c+1) = tau (w_raw - W z1 - V z2) + alpha' v   and store the reflector v itself in A(:, j+1),
//   * bring column c (global j) up to date with the panel's reflectors, write it unscaled to xbuf, partial norms.
// This phase is bound by instruction issue and L2 latency, not by bandwidth, so the work is split by role:
//   * the 16 worker warps own one contiguous block of rows per CTA (one row per lane, the warps of a row group
//     split the partial-sum slots and the panel columns) and issue all their independent loads up front;
//   * warp 16 (the TMA producer, idle here) gathers everything that is common to all rows -- z1, z2, row j of V
//     and W, the slots of row j, tau, scale -- and derives rho, alpha', W(j, c+1) while the workers run their
//     V/W loop; the results travel through shared memory.
template <typename T, bool MG>
__device__ void phase_a(const TrdP<T>& p, int c, PanelSmem<T>& sm, const ColDesc& cd, const ColDesc& cdn,
                        const CUtensorMap* tmap, int units_prev) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int nbp = p.nbp;
  const int j = p.i0 + c;                       // column brought up to date in this phase (c may be -1)
  const int jp = j + 1;                         // order of the product done for column c+1
  const int cprev = c + 1;
  const bool have_prev = (cprev <= nbp - 1) && (jp >= 1);
  const bool mg = MG && p.P > 1;
  const int Pn = mg ? p.P : 1, rk = mg ? p.rank : 0;
  PhaseASmem<T>& S = sm.u.a;
  const unsigned long long seqp = have_prev ? col_seq(p, cprev) : 0ull;
  const unsigned parp = (unsigned)(seqp & 1ull);
  // ---- rows [rbeg, rend) of this CTA in groups of 32 (one row per lane); GC = 2^gcs groups run side by side,
  //      each split over WPG = 16 / GC warps
  const int nrows = jp > 0 ? jp : 0;
  int R;
  if (have_prev) R = cd.R;
  else { R = (nrows + G - 1) / G; R = (R + 7) & ~7; }          // first column of a panel: once per launch
  const int Tn = cd.Tn;                                        // (meaningful when have_prev)
  const int rbeg = cta * R;
  const int rend = nrows < rbeg + R ? nrows : rbeg + R;
  const int ngrp = rend > rbeg ? (rend - rbeg + 31) >> 5 : 0;
  // multi-GPU: the rows are dealt to the CTAs in the same way on every rank (R depends on the order and the grid only), so
  // CTA i of rank q delivers its rows' partial sums to CTA i of every other rank
  const int gcs = ngrp <= 1 ? 0 : (ngrp == 2 ? 1 : 2);
  const int GC = 1 << gcs, wpgs = 4 - gcs, WPG = 1 << wpgs;
  const int grp = warp >> wpgs, sub = warp & (WPG - 1);        // warp 16: grp == GC -> no row work

  double nrm = 0.0;
  for (int g0 = 0, round = 0; round == 0 || g0 < ngrp; g0 += GC, ++round) {
    const int r = rbeg + (g0 + grp) * 32 + lane;
    const bool rv = warp < NW && (g0 + grp) < ngrp && r < rend;
    T wr = zero_<T>(), xb = zero_<T>(), acol = zero_<T>();
    T pv[4], pw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { pv[k] = zero_<T>(); pw[k] = zero_<T>(); }
    if (round > 0) __syncthreads();              // ared is reused
    if (warp == NW) {
      // ===================== scalar warp (round 0 only) =====================
      if (round == 0 && have_prev) {
        const T tau_p = ldcg_(p.tau + j);        // written by CTA 0 in the previous phase B (before a barrier)
        const T scale_p = ldcg_(p.scale_slot);
        const double beta_p = __ldcg(p.beta_slot), ajj_p = __ldcg(p.beta_slot + 1);
        T a1[4], a2[4], a3[4], a4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cc = cprev + 1 + lane + 32 * k;
          a1[k] = a2[k] = a3[k] = a4[k] = zero_<T>();
          if (cc < nbp) {
            a3[k] = ldcg_(p.A + j + (int64_t)(p.i0 + cc) * p.lda); a4[k] = ldcg_(p.W + j + (int64_t)cc * p.ldw);
            // z = scale * (.)^H x~ = scale * ((.)^H x - beta conj(row j of (.)))
            a1[k] = mul_(scale_p, sub_(ldcg_(p.zfin + cc), scale_(conj_(a3[k]), beta_p)));
            a2[k] = mul_(scale_p, sub_(ldcg_(p.zfin + NBMAX + cc), scale_(conj_(a4[k]), beta_p)));
          }
        }
        // partial-sum slots of row j (multi-GPU: the P exchange slots, after sync #1)
        T wj = zero_<T>();
        if (mg && cta == G - 1) {
          // This rank's share of (A x)(j), delivered to every rank (its own copy included) by the scalar warp of the last
          // CTA.  It must not come from the worker warp that owns row j: with more than four row groups per CTA (orders
          // above ~19000) that warp reaches row j in a later round, i.e. after the CTA barrier at which the scalar warps
          // -- this CTA's too -- wait for this very value.
          const int Ij = j >> 6;
          int J0 = Ij + 1;
          J0 += ((rk - J0) % Pn + Pn) % Pn;
          const int nd = J0 < Tn ? (Tn - J0 + Pn - 1) / Pn : 0;
          const int nt = (Ij % Pn == rk) ? ((Ij * cd.rcpC) >> 16) + 1 : 0;
          T wl = zero_<T>();
          for (int q0 = lane; q0 < nd + nt; q0 += 128) {
            T v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int q = q0 + 32 * k;
              const T* src = (q < nd) ? (p.Pd + (int64_t)(J0 + q * Pn) * p.ldp) : (p.Pt + (int64_t)(q - nd) * p.ldp);
              v[k] = q < nd + nt ? ldcg_(src + j) : zero_<T>();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) wl = add_(wl, v[k]);
          }
          wl = warp_sum(wl);
          if (lane < p.P) ex_slot(p, lane, p.rank, parp)[p.wstride - 2] = wl;
        }
        if (!mg) {
          const int ndj = cd.ndj, nsj = cd.nsj, Ij = j >> 6;
          for (int q0 = lane; q0 < nsj; q0 += 128) {
            T v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int q = q0 + 32 * k;
              const T* src = (q < ndj) ? (p.Pd + (int64_t)(Ij + 1 + q) * p.ldp) : (p.Pt + (int64_t)(q - ndj) * p.ldp);
              v[k] = q < nsj ? ldcg_(src + j) : zero_<T>();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) wj = add_(wj, v[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cc = cprev + 1 + lane + 32 * k;
          if (cc < nbp) { S.z1[cc] = a1[k]; S.z2[cc] = a2[k]; S.rowV[cc] = a3[k]; S.rowW[cc] = a4[k]; }
        }
        // z1^H z2 and the row-j correction  W(j,:) z1 + V(j,:) z2
        double zz = 0.0;
        T part = zero_<T>();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          T t = zero_<T>(); fmac_(t, a1[k], a2[k]);
          zz += real_(t);
          fma_(part, a4[k], a1[k]);
          fma_(part, a3[k], a2[k]);
        }
        zz = warp_sum(zz);
        part = warp_sum(part);
        T wj_raw = warp_sum(wj);                  // (A x)(j)
        __syncthreads();                          // (#1) workers' shares of x~^H A x~ are in dscal
        double vv = 0.0;
        if (!mg) {
#pragma unroll
          for (int w = 0; w < NW; ++w) vv += sm.dscal[w];
        } else {
          // row j's partial sums and the v^H A v shares of all ranks (this rank's own included: every rank adds the same P
          // values in the same order, so the replicated panel stays bitwise identical across the ranks)
          T wq = zero_<T>();
          double vq = 0.0;
          if (lane < p.P) {
            const T* slot = ex_slot(p, p.rank, lane, parp);
            T vqt;
            unsigned long long spins = 0;
            do {
              wq = ld_poll(slot + p.wstride - 2);
              vqt = ld_poll(slot + p.wstride - 1);
              if (++spins > (1ull << 24)) { atomicExch(p.status, 78); __trap(); }
            } while (unset_(wq) || unset_(vqt));
            vq = real_(vqt);
          }
          wj_raw = zero_<T>();
          for (int q = 0; q < p.P; ++q) { wj_raw = add_(wj_raw, shfl_(wq, q)); vv += __shfl_sync(0xffffffffu, vq, q); }
        }
        wj = mul_(scale_p, sub_(wj_raw, from_real<T>(beta_p * ajj_p)));    // (A v)(j) = scale * ((A x)(j) - beta A(j,j))
        vv += beta_p * (beta_p * ajj_p - 2.0 * real_(wj_raw));     // x~^H A x~ from x^H A x
        vv *= abs2_(scale_p);                     // v^H A v = |scale|^2 x~^H A x~
        // rho = v^H A v - 2 Re(z1^H z2), alpha' = -1/2 |tau|^2 rho, W(j, c+1) = tau (w_j - part) + alpha' (v(j) = 1)
        const double rho = vv - 2.0 * zz;
        const double alpha_p = -0.5 * abs2_(tau_p) * rho;
        if (lane == 0) {
          S.s_tau = tau_p; S.s_scale = scale_p; S.s_alpha = alpha_p; S.s_beta = beta_p;
          S.s_wj = add_(mul_(tau_p, sub_(wj, part)), from_real<T>(alpha_p));
        }
      } else if (round == 0) {
        __syncthreads();                          // (#1)
      }
      __syncthreads();                            // (#2)
      if (round == 0 && tmap != nullptr && p.npf > 0 && c >= 0 && !mg && cdn.j > 0 &&
          (int64_t)cdn.Tn * (cdn.Tn + 1) * (int64_t)(TB * TB * sizeof(T) / 2) > (int64_t)(96 << 20)) {
        // HBM idles during this phase, the barrier and the Householder scalars: pull the first tiles this CTA
        // will be handed in the coming product into L2 (they were the first ones used for the previous column,
        // i.e. the ones most certainly evicted since)
        constexpr int NBOX = RingCfg<T>::NBOX, DPE = RingCfg<T>::DPE;
        UnitMap um;
        um.Tn = cdn.Tn; um.C = cdn.C; um.rcpC = cdn.rcpC; um.rank = 0; um.P = 1; um.TnO = cdn.Tn; um.KB = cdn.KB;
        um.NF = cdn.NF; um.total = cdn.total; um.bstart = nullptr;
        int left = p.npf;
        for (int unit = cta; unit < um.total && left > 0; unit += G) {
          int J, I0, I1; bool hd;
          um.decode(unit, J, I0, I1, hd);
          const int ntile = (I1 - I0) + (hd ? 1 : 0);
          for (int t = 0; t < ntile && left > 0; ++t, --left) {
            const int I = (hd && t == ntile - 1) ? J : I0 + t;
            if (lane < NBOX) tma_prefetch_2d(tmap, (I * TB) * DPE + lane * 16, J * TB);
          }
        }
      }
      continue;
    }
    // ===================== worker warps =====================
    if (rv && sub == 0) {
      if (have_prev) xb = ldcg_(p.xbuf + r);
      if (c >= 0 || have_prev) acol = ldcg_(p.A + r + (int64_t)j * p.lda);      // (also the correction column of the product)
    }
    const T scale_w = have_prev ? ldcg_(p.scale_slot) : zero_<T>();
    if (have_prev) {
      if (round == 0) {
        // this thread's share of the per-unit v^H A v slots (multi-GPU: of this rank's units)
        const int total = mg ? units_prev : cd.total;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (tid < total) v0 = __ldcg(p.vavunit + tid);
        if (tid + NT < total) v1 = __ldcg(p.vavunit + tid + NT);
        if (tid + 2 * NT < total) v2 = __ldcg(p.vavunit + tid + 2 * NT);
        for (int t = tid + 3 * NT; t < total; t += NT) v0 += __ldcg(p.vavunit + t);
        const double vs = warp_sum((v0 + v1) + v2);
        if (lane == 0) sm.dscal[warp] = vs;
      }
      if (rv) {
        // partial-sum slots of row r: direct J = I+1 .. Tn-1 (multi-GPU: the owned ones, J = rank mod P), then bands
        // k = 0 .. I/Cp (multi-GPU: only when this rank owns tile column I); this warp takes every WPG-th
        const int I = r >> 6;
        int J0 = I + 1;
        if (mg) J0 += ((rk - J0) % Pn + Pn) % Pn;
        const int nd = J0 < Tn ? (Tn - J0 + Pn - 1) / Pn : 0;
        const int nt = (!mg || I % Pn == rk) ? ((I * cd.rcpC) >> 16) + 1 : 0;
        {
          const int64_t step = ((int64_t)p.ldp * Pn) << wpgs;
          const T* ptr = p.Pd + (int64_t)(J0 + sub * Pn) * p.ldp + r;
          int cnt = nd > sub ? (nd - sub + WPG - 1) >> wpgs : 0;
          for (; cnt > 0; cnt -= 8) {
            T v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = cnt > k ? ldcg_(ptr + k * step) : zero_<T>();
            ptr += 8 * step;
            wr = add_(wr, add_(add_(add_(v[0], v[1]), add_(v[2], v[3])), add_(add_(v[4], v[5]), add_(v[6], v[7]))));
          }
        }
        {
          const int64_t step = (int64_t)p.ldp << wpgs;
          const T* ptr = p.Pt + (int64_t)sub * p.ldp + r;
          int cnt = nt > sub ? (nt - sub + WPG - 1) >> wpgs : 0;
          for (; cnt > 0; cnt -= 4) {
            const T v0 = ldcg_(ptr);
            const T v1 = cnt > 1 ? ldcg_(ptr + step) : zero_<T>();
            const T v2 = cnt > 2 ? ldcg_(ptr + 2 * step) : zero_<T>();
            const T v3 = cnt > 3 ? ldcg_(ptr + 3 * step) : zero_<T>();
            ptr += 4 * step;
            wr = add_(wr, add_(add_(v0, v1), add_(v2, v3)));
          }
        }
      }
    }
    // the first four (V, W) pairs of this warp's panel columns are in flight across the barrier
    if (rv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = cprev + 1 + sub + (k << wpgs);
        if (cc < nbp) {
          pv[k] = ldcg_(p.A + r + (int64_t)(p.i0 + cc) * p.lda);
          pw[k] = ldcg_(p.W + r + (int64_t)cc * p.ldw);
        }
      }
    }
    T wloc = zero_<T>();                          // multi-GPU: this rank's complete partial sum of row r (sub == 0 warps)
    if (mg && have_prev) {
      // combine the WPG slices of a row group, then deliver the 32 sums to every other rank (plain stores over NVLink into
      // the peers' exchange buffers); CTA 0 / warp 0 also delivers this rank's v^H A v
      if (round == 0 && cta == 0 && warp == 0 && lane < p.P) {
        // the scalar words of the other parity were read by all CTAs two grid barriers ago: mark them unset again
        T* slot = ex_slot(p, p.rank, lane, parp ^ 1u);
        slot[p.wstride - 2] = unset_value<T>(); slot[p.wstride - 1] = unset_value<T>();
      }
      S.ared[(warp * 32 + lane) * 2] = wr;
      tstamp(p, c, 8);
      consumer_barrier();
      if (sub == 0) {
        for (int w = 0; w < WPG; ++w) wloc = add_(wloc, S.ared[(((grp << wpgs) + w) * 32 + lane) * 2]);
        if (rv) {
          for (int q = 0; q < p.P; ++q) if (q != p.rank) ex_slot(p, q, p.rank, parp)[r] = wloc;
        }
        if (round == 0 && cta == 0 && warp == 0 && lane < p.P) {
          double vloc = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) vloc += sm.dscal[w];
          ex_slot(p, lane, p.rank, parp)[p.wstride - 1] = from_real<T>(vloc);
        }
        tstamp(p, c, 6);
      }
      if (round > 0) consumer_barrier();          // ared is written again below (round 0: sync #1 separates the two uses)
      wr = zero_<T>();
    }
    tstamp(p, c, 5);
    if (round == 0) __syncthreads();              // (#1) z1, z2, rowV, rowW are in shared memory
    tstamp(p, c, 7);
    T t1 = zero_<T>(), t2 = zero_<T>();
    if (rv) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = cprev + 1 + sub + (k << wpgs);
        if (cc < nbp) {
          if (have_prev) { fma_(t1, pw[k], S.z1[cc]); fma_(t1, pv[k], S.z2[cc]); }
          if (c >= 0) { fma_(t2, pv[k], conj_(S.rowW[cc])); fma_(t2, pw[k], conj_(S.rowV[cc])); }
        }
      }
#pragma unroll 2
      for (int cc = cprev + 1 + sub + (4 << wpgs); cc < nbp; cc += WPG) {
        const T vv = ldcg_(p.A + r + (int64_t)(p.i0 + cc) * p.lda);
        const T ww = ldcg_(p.W + r + (int64_t)cc * p.ldw);
        if (have_prev) { fma_(t1, ww, S.z1[cc]); fma_(t1, vv, S.z2[cc]); }
        if (c >= 0) { fma_(t2, vv, conj_(S.rowW[cc])); fma_(t2, ww, conj_(S.rowV[cc])); }
      }
    }
    {
      T* ar = S.ared + (warp * 32 + lane) * 2;
      ar[0] = sub_(mul_(scale_w, wr), t1); ar[1] = t2;
    }
    tstamp(p, c, 9);
    __syncthreads();                              // (#2) also: the scalar warp's results
    tstamp(p, c, 10);
    tstamp(p, c, 11);
    if (rv && sub == 0) {
      T u = zero_<T>(), t2s = zero_<T>();
      for (int w = 0; w < WPG; ++w) {
        const T* a2 = S.ared + (((grp << wpgs) + w) * 32 + lane) * 2;
        u = add_(u, a2[0]); t2s = add_(t2s, a2[1]);
      }
      if (mg && have_prev) {
        // the other ranks' partial sums of this row, added in rank order (this rank's own share comes from registers);
        // every word is polled until the sender's store has landed, then marked unset for the column after the next
        T part[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) part[q] = (q < p.P && q != p.rank) ? ld_poll(ex_slot(p, p.rank, q, parp) + r) : zero_<T>();
        unsigned long long spins = 0;
        for (bool again = true; again;) {
          again = false;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < p.P && q != p.rank && unset_(part[q])) { part[q] = ld_poll(ex_slot(p, p.rank, q, parp) + r); again = true; }
          if (++spins > (1ull << 24)) { atomicExch(p.status, 79); __trap(); }
        }
        T wtot = zero_<T>();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < p.P && q != p.rank) ex_slot(p, p.rank, q, parp)[r] = unset_value<T>();
          if (q < p.P) wtot = add_(wtot, q == p.rank ? wloc : part[q]);
        }
        u = add_(u, mul_(S.s_scale, wtot));
      }
      if (have_prev) {
        const T tau_p = S.s_tau, scale_p = S.s_scale, wjfin = S.s_wj;
        const double alpha_p = S.s_alpha;
        // the product ran on the raw column: (A x~)(r) = (A x)(r) - beta A(r, j)   (rows r < j; row j comes from the scalar warp)
        if (r < j) u = sub_(u, mul_(scale_p, scale_(acol, S.s_beta)));
        // the reflector generated in the previous phase B: v(r) = scale * x(r), v(j) = 1
        const T vnew = (r == j) ? from_real<T>(1.0) : mul_(scale_p, xb);
        p.A[r + (int64_t)(j + 1) * p.lda] = vnew;
        const T wnew = (r == j) ? wjfin : add_(mul_(tau_p, u), scale_(vnew, alpha_p));
        p.W[r + (int64_t)cprev * p.ldw] = wnew;
        if (c >= 0) { fma_(t2s, vnew, conj_(wjfin)); t2s = add_(t2s, wnew); }    // V(j, c+1) = 1
      }
      if (c >= 0) {
        T a = sub_(acol, t2s);
        if (r == j) {
          a = from_real<T>(real_(a));
          p.d[j] = real_(a);
          p.A[j + (int64_t)j * p.lda] = a;
          p.xbuf[j] = zero_<T>();       // x is zero from row j on (the tile engine reads whole 64-row slices)
        } else {
          p.xbuf[r] = a;
          if (r < j - 1) nrm += abs2_(a);
          if (r == j - 1) *p.alpha_slot = a;
        }
      }
    }
  }
  // per-warp partial norms; the caller (grid barrier or the stand-alone kernel) adds them up after a CTA barrier
  nrm = warp_sum(nrm);
  if (lane == 0) sm.nred[warp] = nrm;
}

// sum of the per-warp partial norms of phase A -> npart[cta]; one thread, after a CTA barrier
template <typename T>
__device__ __forceinline__ void store_npart(const TrdP<T>& p, const PanelSmem<T>& sm) {
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NWT; ++w) s += sm.nred[w];
  p.npart[blockIdx.x] = s;
}

// Phase B (tile phase).  Warp 0 derives the Householder scalars (fixed summation order => identical in every
// CTA) while the producer warp already fetches tiles (they do not depend on v; x travels unscaled and is scaled
// when it is used); the consumer warps of the CTAs at the far end of the grid then take the z-dot units
// (z1 = V^H v, z2 = W^H v, one (matrix, column) pair per CTA, complete dot product) before joining the tile queue,
// which absorbs the difference.  Returns the number of tile units of this product.
template <typename T, bool MG>
__device__ int phase_b(const TrdP<T>& p, int c, PanelSmem<T>& sm, T* ring, RingState& rs, const CUtensorMap* tmap,
                       const ColDesc& cd, ColDesc* next_cd) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int j = p.i0 + c;            // order of the product; reflector index j-1
  if (j <= 0) return 0;
  const int Pn = MG ? p.P : 1, rk = MG ? p.rank : 0;
  UnitMap um;
  if (Pn > 1) {
    um = engine_prepare<T>(j, cd.C, rk, Pn, sm.u.e);      // tabulates this rank's units (CTA barriers inside)
  } else {
    um.Tn = cd.Tn; um.C = cd.C; um.rank = 0; um.P = 1; um.TnO = cd.Tn; um.KB = cd.KB; um.NF = cd.NF; um.total = cd.total;
    um.bstart = nullptr;
  }
  um.rcpC = cd.rcpC;
  // cache hints only while the triangle of this product exceeds what L2 can hold anyway
  um.keepI = ((int64_t)cd.Tn * cd.Tn * (int64_t)(TB * TB * sizeof(T) / 2) > ((int64_t)96 << 20)) ? p.keepI : 0;
  // The product runs on the RAW updated column x (x(j-1) = alpha), not on the Householder vector: with
  //   x~ = x - beta e_(j-1)   (so that v = scale * x~, scale = 1 / (alpha - beta)),
  // A x~ = A x - beta A(:, j-1),  x~^H A x~ = x^H A x - 2 beta Re (A x)(j-1) + beta^2 A(j-1, j-1)  and
  // V^H x~ = V^H x - beta conj(V(j-1, :)).  The consumers of the partial sums (the next phase A) apply these rank-one
  // corrections and the scale once per row / scalar -- column j-1 of A is the very column they load anyway.  Nothing in
  // the tile phase therefore waits for the Householder scalars: beta, tau, scale are formed by the producer warp of CTA 0
  // once its first ring-full of tiles is in flight (fixed summation order; every CTA reads them back after the barrier).
  auto scalars_hook = [&]() {
    if (cta != 0) return;
    double xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { const int g = lane + 32 * k; xs[k] = g < G ? __ldcg(p.npart + g) : 0.0; }
    const T alpha = ldcg_(p.alpha_slot);
    const double ajj = real_(ldcg_(p.A + (j - 1) + (int64_t)(j - 1) * p.lda));     // stored diagonal the tiles see
    double x2 = ((xs[0] + xs[1]) + (xs[2] + xs[3])) + xs[4];
    for (int g = lane + 160; g < G; g += 32) x2 += __ldcg(p.npart + g);
    x2 = warp_sum(x2);
    double beta; T tau, scale;
    larfg_scalars(alpha, x2, beta, tau, scale);
    if (lane == 0) {
      p.e[j - 1] = beta; p.tau[j - 1] = tau; *p.scale_slot = scale;
      p.beta_slot[0] = beta; p.beta_slot[1] = ajj;
    }
  };
  tstamp(p, c, 12);
  auto xfix = [](int, T raw) -> T { return raw; };      // xbuf holds x with zeros from row j on (phase A)
  // -- z1 = V^H v, z2 = W^H v: pair q = (which, cc) is done completely by the consumer warps of CTA G-1-q
  const int nf = p.nbp - 1 - c;      // finished columns cc in (c, nbp)
  if (nf > 0 && warp < NW) {
    T* zr = &sm.u.e.ydiag[0][0];     // per-warp partial dots (the engine's scratch is not in use yet)
    for (int q = G - 1 - cta; q < 2 * nf; q += G) {
      const int which = q >= nf ? 1 : 0, cc = c + 1 + (q - which * nf);
      const T* col = which ? (p.W + (int64_t)cc * p.ldw) : (p.A + (int64_t)(p.i0 + cc) * p.lda);
      T s0 = zero_<T>(), s1 = zero_<T>();
      for (int r0 = tid; r0 < j; r0 += 4 * NT) {
        T cv[4], xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = r0 + k * NT;
          cv[k] = r < j ? ldcg_(col + r) : zero_<T>();
          xv[k] = r < j ? ldcg_(p.xbuf + r) : zero_<T>();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) fmac_((k & 1) ? s1 : s0, cv[k], xv[k]);
      }
      s0 = warp_sum(add_(s0, s1));
      if (lane == 0) zr[warp] = s0;
      consumer_barrier();
      if (tid == 0) {
        T s = zero_<T>();
#pragma unroll
        for (int w = 0; w < NW; ++w) s = add_(s, zr[w]);
        p.zfin[which * NBMAX + cc] = s;
      }
      consumer_barrier();
    }
  }
  tstamp(p, c, 13);
  // -- the tile engine: w_raw partials and v^H A v
  engine_run<T>(um, p.A, p.lda, j, p.xbuf, xfix, p.Pd, p.Pt, p.ldp, p.vavunit, p.qctr + c, cta, G, p.use_tma != 0, ring,
                sm.full, sm.empty, sm.meta, rs, sm.u.e, tmap, next_cd, p.i0 + c - 1, Pn, p.upc, p.ctab, scalars_hook);
  return um.total;
}

template <typename T, bool MG>
__global__ void __launch_bounds__(NTT, 1) panel_coop_kernel(const __grid_constant__ CUtensorMap tmap, TrdP<T> p) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ PanelSmem<T> sm;
  T* ring = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
  RingState rs;
  const int Pn = MG ? p.P : 1;
  if (threadIdx.x == NT) compute_desc(sm.cd[(p.nbp - 1) & 1], p.i0 + p.nbp - 1, gridDim.x, Pn, p.upc, p.ctab);
  ring_init(sm.full, sm.empty, RingCfg<T>::STAGES, rs);      // (CTA barrier inside)
  unsigned target = 0;
  int units_prev = 0;                  // tile units of the previous product (multi-GPU: of this rank)
  auto stamp = [&](int c, int k) { tstamp(p, c, k); };
  for (int c = p.nbp - 1; c >= -1; --c) {
    if (c >= 0) stamp(c, 0);
    phase_a<T, MG>(p, c, sm, sm.cd[(c + 1) & 1], sm.cd[c & 1], &tmap, units_prev);
    if (c < 0) break;
    stamp(c, 1);
    target += gridDim.x;
    grid_barrier(p.barrier, target, p.status, false, [&]() { store_npart(p, sm); });
    stamp(c, 2);
    const bool spread = p.trace != nullptr && p.i0 + c == p.etrace_j && threadIdx.x == 0;   // per-CTA begin/end of phase B
    if (spread) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.trace[p.etrace_off + 512 + blockIdx.x] = t; }
    units_prev = phase_b<T, MG>(p, c, sm, ring, rs, &tmap, sm.cd[c & 1], &sm.cd[(c + 1) & 1]);
    if (spread) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.trace[p.etrace_off + 768 + blockIdx.x] = t; }
    stamp(c, 3);
    target += gridDim.x;
    // (multi-GPU: the exchange of the partial sums is part of the next phase A -- per-group flags, no third barrier)
    grid_barrier(p.barrier, target, p.status, false, []() {});
    stamp(c, 4);
  }
}
template <typename T>
__global__ void __launch_bounds__(NTT, 1) phase_a_kernel(TrdP<T> p, int c) {
  __shared__ PanelSmem<T> sm;
  if (threadIdx.x == 0) compute_desc(sm.cd[0], p.i0 + c + 1, gridDim.x, 1, p.upc, p.ctab);
  __syncthreads();
  phase_a<T, false>(p, c, sm, sm.cd[0], sm.cd[0], nullptr, 0);
  __syncthreads();
  if (threadIdx.x == 0 && c >= 0) store_npart(p, sm);
}
template <typename T>
__global__ void __launch_bounds__(NTT, 1) phase_b_kernel(const __grid_constant__ CUtensorMap tmap, TrdP<T> p, int c) {
  extern __shared__ __align__(1024) unsigned char dyn_smem[];
  __shared__ PanelSmem<T> sm;
  T* ring = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
  RingState rs;
  if (threadIdx.x == NT) compute_desc(sm.cd[0], p.i0 + c, gridDim.x, 1, p.upc, p.ctab);
  ring_init(sm.full, sm.empty, RingCfg<T>::STAGES, rs);
  phase_b<T, false>(p, c, sm, ring, rs, &tmap, sm.cd[0], &sm.cd[1]);
}

template <typename T>
int panel_grid(int& grid, size_t dyn_smem) {
  static OncePerDevice once;
  if (once.need()) {
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(panel_coop_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes<T>()));
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(panel_coop_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes<T>()));
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(phase_b_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes<T>()));
    EIGB_CUDA_CHECK(cudaFuncSetAttribute(hemv_tiles_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes<T>()));
    once.done();
  }
  int per_sm = 0;
  EIGB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, panel_coop_kernel<T, true>, NTT, dyn_smem));
  if (per_sm < 1) { set_last_error("panel kernel does not fit on an SM"); return -1; }
  grid = ctx().num_sms;      // one persistent CTA per SM
  return 0;
}

// scratch layout helper shared by hemv() and hetrd()
template <typename T>
size_t partial_elems(int n, int64_t& ldp) {
  ldp = ((int64_t)n + 63) & ~int64_t(63);
  int Tn = (n + TB - 1) / TB;
  return (size_t)ldp * (size_t)(Tn > 0 ? Tn : 1);
}

}  // namespace

template <typename T>
int hemv_upper(cudaStream_t s, int n, const T* A, int64_t lda, const T* x, T* y) {
  if (n <= 0) return 0;
  int64_t ldp;
  size_t pe = partial_elems<T>(n, ldp);
  const int npad = ((n + 63) & ~63) + 64;
  const size_t Tn = (size_t)(n + TB - 1) / TB;
  const size_t nunits = Tn * (Tn + 1) / 2 + 64;
  void* scr = ctx_scratch((2 * pe + npad) * sizeof(T) + nunits * sizeof(double) + 4096);
  if (!scr) return -1;
  Arena ar(scr, ctx().scratch_bytes);
  T* Pd = ar.take<T>(pe);
  T* Pt = ar.take<T>(pe);
  T* xpad = ar.take<T>(npad);
  double* vavunit = ar.take<double>(nunits);
  unsigned* qctr = ar.take<unsigned>(64);
  if (!qctr) { set_last_error("hemv: scratch arena too small"); return -1; }
  const int vec_ok = is_cplx<T>::value ? 1 : ((((uintptr_t)A & 15) == 0 && (lda & 1) == 0) ? 1 : 0);
  int tma = (opts().symv_tma != 0 && vec_ok && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (tma && make_tmap_f64_box16x64(&tmap, A, (uint64_t)n * RingCfg<T>::DPE, (uint64_t)n, (uint64_t)lda * sizeof(T)) != 0)
    tma = 0;
  int grid = 0;
  if (panel_grid<T>(grid, tma ? ring_bytes<T>() : 0) != 0) return -1;
  const int C = strip_len(n, grid, 1, opts().trd_upc);
  EIGB_CUDA_CHECK(cudaMemsetAsync(qctr, 0, sizeof(unsigned), s));
  pad_copy_kernel<T><<<cdiv(npad, 256), 256, 0, s>>>(x, n, xpad, npad);
  hemv_tiles_kernel<T><<<grid, NTT, tma ? ring_bytes<T>() : 0, s>>>(tmap, A, lda, n, xpad, Pd, Pt, ldp, C, tma, vavunit, qctr);
  EIGB_LAUNCH_CHECK();
  hemv_reduce_kernel<T><<<cdiv(n, 256), 256, 0, s>>>(Pd, Pt, ldp, n, C, y);
  EIGB_LAUNCH_CHECK();
  return 0;
}

// Tridiagonalization driver.  A: n x n (upper read/written), outputs d(n), e(n-1), tau(n-1) on device.
// On exit the reflectors v_j (j = 1..n-1, 1-based) are in A(1:j-1, j+1) with the unit element stored
// explicitly, exactly as the reference leaves them (zhetrd_gpu.F90:92).
template <typename T>
int hetrd_upper(cudaStream_t s, int n, T* A, int64_t lda, double* d, double* e, T* tau, bool sync_status) {
  if (n <= 0) return 0;
  Context& c = ctx();
  const int nb = opts().trd_nb < NBMAX ? opts().trd_nb : NBMAX;
  const int vec_ok = is_cplx<T>::value ? 1 : ((((uintptr_t)A & 15) == 0 && (lda & 1) == 0) ? 1 : 0);
  int use_tma = (opts().symv_tma != 0 && vec_ok && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (use_tma && make_tmap_f64_box16x64(&tmap, A, (uint64_t)n * RingCfg<T>::DPE, (uint64_t)n, (uint64_t)lda * sizeof(T)) != 0)
    use_tma = 0;
  const size_t dyn = use_tma ? (size_t)ring_bytes<T>() : 0;
  int grid = 0;
  if (panel_grid<T>(grid, dyn) != 0) return -1;
  int64_t ldp;
  size_t pe = partial_elems<T>(n, ldp);
  const size_t Tnn = (size_t)(n + TB - 1) / TB;
  const size_t nunits = Tnn * (Tnn + 1) / 2 + 64;
  size_t bytes = (2 * pe + (size_t)n * nb + (size_t)n + 256 + (size_t)2 * NBMAX + 64) * sizeof(T) +
                 nunits * sizeof(double) + 4096 +
                 ((size_t)(n / TB + 2) * (size_t)(n / TB + 2) / 2 + 2 * (size_t)(n / TB + 2) + 16) * sizeof(GemmParams<T>) +
                 (size_t)(2 * grid + 64) * sizeof(double) + 4096 + 16 * 256 + 2 * (Tnn + 2 + 256);
  void* scr = ctx_scratch(bytes);
  if (!scr) return -1;
  Arena ar(scr, c.scratch_bytes);
  TrdP<T> p{};
  p.A = A; p.lda = lda; p.d = d; p.e = e; p.tau = tau;
  p.Pd = ar.take<T>(pe); p.Pt = ar.take<T>(pe); p.ldp = ldp;
  p.W = ar.take<T>((size_t)n * nb); p.ldw = n;
  p.xbuf = ar.take<T>((size_t)n + 128);
  p.zfin = ar.take<T>((size_t)2 * NBMAX);
  p.alpha_slot = ar.take<T>(16);
  p.scale_slot = ar.take<T>(16);
  p.beta_slot = ar.take<double>(16);
  p.npart = ar.take<double>(grid);
  p.vavunit = ar.take<double>(nunits);
  p.barrier = ar.take<unsigned>(64 + NBMAX);      // barrier word + the per-column tile queue heads (one memset)
  p.status = c.d_info;
  if (!p.barrier) { set_last_error("hetrd: scratch arena too small"); return -1; }
  p.qctr = p.barrier + 64;
  p.use_tma = use_tma;
  p.upc = opts().trd_upc;
  {
    // tiles of the top keepI tile rows: keepI * Tn * tile_bytes ~ trd_l2keep_mb; only when the triangle exceeds L2
    const double tile_mb = (double)(TB * TB * sizeof(T)) / (1 << 20);
    const double tri_mb = 0.5 * Tnn * Tnn * tile_mb;
    p.keepI = 0;
    if (use_tma && opts().trd_l2keep_mb > 0 && tri_mb > 100.0) {
      p.keepI = (int)(opts().trd_l2keep_mb / (tile_mb * (double)Tnn));
      if (p.keepI < 1) p.keepI = 1;
    }
  }
  p.npf = use_tma ? (opts().trd_prefetch >= 0 ? opts().trd_prefetch : (is_cplx<T>::value ? 4 : 8)) : 0;
  p.trace = nullptr;
  MgConfig& M = mg();
  p.rank = 0; p.P = 1;
  GemmParams<T>* GP = nullptr;
  const bool dist = M.P > 1 && M.active;
  // strip length per number of tile rows: replay of the unit queue on the host (cached per grid / rank count / type)
  const unsigned char* ctab_1 = nullptr; const unsigned char* ctab_p = nullptr;
  if (opts().trd_ctab) {
    const int cmax = (p.upc >> 8) ? (p.upc >> 8) : 8;
    const double ov = is_cplx<T>::value ? 0.5 : 1.0;       // per-unit cost in tile times (real tiles are half the size)
    const int Pm = (M.P > 1 && M.active) ? M.P : 1;
    for (int pass = 0; pass < (Pm > 1 ? 2 : 1); ++pass) {
      const int Pt = pass == 0 ? 1 : Pm;
      static std::vector<unsigned char> cache[2][9];
      static int cache_key[2][9][2];
      std::vector<unsigned char>& tab = cache[is_cplx<T>::value ? 1 : 0][Pt];
      int* key = cache_key[is_cplx<T>::value ? 1 : 0][Pt];
      if ((int)tab.size() < (int)Tnn + 2 || key[0] != grid || key[1] != cmax) {
        tab.assign(Tnn + 2, 1);
        tile::build_strip_table(grid, Pt, (int)Tnn + 1, cmax, ov, tab.data());
        key[0] = grid; key[1] = cmax;
      }
      unsigned char* dt = ar.take<unsigned char>(Tnn + 2);
      if (!dt) { set_last_error("hetrd: scratch arena too small"); return -1; }
      EIGB_CUDA_CHECK(cudaMemcpyAsync(dt, tab.data(), Tnn + 2, cudaMemcpyHostToDevice, s));
      (pass == 0 ? ctab_1 : ctab_p) = dt;
    }
  }
  p.ctab = ctab_1;
  if (dist) {
    if (!opts().trd_coop) { set_last_error("hetrd: multi-GPU needs the cooperative panel kernel"); return -1; }
    if (nb != TB) { set_last_error("hetrd: multi-GPU needs trd_nb == 64 (panels aligned to the tile columns)"); return -1; }
    p.rank = M.rank; p.P = M.P;
    p.wstride = M.wbuf_bytes / ((int64_t)M.P * 2 * (int64_t)sizeof(T));
    if (p.wstride < (int64_t)n + 2) { set_last_error("hetrd: multi-GPU exchange buffer too small"); return -1; }
    for (int q = 0; q < M.P; ++q) { p.peer_w[q] = (T*)M.wbuf[q]; p.peer_flag[q] = M.flags[q]; }
    p.fstride = M.flag_stride;
    (void)p.fstride;      // (the arrival flags are no longer used: the exchanged words validate themselves)
    // parameter blocks of the per-tile-column rank-2k updates of ALL panels: built once, uploaded once (no host
    // synchronisation inside the panel loop)
    const size_t ntc = (size_t)(n / TB + 2);
    GP = ar.take<GemmParams<T>>(ntc * ntc / 2 / M.P + 2 * ntc + 8);
    if (!GP) { set_last_error("hetrd: scratch arena too small (multi-GPU)"); return -1; }
  }
  if (opts().trd_trace) {
    if (cudaMalloc(&p.trace, ((size_t)n * TRSLOTS + 1024) * sizeof(unsigned long long)) != cudaSuccess) p.trace = nullptr;
    else cudaMemsetAsync(p.trace, 0, ((size_t)n * TRSLOTS + 1024) * sizeof(unsigned long long), s);
    p.etrace_j = opts().trd_trace > 1 ? opts().trd_trace : -1;
    p.etrace_off = (int64_t)n * TRSLOTS;
  }
  EIGB_CUDA_CHECK(cudaMemsetAsync(p.status, 0, sizeof(int), s));
  EIGB_CUDA_CHECK(cudaMemsetAsync(p.xbuf, 0, ((size_t)n + 128) * sizeof(T), s));     // rows >= order must read as zero
  const bool coop = opts().trd_coop != 0;
  std::vector<GemmParams<T>> hp_all;          // multi-GPU: rank-2k parameter blocks, panel after panel
  std::vector<int> hp_off;
  if (dist) {
    int hi2 = n;
    while (hi2 > 0) {
      int nbp2 = hi2 < nb ? hi2 : nb;
      if (hi2 > nb && (hi2 % nb) != 0) nbp2 = hi2 % nb;
      const int i02 = hi2 - nbp2;
      hp_off.push_back((int)hp_all.size());
      const T* V = A + (int64_t)i02 * lda;
      for (int J = M.rank; J < i02 / TB; J += M.P) {
        GemmParams<T> q;
        memset(&q, 0, sizeof(q));
        q.M = (J + 1) * TB; q.N = TB; q.nseg = 2;
        q.A[0] = V; q.lda[0] = lda; q.B[0] = p.W + J * TB; q.ldb[0] = p.ldw; q.K[0] = nbp2;
        q.A[1] = p.W; q.lda[1] = p.ldw; q.B[1] = V + J * TB; q.ldb[1] = lda; q.K[1] = nbp2;
        q.sa[0] = q.sa[1] = 1.0; q.sb[0] = q.sb[1] = -1.0;
        q.C = A + (int64_t)J * TB * lda; q.ldc = lda;
        q.alpha = -1.0; q.beta = 1.0; q.mode = 1; q.real_diag = 1; q.colmap = nullptr; q.diag_off = J * TB;
        hp_all.push_back(q);
      }
      hi2 = i02;
    }
    hp_off.push_back((int)hp_all.size());
    if (!hp_all.empty()) {
      EIGB_CUDA_CHECK(cudaMemcpyAsync(GP, hp_all.data(), sizeof(GemmParams<T>) * hp_all.size(), cudaMemcpyHostToDevice, s));
      EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
  }
  int panel_idx = 0;

  int hi = n;                         // columns [0, hi) still to reduce
  while (hi > 0) {
    int nbp = hi < nb ? hi : nb;
    // the first (rightmost) panel absorbs the remainder so that the others are aligned to nb
    if (hi > nb && (hi % nb) != 0) nbp = hi % nb;
    p.i0 = hi - nbp; p.nbp = nbp;
    if (dist && p.P > 1 && hi <= (opts().mg_switch_n >= 0 ? opts().mg_switch_n : (M.P <= 2 ? 3072 : 2048))) {
      // small trailing matrix: the per-column exchange costs more than the tiles it saves.  Make the leading hi
      // columns current everywhere (owner = -1: "gather all tile columns") and finish replicated (deterministic).
      if (M.hook) M.hook(hi, 0, -1);
      else if (mg_bcast_columns(s, A, lda, hi, 0, -1, (int)sizeof(T)) != 0) return -1;
      p.P = 1; p.rank = 0;
    }
    p.ctab = p.P > 1 ? ctab_p : ctab_1;
    if (p.P > 1) {
      p.seq_base = M.seq;
      M.seq += (unsigned long long)nbp;
      // the panel's columns are current only on the rank that owns this tile column: broadcast them.  So is column
      // i0-1 (owned by the previous tile column's rank): the panel's last product (order i0) runs on the raw column, and
      // its consumers correct the partial sums with the STORED column i0-1 and diagonal element (phase A of c = -1)
      const int own = (p.i0 / TB) % M.P, own1 = p.i0 > 0 ? ((p.i0 - 1) / TB) % M.P : 0;
      if (M.hook) {
        M.hook(p.i0, nbp, own);
        if (p.i0 > 0) M.hook(p.i0 - 1, 1, own1);
      } else {
        if (mg_group(true) != 0) return -1;
        int rcb = mg_bcast_columns(s, A, lda, p.i0, nbp, own, (int)sizeof(T));
        if (rcb == 0 && p.i0 > 0) rcb = mg_bcast_columns(s, A, lda, p.i0 - 1, 1, own1, (int)sizeof(T));
        if (mg_group(false) != 0 || rcb != 0) return -1;
      }
    }
    prof_begin(PROF_PANEL, s);
    EIGB_CUDA_CHECK(cudaMemsetAsync(p.barrier, 0, (64 + NBMAX) * sizeof(unsigned), s));
    if (coop) {
      void* args[] = {&tmap, &p};
      void* kfn = (p.P > 1) ? (void*)panel_coop_kernel<T, true> : (void*)panel_coop_kernel<T, false>;
      EIGB_CUDA_CHECK(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(NTT), args, dyn, s));
      count_launch(1);
    } else {
      for (int cc = nbp - 1; cc >= -1; --cc) {
        phase_a_kernel<T><<<grid, NTT, 0, s>>>(p, cc);
        if (cc < 0) break;
        phase_b_kernel<T><<<grid, NTT, dyn, s>>>(tmap, p, cc);
      }
      count_launch(2 * nbp + 1);
      EIGB_CUDA_CHECK(cudaGetLastError());
    }
    prof_end(PROF_PANEL, s);
    // trailing update A(0:i0, 0:i0) -= V W^H + W V^H   (zhetrd_gpu.F90:67 / dsytrd_gpu.F90:66)
    if (p.i0 > 0 && p.P == 1) {
      ProfScope ps(PROF_HER2K, s);
      if (her2k_upper<T>(s, 'N', p.i0, nbp, -1.0, A + (int64_t)p.i0 * lda, lda, p.W, p.ldw, 1.0, A, lda) != 0)
        return -1;
    } else if (p.i0 > 0) {
      // multi-GPU: only the tile columns J = rank (mod P) of the trailing matrix are kept current on this rank
      ProfScope ps(PROF_HER2K, s);
      const int cnt = hp_off[panel_idx + 1] - hp_off[panel_idx];
      if (cnt > 0) {
        GemmParams<T> dummy{};
        if (gemm_launch<T>(s, false, false, dummy, GP + hp_off[panel_idx], cnt, p.i0, TB) != 0) return -1;
      }
    }
    ++panel_idx;
    hi = p.i0;
  }
  int st = 0;
  if (sync_status || p.trace) {
    EIGB_CUDA_CHECK(cudaMemcpyAsync(&st, p.status, sizeof(int), cudaMemcpyDeviceToHost, s));
    EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  if (p.trace) {
    trace_store().resize((size_t)n * TRSLOTS + 1024);
    cudaMemcpy(trace_store().data(), p.trace, ((size_t)n * TRSLOTS + 1024) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
  }
  if (st != 0) { set_last_error("hetrd: device status %d (grid barrier watchdog)", st); return -1; }
  return 0;
}

template int hemv_upper<double>(cudaStream_t, int, const double*, int64_t, const double*, double*);
template int hemv_upper<double2>(cudaStream_t, int, const double2*, int64_t, const double2*, double2*);
template int hetrd_upper<double>(cudaStream_t, int, double*, int64_t, double*, double*, double*, bool);
template int hetrd_upper<double2>(cudaStream_t, int, double2*, int64_t, double*, double*, double2*, bool);

}  // namespace eigb200
