// eigb200 -- shared device/host helpers for the B200 (sm_100a) generalized eigensolver hot path.
// Real FP64 is `double`, complex FP64 is `double2` (x = re, y = im), column-major everywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace eigb200 {

typedef double2 zdouble;

template <typename T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<double2> { static constexpr bool value = true; };

// ---- scalar algebra on T in {double, double2} -------------------------------------------------------
__host__ __device__ __forceinline__ double2 mkz(double re, double im) { double2 r; r.x = re; r.y = im; return r; }
template <typename T> __host__ __device__ __forceinline__ T zero_();
template <> __host__ __device__ __forceinline__ double zero_<double>() { return 0.0; }
template <> __host__ __device__ __forceinline__ double2 zero_<double2>() { return mkz(0.0, 0.0); }
template <typename T> __host__ __device__ __forceinline__ T from_real(double r);
template <> __host__ __device__ __forceinline__ double from_real<double>(double r) { return r; }
template <> __host__ __device__ __forceinline__ double2 from_real<double2>(double r) { return mkz(r, 0.0); }

__host__ __device__ __forceinline__ double conj_(double a) { return a; }
__host__ __device__ __forceinline__ double2 conj_(double2 a) { return mkz(a.x, -a.y); }
__host__ __device__ __forceinline__ double real_(double a) { return a; }
__host__ __device__ __forceinline__ double real_(double2 a) { return a.x; }
__host__ __device__ __forceinline__ double imag_(double) { return 0.0; }
__host__ __device__ __forceinline__ double imag_(double2 a) { return a.y; }
__host__ __device__ __forceinline__ double add_(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ double2 add_(double2 a, double2 b) { return mkz(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double sub_(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ double2 sub_(double2 a, double2 b) { return mkz(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double mul_(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ double2 mul_(double2 a, double2 b) {
  return mkz(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double scale_(double a, double s) { return a * s; }
__host__ __device__ __forceinline__ double2 scale_(double2 a, double s) { return mkz(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ double neg_(double a) { return -a; }
__host__ __device__ __forceinline__ double2 neg_(double2 a) { return mkz(-a.x, -a.y); }
// acc += a * b
__host__ __device__ __forceinline__ void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
__host__ __device__ __forceinline__ void fma_(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__host__ __device__ __forceinline__ void fmac_(double& acc, double a, double b) { acc = fma(a, b, acc); }
__host__ __device__ __forceinline__ void fmac_(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__host__ __device__ __forceinline__ double abs2_(double a) { return a * a; }
__host__ __device__ __forceinline__ double abs2_(double2 a) { return a.x * a.x + a.y * a.y; }

// ---- warp helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double2 warp_sum(double2 v) {
  v.x = warp_sum(v.x); v.y = warp_sum(v.y); return v;
}

// ---- error plumbing ---------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define EIGB_CUDA_CHECK(expr)                                                                 \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::eigb200::set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,          \
                                cudaGetErrorString(_e));                                      \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)
#define EIGB_LAUNCH_CHECK() do { ::eigb200::count_launch(1); EIGB_CUDA_CHECK(cudaGetLastError()); } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- context (replaces the reference's process-global eigsolve_vars, eigsolve_vars.F90:25-61) ----------
struct Context {
  bool initialized = false;
  int device = -1;
  int num_sms = 0;
  cudaStream_t stream = 0;       // all hot-path work is issued here (default: legacy default stream,
                                 // like the reference's cuBLAS-legacy/default-stream ordering)
  cudaStream_t stream2 = nullptr;  // side stream for overlap (created lazily)
  cudaStream_t stream_hi = nullptr;  // highest-priority side stream: latency-bound critical-path work next to a bulk update
  cudaEvent_t ev1 = nullptr, ev2 = nullptr;
  cudaEvent_t a_ready = nullptr;   // one-shot: the generalized driver waits for it before touching A
  void* scratch = nullptr;       // growable device scratch (stedc, panel partials, ...)
  size_t scratch_bytes = 0;
  int* d_info = nullptr;         // device-side status words: [0] hetrd watchdog, [1] potrf pivot, [2] stedc convergence
  int* h_status = nullptr;       // pinned host mirror of d_info (read back once, at the end of a driver call)
  int epoch = 0;                 // bumped whenever the device changes: per-device one-time setup (kernel attributes) is redone
  int verbose = 0;
};
// status word indices in Context::d_info
enum StatusWord { ST_HETRD = 0, ST_POTRF = 1, ST_STEDC = 2, ST_NWORDS = 4 };
// true exactly once per (call site, device epoch): guards cudaFuncSetAttribute-style per-device setup
struct OncePerDevice {
  int seen = -1;
  bool need() { if (seen == ctx_epoch()) return false; return true; }
  void done() { seen = ctx_epoch(); }
  static int ctx_epoch();
};
Context& ctx();
int ctx_init();
// returns a device pointer to at least `bytes` of scratch (grown if needed; contents undefined)
void* ctx_scratch(size_t bytes);

// ---- optional stage profiling (CUDA events on the launching stream) and launch counting ------------------
// Categories: 0 potrf, 1 hegst, 2 hetrd_panel (persistent panel kernel), 3 hetrd_her2k, 4 stedc, 5 ormtr,
// 6 trsm, 7 other.
enum ProfCat { PROF_POTRF = 0, PROF_HEGST, PROF_PANEL, PROF_HER2K, PROF_STEDC, PROF_ORMTR, PROF_TRSM, PROF_OTHER, PROF_NCAT };
void prof_begin(int cat, cudaStream_t s);
void prof_end(int cat, cudaStream_t s);
void count_launch(int n = 1);
struct ProfScope {
  int cat; cudaStream_t s;
  ProfScope(int c, cudaStream_t st) : cat(c), s(st) { prof_begin(cat, s); }
  ~ProfScope() { prof_end(cat, s); }
};

// simple bump allocator over the context scratch
struct Arena {
  char* base; size_t cap; size_t off;
  Arena(void* p, size_t c) : base((char*)p), cap(c), off(0) {}
  template <typename U> U* take(size_t count) {
    size_t bytes = (count * sizeof(U) + 255) & ~size_t(255);
    if (off + bytes > cap) return nullptr;
    U* r = (U*)(base + off); off += bytes; return r;
  }
};

}  // namespace eigb200
