// eigb200 -- process context: device properties, stream, growable scratch (replaces eigsolve_vars.F90:25-61).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace eigb200 {

static char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  if (ctx().verbose) fprintf(stderr, "eigb200 error: %s\n", g_last_error);
}
const char* last_error() { return g_last_error; }

Context& ctx() {
  static Context c;
  return c;
}

int ctx_init() {
  Context& c = ctx();
  int dev = 0;
  EIGB_CUDA_CHECK(cudaGetDevice(&dev));
  if (c.initialized && c.device == dev) return 0;
  if (c.initialized) {
    // device changed: release the old device's resources with that device current, then start over.  The caller's
    // stream belonged to the old device as well: fall back to the legacy default stream until set_stream is called.
    cudaSetDevice(c.device);
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr; c.scratch_bytes = 0;
    if (c.d_info) cudaFree(c.d_info);
    c.d_info = nullptr;
    if (c.stream2) cudaStreamDestroy(c.stream2);
    if (c.stream_hi) cudaStreamDestroy(c.stream_hi);
    c.stream_hi = nullptr;
    if (c.ev1) cudaEventDestroy(c.ev1);
    if (c.ev2) cudaEventDestroy(c.ev2);
    c.stream2 = nullptr; c.ev1 = c.ev2 = nullptr;
    c.stream = 0; c.a_ready = nullptr;
    c.initialized = false;
    cudaSetDevice(dev);
  }
  c.epoch += 1;
  cudaDeviceProp prop;
  EIGB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    set_last_error("eigb200 requires an sm_100a (B200) device, found sm_%d%d", prop.major, prop.minor);
    return -1;
  }
  c.device = dev;
  c.num_sms = prop.multiProcessorCount;
  EIGB_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    EIGB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    EIGB_CUDA_CHECK(cudaStreamCreateWithPriority(&c.stream_hi, cudaStreamNonBlocking, hi));
  }
  EIGB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev1, cudaEventDisableTiming));
  EIGB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev2, cudaEventDisableTiming));
  EIGB_CUDA_CHECK(cudaMalloc(&c.d_info, 64 * sizeof(int)));
  EIGB_CUDA_CHECK(cudaMemset(c.d_info, 0, 64 * sizeof(int)));
  if (!c.h_status) EIGB_CUDA_CHECK(cudaHostAlloc((void**)&c.h_status, 64 * sizeof(int), cudaHostAllocPortable));
  const char* v = getenv("EIGB200_VERBOSE");
  c.verbose = v ? atoi(v) : 0;
  c.initialized = true;
  return 0;
}

int OncePerDevice::ctx_epoch() { return ctx().epoch; }

void* ctx_scratch(size_t bytes) {
  Context& c = ctx();
  if (bytes <= c.scratch_bytes) return c.scratch;
  if (c.scratch) {
    cudaStreamSynchronize(c.stream);
    cudaFree(c.scratch);
    c.scratch = nullptr; c.scratch_bytes = 0;
  }
  size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
  if (cudaMalloc(&c.scratch, want) != cudaSuccess) {
    set_last_error("eigb200: cannot allocate %zu bytes of device scratch", want);
    c.scratch = nullptr;
    return nullptr;
  }
  c.scratch_bytes = want;
  return c.scratch;
}

}  // namespace eigb200

#include "stages.cuh"
#include <nvtx3/nvToolsExt.h>
#include <vector>
namespace eigb200 {

// ---- profiling / launch counting ------------------------------------------------------------------------
static bool g_prof_on = false;
static long long g_launches = 0;
struct ProfPair { cudaEvent_t a, b; int cat; };
static std::vector<ProfPair> g_pairs;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t g_open[PROF_NCAT];
static double g_ms[PROF_NCAT];
static int g_cnt[PROF_NCAT];

void count_launch(int n) { g_launches += n; }
static cudaEvent_t prof_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
// NVTX3 ranges around the stages (header-only: the injection library is loaded by the tools, nothing is linked) --
// the counterpart of the reference's nvtx_inters module (toolbox.F90:25-99), including its optional device sync
static const char* kStageNames[PROF_NCAT] = {"potrf", "hegst", "hetrd_panel", "hetrd_her2k", "stedc", "ormtr", "trsm", "other"};
void prof_begin(int cat, cudaStream_t s) {
  if (opts().nvtx) { if (opts().nvtx > 1) cudaStreamSynchronize(s); nvtxRangePushA(kStageNames[cat]); }
  if (!g_prof_on) return;
  g_open[cat] = prof_event();
  cudaEventRecord(g_open[cat], s);
}
void prof_end(int cat, cudaStream_t s) {
  if (opts().nvtx) { if (opts().nvtx > 1) cudaStreamSynchronize(s); nvtxRangePop(); }
  if (!g_prof_on) return;
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, s);
  g_pairs.push_back({g_open[cat], b, cat});
}
void prof_enable(int on) { g_prof_on = on != 0; }
void prof_reset() {
  for (auto& p : g_pairs) { g_pool.push_back(p.a); g_pool.push_back(p.b); }
  g_pairs.clear();
  for (int i = 0; i < PROF_NCAT; ++i) { g_ms[i] = 0; g_cnt[i] = 0; }
  g_launches = 0;
}
// resolves all recorded pairs (synchronises the device) and returns accumulated ms / count per category
void prof_collect(double* ms, int* cnt, long long* launches) {
  cudaDeviceSynchronize();
  for (auto& p : g_pairs) {
    float t = 0;
    if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) { g_ms[p.cat] += t; g_cnt[p.cat] += 1; }
    g_pool.push_back(p.a); g_pool.push_back(p.b);
  }
  g_pairs.clear();
  for (int i = 0; i < PROF_NCAT; ++i) { ms[i] = g_ms[i]; cnt[i] = g_cnt[i]; }
  *launches = g_launches;
}

int status_fetch(cudaStream_t s) {
  Context& c = ctx();
  EIGB_CUDA_CHECK(cudaMemcpyAsync(c.h_status, c.d_info, ST_NWORDS * sizeof(int), cudaMemcpyDeviceToHost, s));
  EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  return 0;
}
int status_check(const char* who, int* potrf_pivot) {
  const int* h = ctx().h_status;
  if (potrf_pivot) *potrf_pivot = h[ST_POTRF];
  if (h[ST_POTRF] != 0) { set_last_error("%s error: potrf failed (B not positive definite at pivot %d)", who, h[ST_POTRF]); return -1; }
  if (h[ST_HETRD] != 0) { set_last_error("%s error: hetrd device status %d (grid barrier / exchange watchdog)", who, h[ST_HETRD]); return -1; }
  if (h[ST_STEDC] != 0) {
    set_last_error("%s error: stedc failed (%s)", who, h[ST_STEDC] == 1 ? "leaf QL iteration did not converge" :
                   h[ST_STEDC] == 2 ? "secular equation did not converge" : "non-finite eigenvalue");
    return -1;
  }
  return 0;
}

Options& opts() { static Options o; return o; }
MgConfig& mg() { static MgConfig m; return m; }
int set_option(const char* name, int value) {
  Options& o = opts();
  if (!strcmp(name, "trd_nb")) { if (value < 1 || value > 128) return -1; o.trd_nb = value; return 0; }
  if (!strcmp(name, "bt_nb")) { if (value < 1 || value > 256) return -1; o.bt_nb = value; return 0; }
  if (!strcmp(name, "symv_tma")) { o.symv_tma = value; return 0; }
  if (!strcmp(name, "trd_coop")) { o.trd_coop = value; return 0; }
  if (!strcmp(name, "trd_trace")) { o.trd_trace = value; return 0; }
  if (!strcmp(name, "nvtx")) { o.nvtx = value; return 0; }
  if (!strcmp(name, "hegst_hb")) { if (value < 0) return -1; o.hegst_hb = value; return 0; }
  if (!strcmp(name, "trd_l2keep_mb")) { if (value < 0 || value > 4096) return -1; o.trd_l2keep_mb = value; return 0; }
  if (!strcmp(name, "trsm_leaf256")) { o.trsm_leaf256 = value; return 0; }
  if (!strcmp(name, "potrf_pb")) { if (value > 8192) return -1; o.potrf_pb = value; return 0; }
  if (!strcmp(name, "gemm_tma")) { o.gemm_tma = value; return 0; }
  if (!strcmp(name, "gemm_tma_dbg")) { o.gemm_tma_dbg = value; return 0; }
  if (!strcmp(name, "trd_ctab")) { o.trd_ctab = value != 0; return 0; }
  if (!strcmp(name, "trd_upc")) { if ((value & 255) < 1 || (value & 255) > 64 || (value >> 8) > 32) return -1; o.trd_upc = value; return 0; }
  if (!strcmp(name, "trd_prefetch")) { if (value < -1 || value > 64) return -1; o.trd_prefetch = value; return 0; }
  if (!strcmp(name, "mg_switch_n")) { o.mg_switch_n = value; return 0; }
  if (!strcmp(name, "mg_dist_min_n")) { o.mg_dist_min_n = value; return 0; }
  if (!strcmp(name, "mg_gather_z")) { o.mg_gather_z = value; return 0; }
  if (!strcmp(name, "mg_potrf_min_n")) { o.mg_potrf_min_n = value; return 0; }
  if (!strcmp(name, "verbose")) { ctx().verbose = value; return 0; }
  return -1;
}
int get_option(const char* name) {
  Options& o = opts();
  if (!strcmp(name, "trd_nb")) return o.trd_nb;
  if (!strcmp(name, "bt_nb")) return o.bt_nb;
  if (!strcmp(name, "symv_tma")) return o.symv_tma;
  if (!strcmp(name, "trd_coop")) return o.trd_coop;
  if (!strcmp(name, "nvtx")) return o.nvtx;
  if (!strcmp(name, "hegst_hb")) return o.hegst_hb;
  if (!strcmp(name, "trd_l2keep_mb")) return o.trd_l2keep_mb;
  if (!strcmp(name, "trsm_leaf256")) return o.trsm_leaf256;
  if (!strcmp(name, "potrf_pb")) return o.potrf_pb;
  if (!strcmp(name, "gemm_tma")) return o.gemm_tma;
  if (!strcmp(name, "trd_upc")) return o.trd_upc;
  if (!strcmp(name, "trd_ctab")) return o.trd_ctab;
  if (!strcmp(name, "trd_prefetch")) return o.trd_prefetch;
  if (!strcmp(name, "mg_switch_n")) return o.mg_switch_n;
  if (!strcmp(name, "mg_dist_min_n")) return o.mg_dist_min_n;
  if (!strcmp(name, "mg_gather_z")) return o.mg_gather_z;
  if (!strcmp(name, "mg_potrf_min_n")) return o.mg_potrf_min_n;
  if (!strcmp(name, "verbose")) return ctx().verbose;
  return -1;
}
}  // namespace eigb200
