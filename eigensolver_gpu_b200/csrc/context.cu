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
  if (c.initialized) {   // device changed: drop per-device resources
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr; c.scratch_bytes = 0;
    if (c.d_info) cudaFree(c.d_info);
    c.d_info = nullptr;
    c.stream2 = nullptr; c.ev1 = c.ev2 = nullptr;
  }
  cudaDeviceProp prop;
  EIGB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    set_last_error("eigb200 requires an sm_100a (B200) device, found sm_%d%d", prop.major, prop.minor);
    return -1;
  }
  c.device = dev;
  c.num_sms = prop.multiProcessorCount;
  EIGB_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  EIGB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev1, cudaEventDisableTiming));
  EIGB_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev2, cudaEventDisableTiming));
  EIGB_CUDA_CHECK(cudaMalloc(&c.d_info, 64 * sizeof(int)));
  EIGB_CUDA_CHECK(cudaMemset(c.d_info, 0, 64 * sizeof(int)));
  const char* v = getenv("EIGB200_VERBOSE");
  c.verbose = v ? atoi(v) : 0;
  c.initialized = true;
  return 0;
}

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
namespace eigb200 {
Options& opts() { static Options o; return o; }
int set_option(const char* name, int value) {
  Options& o = opts();
  if (!strcmp(name, "trd_nb")) { if (value < 1 || value > 128) return -1; o.trd_nb = value; return 0; }
  if (!strcmp(name, "bt_nb")) { if (value < 1 || value > 256) return -1; o.bt_nb = value; return 0; }
  if (!strcmp(name, "symv_tma")) { o.symv_tma = value; return 0; }
  if (!strcmp(name, "trd_coop")) { o.trd_coop = value; return 0; }
  if (!strcmp(name, "verbose")) { ctx().verbose = value; return 0; }
  return -1;
}
int get_option(const char* name) {
  Options& o = opts();
  if (!strcmp(name, "trd_nb")) return o.trd_nb;
  if (!strcmp(name, "bt_nb")) return o.bt_nb;
  if (!strcmp(name, "symv_tma")) return o.symv_tma;
  if (!strcmp(name, "trd_coop")) return o.trd_coop;
  if (!strcmp(name, "verbose")) return ctx().verbose;
  return -1;
}
}  // namespace eigb200
