// eigb200 -- live peak probes for the roofline denominators (bench.py): FP64 tensor (DMMA) and FMA throughput, HBM
// read-only and copy bandwidth.  MEASURED_PEAKS.json (driver-written) carries the HBM copy figure and a bf16 number
// but no FP64 one, so the FP64 denominator is measured on the device the bench runs on, right before it is used.
#include "common.cuh"
#include "stages.cuh"

namespace eigb200 {
namespace {

__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(512) probe_copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void __launch_bounds__(512) probe_read_kernel(const double2* __restrict__ in, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  if (s == 1.2345) out[0] = s;
}

template <class F>
float time_ms(cudaStream_t s, F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return ms / reps;
}

}  // namespace

// out[0] = DMMA TFLOP/s, out[1] = DFMA TFLOP/s, out[2] = HBM read GB/s, out[3] = HBM copy GB/s (read + write bytes)
int probe_peaks(cudaStream_t s, double* out) {
  const int sms = ctx().num_sms;
  const size_t n = (size_t)1 << 26;                      // 1 GiB per buffer of double2: 8x the L2
  void* scr = ctx_scratch(2 * n * sizeof(double2) + (size_t)sms * 2 * 256 * sizeof(double) + 4096);
  if (!scr) return -1;
  double2* a = (double2*)scr;
  double2* b = a + n;
  double* o = (double*)(b + n);
  EIGB_CUDA_CHECK(cudaMemsetAsync(a, 0, 2 * n * sizeof(double2), s));
  const int blocks = sms * 2, iters = 20000;
  float ms = time_ms(s, [&] { probe_dmma_kernel<<<blocks, 256, 0, s>>>(o, iters, 1.0000001, 1e-9); }, 3);
  out[0] = 2.0 * 256 * 16 * iters * 8.0 * blocks / ms * 1e-9;
  ms = time_ms(s, [&] { probe_dfma_kernel<<<blocks, 256, 0, s>>>(o, iters, 1.0000001, 1e-9); }, 3);
  out[1] = 2.0 * 16 * iters * 256.0 * blocks / ms * 1e-9;
  ms = time_ms(s, [&] { probe_read_kernel<<<sms * 8, 512, 0, s>>>(a, o, n); }, 10);
  out[2] = (double)n * 16 / ms * 1e-6;
  ms = time_ms(s, [&] { probe_copy_kernel<<<sms * 8, 512, 0, s>>>(a, b, n); }, 10);
  out[3] = 2.0 * n * 16 / ms * 1e-6;
  EIGB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace eigb200
