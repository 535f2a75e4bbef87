// Probe: FP64 peak on B200 through (a) DFMA chains, (b) mma.sync m8n8k4 f64 (DMMA), (c) m16n8k4/k8/k16 variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
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

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void dmma16_kernel(double* out, int iters, double a0, double b0) {
  double c[8][4]; double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = a0 + i; for (int j = 0; j < 4; ++j) c[i][j] = i + j; }
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = b0 * j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void read_kernel(const double2* __restrict__ in, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  if (s == 1.2345) out[0] = s;
}

template <class F> float timeit(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d clock %d kHz smem/block optin %zu l2 %d\n", p.name, p.multiProcessorCount, p.clockRate,
         p.sharedMemPerBlockOptin, p.l2CacheSize);
  int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, 0); printf("cooperative launch %d\n", coop);
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024);
  int sms = p.multiProcessorCount;
  for (int wpb : {4, 8, 16, 32}) {
    int threads = wpb * 32, blocks = sms * (wpb <= 8 ? 2 : 1), iters = 20000;
    float ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
    double fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("DFMA  threads/blk %4d blocks %4d: %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, fl / ms * 1e-9);
    ms = timeit([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
    fl = 2.0 * 256 * 16 * iters * (double)wpb * blocks;
    printf("DMMA884 threads/blk %4d blocks %4d: %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, fl / ms * 1e-9);
    ms = timeit([&] { dmma16_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, 3);
    fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)wpb * blocks;
    printf("DMMA16816 threads/blk %4d blocks %4d: %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, fl / ms * 1e-9);
  }
  // sustained: 2 seconds of DMMA
  {
    int threads = 256, blocks = sms * 2, iters = 200000;
    float ms = timeit([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    double fl = 2.0 * 256 * 16 * iters * 8.0 * blocks;
    printf("DMMA884 sustained: %.3f ms/launch  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 10);
    fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("DFMA sustained: %.3f ms/launch  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  // HBM copy / read bandwidth
  {
    size_t n = (size_t)1 << 27;  // 2 GiB per buffer of double2
    double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16);
    float ms = timeit([&] { copy_kernel<<<sms * 8, 512>>>(a, b, n); }, 5);
    printf("copy 2x%.1f GB: %.3f ms  %.1f GB/s\n", n * 16e-9, ms, 2.0 * n * 16 / ms * 1e-6);
    ms = timeit([&] { read_kernel<<<sms * 8, 512>>>(a, out, n); }, 5);
    printf("read %.1f GB: %.3f ms  %.1f GB/s\n", n * 16e-9, ms, 1.0 * n * 16 / ms * 1e-6);
    // L2-resident read: 48 MB
    size_t n2 = (size_t)3 << 20;
    ms = timeit([&] { read_kernel<<<sms * 8, 512>>>(a, out, n2); }, 50);
    printf("read L2-resident %.1f MB: %.4f ms  %.1f GB/s\n", n2 * 16e-6, ms, 1.0 * n2 * 16 / ms * 1e-6);
  }
  // launch latency: empty kernels back-to-back
  {
    float ms = timeit([&] { for (int i = 0; i < 1000; ++i) read_kernel<<<1, 32>>>(nullptr, out, 0); }, 3);
    printf("1000 dependent tiny launches: %.3f ms => %.2f us/launch\n", ms, ms);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
