// eigb200 -- multi-GPU drivers behind the C ABI: eigb200_{dsygvdx,zhegvdx}_mg.  One process per GPU, ONE problem.
//
// The reference is single-GPU ("replicas only"); the partition follows BASELINE.json's north star:
//   * tridiagonalization: trailing matrix 1-D block-cyclic by 64-wide tile columns, the per-column partial products
//     exchanged INSIDE the persistent panel kernel through peer-mapped buffers over NVLink (sytrd.cu, MG variant);
//     once per panel the owner broadcasts the panel's columns (NCCL);
//   * reduction to standard form, back-transformation and the final solve with U: split by right-hand-side columns,
//     one exchange of column blocks each (NCCL broadcasts, grouped);
//   * Cholesky of B and the tridiagonal divide & conquer: replicated (bitwise deterministic kernels, so replicas agree).
// NCCL is used for setup and bulk exchanges only and is resolved at run time with dlopen("libnccl.so.2") -- inside a
// PyTorch process that is the copy torch already loaded, for a Fortran/MPI caller the system one; libeigb200.so itself
// keeps linking libcudart only.  The caller owns the rendezvous: rank 0 asks for the 128-byte NCCL id
// (eigb200_mg_unique_id), distributes it (MPI_Bcast / torch.distributed) and every rank calls eigb200_mg_init.
#include "common.cuh"
#include "gemm.cuh"
#include "stages.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <vector>

namespace eigb200 {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) { set_last_error("eigb200 multi-GPU: libnccl.so.2 not found (%s)", dlerror()); return nullptr; }
#define EIGB_SYM(field, name)                                                                \
  *(void**)(&api.field) = dlsym(api.handle, name);                                           \
  if (!api.field) { set_last_error("eigb200 multi-GPU: symbol %s missing from libnccl", name); api.handle = nullptr; return nullptr; }
  EIGB_SYM(GetUniqueId, "ncclGetUniqueId")
  EIGB_SYM(CommInitRank, "ncclCommInitRank")
  EIGB_SYM(CommDestroy, "ncclCommDestroy")
  EIGB_SYM(Broadcast, "ncclBroadcast")
  EIGB_SYM(AllGather, "ncclAllGather")
  EIGB_SYM(AllReduce, "ncclAllReduce")
  EIGB_SYM(GroupStart, "ncclGroupStart")
  EIGB_SYM(GroupEnd, "ncclGroupEnd")
  EIGB_SYM(GetErrorString, "ncclGetErrorString")
#undef EIGB_SYM
  return &api;
}

#define EIGB_NCCL_CHECK(expr)                                                                         \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess) {                                                                          \
      set_last_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, nccl_api()->GetErrorString(_r)); \
      return -1;                                                                                      \
    }                                                                                                 \
  } while (0)

// out(r, c) = conj(in(c0 + c, r)),  r in [0, n), c in [0, nc): the conjugate-transposed row block [c0, c0+nc) of `in`
template <typename T>
__global__ void __launch_bounds__(256) conj_transpose_rows_kernel(const T* __restrict__ in, int64_t ldi, int n, int c0, int nc,
                                                                  T* __restrict__ out, int64_t ldo) {
  __shared__ T tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;     // bx: rows of `in` block (0..nc), by: columns of `in` (0..n)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int ri = bx + tx, ci = by + k;
    if (ri < nc && ci < n) tile[k][tx] = in[(int64_t)(c0 + ri) + (int64_t)ci * ldi];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int ro = by + tx, co = bx + k;                    // out row = column of in, out col = row of in block
    if (ro < n && co < nc) out[(int64_t)ro + (int64_t)co * ldo] = conj_(tile[tx][k]);
  }
}

template <typename T>
__global__ void select_columns_mg_kernel(const double* __restrict__ Q, int64_t ldq, int n, int c0, int m, T* Z, int64_t ldz) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r < n && c < m) Z[r + (int64_t)c * ldz] = from_real<T>(Q[r + (int64_t)(c0 + c) * ldq]);
}

}  // namespace

// contiguous column ranges [c0, c1) per rank, boundaries aligned to 64 (the TRSM / tile block size); the same
// formula as multi_gpu.column_ranges (tests/test_multi_gpu_cpu.py exercises the partition on gloo)
void mg_column_range(int ncols, int world, int rank, int& c0, int& c1) {
  const int align = 64;
  const long long nblk = (ncols + align - 1) / align;
  const long long b0 = (nblk * rank) / world, b1 = (nblk * (rank + 1)) / world;
  c0 = (int)(b0 * align < ncols ? b0 * align : ncols);
  c1 = (int)(b1 * align < ncols ? b1 * align : ncols);
}

// flags per source rank for order <= n: one per (CTA, 32-row group of the CTA's rows) + one for v^H A v; the rows are dealt
// to G <= 256 CTAs in blocks of R = roundup8(ceil(n / G)) rows, so G * ceil(R / 32) <= (n + 39 G) / 32
int mg_flag_stride(int n) { return (n + 31) / 32 + 328; }

int mg_unique_id(char* id128) {
  NcclApi* N = nccl_api();
  if (!N) return -1;
  ncclUniqueId id;
  EIGB_NCCL_CHECK(N->GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}

static void mg_release_exchange() {
  MgConfig& M = mg();
  if (!M.own_exchange) return;
  for (int q = 0; q < M.P; ++q) {
    if (q == M.rank) continue;
    if (M.wbuf[q]) cudaIpcCloseMemHandle(M.wbuf[q]);
    if (M.flags[q]) cudaIpcCloseMemHandle(M.flags[q]);
  }
  if (M.wbuf[M.rank]) cudaFree(M.wbuf[M.rank]);
  if (M.flags[M.rank]) cudaFree(M.flags[M.rank]);
  for (int q = 0; q < 8; ++q) { M.wbuf[q] = nullptr; M.flags[q] = nullptr; }
  M.wbuf_bytes = 0;
  M.own_exchange = false;
}

int mg_finalize() {
  MgConfig& M = mg();
  cudaDeviceSynchronize();
  mg_release_exchange();
  if (M.comm) { NcclApi* N = nccl_api(); if (N) N->CommDestroy((ncclComm_t)M.comm); M.comm = nullptr; }
  M.rank = 0; M.P = 1; M.hook = nullptr; M.active = false; M.seq = 0;
  return 0;
}

int mg_init(int rank, int world, const char* id128) {
  NcclApi* N = nccl_api();
  if (!N) return -1;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) { set_last_error("eigb200_mg_init: bad rank/world (world <= 8)"); return -1; }
  mg_finalize();
  MgConfig& M = mg();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  EIGB_NCCL_CHECK(N->CommInitRank(&comm, world, id, rank));
  M.comm = comm; M.rank = rank; M.P = world; M.hook = nullptr; M.active = false; M.seq = 0;
  return 0;
}

// Exchange buffers of the distributed tridiagonalization for order <= n: [P][2][n + 64] complex elements + flags per rank,
// exported with CUDA IPC, handles all-gathered through the communicator, peers mapped.  Collective (all ranks, same n).
int mg_ensure_exchange(cudaStream_t s, int n) {
  MgConfig& M = mg();
  NcclApi* N = nccl_api();
  if (!N || !M.comm) { set_last_error("eigb200 multi-GPU: eigb200_mg_init has not been called"); return -1; }
  const int64_t need = (int64_t)M.P * 2 * ((int64_t)n + 64) * 16;
  if (M.own_exchange && M.wbuf_bytes >= need) return 0;
  EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  mg_release_exchange();
  void *w = nullptr, *f = nullptr;
  const int fstride = mg_flag_stride(n);
  const size_t fbytes = (size_t)M.P * fstride * sizeof(unsigned long long);
  EIGB_CUDA_CHECK(cudaMalloc(&w, (size_t)need));
  EIGB_CUDA_CHECK(cudaMalloc(&f, fbytes));
  EIGB_CUDA_CHECK(cudaMemset(w, 0xFF, (size_t)need));      // every word "unset" (sytrd.cu, multi-GPU exchange)
  EIGB_CUDA_CHECK(cudaMemset(f, 0, fbytes));
  struct Handles { cudaIpcMemHandle_t hw, hf; } mine;
  EIGB_CUDA_CHECK(cudaIpcGetMemHandle(&mine.hw, w));
  EIGB_CUDA_CHECK(cudaIpcGetMemHandle(&mine.hf, f));
  char* dbuf = nullptr;
  EIGB_CUDA_CHECK(cudaMalloc((void**)&dbuf, sizeof(Handles) * (size_t)(M.P + 1)));
  EIGB_CUDA_CHECK(cudaMemcpyAsync(dbuf, &mine, sizeof(Handles), cudaMemcpyHostToDevice, s));
  EIGB_NCCL_CHECK(N->AllGather(dbuf, dbuf + sizeof(Handles), sizeof(Handles), ncclChar, (ncclComm_t)M.comm, s));
  std::vector<Handles> all(M.P);
  EIGB_CUDA_CHECK(cudaMemcpyAsync(all.data(), dbuf + sizeof(Handles), sizeof(Handles) * (size_t)M.P, cudaMemcpyDeviceToHost, s));
  EIGB_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(dbuf);
  for (int q = 0; q < M.P; ++q) {
    if (q == M.rank) { M.wbuf[q] = w; M.flags[q] = (unsigned long long*)f; continue; }
    void *pw = nullptr, *pf = nullptr;
    EIGB_CUDA_CHECK(cudaIpcOpenMemHandle(&pw, all[q].hw, cudaIpcMemLazyEnablePeerAccess));
    EIGB_CUDA_CHECK(cudaIpcOpenMemHandle(&pf, all[q].hf, cudaIpcMemLazyEnablePeerAccess));
    M.wbuf[q] = pw; M.flags[q] = (unsigned long long*)pf;
  }
  M.wbuf_bytes = need;
  M.flag_stride = fstride;
  M.own_exchange = true;
  M.seq = 0;
  return 0;
}

// broadcast of columns [c0, c0+nc) (all ld rows: the columns are contiguous) from `owner`; owner < 0: all-gather of the
// 64-wide tile columns of the leading c0 columns from their block-cyclic owners (one NCCL group)
int mg_bcast_columns(cudaStream_t s, void* A, int64_t ld, int c0, int nc, int owner, int elem_bytes) {
  MgConfig& M = mg();
  NcclApi* N = nccl_api();
  if (!N || !M.comm) return -1;
  char* base = (char*)A;
  if (owner >= 0) {
    if (nc <= 0) return 0;
    EIGB_NCCL_CHECK(N->Broadcast(base + (size_t)c0 * ld * elem_bytes, base + (size_t)c0 * ld * elem_bytes,
                                 (size_t)nc * ld * elem_bytes, ncclChar, owner, (ncclComm_t)M.comm, s));
    return 0;
  }
  EIGB_NCCL_CHECK(N->GroupStart());
  for (int c = 0; c < c0; c += 64) {
    const int w = c0 - c < 64 ? c0 - c : 64;
    char* p = base + (size_t)c * ld * elem_bytes;
    ncclResult_t r = N->Broadcast(p, p, (size_t)w * ld * elem_bytes, ncclChar, (c / 64) % M.P, (ncclComm_t)M.comm, s);
    if (r != ncclSuccess) { N->GroupEnd(); set_last_error("ncclBroadcast failed: %s", N->GetErrorString(r)); return -1; }
  }
  EIGB_NCCL_CHECK(N->GroupEnd());
  return 0;
}

// plain pieces for stage code that packs its own buffers (distributed Cholesky): broadcast of a contiguous buffer, NCCL
// group brackets, max-reduction of one device int (status words)
int mg_bcast(cudaStream_t s, void* ptr, size_t bytes, int root) {
  MgConfig& M = mg();
  NcclApi* N = nccl_api();
  if (!N || !M.comm) return -1;
  if (bytes == 0) return 0;
  EIGB_NCCL_CHECK(N->Broadcast(ptr, ptr, bytes, ncclChar, root, (ncclComm_t)M.comm, s));
  return 0;
}
int mg_group(bool start) {
  NcclApi* N = nccl_api();
  if (!N) return -1;
  EIGB_NCCL_CHECK(start ? N->GroupStart() : N->GroupEnd());
  return 0;
}
int mg_allreduce_max_int(cudaStream_t s, int* dptr) {
  MgConfig& M = mg();
  NcclApi* N = nccl_api();
  if (!N || !M.comm) return -1;
  EIGB_NCCL_CHECK(N->AllReduce(dptr, dptr, 1, ncclInt32, ncclMax, (ncclComm_t)M.comm, s));
  return 0;
}

// every rank owns the contiguous column block mg_column_range(ncols, P, rank) of the ld x ncols matrix; after the call
// all ranks hold all blocks (in place; one NCCL group of P broadcasts)
int mg_allgather_columns(cudaStream_t s, void* A, int64_t ld, int ncols, int elem_bytes) {
  MgConfig& M = mg();
  NcclApi* N = nccl_api();
  if (!N || !M.comm) { set_last_error("eigb200 multi-GPU: eigb200_mg_init has not been called"); return -1; }
  if (M.P == 1) return 0;
  char* base = (char*)A;
  EIGB_NCCL_CHECK(N->GroupStart());
  for (int q = 0; q < M.P; ++q) {
    int c0, c1;
    mg_column_range(ncols, M.P, q, c0, c1);
    if (c1 <= c0) continue;
    char* p = base + (size_t)c0 * ld * elem_bytes;
    ncclResult_t r = N->Broadcast(p, p, (size_t)(c1 - c0) * ld * elem_bytes, ncclChar, q, (ncclComm_t)M.comm, s);
    if (r != ncclSuccess) { N->GroupEnd(); set_last_error("ncclBroadcast failed: %s", N->GetErrorString(r)); return -1; }
  }
  EIGB_NCCL_CHECK(N->GroupEnd());
  return 0;
}

// Distributed generalized solve.  Same argument list and checks as the single-GPU driver; every rank passes the SAME
// A and B (replicated device inputs).  On exit, on every rank: B <- U, w(1:N) all eigenvalues, Z(:, 1:m) the
// eigenvectors il..iu (all of them: the column blocks are gathered), A destroyed (both triangles -- the single-GPU
// driver's "strict lower triangle preserved" needs Z as a save area, which this driver uses as workspace).
template <typename T>
int hegvdx_mg_driver(int n, T* A, int lda, T* B, int ldb, T* Z, int ldz, int il, int iu, double* w, T* work, int lwork,
                     double* rwork, int lrwork, int lwork_h, int lrwork_h, int liwork_h, T* Z_h, int ldz_h, double* w_h,
                     int* info, int skip_host_copy) {
  MgConfig& M = mg();
  if (M.comm == nullptr || M.P <= 1)
    return hegvdx_driver<T>(n, A, lda, B, ldb, Z, ldz, il, iu, w, work, lwork, rwork, lrwork, lwork_h, lrwork_h, liwork_h, Z_h,
                            ldz_h, w_h, info, skip_host_copy);
  const bool cplx = is_cplx<T>::value;
  const char* name = cplx ? "zhegvdx_gpu (multi-GPU)" : "dsygvdx_gpu (multi-GPU)";
  *info = 0;
  ctx().a_ready = nullptr;
  if (hegvdx_check_args<T>(n, lda, ldb, ldz, il, iu, lwork, lrwork, lwork_h, lrwork_h, liwork_h, info) != 0) return -1;
  if (n == 0) return 0;
#define EIGB_MG_FAIL(cond, what)                                                             \
  if (cond) { printf(" %s error: %s failed: %s\n", name, what, "see eigb200_last_error()"); *info = -1; M.active = false; return -1; }
  cudaStream_t s = ctx().stream;
  const int m = iu - il + 1;
  const int P = M.P, rank = M.rank;
  const int es = (int)sizeof(T);
  // distributed tridiagonalization only when the per-column tile work outweighs the exchange latency
  const int min_n = opts().mg_dist_min_n >= 0 ? opts().mg_dist_min_n : (P <= 2 ? 6144 : 4096);
  const bool dist_trd = n >= min_n;
  if (dist_trd) EIGB_MG_FAIL(mg_ensure_exchange(s, n) != 0, "exchange buffer setup");
  // 1. Cholesky: block columns dealt cyclically over the ranks from mg_potrf_min_n on (trsm.cu, potrf_upper_mg), else
  //    replicated (deterministic => identical U everywhere)
  prof_begin(PROF_POTRF, s);
  int rc = (opts().mg_potrf_min_n >= 0 && n >= opts().mg_potrf_min_n) ? potrf_upper_mg<T>(s, n, B, ldb)
                                                                       : potrf_upper<T>(s, n, B, ldb, nullptr, /*sync_status=*/false);
  prof_end(PROF_POTRF, s);
  EIGB_MG_FAIL(rc != 0, "potrf");
  // 2. C = U^-H A U^-1 by two column-parallel left solves: Y = U^-H A on this rank's columns, exchange, then
  //    C^H(:, cols) = U^-H Y^H(:, cols) with C = C^H (both triangles of the result are formed)
  prof_begin(PROF_HEGST, s);
  int c0, c1;
  mg_column_range(n, P, rank, c0, c1);
  const int nc = c1 - c0;
  rc = symmetrize_from_upper<T>(s, n, A, lda, (T*)nullptr, 0);
  if (rc == 0 && nc > 0) rc = trsm_upper<T>(s, 'L', 'C', n, nc, B, ldb, A + (int64_t)c0 * lda, lda);
  if (rc == 0) rc = mg_allgather_columns(s, A, lda, n, es);
  if (rc == 0 && nc > 0) {
    conj_transpose_rows_kernel<T><<<dim3(cdiv(nc, 32), cdiv(n, 32)), 256, 0, s>>>(A, lda, n, c0, nc, Z, ldz);
    count_launch(1);
    rc = trsm_upper<T>(s, 'L', 'C', n, nc, B, ldb, Z, ldz);
    if (rc == 0 && cudaMemcpy2DAsync(A + (int64_t)c0 * lda, (size_t)lda * es, Z, (size_t)ldz * es, (size_t)n * es, nc,
                                     cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = -1;
  }
  if (rc == 0) rc = mg_allgather_columns(s, A, lda, n, es);
  prof_end(PROF_HEGST, s);
  EIGB_MG_FAIL(rc != 0, "reduction to standard form");
  // 3. tridiagonalization: distributed trailing matrix (identical d, e, tau and reflectors on all ranks)
  double* d_e = cplx ? rwork : reinterpret_cast<double*>(work);
  T* d_tau = cplx ? work : work + n;
  M.active = dist_trd;
  rc = hetrd_upper<T>(s, n, A, lda, w, d_e, d_tau, /*sync_status=*/false);
  M.active = false;
  EIGB_MG_FAIL(rc != 0, "hetrd");
  // 4. divide & conquer, replicated; then this rank's block of eigenvector columns
  const size_t nn = (size_t)n * n * sizeof(double);
  int z0, z1;
  mg_column_range(m, P, rank, z0, z1);
  const int mz = z1 - z0;
  const size_t need_dc = nn + 256 + stedc_scratch_bytes(n);
  const size_t need_bt = ormtr_scratch_bytes(n, mz > 0 ? mz : 1, sizeof(T));
  char* scr = (char*)ctx_scratch(need_dc > need_bt ? need_dc : need_bt);
  EIGB_MG_FAIL(scr == nullptr, "scratch allocation");
  double* Qt = (double*)scr;
  const size_t qoff = (nn + 255) & ~size_t(255);
  // (eigenvectors of T only for this rank's columns: the root merge of the divide & conquer is thereby split over the ranks)
  EIGB_MG_FAIL(stedc_device(s, n, w, d_e, Qt, n, scr + qoff, ctx().scratch_bytes - qoff, il - 1 + z0, il - 1 + z1) != 0, "stedc");
  if (mz > 0) {
    T* Zb = Z + (int64_t)z0 * ldz;
    select_columns_mg_kernel<T><<<dim3(cdiv(n, 256), mz), 256, 0, s>>>(Qt, n, n, il - 1 + z0, mz, Zb, ldz);
    count_launch(1);
    EIGB_MG_FAIL(ormtr_upper<T>(s, n, mz, A, lda, d_tau, Zb, ldz, scr, ctx().scratch_bytes) != 0, "back-transformation");
    prof_begin(PROF_TRSM, s);
    rc = trsm_upper<T>(s, 'L', 'N', n, mz, B, ldb, Zb, ldz);
    prof_end(PROF_TRSM, s);
    EIGB_MG_FAIL(rc != 0, "final solve");
  }
  if (opts().mg_gather_z) EIGB_MG_FAIL(mg_allgather_columns(s, Z, ldz, m, es) != 0, "gather of Z");
  // host copies (every rank gets the full result) + the deferred status words
  if (w_h && cudaMemcpyAsync(w_h, w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = -1;
  if (!skip_host_copy && Z_h &&
      cudaMemcpy2DAsync(Z_h, (size_t)ldz_h * es, Z, (size_t)ldz * es, (size_t)n * es, m, cudaMemcpyDeviceToHost, s) != cudaSuccess)
    rc = -1;
  EIGB_MG_FAIL(rc != 0 || status_fetch(s) != 0, "copy to host");
  if (status_check(name) != 0) { printf(" %s error: %s\n", name, "see eigb200_last_error()"); *info = -1; return -1; }
#undef EIGB_MG_FAIL
  return 0;
}

template int hegvdx_mg_driver<double>(int, double*, int, double*, int, double*, int, int, int, double*, double*, int,
                                      double*, int, int, int, int, double*, int, double*, int*, int);
template int hegvdx_mg_driver<double2>(int, double2*, int, double2*, int, double2*, int, int, int, double*, double2*, int,
                                       double*, int, int, int, int, double2*, int, double*, int*, int);

}  // namespace eigb200
