// Microbenchmark (2 GPUs, one process): cost per kernel of a ring halo exchange done INSIDE back-to-back kernels.
// Each kernel: G edge blocks wait for the neighbour's rows of the previous exchange, then store their own rows into the
// neighbour's buffer.  Variant 0: data stores + __threadfence_system + block ticket + st.release.sys flag (the r2
// protocol).  Variant 1: LL protocol -- every 8-byte value travels with its 8-byte sequence tag in ONE 16-byte store,
// the consumer polls the elements it needs; no fences, no tickets.  Variant 2: like 0 but only lane 0 of each block fences.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o halo_latency halo_latency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_ll(void* p, double v, unsigned long long tag) {
  asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory"); }
__device__ __forceinline__ void ld_ll(const void* p, double& v, unsigned long long& tag) {
  long long a; asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(tag) : "l"(p) : "memory");
  v = __longlong_as_double(a); }

struct Buf { unsigned long long* flag; double* data; double* ll; unsigned int* ticket; unsigned long long* seq; };

// n columns, 4 fields; G = n/120 blocks of 128 threads, each thread <= 1 column
template <int VAR>
__global__ void k_exchange(Buf mine, Buf peer, int n, double* sink, int work_iters) {
  const unsigned long long s = *mine.seq + 1ull;   // this kernel produces exchange s and consumes s-1
  const int slot = (int)(s & 1ull), pslot = slot ^ 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (VAR == 1) {
    if (c < n && s > 1ull)
      for (int k = 0; k < 4; ++k) {
        double v; unsigned long long t;
        do { ld_ll(mine.ll + 2 * ((size_t)(pslot * 4 + k) * n + c), v, t); } while (t != s - 1ull);
        acc += v;
      }
  } else {
    if ((threadIdx.x & 31) == 0 && s > 1ull) while (ld_acq(mine.flag + pslot * 16) < s - 1ull) {}
    __syncwarp();
    if (c < n && s > 1ull) for (int k = 0; k < 4; ++k) acc += mine.data[(size_t)(pslot * 4 + k) * n + c];
  }
  for (int i = 0; i < work_iters; ++i) acc = acc * 1.0000001 + 1e-9;   // stands for the strip's march
  if (VAR == 1) {
    if (c < n) for (int k = 0; k < 4; ++k) st_ll(peer.ll + 2 * ((size_t)(slot * 4 + k) * n + c), acc + k, s);
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(mine.ticket, 1u) == gridDim.x - 1) { *mine.ticket = 0u; *mine.seq = s; }
  } else {
    if (c < n) for (int k = 0; k < 4; ++k) peer.data[(size_t)(slot * 4 + k) * n + c] = acc + k;
    if (VAR == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (VAR == 2) __threadfence_system();
      if (atomicAdd(mine.ticket, 1u) == gridDim.x - 1) {
        *mine.ticket = 0u; *mine.seq = s;
        __threadfence_system();
        st_rel(peer.flag + slot * 16, s);
      }
    }
  }
  if (sink && acc == 123.456) sink[0] = acc;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4096;
  const int K = 400;
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  Buf b[2]; cudaStream_t st[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0)); CK(cudaStreamCreate(&st[d]));
    CK(cudaMalloc(&b[d].flag, 4096)); CK(cudaMalloc(&b[d].data, sizeof(double) * 8 * n)); CK(cudaMalloc(&b[d].ll, sizeof(double) * 16 * n));
    CK(cudaMalloc(&b[d].ticket, 64)); CK(cudaMalloc(&b[d].seq, 64));
  }
  const int G = (n + 127) / 128;
  for (int var = 0; var < 3; ++var)
    for (int work : {0, 20000}) {
      for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d)); CK(cudaMemset(b[d].flag, 0, 4096)); CK(cudaMemset(b[d].ll, 0, sizeof(double) * 16 * n));
        CK(cudaMemset(b[d].ticket, 0, 64)); CK(cudaMemset(b[d].seq, 0, 64)); CK(cudaDeviceSynchronize());
      }
      cudaGraphExec_t ge[2];
      for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        cudaGraph_t g; CK(cudaStreamBeginCapture(st[d], cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < K; ++i) {
          if (var == 0) k_exchange<0><<<G, 128, 0, st[d]>>>(b[d], b[1 - d], n, nullptr, work);
          if (var == 1) k_exchange<1><<<G, 128, 0, st[d]>>>(b[d], b[1 - d], n, nullptr, work);
          if (var == 2) k_exchange<2><<<G, 128, 0, st[d]>>>(b[d], b[1 - d], n, nullptr, work);
        }
        CK(cudaStreamEndCapture(st[d], &g)); CK(cudaGraphInstantiate(&ge[d], g, 0)); CK(cudaGraphDestroy(g));
      }
      float ms[2] = {0, 0};
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0[2], e1[2];
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d])); }
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); CK(cudaGraphLaunch(ge[d], st[d])); CK(cudaEventRecord(e1[d], st[d])); }
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); CK(cudaEventElapsedTime(&ms[d], e0[d], e1[d])); }
      }
      printf("variant %d (%s) n=%d blocks=%d work=%d: %.2f / %.2f us per kernel\n", var,
             var == 0 ? "fence by all threads + flag" : var == 1 ? "LL 16-byte value+tag" : "fence by one thread + flag", n, G, work,
             ms[0] * 1e3 / K, ms[1] * 1e3 / K);
      for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaGraphExecDestroy(ge[d])); }
    }
  return 0;
}
