// Krylov vector kernels (fp64): deterministic two-stage reductions (warp shuffle -> block -> last
// block sums the per-block partials in a fixed order) and streaming axpy-type updates.
// Replaces np.dot / np.linalg.norm / vector updates inside the reference's Krylov solvers
// (solve.py:167-168, :285; utils.py:7-17).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ll.cuh"

namespace mpbp {

constexpr int kRedThreads = 256;
constexpr int kMaxRedBlocks = 1184;  // 148 SMs x 8
constexpr int kMaxMulti = 8;         // vectors handled per multi-dot / multi-axpy launch

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sums `v` over the block; result valid in thread 0. smem: kRedThreads/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < kRedThreads / 32) ? smem[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < kRedThreads / 32) ? smem[lane] : 0.0;
    v = warp_max(v);
  }
  return v;
}

// last-arriving block detection
__device__ __forceinline__ bool last_block(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  return is_last;
}

// ---------------------------------------------------------------------------------------------------------
// All-reduce (sum) of a few scalars over the ranks through peer memory, fused into the reduction kernels: the last
// block of the local reduction stores its partial sums into EVERY rank's comm buffer (NVLink peer stores) as LL
// elements (value + operation number in one 16-byte store, ll.cuh), polls its own buffer until all ranks' elements of
// this operation have arrived and adds them in rank order -- the result is bitwise identical on all ranks, no NCCL
// kernel is launched (an 8-byte ncclAllReduce costs 10-20 us) and no memory fence sits on the critical path.
// Operation number *rseq (the same on all ranks) alternates between two slots; a rank can only reach operation k+2
// after it has received every rank's elements of k+1, which each rank sent after it had finished reading slot k&1.
constexpr int kRedMaxVals = 16;
constexpr int kRedMaxRanks = 8;
constexpr size_t kRedBytes = 2 * kRedMaxRanks * kRedMaxVals * sizeof(LLElem);
struct RedCtx {
  char* peers[kRedMaxRanks];  // every rank's reduction area (mapped peer memory); peers[rank] is this rank's own
  int rank, nranks;           // nranks <= 1: no exchange
  unsigned long long* rseq;   // device-resident operation counter
};
__device__ __forceinline__ LLElem* red_vals(char* area, int slot, int src) {
  return reinterpret_cast<LLElem*>(area) + (size_t)(slot * kRedMaxRanks + src) * kRedMaxVals;
}
// called by every thread of ONE block (>= nranks * count threads); vals: `count` block-visible doubles (shared
// memory), overwritten by the sums
__device__ __forceinline__ void p2p_allreduce_sum(const RedCtx& rc, double* vals, int count) {
  if (rc.nranks <= 1) return;
  __shared__ double red_in[kRedMaxRanks * kRedMaxVals];
  const unsigned long long seq = *rc.rseq + 1ull;
  const int slot = (int)(seq & 1ull);
  const int t = threadIdx.x;
  int r = 0, i = 0;
  if (t < rc.nranks * count) {
    r = t / count;
    i = t - r * count;
    st_ll(red_vals(rc.peers[r], slot, rc.rank) + i, vals[i], seq);
  }
  __syncthreads();
  if (t < rc.nranks * count) red_in[r * count + i] = ll_wait(red_vals(rc.peers[rc.rank], slot, r) + i, seq);
  __syncthreads();
  if (t < count) {
    double s = 0.0;
    for (int q = 0; q < rc.nranks; ++q) s += red_in[q * count + t];
    vals[t] = s;
  }
  __syncthreads();
  if (t == 0) *rc.rseq = seq;
}

// out[k] = sum_i V_k[i] * w[i], k < NV (V_k = V + k*ld).  partial: gridDim.x * NV doubles.
// post: 0 none, 1 sqrt (nrm2 on one rank)
template <int NV>
__global__ void __launch_bounds__(kRedThreads) k_multi_dot(const double* __restrict__ V, size_t ld,
                                                           const double* __restrict__ w, size_t len,
                                                           double* __restrict__ partial, unsigned int* counter,
                                                           double* __restrict__ out, int post, RedCtx rc) {
  __shared__ double smem[kRedThreads / 32];
  __shared__ double red_sh[kRedMaxVals];
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (NV <= 2) {
    // four independent loads in flight per stream; the accumulation order is the plain loop's
    for (; i + 3 * stride < len; i += 4 * stride) {
      double wv[4], vv[NV][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = w[i + u * stride];
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int u = 0; u < 4; ++u) vv[k][u] = V[k * ld + i + u * stride];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = fma(vv[k][u], wv[u], acc[k]);
    }
  }
  for (; i < len; i += stride) {
    const double wi = w[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = fma(V[k * ld + i], wi, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = block_sum(acc[k], smem);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NV + k] = s;
  }
  if (last_block(counter)) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) s += partial[(size_t)bI * NV + k];
      s = block_sum(s, smem);
      if (threadIdx.x == 0) red_sh[k] = s;
    }
    __syncthreads();
    p2p_allreduce_sum(rc, red_sh, NV);
    if (threadIdx.x < NV) out[threadIdx.x] = (post == 1) ? sqrt(red_sh[threadIdx.x]) : red_sh[threadIdx.x];
    if (threadIdx.x == 0) *counter = 0u;
  }
}

struct Alphas {
  double a[kMaxMulti];
};

// Two multi-dots in one pass over the basis (low-synchronisation Gram-Schmidt, see gmres_right in plan.cu):
//   out[k] = <V_k, w>,   out[NV + k] = <V_k, u>,   k < NV      (u = the newest basis vector: its Gram column)
template <int NV>
__global__ void __launch_bounds__(kRedThreads) k_multi_dot2(const double* __restrict__ V, size_t ld,
                                                            const double* __restrict__ w, const double* __restrict__ u,
                                                            size_t len, double* __restrict__ partial,
                                                            unsigned int* counter, double* __restrict__ out, RedCtx rc) {
  __shared__ double smem[kRedThreads / 32];
  __shared__ double red_sh[kRedMaxVals];
  double aw[NV], au[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) aw[k] = au[k] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < len; i += 2 * stride) {  // two independent rows of loads in flight
    const double w0 = w[i], w1 = w[i + stride], u0 = u[i], u1 = u[i + stride];
    double v0[NV], v1[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      v0[k] = V[k * ld + i];
      v1[k] = V[k * ld + i + stride];
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      aw[k] = fma(v0[k], w0, aw[k]);
      au[k] = fma(v0[k], u0, au[k]);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      aw[k] = fma(v1[k], w1, aw[k]);
      au[k] = fma(v1[k], u1, au[k]);
    }
  }
  for (; i < len; i += stride) {
    const double wi = w[i], ui = u[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double vk = V[k * ld + i];
      aw[k] = fma(vk, wi, aw[k]);
      au[k] = fma(vk, ui, au[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = block_sum(aw[k], smem);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 2 * NV + k] = s;
    const double t = block_sum(au[k], smem);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * 2 * NV + NV + k] = t;
  }
  if (last_block(counter)) {
#pragma unroll
    for (int k = 0; k < 2 * NV; ++k) {
      double s = 0.0;
      for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) s += partial[(size_t)bI * 2 * NV + k];
      s = block_sum(s, smem);
      if (threadIdx.x == 0) red_sh[k] = s;
    }
    __syncthreads();
    p2p_allreduce_sum(rc, red_sh, 2 * NV);
    if (threadIdx.x < 2 * NV) out[threadIdx.x] = red_sh[threadIdx.x];
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// w -= sum_k h[k] V_k (h by value) fused with nrm = ||w_new|| (only the LAST launch of a chain asks for the norm)
template <int NV, bool NRM>
__global__ void __launch_bounds__(kRedThreads) k_multi_axpy_nrm(const double* __restrict__ V, size_t ld, Alphas h,
                                                                double* __restrict__ w, size_t len,
                                                                double* __restrict__ partial, unsigned int* counter,
                                                                double* __restrict__ out_nrm, RedCtx rc) {
  __shared__ double smem[kRedThreads / 32];
  __shared__ double red_sh[kRedMaxVals];
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    double wi = w[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) wi = fma(-h.a[k], V[k * ld + i], wi);
    w[i] = wi;
    if (NRM) acc = fma(wi, wi, acc);
  }
  if (!NRM) return;
  const double s = block_sum(acc, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_block(counter)) {
    double t = 0.0;
    for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) t += partial[bI];
    t = block_sum(t, smem);
    if (threadIdx.x == 0) red_sh[0] = t;
    __syncthreads();
    p2p_allreduce_sum(rc, red_sh, 1);
    if (threadIdx.x == 0) {
      out_nrm[0] = sqrt(red_sh[0]);
      *counter = 0u;
    }
  }
}

// One fused step of modified Gram-Schmidt (the `for k: h = dot(v_k, w); w -= h v_k` loop of scipy's gmres /
// the Arnoldi loop of fgmres):   w <- w - alpha_prev * v_prev ;  out_dot = <v_next, w> ;  out_nrm = ||w||^2
// (each part optional).  4 vector passes per basis vector instead of 5, and the per-thread summation
// order is exactly k_multi_dot<1>'s, so the result is bitwise the unfused dot-then-axpy sequence.
template <bool HAS_PREV, bool HAS_NEXT, bool NRM>
__global__ void __launch_bounds__(kRedThreads) k_mgs_fused(const double* __restrict__ alpha_dev,
                                                           const double* __restrict__ v_prev,
                                                           const double* __restrict__ v_next, double* __restrict__ w,
                                                           size_t len, double* __restrict__ partial,
                                                           unsigned int* counter, double* __restrict__ out_dot,
                                                           double* __restrict__ out_nrm, int post_sqrt, RedCtx rc) {
  __shared__ double smem[kRedThreads / 32];
  __shared__ double red_sh[kRedMaxVals];
  double acc = 0.0, accn = 0.0;
  double al = 0.0;
  if (HAS_PREV) al = -alpha_dev[0];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < len; i += 4 * stride) {  // four independent loads per stream in flight
    double wv[4], pv[4], nv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = w[i + u * stride];
    if (HAS_PREV) {
#pragma unroll
      for (int u = 0; u < 4; ++u) pv[u] = v_prev[i + u * stride];
    }
    if (HAS_NEXT) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nv[u] = v_next[i + u * stride];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double wi = wv[u];
      if (HAS_PREV) {
        wi = fma(al, pv[u], wi);
        w[i + u * stride] = wi;
      }
      if (HAS_NEXT) acc = fma(nv[u], wi, acc);
      if (NRM) accn = fma(wi, wi, accn);
    }
  }
  for (; i < len; i += stride) {
    double wi = w[i];
    if (HAS_PREV) {
      wi = fma(al, v_prev[i], wi);
      w[i] = wi;
    }
    if (HAS_NEXT) acc = fma(v_next[i], wi, acc);
    if (NRM) accn = fma(wi, wi, accn);
  }
  if (HAS_NEXT) {
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) partial[2 * blockIdx.x] = s;
  }
  if (NRM) {
    const double s = block_sum(accn, smem);
    if (threadIdx.x == 0) partial[2 * blockIdx.x + 1] = s;
  }
  if (!HAS_NEXT && !NRM) return;
  if (last_block(counter)) {
    if (HAS_NEXT) {
      double s = 0.0;
      for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) s += partial[2 * bI];
      s = block_sum(s, smem);
      if (threadIdx.x == 0) red_sh[0] = s;
    }
    if (NRM) {
      double s = 0.0;
      for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) s += partial[2 * bI + 1];
      s = block_sum(s, smem);
      if (threadIdx.x == 0) red_sh[1] = s;
    }
    __syncthreads();
    if (rc.nranks > 1) {
      if (!HAS_NEXT && threadIdx.x == 0) red_sh[0] = 0.0;
      if (!NRM && threadIdx.x == 0) red_sh[1] = 0.0;
      __syncthreads();
      p2p_allreduce_sum(rc, red_sh, 2);
    }
    if (threadIdx.x == 0) {
      if (HAS_NEXT) out_dot[0] = red_sh[0];
      if (NRM) out_nrm[0] = post_sqrt ? sqrt(red_sh[1]) : red_sh[1];
      *counter = 0u;
    }
  }
}

// out[0] = sum x  (used for the mean removal, solve.py:260-264)
__global__ void __launch_bounds__(kRedThreads) k_sum(const double* __restrict__ x, size_t len,
                                                     double* __restrict__ partial, unsigned int* counter,
                                                     double* __restrict__ out, RedCtx rc) {
  __shared__ double smem[kRedThreads / 32];
  __shared__ double red_sh[kRedMaxVals];
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) acc += x[i];
  const double s = block_sum(acc, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
  if (last_block(counter)) {
    double t = 0.0;
    for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) t += partial[bI];
    t = block_sum(t, smem);
    if (threadIdx.x == 0) red_sh[0] = t;
    __syncthreads();
    p2p_allreduce_sum(rc, red_sh, 1);
    if (threadIdx.x == 0) {
      out[0] = red_sh[0];
      *counter = 0u;
    }
  }
}

// weighted_L1 / weighted_L2 / max_norm of a-b (utils.py:7-17): out = {sum|q|, sum q^2, max|q|} (unweighted)
__global__ void __launch_bounds__(kRedThreads) k_diffnorms(const double* __restrict__ a, const double* __restrict__ b,
                                                           size_t len, double* __restrict__ partial,
                                                           unsigned int* counter, double* __restrict__ out) {
  __shared__ double smem[kRedThreads / 32];
  double s1 = 0.0, s2 = 0.0, mx = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const double q = fabs(a[i] - b[i]);
    s1 += q;
    s2 = fma(q, q, s2);
    mx = fmax(mx, q);
  }
  s1 = block_sum(s1, smem);
  s2 = block_sum(s2, smem);
  mx = block_max(mx, smem);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = s1;
    partial[3 * blockIdx.x + 1] = s2;
    partial[3 * blockIdx.x + 2] = mx;
  }
  if (last_block(counter)) {
    double t1 = 0.0, t2 = 0.0, tm = 0.0;
    for (int bI = threadIdx.x; bI < (int)gridDim.x; bI += blockDim.x) {
      t1 += partial[3 * bI];
      t2 += partial[3 * bI + 1];
      tm = fmax(tm, partial[3 * bI + 2]);
    }
    t1 = block_sum(t1, smem);
    t2 = block_sum(t2, smem);
    tm = block_max(tm, smem);
    if (threadIdx.x == 0) {
      out[0] = t1;
      out[1] = t2;
      out[2] = tm;
      *counter = 0u;
    }
  }
}

__global__ void k_sqrt_inplace(double* v, int nv) {
  const int i = threadIdx.x;
  if (i < nv) v[i] = sqrt(v[i]);
}

// y += sum_k alpha[k] * V_k  (alpha by value)
template <int NV>
__global__ void __launch_bounds__(256) k_multi_axpy(const double* __restrict__ V, size_t ld, Alphas al,
                                                    double* __restrict__ y, size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    double acc = y[i];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc = fma(al.a[k], V[k * ld + i], acc);
    y[i] = acc;
  }
}

// y += sign * (*alpha_dev) * x   (alpha stays on the device: no host round trip inside Gram-Schmidt)
__global__ void __launch_bounds__(256) k_axpy_dev(const double* __restrict__ alpha_dev, double sign,
                                                  const double* __restrict__ x, double* __restrict__ y, size_t len) {
  const double al = sign * alpha_dev[0];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) y[i] = fma(al, x[i], y[i]);
}

// y = x * s  with s = *s_dev or 1 / *s_dev (skipped when the scalar is 0: GMRES breakdown)
__global__ void __launch_bounds__(256) k_scale_dev(const double* s_dev, int reciprocal, const double* x, double* y,
                                                   size_t len) {
  double s = s_dev[0];
  if (reciprocal) s = (s != 0.0) ? 1.0 / s : 1.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) y[i] = x[i] * s;
}

// z = a*x + b*y (any of z may alias x or y)
__global__ void __launch_bounds__(256) k_axpby(double a, const double* x, double b, const double* y, double* z,
                                               size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) z[i] = a * x[i] + b * y[i];
}

// x -= (*sum_dev) * inv_count   (mean removal)
__global__ void __launch_bounds__(256) k_shift_dev(const double* __restrict__ sum_dev, double inv_count,
                                                   double* __restrict__ x, size_t len) {
  const double m = sum_dev[0] * inv_count;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) x[i] -= m;
}

// Chebyshev update: d = a*d + b*z ; x += d
__global__ void __launch_bounds__(256) k_cheb_update(double a, double b, const double* __restrict__ z,
                                                     double* __restrict__ d, double* __restrict__ x, size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const double dn = a * d[i] + b * z[i];
    d[i] = dn;
    x[i] += dn;
  }
}

// Chebyshev / iteration update as a stand-alone pass (only where no smoothing sweep can carry it as its epilogue:
// single-level hierarchies): d = ca*d + cb*z ; xk += d
__global__ void __launch_bounds__(256) k_cheb_ep(double ca, double cb, const double* __restrict__ z, double* __restrict__ d,
                                                 double* __restrict__ xk, int read_d, int read_x, int write_d, size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const double dn = ca * (read_d ? d[i] : 0.0) + cb * z[i];
    if (write_d) d[i] = dn;
    xk[i] = (read_x ? xk[i] : 0.0) + dn;
  }
}

// z = [w - y ; xp]: the last two lines of approx_schur_op (solve.py:275-276)
__global__ void __launch_bounds__(256) k_combine(const double* __restrict__ w, const double* __restrict__ y,
                                                 const double* __restrict__ xp, double* __restrict__ z, size_t n4, size_t n1) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4 + n1; i += stride)
    z[i] = (i < n4) ? 1.0 * w[i] + (-1.0) * y[i] : xp[i - n4];
}

}  // namespace mpbp
