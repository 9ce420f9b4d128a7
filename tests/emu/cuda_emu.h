// cuda_emu.h -- a minimal SIMT-on-CPU shim so the CUDA kernels of csrc/*.cuh can be compiled with g++ and
// executed on the build container (which has no GPU) for LOGIC checks against the oracle.
//
// TEST INFRASTRUCTURE ONLY.  Every CUDA thread of a block runs as one std::thread; warp shuffles,
// __syncwarp and __syncthreads are barrier exchanges, so divergence-free warp-synchronous code (all the
// kernels here) behaves as on the device.  Blocks run one after another.  It says nothing about speed and
// nothing about memory-model subtleties; GPU parity tests remain the gate for the product.
#pragma once
#ifndef MPBP_EMU
#error "cuda_emu.h is only for -DMPBP_EMU host builds"
#endif

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static /* blocks run one after another, so a function-local static is block-shared */

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3() = default;
  dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef void* cudaStream_t;

namespace emu {

// reusable barrier that tolerates participants leaving (a CUDA thread returning early)
class Barrier {
 public:
  explicit Barrier(int n) : count_(n), waiting_(0), gen_(0) {}
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned long g = gen_;
    if (++waiting_ == count_) {
      waiting_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return g != gen_; });
    }
  }
  void drop() {
    std::unique_lock<std::mutex> lk(m_);
    --count_;
    if (count_ > 0 && waiting_ == count_) {
      waiting_ = 0;
      ++gen_;
      cv_.notify_all();
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int count_, waiting_;
  unsigned long gen_;
};

struct WarpCtx {
  Barrier bar{32};
  double xd[32];
  explicit WarpCtx(int lanes) : bar(lanes) {}
};
struct BlockCtx {
  Barrier bar;
  std::vector<WarpCtx*> warps;
  explicit BlockCtx(int threads) : bar(threads) {}
};

struct ThreadState {
  dim3 tid, bid, bdim, gdim;
  WarpCtx* warp = nullptr;
  BlockCtx* block = nullptr;
};
inline ThreadState& ts() {
  static thread_local ThreadState s;
  return s;
}

template <class F>
void run_block(dim3 grid, dim3 block, unsigned bx, unsigned by, F& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  BlockCtx bc(nthreads);
  const int nwarps = (nthreads + 31) / 32;
  for (int w = 0; w < nwarps; ++w) bc.warps.push_back(new WarpCtx(std::min(32, nthreads - 32 * w)));
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    th.emplace_back([&, t]() {
      ThreadState& s = ts();
      s.tid = dim3((unsigned)t);
      s.bid = dim3(bx, by);
      s.bdim = block;
      s.gdim = grid;
      s.block = &bc;
      s.warp = bc.warps[t / 32];
      body();
      s.warp->bar.drop();
      s.block->bar.drop();
    });
  }
  for (auto& x : th) x.join();
  for (auto* w : bc.warps) delete w;
}

template <class F>
void launch(dim3 grid, dim3 block, F body) {
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) run_block(grid, block, bx, by, body);
}

// All blocks of the launch run at the same time (one host thread per block): for kernels whose blocks wait for each
// other or for another rank's blocks (small grids only; kernels with __shared__ statics cannot use it).
template <class F>
void launch_concurrent(dim3 grid, dim3 block, F body) {
  std::vector<std::thread> blocks;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) blocks.emplace_back([&, bx, by] { run_block(grid, block, bx, by, body); });
  for (auto& b : blocks) b.join();
}

}  // namespace emu

#define threadIdx (emu::ts().tid)
#define blockIdx (emu::ts().bid)
#define blockDim (emu::ts().bdim)
#define gridDim (emu::ts().gdim)

inline double __shfl_up_sync(unsigned, double v, int d) {
  emu::WarpCtx* w = emu::ts().warp;
  const int lane = (int)(emu::ts().tid.x & 31);
  w->xd[lane] = v;
  w->bar.arrive_and_wait();
  const double r = lane >= d ? w->xd[lane - d] : v;
  w->bar.arrive_and_wait();
  return r;
}
inline double __shfl_down_sync(unsigned, double v, int d) {
  emu::WarpCtx* w = emu::ts().warp;
  const int lane = (int)(emu::ts().tid.x & 31);
  w->xd[lane] = v;
  w->bar.arrive_and_wait();
  const double r = lane + d < 32 ? w->xd[lane + d] : v;
  w->bar.arrive_and_wait();
  return r;
}
inline double __shfl_xor_sync(unsigned, double v, int m) {
  emu::WarpCtx* w = emu::ts().warp;
  const int lane = (int)(emu::ts().tid.x & 31);
  w->xd[lane] = v;
  w->bar.arrive_and_wait();
  const double r = w->xd[lane ^ m];
  w->bar.arrive_and_wait();
  return r;
}
inline int __shfl_sync(unsigned, int v, int src) {
  emu::WarpCtx* w = emu::ts().warp;
  const int lane = (int)(emu::ts().tid.x & 31);
  w->xd[lane] = (double)v;
  w->bar.arrive_and_wait();
  const int r = (int)w->xd[src & 31];
  w->bar.arrive_and_wait();
  return r;
}
inline void __syncwarp() { emu::ts().warp->bar.arrive_and_wait(); }
inline void __syncthreads() { emu::ts().block->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __sync_fetch_and_add(p, v); }
using std::fma;
using std::fmax;
using std::min;
using std::sqrt;
