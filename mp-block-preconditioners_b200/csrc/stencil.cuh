// Matrix-free MAC-grid stencil kernels (fp64, sm_100a) for the two-phase Stokes hot path.
//
// All operators are written in flux (stress-divergence) form, which is algebraically the dense
// matrices of preconditioner.py:86-349 ("flux form"):
//   Q = theta      * [(u[r,c+1]-u[r,c]) + (v[r+1,c]-v[r,c])]     at cell centres
//   T = theta_node * [(u[r-1,c]-u[r,c]) + (v[r,c]-v[r,c-1])]     at cell corners (top-left of (r,c))
//   (L u)_u = (Q[r,c]-Q[r,c-1]) + (T[r,c]-T[r+1,c])              preconditioner.py:127-179
//   (L u)_v = (T[r,c+1]-T[r,c]) + (Q[r,c]-Q[r-1,c])              preconditioner.py:242-295
//
// Thread mapping ("column marching"): one lane owns one grid column and walks down a strip of rows
// keeping a 3-row window in registers; horizontal neighbours are exchanged with warp shuffles.
// A warp spans 32 columns and produces 30 (lanes 1..30); lanes 0/31 are read-only halo lanes, so
// no shared memory and no block barrier is needed.  Rows above/below the slab come through the
// `top`/`bot` halo pointers (periodic wrap on one GPU, neighbour rows after a halo exchange).
#pragma once
#ifdef MPBP_EMU  // host-side logic checks of these kernels without a GPU (tests/emu/, test infrastructure only)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "ll.cuh"

namespace mpbp {

constexpr int kWarpCols = 30;   // output columns per warp
constexpr int kBlockWarps = 4;  // warps per block, side by side in x
constexpr int kBlockThreads = 32 * kBlockWarps;
constexpr unsigned kFull = 0xffffffffu;

struct Phys {
  double xi, c, d_u;
  double kap_n, kap_s;  // d_u*eta/h^2
  double dp_h;          // d_p/h
  double ddiv_h;        // d_div/h
  double inv_h;         // 1/h
  double dp_h2;         // d_p/h^2
  int mass_mode;        // 1: analytic separable face tables (level 0), 0: two-cell face average
  const double *sxf, *sxc, *syf, *syc;
};

// input vector view: field k at x + k*fs; halo rows (row -1 / row `rows`) at top + k*hs, bot + k*hs
struct VecIn {
  const double* x;
  const double* top;
  const double* bot;
  size_t fs, hs;
  // multi-GPU: the halo rows are pushed into this rank's comm buffer by the ring neighbours (peer stores over
  // NVLink) as (value, sequence tag) pairs.  The exchange sequence number lives on the device (`dseq`, bumped by the
  // pushing kernel), so kernel arguments never change between calls and whole applies replay as CUDA graphs.  The
  // threads of a consumer's edge strips poll the elements they need until the tag matches, land the values in the
  // plain per-level buffer `land` ([2][5][n]: top rows, then bottom rows) and read them from there: the host sets
  // top = land, bot = land + 5n, hs = n, so the view is never modified on the device (its fields stay kernel-parameter
  // constants on the uniform datapath).
  const unsigned long long* dseq;   // null: single GPU / replicated level (top/bot given directly)
  char* comm;                       // this rank's comm buffer
  size_t area;                      // elements per (slot, direction) halo area
  double* land;                     // landing buffer of the fetched rows
};

// ---- LL halo protocol ---------------------------------------------------------------------------------------------
// Every 8-byte value travels with its 8-byte sequence tag in ONE 16-byte store (the idea of NCCL's LL protocol): a
// consumer that reads a matching tag has the value, so the exchange needs no memory fence, no flag and no "last block"
// ticket on the critical path -- measured 3 us per back-to-back exchanging kernel against 10-13 us for data stores +
// __threadfence_system + release flag (profiles/micro/halo_latency.cu).
// Exchange number s (identical on all ranks: every rank runs the same kernel sequence) lives in slot s % 3 of the
// receiver's comm buffer, layout [slot][dir][5 fields * n0] elements.  Rules:
//  * consumer: a thread reading exchange s polls its elements until tag == s;
//  * producer (credit): before a block writes exchange q into neighbour X's slot q % 3 -- which still holds q-3 -- it
//    waits until SOME element of X's exchange q-1 has arrived here.  X's kernel producing q-1 has then started, so every
//    earlier kernel of X has completed, in particular the producer of q-2 and with it every reader of q-3 (readers of
//    q-3 are at or before the producer of q-2; the producer of q-1 may itself still be reading q-2, hence three slots).
//    X cannot overwrite q-1 with q+2 before it has seen this rank's q+1, so the tag is exact.
constexpr int kHaloSlots = 3;
__host__ __device__ __forceinline__ size_t comm_halo_bytes(size_t area) { return (size_t)kHaloSlots * 2 * area * sizeof(LLElem); }
__host__ __device__ __forceinline__ LLElem* comm_halo(char* base, size_t area, int slot, int dir) {
  return reinterpret_cast<LLElem*>(base) + (size_t)(slot * 2 + dir) * area;
}
// One thread fetches the halo elements of column `col` (fields k0, k0 + kstep, ... < k0 + nk * kstep; row length n) that
// IT will read and lands them in `land`; the same thread reads them back later (program order), so no barrier is
// involved.  WARP: called by all 32 lanes of a warp with warp-uniform arguments except `col` (warp-uniform wait loop).
template <bool WARP>
__device__ __forceinline__ void halo_fetch_cols(double* land, char* comm, size_t area, const unsigned long long* dseq, int k0,
                                              int nk, int kstep, int n, int col, bool need_top, bool need_bot) {
  const unsigned long long s = *dseq;
  const int slot = (int)(s % (unsigned long long)kHaloSlots);
  const LLElem* ll_top = comm_halo(comm, area, slot, 0);
  const LLElem* ll_bot = comm_halo(comm, area, slot, 1);
#pragma unroll 1
  for (int i = 0, k = k0; i < nk; ++i, k += kstep) {
    const int e = k * n + col;
    if (need_top) land[e] = WARP ? ll_wait_warp(ll_top + e, s) : ll_wait(ll_top + e, s);
    if (need_bot) land[e + 5 * n] = WARP ? ll_wait_warp(ll_bot + e, s) : ll_wait(ll_bot + e, s);
  }
}
// the light kernels (k_poisson, k_div, k_grad: 32 registers, 16 blocks/SM) fetch one or four fields of their own column.
// Inlined on purpose: an out-of-line call in a kernel takes its loop off the uniform datapath (measured on k_poisson<2>:
// 18 constant-bank loads per iteration instead of uniform registers, 130 instead of 105 us at 4096^2).
template <bool WARP = false>
__device__ __forceinline__ void halo_fetch(const VecIn& v, int k0, int nk, int kstep, int n, int col, bool need_top,
                                           bool need_bot) {
  if (v.dseq != nullptr) halo_fetch_cols<WARP>(v.land, v.comm, v.area, v.dseq, k0, nk, kstep, n, col, need_top, need_bot);
}
// producer credit (see above): lane 0 of a pushing warp, before its first remote store of exchange q
__device__ __forceinline__ void halo_credit(char* my_comm, size_t area, unsigned long long q, bool to_prev, bool to_next) {
  if (q < 2ull) return;
  const int slot = (int)((q - 1ull) % (unsigned long long)kHaloSlots);
  if (to_prev) ll_wait(comm_halo(my_comm, area, slot, 0), q - 1ull);
  if (to_next) ll_wait(comm_halo(my_comm, area, slot, 1), q - 1ull);
}

// Halo push: copies this rank's first / last slab row of every field into the ring neighbours' halo areas (peer
// memory over NVLink).  One thread per (field, column).  Used where the producer of a vector is not a stencil kernel.
__global__ void __launch_bounds__(256) k_halo_push(const double* __restrict__ x, int nf, size_t fs, int rows, int n,
                                                   char* prev_comm, char* next_comm, char* my_comm, size_t area,
                                                   unsigned long long* dseq, unsigned int* done_counter) {
  const unsigned long long seq = *dseq + 1ull;  // read before this block's ticket, written only by the last block
  const int slot = (int)(seq % (unsigned long long)kHaloSlots);
  if (threadIdx.x == 0) halo_credit(my_comm, area, seq, true, true);
  __syncthreads();
  LLElem* prev_bot = comm_halo(prev_comm, area, slot, 1);
  LLElem* next_top = comm_halo(next_comm, area, slot, 0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf * n) {
    const int k = i / n, c = i - k * n;
    st_ll(prev_bot + i, x[k * fs + c], seq);                          // my first row is the previous rank's row `rows`
    st_ll(next_top + i, x[k * fs + (size_t)(rows - 1) * n + c], seq);  // my last row is the next rank's row -1
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(done_counter, 1u);
    if (t == gridDim.x - 1) {
      *done_counter = 0u;
      *dseq = seq;
    }
  }
}

// consumer side of one halo exchange and nothing else (communication probe: exchange latency without a stencil)
__global__ void k_halo_consume(VecIn v, int nf, int n, double* sink) {
  double acc = 0.0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    halo_fetch(v, 0, nf, 1, n, c, true, true);
    acc += v.land[c];
  }
  if (sink && acc == 1.2345e300) sink[0] = acc;
}

// Chebyshev epilogue of the last smoothing sweep of a V-cycle: the sweep's result is the cycle's output z, which is
// consumed in registers instead of being stored:  d = ca*d + cb*z ;  xk += d.
struct ChebEp {
  double ca, cb;     // d = ca*d + cb*z
  double* d;         // Chebyshev direction (in/out)
  double* xk;        // iterate (in/out)
  int read_d;        // 0: first cycle (d = cb*z)
  int read_x;        // 0: first cycle (xk = d)
  int write_d;       // 0: last cycle (d is dead)
};

struct Geo {
  int n;     // columns (= global grid size of the level)
  int rows;  // local rows of the slab
  int row0;  // global index of local row 0
  int rs;    // rows per strip
  int pf;    // L2 prefetch distance in rows (0 = off)
  int re;    // > 0: the first and last strip are `re` rows short strips and the rows in between are cut into strips of
             // `rs` rows (single-wave decomposition, see strip_rows); 0: uniform strips of `rs` rows
  int pfc;   // 1: the prolongation + sweep variant also prefetches the coarse correction's rows into L2 (MPBP_PFC)
};
// number of strips (= gridDim.y) of a geometry
__host__ __device__ __forceinline__ int strip_count(const Geo& g) {
  if (g.re > 0) return 2 + (g.rows - 2 * g.re + g.rs - 1) / g.rs;
  return (g.rows + g.rs - 1) / g.rs;
}

// Software prefetch into L2 of the row `pf` rows ahead: three lanes (0, 16, 31) cover the <= 3 cache
// lines a warp's 32 columns touch.  Costs no registers, turns the later LDG into an L2 hit.
#ifdef MPBP_EMU
__device__ __forceinline__ void pf_l2(const double*) {}
#else
__device__ __forceinline__ void pf_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
__device__ __forceinline__ bool pf_lane() {
  const int lane = threadIdx.x & 31;
  return lane == 0 || lane == 16 || lane == 31;
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(kFull, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(kFull, v, 1); }

__device__ __forceinline__ const double* row_ptr(const VecIn& v, int k, int r, int rows, int n) {
  if (r < 0) return v.top + k * v.hs;
  if (r >= rows) return v.bot + k * v.hs;
  return v.x + k * v.fs + (size_t)r * n;
}

// theta is stored padded: row r (r in [-1, rows]) at th + (r+1)*n
__device__ __forceinline__ const double* th_row(const double* th, int r, int n) { return th + (size_t)(r + 1) * n; }

__device__ __forceinline__ double fast_rcp(double d) {
#ifdef MPBP_EMU
  return 1.0 / d;
#endif
  double r;
#ifndef MPBP_EMU
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#endif
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

struct LaneGeom {
  int cc;       // wrapped column this lane reads
  int j;        // unwrapped output column
  bool store;   // lane writes output
  bool alive;   // warp has at least one output column
};

__device__ __forceinline__ LaneGeom lane_geom(int n) {
  LaneGeom g;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int j0 = (blockIdx.x * kBlockWarps + warp) * kWarpCols;
  g.alive = j0 < n;
  g.j = j0 + lane - 1;
  int cc = g.j % n;
  if (cc < 0) cc += n;
  g.cc = cc;
  g.store = (lane >= 1) && (lane <= kWarpCols) && (g.j < n);
  return g;
}

// Fused halo push (MPBP_PUSH_FUSED): the producing stencil kernel itself stores its first and last output rows into
// the ring neighbours' comm buffers (peer memory over NVLink, LL elements) -- the compute kernel IS the halo exchange,
// and because the edge strips are scheduled first the transfer overlaps the interior of the slab.
struct PushOut {
  char* prev_comm;            // neighbours' comm buffers (mapped peer memory)
  char* next_comm;
  char* my_comm;              // this rank's comm buffer (credit)
  size_t area;                // elements per (slot, direction) halo area
  unsigned long long* dseq;   // this rank's exchange counter: output rows go out under sequence *dseq + 1
  unsigned int* counters;     // [2] edge blocks done
};

// A producing kernel's fused push.  Blocks that own the slab's first row (`first`) write it into the previous
// rank's bottom halo area, blocks that own the last row (`last`) into the next rank's top area, both tagged with
// sequence *dseq + 1; the last edge block of the launch bumps *dseq.  push_begin: every warp of an edge block (credit
// wait before the first remote store); push_end: every non-exited thread of an edge block.
struct PushCtx {
  LLElem* prev;  // neighbour's bot area: receives my row 0
  LLElem* next;  // neighbour's top area: receives my row rows-1
  unsigned long long seq;
};
__device__ __forceinline__ PushCtx push_begin(const PushOut& po, bool first, bool last) {
  PushCtx c;
  c.seq = *po.dseq + 1ull;  // bumped only after every edge block of this launch has finished
  const int slot = (int)(c.seq % (unsigned long long)kHaloSlots);
  c.prev = comm_halo(po.prev_comm, po.area, slot, 1);
  c.next = comm_halo(po.next_comm, po.area, slot, 0);
  if (first || last) {
    if ((threadIdx.x & 31) == 0) halo_credit(po.my_comm, po.area, c.seq, first, last);
    __syncwarp();
  }
  return c;
}
// n_first / n_last: number of blocks of this launch that own the first / last row
__device__ __forceinline__ void push_end(const PushOut& po, const PushCtx& c, bool first, bool last, unsigned int n_first,
                                         unsigned int n_last, bool same_blocks) {
  if (!(first || last)) return;
  __syncthreads();  // every warp of the block has read *dseq
  if (threadIdx.x == 0) {
    const unsigned int total = same_blocks ? n_first : n_first + n_last;
    if (atomicAdd(&po.counters[2], 1u) == total - 1) {
      po.counters[2] = 0u;
      *po.dseq = c.seq;
    }
  }
}
// strip order of the marching kernels: first strip, LAST strip, then the interior -- both edge strips (the ones that
// wait for the ring neighbours' rows and push this rank's own) run in the first wave, the interior hides the transfer
__device__ __forceinline__ int strip_of_block() {
  const int S = (int)gridDim.y, sy = (int)blockIdx.y;
  return (S > 2) ? (sy == 0 ? 0 : (sy == 1 ? S - 1 : sy - 1)) : sy;
}
// Rows [r0, r1) of this block's strip.  Large levels use the SINGLE-WAVE decomposition: the number of interior strips
// is chosen so that all blocks of the launch are resident at once (no tail wave: with uniform 32-row strips a
// 4096^2 sweep was 6.05 waves of 5 blocks/SM and paid for 7), and two SHORT edge strips, scheduled first, carry the
// halo waits / pushes and the general (periodic / halo) code path, so their slots are recycled almost immediately.
__device__ __forceinline__ bool strip_rows(const Geo& g, int& r0, int& r1) {
  const int s = strip_of_block();
  if (g.re > 0) {
    const int S = (int)gridDim.y;
    if (s == 0) { r0 = 0; r1 = g.re; }
    else if (s == S - 1) { r0 = g.rows - g.re; r1 = g.rows; }
    else { r0 = g.re + (s - 1) * g.rs; r1 = min(r0 + g.rs, g.rows - g.re); }
  } else {
    r0 = s * g.rs;
    r1 = min(r0 + g.rs, g.rows);
  }
  return r0 < r1;
}

// (the velocity-block / full-system marching kernel with its fused variants lives in stokes.cuh)

__global__ void k_fill(double* __restrict__ x, double v, size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) x[i] = v;
}

// x = omega * b / diag(F): first Jacobi sweep from a zero initial guess (solve.py:149-159 with x=0)
__global__ void __launch_bounds__(kBlockThreads) k_jacobi0_F(const double* __restrict__ th, const double* __restrict__ b,
                                                             double* __restrict__ y, size_t fs, Geo g, Phys ph,
                                                             double omega) {
  const LaneGeom lg = lane_geom(g.n);
  if (!lg.alive) return;
  const int n = g.n, c = lg.cc;
  int r0, r1;
  if (!strip_rows(g, r0, r1)) return;
  double sxf = 0.0, sxc = 0.0;
  if (ph.mass_mode) {
    sxf = ph.sxf[c];
    sxc = ph.sxc[c];
  }
  double th_m = th_row(th, r0 - 1, n)[c];
  double th_c = th_row(th, r0, n)[c];
  const double a_m = th_m + shfl_up1(th_m);
  double a_c = th_c + shfl_up1(th_c);
  double node_c = 0.25 * (a_c + a_m);
  const bool pfl = pf_lane();
  for (int r = r0; r < r1; ++r) {
    if (g.pf > 0 && pfl) {
      const int rp = r + g.pf;
      if (rp < r1) {
        pf_l2(th_row(th, rp + 1, n) + c);
        const size_t offp = (size_t)rp * n + c;
#pragma unroll
        for (int k = 0; k < 4; ++k) pf_l2(b + offp + k * fs);
      }
    }
    const double th_p = th_row(th, r + 1, n)[c];
    const double a_p = th_p + shfl_up1(th_p);
    const double node_p = 0.25 * (a_p + a_c);
    const double node_e = shfl_dn1(node_c);
    const double fu_c = 0.5 * a_c, fv_c = 0.5 * (th_c + th_m);
    double mu, mv;
    if (ph.mass_mode) {
      const int gr = g.row0 + r;
      mu = 0.25 * sxf * ph.syc[gr] + 0.5;
      mv = 0.25 * sxc * ph.syf[gr] + 0.5;
    } else {
      mu = fu_c;
      mv = fv_c;
    }
    const double dXu = ph.d_u * (ph.xi * fu_c * (1.0 - fu_c));
    const double dXv = ph.d_u * (ph.xi * fv_c * (1.0 - fv_c));
    const double cmu = ph.c * mu, cmv = ph.c * mv;
    const double su = a_c + node_c + node_p;
    const double sv = th_m + th_c + node_c + node_e;
    const double d_un = cmu - dXu - ph.kap_n * su;
    const double d_us = (ph.c - cmu) - dXu - ph.kap_s * (4.0 - su);
    const double d_vn = cmv - dXv - ph.kap_n * sv;
    const double d_vs = (ph.c - cmv) - dXv - ph.kap_s * (4.0 - sv);
    if (lg.store) {
      const size_t off = (size_t)r * n + c;
      y[off] = omega * b[off] * fast_rcp(d_un);
      y[off + fs] = omega * b[off + fs] * fast_rcp(d_vn);
      y[off + 2 * fs] = omega * b[off + 2 * fs] * fast_rcp(d_us);
      y[off + 3 * fs] = omega * b[off + 3 * fs] * fast_rcp(d_vs);
    }
    th_m = th_c; th_c = th_p; a_c = a_p; node_c = node_p;
  }
}

// ------------------------------------------------------------------------------------------
// Pressure-Poisson operator GtG = -D G (solve.py:246-247): 5-point, face weights wu = fu_n^2+fu_s^2.
//   MODE 0: y = GtG p ; 1: y = b - GtG p ; 2: y = p + omega (b - GtG p)/diag ; 3: y = omega b/diag
// ------------------------------------------------------------------------------------------
//   CHEB (MODE 2): the sweep's result z goes through the Chebyshev epilogue (ChebEp) instead of being stored
//   PUSH: the result's first / last rows also go to the ring neighbours (see PushOut)
//   (register caps: 32 for the plain variants = 16 resident blocks per SM, 48 / 56 / 64 with the Chebyshev epilogue /
//   the fused push / both: these kernels consume their loads in the iteration that issues them and live off occupancy)
template <int MODE, bool CHEB, bool PUSH, bool FETCH>
__device__ __forceinline__ void poisson_march(const VecIn& pin_, const double* __restrict__ th,
                                              const double* __restrict__ b, double* __restrict__ y, const Geo& g,
                                              const Phys& ph, double omega, const ChebEp& ce, const PushOut& po,
                                              const LaneGeom& lg, int r0, int r1) {
  const int n = g.n, rows = g.rows, c = lg.cc;
  // A local copy of the view whose halo pointers are (re)assigned under a uniform condition -- semantically a no-op
  // (make_view already points top/bot at the landing buffer) but it makes the compiler keep the view in uniform registers:
  // reading it straight from the parameter bank cost the loop 16 constant loads per iteration and, under the 32-register
  // cap, spills (k_poisson<2> at 4096^2: 130 us instead of 105).
  VecIn pin = pin_;
  if (pin.dseq != nullptr) {
    pin.top = pin.land;
    pin.bot = pin.land + (size_t)5 * n;
  }
  if (FETCH) halo_fetch<true>(pin, 0, 1, 1, n, c, r0 == 0, r1 == rows);
  PushCtx pc{};
  if (PUSH) pc = push_begin(po, r0 == 0, r1 == rows);
  double th_m = th_row(th, r0 - 1, n)[c];
  double th_c = th_row(th, r0, n)[c];
  double p_m = 0.0, p_c = 0.0;
  if (MODE != 3) {
    p_m = row_ptr(pin, 0, r0 - 1, rows, n)[c];
    p_c = row_ptr(pin, 0, r0, rows, n)[c];
  }
  double fv = 0.5 * (th_c + th_m);
  double wv_c = fv * fv + (1.0 - fv) * (1.0 - fv);
  double Hy_c = wv_c * (p_m - p_c);
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    const double th_p = th_row(th, r + 1, n)[c];
    double p_p = 0.0;
    if (MODE != 3) p_p = row_ptr(pin, 0, r + 1, rows, n)[c];
    const double fu = 0.5 * (th_c + shfl_up1(th_c));
    const double wu = fu * fu + (1.0 - fu) * (1.0 - fu);
    const double Hx = wu * (p_c - shfl_up1(p_c));
    fv = 0.5 * (th_p + th_c);
    const double wv_p = fv * fv + (1.0 - fv) * (1.0 - fv);
    const double Hy_p = wv_p * (p_c - p_p);
    double out = -ph.dp_h2 * ((shfl_dn1(Hx) - Hx) + (Hy_c - Hy_p));
    const size_t off = (size_t)r * n + c;
    if (MODE == 1) out = b[off] - out;
    if (MODE == 2 || MODE == 3) {
      const double dg = ph.dp_h2 * (shfl_dn1(wu) + wu + wv_c + wv_p);
      const double rinv = fast_rcp(dg);
      if (MODE == 2) out = p_c + omega * (b[off] - out) * rinv;
      else out = omega * b[off] * rinv;
    }
    if (lg.store) {
      if (CHEB) {
        const double dk = ce.read_d ? ce.d[off] : 0.0;
        const double dn = ce.ca * dk + ce.cb * out;
        if (ce.write_d) ce.d[off] = dn;
        out = (ce.read_x ? ce.xk[off] : 0.0) + dn;
        ce.xk[off] = out;
      } else {
        y[off] = out;
      }
      if (PUSH) {
        if (r == 0) st_ll(pc.prev + c, out, pc.seq);
        if (r == rows - 1) st_ll(pc.next + c, out, pc.seq);
      }
    }
    th_m = th_c; th_c = th_p; p_m = p_c; p_c = p_p; wv_c = wv_p; Hy_c = Hy_p;
  }
  if (PUSH) push_end(po, pc, r0 == 0, r1 == rows, gridDim.x, gridDim.x, gridDim.y == 1);
}
// FETCH (slab edge strips of a distributed level) is a separate instantiation: the polling code in front of the loop
// takes it off the uniform datapath (18 constant-bank loads per iteration), which the interior strips and every
// single-GPU launch must not pay for.
template <int MODE, bool CHEB = false, bool PUSH = false>
__global__ void __launch_bounds__(kBlockThreads, (CHEB && PUSH) ? 8 : (PUSH ? 9 : (CHEB ? 10 : 16)))
k_poisson(VecIn pin, const double* __restrict__ th, const double* __restrict__ b, double* __restrict__ y, Geo g, Phys ph,
          double omega, ChebEp ce = ChebEp{}, PushOut po = PushOut{}) {
  const LaneGeom lg = lane_geom(g.n);
  if (!lg.alive) return;
  int r0, r1;
  if (!strip_rows(g, r0, r1)) return;
  if (MODE != 3 && pin.dseq != nullptr && (r0 == 0 || r1 == g.rows))
    poisson_march<MODE, CHEB, PUSH, true>(pin, th, b, y, g, ph, omega, ce, po, lg, r0, r1);
  else
    poisson_march<MODE, CHEB, PUSH, false>(pin, th, b, y, g, ph, omega, ce, po, lg, r0, r1);
}

// r = D w (+ add): un-negated divergence of both phases (preconditioner.py:221-238, :311; solve.py:259)
__global__ void __launch_bounds__(kBlockThreads) k_div(VecIn win, const double* __restrict__ th,
                                                       const double* __restrict__ add, double* __restrict__ y, Geo g,
                                                       Phys ph) {
  const LaneGeom lg = lane_geom(g.n);
  if (!lg.alive) return;
  const int n = g.n, rows = g.rows, c = lg.cc;
  int r0, r1;
  if (!strip_rows(g, r0, r1)) return;
  if (r0 == 0 || r1 == rows) {
    halo_fetch<true>(win, 0, 4, 1, n, c, r0 == 0, r1 == rows);
  }
  double th_m = th_row(th, r0 - 1, n)[c];
  double th_c = th_row(th, r0, n)[c];
  double vn_c = row_ptr(win, 1, r0, rows, n)[c], vs_c = row_ptr(win, 3, r0, rows, n)[c];
  double fv = 0.5 * (th_c + th_m);
  double Vsum_c = vs_c + fv * (vn_c - vs_c);
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    const double th_p = th_row(th, r + 1, n)[c];
    const double vn_p = row_ptr(win, 1, r + 1, rows, n)[c], vs_p = row_ptr(win, 3, r + 1, rows, n)[c];
    const double un_c = row_ptr(win, 0, r, rows, n)[c], us_c = row_ptr(win, 2, r, rows, n)[c];
    const double fu = 0.5 * (th_c + shfl_up1(th_c));
    const double Usum = us_c + fu * (un_c - us_c);
    fv = 0.5 * (th_p + th_c);
    const double Vsum_p = vs_p + fv * (vn_p - vs_p);
    double out = ph.inv_h * ((shfl_dn1(Usum) - Usum) + (Vsum_c - Vsum_p));
    const size_t off = (size_t)r * n + c;
    if (add != nullptr) out += add[off];
    if (lg.store) y[off] = out;
    th_m = th_c; th_c = th_p; Vsum_c = Vsum_p;
  }
}

// y = G p for both phases (preconditioner.py:203-219, :313; solve.py:273)
template <bool PUSH = false>
__global__ void __launch_bounds__(kBlockThreads) k_grad(VecIn pin, const double* __restrict__ th,
                                                        double* __restrict__ y, size_t fs, Geo g, Phys ph,
                                                        PushOut po = PushOut{}) {
  const LaneGeom lg = lane_geom(g.n);
  if (!lg.alive) return;
  const int n = g.n, rows = g.rows, c = lg.cc;
  int r0, r1;
  if (!strip_rows(g, r0, r1)) return;
  if (r0 == 0 || r1 == rows) {
    halo_fetch<true>(pin, 0, 1, 1, n, c, r0 == 0, r1 == rows);
  }
  PushCtx pc{};
  if (PUSH) pc = push_begin(po, r0 == 0, r1 == rows);
  double th_m = th_row(th, r0 - 1, n)[c];
  double p_m = row_ptr(pin, 0, r0 - 1, rows, n)[c];
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    const double th_c = th_row(th, r, n)[c];
    const double p_c = row_ptr(pin, 0, r, rows, n)[c];
    const double fu = 0.5 * (th_c + shfl_up1(th_c));
    const double fv = 0.5 * (th_c + th_m);
    const double gx = ph.dp_h * (p_c - shfl_up1(p_c));
    const double gy = ph.dp_h * (p_m - p_c);
    if (lg.store) {
      const size_t off = (size_t)r * n + c;
      const double g0 = fu * gx, g1 = fv * gy, g2 = (1.0 - fu) * gx, g3 = (1.0 - fv) * gy;
      y[off] = g0;
      y[off + fs] = g1;
      y[off + 2 * fs] = g2;
      y[off + 3 * fs] = g3;
      if (PUSH) {
        if (r == 0) {
          st_ll(pc.prev + c, g0, pc.seq); st_ll(pc.prev + n + c, g1, pc.seq);
          st_ll(pc.prev + 2 * n + c, g2, pc.seq); st_ll(pc.prev + 3 * n + c, g3, pc.seq);
        }
        if (r == rows - 1) {
          st_ll(pc.next + c, g0, pc.seq); st_ll(pc.next + n + c, g1, pc.seq);
          st_ll(pc.next + 2 * n + c, g2, pc.seq); st_ll(pc.next + 3 * n + c, g3, pc.seq);
        }
      }
    }
    th_m = th_c; p_m = p_c;
  }
  if (PUSH) push_end(po, pc, r0 == 0, r1 == rows, gridDim.x, gridDim.x, gridDim.y == 1);
}

// ------------------------------------------------------------------------------------------
// grid transfers (thread per output cell; coarse cell (R,C) covers fine (2R..2R+1, 2C..2C+1))
// ------------------------------------------------------------------------------------------
// full weighting of the four face fields: u: (1/4,1/2,1/4) over columns x (1/2,1/2) over rows; v transposed
template <bool PUSH = false>
__global__ void k_restrict_F(VecIn fin, double* __restrict__ yc, int nf, int rows_f, PushOut po = PushOut{}) {
  const int nc = nf >> 1, rows_c = rows_f >> 1;
  const int C = blockIdx.x * blockDim.x + threadIdx.x;
  const int R = blockIdx.y;
  if (R == 0 && C < nc) {
    // only the v-type fields of fine row -1 are read (columns 2C, 2C+1)
    halo_fetch(fin, 1, 2, 2, nf, 2 * C, true, false);
    halo_fetch(fin, 1, 2, 2, nf, 2 * C + 1, true, false);
  }
  PushCtx pc{};
  if (PUSH) pc = push_begin(po, R == 0, R == rows_c - 1);
  if (C < nc && R < rows_c) {
  const size_t fsc = (size_t)rows_c * nc;
  const int c0 = 2 * C, cm = (c0 == 0) ? nf - 1 : c0 - 1, cp = c0 + 1;
  const int ra = 2 * R, rb = 2 * R + 1;
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
    const double* ua = row_ptr(fin, 2 * ph, ra, rows_f, nf);
    const double* ub = row_ptr(fin, 2 * ph, rb, rows_f, nf);
    const double um = 0.5 * (ua[cm] + ub[cm]), u0 = 0.5 * (ua[c0] + ub[c0]), up = 0.5 * (ua[cp] + ub[cp]);
    const double cu = 0.25 * um + 0.5 * u0 + 0.25 * up;
    yc[(2 * ph) * fsc + (size_t)R * nc + C] = cu;
    const double* vm = row_ptr(fin, 2 * ph + 1, ra - 1, rows_f, nf);
    const double* v0 = row_ptr(fin, 2 * ph + 1, ra, rows_f, nf);
    const double* vp = row_ptr(fin, 2 * ph + 1, rb, rows_f, nf);
    const double wm = 0.5 * (vm[c0] + vm[cp]), w0 = 0.5 * (v0[c0] + v0[cp]), wp = 0.5 * (vp[c0] + vp[cp]);
    const double cv = 0.25 * wm + 0.5 * w0 + 0.25 * wp;
    yc[(2 * ph + 1) * fsc + (size_t)R * nc + C] = cv;
    if (PUSH) {
      if (R == 0) { st_ll(pc.prev + (2 * ph) * nc + C, cu, pc.seq); st_ll(pc.prev + (2 * ph + 1) * nc + C, cv, pc.seq); }
      if (R == rows_c - 1) { st_ll(pc.next + (2 * ph) * nc + C, cu, pc.seq); st_ll(pc.next + (2 * ph + 1) * nc + C, cv, pc.seq); }
    }
  }
  }
  if (PUSH) push_end(po, pc, R == 0, R == rows_c - 1, gridDim.x, gridDim.x, gridDim.y == 1);
}

// x_f += P x_c, P = 4 R^T: u linear in x / constant in y, v linear in y / constant in x
__global__ void k_prolong_add_F(VecIn cin, double* __restrict__ xf, int nf, int rows_f) {
  const int nc = nf >> 1, rows_c = rows_f >> 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= nf || r >= rows_f) return;
  const size_t fsf = (size_t)rows_f * nf;
  const int C = c >> 1, R = r >> 1;
  const int Cp = (C + 1 == nc) ? 0 : C + 1;
  if (r == rows_f - 1) {
    // only the v-type fields of coarse row rows_c are read (column C)
    halo_fetch(cin, 1, 2, 2, nc, C, false, true);
  }
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
    const double* uc = row_ptr(cin, 2 * ph, R, rows_c, nc);
    const double eu = (c & 1) ? 0.5 * (uc[C] + uc[Cp]) : uc[C];
    xf[(2 * ph) * fsf + (size_t)r * nf + c] += eu;
    const double* v0 = row_ptr(cin, 2 * ph + 1, R, rows_c, nc);
    double ev = v0[C];
    if (r & 1) ev = 0.5 * (ev + row_ptr(cin, 2 * ph + 1, R + 1, rows_c, nc)[C]);
    xf[(2 * ph + 1) * fsf + (size_t)r * nf + c] += ev;
  }
}

// 4-cell average of a cell-centred field
__global__ void k_restrict_P(const double* __restrict__ f, double* __restrict__ yc, int nf, int rows_f) {
  const int nc = nf >> 1, rows_c = rows_f >> 1;
  const int C = blockIdx.x * blockDim.x + threadIdx.x;
  const int R = blockIdx.y;
  if (C >= nc || R >= rows_c) return;
  const double* a = f + (size_t)(2 * R) * nf + 2 * C;
  const double* bq = a + nf;
  yc[(size_t)R * nc + C] = 0.25 * ((a[0] + a[1]) + (bq[0] + bq[1]));
}

// x_f += piecewise-constant prolongation of x_c
template <bool PUSH = false>
__global__ void k_prolong_add_P(const double* __restrict__ xc, double* __restrict__ xf, int nf, int rows_f,
                                PushOut po = PushOut{}) {
  const int nc = nf >> 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  PushCtx pc{};
  if (PUSH) pc = push_begin(po, r == 0, r == rows_f - 1);
  if (c < nf && r < rows_f) {
    const double v = xf[(size_t)r * nf + c] + xc[(size_t)(r >> 1) * nc + (c >> 1)];
    xf[(size_t)r * nf + c] = v;
    if (PUSH) {
      if (r == 0) st_ll(pc.prev + c, v, pc.seq);
      if (r == rows_f - 1) st_ll(pc.next + c, v, pc.seq);
    }
  }
  if (PUSH) push_end(po, pc, r == 0, r == rows_f - 1, gridDim.x, gridDim.x, gridDim.y == 1);
}

#ifndef MPBP_EMU
// y = M x for the coarsest-level dense (pseudo-)inverse, M row-major: one warp per row (coalesced row read, fixed
// summation order), m/8 blocks -- at n_coarse = 16 (m = 1024, 8 MB, L2-resident) one launch of a few microseconds
// replaces three multigrid levels of latency-bound kernels
__global__ void __launch_bounds__(256) k_dense_matvec(const double* __restrict__ M, const double* __restrict__ x,
                                                      double* __restrict__ y, int m) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= m) return;
  const double* __restrict__ mr = M + (size_t)row * m;
  double acc = 0.0;
  if ((m & 63) == 0 && ((reinterpret_cast<uintptr_t>(M) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
    // 16-byte loads, eight of them in flight per lane: the matrix is usually NOT L2-resident (a V-cycle streams
    // gigabytes between two coarse solves), so the kernel is one DRAM latency plus 8 MB / all SMs
    const double2* __restrict__ m2 = reinterpret_cast<const double2*>(mr);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
    const int h = m >> 1;  // double2 elements per row
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int k = lane; k < h; k += 256) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (k + 32 * u < h) ? m2[k + 32 * u] : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        if (k + 32 * u < h) {
          const double2 xa = x2[k + 32 * u];
          a0 = fma(v[u].x, xa.x, a0);
          a1 = fma(v[u].y, xa.y, a1);
        }
        if (k + 32 * (u + 1) < h) {
          const double2 xb = x2[k + 32 * (u + 1)];
          a2 = fma(v[u + 1].x, xb.x, a2);
          a3 = fma(v[u + 1].y, xb.y, a3);
        }
      }
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (int k = lane; k < m; k += 32) acc = fma(mr[k], x[k], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if (lane == 0) y[row] = acc;
}
#endif  // MPBP_EMU

// manufactured solution and right-hand side of solve.main (solve.py:52-78 via utils.py:159-210)
__global__ void k_fill_manufactured(double* __restrict__ u_vec, double* __restrict__ b_vec, int n, int rows, int row0,
                                    double c, double d, double xi, double etan, double etas, double b_p_sign) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (col >= n || r >= rows) return;
  const double PI = 3.141592653589793;
  const double h = 1.0 / n, nu = 1.0;
  const int gr = row0 + r;
  const size_t fs = (size_t)rows * n, off = (size_t)r * n + col;
  const double yu = -(gr + 0.5) * h, xu = col * h;          // utils.py:187
  const double yv = -gr * h, xv = (col + 0.5) * h;          // utils.py:188
  const double yp = -(gr + 0.5) * h, xp = (col + 0.5) * h;  // utils.py:193
  {
    const double sx = sin(2 * PI * xu), cy = cos(2 * PI * yu), sy = sin(2 * PI * yu);
    const double ux = sx * cy;  // solve.py:52
    const double common = 2 * nu * sx * sy;
    const double s2 = sx * sx * sy * sy;
    const double bn = (cy * sx * (4 * c * nu - 4 * d * (8 * etan * nu * PI * PI + xi) +
                                  (c - 16 * d * etan * PI * PI) * common + d * xi * s2)) / (8 * nu);  // solve.py:72
    const double bs = (cy * sx * (-4 * c * nu + 4 * d * (8 * etas * nu * PI * PI + xi) +
                                  (c - 16 * d * etas * PI * PI) * common - d * xi * s2)) / (8 * nu);  // solve.py:75
    if (u_vec) { u_vec[off] = ux; u_vec[off + 2 * fs] = -ux; }
    if (b_vec) { b_vec[off] = bn; b_vec[off + 2 * fs] = bs; }
  }
  {
    const double cx = cos(2 * PI * xv), sx = sin(2 * PI * xv), sy = sin(2 * PI * yv);
    const double uy = cx * sy;  // solve.py:53
    const double common = 2 * nu * sx * sy;
    const double s2 = sx * sx * sy * sy;
    const double bn = (cx * sy * (4 * c * nu - 4 * d * (8 * etan * nu * PI * PI + xi) +
                                  (c - 16 * d * etan * PI * PI) * common + d * xi * s2)) / (8 * nu);  // solve.py:73
    const double bs = (cx * sy * (-4 * c * nu + 4 * d * (8 * etas * nu * PI * PI + xi) +
                                  (c - 16 * d * etas * PI * PI) * common - d * xi * s2)) / (8 * nu);  // solve.py:76
    if (u_vec) { u_vec[off + fs] = uy; u_vec[off + 3 * fs] = -uy; }
    if (b_vec) { b_vec[off + fs] = bn; b_vec[off + 3 * fs] = bs; }
  }
  if (u_vec) u_vec[off + 4 * fs] = 0.0;                                                    // solve.py:58
  if (b_vec) b_vec[off + 4 * fs] = b_p_sign * PI * sin(4 * PI * xp) * sin(4 * PI * yp);  // solve.py:78
}

}  // namespace mpbp
