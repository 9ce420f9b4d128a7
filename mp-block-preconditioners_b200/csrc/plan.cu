// libmpbp.so -- host side of the C ABI declared in include/mpbp.h: plan (multigrid hierarchy,
// workspace, slab decomposition), sub-solvers, the block preconditioner apply (solve.py:257-277)
// and the Krylov drivers (scipy gmres / pyamg fgmres semantics).  All arithmetic on vectors runs in
// the CUDA kernels of stencil.cuh / blas1.cuh; there is no CPU fallback.
#include "../../include/mpbp.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "blas1.cuh"
#include "coarse.cuh"
#include "stencil.cuh"
#include "stokes.cuh"
#include "cell.cuh"
#include "poisson.cuh"

using namespace mpbp;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return set_err((int)e_ > 0 ? (int)e_ : 999, "%s:%d CUDA error %d: %s", __FILE__, __LINE__, (int)e_, \
                     cudaGetErrorString(e_));                                                            \
  } while (0)
#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t e_ = (call);                                                                                 \
    if (e_ != ncclSuccess)                                                                                    \
      return set_err(1000 + (int)e_, "%s:%d NCCL error %d: %s", __FILE__, __LINE__, (int)e_, ncclGetErrorString(e_)); \
  } while (0)
#define RET(call)          \
  do {                     \
    int r_ = (call);       \
    if (r_ != 0) return r_; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct Level {
  int n = 0, rows = 0, row0 = 0;
  bool dist = false;  // slab-distributed over ranks (needs halo exchange); false = whole grid on this rank
  Geo geo{};   // strip geometry for the register-heavy k_stokes kernels (~5 blocks/SM)
  Geo geo4{};  // ... for the 128-register variants (prolongation fused into the sweep: 4 blocks/SM)
  Geo geoR{};  // ... for the residual + restriction variant (28-column warp tiles: more blocks per strip)
  Geo geoL{};  // strip geometry for the light kernels (k_poisson, k_div, k_grad, k_jacobi0_F: 12-16 blocks/SM)
  Phys ph{};
  double* th = nullptr;    // padded theta: (rows+2) x n
  double* halo = nullptr;  // [2][5][n] receive rows (dist only): NCCL halos; peer-memory mode: the static stash of x's rows
  double* hland = nullptr; // [2][5][n] transient landing buffer of the rows fetched from the comm buffer (dist only)
  double *bF = nullptr, *xF = nullptr, *tF = nullptr, *rF = nullptr;  // 4*rows*n each
  double *bP = nullptr, *xP = nullptr, *tP = nullptr, *rP = nullptr;  // rows*n each
  double *gF = nullptr, *gP = nullptr;  // restricted slab before the all-gather (first replicated level only)
  double* wdF = nullptr;                // omega / diag(F), 4N (fused pre-smoothing)
  double* wdh = nullptr;                // dist: the ring neighbours' boundary rows of wdF, [2][4][n] (static)
  double* wdP = nullptr;                // 1 / diag(GtG), N (fused pressure pre-smoothing; whole-grid levels)
  size_t fs() const { return (size_t)rows * n; }
};

struct Bump {
  char* base = nullptr;
  size_t off = 0, cap = 0;
  bool dry = true;
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return p;
  }
};

struct mpbp_plan {
  mpbp_config cfg{};
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  std::vector<Level> lev;
  int first_repl = -1;  // index of the first replicated level when nranks > 1 (else -1)
  // coarsest dense (pseudo-)inverses, row-major on the device
  double *FinvT = nullptr, *PinvT = nullptr;
  int mF = 0, mP = 0;
  // tables for the analytic mass term (level 0)
  double *sxf = nullptr, *sxc = nullptr, *syf = nullptr, *syc = nullptr;
  // level-0 scratch
  double *w = nullptr, *g = nullptr, *t2 = nullptr;                 // 4N
  double *rinF = nullptr, *zF = nullptr, *dvF = nullptr;            // 4N
  double *rhs = nullptr, *xa = nullptr, *xb = nullptr, *xp = nullptr;  // N
  double* vin = nullptr;                                            // 5N: copy of the apply's input (graph-fixed address)
  double *rinP = nullptr, *zP = nullptr, *dvP = nullptr;            // N
  // reductions
  double* partial = nullptr;
  unsigned int* counter = nullptr;
  double* scal = nullptr;  // device scalars
  double* hscal = nullptr; // pinned host mirror
  int red_blocks = 592;
  // memory ownership
  void* owned = nullptr;
  void* kry_owned = nullptr;
  size_t kry_owned_bytes = 0;
  double* hostbuf_dev[2] = {nullptr, nullptr};  // staging for the *_host entry points
  long long launches = 0;
  cudaStream_t st = nullptr;    // the plan's own stream: all work runs here, fenced against the caller's stream
  cudaStream_t ust = nullptr;   // caller's stream of the current API call
  cudaStream_t own = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // V-cycles (rin -> z) replay as CUDA graphs: ~100 small launches per cycle become one graph launch
  // whole-grid levels, MPBP_FUSE overrides: bit 0 fused pre-smoothing pair, bit 1 fused prolongation + first
  // post-sweep, bit 2 fused residual + restriction (csrc/stokes.cuh)
  int fuse = 15;  // bit 3: the same three fusions on the pressure-Poisson cycle (whole-grid levels, csrc/poisson.cuh)
  // experimental (MPBP_PUSH_FUSED=1): smoothing / residual kernels on distributed levels push their own boundary
  // rows to the neighbours; the next stencil kernel on that vector then skips its k_halo_push
  bool push_fused = true;
  const double* pending_push = nullptr;  // vector whose halo rows the last kernel already pushed
  int cell_n = 256;  // whole-grid levels below level 0 with n <= cell_n use the cell-parallel kernels of cell.cuh (MPBP_CELL)
  int coarse_n = 0;  // experimental (MPBP_COARSE=<n>): whole-grid levels with n <= coarse_n run as ONE persistent kernel
  bool fused_mgs = true;
  bool lowsync = true;  // FGMRES orthogonalisation: low-synchronisation Gram-Schmidt (MPBP_ORTH=mgs: modified Gram-Schmidt)
  int jac_minb = 0;  // __launch_bounds__ min blocks/SM variant of the Jacobi kernel (register cap)
  bool use_graph = true;
  cudaGraphExec_t gexec_apply = nullptr;  // the whole preconditioner apply (vin -> w, t2, xp)
  long long glaunches_apply = 0;
  // ---- peer-memory halo exchange (nranks > 1): ring neighbours push rows into this rank's comm buffer ----
  bool p2p = false;
  char* comm_local = nullptr;               // cudaMalloc'd, exported with cudaIpc
  char *comm_prev = nullptr, *comm_next = nullptr;  // neighbours' comm buffers mapped into this process
  size_t comm_area = 0;                     // doubles per (slot, direction) halo area
  unsigned long long* dseq = nullptr;       // device-resident exchange counter (identical on all ranks)
  std::vector<char*> comm_all;              // every rank's comm buffer mapped here (index = rank; own = comm_local)
  std::vector<double> last_H;               // un-rotated Hessenberg of the last Arnoldi cycle, (last_m+1) x last_m row-major
  int last_m = 0, last_k = 0;               // allocated / used cycle length
  RedCtx red{};                             // peer-memory all-reduce of the reduction kernels (nranks <= 1: off)
};


static constexpr int kScal = 1024;

static int build_levels_shape(const mpbp_config& c, std::vector<Level>& lev, int& first_repl) {
  lev.clear();
  first_repl = -1;
  const int P = c.nranks;
  int n = c.n;
  if (P > 1 && (n % P != 0)) return set_err(MPBP_E_ARG, "n=%d not divisible by nranks=%d", n, P);
  bool dist = P > 1;
  int rows = n / P;
  while (true) {
    Level L;
    L.n = n;
    L.dist = dist;
    L.rows = dist ? rows : n;
    L.row0 = dist ? c.rank * rows : 0;
    lev.push_back(L);
    const bool last = c.operators_only || (n <= c.n_coarse) || (n % 2) || (n / 2 < 2);
    if (last) break;
    if (dist) {
      // the next level stays distributed only if it keeps >= 2 rows per rank, is not the coarsest
      // level (the dense solve is replicated) and this level's slab can be restricted locally
      if (rows % 2) return set_err(MPBP_E_ARG, "slab of %d rows at level n=%d cannot be coarsened", rows, n);
      const int nn = n / 2, nrows = rows / 2;
      const bool next_last = (nn <= c.n_coarse) || (nn % 2) || (nn / 2 < 2);
      // default replication boundary: with the LL halo protocol a distributed small level costs a replicated one plus
      // ~2 us per kernel, while the all-gather at the boundary shrinks 4x per level (8 ranks: 103.1 vs 100.5 its/s)
      int dmin = c.dist_min_n > 0 ? c.dist_min_n : 512;
      if (const char* e = getenv("MPBP_DIST_MIN_N")) dmin = std::max(8, atoi(e));  // tuning knob (all ranks must agree)
      if (nrows < 2 || (nrows % 2) || next_last || nn < dmin) {
        dist = false;
        first_repl = (int)lev.size();
      }
      rows = nrows;
    }
    n /= 2;
  }
  if (!c.operators_only && lev.back().n > 16)
    return set_err(MPBP_E_ARG, "n=%d coarsens only down to %d (> 16): n needs more factors of 2", c.n, lev.back().n);
  if (P > 1 && lev.back().dist && !c.operators_only)
    return set_err(MPBP_E_UNSUPPORTED, "nranks>1 needs at least one coarsening step (n=%d, n_coarse=%d)", c.n, c.n_coarse);
  if ((int)lev.size() > 1 && (c.nu1 < 1 || c.nu2 < 0)) return set_err(MPBP_E_ARG, "need nu1>=1, nu2>=0");
  if (5.0 * (double)lev[0].rows * (double)lev[0].n >= 2147483647.0)
    return set_err(MPBP_E_UNSUPPORTED, "slab of %d x %d cells exceeds the 32-bit element offsets of the kernels", lev[0].rows, lev[0].n);
  return 0;
}

static void carve(mpbp_plan* p, Bump& B) {
  const int L = (int)p->lev.size();
  for (int l = 0; l < L; ++l) {
    Level& v = p->lev[l];
    const size_t fs = v.fs();
    v.th = B.take<double>((size_t)(v.rows + 2) * v.n);
    v.halo = v.dist ? B.take<double>((size_t)2 * 5 * v.n) : nullptr;
    v.hland = v.dist ? B.take<double>((size_t)2 * 5 * v.n) : nullptr;
    v.tF = B.take<double>(4 * fs);
    v.rF = B.take<double>(4 * fs);
    v.tP = B.take<double>(fs);
    v.rP = B.take<double>(fs);
    if (l > 0) {
      v.bF = B.take<double>(4 * fs);
      v.xF = B.take<double>(4 * fs);
      v.bP = B.take<double>(fs);
      v.xP = B.take<double>(fs);
    }
    if (l < L - 1 && !p->cfg.operators_only) {
      v.wdF = B.take<double>(4 * fs);
      if (v.dist) v.wdh = B.take<double>((size_t)2 * 4 * v.n);
      if (!v.dist) v.wdP = B.take<double>(fs);
    }
    if (l == p->first_repl) {
      const Level& f = p->lev[l - 1];
      v.gF = B.take<double>(4 * (f.fs() / 4));
      v.gP = B.take<double>(f.fs() / 4);
    }
  }
  const size_t fs0 = p->lev[0].fs();
  p->w = B.take<double>(4 * fs0);
  p->g = B.take<double>(4 * fs0);
  p->t2 = B.take<double>(4 * fs0);
  p->rinF = B.take<double>(4 * fs0);
  p->zF = B.take<double>(4 * fs0);
  p->dvF = B.take<double>(4 * fs0);
  p->rhs = B.take<double>(fs0);
  p->xa = B.take<double>(fs0);
  p->xb = B.take<double>(fs0);
  p->xp = B.take<double>(fs0);
  p->vin = B.take<double>(5 * fs0);
  p->rinP = B.take<double>(fs0);
  p->zP = B.take<double>(fs0);
  p->dvP = B.take<double>(fs0);
  const Level& c = p->lev.back();
  p->mF = p->cfg.operators_only ? 0 : 4 * c.n * c.n;
  p->mP = p->cfg.operators_only ? 0 : c.n * c.n;
  p->FinvT = B.take<double>((size_t)p->mF * p->mF);
  p->PinvT = B.take<double>((size_t)p->mP * p->mP);
  const int n0 = p->lev[0].n;
  p->sxf = B.take<double>(n0);
  p->sxc = B.take<double>(n0);
  p->syf = B.take<double>(n0);
  p->syc = B.take<double>(n0);
  p->partial = B.take<double>((size_t)kMaxRedBlocks * 2 * kMaxMulti);
  p->counter = B.take<unsigned int>(64);
  p->scal = B.take<double>(kScal);
  p->dseq = B.take<unsigned long long>(16);
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
static inline dim3 stencil_grid(const Level& v, const Geo& g) {
  return dim3((unsigned)((v.n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps)), (unsigned)strip_count(g));
}
// Rows per strip: long strips amortise the two re-read halo rows, but the block count should fill whole
// waves of `cap` resident blocks (wave quantisation costs up to 2x on the slab sizes of 4-8 GPUs).
static int choose_rs(int gx, int rows, int cap) {
  if (rows <= 8) return rows;
  const int s32 = (rows + 31) / 32;
  if ((long long)gx * s32 >= 4LL * cap) return 32;
  double best = -1.0;
  int best_rs = std::min(rows, 32);
  const int inc = (rows & 1) ? 1 : 2;  // even strips on even grids: the row-pair kernels (stokes.cuh IN 2 / EP 2) need them
  for (int rs = 4; rs <= std::min(rows, 64); rs += inc) {
    const int S = (rows + rs - 1) / rs;
    const long long B = (long long)gx * S;
    const long long waves = (B + cap - 1) / cap;
    const double eff = (double)B / (double)(waves * cap);
    const double score = eff * (double)rs / (double)(rs + 2);  // halo rows re-read per strip
    if (score > best + 1e-12) best = score, best_rs = rs;
  }
  return best_rs;
}
// Strip decomposition of a level for kernels with `cap` resident blocks on the device.  Large levels: SINGLE WAVE --
// as many interior strips as fit next to each other (gx * S <= cap), so the launch has no tail wave, plus two short
// edge strips that are scheduled first (halo waits / pushes, general code path).  Small levels keep the wave-aware
// uniform strips of choose_rs.
static void set_strips(Geo& g, int gx, int rows, int cap) {
  g.re = 0;
  const int e = 8;
  const int S = cap / std::max(gx, 1);
  if (!(rows & 1) && rows >= 64 && S >= 1) {
    int rs = ((rows - 2 * e) + S - 1) / S;
    rs += rs & 1;
    if (rs >= 16) {
      g.rs = rs;
      g.re = e;
      return;
    }
  }
  g.rs = choose_rs(gx, rows, cap);
  if (!(rows & 1) && (g.rs & 1)) g.rs += 1;
}
static inline int ew_blocks(size_t len) { return (int)std::min<size_t>((len + 255) / 256, 148 * 16); }

#define LAUNCH_CHECK(p)            \
  do {                             \
    (p)->launches++;               \
    (p)->pending_push = nullptr;   \
    CU(cudaGetLastError());        \
  } while (0)

// neighbour rows of a distributed level's vector (nf fields, field stride fs) into lev.halo
static int halo_exchange(mpbp_plan* p, Level& v, const double* x, int nf, size_t fs) {
  const int P = p->nranks, prev = (p->rank + P - 1) % P, next = (p->rank + 1) % P;
  const size_t n = v.n;
  if (p->p2p) {
    // push my boundary rows straight into the neighbours' halo areas (dir 0 = top, 1 = bot) and release
    // their flags; consumers wait on the flags inside the stencil kernel (edge strips only)
    k_halo_push<<<(nf * v.n + 255) / 256, 256, 0, p->st>>>(x, nf, fs, v.rows, v.n, p->comm_prev, p->comm_next,
                                                           p->comm_local, p->comm_area, p->dseq, p->counter + 32);
    p->launches++;
    CU(cudaGetLastError());
    return 0;
  }
  double* top = v.halo;
  double* bot = v.halo + 5 * n;
  NC(ncclGroupStart());
  for (int k = 0; k < nf; ++k) {
    // order matters when prev == next (2 ranks): my last row is the peer's top halo
    NC(ncclSend(x + k * fs + (size_t)(v.rows - 1) * n, n, ncclDouble, next, p->comm, p->st));
    NC(ncclSend(x + k * fs, n, ncclDouble, prev, p->comm, p->st));
    NC(ncclRecv(top + k * n, n, ncclDouble, prev, p->comm, p->st));
    NC(ncclRecv(bot + k * n, n, ncclDouble, next, p->comm, p->st));
  }
  NC(ncclGroupEnd());
  return 0;
}

// one-off (plan creation) fetch of the ring neighbours' boundary rows of a static per-level field into `dst`
// ([2][nf][n]: top rows, then bottom rows) with NCCL send/recv
static int fetch_static_halo(mpbp_plan* p, Level& v, const double* x, int nf, double* dst) {
  const int P = p->nranks, prev = (p->rank + P - 1) % P, next = (p->rank + 1) % P;
  const size_t n = v.n, fs = v.fs();
  NC(ncclGroupStart());
  for (int k = 0; k < nf; ++k) {
    NC(ncclSend(x + k * fs + (size_t)(v.rows - 1) * n, n, ncclDouble, next, p->comm, p->st));
    NC(ncclSend(x + k * fs, n, ncclDouble, prev, p->comm, p->st));
    NC(ncclRecv(dst + k * n, n, ncclDouble, prev, p->comm, p->st));
    NC(ncclRecv(dst + (size_t)nf * n + k * n, n, ncclDouble, next, p->comm, p->st));
  }
  NC(ncclGroupEnd());
  return 0;
}

// view of a level vector for stencil kernels (performs the halo exchange when distributed)
static int make_view(mpbp_plan* p, Level& v, const double* x, int nf, VecIn& out, bool stash = false) {
  const size_t fs = v.fs();
  out.x = x;
  out.fs = fs;
  if (v.dist) {
    // the kernel that produced x may already have pushed its boundary rows (fused push): nothing to exchange then
    const bool already = p->p2p && p->pending_push != nullptr && p->pending_push == x;
    if (!already) RET(halo_exchange(p, v, x, nf, fs));
    p->pending_push = nullptr;
    if (p->p2p) {
      out.dseq = p->dseq;
      out.comm = p->comm_local;
      out.area = p->comm_area;
      out.land = stash ? v.halo : v.hland;  // the consumer's edge strips land the fetched rows here ...
      out.top = out.land;                    // ... and read them back through the ordinary halo pointers
      out.bot = out.land + 5 * (size_t)v.n;
    } else {
      out.top = v.halo;
      out.bot = v.halo + 5 * (size_t)v.n;
    }
    out.hs = v.n;
  } else {
    out.top = x + (size_t)(v.rows - 1) * v.n;
    out.bot = x;
    out.hs = fs;
  }
  return 0;
}

static int allreduce_scal(mpbp_plan* p, double* dev, int count) {
  if (p->nranks > 1) NC(ncclAllReduce(dev, dev, count, ncclDouble, ncclSum, p->comm, p->st));
  return 0;
}

// ---- operator launches -----------------------------------------------------------------------
static inline PushOut push_out(mpbp_plan* p) {
  return PushOut{p->comm_prev, p->comm_next, p->comm_local, p->comm_area, p->dseq, p->counter + 40};
}
// one launch of the unified marching kernel (csrc/stokes.cuh); `a` carries the variant's extra arguments
struct SxKind {
  int in = 0, mode = 0, ep = 0;
  bool with_p = false, push = false;
};
template <int IN, int MODE, bool WP, int EP, bool PUSH, int MINB>
static void sx_launch(mpbp_plan* p, dim3 grid, const StokesArgs& a) {
  k_stokes_x<IN, MODE, WP, EP, PUSH, MINB><<<grid, kBlockThreads, 0, p->st>>>(a);
}
static int launch_sx(mpbp_plan* p, int l, SxKind k, StokesArgs& a) {
  Level& v = p->lev[l];
  a.th = v.th;
  a.g = (k.in == 2) ? v.geo4 : (k.ep == 2 ? v.geoR : v.geo);
  a.ph = v.ph;
  const int wc = (k.ep == 2) ? WarpTile<2>::cols : WarpTile<0>::cols;
  if ((k.in == 2 || k.ep == 2) && ((v.rows & 1) || (a.g.rs & 1) || (a.g.re & 1) || (v.dist && k.ep == 2 && !k.push)))
    return set_err(MPBP_E_STATE, "internal: row-pair kernel on an odd / distributed level (rows %d, rs %d)", v.rows, a.g.rs);
  const dim3 grid((unsigned)((v.n + wc * kBlockWarps - 1) / (wc * kBlockWarps)), (unsigned)strip_count(a.g));
  if (k.push) a.po = push_out(p);
  const int key = k.in * 10000 + k.mode * 1000 + (k.with_p ? 100 : 0) + k.ep * 10 + (k.push ? 1 : 0);
  switch (key) {
    case 100: sx_launch<0, 0, true, 0, false, 5>(p, grid, a); break;    // y = A x
    case 0: sx_launch<0, 0, false, 0, false, 5>(p, grid, a); break;      // y = F x
    case 1000: sx_launch<0, 1, false, 0, false, 5>(p, grid, a); break;   // residual
    case 1001: sx_launch<0, 1, false, 0, true, 5>(p, grid, a); break;
    case 2000: sx_launch<0, 2, false, 0, false, 5>(p, grid, a); break;   // Jacobi sweep
    case 2001: sx_launch<0, 2, false, 0, true, 5>(p, grid, a); break;
    case 2010: sx_launch<0, 2, false, 1, false, 5>(p, grid, a); break;   // last sweep + Chebyshev update
    case 2011: sx_launch<0, 2, false, 1, true, 5>(p, grid, a); break;
    case 12000: sx_launch<1, 2, false, 0, false, 5>(p, grid, a); break;  // pre-smoothing pair from b
    case 12001: sx_launch<1, 2, false, 0, true, 5>(p, grid, a); break;
    case 1020: sx_launch<0, 1, false, 2, false, 5>(p, grid, a); break;   // residual + restriction
    case 1021: sx_launch<0, 1, false, 2, true, 5>(p, grid, a); break;    // ... on a slab of a distributed level
    case 22000: sx_launch<2, 2, false, 0, false, 4>(p, grid, a); break;  // prolongation + first post-sweep
    case 22001: sx_launch<2, 2, false, 0, true, 4>(p, grid, a); break;
    case 22010: sx_launch<2, 2, false, 1, false, 4>(p, grid, a); break;  // ... which is also the last one
    case 22011: sx_launch<2, 2, false, 1, true, 4>(p, grid, a); break;
    default: return set_err(MPBP_E_STATE, "internal: no stokes kernel variant %d", key);
  }
  LAUNCH_CHECK(p);
  return 0;
}

// ---- small whole-grid levels: cell-parallel kernels (csrc/cell.cuh), one load round trip per launch ----
static inline bool use_cell(const mpbp_plan* p, int l) {
  const Level& v = p->lev[l];
  return l > 0 && !v.dist && v.n <= p->cell_n && !(v.n & 1) && v.ph.mass_mode == 0;
}
static CellArgs cell_args(mpbp_plan* p, int l) {
  Level& v = p->lev[l];
  CellArgs a{};
  a.th = v.th;
  a.ph = v.ph;
  a.n = v.n;
  a.omega = p->cfg.omega;
  return a;
}
static inline dim3 cell_grid(int nx, int ny) { return dim3((nx + kCellBX - 1) / kCellBX, (ny + kCellBY - 1) / kCellBY); }
// in: 0 sweep of x, 1 pre-smoothing pair from b, 2 prolongation + sweep, 3 residual + restriction
static int cell_launch(mpbp_plan* p, int l, int in, const double* x, const double* b, double* y) {
  Level& v = p->lev[l];
  CellArgs a = cell_args(p, l);
  a.x = x;
  a.b = b;
  a.y = y;
  a.wd = v.wdF;
  const dim3 block(kCellBX * kCellBY);
  if (in == 3) {
    a.bc = p->lev[l + 1].bF;
    k_cell_rr<<<cell_grid(v.n / 2, v.n / 2), block, 0, p->st>>>(a);
  } else if (in == 2) {
    a.ec = p->lev[l + 1].xF;
    k_cell_sweep<2><<<cell_grid(v.n, v.n), block, 0, p->st>>>(a);
  } else if (in == 1) {
    k_cell_sweep<1><<<cell_grid(v.n, v.n), block, 0, p->st>>>(a);
  } else {
    k_cell_sweep<0><<<cell_grid(v.n, v.n), block, 0, p->st>>>(a);
  }
  LAUNCH_CHECK(p);
  return 0;
}

// y = Op x (mode 0), b - F x (mode 1), x + omega (b - F x)/diag (mode 2); optional Chebyshev epilogue on mode 2.
// On distributed levels the smoothing / residual kernels push their own boundary rows to the ring neighbours.
static int op_stokes(mpbp_plan* p, int l, int mode, bool with_p, const double* x, const double* b, double* y,
                     double omega, const ChebEp* ce = nullptr, bool stash = false) {
  Level& v = p->lev[l];
  if (mode == 2 && !with_p && !ce && omega == p->cfg.omega && use_cell(p, l)) return cell_launch(p, l, 0, x, b, y);
  StokesArgs a{};
  RET(make_view(p, v, x, with_p ? 5 : 4, a.xin, stash));
  a.b = b;
  a.y = y;
  a.omega = omega;
  SxKind k;
  k.mode = mode;
  k.with_p = with_p;
  if (ce) {
    a.ce = *ce;
    k.ep = 1;
  }
  k.push = v.dist && p->p2p && p->push_fused && mode != 0;
  RET(launch_sx(p, l, k, a));
  if (k.push) p->pending_push = ce ? ce->xk : y;
  return 0;
}
static int op_jacobi0_F(mpbp_plan* p, int l, const double* b, double* y, double omega) {
  Level& v = p->lev[l];
  k_jacobi0_F<<<stencil_grid(v, v.geoL), kBlockThreads, 0, p->st>>>(v.th, b, y, v.fs(), v.geoL, v.ph, omega);
  LAUNCH_CHECK(p);
  return 0;
}
// x2 = x1 + wd (b - F x1), x1 = wd b: the two pre-smoothing sweeps from a zero guess in one pass over b
static int op_presmooth_pair(mpbp_plan* p, int l, const double* b, double* y) {
  Level& v = p->lev[l];
  if (use_cell(p, l)) return cell_launch(p, l, 1, nullptr, b, y);
  StokesArgs a{};
  RET(make_view(p, v, b, 4, a.xin));
  a.wd.x = v.wdF;
  a.wd.fs = v.fs();
  if (v.dist) {  // the neighbours' rows of omega/diag(F) were fetched once at plan creation
    a.wd.top = v.wdh;
    a.wd.bot = v.wdh + (size_t)4 * v.n;
    a.wd.hs = v.n;
  } else {
    a.wd.top = v.wdF + (size_t)(v.rows - 1) * v.n;
    a.wd.bot = v.wdF;
    a.wd.hs = v.fs();
  }
  a.y = y;
  SxKind k;
  k.in = 1;
  k.mode = 2;
  k.push = v.dist && p->p2p && p->push_fused;
  RET(launch_sx(p, l, k, a));
  if (k.push) p->pending_push = y;
  return 0;
}
// coarse rhs b_{l+1} = R (b - F x): residual and full-weighting restriction in one pass.  On a distributed level the
// kernel also exchanges what the restriction needs across the slab boundary (the previous rank's last-row half-sums),
// pushes the coarse rhs's own boundary rows to the ring neighbours (it is the next level's pre-smoother input) and
// lands x's halo rows in the level's stash for the prolongation + sweep kernel of the same cycle.
static int op_residual_restrict(mpbp_plan* p, int l, const double* x, const double* b, bool stash) {
  Level& v = p->lev[l];
  Level& c = p->lev[l + 1];
  if (use_cell(p, l)) return cell_launch(p, l, 3, x, b, nullptr);
  const bool gather = v.dist && (l + 1 == p->first_repl);
  StokesArgs a{};
  RET(make_view(p, v, x, 4, a.xin, stash));
  a.b = b;
  a.bc = gather ? c.gF : c.bF;
  a.nc = c.n;
  a.rows_c = v.rows / 2;
  SxKind k;
  k.mode = 1;
  k.ep = 2;
  k.push = v.dist;
  RET(launch_sx(p, l, k, a));
  if (k.push && c.dist) p->pending_push = c.bF;
  if (gather) {
    const size_t cnt = (size_t)(v.rows / 2) * c.n;
    NC(ncclGroupStart());
    for (int f = 0; f < 4; ++f) NC(ncclAllGather(c.gF + f * cnt, c.bF + f * c.fs(), cnt, ncclDouble, p->comm, p->st));
    NC(ncclGroupEnd());
  }
  return 0;
}
// y = xt + omega (b - F xt)/diag, xt = x + P x_{l+1}: coarse-grid correction + first post-smoothing sweep in one pass
static int op_prolong_sweep(mpbp_plan* p, int l, const double* x, const double* b, double* y, double omega,
                            const ChebEp* ce) {
  Level& v = p->lev[l];
  Level& c = p->lev[l + 1];
  if (!ce && omega == p->cfg.omega && use_cell(p, l)) return cell_launch(p, l, 2, x, b, y);
  StokesArgs a{};
  if (v.dist) {
    // x's halo rows were stashed by the residual kernel of this cycle (the comm slots have been recycled since)
    a.xin.x = x;
    a.xin.fs = v.fs();
    a.xin.hs = v.n;
    a.xin.top = v.halo;
    a.xin.bot = v.halo + (size_t)5 * v.n;
    if (l + 1 == p->first_repl) {
      // coarse level is replicated: address this rank's rows inside the full coarse grid
      const int R0 = v.row0 / 2, rows_c = v.rows / 2, nc = c.n;
      a.cin.x = c.xF + (size_t)R0 * nc;
      a.cin.fs = a.cin.hs = c.fs();
      a.cin.top = c.xF + (size_t)((R0 + nc - 1) % nc) * nc;
      a.cin.bot = c.xF + (size_t)((R0 + rows_c) % nc) * nc;
    } else {
      RET(make_view(p, c, c.xF, 4, a.cin));
    }
  } else {
    RET(make_view(p, v, x, 4, a.xin));
    a.cin.x = c.xF;
    a.cin.fs = a.cin.hs = c.fs();
    a.cin.top = c.xF + (size_t)(c.rows - 1) * c.n;
    a.cin.bot = c.xF;
  }
  a.nc = c.n;
  a.rows_c = v.rows / 2;
  a.b = b;
  a.y = y;
  a.omega = omega;
  SxKind k;
  k.in = 2;
  k.mode = 2;
  if (ce) {
    a.ce = *ce;
    k.ep = 1;
  }
  k.push = v.dist && p->p2p && p->push_fused;
  RET(launch_sx(p, l, k, a));
  if (k.push) p->pending_push = ce ? ce->xk : y;
  return 0;
}
static int op_poisson(mpbp_plan* p, int l, int mode, const double* x, const double* b, double* y, double omega,
                      const ChebEp* ce = nullptr) {
  Level& v = p->lev[l];
  VecIn in{};
  if (mode != 3) RET(make_view(p, v, x, 1, in));
  const dim3 grid = stencil_grid(v, v.geoL), block(kBlockThreads);
  // distributed levels: the sweeps push their own boundary rows (the residual's consumer, the cell-average
  // restriction, needs no halo)
  const bool push = v.dist && p->p2p && p->push_fused && mode >= 2;
  const PushOut po = push ? push_out(p) : PushOut{};
  if (ce) {
    if (mode != 2) return set_err(MPBP_E_STATE, "internal: Chebyshev epilogue on a non-sweep");
    if (push) k_poisson<2, true, true><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, *ce, po);
    else k_poisson<2, true, false><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, *ce, po);
    LAUNCH_CHECK(p);
    if (push) p->pending_push = ce->xk;
    return 0;
  }
  const ChebEp none{};
  switch (mode) {
    case 0: k_poisson<0><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po); break;
    case 1: k_poisson<1><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po); break;
    case 2:
      if (push) k_poisson<2, false, true><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po);
      else k_poisson<2><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po);
      break;
    default:
      if (push) k_poisson<3, false, true><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po);
      else k_poisson<3><<<grid, block, 0, p->st>>>(in, v.th, b, y, v.geoL, v.ph, omega, none, po);
      break;
  }
  LAUNCH_CHECK(p);
  if (push) p->pending_push = y;
  return 0;
}
// fused pressure-Poisson kernels on whole-grid levels (csrc/poisson.cuh): in 1 pre-smoothing pair from b, 3 residual +
// restriction (coarse rhs of level l+1), 2 prolongation of level l+1's correction + first post-sweep (optionally the last)
static int op_poisson_f(mpbp_plan* p, int l, int in, const double* x, const double* b, double* y, const ChebEp* ce) {
  Level& v = p->lev[l];
  PoissonFArgs a{};
  a.x = x;
  a.b = b;
  a.wd = v.wdP;
  a.th = v.th;
  a.y = y;
  a.g = v.geoL;
  a.ph = v.ph;
  a.omega = p->cfg.omega;
  const dim3 grid = stencil_grid(v, v.geoL), block(kBlockThreads);
  if (in == 1) {
    k_poisson_f<1, 2, 0><<<grid, block, 0, p->st>>>(a);
  } else if (in == 3) {
    a.bc = p->lev[l + 1].bP;
    k_poisson_f<0, 1, 2><<<grid, block, 0, p->st>>>(a);
  } else {
    a.ec = p->lev[l + 1].xP;
    if (ce) {
      a.ce = *ce;
      k_poisson_f<2, 2, 1><<<grid, block, 0, p->st>>>(a);
    } else {
      k_poisson_f<2, 2, 0><<<grid, block, 0, p->st>>>(a);
    }
  }
  LAUNCH_CHECK(p);
  return 0;
}
// r = scale * D w + add
static int op_div(mpbp_plan* p, int l, const double* w, const double* add, double* r, double scale) {
  Level& v = p->lev[l];
  VecIn in{};
  RET(make_view(p, v, w, 4, in));
  Phys ph = v.ph;
  ph.inv_h *= scale;
  k_div<<<stencil_grid(v, v.geoL), kBlockThreads, 0, p->st>>>(in, v.th, add, r, v.geoL, ph);
  LAUNCH_CHECK(p);
  return 0;
}
static int op_grad(mpbp_plan* p, int l, const double* pr, double* y) {
  Level& v = p->lev[l];
  VecIn in{};
  RET(make_view(p, v, pr, 1, in));
  const bool push = v.dist && p->p2p && p->push_fused;
  if (push) k_grad<true><<<stencil_grid(v, v.geoL), kBlockThreads, 0, p->st>>>(in, v.th, y, v.fs(), v.geoL, v.ph, push_out(p));
  else k_grad<false><<<stencil_grid(v, v.geoL), kBlockThreads, 0, p->st>>>(in, v.th, y, v.fs(), v.geoL, v.ph, PushOut{});
  LAUNCH_CHECK(p);
  if (push) p->pending_push = y;
  return 0;
}
static int op_dense(mpbp_plan* p, const double* M, const double* x, double* y, int m) {
  k_dense_matvec<<<(m * 32 + 255) / 256, 256, 0, p->st>>>(M, x, y, m);
  LAUNCH_CHECK(p);
  return 0;
}

// ---- grid transfers ---------------------------------------------------------------------------
// b_{l+1} = R r_l
static int op_restrict(mpbp_plan* p, int l, bool isF, const double* r) {
  Level& f = p->lev[l];
  Level& c = p->lev[l + 1];
  const bool gather = (l + 1 == p->first_repl);
  const int rows_c = f.rows / 2, nc = f.n / 2;
  const dim3 block(128), grid((nc + 127) / 128, rows_c);
  if (isF) {
    VecIn in{};
    RET(make_view(p, f, r, 4, in));
    double* dst = gather ? c.gF : c.bF;
    const bool push = c.dist && p->p2p && p->push_fused;  // the coarse rhs is the input of the next level's pre-smoother
    if (push) k_restrict_F<true><<<grid, block, 0, p->st>>>(in, dst, f.n, f.rows, push_out(p));
    else k_restrict_F<false><<<grid, block, 0, p->st>>>(in, dst, f.n, f.rows, PushOut{});
    LAUNCH_CHECK(p);
    if (push) p->pending_push = dst;
    if (gather) {
      const size_t cnt = (size_t)rows_c * nc;
      NC(ncclGroupStart());
      for (int k = 0; k < 4; ++k)
        NC(ncclAllGather(dst + k * cnt, c.bF + k * c.fs(), cnt, ncclDouble, p->comm, p->st));
      NC(ncclGroupEnd());
    }
  } else {
    double* dst = gather ? c.gP : c.bP;
    k_restrict_P<<<grid, block, 0, p->st>>>(r, dst, f.n, f.rows);
    LAUNCH_CHECK(p);
    if (gather) NC(ncclAllGather(dst, c.bP, (size_t)rows_c * nc, ncclDouble, p->comm, p->st));
  }
  return 0;
}
// x_l += P x_{l+1}
static int op_prolong_add(mpbp_plan* p, int l, bool isF, double* x) {
  Level& f = p->lev[l];
  Level& c = p->lev[l + 1];
  const dim3 block(128), grid((f.n + 127) / 128, f.rows);
  if (isF) {
    VecIn in{};
    if (l + 1 == p->first_repl) {
      // coarse level is replicated: address this rank's rows inside the full coarse grid
      const int R0 = f.row0 / 2, rows_c = f.rows / 2, nc = c.n;
      in.x = c.xF + (size_t)R0 * nc;
      in.fs = in.hs = c.fs();
      in.top = c.xF + (size_t)((R0 + nc - 1) % nc) * nc;
      in.bot = c.xF + (size_t)((R0 + rows_c) % nc) * nc;
    } else {
      RET(make_view(p, c, c.xF, 4, in));
    }
    k_prolong_add_F<<<grid, block, 0, p->st>>>(in, x, f.n, f.rows);
  } else {
    const double* xc = c.xP;
    if (l + 1 == p->first_repl) xc += (size_t)(f.row0 / 2) * c.n;
    const bool push = f.dist && p->p2p && p->push_fused;
    if (push) k_prolong_add_P<true><<<grid, block, 0, p->st>>>(xc, x, f.n, f.rows, push_out(p));
    else k_prolong_add_P<false><<<grid, block, 0, p->st>>>(xc, x, f.n, f.rows, PushOut{});
    LAUNCH_CHECK(p);
    if (push) p->pending_push = x;
    return 0;
  }
  LAUNCH_CHECK(p);
  return 0;
}

// ---- vector helpers ---------------------------------------------------------------------------
static int v_axpby(mpbp_plan* p, double a, const double* x, double b, const double* y, double* z, size_t len) {
  k_axpby<<<ew_blocks(len), 256, 0, p->st>>>(a, x, b, y, z, len);
  LAUNCH_CHECK(p);
  return 0;
}
static int v_copy(mpbp_plan* p, const double* x, double* y, size_t len) {
  p->pending_push = nullptr;
  CU(cudaMemcpyAsync(y, x, len * sizeof(double), cudaMemcpyDeviceToDevice, p->st));
  return 0;
}
// out_dev[k] = <V_k, w> summed over ranks; post_sqrt applies sqrt to each result
static int v_multi_dot(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* w, size_t len,
                       double* out_dev, bool post_sqrt) {
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  const bool fused_ar = p->red.nranks > 1;  // the all-reduce happens inside the kernel (peer memory)
  const int post = (post_sqrt && (p->nranks == 1 || fused_ar)) ? 1 : 0;
  for (int k0 = 0; k0 < nvec; k0 += kMaxMulti) {
    const int nv = std::min(kMaxMulti, nvec - k0);
    const double* Vk = V + (size_t)k0 * ld;
    double* o = out_dev + k0;
    switch (nv) {
      case 1: k_multi_dot<1><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 2: k_multi_dot<2><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 3: k_multi_dot<3><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 4: k_multi_dot<4><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 5: k_multi_dot<5><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 6: k_multi_dot<6><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      case 7: k_multi_dot<7><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
      default: k_multi_dot<8><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, len, p->partial, p->counter, o, post, p->red); break;
    }
    LAUNCH_CHECK(p);
  }
  if (p->nranks > 1 && !fused_ar) {
    RET(allreduce_scal(p, out_dev, nvec));
    if (post_sqrt) {
      k_sqrt_inplace<<<1, 256, 0, p->st>>>(out_dev, nvec);
      LAUNCH_CHECK(p);
    }
  }
  return 0;
}
static int v_dot(mpbp_plan* p, const double* x, const double* y, size_t len, double* out_dev) {
  return v_multi_dot(p, x, 0, 1, y, len, out_dev, false);
}
static int v_nrm2(mpbp_plan* p, const double* x, size_t len, double* out_dev) {
  return v_multi_dot(p, x, 0, 1, x, len, out_dev, true);
}
static int v_multi_axpy(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* alpha_host, double* y,
                        size_t len) {
  for (int k0 = 0; k0 < nvec; k0 += kMaxMulti) {
    const int nv = std::min(kMaxMulti, nvec - k0);
    Alphas al{};
    for (int k = 0; k < nv; ++k) al.a[k] = alpha_host[k0 + k];
    const double* Vk = V + (size_t)k0 * ld;
    const int blocks = ew_blocks(len);
    switch (nv) {
      case 1: k_multi_axpy<1><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 2: k_multi_axpy<2><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 3: k_multi_axpy<3><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 4: k_multi_axpy<4><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 5: k_multi_axpy<5><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 6: k_multi_axpy<6><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      case 7: k_multi_axpy<7><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
      default: k_multi_axpy<8><<<blocks, 256, 0, p->st>>>(Vk, ld, al, y, len); break;
    }
    LAUNCH_CHECK(p);
  }
  return 0;
}
// Fused modified Gram-Schmidt of w against V_0..V_j:  hd[k] = <V_k, w_k>, w_{k+1} = w_k - hd[k] V_k,
// nrm_after = ||w_{j+1}||; optionally nrm_before = ||w_0|| (scipy's h0).  All scalars stay on the device.
static int v_mgs(mpbp_plan* p, const double* V, size_t ld, int j, double* w, size_t len, double* hd, double* nrm_after,
                 double* nrm_before) {
  if (!p->fused_mgs) {
    if (nrm_before) RET(v_nrm2(p, w, len, nrm_before));
    for (int k = 0; k <= j; ++k) {
      RET(v_dot(p, V + (size_t)k * ld, w, len, hd + k));
      k_axpy_dev<<<ew_blocks(len), 256, 0, p->st>>>(hd + k, -1.0, V + (size_t)k * ld, w, len);
      LAUNCH_CHECK(p);
    }
    return v_nrm2(p, w, len, nrm_after);
  }
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  const bool fused_ar = p->red.nranks > 1;
  const int post = (p->nranks == 1 || fused_ar) ? 1 : 0;
  auto finish = [&](double* dot, double* nrm) -> int {
    if (p->nranks > 1 && !fused_ar) {
      if (dot) RET(allreduce_scal(p, dot, 1));
      if (nrm) {
        RET(allreduce_scal(p, nrm, 1));
        k_sqrt_inplace<<<1, 32, 0, p->st>>>(nrm, 1);
        LAUNCH_CHECK(p);
      }
    }
    return 0;
  };
  // first: h_0 = <V_0, w> (+ ||w||)
  if (nrm_before)
    k_mgs_fused<false, true, true><<<blocks, kRedThreads, 0, p->st>>>(nullptr, nullptr, V, w, len, p->partial, p->counter,
                                                                      hd, nrm_before, post, p->red);
  else
    k_mgs_fused<false, true, false><<<blocks, kRedThreads, 0, p->st>>>(nullptr, nullptr, V, w, len, p->partial,
                                                                       p->counter, hd, nullptr, post, p->red);
  LAUNCH_CHECK(p);
  RET(finish(hd, nrm_before));
  for (int k = 1; k <= j; ++k) {
    k_mgs_fused<true, true, false><<<blocks, kRedThreads, 0, p->st>>>(hd + k - 1, V + (size_t)(k - 1) * ld,
                                                                      V + (size_t)k * ld, w, len, p->partial, p->counter,
                                                                      hd + k, nullptr, post, p->red);
    LAUNCH_CHECK(p);
    RET(finish(hd + k, nullptr));
  }
  k_mgs_fused<true, false, true><<<blocks, kRedThreads, 0, p->st>>>(hd + j, V + (size_t)j * ld, nullptr, w, len, p->partial,
                                                                    p->counter, nullptr, nrm_after, post, p->red);
  LAUNCH_CHECK(p);
  return finish(nullptr, nrm_after);
}

// Low-synchronisation Gram-Schmidt, pass 1: r = V^T w and g = V^T u in ONE pass over the basis (u = newest basis
// vector).  out_dev: chunks of 8 vectors, chunk c at out_dev + 16c holds [r (nv values) | g (nv values)].
static int v_multi_dot2(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* w, const double* u, size_t len,
                        double* out_dev) {
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  for (int k0 = 0, c = 0; k0 < nvec; k0 += kMaxMulti, ++c) {
    const int nv = std::min(kMaxMulti, nvec - k0);
    const double* Vk = V + (size_t)k0 * ld;
    double* o = out_dev + 16 * c;
    switch (nv) {
      case 1: k_multi_dot2<1><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 2: k_multi_dot2<2><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 3: k_multi_dot2<3><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 4: k_multi_dot2<4><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 5: k_multi_dot2<5><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 6: k_multi_dot2<6><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      case 7: k_multi_dot2<7><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
      default: k_multi_dot2<8><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, w, u, len, p->partial, p->counter, o, p->red); break;
    }
    LAUNCH_CHECK(p);
  }
  return 0;
}
// pass 2: w -= V h (h on the host), nrm_dev = ||w_new|| (summed over ranks inside the kernel)
static int v_multi_axpy_nrm(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* h_host, double* w, size_t len,
                            double* nrm_dev) {
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  for (int k0 = 0; k0 < nvec; k0 += kMaxMulti) {
    const int nv = std::min(kMaxMulti, nvec - k0);
    const bool last = k0 + nv >= nvec;
    Alphas al{};
    for (int k = 0; k < nv; ++k) al.a[k] = h_host[k0 + k];
    const double* Vk = V + (size_t)k0 * ld;
#define MPBP_AXN(NV)                                                                                                        \
  if (last) k_multi_axpy_nrm<NV, true><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, al, w, len, p->partial, p->counter, nrm_dev, p->red); \
  else k_multi_axpy_nrm<NV, false><<<blocks, kRedThreads, 0, p->st>>>(Vk, ld, al, w, len, p->partial, p->counter, nrm_dev, p->red);
    switch (nv) {
      case 1: MPBP_AXN(1) break;
      case 2: MPBP_AXN(2) break;
      case 3: MPBP_AXN(3) break;
      case 4: MPBP_AXN(4) break;
      case 5: MPBP_AXN(5) break;
      case 6: MPBP_AXN(6) break;
      case 7: MPBP_AXN(7) break;
      default: MPBP_AXN(8) break;
    }
#undef MPBP_AXN
    LAUNCH_CHECK(p);
  }
  return 0;
}

// fetch `count` device scalars to the pinned host mirror (synchronises the stream)
static int fetch_scal(mpbp_plan* p, const double* dev, int count, double* host) {
  CU(cudaMemcpyAsync(host, dev, count * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  CU(cudaStreamSynchronize(p->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// sub-solvers
// ---------------------------------------------------------------------------------------------
// `sweeps` damped Jacobi sweeps starting from x (in place); tmp is a same-size scratch buffer
static int jacobi_sweeps(mpbp_plan* p, int l, bool isF, const double* b, double* x, double* tmp, int sweeps,
                         double omega) {
  Level& v = p->lev[l];
  const size_t len = (isF ? 4 : 1) * v.fs();
  double* cur = x;
  double* oth = tmp;
  for (int s = 0; s < sweeps; ++s) {
    if (isF) RET(op_stokes(p, l, 2, false, cur, b, oth, omega));
    else RET(op_poisson(p, l, 2, cur, b, oth, omega));
    std::swap(cur, oth);
  }
  if (cur != x) RET(v_copy(p, cur, x, len));
  return 0;
}

// experimental: the whole sub-hierarchy from level l down as one single-CTA kernel (csrc/coarse.cuh)
static int coarse_vcycle_launch(mpbp_plan* p, int l, bool isF, const double* b, double* x) {
  const int L = (int)p->lev.size();
  CoarseArgs a{};
  a.nlev = L - l;
  for (int i = 0; i < a.nlev; ++i) {
    Level& v = p->lev[l + i];
    CoarseLevel& cl = a.lev[i];
    cl.n = v.n;
    cl.ph = v.ph;
    cl.th = v.th;
    cl.b = isF ? v.bF : v.bP;
    cl.x = isF ? v.xF : v.xP;
    cl.t = isF ? v.tF : v.tP;
    cl.r = isF ? v.rF : v.rP;
  }
  a.Minv_t = isF ? p->FinvT : p->PinvT;
  a.m = isF ? p->mF : p->mP;
  a.b_in = b;
  a.x_out = x;
  a.omega = p->cfg.omega;
  a.nu1 = p->cfg.nu1;
  a.nu2 = p->cfg.nu2;
  if (isF) k_coarse_vcycle<true><<<1, kCoarseThreadsF, 0, p->st>>>(a);
  else k_coarse_vcycle<false><<<1, kCoarseThreadsP, 0, p->st>>>(a);
  LAUNCH_CHECK(p);
  return 0;
}

// x = V b: one V(nu1,nu2) cycle from a zero guess.  x must not alias the level's t/r buffers.
// With `ce` (level 0 of a sub-solve) the cycle's output z is never stored: the last smoothing sweep feeds it
// straight into the Chebyshev / iteration update d = ca d + cb z, xk += d, and `x` is only scratch.
static int vcycle(mpbp_plan* p, int l, bool isF, const double* b, double* x, const ChebEp* ce = nullptr) {
  const mpbp_config& c = p->cfg;
  if (c.operators_only) return set_err(MPBP_E_STATE, "plan was created with operators_only");
  Level& v = p->lev[l];
  const int L = (int)p->lev.size();
  if (!ce && p->coarse_n > 0 && !v.dist && v.n <= p->coarse_n && l < L - 1 && (L - l) <= kCoarseMaxLevels)
    return coarse_vcycle_launch(p, l, isF, b, x);
  if (l == L - 1) {
    // coarsest level: dense inverse (F) / pseudo-inverse (GtG)
    RET(isF ? op_dense(p, p->FinvT, b, x, p->mF) : op_dense(p, p->PinvT, b, x, p->mP));
    if (ce) {  // single-level hierarchy: no sweep to carry the update
      const size_t len = (isF ? 4 : 1) * v.fs();
      k_cheb_ep<<<ew_blocks(len), 256, 0, p->st>>>(ce->ca, ce->cb, x, ce->d, ce->xk, ce->read_d, ce->read_x, ce->write_d, len);
      LAUNCH_CHECK(p);
    }
    return 0;
  }
  double* t = isF ? v.tF : v.tP;
  double* r = isF ? v.rF : v.rP;
  const bool even = !(v.rows & 1) && !(v.geo.rs & 1) && !(v.geo4.rs & 1) && !(v.geoR.rs & 1);
  const bool dist_ok = !v.dist || (p->p2p && p->push_fused);  // distributed levels: fused variants need the peer-memory halos
  const bool fuseP = !isF && (p->fuse & 8) && !v.dist && !(v.rows & 1) && !(v.geoL.rs & 1) && v.geoL.re == 0;
  const bool fuse_pre = isF ? (p->fuse & 1) && dist_ok && c.nu1 == 2 && v.wdF != nullptr : fuseP && c.nu1 == 2 && v.wdP != nullptr;
  // residual + restriction (slabs: peer-memory halos and at least two coarse rows per rank)
  const bool fuse_rr = isF ? (p->fuse & 4) && even && (!v.dist || (dist_ok && v.rows >= 4)) : fuseP;
  const bool fuse_post = isF ? (p->fuse & 2) && dist_ok && even && c.nu2 >= 1 : fuseP && c.nu2 >= 1;  // prolongation + first post-sweep
  // number of kernels that write the iterate into a fresh buffer (they ping-pong between x and t); the last one
  // must land in x unless it ends in the epilogue
  const int pre_w = fuse_pre ? 1 : c.nu1;
  const int post_w = c.nu2;  // unfused: prolongation is in place, then nu2 sweeps; fused: nu2 kernels as well
  const int W = pre_w + post_w - ((ce && post_w > 0) ? 1 : 0);
  double* cur = (W % 2 == 1) ? x : t;
  double* oth = (cur == x) ? t : x;
  if (c.nu2 == 0 && ce) return set_err(MPBP_E_UNSUPPORTED, "nu2 = 0 is not supported inside the sub-solves");
  // ---- pre-smoothing from a zero guess ----
  if (fuse_pre) {
    if (isF) RET(op_presmooth_pair(p, l, b, cur));
    else RET(op_poisson_f(p, l, 1, nullptr, b, cur, nullptr));
  } else {
    if (isF) RET(op_jacobi0_F(p, l, b, cur, c.omega));
    else RET(op_poisson(p, l, 3, nullptr, b, cur, c.omega));
    for (int s = 1; s < c.nu1; ++s) {
      if (isF) RET(op_stokes(p, l, 2, false, cur, b, oth, c.omega));
      else RET(op_poisson(p, l, 2, cur, b, oth, c.omega));
      std::swap(cur, oth);
    }
  }
  // ---- coarse-grid correction ----
  if (fuse_rr) {
    if (isF) RET(op_residual_restrict(p, l, cur, b, /*stash=*/fuse_post));
    else RET(op_poisson_f(p, l, 3, cur, b, nullptr, nullptr));
  } else {
    if (isF) RET(op_stokes(p, l, 1, false, cur, b, r, 0.0, nullptr, /*stash=*/fuse_post));
    else RET(op_poisson(p, l, 1, cur, b, r, 0.0));
    RET(op_restrict(p, l, isF, r));
  }
  Level& cl = p->lev[l + 1];
  RET(vcycle(p, l + 1, isF, isF ? cl.bF : cl.bP, isF ? cl.xF : cl.xP));
  // ---- post-smoothing ----
  int post = c.nu2;
  if (fuse_post) {
    const bool last = (post == 1);
    if (isF) RET(op_prolong_sweep(p, l, cur, b, oth, c.omega, (last && ce) ? ce : nullptr));
    else RET(op_poisson_f(p, l, 2, cur, b, oth, (last && ce) ? ce : nullptr));
    std::swap(cur, oth);
    post--;
  } else {
    RET(op_prolong_add(p, l, isF, cur));
  }
  for (int s = 0; s < post; ++s) {
    const ChebEp* e = (s == post - 1) ? ce : nullptr;
    if (isF) RET(op_stokes(p, l, 2, false, cur, b, oth, c.omega, e));
    else RET(op_poisson(p, l, 2, cur, b, oth, c.omega, e));
    std::swap(cur, oth);
  }
  if (!ce && cur != x) return set_err(MPBP_E_STATE, "internal: V-cycle ping-pong parity");
  return 0;
}

static int project_mean(mpbp_plan* p, double* x) {
  Level& v = p->lev[0];
  const size_t len = v.fs();
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  double* s = p->scal + kScal - 1;
  k_sum<<<blocks, kRedThreads, 0, p->st>>>(x, len, p->partial, p->counter, s, p->red);
  LAUNCH_CHECK(p);
  if (p->red.nranks <= 1) RET(allreduce_scal(p, s, 1));
  k_shift_dev<<<ew_blocks(len), 256, 0, p->st>>>(s, 1.0 / ((double)v.n * (double)v.n), x, len);
  LAUNCH_CHECK(p);
  return 0;
}

// x = F~^-1 b  /  x = (GtG)~^-1 b  (fixed linear operators; zero initial guess).  b is never copied and the V-cycle
// outputs are never materialised: cycle k runs on rhs_k (b, then the residual b - Op x) and its last sweep applies
// the update x += d_k directly (plain iteration: d = z; Chebyshev semi-iteration over [lmin, lmax]: d = ca d + cb z).
static int sub_solve(mpbp_plan* p, bool isF, const double* b, double* x) {
  const mpbp_config& c = p->cfg;
  Level& v = p->lev[0];
  const int kind = isF ? c.F_kind : c.P_kind;
  if (c.operators_only) return set_err(MPBP_E_STATE, "plan was created with operators_only");
  if (kind == MPBP_SUB_JACOBI) {
    const int sweeps = isF ? c.F_sweeps : c.P_sweeps;
    if (sweeps < 1) return set_err(MPBP_E_ARG, "sweeps must be >= 1");
    if (isF) RET(op_jacobi0_F(p, 0, b, x, c.omega));
    else RET(op_poisson(p, 0, 3, nullptr, b, x, c.omega));
    RET(jacobi_sweeps(p, 0, isF, b, x, isF ? v.tF : v.tP, sweeps - 1, c.omega));
  } else {
    const int cycles = isF ? c.F_cycles : c.P_cycles;
    if (cycles < 1) return set_err(MPBP_E_ARG, "cycles must be >= 1");
    double* rin = isF ? p->rinF : p->rinP;
    double* z = isF ? p->zF : p->zP;  // scratch of the level-0 cycle
    double* dv = isF ? p->dvF : p->dvP;
    const double th = 0.5 * (c.lmax + c.lmin), de = 0.5 * (c.lmax - c.lmin);
    const double sig = th / de;
    double rho_k = 1.0 / sig;
    for (int k = 0; k < cycles; ++k) {
      ChebEp ce{};
      ce.d = dv;
      ce.xk = x;
      ce.read_x = k > 0;
      if (!c.cheb) {
        ce.ca = 0.0;
        ce.cb = 1.0;  // x += z
        ce.read_d = 0;
        ce.write_d = 0;
      } else if (k == 0) {
        ce.ca = 0.0;
        ce.cb = 1.0 / th;  // d = z / theta ; x = d
        ce.read_d = 0;
        ce.write_d = cycles > 1;
      } else {
        const double rho_n = 1.0 / (2.0 * sig - rho_k);
        ce.ca = rho_n * rho_k;
        ce.cb = 2.0 * rho_n / de;
        ce.read_d = 1;
        ce.write_d = k < cycles - 1;
        rho_k = rho_n;
      }
      const double* rhs = b;
      if (k > 0) {
        if (isF) RET(op_stokes(p, 0, 1, false, x, b, rin, 0.0));
        else RET(op_poisson(p, 0, 1, x, b, rin, 0.0));
        rhs = rin;
      }
      RET(vcycle(p, 0, isF, rhs, z, &ce));
    }
  }
  if (!isF && c.project) RET(project_mean(p, x));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// the preconditioner apply: approx_schur_op, solve.py:257-277
// ---------------------------------------------------------------------------------------------
// everything of the apply that runs on plan-internal buffers: vin -> (w, t2, xp)
static int precond_core(mpbp_plan* p) {
  const size_t fs = p->lev[0].fs();
  const double* v = p->vin;
  RET(sub_solve(p, true, v, p->w));                       // :258  Finv_v = F_inv @ v[:4N]
  RET(op_div(p, 0, p->w, v + 4 * fs, p->rhs, 1.0));       // :259  rhs_interim = D Finv_v + v[4N:]
  RET(sub_solve(p, false, p->rhs, p->xa));                // :265  x_a = (GtG)~^-1 rhs_interim
  RET(op_grad(p, 0, p->xa, p->g));                        // :267  x_b = Gt_F_G x_a = -D F G x_a   (:246-249)
  RET(op_stokes(p, 0, 0, false, p->g, nullptr, p->t2, 0.0));
  RET(op_div(p, 0, p->t2, nullptr, p->xb, -1.0));
  RET(sub_solve(p, false, p->xb, p->xp));                 // :271  x_p = (GtG)~^-1 x_b
  RET(op_grad(p, 0, p->xp, p->g));                        // :273  G_xp = G x_p
  RET(sub_solve(p, true, p->g, p->t2));                   // :274  Finv_G_xp = F_inv @ G_xp
  return 0;
}
// The apply is one CUDA graph (about 1800 kernel nodes at 4096^2) between a copy-in of v and the final combine:
// every kernel inside works on plan-owned buffers, so its arguments never change and the graph is captured once.
static int precond_apply(mpbp_plan* p, const double* v, double* z) {
  const size_t fs = p->lev[0].fs();
  CU(cudaMemcpyAsync(p->vin, v, 5 * fs * sizeof(double), cudaMemcpyDeviceToDevice, p->st));
  p->pending_push = nullptr;
  if (!p->use_graph) {
    RET(precond_core(p));
  } else {
    if (!p->gexec_apply) {
      const long long l0 = p->launches;
      CU(cudaStreamBeginCapture(p->st, cudaStreamCaptureModeThreadLocal));
      const int rc = precond_core(p);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(p->st, &g);
      if (rc != 0) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      CU(e);
      CU(cudaGraphInstantiate(&p->gexec_apply, g, 0));
      cudaGraphDestroy(g);
      p->glaunches_apply = p->launches - l0;
      p->launches = l0;
    }
    CU(cudaGraphLaunch(p->gexec_apply, p->st));
    p->launches += p->glaunches_apply;
    p->pending_push = nullptr;
  }
  // :275  u = Finv_v - Finv_G_xp ; :276  concat(u, x_p)
  k_combine<<<ew_blocks(5 * fs), 256, 0, p->st>>>(p->w, p->t2, p->xp, z, 4 * fs, fs);
  LAUNCH_CHECK(p);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// plan creation
// ---------------------------------------------------------------------------------------------
extern "C" int mpbp_config_default(mpbp_config* c) {
  if (!c) return set_err(MPBP_E_ARG, "null config");
  memset(c, 0, sizeof(*c));
  c->n = 16;  // solve.py:291
  c->xi = 1.0;
  c->eta_n = 100.0;
  c->eta_s = 1.0;  // solve.py:295-297
  c->c = 1.0;
  c->d_u = -1.0;  // solve.py:292-293
  c->d_p = 1.0;
  c->d_div = -1.0;  // preconditioner.py:299
  c->rank = 0;
  c->nranks = 1;
  c->F_kind = c->P_kind = MPBP_SUB_MG;
  c->F_sweeps = c->P_sweeps = 20;
  c->F_cycles = 4;
  c->P_cycles = 2;
  c->omega = 0.8;
  c->nu1 = c->nu2 = 2;
  c->n_coarse = 4;
  c->cheb = 1;
  c->lmin = 0.75;
  c->lmax = 1.2;
  c->project = 1;
  return 0;
}

static int check_cfg(const mpbp_config* c) {
  if (!c) return set_err(MPBP_E_ARG, "null config");
  if (c->n < 2) return set_err(MPBP_E_ARG, "n must be >= 2 (got %d)", c->n);
  if (c->nranks < 1 || c->rank < 0 || c->rank >= c->nranks) return set_err(MPBP_E_ARG, "bad rank/nranks");
  if (c->nranks > 1 && !c->nccl_unique_id) return set_err(MPBP_E_ARG, "nranks>1 needs nccl_unique_id");
  if (c->n_coarse < 2 || c->n_coarse > 16) return set_err(MPBP_E_ARG, "n_coarse must be in [2,16]");
  if (!(c->omega > 0.0)) return set_err(MPBP_E_ARG, "omega must be > 0");
  if (c->cheb && !(c->lmax > c->lmin && c->lmin > 0.0)) return set_err(MPBP_E_ARG, "need 0 < lmin < lmax");
  return 0;
}

extern "C" int mpbp_plan_workspace_bytes(const mpbp_config* cfg, size_t* bytes) {
  RET(check_cfg(cfg));
  if (!bytes) return set_err(MPBP_E_ARG, "null bytes");
  mpbp_plan tmp;
  tmp.cfg = *cfg;
  RET(build_levels_shape(*cfg, tmp.lev, tmp.first_repl));
  Bump B;
  B.dry = true;
  carve(&tmp, B);
  *bytes = B.off + 256;
  return 0;
}

// dense inverse by Gauss-Jordan with partial pivoting (row-major, in place); returns false if singular
static bool invert_dense(std::vector<double>& A, int m) {
  double amax = 0.0;
  for (double v : A) amax = std::max(amax, std::fabs(v));
  const double tiny = 1e-13 * amax;  // relative pivot threshold: a numerically singular coarse operator is an error
  std::vector<double> I((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) I[(size_t)i * m + i] = 1.0;
  for (int col = 0; col < m; ++col) {
    int piv = col;
    double best = std::fabs(A[(size_t)col * m + col]);
    for (int r = col + 1; r < m; ++r) {
      const double v = std::fabs(A[(size_t)r * m + col]);
      if (v > best) best = v, piv = r;
    }
    if (!(best > tiny)) return false;
    if (piv != col)
      for (int k = 0; k < m; ++k) {
        std::swap(A[(size_t)piv * m + k], A[(size_t)col * m + k]);
        std::swap(I[(size_t)piv * m + k], I[(size_t)col * m + k]);
      }
    const double d = 1.0 / A[(size_t)col * m + col];
    for (int k = 0; k < m; ++k) {
      A[(size_t)col * m + k] *= d;
      I[(size_t)col * m + k] *= d;
    }
    for (int r = 0; r < m; ++r) {
      if (r == col) continue;
      const double f = A[(size_t)r * m + col];
      if (f == 0.0) continue;
      double* __restrict__ Ar = &A[(size_t)r * m];
      double* __restrict__ Ir = &I[(size_t)r * m];
      const double* __restrict__ Ac = &A[(size_t)col * m];
      const double* __restrict__ Ic = &I[(size_t)col * m];
      for (int k = col; k < m; ++k) Ar[k] -= f * Ac[k];  // columns < col of the pivot row are already zero
      for (int k = 0; k < m; ++k) Ir[k] -= f * Ic[k];
    }
  }
  A.swap(I);
  return true;
}

static int build_coarse_inverses(mpbp_plan* p) {
  const int l = (int)p->lev.size() - 1;
  Level& v = p->lev[l];
  // scratch: two vectors of the coarsest level.  With a single level the per-level x/b buffers do
  // not exist, so borrow level-0 scratch.
  double* xin = (l > 0) ? v.bF : p->w;
  double* yout = (l > 0) ? v.xF : p->g;
  for (int pass = 0; pass < 2; ++pass) {
    const bool isF = pass == 0;
    const int m = isF ? p->mF : p->mP;
    std::vector<double> M((size_t)m * m), col(m);
    const double one = 1.0;
    for (int j = 0; j < m; ++j) {
      CU(cudaMemsetAsync(xin, 0, m * sizeof(double), p->st));
      CU(cudaMemcpyAsync(xin + j, &one, sizeof(double), cudaMemcpyHostToDevice, p->st));
      if (isF) RET(op_stokes(p, l, 0, false, xin, nullptr, yout, 0.0));
      else RET(op_poisson(p, l, 0, xin, nullptr, yout, 0.0));
      CU(cudaMemcpyAsync(col.data(), yout, m * sizeof(double), cudaMemcpyDeviceToHost, p->st));
      CU(cudaStreamSynchronize(p->st));
      for (int i = 0; i < m; ++i) M[(size_t)i * m + j] = col[i];
    }
    if (!isF) {
      // GtG is singular (constants): pseudo-inverse = (M + e e^T)^-1 - e e^T, e = 1/sqrt(m)
      const double ee = 1.0 / m;
      for (size_t i = 0; i < M.size(); ++i) M[i] += ee;
      if (!invert_dense(M, m)) return set_err(MPBP_E_STATE, "coarse GtG + ee^T is singular");
      for (size_t i = 0; i < M.size(); ++i) M[i] -= ee;
    } else if (!invert_dense(M, m)) {
      return set_err(MPBP_E_STATE, "coarse F is singular");
    }
    CU(cudaMemcpyAsync(isF ? p->FinvT : p->PinvT, M.data(), M.size() * sizeof(double), cudaMemcpyHostToDevice, p->st));
    CU(cudaStreamSynchronize(p->st));
  }
  return 0;
}

// Peer-memory halo exchange setup: every rank cudaMallocs a small comm buffer, the cudaIpc handles are
// all-gathered with NCCL and the two ring neighbours' buffers are mapped into this process.
static int setup_p2p(mpbp_plan* p) {
  const int P = p->nranks, prev = (p->rank + P - 1) % P, next = (p->rank + 1) % P;
  p->comm_area = (size_t)5 * p->lev[0].n;
  const size_t halo_bytes = (comm_halo_bytes(p->comm_area) + 255) & ~size_t(255);
  const size_t bytes = halo_bytes + kRedBytes;
  CU(cudaMalloc(&p->comm_local, bytes));
  CU(cudaMemset(p->comm_local, 0, bytes));
  cudaIpcMemHandle_t mine;
  CU(cudaIpcGetMemHandle(&mine, p->comm_local));
  char* dh = nullptr;
  CU(cudaMalloc(&dh, (size_t)(P + 1) * sizeof(mine)));
  CU(cudaMemcpy(dh + (size_t)P * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
  NC(ncclAllGather(dh + (size_t)P * sizeof(mine), dh, sizeof(mine), ncclChar, p->comm, nullptr));
  CU(cudaStreamSynchronize(nullptr));
  std::vector<cudaIpcMemHandle_t> all(P);
  CU(cudaMemcpy(all.data(), dh, (size_t)P * sizeof(mine), cudaMemcpyDeviceToHost));
  CU(cudaFree(dh));
  // every rank's buffer is mapped: the ring neighbours' for the halo rows, all of them for the fused all-reduce
  p->comm_all.assign(P, nullptr);
  for (int r = 0; r < P; ++r) {
    if (r == p->rank) {
      p->comm_all[r] = p->comm_local;
      continue;
    }
    void* pp = nullptr;
    CU(cudaIpcOpenMemHandle(&pp, all[r], cudaIpcMemLazyEnablePeerAccess));
    p->comm_all[r] = (char*)pp;
  }
  p->comm_prev = p->comm_all[prev];
  p->comm_next = p->comm_all[next];
  if (P <= kRedMaxRanks && !getenv("MPBP_NCCL_ALLREDUCE")) {
    p->red.rank = p->rank;
    p->red.nranks = P;
    p->red.rseq = p->dseq + 8;
    for (int r = 0; r < P; ++r) p->red.peers[r] = p->comm_all[r] + halo_bytes;
  }
  // nobody may push before every rank has zeroed and mapped its buffers
  double* tok = nullptr;
  CU(cudaMalloc(&tok, sizeof(double)));
  CU(cudaMemset(tok, 0, sizeof(double)));
  NC(ncclAllReduce(tok, tok, 1, ncclDouble, ncclSum, p->comm, nullptr));
  CU(cudaStreamSynchronize(nullptr));
  CU(cudaFree(tok));
  p->p2p = true;
  return 0;
}

extern "C" int mpbp_plan_create(mpbp_plan** out, const mpbp_config* cfg) {
  if (!out) return set_err(MPBP_E_ARG, "null plan pointer");
  *out = nullptr;
  RET(check_cfg(cfg));
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return set_err(MPBP_E_STATE, "no CUDA device: libmpbp has no CPU fallback");
  mpbp_plan* p = new mpbp_plan();
  p->cfg = *cfg;
  p->cfg.theta_host = nullptr;
  p->cfg.nccl_unique_id = nullptr;
  p->nranks = cfg->nranks;
  p->rank = cfg->rank;
  int rc = build_levels_shape(*cfg, p->lev, p->first_repl);
  if (rc) { delete p; return rc; }
  Bump B;
  B.dry = true;
  carve(p, B);
  const size_t need = B.off + 256;
  char* base = nullptr;
  if (cfg->workspace) {
    if (cfg->workspace_bytes < need) {
      delete p;
      return set_err(MPBP_E_NOMEM, "workspace too small: %zu < %zu", cfg->workspace_bytes, need);
    }
    base = (char*)cfg->workspace;
  } else {
    cudaError_t e = cudaMalloc(&p->owned, need);
    if (e != cudaSuccess) {
      delete p;
      return set_err((int)e, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
    }
    base = (char*)p->owned;
  }
  B = Bump();
  B.base = (char*)(((uintptr_t)base + 255) & ~uintptr_t(255));
  B.dry = false;
  carve(p, B);
  {
    cudaError_t e = cudaMallocHost(&p->hscal, kScal * sizeof(double));
    if (e != cudaSuccess) { mpbp_plan_destroy(p); return set_err((int)e, "cudaMallocHost failed"); }
  }
  auto fail = [&](int code) { mpbp_plan_destroy(p); return code; };
  if (cudaStreamCreateWithFlags(&p->own, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming) != cudaSuccess)
    return fail(set_err(999, "stream/event creation failed"));
  p->st = p->own;
  if (const char* e = getenv("MPBP_GRAPH")) p->use_graph = atoi(e) != 0;
  if (const char* e = getenv("MPBP_JAC_MINB")) p->jac_minb = atoi(e);
  if (const char* e = getenv("MPBP_FUSED_MGS")) p->fused_mgs = atoi(e) != 0;
  if (const char* e = getenv("MPBP_ORTH")) p->lowsync = strcmp(e, "mgs") != 0;
  if (const char* e = getenv("MPBP_FUSE")) p->fuse = atoi(e);
  if (const char* e = getenv("MPBP_COARSE")) p->coarse_n = atoi(e);
  if (const char* e = getenv("MPBP_CELL")) p->cell_n = atoi(e);  // 0: marching kernels on every level
  if (const char* e = getenv("MPBP_PUSH_FUSED")) p->push_fused = atoi(e) != 0;
  if (cudaMemsetAsync(p->counter, 0, 64 * sizeof(unsigned int), nullptr) != cudaSuccess ||
      cudaMemsetAsync(p->dseq, 0, 16 * sizeof(unsigned long long), nullptr) != cudaSuccess)
    return fail(set_err(999, "memset failed"));

  if (p->nranks > 1) {
    ncclUniqueId id;
    memcpy(&id, cfg->nccl_unique_id, sizeof(id));
    ncclResult_t e = ncclCommInitRank(&p->comm, p->nranks, id, p->rank);
    if (e != ncclSuccess) return fail(set_err(1000 + (int)e, "ncclCommInitRank: %s", ncclGetErrorString(e)));
    const char* mode = getenv("MPBP_HALO");
    if (!(mode && strcmp(mode, "nccl") == 0)) {
      rc = setup_p2p(p);
      if (rc) return fail(rc);
    }
  }

  // ---- coefficient fields on the host: theta_n per level (4-cell averages), mass-term tables ----
  const int n0 = cfg->n;
  const double PI = 3.141592653589793;
  {
    std::vector<double> sxf(n0), sxc(n0), syf(n0), syc(n0);
    const double h = 1.0 / n0;
    for (int i = 0; i < n0; ++i) {
      sxf[i] = std::sin(2 * PI * (i * h));            // x = c h            (preconditioner.py:325)
      sxc[i] = std::sin(2 * PI * ((i + 0.5) * h));    // x = (c+1/2) h      (preconditioner.py:326)
      syf[i] = std::sin(2 * PI * (-i * h));           // y = -r h
      syc[i] = std::sin(2 * PI * (-(i + 0.5) * h));   // y = -(r+1/2) h
    }
    std::vector<double> th((size_t)n0 * n0);
    if (cfg->theta_host) {
      memcpy(th.data(), cfg->theta_host, th.size() * sizeof(double));
    } else {
      for (int r = 0; r < n0; ++r)
        for (int c = 0; c < n0; ++c) th[(size_t)r * n0 + c] = 0.25 * sxc[c] * syc[r] + 0.5;  // preconditioner.py:10
    }
    cudaMemcpy(p->sxf, sxf.data(), n0 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(p->sxc, sxc.data(), n0 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(p->syf, syf.data(), n0 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(p->syc, syc.data(), n0 * sizeof(double), cudaMemcpyHostToDevice);
    for (size_t l = 0; l < p->lev.size(); ++l) {
      Level& v = p->lev[l];
      const int n = v.n;
      if (l > 0) {
        std::vector<double> tc((size_t)n * n);
        const int nf = 2 * n;
        for (int R = 0; R < n; ++R)
          for (int C = 0; C < n; ++C) {
            const double* a = &th[(size_t)(2 * R) * nf + 2 * C];
            tc[(size_t)R * n + C] = 0.25 * (a[0] + a[nf] + a[1] + a[nf + 1]);
          }
        th.swap(tc);
      }
      // upload rows row0-1 .. row0+rows (periodic)
      std::vector<double> pad((size_t)(v.rows + 2) * n);
      for (int r = -1; r <= v.rows; ++r) {
        const int gr = ((v.row0 + r) % n + n) % n;
        memcpy(&pad[(size_t)(r + 1) * n], &th[(size_t)gr * n], n * sizeof(double));
      }
      if (cudaMemcpy(v.th, pad.data(), pad.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(set_err(999, "theta upload failed: %s", cudaGetErrorString(cudaGetLastError())));
      const double h = 1.0 / n;
      Phys& ph = v.ph;
      ph.xi = cfg->xi;
      ph.c = cfg->c;
      ph.d_u = cfg->d_u;
      ph.kap_n = cfg->d_u * cfg->eta_n / (h * h);
      ph.kap_s = cfg->d_u * cfg->eta_s / (h * h);
      ph.dp_h = cfg->d_p / h;
      ph.ddiv_h = cfg->d_div / h;
      ph.inv_h = 1.0 / h;
      ph.dp_h2 = cfg->d_p / (h * h);
      ph.mass_mode = (l == 0 && !cfg->theta_host) ? 1 : 0;
      ph.sxf = p->sxf;
      ph.sxc = p->sxc;
      ph.syf = p->syf;
      ph.syc = p->syc;
      v.geo.n = n;
      v.geo.rows = v.rows;
      v.geo.row0 = v.row0;
      const int gx = (n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps);
      int sms_ = 148, dev_ = 0;
      cudaGetDevice(&dev_);
      cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev_);
      v.geo.pf = 3;  // measured best on B200 at 4096^2 (profiles/r1_tuning.txt)
      if (const char* e = getenv("MPBP_PF")) v.geo.pf = std::max(0, std::min(atoi(e), 64));
      v.geo.pfc = 1;
      if (const char* e = getenv("MPBP_PFC")) v.geo.pfc = atoi(e) != 0;
      v.geo4 = v.geo;
      v.geoL = v.geo;
      v.geoR = v.geo;
      const int gxR = (n + WarpTile<2>::cols * kBlockWarps - 1) / (WarpTile<2>::cols * kBlockWarps);
      // single-wave decomposition (set_strips): measured SLOWER than uniform 32-row strips on one B200 at 4096^2
      // (Jacobi sweep 0.344 ms vs 0.290 ms, profiles/r2_tuning.txt) -- off unless MPBP_WAVE=1
      const bool wave = getenv("MPBP_WAVE") && atoi(getenv("MPBP_WAVE")) != 0;
      if (wave) {
        set_strips(v.geo, gx, v.rows, 5 * sms_);
        set_strips(v.geo4, gx, v.rows, 4 * sms_);
        set_strips(v.geoR, gxR, v.rows, 5 * sms_);
      } else {
        // every variant gets the strip height that fills whole waves of ITS resident-block count and grid width:
        // on the slabs of 2-8 GPUs a launch is one or two waves, and a shared height cost the 4-blocks/SM and the
        // 28-column variants a nearly empty extra wave
        v.geo.rs = choose_rs(gx, v.rows, 5 * sms_);
        v.geo4.rs = choose_rs(gx, v.rows, 4 * sms_);
        v.geoR.rs = choose_rs(gxR, v.rows, 5 * sms_);
        for (Geo* g : {&v.geo, &v.geo4, &v.geoR})
          if (!(v.rows & 1) && (g->rs & 1)) g->rs += 1;
      }
      if (const char* e = getenv("MPBP_RS")) {
        v.geo.rs = std::max(1, std::min(atoi(e), v.rows));
        if (!(v.rows & 1) && (v.geo.rs & 1)) v.geo.rs += 1;
        v.geo4.rs = v.geoR.rs = v.geo.rs;
        v.geo.re = v.geo4.re = v.geoR.re = 0;
      }
      {  // light kernels (16 resident blocks/SM)
        int rs = 32;
        while (rs > 4 && gx * ((v.rows + rs - 1) / rs) < 4 * sms_) rs /= 2;
        v.geoL.rs = std::max(1, std::min(rs, v.rows));
        v.geoL.re = 0;
        if (wave && gx * ((v.rows + 31) / 32) >= 16 * sms_) set_strips(v.geoL, gx, v.rows, 16 * sms_);
      }
      if (const char* e = getenv("MPBP_RSL")) {
        v.geoL.rs = std::max(1, std::min(atoi(e), v.rows));
        v.geoL.re = 0;
      }
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p->red_blocks = std::min(kMaxRedBlocks, sms * 8);
  if (const char* rb = getenv("MPBP_RED_BLOCKS")) {  // testing knob: changes the (deterministic) summation order
    const int v = atoi(rb);
    if (v >= 1 && v <= kMaxRedBlocks) p->red_blocks = v;
  }
  if (!cfg->operators_only) {
    rc = build_coarse_inverses(p);
    if (rc) return fail(rc);
    // omega / diag(F) per whole-grid level: one first-sweep kernel applied to a vector of ones
    for (size_t l = 0; l + 1 < p->lev.size(); ++l) {
      Level& v = p->lev[l];
      if (!v.wdF) continue;
      const size_t len = 4 * v.fs();
      k_fill<<<ew_blocks(len), 256, 0, p->st>>>(v.tF, 1.0, len);
      p->launches++;
      rc = op_jacobi0_F(p, (int)l, v.tF, v.wdF, cfg->omega);
      if (rc) return fail(rc);
      if (v.dist) {
        rc = fetch_static_halo(p, v, v.wdF, 4, v.wdh);
        if (rc) return fail(rc);
      }
      if (v.wdP) {  // 1 / diag(GtG): the first pressure sweep (omega = 1) applied to ones -- the sweeps' own reciprocal
        k_fill<<<ew_blocks(v.fs()), 256, 0, p->st>>>(v.tP, 1.0, v.fs());
        p->launches++;
        rc = op_poisson(p, (int)l, 3, nullptr, v.tP, v.wdP, 1.0);
        if (rc) return fail(rc);
      }
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(set_err(999, "plan setup failed: %s", cudaGetErrorString(cudaGetLastError())));
  p->launches = 0;
  *out = p;
  return 0;
}

extern "C" int mpbp_plan_destroy(mpbp_plan* p) {
  if (!p) return 0;
  cudaDeviceSynchronize();
  // graphs first: captured NCCL collectives hold references on the communicator, and ncclCommDestroy
  // waits for them
  if (p->gexec_apply) cudaGraphExecDestroy(p->gexec_apply);
  cudaDeviceSynchronize();
  for (size_t r = 0; r < p->comm_all.size(); ++r)
    if (p->comm_all[r] && p->comm_all[r] != p->comm_local) cudaIpcCloseMemHandle(p->comm_all[r]);
  // (destroy is not collective: callers finish all solves on every rank before dropping their plans;
  //  ncclCommAbort never waits for the other ranks)
  if (p->comm) ncclCommAbort(p->comm);
  if (p->ev_in) cudaEventDestroy(p->ev_in);
  if (p->ev_out) cudaEventDestroy(p->ev_out);
  if (p->own) cudaStreamDestroy(p->own);
  if (p->comm_local) cudaFree(p->comm_local);
  if (p->owned) cudaFree(p->owned);
  if (p->kry_owned) cudaFree(p->kry_owned);
  for (int i = 0; i < 2; ++i)
    if (p->hostbuf_dev[i]) cudaFree(p->hostbuf_dev[i]);
  if (p->hscal) cudaFreeHost(p->hscal);
  delete p;
  return 0;
}

extern "C" const char* mpbp_last_error_string(void) { return g_err; }

#ifndef MPBP_BUILD_ID_STR
#define MPBP_BUILD_ID_STR "unknown"
#endif
// hash of the sources this binary was compiled from (written by _build.py); the marker makes it greppable in the file
static const char g_build_id[] = "MPBP_BUILD_ID=" MPBP_BUILD_ID_STR;
extern "C" const char* mpbp_build_id(void) { return g_build_id + 14; }

extern "C" int mpbp_nccl_unique_id(void* out128) {
  if (!out128) return set_err(MPBP_E_ARG, "null output");
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, sizeof(id));
  return 0;
}

extern "C" int mpbp_plan_rows_local(const mpbp_plan* p) { return p ? p->lev[0].rows : MPBP_E_ARG; }
extern "C" int mpbp_plan_num_levels(const mpbp_plan* p) { return p ? (int)p->lev.size() : MPBP_E_ARG; }
extern "C" long long mpbp_plan_launches(const mpbp_plan* p) { return p ? p->launches : -1; }

// ---------------------------------------------------------------------------------------------
// exported operator applies
// ---------------------------------------------------------------------------------------------
// every export runs on the plan's own stream between two event fences against the caller's stream
static int enter(mpbp_plan* p, void* stream) {
  p->ust = (cudaStream_t)stream;
  p->st = p->own;
  CU(cudaEventRecord(p->ev_in, p->ust));
  CU(cudaStreamWaitEvent(p->own, p->ev_in, 0));
  return 0;
}
static int leave(mpbp_plan* p, int rc) {
  cudaError_t e = cudaEventRecord(p->ev_out, p->own);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(p->ust, p->ev_out, 0);
  if (rc == 0 && e != cudaSuccess) return set_err((int)e, "stream fence failed: %s", cudaGetErrorString(e));
  return rc;
}
#define ENTER(p, stream)                               \
  if (!(p)) return set_err(MPBP_E_ARG, "null plan");   \
  RET(enter((p), (stream)));

extern "C" int mpbp_apply_A(mpbp_plan* p, const double* x, double* y, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!x || !y || x == y) return set_err(MPBP_E_ARG, "apply_A: bad pointers");
  return op_stokes(p, 0, 0, true, x, nullptr, y, 0.0);
  };
  return leave(p, body());
}
extern "C" int mpbp_apply_F(mpbp_plan* p, const double* x, double* y, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!x || !y || x == y) return set_err(MPBP_E_ARG, "apply_F: bad pointers");
  return op_stokes(p, 0, 0, false, x, nullptr, y, 0.0);
  };
  return leave(p, body());
}
extern "C" int mpbp_apply_G(mpbp_plan* p, const double* pr, double* y, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!pr || !y) return set_err(MPBP_E_ARG, "apply_G: bad pointers");
  return op_grad(p, 0, pr, y);
  };
  return leave(p, body());
}
extern "C" int mpbp_apply_D(mpbp_plan* p, const double* w, const double* add, double* r, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!w || !r) return set_err(MPBP_E_ARG, "apply_D: bad pointers");
  return op_div(p, 0, w, add, r, 1.0);
  };
  return leave(p, body());
}
extern "C" int mpbp_apply_GtG(mpbp_plan* p, const double* pr, double* y, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!pr || !y || pr == y) return set_err(MPBP_E_ARG, "apply_GtG: bad pointers");
  return op_poisson(p, 0, 0, pr, nullptr, y, 0.0);
  };
  return leave(p, body());
}
extern "C" int mpbp_apply_GtFG(mpbp_plan* p, const double* pr, double* y, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!pr || !y) return set_err(MPBP_E_ARG, "apply_GtFG: bad pointers");
  RET(op_grad(p, 0, pr, p->g));
  RET(op_stokes(p, 0, 0, false, p->g, nullptr, p->t2, 0.0));
  return op_div(p, 0, p->t2, nullptr, y, -1.0);
  };
  return leave(p, body());
}
extern "C" int mpbp_jacobi_F(mpbp_plan* p, const double* b, double* x, int sweeps, double omega, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || sweeps < 0) return set_err(MPBP_E_ARG, "jacobi_F: bad arguments");
  return jacobi_sweeps(p, 0, true, b, x, p->lev[0].tF, sweeps, omega);
  };
  return leave(p, body());
}
extern "C" int mpbp_jacobi_P(mpbp_plan* p, const double* b, double* x, int sweeps, double omega, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || sweeps < 0) return set_err(MPBP_E_ARG, "jacobi_P: bad arguments");
  return jacobi_sweeps(p, 0, false, b, x, p->lev[0].tP, sweeps, omega);
  };
  return leave(p, body());
}
extern "C" int mpbp_vcycle_F(mpbp_plan* p, const double* b, double* x, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || b == x) return set_err(MPBP_E_ARG, "vcycle_F: bad pointers");
  return vcycle(p, 0, true, b, x);
  };
  return leave(p, body());
}
extern "C" int mpbp_vcycle_P(mpbp_plan* p, const double* b, double* x, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || b == x) return set_err(MPBP_E_ARG, "vcycle_P: bad pointers");
  return vcycle(p, 0, false, b, x);
  };
  return leave(p, body());
}
extern "C" int mpbp_solve_F(mpbp_plan* p, const double* b, double* x, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || b == x) return set_err(MPBP_E_ARG, "solve_F: bad pointers");
  return sub_solve(p, true, b, x);
  };
  return leave(p, body());
}
extern "C" int mpbp_solve_P(mpbp_plan* p, const double* b, double* x, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || b == x) return set_err(MPBP_E_ARG, "solve_P: bad pointers");
  return sub_solve(p, false, b, x);
  };
  return leave(p, body());
}
extern "C" int mpbp_precond_apply(mpbp_plan* p, const double* v, double* z, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!v || !z || v == z) return set_err(MPBP_E_ARG, "precond_apply: bad pointers");
  return precond_apply(p, v, z);
  };
  return leave(p, body());
}

// Communication probe (bench.py --workload apply8192): average time of one level-0 halo exchange of a 5-field vector
// (push kernel + a consumer kernel that only fetches the rows) and of one scalar all-reduce (a one-block reduction
// kernel with the fused peer-memory all-reduce), in microseconds.  The repetitions are replayed from a CUDA graph --
// the regime the preconditioner apply runs in -- and timed with CUDA events.
extern "C" int mpbp_comm_probe(mpbp_plan* p, int reps, double* halo_us, double* allreduce_us, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
    if (!halo_us || !allreduce_us || reps < 1) return set_err(MPBP_E_ARG, "comm_probe: bad arguments");
    *halo_us = *allreduce_us = 0.0;
    if (p->nranks == 1) return 0;
    Level& v = p->lev[0];
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int what = 0; what < 2; ++what) {
      CU(cudaStreamBeginCapture(p->st, cudaStreamCaptureModeThreadLocal));
      int rc = 0;
      for (int i = 0; i < reps && rc == 0; ++i) {
        if (what == 0) {
          VecIn in{};
          p->pending_push = nullptr;
          rc = make_view(p, v, p->vin, 5, in);
          if (rc == 0 && p->p2p) {
            k_halo_consume<<<(v.n + 255) / 256, 256, 0, p->st>>>(in, 5, v.n, p->scal + kScal - 2);
            p->launches++;
          }
        } else {
          k_sum<<<1, kRedThreads, 0, p->st>>>(p->vin, 1024, p->partial, p->counter, p->scal + kScal - 2, p->red);
          p->launches++;
          if (p->red.nranks <= 1) rc = allreduce_scal(p, p->scal + kScal - 2, 1);
        }
      }
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(p->st, &g);
      if (rc != 0) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      CU(ce);
      cudaGraphExec_t ge = nullptr;
      CU(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      float ms = 0.f;
      for (int pass = 0; pass < 2; ++pass) {  // pass 0: warm-up
        CU(cudaEventRecord(e0, p->st));
        CU(cudaGraphLaunch(ge, p->st));
        CU(cudaEventRecord(e1, p->st));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
      }
      cudaGraphExecDestroy(ge);
      (what == 0 ? *halo_us : *allreduce_us) = 1e3 * ms / reps;
    }
    p->pending_push = nullptr;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
  };
  return leave(p, body());
}

static int ensure_hostbuf(mpbp_plan* p) {
  const size_t bytes = 5 * p->lev[0].fs() * sizeof(double);
  for (int i = 0; i < 2; ++i)
    if (!p->hostbuf_dev[i]) CU(cudaMalloc(&p->hostbuf_dev[i], bytes));
  return 0;
}
extern "C" int mpbp_precond_apply_host(mpbp_plan* p, const double* v_host, double* z_host, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!v_host || !z_host) return set_err(MPBP_E_ARG, "precond_apply_host: bad pointers");
  RET(ensure_hostbuf(p));
  const size_t bytes = 5 * p->lev[0].fs() * sizeof(double);
  CU(cudaMemcpyAsync(p->hostbuf_dev[0], v_host, bytes, cudaMemcpyHostToDevice, p->st));
  RET(precond_apply(p, p->hostbuf_dev[0], p->hostbuf_dev[1]));
  CU(cudaMemcpyAsync(z_host, p->hostbuf_dev[1], bytes, cudaMemcpyDeviceToHost, p->st));
  CU(cudaStreamSynchronize(p->st));
  return 0;
  };
  return leave(p, body());
}

// algorithmic bytes (SURVEY 8d accounting: every input read once, every output written once, one
// coefficient field per stencil kernel) of one preconditioner apply under the current configuration.
// Only the passes the fused kernels actually make are counted: no copies, no materialised V-cycle outputs.
static double vcycle_bytes(const mpbp_plan* p, int l, bool isF, bool with_ep) {
  const mpbp_config& c = p->cfg;
  const int L = (int)p->lev.size();
  const double N = (double)p->lev[l].fs();
  if (l == L - 1) return 0.0;  // dense coarse solve: negligible
  const Level& v = p->lev[l];
  const bool even = !(v.rows & 1) && !(v.geo.rs & 1) && !(v.geo4.rs & 1) && !(v.geoR.rs & 1);
  double by = 0.0;
  // the epilogue replaces the write of z (4N or N doubles) by read d, x + write d, x (upper bound: middle cycles)
  const double ep_extra = with_ep ? (isF ? 4 : 1) * 8.0 * N * 3.0 : 0.0;
  if (isF) {
    const bool dist_ok = !v.dist || (p->p2p && p->push_fused);
    const bool fuse_pre = (p->fuse & 1) && dist_ok && c.nu1 == 2 && v.wdF != nullptr;
    const bool fuse_rr = (p->fuse & 4) && even && (!v.dist || (dist_ok && v.rows >= 4));
    const bool fuse_post = (p->fuse & 2) && dist_ok && even && c.nu2 >= 1;
    by += fuse_pre ? 104 * N : 72 * N + (c.nu1 - 1) * 104.0 * N;   // pre-smoothing (pair: reads b, omega/diag, theta)
    by += fuse_rr ? 80 * N : (104 + 40) * N;                         // residual (+ restriction: writes N instead of 4N)
    by += fuse_post ? 112 * N + (c.nu2 - 1) * 104.0 * N              // prolongation fused into the first post-sweep
                    : 72 * N + c.nu2 * 104.0 * N;                    // x += P e ; nu2 sweeps
  } else {
    const bool fuseP = (p->fuse & 8) && !v.dist && !(v.rows & 1) && !(v.geoL.rs & 1) && v.geoL.re == 0;
    by += (fuseP && c.nu1 == 2 && v.wdP) ? 32 * N : 24 * N + (c.nu1 - 1) * 32.0 * N;   // pair: b, omega/diag, theta, write
    by += fuseP ? 26 * N : 32 * N + 10 * N;                                            // residual (+ restriction)
    by += (fuseP && c.nu2 >= 1) ? 34 * N + (c.nu2 - 1) * 32.0 * N : 18 * N + c.nu2 * 32.0 * N;
  }
  return by + ep_extra + vcycle_bytes(p, l + 1, isF, false);
}
static double solve_bytes(const mpbp_plan* p, bool isF) {
  const mpbp_config& c = p->cfg;
  const double N = (double)p->lev[0].fs();
  const int kind = isF ? c.F_kind : c.P_kind;
  double by = 0.0;
  if (kind == MPBP_SUB_JACOBI) {
    const int s = isF ? c.F_sweeps : c.P_sweeps;
    by = isF ? (72 * N + (s - 1) * 104.0 * N) : (24 * N + (s - 1) * 32.0 * N);
  } else {
    const int k = isF ? c.F_cycles : c.P_cycles;
    const double vec = isF ? 32 * N : 8 * N;  // one vector pass
    by = k * vcycle_bytes(p, 0, isF, true);
    by -= 2 * vec;                                            // first cycle reads neither d nor x
    by -= (c.cheb ? 1 : 2 * k - 1) * vec;                     // last cycle does not write d (plain iteration: never d)
    by += (k - 1) * (isF ? 104 * N : 32 * N);                 // residual before every extra cycle
  }
  if (!isF && c.project) by += 3 * 8 * N;
  return by;
}
extern "C" int mpbp_precond_bytes(const mpbp_plan* p, double* bytes) {
  if (!p || !bytes) return set_err(MPBP_E_ARG, "null argument");
  const double N = (double)p->lev[0].fs();
  double by = 2 * solve_bytes(p, true) + 2 * solve_bytes(p, false);
  by += 56 * N;             // K5  r = D w + v_p
  by += 48 * N + 72 * N + 48 * N;  // K7 -> K2 -> K5 chain for GtFG
  by += 48 * N;             // K7  G x_p
  by += 80 * N;             // copy-in of v (the graph's fixed input address)
  by += 112 * N;            // K8  z = [w - y ; x_p]
  *bytes = by;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// exported vector kernels
// ---------------------------------------------------------------------------------------------
extern "C" int mpbp_dot(mpbp_plan* p, const double* x, const double* y, size_t len, double* result, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!x || !y || !result) return set_err(MPBP_E_ARG, "dot: bad pointers");
  RET(v_dot(p, x, y, len, p->scal));
  RET(fetch_scal(p, p->scal, 1, p->hscal));
  *result = p->hscal[0];
  return 0;
  };
  return leave(p, body());
}
extern "C" int mpbp_nrm2(mpbp_plan* p, const double* x, size_t len, double* result, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!x || !result) return set_err(MPBP_E_ARG, "nrm2: bad pointers");
  RET(v_nrm2(p, x, len, p->scal));
  RET(fetch_scal(p, p->scal, 1, p->hscal));
  *result = p->hscal[0];
  return 0;
  };
  return leave(p, body());
}
extern "C" int mpbp_axpy(mpbp_plan* p, double alpha, const double* x, double* y, size_t len, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!x || !y) return set_err(MPBP_E_ARG, "axpy: bad pointers");
  return v_multi_axpy(p, x, 0, 1, &alpha, y, len);
  };
  return leave(p, body());
}
extern "C" int mpbp_multi_dot(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* w, size_t len,
                              double* out, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!V || !w || !out || nvec < 1 || nvec > kScal - 8) return set_err(MPBP_E_ARG, "multi_dot: bad arguments");
  RET(v_multi_dot(p, V, ld, nvec, w, len, p->scal, false));
  RET(fetch_scal(p, p->scal, nvec, p->hscal));
  memcpy(out, p->hscal, nvec * sizeof(double));
  return 0;
  };
  return leave(p, body());
}
extern "C" int mpbp_multi_axpy(mpbp_plan* p, const double* V, size_t ld, int nvec, const double* alpha, double* w,
                               size_t len, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!V || !w || !alpha || nvec < 1) return set_err(MPBP_E_ARG, "multi_axpy: bad arguments");
  return v_multi_axpy(p, V, ld, nvec, alpha, w, len);
  };
  return leave(p, body());
}
extern "C" int mpbp_wnorms(mpbp_plan* p, const double* a, const double* b, size_t len, double w, double* out3,
                           void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!a || !b || !out3) return set_err(MPBP_E_ARG, "wnorms: bad pointers");
  const int blocks = (int)std::min<size_t>((len + kRedThreads - 1) / kRedThreads, (size_t)p->red_blocks);
  k_diffnorms<<<blocks, kRedThreads, 0, p->st>>>(a, b, len, p->partial, p->counter, p->scal);
  LAUNCH_CHECK(p);
  if (p->nranks > 1) {
    NC(ncclAllReduce(p->scal, p->scal, 2, ncclDouble, ncclSum, p->comm, p->st));
    NC(ncclAllReduce(p->scal + 2, p->scal + 2, 1, ncclDouble, ncclMax, p->comm, p->st));
  }
  RET(fetch_scal(p, p->scal, 3, p->hscal));
  out3[0] = w * p->hscal[0];        // weighted_L1, utils.py:11-14
  out3[1] = sqrt(w * p->hscal[1]);  // weighted_L2, utils.py:7-9
  out3[2] = p->hscal[2];            // max_norm,    utils.py:16-17
  return 0;
  };
  return leave(p, body());
}
extern "C" int mpbp_fill_manufactured(mpbp_plan* p, double* u_vec, double* b_vec, double b_p_sign, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  const Level& v = p->lev[0];
  const mpbp_config& c = p->cfg;
  const dim3 block(128), grid((v.n + 127) / 128, v.rows);
  k_fill_manufactured<<<grid, block, 0, p->st>>>(u_vec, b_vec, v.n, v.rows, v.row0, c.c, c.d_u, c.xi, c.eta_n, c.eta_s,
                                                  b_p_sign);
  LAUNCH_CHECK(p);
  return 0;
  };
  return leave(p, body());
}

// ---------------------------------------------------------------------------------------------
// Krylov
// ---------------------------------------------------------------------------------------------
extern "C" int mpbp_gmres_opts_default(mpbp_gmres_opts* o) {
  if (!o) return set_err(MPBP_E_ARG, "null opts");
  memset(o, 0, sizeof(*o));
  o->rtol = 1e-8;     // solve.py:285
  o->restart = 20;    // scipy default
  o->maxiter = 150;   // solve.py:285
  o->side = MPBP_SIDE_RIGHT;
  o->use_precond = 1;
  return 0;
}
static size_t kry_vectors(const mpbp_gmres_opts* o) {
  // V: restart+1, w, tmp ; right: + Z: restart
  return (size_t)(o->restart + 1) + 2 + (o->side == MPBP_SIDE_RIGHT ? (size_t)o->restart : 0);
}
extern "C" int mpbp_gmres_workspace_bytes(const mpbp_plan* p, const mpbp_gmres_opts* o, size_t* bytes) {
  if (!p || !o || !bytes) return set_err(MPBP_E_ARG, "null argument");
  if (o->restart < 1) return set_err(MPBP_E_ARG, "restart must be >= 1");
  *bytes = kry_vectors(o) * 5 * p->lev[0].fs() * sizeof(double) + 256;
  return 0;
}

// LAPACK dlartg: [c s; -s c] [f; g] = [r; 0]
static void lartg(double f, double g, double& c, double& s, double& r) {
  if (g == 0.0) { c = 1.0; s = 0.0; r = f; return; }
  if (f == 0.0) { c = 0.0; s = (g < 0 ? -1.0 : 1.0); r = std::fabs(g); return; }
  const double d = std::sqrt(f * f + g * g);
  c = std::fabs(f) / d;
  r = (f < 0 ? -d : d);
  s = g / r;
}

static int psolve(mpbp_plan* p, bool use_pc, const double* in, double* out, size_t len) {
  if (use_pc) return precond_apply(p, in, out);
  return v_copy(p, in, out, len);
}

// scipy.sparse.linalg.gmres (scipy/sparse/linalg/_isolve/iterative.py) restated: left
// preconditioning, modified Gram-Schmidt, Givens rotations, restart, presid/ptol tolerance control.
static int gmres_left(mpbp_plan* p, const double* b, double* x, const mpbp_gmres_opts* o, double* ws, double* hist,
                      int hist_cap, int* n_iters, int* info) {
  const size_t len = 5 * p->lev[0].fs();
  const int m = o->restart;
  double* V = ws;                       // (m+1) x len
  double* w = V + (size_t)(m + 1) * len;
  double* tmp = w + len;
  double* hs = p->hscal;
  double* ds = p->scal;
  const double eps = std::numeric_limits<double>::epsilon();
  const bool pc = o->use_precond != 0;
  const bool forced = o->force_iters > 0;
  int inner_iter = 0;
  *n_iters = 0;
  *info = 0;

  RET(v_nrm2(p, b, len, ds));
  RET(fetch_scal(p, ds, 1, hs));
  const double bnrm2 = hs[0];
  const double atol = std::max(0.0, o->rtol * bnrm2);
  if (bnrm2 == 0.0) { RET(v_copy(p, b, x, len)); return 0; }
  if (!o->x0_nonzero) CU(cudaMemsetAsync(x, 0, len * sizeof(double), p->st));
  RET(psolve(p, pc, b, tmp, len));
  RET(v_nrm2(p, tmp, len, ds));
  RET(fetch_scal(p, ds, 1, hs));
  const double Mb_nrm2 = hs[0];
  double ptol_max_factor = 1.0;
  double ptol = Mb_nrm2 * std::min(ptol_max_factor, atol / bnrm2);
  double presid = 0.0, rnorm = 0.0;
  std::vector<double> H((size_t)m * (m + 1), 0.0), giv((size_t)m * 2, 0.0), S(m + 1), yv(m);
  double* r = tmp;  // residual lives in tmp between outer iterations
  p->last_H.assign((size_t)(m + 1) * m, 0.0);
  p->last_m = m;
  p->last_k = 0;

  for (int iteration = 0; iteration < o->maxiter; ++iteration) {
    if (iteration == 0) {
      if (o->x0_nonzero) {
        RET(op_stokes(p, 0, 0, true, x, nullptr, w, 0.0));
        RET(v_axpby(p, 1.0, b, -1.0, w, r, len));
      } else {
        RET(v_copy(p, b, r, len));
      }
      RET(v_nrm2(p, r, len, ds));
      RET(fetch_scal(p, ds, 1, hs));
      if (hs[0] < atol && !forced) return 0;
    }
    RET(psolve(p, pc, r, V, len));
    RET(v_nrm2(p, V, len, ds));
    k_scale_dev<<<ew_blocks(len), 256, 0, p->st>>>(ds, 1, V, V, len);
    LAUNCH_CHECK(p);
    RET(fetch_scal(p, ds, 1, hs));
    std::fill(S.begin(), S.end(), 0.0);
    S[0] = hs[0];
    bool breakdown = false;
    int col = 0;
    for (col = 0; col < m; ++col) {
      double* vc = V + (size_t)col * len;
      double* vn = V + (size_t)(col + 1) * len;
      RET(op_stokes(p, 0, 0, true, vc, nullptr, tmp, 0.0));  // av = A v[col]
      RET(psolve(p, pc, tmp, w, len));                       // w = M av
      // modified Gram-Schmidt with device-resident coefficients: ds[0]=h0, ds[1+k]=h[col][k], ds[col+2]=h1
      RET(v_mgs(p, V, len, col, w, len, ds + 1, ds + col + 2, ds));
      k_scale_dev<<<ew_blocks(len), 256, 0, p->st>>>(ds + col + 2, 1, w, vn, len);
      LAUNCH_CHECK(p);
      RET(fetch_scal(p, ds, col + 3, hs));
      const double h0 = hs[0], h1 = hs[col + 2];
      double* h = &H[(size_t)col * (m + 1)];
      for (int k = 0; k <= col; ++k) h[k] = hs[1 + k];
      h[col + 1] = h1;
      if (col == 0) std::fill(p->last_H.begin(), p->last_H.end(), 0.0);
      for (int k = 0; k <= col + 1; ++k) p->last_H[(size_t)k * m + col] = hs[1 + k];
      p->last_k = col + 1;
      if (h1 <= eps * h0) {
        h[col + 1] = 0.0;
        breakdown = true;
      }
      for (int k = 0; k < col; ++k) {
        const double c = giv[2 * k], s = giv[2 * k + 1];
        const double n0 = h[k], n1 = h[k + 1];
        h[k] = c * n0 + s * n1;
        h[k + 1] = -s * n0 + c * n1;
      }
      double c, s, mag;
      lartg(h[col], h[col + 1], c, s, mag);
      giv[2 * col] = c;
      giv[2 * col + 1] = s;
      h[col] = mag;
      h[col + 1] = 0.0;
      const double t = -s * S[col];
      S[col] = c * S[col];
      S[col + 1] = t;
      presid = std::fabs(t);
      inner_iter++;
      if (hist && inner_iter <= hist_cap) hist[inner_iter - 1] = presid / bnrm2;
      if (forced) {
        if (inner_iter >= o->force_iters) break;
      } else if (presid <= ptol || breakdown) {
        break;
      }
    }
    if (col == m) col = m - 1;
    if (H[(size_t)col * (m + 1) + col] == 0.0) S[col] = 0.0;
    for (int k = 0; k <= col; ++k) yv[k] = S[k];
    for (int k = col; k > 0; --k) {
      if (yv[k] != 0.0) {
        yv[k] /= H[(size_t)k * (m + 1) + k];
        const double t = yv[k];
        for (int i = 0; i < k; ++i) yv[i] -= t * H[(size_t)k * (m + 1) + i];
      }
    }
    if (yv[0] != 0.0) yv[0] /= H[0];
    RET(v_multi_axpy(p, V, len, col + 1, yv.data(), x, len));
    RET(op_stokes(p, 0, 0, true, x, nullptr, w, 0.0));
    RET(v_axpby(p, 1.0, b, -1.0, w, r, len));
    RET(v_nrm2(p, r, len, ds));
    RET(fetch_scal(p, ds, 1, hs));
    rnorm = hs[0];
    if (forced) {
      if (inner_iter >= o->force_iters) break;
      continue;
    }
    if (rnorm <= atol) break;
    else if (breakdown) break;
    else if (presid <= ptol) ptol_max_factor = std::max(eps, 0.25 * ptol_max_factor);
    else ptol_max_factor = std::min(1.0, 1.5 * ptol_max_factor);
    ptol = presid * std::min(ptol_max_factor, atol / rnorm);
  }
  *n_iters = inner_iter;
  *info = (rnorm <= atol) ? 0 : o->maxiter;
  return 0;
}

// Right-preconditioned flexible GMRES with the call shape of pyamg.krylov.fgmres as used at
// solve.py:285 (tol on ||r||/||b||, maxiter = total inner iterations, one cycle unless restart < maxiter).
static int gmres_right(mpbp_plan* p, const double* b, double* x, const mpbp_gmres_opts* o, double* ws, double* hist,
                       int hist_cap, int* n_iters, int* info) {
  const size_t len = 5 * p->lev[0].fs();
  const int m = o->restart;
  double* V = ws;                              // (m+1) x len
  double* Z = V + (size_t)(m + 1) * len;       // m x len
  double* w = Z + (size_t)m * len;
  double* tmp = w + len;
  double* hs = p->hscal;
  double* ds = p->scal;
  const bool pc = o->use_precond != 0;
  const bool forced = o->force_iters > 0;
  const int maxit = forced ? o->force_iters : o->maxiter;
  *n_iters = 0;
  *info = maxit;
  RET(v_nrm2(p, b, len, ds));
  RET(fetch_scal(p, ds, 1, hs));
  double bn = hs[0];
  if (bn == 0.0) bn = 1.0;
  if (!o->x0_nonzero) CU(cudaMemsetAsync(x, 0, len * sizeof(double), p->st));
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gv(m + 1), yv(m);
  auto Hm = [&](int i, int j) -> double& { return H[(size_t)i * m + j]; };
  p->last_H.assign((size_t)(m + 1) * m, 0.0);
  p->last_m = m;
  p->last_k = 0;
  // low-synchronisation orthogonalisation needs the all-reduce inside the kernels (or a single rank)
  const bool lowsync = p->lowsync && (p->nranks == 1 || p->red.nranks > 1);
  std::vector<double> Lg((size_t)m * m, 0.0), hcoef(m + 1, 0.0);
  int it = 0;
  bool first = true;
  bool stop = false;
  while (it < maxit && !stop) {
    if (first && !o->x0_nonzero) {
      RET(v_copy(p, b, tmp, len));
    } else {
      RET(op_stokes(p, 0, 0, true, x, nullptr, w, 0.0));
      RET(v_axpby(p, 1.0, b, -1.0, w, tmp, len));
    }
    first = false;
    RET(v_nrm2(p, tmp, len, ds));
    k_scale_dev<<<ew_blocks(len), 256, 0, p->st>>>(ds, 1, tmp, V, len);
    LAUNCH_CHECK(p);
    RET(fetch_scal(p, ds, 1, hs));
    const double beta = hs[0];
    if (beta < o->rtol * bn && !forced) { *info = 0; break; }
    std::fill(gv.begin(), gv.end(), 0.0);
    gv[0] = beta;
    int jdone = 0;
    double res = beta;
    for (int j = 0; j < m; ++j) {
      double* vj = V + (size_t)j * len;
      double* zj = Z + (size_t)j * len;
      RET(psolve(p, pc, vj, zj, len));                       // Z_j = M v_j
      RET(op_stokes(p, 0, 0, true, zj, nullptr, w, 0.0));    // w = A Z_j
      if (lowsync) {
        // Low-synchronisation modified Gram-Schmidt: ONE pass over the basis computes r = V^T w together with the
        // Gram column g = V^T v_j of the newest basis vector; with L = strictly lower part of V^T V the MGS
        // coefficients are the forward substitution (I + L) h = r (h_k = <v_k, w - sum_{i<k} h_i v_i> exactly), and a
        // second pass applies w -= V h and returns ||w||.  2j+6 vector passes and 2 reductions per iteration instead
        // of 4j+6 passes and j+2 reductions.
        const int nc = (j + 1 + kMaxMulti - 1) / kMaxMulti;
        RET(v_multi_dot2(p, V, len, j + 1, w, vj, len, ds));
        RET(fetch_scal(p, ds, 16 * nc, hs));
        for (int k = 0; k <= j; ++k) {
          const int c = k / kMaxMulti, o = k % kMaxMulti, nv = std::min(kMaxMulti, j + 1 - c * kMaxMulti);
          hcoef[k] = hs[16 * c + o];
          if (k < j) Lg[(size_t)j * m + k] = hs[16 * c + nv + o];  // <v_k, v_j>
        }
        for (int k = 0; k <= j; ++k) {
          double sacc = hcoef[k];
          for (int i = 0; i < k; ++i) sacc -= Lg[(size_t)k * m + i] * hcoef[i];
          hcoef[k] = sacc;
        }
        RET(v_multi_axpy_nrm(p, V, len, j + 1, hcoef.data(), w, len, ds));
        k_scale_dev<<<ew_blocks(len), 256, 0, p->st>>>(ds, 1, w, V + (size_t)(j + 1) * len, len);
        LAUNCH_CHECK(p);
        RET(fetch_scal(p, ds, 1, hs));
        hs[j + 1] = hs[0];
        for (int i = 0; i <= j; ++i) hs[i] = hcoef[i];
      } else {
        RET(v_mgs(p, V, len, j, w, len, ds, ds + j + 1, nullptr));
        k_scale_dev<<<ew_blocks(len), 256, 0, p->st>>>(ds + j + 1, 1, w, V + (size_t)(j + 1) * len, len);
        LAUNCH_CHECK(p);
        RET(fetch_scal(p, ds, j + 2, hs));
      }
      for (int i = 0; i <= j + 1; ++i) Hm(i, j) = hs[i];
      if (j == 0) std::fill(p->last_H.begin(), p->last_H.end(), 0.0);
      for (int i = 0; i <= j + 1; ++i) p->last_H[(size_t)i * m + j] = hs[i];
      p->last_k = j + 1;
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * Hm(i, j) + sn[i] * Hm(i + 1, j);
        Hm(i + 1, j) = -sn[i] * Hm(i, j) + cs[i] * Hm(i + 1, j);
        Hm(i, j) = t;
      }
      const double den = std::hypot(Hm(j, j), Hm(j + 1, j));
      cs[j] = Hm(j, j) / den;
      sn[j] = Hm(j + 1, j) / den;
      Hm(j, j) = den;
      Hm(j + 1, j) = 0.0;
      gv[j + 1] = -sn[j] * gv[j];
      gv[j] = cs[j] * gv[j];
      it++;
      jdone = j + 1;
      res = std::fabs(gv[j + 1]);
      if (hist && it <= hist_cap) hist[it - 1] = res / bn;
      if (o->iter_cb && o->xk_buf) {
        // the iterate of this iteration, x_k = x + Z y_k (what pyamg hands to callback=, solve.py:285): one
        // back substitution on the host and one multi-axpy -- no extra preconditioner apply
        for (int i = jdone - 1; i >= 0; --i) {
          double s = gv[i];
          for (int k = i + 1; k < jdone; ++k) s -= Hm(i, k) * yv[k];
          yv[i] = s / Hm(i, i);
        }
        double* xk = (double*)o->xk_buf;
        RET(v_copy(p, x, xk, len));
        RET(v_multi_axpy(p, Z, len, jdone, yv.data(), xk, len));
        CU(cudaStreamSynchronize(p->st));
        if (o->iter_cb(o->cb_user, it, res / bn) != 0) stop = true;
      }
      if ((!forced && res < o->rtol * bn) || it >= maxit || stop) break;
    }
    // back substitution on the jdone x jdone triangle
    for (int i = jdone - 1; i >= 0; --i) {
      double s = gv[i];
      for (int k = i + 1; k < jdone; ++k) s -= Hm(i, k) * yv[k];
      yv[i] = s / Hm(i, i);
    }
    RET(v_multi_axpy(p, Z, len, jdone, yv.data(), x, len));
    if (!forced && res < o->rtol * bn) { *info = 0; break; }
  }
  *n_iters = it;
  return 0;
}

static int gmres_dispatch(mpbp_plan* p, const double* b, double* x, const mpbp_gmres_opts* o, double* hist,
                          int hist_cap, int* n_iters, int* info) {
  if (!o || o->restart < 1 || o->maxiter < 1) return set_err(MPBP_E_ARG, "gmres: bad options");
  if (o->restart > kScal - 8) return set_err(MPBP_E_ARG, "gmres: restart must be <= %d", kScal - 8);
  size_t need = 0;
  RET(mpbp_gmres_workspace_bytes(p, o, &need));
  double* ws = nullptr;
  if (o->workspace) {
    if (o->workspace_bytes < need) return set_err(MPBP_E_NOMEM, "gmres workspace too small: %zu < %zu", o->workspace_bytes, need);
    ws = (double*)(((uintptr_t)o->workspace + 255) & ~uintptr_t(255));
  } else {
    if (p->kry_owned_bytes < need) {
      if (p->kry_owned) cudaFree(p->kry_owned);
      p->kry_owned = nullptr;
      p->kry_owned_bytes = 0;
      CU(cudaMalloc(&p->kry_owned, need));
      p->kry_owned_bytes = need;
    }
    ws = (double*)p->kry_owned;
  }
  int ni = 0, inf = 0;
  int rc = (o->side == MPBP_SIDE_LEFT) ? gmres_left(p, b, x, o, ws, hist, hist_cap, &ni, &inf)
                                       : gmres_right(p, b, x, o, ws, hist, hist_cap, &ni, &inf);
  if (n_iters) *n_iters = ni;
  if (info) *info = inf;
  if (rc) return rc;
  CU(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int mpbp_gmres_last_hessenberg(const mpbp_plan* p, double* H, int ldh, int* k) {
  if (!p || !k) return set_err(MPBP_E_ARG, "null argument");
  *k = p->last_k;
  if (!H) return 0;  // size query
  if (ldh < p->last_k) return set_err(MPBP_E_ARG, "ldh=%d < cycle length %d", ldh, p->last_k);
  for (int i = 0; i <= p->last_k; ++i)
    for (int j = 0; j < p->last_k; ++j) H[(size_t)i * ldh + j] = p->last_H[(size_t)i * p->last_m + j];
  return 0;
}

extern "C" int mpbp_gmres(mpbp_plan* p, const double* b, double* x, const mpbp_gmres_opts* o, double* hist, int hist_cap,
                          int* n_iters, int* info, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b || !x || b == x) return set_err(MPBP_E_ARG, "gmres: bad pointers");
  return gmres_dispatch(p, b, x, o, hist, hist_cap, n_iters, info);
  };
  return leave(p, body());
}
extern "C" int mpbp_gmres_host(mpbp_plan* p, const double* b_host, double* x_host, const mpbp_gmres_opts* o, double* hist,
                               int hist_cap, int* n_iters, int* info, void* stream) {
  ENTER(p, stream);
  auto body = [&]() -> int {
  if (!b_host || !x_host) return set_err(MPBP_E_ARG, "gmres_host: bad pointers");
  RET(ensure_hostbuf(p));
  const size_t bytes = 5 * p->lev[0].fs() * sizeof(double);
  CU(cudaMemcpyAsync(p->hostbuf_dev[0], b_host, bytes, cudaMemcpyHostToDevice, p->st));
  if (o && o->x0_nonzero) CU(cudaMemcpyAsync(p->hostbuf_dev[1], x_host, bytes, cudaMemcpyHostToDevice, p->st));
  RET(gmres_dispatch(p, p->hostbuf_dev[0], p->hostbuf_dev[1], o, hist, hist_cap, n_iters, info));
  CU(cudaMemcpyAsync(x_host, p->hostbuf_dev[1], bytes, cudaMemcpyDeviceToHost, p->st));
  CU(cudaStreamSynchronize(p->st));
  return 0;
  };
  return leave(p, body());
}
