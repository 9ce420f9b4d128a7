// Persistent coarse-grid V-cycle: ONE single-CTA kernel runs the whole multigrid sub-hierarchy below a
// threshold grid size (smoothing, residual, restriction, dense coarsest solve, prolongation, post-smoothing),
// replacing ~9 launches per level.  On the small levels every kernel is pure launch latency (a few us each,
// identical on 1 or 8 GPUs), so this is what the multi-GPU scaling and the replicated coarse levels need.
//
// EXPERIMENTAL in round 1 (off unless MPBP_COARSE=<n>): parity-tested on the B200 and on the SIMT-on-CPU shim
// (tests/emu), not yet timed.
//
// The per-cell operator is written with the reference's coefficient table (preconditioner.py:127-179,
// :242-295), independent of the flux form used by the marching kernels.
#pragma once
#include "stencil.cuh"

namespace mpbp {

constexpr int kCoarseMaxLevels = 6;
constexpr int kCoarseThreadsF = 512;   // velocity block: 4 fields per cell, needs > 64 registers
constexpr int kCoarseThreadsP = 1024;  // pressure Poisson

struct CoarseLevel {
  int n;
  Phys ph;
  const double* th;  // padded theta, (n+2) x n
  double *b, *x, *t, *r;
};
struct CoarseArgs {
  int nlev;  // levels of the sub-hierarchy; the last one is solved with the dense (pseudo-)inverse
  CoarseLevel lev[kCoarseMaxLevels];
  const double* Minv_t;  // row-major dense inverse of the coarsest level
  int m;
  const double* b_in;  // rhs on the first level of the sub-hierarchy
  double* x_out;       // result on the first level
  double omega;
  int nu1, nu2;
};

namespace coarse {

__device__ __forceinline__ int wrapi(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

struct LevelView {
  int n;
  size_t N;
  const double* th;
  __device__ __forceinline__ double T(int r, int c) const { return th[(size_t)(r + 1) * n + wrapi(c, n)]; }  // r in [-1, n]
  __device__ __forceinline__ double node(int r, int c) const {  // top-left corner of cell (r,c), r in [0, n]
    return 0.25 * (T(r, c) + T(r, c - 1) + T(r - 1, c) + T(r - 1, c - 1));
  }
};
__device__ __forceinline__ double phs(double t, int s) { return s ? 1.0 - t : t; }
__device__ __forceinline__ double at(const double* f, int n, int r, int c) {
  return f[(size_t)wrapi(r, n) * n + wrapi(c, n)];
}

struct Cell4 {
  double un, vn, us, vs;
};

// (L [u v])_u and (L [u v])_v at cell (r,c) for phase s, coefficient-table form (x h^2)
__device__ __forceinline__ double Lu(const LevelView& V, int s, const double* u, const double* v, int r, int c) {
  const int n = V.n;
  const double tE = phs(V.T(r, c), s), tW = phs(V.T(r, c - 1), s);
  const double nN = phs(V.node(r, c), s), nS = phs(V.node(r + 1, c), s);
  return tW * at(u, n, r, c - 1) + tE * at(u, n, r, c + 1) + nN * at(u, n, r - 1, c) + nS * at(u, n, r + 1, c) -
         (tE + tW + nN + nS) * at(u, n, r, c) + (nN - tE) * at(v, n, r, c) + (tW - nN) * at(v, n, r, c - 1) +
         (nS - tW) * at(v, n, r + 1, c - 1) + (tE - nS) * at(v, n, r + 1, c);
}
__device__ __forceinline__ double Lv(const LevelView& V, int s, const double* u, const double* v, int r, int c) {
  const int n = V.n;
  const double tC = phs(V.T(r, c), s), tN = phs(V.T(r - 1, c), s);
  const double nL = phs(V.node(r, c), s), nR = phs(V.node(r, c + 1), s);
  return nL * at(v, n, r, c - 1) + nR * at(v, n, r, c + 1) + tN * at(v, n, r - 1, c) + tC * at(v, n, r + 1, c) -
         (tN + tC + nL + nR) * at(v, n, r, c) + (nL - tC) * at(u, n, r, c) + (tC - nR) * at(u, n, r, c + 1) +
         (tN - nL) * at(u, n, r - 1, c) + (nR - tN) * at(u, n, r - 1, c + 1);
}

// diagonal of F at one cell (preconditioner.py:127, :242 with the mass/drag terms of :331-337)
__device__ __forceinline__ void diag_cell(const CoarseLevel& L, const LevelView& V, int r, int c, Cell4& dg) {
  const Phys& ph = L.ph;
  const double fu = 0.5 * (V.T(r, c) + V.T(r, c - 1));
  const double fv = 0.5 * (V.T(r, c) + V.T(r - 1, c));
  double mu = fu, mv = fv;
  if (ph.mass_mode) {
    mu = 0.25 * ph.sxf[c] * ph.syc[r] + 0.5;
    mv = 0.25 * ph.sxc[c] * ph.syf[r] + 0.5;
  }
  const double dXu = ph.d_u * (ph.xi * fu * (1.0 - fu)), dXv = ph.d_u * (ph.xi * fv * (1.0 - fv));
  const double cmu = ph.c * mu, cmv = ph.c * mv;
  const double su = V.T(r, c) + V.T(r, c - 1) + V.node(r, c) + V.node(r + 1, c);
  const double sv = V.T(r - 1, c) + V.T(r, c) + V.node(r, c) + V.node(r, c + 1);
  dg.un = cmu - dXu - ph.kap_n * su;
  dg.us = (ph.c - cmu) - dXu - ph.kap_s * (4.0 - su);
  dg.vn = cmv - dXv - ph.kap_n * sv;
  dg.vs = (ph.c - cmv) - dXv - ph.kap_s * (4.0 - sv);
}

// F x at one cell, and the diagonal of F there
__device__ __forceinline__ void stokes_cell(const CoarseLevel& L, const LevelView& V, const double* x, int r, int c,
                                            Cell4& Fx, Cell4& dg) {
  const int n = V.n;
  const size_t N = V.N, k = (size_t)r * n + c;
  const double *un = x, *vn = x + N, *us = x + 2 * N, *vs = x + 3 * N;
  const Phys& ph = L.ph;
  const double fu = 0.5 * (V.T(r, c) + V.T(r, c - 1));
  const double fv = 0.5 * (V.T(r, c) + V.T(r - 1, c));
  double mu = fu, mv = fv;
  if (ph.mass_mode) {
    mu = 0.25 * ph.sxf[c] * ph.syc[r] + 0.5;
    mv = 0.25 * ph.sxc[c] * ph.syf[r] + 0.5;
  }
  const double dXu = ph.d_u * (ph.xi * fu * (1.0 - fu)), dXv = ph.d_u * (ph.xi * fv * (1.0 - fv));
  const double cmu = ph.c * mu, cmv = ph.c * mv;
  Fx.un = (cmu - dXu) * un[k] + dXu * us[k] + ph.kap_n * Lu(V, 0, un, vn, r, c);
  Fx.vn = (cmv - dXv) * vn[k] + dXv * vs[k] + ph.kap_n * Lv(V, 0, un, vn, r, c);
  Fx.us = ((ph.c - cmu) - dXu) * us[k] + dXu * un[k] + ph.kap_s * Lu(V, 1, us, vs, r, c);
  Fx.vs = ((ph.c - cmv) - dXv) * vs[k] + dXv * vn[k] + ph.kap_s * Lv(V, 1, us, vs, r, c);
  diag_cell(L, V, r, c, dg);
}

// GtG p at one cell and its diagonal
__device__ __forceinline__ void poisson_cell(const CoarseLevel& L, const LevelView& V, const double* p, int r, int c,
                                             double& Ap, double& dg) {
  const int n = V.n;
  const double fu = 0.5 * (V.T(r, c) + V.T(r, c - 1)), fuE = 0.5 * (V.T(r, c + 1) + V.T(r, c));
  const double fv = 0.5 * (V.T(r, c) + V.T(r - 1, c)), fvS = 0.5 * (V.T(r + 1, c) + V.T(r, c));
  const double wu = fu * fu + (1 - fu) * (1 - fu), wuE = fuE * fuE + (1 - fuE) * (1 - fuE);
  const double wv = fv * fv + (1 - fv) * (1 - fv), wvS = fvS * fvS + (1 - fvS) * (1 - fvS);
  const double pc = at(p, n, r, c);
  Ap = -L.ph.dp_h2 * (wuE * (at(p, n, r, c + 1) - pc) - wu * (pc - at(p, n, r, c - 1)) + wv * (at(p, n, r - 1, c) - pc) -
                      wvS * (pc - at(p, n, r + 1, c)));
  dg = L.ph.dp_h2 * (wuE + wu + wv + wvS);
}

// mode 0: y = omega b / diag ; 1: y = x + omega (b - A x)/diag ; 2: y = b - A x        (block-stride over cells)
template <bool IS_F>
__device__ void smooth_or_residual(const CoarseLevel& L, int mode, const double* x, const double* b, double* y,
                                   double omega) {
  const LevelView V{L.n, (size_t)L.n * L.n, L.th};
  const int n = L.n;
  const size_t N = V.N;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int r = idx / n, c = idx - r * n;
    if (IS_F) {
      Cell4 Fx{0, 0, 0, 0}, dg;
      if (mode == 0) {
        diag_cell(L, V, r, c, dg);
        const size_t k = idx;
        y[k] = omega * b[k] / dg.un;
        y[k + N] = omega * b[k + N] / dg.vn;
        y[k + 2 * N] = omega * b[k + 2 * N] / dg.us;
        y[k + 3 * N] = omega * b[k + 3 * N] / dg.vs;
        continue;
      }
      stokes_cell(L, V, x, r, c, Fx, dg);
      const size_t k = idx;
      if (mode == 1) {
        y[k] = x[k] + omega * (b[k] - Fx.un) / dg.un;
        y[k + N] = x[k + N] + omega * (b[k + N] - Fx.vn) / dg.vn;
        y[k + 2 * N] = x[k + 2 * N] + omega * (b[k + 2 * N] - Fx.us) / dg.us;
        y[k + 3 * N] = x[k + 3 * N] + omega * (b[k + 3 * N] - Fx.vs) / dg.vs;
      } else {
        y[k] = b[k] - Fx.un;
        y[k + N] = b[k + N] - Fx.vn;
        y[k + 2 * N] = b[k + 2 * N] - Fx.us;
        y[k + 3 * N] = b[k + 3 * N] - Fx.vs;
      }
    } else {
      double Ap = 0.0, dg;
      if (mode == 0) {
        double tmp;
        poisson_cell(L, V, b, r, c, tmp, dg);  // diagonal only
        y[idx] = omega * b[idx] / dg;
        continue;
      }
      poisson_cell(L, V, x, r, c, Ap, dg);
      y[idx] = (mode == 1) ? x[idx] + omega * (b[idx] - Ap) / dg : b[idx] - Ap;
    }
  }
}

__device__ __forceinline__ void copy_vec(const double* x, double* y, size_t len) {
  for (size_t i = threadIdx.x; i < len; i += blockDim.x) y[i] = x[i];
}

template <bool IS_F>
__device__ void restrict_level(const double* f, double* yc, int nf) {
  const int nc = nf >> 1, n = nf;
  const size_t Nf = (size_t)nf * nf, Nc = (size_t)nc * nc;
  for (int idx = threadIdx.x; idx < nc * nc; idx += blockDim.x) {
    const int R = idx / nc, C = idx - R * nc;
    const int c0 = 2 * C, ra = 2 * R, rb = 2 * R + 1;
    if (IS_F) {
      for (int s = 0; s < 2; ++s) {
        const double* u = f + (size_t)(2 * s) * Nf;
        const double* v = f + (size_t)(2 * s + 1) * Nf;
        const double um = 0.5 * (at(u, n, ra, c0 - 1) + at(u, n, rb, c0 - 1)), u0 = 0.5 * (at(u, n, ra, c0) + at(u, n, rb, c0)),
                     up = 0.5 * (at(u, n, ra, c0 + 1) + at(u, n, rb, c0 + 1));
        yc[(size_t)(2 * s) * Nc + idx] = 0.25 * um + 0.5 * u0 + 0.25 * up;
        const double wm = 0.5 * (at(v, n, ra - 1, c0) + at(v, n, ra - 1, c0 + 1)),
                     w0 = 0.5 * (at(v, n, ra, c0) + at(v, n, ra, c0 + 1)), wp = 0.5 * (at(v, n, rb, c0) + at(v, n, rb, c0 + 1));
        yc[(size_t)(2 * s + 1) * Nc + idx] = 0.25 * wm + 0.5 * w0 + 0.25 * wp;
      }
    } else {
      const double* a = f + (size_t)ra * nf + c0;
      yc[idx] = 0.25 * ((a[0] + a[1]) + (a[nf] + a[nf + 1]));
    }
  }
}

template <bool IS_F>
__device__ void prolong_add_level(const double* xc, double* xf, int nf) {
  const int nc = nf >> 1;
  const size_t Nf = (size_t)nf * nf, Nc = (size_t)nc * nc;
  for (int idx = threadIdx.x; idx < nf * nf; idx += blockDim.x) {
    const int r = idx / nf, c = idx - r * nf;
    const int R = r >> 1, C = c >> 1;
    if (IS_F) {
      const int Cp = (C + 1 == nc) ? 0 : C + 1, Rp = (R + 1 == nc) ? 0 : R + 1;
      for (int s = 0; s < 2; ++s) {
        const double* uc = xc + (size_t)(2 * s) * Nc;
        const double* vc = xc + (size_t)(2 * s + 1) * Nc;
        const double eu = (c & 1) ? 0.5 * (uc[(size_t)R * nc + C] + uc[(size_t)R * nc + Cp]) : uc[(size_t)R * nc + C];
        const double ev = (r & 1) ? 0.5 * (vc[(size_t)R * nc + C] + vc[(size_t)Rp * nc + C]) : vc[(size_t)R * nc + C];
        xf[(size_t)(2 * s) * Nf + idx] += eu;
        xf[(size_t)(2 * s + 1) * Nf + idx] += ev;
      }
    } else {
      xf[idx] += xc[(size_t)R * nc + C];
    }
  }
}

}  // namespace coarse

// One V(nu1,nu2) cycle over levels lev[0..nlev-1] of the sub-hierarchy, zero initial guess: x_out = V b_in.
template <bool IS_F>
__global__ void __launch_bounds__(IS_F ? kCoarseThreadsF : kCoarseThreadsP) k_coarse_vcycle(CoarseArgs a) {
  using namespace coarse;
  const int nf_ = IS_F ? 4 : 1;
  const int last = a.nlev - 1;
  // ---- down ----
  for (int l = 0; l < last; ++l) {
    const CoarseLevel& L = a.lev[l];
    const double* b = (l == 0) ? a.b_in : L.b;
    double* x = (l == 0) ? a.x_out : L.x;
    const size_t len = (size_t)nf_ * L.n * L.n;
    smooth_or_residual<IS_F>(L, 0, nullptr, b, x, a.omega);
    __syncthreads();
    for (int s = 1; s < a.nu1; ++s) {
      smooth_or_residual<IS_F>(L, 1, x, b, L.t, a.omega);
      __syncthreads();
      copy_vec(L.t, x, len);
      __syncthreads();
    }
    smooth_or_residual<IS_F>(L, 2, x, b, L.r, a.omega);
    __syncthreads();
    restrict_level<IS_F>(L.r, a.lev[l + 1].b, L.n);
    __syncthreads();
  }
  // ---- coarsest level: dense (pseudo-)inverse ----
  {
    const CoarseLevel& L = a.lev[last];
    const double* b = (last == 0) ? a.b_in : L.b;
    double* x = (last == 0) ? a.x_out : L.x;
    for (int i = threadIdx.x; i < a.m; i += blockDim.x) {
      double acc = 0.0;
      for (int k = 0; k < a.m; ++k) acc = fma(a.Minv_t[(size_t)i * a.m + k], b[k], acc);
      x[i] = acc;
    }
    __syncthreads();
  }
  // ---- up ----
  for (int l = last - 1; l >= 0; --l) {
    const CoarseLevel& L = a.lev[l];
    const double* b = (l == 0) ? a.b_in : L.b;
    double* x = (l == 0) ? a.x_out : L.x;
    const size_t len = (size_t)nf_ * L.n * L.n;
    prolong_add_level<IS_F>(a.lev[l + 1].x, x, L.n);
    __syncthreads();
    for (int s = 0; s < a.nu2; ++s) {
      smooth_or_residual<IS_F>(L, 1, x, b, L.t, a.omega);
      __syncthreads();
      copy_vec(L.t, x, len);
      __syncthreads();
    }
  }
}

}  // namespace mpbp
