// Cell-parallel velocity-block kernels for the SMALL whole-grid multigrid levels (n <= 512, fp64, sm_100a).
//
// The marching kernels of stokes.cuh stream a strip of rows through a register window: every global load is issued
// exactly once, which is what an HBM-bound level needs.  On the small levels (a few thousand cells, all L2-resident,
// replicated on every rank of a multi-GPU run) a launch is pure LATENCY instead: the march is a chain of dependent
// rows and a 4-row strip takes 5-9 us.  Here one thread owns one cell and evaluates the whole stencil from (cached)
// global loads that are all issued up front -- one load round trip per launch, 2-3 us -- at the price of re-reading
// neighbours, which costs nothing at these sizes.  Same operator, same flux form and evaluation order as stokes.cuh
// (DESIGN.md section 2), same fused variants:
//   k_cell_sweep<IN 0>   y = x + omega (b - F x) / diag(F)                 damped Jacobi, solve.py:149-159
//   k_cell_sweep<IN 1>   x1 = wd .* b on the fly; y = x1 + wd (b - F x1)   pre-smoothing pair from a zero guess
//   k_cell_sweep<IN 2>   xt = x + P e_c on the fly; y = xt + omega (b - F xt)/diag   prolongation + first post-sweep
//   k_cell_rr            b_c = R (b - F x)                                 residual + full-weighting restriction
// Whole-grid (periodic) levels below level 0 only: theta comes padded by one row, the mass term is the face average.
#pragma once
#include "stencil.cuh"

namespace mpbp {

struct CellArgs {
  const double* th;   // padded theta, (n+2) x n
  Phys ph;
  int n;
  const double* x;    // 4 fields, stride n*n (IN 0 / IN 2, residual)
  const double* b;    // rhs
  const double* wd;   // omega / diag(F) (IN 1)
  const double* ec;   // coarse correction, 4 fields, stride (n/2)^2 (IN 2)
  double* y;          // output (sweeps)
  double* bc;         // coarse rhs (k_cell_rr), 4 fields, stride (n/2)^2
  double omega;
};

namespace cell {

__device__ __forceinline__ int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// value of the (possibly virtual) input field k at cell (r, c), indices already wrapped into [0, n)
template <int IN>
__device__ __forceinline__ double xval(const CellArgs& a, int k, int r, int c) {
  const int n = a.n;
  const int idx = k * n * n + r * n + c;
  if (IN == 1) return a.b[idx] * a.wd[idx];
  double v = a.x[idx];
  if (IN == 2) {
    // P = 4 R^T (k_restrict_F's full weighting): u-type fields constant in y / linear in x, v-type the transpose
    const int nc = n >> 1, R = r >> 1, C = c >> 1;
    const double* e = a.ec + k * nc * nc;
    // both coarse values are loaded unconditionally (no load hides behind a branch: one round trip for the launch)
    const double e0 = e[R * nc + C];
    const double e1 = ((k & 1) == 0) ? e[R * nc + (C + 1 == nc ? 0 : C + 1)] : e[(R + 1 == nc ? 0 : R + 1) * nc + C];
    const bool odd = ((k & 1) == 0) ? (c & 1) : (r & 1);
    v += odd ? 0.5 * (e0 + e1) : e0;
  }
  return v;
}

struct Th {  // theta around a cell: t[dr+1][dc+1], dr, dc in {-1, 0, 1}
  double t[3][3];
};
__device__ __forceinline__ Th load_theta(const CellArgs& a, int r, int cm, int c, int cp) {
  Th T;
  const int n = a.n;
#pragma unroll
  for (int dr = 0; dr < 3; ++dr) {
    const double* row = a.th + (size_t)(r + dr) * n;  // padded: row r-1+dr of the grid is row r+dr of the array
    T.t[dr][0] = row[cm];
    T.t[dr][1] = row[c];
    T.t[dr][2] = row[cp];
  }
  return T;
}

// F x at cell (r, c) for all four face fields, and (DIAG) the diagonal of F there.  Same expressions, in the same
// order, as one marching step of stokes.cuh.
template <int IN, bool DIAG>
__device__ __forceinline__ void F_cell(const CellArgs& a, int r, int c, double (&y)[4], double (&xc)[4], double (&dg)[4]) {
  const int n = a.n;
  const int rm = wrap(r - 1, n), rp = wrap(r + 1, n), cm = wrap(c - 1, n), cp = wrap(c + 1, n);
  const Th T = load_theta(a, r, cm, c, cp);
  const Phys& ph = a.ph;
  const double th_c = T.t[1][1], th_m = T.t[0][1], th_p = T.t[2][1];
  const double a_c = th_c + T.t[1][0];          // theta[r][c] + theta[r][c-1]
  const double a_m = th_m + T.t[0][0];          // row r-1
  const double a_p = th_p + T.t[2][0];          // row r+1
  const double a_ce = T.t[1][2] + th_c;         // column c+1: theta[r][c+1] + theta[r][c]
  const double a_me = T.t[0][2] + th_m;
  const double node_c = 0.25 * (a_c + a_m);     // corner (r, c)
  const double node_p = 0.25 * (a_p + a_c);     // corner (r+1, c)
  const double node_e = 0.25 * (a_ce + a_me);   // corner (r, c+1)
  const double th_w = T.t[1][0];                // theta[r][c-1]
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int ku = 2 * s, kv = 2 * s + 1;
    const double u_c = xval<IN>(a, ku, r, c), u_e = xval<IN>(a, ku, r, cp), u_w = xval<IN>(a, ku, r, cm);
    const double u_m = xval<IN>(a, ku, rm, c), u_me = xval<IN>(a, ku, rm, cp), u_p = xval<IN>(a, ku, rp, c);
    const double v_c = xval<IN>(a, kv, r, c), v_p = xval<IN>(a, kv, rp, c), v_w = xval<IN>(a, kv, r, cm);
    const double v_pw = xval<IN>(a, kv, rp, cm), v_m = xval<IN>(a, kv, rm, c), v_e = xval<IN>(a, kv, r, cp);
    const double wc = s ? 1.0 - th_c : th_c, ww = s ? 1.0 - th_w : th_w, wm = s ? 1.0 - th_m : th_m;
    const double nc_ = s ? 1.0 - node_c : node_c, np_ = s ? 1.0 - node_p : node_p, ne_ = s ? 1.0 - node_e : node_e;
    const double Q_c = wc * ((u_e - u_c) + (v_p - v_c));     // cell (r, c)
    const double Q_w = ww * ((u_c - u_w) + (v_pw - v_w));    // cell (r, c-1)
    const double Q_m = wm * ((u_me - u_m) + (v_c - v_m));    // cell (r-1, c)
    const double T_c = nc_ * ((u_m - u_c) + (v_c - v_w));    // corner (r, c)
    const double T_p = np_ * ((u_c - u_p) + (v_p - v_pw));   // corner (r+1, c)
    const double T_e = ne_ * ((u_me - u_e) + (v_e - v_c));   // corner (r, c+1)
    y[ku] = (Q_c - Q_w) + (T_c - T_p);                       // (L u)_u
    y[kv] = (T_e - T_c) + (Q_c - Q_m);                       // (L u)_v
    xc[ku] = u_c;
    xc[kv] = v_c;
  }
  const double fu_c = 0.5 * a_c, fv_c = 0.5 * (th_c + th_m);
  const double dXu = ph.d_u * (ph.xi * fu_c * (1.0 - fu_c));  // preconditioner.py:124
  const double dXv = ph.d_u * (ph.xi * fv_c * (1.0 - fv_c));  // preconditioner.py:125
  const double cmu = ph.c * fu_c, cmv = ph.c * fv_c;          // face-average mass term (levels below 0)
  const double du = xc[0] - xc[2], dv = xc[1] - xc[3];
  const double Lu_n = y[0], Lv_n = y[1], Lu_s = y[2], Lv_s = y[3];
  y[0] = cmu * xc[0] - dXu * du + ph.kap_n * Lu_n;
  y[2] = (ph.c - cmu) * xc[2] + dXu * du + ph.kap_s * Lu_s;
  y[1] = cmv * xc[1] - dXv * dv + ph.kap_n * Lv_n;
  y[3] = (ph.c - cmv) * xc[3] + dXv * dv + ph.kap_s * Lv_s;
  if (DIAG) {
    const double su = a_c + node_c + node_p;            // tE+tW+nN+nS, preconditioner.py:127
    const double sv = th_m + th_c + node_c + node_e;    // tN+tC+nL+nR, preconditioner.py:242
    dg[0] = cmu - dXu - ph.kap_n * su;
    dg[2] = (ph.c - cmu) - dXu - ph.kap_s * (4.0 - su);
    dg[1] = cmv - dXv - ph.kap_n * sv;
    dg[3] = (ph.c - cmv) - dXv - ph.kap_s * (4.0 - sv);
  }
}

}  // namespace cell

constexpr int kCellBX = 32, kCellBY = 4;  // cells per block: 32 columns x 4 rows

template <int IN>
__global__ void __launch_bounds__(kCellBX * kCellBY) k_cell_sweep(const __grid_constant__ CellArgs a) {
  const int n = a.n;
  const int c = blockIdx.x * kCellBX + (threadIdx.x & 31), r = blockIdx.y * kCellBY + (threadIdx.x >> 5);
  if (c >= n || r >= n) return;
  const int off = r * n + c, fs = n * n;
  double y[4], xc[4], dg[4];
  cell::F_cell<IN, IN != 1>(a, r, c, y, xc, dg);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double bk = a.b[off + k * fs];
    // IN 1: omega/diag(F) is the streamed field wd; otherwise the diagonal is recomputed (as in the marching kernels)
    const double out = (IN == 1) ? xc[k] + (bk - y[k]) * a.wd[off + k * fs] : xc[k] + a.omega * (bk - y[k]) * fast_rcp(dg[k]);
    a.y[off + k * fs] = out;
  }
}

// One thread per COARSE cell (R, C): the residual b - F x at the fine cells under the full-weighting stencils
//   u-type: (1/4, 1/2, 1/4) over columns 2C-1, 2C, 2C+1 x (1/2, 1/2) over rows 2R, 2R+1
//   v-type: (1/2, 1/2) over columns 2C, 2C+1 x (1/4, 1/2, 1/4) over rows 2R-1, 2R, 2R+1
// is evaluated on the fly; nothing but the coarse rhs is stored.
__global__ void __launch_bounds__(kCellBX* kCellBY) k_cell_rr(const __grid_constant__ CellArgs a) {
  const int n = a.n, nc = n >> 1;
  const int C = blockIdx.x * kCellBX + (threadIdx.x & 31), R = blockIdx.y * kCellBY + (threadIdx.x >> 5);
  if (C >= nc || R >= nc) return;
  const int fs = n * n;
  // residuals on the 4 x 4 fine cells rows 2R-1 .. 2R+2, columns 2C-1 .. 2C+2 (only the ones the weights touch)
  double ru[2][2][3];  // [phase][row 2R, 2R+1][col 2C-1, 2C, 2C+1]
  double rv[2][3][2];  // [phase][row 2R-1, 2R, 2R+1][col 2C, 2C+1]
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      const bool need_u = dr >= 0;   // rows 2R, 2R+1
      const bool need_v = dc >= 0;   // columns 2C, 2C+1
      if (!need_u && !need_v) continue;
      const int r = cell::wrap(2 * R + dr, n), c = cell::wrap(2 * C + dc, n);
      double y[4], xc[4], dg[4];
      cell::F_cell<0, false>(a, r, c, y, xc, dg);
      const int off = r * n + c;
      if (need_u) {
        ru[0][dr][dc + 1] = a.b[off] - y[0];
        ru[1][dr][dc + 1] = a.b[off + 2 * fs] - y[2];
      }
      if (need_v) {
        rv[0][dr + 1][dc] = a.b[off + fs] - y[1];
        rv[1][dr + 1][dc] = a.b[off + 3 * fs] - y[3];
      }
    }
  }
  const int oc = R * nc + C, fsc = nc * nc;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const double sm = 0.5 * (ru[s][0][0] + ru[s][1][0]), s0 = 0.5 * (ru[s][0][1] + ru[s][1][1]),
                 sp = 0.5 * (ru[s][0][2] + ru[s][1][2]);
    a.bc[oc + (2 * s) * fsc] = 0.25 * sm + 0.5 * s0 + 0.25 * sp;
    const double tm = 0.5 * (rv[s][0][0] + rv[s][0][1]), t0 = 0.5 * (rv[s][1][0] + rv[s][1][1]),
                 tp = 0.5 * (rv[s][2][0] + rv[s][2][1]);
    a.bc[oc + (2 * s + 1) * fsc] = 0.25 * tm + 0.5 * t0 + 0.25 * tp;
  }
}

}  // namespace mpbp
