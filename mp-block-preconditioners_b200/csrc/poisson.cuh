// Fused marching kernels of the pressure-Poisson V-cycle on whole-grid (periodic) levels (fp64, sm_100a).
//
// GtG = -D G is the 5-point operator with face weights fu_n^2 + fu_s^2 (solve.py:246-247; k_poisson in stencil.cuh).
// Its V(2,2) cycle used seven passes per level (first sweep from zero, sweep, residual, restriction, prolongation, two
// sweeps); the same three fusions as on the velocity block (stokes.cuh) leave four:
//   IN 1, MODE 2        pre-smoothing PAIR from a zero guess: x1 = omega b .* wd (wd = 1/diag precomputed per level) is
//                       formed on the fly, x2 = x1 + omega (b - GtG x1) .* wd           32N bytes instead of 24N + 32N
//   IN 0, MODE 1, EP 2  residual + 4-cell-average restriction in registers               26N instead of 32N + 10N
//   IN 2, MODE 2        piecewise-constant prolongation x + P e_c applied while loading + first post-sweep
//                       (EP 1: ... which is also the last one: Chebyshev epilogue)       34N instead of 18N + 32N
// Every variant performs the arithmetic of the passes it replaces in the same order (bit-identical on the CPU shim; on
// the device nvcc contracts multiply-adds per instantiation, so fused and unfused cycles agree to rounding).
// Thread mapping as in stencil.cuh (column marching: a lane owns a column, 3-row window, shuffles for the neighbours).
#pragma once
#include "stencil.cuh"

namespace mpbp {

struct PoissonFArgs {
  const double* x;    // iterate, n x n (IN 0 / IN 2)
  const double* b;    // rhs
  const double* wd;   // 1 / diag(GtG) exactly as the sweeps compute it (IN 1)
  const double* ec;   // coarse correction, (n/2) x (n/2) (IN 2)
  const double* th;   // padded theta
  double* y;          // EP 0
  double* bc;         // EP 2: coarse rhs, (n/2) x (n/2)
  Geo g;
  Phys ph;
  double omega;
  ChebEp ce;          // EP 1
};

template <int IN, int MODE, int EP>
__global__ void __launch_bounds__(kBlockThreads) k_poisson_f(const __grid_constant__ PoissonFArgs a) {
  static_assert(!(EP == 2 && MODE != 1), "restriction epilogue belongs to the residual");
  static_assert(!(EP == 1 && MODE != 2), "Chebyshev epilogue belongs to the last sweep");
  const LaneGeom lg = lane_geom(a.g.n);
  if (!lg.alive) return;
  const int n = a.g.n, c = lg.cc, nc = n >> 1;
  int r0, r1;
  if (!strip_rows(a.g, r0, r1)) return;
  const double* __restrict__ th = a.th;
  const double* __restrict__ b = a.b;
  const Phys& ph = a.ph;
  // value of the (possibly virtual) iterate at (row rr, my column), rr in [-1, n]: periodic rows
  auto pval = [&](int rr) -> double {
    const int w = rr < 0 ? rr + n : (rr >= n ? rr - n : rr);
    const int idx = w * n + c;
    if (IN == 1) return a.omega * b[idx] * a.wd[idx];  // (omega b) / diag: bitwise the unfused first sweep
    double v = a.x[idx];
    if (IN == 2) v += a.ec[(w >> 1) * nc + (c >> 1)];
    return v;
  };
  double th_m = th_row(th, r0 - 1, n)[c];
  double th_c = th_row(th, r0, n)[c];
  double p_m = pval(r0 - 1), p_c = pval(r0);
  double fv = 0.5 * (th_c + th_m);
  double wv_c = fv * fv + (1.0 - fv) * (1.0 - fv);
  double Hy_c = wv_c * (p_m - p_c);
  double s_even = 0.0;  // EP 2: column-pair sum of the residual on the even row of the current pair
#pragma unroll 2
  for (int r = r0; r < r1; ++r) {
    const double th_p = th_row(th, r + 1, n)[c];
    const double p_p = pval(r + 1);
    const double fu = 0.5 * (th_c + shfl_up1(th_c));
    const double wu = fu * fu + (1.0 - fu) * (1.0 - fu);
    const double Hx = wu * (p_c - shfl_up1(p_c));
    fv = 0.5 * (th_p + th_c);
    const double wv_p = fv * fv + (1.0 - fv) * (1.0 - fv);
    const double Hy_p = wv_p * (p_c - p_p);
    double out = -ph.dp_h2 * ((shfl_dn1(Hx) - Hx) + (Hy_c - Hy_p));
    const int off = r * n + c;
    if (MODE == 1) out = b[off] - out;
    if (MODE == 2) {
      if (IN == 1) {
        out = p_c + a.omega * (b[off] - out) * a.wd[off];
      } else {
        const double dg = ph.dp_h2 * (shfl_dn1(wu) + wu + wv_c + wv_p);
        out = p_c + a.omega * (b[off] - out) * fast_rcp(dg);
      }
    }
    if (EP == 2) {
      // 4-cell average (k_restrict_P): 0.25 * ((row 2R: c, c+1) + (row 2R+1: c, c+1)), even columns store
      const double s = out + shfl_dn1(out);
      if ((r & 1) == 0) {
        s_even = s;
      } else if (lg.store && (c & 1) == 0) {
        a.bc[(r >> 1) * nc + (c >> 1)] = 0.25 * (s_even + s);
      }
    } else if (lg.store) {
      if (EP == 1) {
        const double dk = a.ce.read_d ? a.ce.d[off] : 0.0;
        const double dn = a.ce.ca * dk + a.ce.cb * out;
        if (a.ce.write_d) a.ce.d[off] = dn;
        a.ce.xk[off] = (a.ce.read_x ? a.ce.xk[off] : 0.0) + dn;
      } else {
        a.y[off] = out;
      }
    }
    th_m = th_c; th_c = th_p; p_m = p_c; p_c = p_p; wv_c = wv_p; Hy_c = Hy_p;
  }
}

}  // namespace mpbp
