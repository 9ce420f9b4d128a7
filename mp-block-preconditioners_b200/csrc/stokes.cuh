// Velocity-block / full-system marching kernel (fp64, sm_100a) with fused prologues and epilogues.
//
// One kernel template covers every stencil pass over the four face fields of the hot path:
//   MODE 0  y = Op x             (Op = F, or A when WITH_P)                      K1 / K2
//   MODE 1  y = b - F x                                                          residual
//   MODE 2  y = x + omega (b - F x) / diag(F)                                    damped Jacobi, solve.py:149-159
// What `x` is (IN) and where the result goes (EP) are fused variants that remove whole passes over HBM:
//   IN 0    x read from memory
//   IN 1    x = wd .* b  (wd = omega/diag(F) precomputed per level): the first sweep from a zero guess is formed
//           on the fly, so MODE 2 yields the pre-smoothing PAIR x2 = x1 + wd (b - F x1) in one pass
//   IN 2    x = x + P e_c  (P = 4 R^T, face-centred linear/constant interpolation): coarse-grid correction fused
//           into the first post-smoothing sweep
//   EP 0    y stored
//   EP 1    Chebyshev epilogue: the result is the V-cycle output z, consumed in registers:
//           d = ca*d + cb*z ; xk += d   (z never touches memory)
//   EP 2    restriction epilogue (MODE 1): the residual is full-weighted onto the coarse grid in registers
//           (shuffles across columns, a two-row carry down the march); only the coarse right-hand side is stored
// Thread mapping as in stencil.cuh ("column marching"): a lane owns one column and walks down a strip of rows
// with a register window; horizontal neighbours come from warp shuffles; lanes at the warp edge are halo lanes.
#pragma once
#include "stencil.cuh"

namespace mpbp {

struct StokesArgs {
  VecIn xin;                 // input view (IN 1: the view of b)
  const double* th;          // padded theta
  const double* b;           // rhs (MODE >= 1)
  double* y;                 // output (EP 0)
  Geo g;
  Phys ph;
  double omega;
  VecIn wd;                  // IN 1
  VecIn cin;                 // IN 2: coarse correction (4 fields), with halos on distributed levels
  int nc, rows_c;            // IN 2 / EP 2: coarse grid columns / local rows
  ChebEp ce;                 // EP 1
  double* bc;                // EP 2: coarse rhs, field stride rows_c*nc
  PushOut po;                // PUSH
  // (distributed levels: xin.land may point at a static per-level stash instead of the transient landing buffer, so a
  // later kernel -- prolongation + sweep -- can read the received rows again after newer exchanges recycled the slots)
};

template <int EP>
struct WarpTile {
  static constexpr int cols = (EP == 2) ? 28 : kWarpCols;  // output columns per warp
  static constexpr int loff = (EP == 2) ? 2 : 1;           // lane of the first output column
};

// EDGE = false is the lean instantiation for strips whose whole register window (rows r0-2 .. r1+1, coarse rows
// included) lies inside the slab: every access is base_k[32-bit offset], no halo pointers, no row selects.  Edge
// strips (the first and last of a slab) run the general instantiation with halo / periodic row lookup.
template <int IN, int MODE, bool WITH_P, int EP, bool PUSH, bool EDGE>
__device__ __forceinline__ void stokes_march(const StokesArgs& a, const int r0, const int r1) {
  static_assert(!(EP == 2 && MODE != 1), "restriction epilogue belongs to the residual");
  static_assert(!(EP == 1 && MODE != 2), "Chebyshev epilogue belongs to the last sweep");
  static_assert(!(WITH_P && (IN != 0 || EP != 0 || MODE != 0)), "the pressure coupling is only used by y = A x");
  constexpr int WC = WarpTile<EP>::cols, LOFF = WarpTile<EP>::loff;
  // EP 2 on whole-grid levels addresses rows periodically in place (it reaches row r0-2); with PUSH the level is a
  // slab of a distributed grid: rows -1 / `rows` are halo rows and the slab's first coarse row is finished with the
  // previous rank's contribution (see the end of this function)
  constexpr bool WRAP = (EP == 2) && !PUSH;
  const int n = a.g.n, rows = a.g.rows;
  const int lane = threadIdx.x & 31;
  const int j0 = (blockIdx.x * kBlockWarps + (threadIdx.x >> 5)) * WC;
  if (j0 >= n) return;
  const int j = j0 + lane - LOFF;
  int c = j % n;
  if (c < 0) c += n;
  const bool store = (lane >= LOFF) && (lane < LOFF + WC) && (j < n);
  const Phys& ph = a.ph;
  VecIn xin = a.xin;
  VecIn cin = a.cin;
  if (EDGE && (r0 == 0 || r1 == rows)) {
    // every lane fetches the halo elements of its own column (the only ones it reads) and lands them
    if (xin.dseq != nullptr) {
      halo_fetch<true>(xin, 0, WITH_P ? 5 : 4, 1, n, c, r0 == 0, r1 == rows);
    }
    if (IN == 2 && cin.dseq != nullptr) {
      const int ncc = a.nc, Cc = c >> 1, Cq = (Cc + 1 == ncc) ? 0 : Cc + 1;
      halo_fetch<true>(cin, 0, 4, 1, ncc, Cc, r0 == 0, r1 == rows);
      halo_fetch<true>(cin, 0, 4, 1, ncc, Cq, r0 == 0, r1 == rows);
    }
  }

  // PUSH: my first / last output rows also go to the ring neighbours' halo areas (peer memory over NVLink)
  PushCtx pc{};
  if (PUSH) pc = push_begin(a.po, EDGE && r0 == 0, EDGE && r1 == rows);
  LLElem* const push_prev = pc.prev;
  LLElem* const push_next = pc.next;
  const unsigned long long push_seq = pc.seq;

  const size_t fs = xin.fs;
  const double* __restrict__ th = a.th;
  const double* __restrict__ b = a.b;

  // ---- input access: value of field k at (row rr, my column), rr in [-1, rows] (EP 2: [-2, rows], whole grid) ----
  const int nc = a.nc;
  const int C = c >> 1, Cp = (C + 1 == nc) ? 0 : C + 1;
  const bool codd = (c & 1) != 0;
  auto wrap_row = [&](int rr) -> int { return rr < 0 ? rr + rows : (rr >= rows ? rr - rows : rr); };
  auto xrow = [&](const VecIn& v, int k, int rr) -> const double* {
    if (WRAP) return v.x + k * v.fs + (size_t)wrap_row(rr) * n;  // whole-grid levels only: periodic in place
    return row_ptr(v, k, rr, rows, n);
  };
  const int fs32 = (int)fs;  // 4 fields of one slab stay below 2^31 elements (checked at plan creation)
  auto ldraw = [&](int k, int rr) -> double {
    if (!EDGE) {
      const int idx = rr * n + c + k * fs32;
      const double raw = xin.x[idx];
      if (IN == 1) return raw * a.wd.x[idx];
      return raw;
    }
    const double raw = xrow(xin, k, rr)[c];
    if (IN == 1) return raw * xrow(a.wd, k, rr)[c];
    return raw;
  };
  // IN 2: the coarse correction P e_c at (row rr, my column).  u-type fields (k even) are constant in y and linear
  // in x, v-type fields linear in y and constant in x (P = 4 R^T of k_restrict_F's full weighting).  The coarse rows
  // are re-read per fine row: each is shared by two fine rows and two lanes, so these are L1 hits.
  const int fsc32 = a.rows_c * nc;   // EP 2: field stride of the coarse rhs
  const int fsci = (int)cin.fs;      // IN 2: field stride of the coarse correction (a replicated coarse level is addressed
                                     // inside its full grid, so this is not rows_c*nc there)
  auto crow = [&](int k, int R) -> const double* {
    if (!EDGE) return cin.x + (k * fsci + R * nc);
    return row_ptr(cin, k, R, a.rows_c, nc);
  };
  auto pe = [&](int k, int rr, bool odd_row) -> double {
    const int R = rr >> 1;  // arithmetic shift: row -1 belongs to coarse row -1 (the top halo)
    const double* e0 = crow(k, R);
    if ((k & 1) == 0) {
      const double e_a = e0[C], e_b = e0[Cp];
      return codd ? 0.5 * (e_a + e_b) : e_a;
    }
    if (odd_row) return 0.5 * (e0[C] + crow(k, R + 1)[C]);
    return e0[C];
  };

  double sxf = 0.0, sxc = 0.0;
  if (ph.mass_mode) {
    sxf = ph.sxf[c];
    sxc = ph.sxc[c];
  }

  // ---- L2 prefetch: ONE instruction per row covers every streamed array (lane -> (array, cache line)) ----
  const double* pfp = nullptr;
  const double* pfp2 = nullptr;
  if (a.g.pf > 0) {
    // row (r + pf) of every array: lane -> (array = lane / 3, one of the <= 3 cache lines the warp's columns touch)
    const int li = lane % 3;
    const int src = (li == 0) ? 0 : (li == 1 ? 16 : 31);
    const int pc = __shfl_sync(kFull, c, src);
    const int arr = lane / 3;  // 0..10
    const size_t pfo = (size_t)a.g.pf * n + pc;
    if (arr == 0) pfp = th + n + pfo;  // theta is padded by one row
    else if (arr <= 4) pfp = (IN == 1 ? a.wd.x : xin.x) + (arr - 1) * fs + pfo;
    else if (arr <= 8) {
      if (MODE != 0 || IN == 1) pfp = (IN == 1 ? xin.x : b) + (arr - 5) * fs + pfo;
    } else if (arr == 9 && WITH_P) {
      pfp = xin.x + 4 * fs + pfo;
    }
    if (EP == 1) {
      if (arr >= 1 && arr <= 4 && a.ce.read_d) pfp2 = a.ce.d + (arr - 1) * fs + pfo;
      else if (arr >= 5 && arr <= 8 && a.ce.read_x) pfp2 = a.ce.xk + (arr - 5) * fs + pfo;
    }
    if (IN == 2 && EP != 1 && !EDGE) {
      // the coarse correction is first touched here as well (every coarse row serves two fine rows and two lanes, but
      // its first read is a DRAM miss in the middle of a step whose fine rows were prefetched): lanes 0..7 cover the
      // <= 2 cache lines per field that the warp's 30 columns map to; the coarse row is added per step
      const int c_first = __shfl_sync(kFull, c, 0), c_last = __shfl_sync(kFull, c, 31);
      if (lane < 8 && a.g.pfc) pfp2 = cin.x + (lane >> 1) * (int)cin.fs + (((lane & 1) ? c_last : c_first) >> 1);
    }
  }

  // EP 2 starts one row early (stores off): the v-type restriction of coarse row r0/2 needs the residual of row r0-1
  // (a slab's first strip cannot: row -1 belongs to the previous rank, which sends its half-sums instead)
  const bool pre_step = (EP == 2) && !(PUSH && r0 == 0);
  int r = pre_step ? r0 - 1 : r0;

  // ---- prologue: rows r-1 and r (r is even whenever row parity matters: IN 2 / EP 2 start strips on even rows,
  //      EP 2 then steps back to the odd row r0-1) ----
  const bool r_odd = (EP == 2);
  auto ldx = [&](int k, int rr, bool odd_row) -> double {
    double v = ldraw(k, rr);
    if (IN == 2) v += pe(k, rr, odd_row);
    return v;
  };
  double th_m = th_row(th, (EDGE && WRAP) ? wrap_row(r - 1) : r - 1, n)[c];
  double th_c = th_row(th, (EDGE && WRAP) ? wrap_row(r) : r, n)[c];
  double un_m = ldx(0, r - 1, !r_odd), vn_m = ldx(1, r - 1, !r_odd), us_m = ldx(2, r - 1, !r_odd), vs_m = ldx(3, r - 1, !r_odd);
  double un_c = ldx(0, r, r_odd), vn_c = ldx(1, r, r_odd), us_c = ldx(2, r, r_odd), vs_c = ldx(3, r, r_odd);
  double p_m = 0.0, p_c = 0.0;
  if (WITH_P) {
    p_m = EDGE ? row_ptr(xin, 4, r - 1, rows, n)[c] : xin.x[(r - 1) * n + c + 4 * fs32];
    p_c = EDGE ? row_ptr(xin, 4, r, rows, n)[c] : xin.x[r * n + c + 4 * fs32];
  }
  const double a_m = th_m + shfl_up1(th_m);
  double a_c = th_c + shfl_up1(th_c);
  double node_c = 0.25 * (a_c + a_m);
  double Tn_c = node_c * ((un_m - un_c) + (vn_c - shfl_up1(vn_c)));
  double Ts_c = (1.0 - node_c) * ((us_m - us_c) + (vs_c - shfl_up1(vs_c)));
  double Qn_m = th_m * ((shfl_dn1(un_m) - un_m) + (vn_c - vn_m));
  double Qs_m = (1.0 - th_m) * ((shfl_dn1(us_m) - us_m) + (vs_c - vs_m));
  double fv_c = 0.5 * (th_c + th_m);
  double Vsum_c = vs_c + fv_c * (vn_c - vs_c);

  // next row (r+1) raw values, loaded one row ahead of use
  double th_p = th_row(th, (EDGE && WRAP) ? wrap_row(r + 1) : r + 1, n)[c];
  double un_p = ldx(0, r + 1, !r_odd), vn_p = ldx(1, r + 1, !r_odd), us_p = ldx(2, r + 1, !r_odd), vs_p = ldx(3, r + 1, !r_odd);
  double p_p = 0.0;
  if (WITH_P) p_p = EDGE ? row_ptr(xin, 4, r + 1, rows, n)[c] : xin.x[(r + 1) * n + c + 4 * fs32];

  // EP 2 carries: half-sums of the residual down the rows
  double ru_n_prev = 0.0, ru_s_prev = 0.0;          // u-type residual of the even row of the current pair
  double tv_n_m1 = 0.0, tv_s_m1 = 0.0;              // v-type column half-sum of row 2R-1
  double tv_n_0 = 0.0, tv_s_0 = 0.0;                // ... of row 2R
  double dtv0_n = 0.0, dtv0_s = 0.0, dtvp_n = 0.0, dtvp_s = 0.0;  // slab: rows 0 / 1 half-sums of the deferred first coarse row

  // One marching step: produce output row r (parity known at compile time per call site), take in row r+2.
  auto step = [&](const bool odd_out, const bool live) {
    // incoming row r+2 (same parity as r), clamped to r1: the last one is unused but stays inside the halo
    const int rq = EDGE ? min(r + 2, r1) : r + 2;
    const double th_q = th_row(th, (EDGE && WRAP) ? wrap_row(rq) : rq, n)[c];
    const double un_q = ldx(0, rq, odd_out), vn_q = ldx(1, rq, odd_out);
    const double us_q = ldx(2, rq, odd_out), vs_q = ldx(3, rq, odd_out);
    double p_q = 0.0;
    if (WITH_P) p_q = EDGE ? row_ptr(xin, 4, rq, rows, n)[c] : xin.x[rq * n + c + 4 * fs32];
    const int off = r * n + c;
    if (a.g.pf > 0 && r + a.g.pf < min(r1 + 1, rows) && r >= 0) {
      if (pfp) pf_l2(pfp + (size_t)r * n);
      if (EP == 1 && pfp2) pf_l2(pfp2 + (size_t)r * n);
      if (IN == 2 && EP != 1 && !EDGE && pfp2) pf_l2(pfp2 + (size_t)((r + a.g.pf) >> 1) * nc);
    }
    double bn_u = 0.0, bn_v = 0.0, bs_u = 0.0, bs_v = 0.0;
    if (MODE != 0) {
      const double* bb = (IN == 1) ? xin.x : b;  // IN 1: the rhs IS the streamed input
      const int ob = (EDGE && WRAP && r < 0) ? wrap_row(r) * n + c : off;
      bn_u = bb[ob];
      bn_v = bb[ob + fs32];
      bs_u = bb[ob + 2 * fs32];
      bs_v = bb[ob + 3 * fs32];
    }

    const double a_p = th_p + shfl_up1(th_p);
    const double node_p = 0.25 * (a_p + a_c);
    const double Tn_p = node_p * ((un_c - un_p) + (vn_p - shfl_up1(vn_p)));
    const double Ts_p = (1.0 - node_p) * ((us_c - us_p) + (vs_p - shfl_up1(vs_p)));
    const double Qn_c = th_c * ((shfl_dn1(un_c) - un_c) + (vn_p - vn_c));
    const double Qs_c = (1.0 - th_c) * ((shfl_dn1(us_c) - us_c) + (vs_p - vs_c));
    const double Lu_n = (Qn_c - shfl_up1(Qn_c)) + (Tn_c - Tn_p);
    const double Lu_s = (Qs_c - shfl_up1(Qs_c)) + (Ts_c - Ts_p);
    const double Lv_n = (shfl_dn1(Tn_c) - Tn_c) + (Qn_c - Qn_m);
    const double Lv_s = (shfl_dn1(Ts_c) - Ts_c) + (Qs_c - Qs_m);

    const double fu_c = 0.5 * a_c;
    double mu, mv;
    if (ph.mass_mode) {
      const int gr = a.g.row0 + ((EDGE && WRAP) ? wrap_row(r) : r);
      mu = 0.25 * sxf * ph.syc[gr] + 0.5;  // thn(-(r+1/2)h, c h), preconditioner.py:325
      mv = 0.25 * sxc * ph.syf[gr] + 0.5;  // thn(-r h, (c+1/2)h), preconditioner.py:326
    } else {
      mu = fu_c;
      mv = fv_c;
    }
    const double dXu = ph.d_u * (ph.xi * fu_c * (1.0 - fu_c));  // preconditioner.py:124
    const double dXv = ph.d_u * (ph.xi * fv_c * (1.0 - fv_c));  // preconditioner.py:125
    const double du = un_c - us_c, dv = vn_c - vs_c;
    const double cmu = ph.c * mu, cmv = ph.c * mv;
    double y_un = cmu * un_c - dXu * du + ph.kap_n * Lu_n;
    double y_us = (ph.c - cmu) * us_c + dXu * du + ph.kap_s * Lu_s;
    double y_vn = cmv * vn_c - dXv * dv + ph.kap_n * Lv_n;
    double y_vs = (ph.c - cmv) * vs_c + dXv * dv + ph.kap_s * Lv_s;
    const double fv_p = 0.5 * (th_p + th_c);
    double Vsum_p = 0.0, y_p = 0.0;
    if (WITH_P) {
      Vsum_p = vs_p + fv_p * (vn_p - vs_p);
      const double gx = ph.dp_h * (p_c - shfl_up1(p_c));  // preconditioner.py:204-210
      const double gy = ph.dp_h * (p_m - p_c);            // preconditioner.py:213-219
      y_un += fu_c * gx;
      y_us += (1.0 - fu_c) * gx;
      y_vn += fv_c * gy;
      y_vs += (1.0 - fv_c) * gy;
      const double Usum_c = us_c + fu_c * du;
      y_p = ph.ddiv_h * ((shfl_dn1(Usum_c) - Usum_c) + (Vsum_c - Vsum_p));  // preconditioner.py:221-238, :312
    }
    if (MODE == 1) {
      y_un = bn_u - y_un;
      y_vn = bn_v - y_vn;
      y_us = bs_u - y_us;
      y_vs = bs_v - y_vs;
    }
    if (MODE == 2) {
      if (IN == 1) {
        // x2 = x1 + wd (b - F x1); wd = omega/diag(F) streamed (its row r was the incoming row two steps ago: L1/L2 hit)
        const double* w0 = a.wd.x;
        y_un = un_c + (bn_u - y_un) * w0[off];
        y_vn = vn_c + (bn_v - y_vn) * w0[off + fs32];
        y_us = us_c + (bs_u - y_us) * w0[off + 2 * fs32];
        y_vs = vs_c + (bs_v - y_vs) * w0[off + 3 * fs32];
      } else {
        const double node_e = shfl_dn1(node_c);
        const double su = a_c + node_c + node_p;           // tE+tW+nN+nS, preconditioner.py:127
        const double sv = th_m + th_c + node_c + node_e;   // tN+tC+nL+nR, preconditioner.py:242
        const double d_un = cmu - dXu - ph.kap_n * su;
        const double d_us = (ph.c - cmu) - dXu - ph.kap_s * (4.0 - su);
        const double d_vn = cmv - dXv - ph.kap_n * sv;
        const double d_vs = (ph.c - cmv) - dXv - ph.kap_s * (4.0 - sv);
        y_un = un_c + a.omega * (bn_u - y_un) * fast_rcp(d_un);
        y_us = us_c + a.omega * (bs_u - y_us) * fast_rcp(d_us);
        y_vn = vn_c + a.omega * (bn_v - y_vn) * fast_rcp(d_vn);
        y_vs = vs_c + a.omega * (bs_v - y_vs) * fast_rcp(d_vs);
      }
    }
    if (EP == 0) {
      if (store && live) {
        double* __restrict__ y = a.y;
        y[off] = y_un;
        y[off + fs32] = y_vn;
        y[off + 2 * fs32] = y_us;
        y[off + 3 * fs32] = y_vs;
        if (WITH_P && MODE == 0) y[off + 4 * fs32] = y_p;
      }
    } else if (EP == 1) {
      if (store && live) {
        // Chebyshev semi-iteration on (V-cycle) o F: z = the sweep's result, never stored
        double* __restrict__ d = a.ce.d;
        double* __restrict__ xk = a.ce.xk;
        const int o0 = off, o1 = off + fs32, o2 = off + 2 * fs32, o3 = off + 3 * fs32;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        if (a.ce.read_d) { d0 = d[o0]; d1 = d[o1]; d2 = d[o2]; d3 = d[o3]; }
        if (a.ce.read_x) { x0 = xk[o0]; x1 = xk[o1]; x2 = xk[o2]; x3 = xk[o3]; }
        d0 = a.ce.ca * d0 + a.ce.cb * y_un;
        d1 = a.ce.ca * d1 + a.ce.cb * y_vn;
        d2 = a.ce.ca * d2 + a.ce.cb * y_us;
        d3 = a.ce.ca * d3 + a.ce.cb * y_vs;
        if (a.ce.write_d) { d[o0] = d0; d[o1] = d1; d[o2] = d2; d[o3] = d3; }
        y_un = x0 + d0;  // (the iterate is what a fused halo push sends on)
        y_vn = x1 + d1;
        y_us = x2 + d2;
        y_vs = x3 + d3;
        xk[o0] = y_un;
        xk[o1] = y_vn;
        xk[o2] = y_us;
        xk[o3] = y_vs;
      }
    } else {
      // EP 2: full weighting onto the coarse grid (k_restrict_F's weights): u: (1/4,1/2,1/4) over columns x (1/2,1/2)
      // over the row pair; v: (1/2,1/2) over the column pair x (1/4,1/2,1/4) over rows 2R-1, 2R, 2R+1
      const double tvn = 0.5 * (y_vn + shfl_dn1(y_vn));
      const double tvs = 0.5 * (y_vs + shfl_dn1(y_vs));
      if (!odd_out) {
        ru_n_prev = y_un;
        ru_s_prev = y_us;
        tv_n_0 = tvn;
        tv_s_0 = tvs;
      } else {
        const double sun = 0.5 * (ru_n_prev + y_un), sus = 0.5 * (ru_s_prev + y_us);
        const double cu_n = 0.25 * shfl_up1(sun) + 0.5 * sun + 0.25 * shfl_dn1(sun);
        const double cu_s = 0.25 * shfl_up1(sus) + 0.5 * sus + 0.25 * shfl_dn1(sus);
        // slab (PUSH): the v-type weights of the first coarse row need fine row -1, which the previous rank owns
        const bool defer = PUSH && EDGE && r == 1;
        const double cv_n = 0.25 * tv_n_m1 + 0.5 * tv_n_0 + 0.25 * tvn;
        const double cv_s = 0.25 * tv_s_m1 + 0.5 * tv_s_0 + 0.25 * tvs;
        if (defer) {
          dtv0_n = tv_n_0; dtv0_s = tv_s_0; dtvp_n = tvn; dtvp_s = tvs;
        }
        if (store && live && !codd) {
          double* __restrict__ bc = a.bc;
          const int oc = (r >> 1) * nc + C;
          bc[oc] = cu_n;
          bc[oc + 2 * fsc32] = cu_s;
          if (!defer) {
            bc[oc + fsc32] = cv_n;
            bc[oc + 3 * fsc32] = cv_s;
          }
          if (PUSH && EDGE) {
            // the coarse rhs is the next level's pre-smoother input: its first / last rows go to the ring neighbours
            if (r == 1) {
              st_ll(push_prev + C, cu_n, push_seq);
              st_ll(push_prev + 2 * nc + C, cu_s, push_seq);
            }
            if (r == rows - 1) {
              st_ll(push_next + C, cu_n, push_seq);
              st_ll(push_next + nc + C, cv_n, push_seq);
              st_ll(push_next + 2 * nc + C, cu_s, push_seq);
              st_ll(push_next + 3 * nc + C, cv_s, push_seq);
              // ... and this row's half-sums complete the next rank's first coarse row
              st_ll(push_next + 4 * nc + C, tvn, push_seq);
              st_ll(push_next + 5 * nc + C, tvs, push_seq);
            }
          }
        }
        tv_n_m1 = tvn;
        tv_s_m1 = tvs;
      }
    }
    if (PUSH && EDGE && EP != 2 && store && live) {
      if (r == 0) {
        st_ll(push_prev + c, y_un, push_seq);
        st_ll(push_prev + n + c, y_vn, push_seq);
        st_ll(push_prev + 2 * n + c, y_us, push_seq);
        st_ll(push_prev + 3 * n + c, y_vs, push_seq);
      }
      if (r == rows - 1) {
        st_ll(push_next + c, y_un, push_seq);
        st_ll(push_next + n + c, y_vn, push_seq);
        st_ll(push_next + 2 * n + c, y_us, push_seq);
        st_ll(push_next + 3 * n + c, y_vs, push_seq);
      }
    }
    // rotate the window
    th_m = th_c; th_c = th_p; th_p = th_q;
    a_c = a_p; node_c = node_p;
    un_c = un_p; vn_c = vn_p; us_c = us_p; vs_c = vs_p;
    un_p = un_q; vn_p = vn_q; us_p = us_q; vs_p = vs_q;
    if (WITH_P) { p_m = p_c; p_c = p_p; p_p = p_q; Vsum_c = Vsum_p; }
    Tn_c = Tn_p; Ts_c = Ts_p; Qn_m = Qn_c; Qs_m = Qs_c;
    fv_c = fv_p;
    ++r;
  };

  if (pre_step) step(true, false);  // row r0-1 (odd): only its v-type half-sums are kept
  if (IN == 2 || EP == 2) {
    // row pairs: strips start on even rows and coarsened levels have even row counts
    while (r < r1) {
      step(false, true);
      step(true, true);
    }
  } else {
    while (r < r1) {
      step(false, true);
      if (r < r1) step(true, true);
    }
  }

  if (EP == 2 && PUSH && EDGE && r0 == 0 && store && !codd) {
    // first coarse row of the slab, v-type fields: the half-sums of fine row -1 arrive from the previous rank (its last
    // strip sends them under this kernel's own exchange number; both edge strips run in the first wave on every rank)
    const LLElem* mine = comm_halo(a.po.my_comm, a.po.area, (int)(push_seq % (unsigned long long)kHaloSlots), 0);
    const double tm_n = ll_wait(mine + 4 * nc + C, push_seq);
    const double tm_s = ll_wait(mine + 5 * nc + C, push_seq);
    const double cv_n = 0.25 * tm_n + 0.5 * dtv0_n + 0.25 * dtvp_n;
    const double cv_s = 0.25 * tm_s + 0.5 * dtv0_s + 0.25 * dtvp_s;
    a.bc[C + fsc32] = cv_n;
    a.bc[C + 3 * fsc32] = cv_s;
    st_ll(push_prev + nc + C, cv_n, push_seq);
    st_ll(push_prev + 3 * nc + C, cv_s, push_seq);
  }
  if (PUSH) push_end(a.po, pc, r0 == 0, r1 == rows, gridDim.x, gridDim.x, gridDim.y == 1);
}

template <int IN, int MODE, bool WITH_P, int EP, bool PUSH, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB) k_stokes_x(const __grid_constant__ StokesArgs a) {
  const int rows = a.g.rows;
  // strip order: first strip, LAST strip, then the interior -- both edge strips (the ones that wait for the ring
  // neighbours' rows and push this rank's own) run in the first wave and their transfers overlap the interior
  int r0, r1;
  if (!strip_rows(a.g, r0, r1)) return;
  // interior strips touch rows r0-2 .. r1+1 (and the coarse rows under them) only: all inside the slab
  if (r0 >= 2 && r1 + 4 <= rows) stokes_march<IN, MODE, WITH_P, EP, PUSH, false>(a, r0, r1);
  else stokes_march<IN, MODE, WITH_P, EP, PUSH, true>(a, r0, r1);
}

}  // namespace mpbp
