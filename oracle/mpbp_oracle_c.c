/*
 * mpbp_oracle_c.c -- multi-threaded (OpenMP) C restatement of the reference's hot path.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (cross-check of the numpy oracle) and by
 * bench.py's cpu_baseline / --impl reference legs (the reference's CPU path "with all the host threads it
 * can use").  The product package never links or loads it.
 *
 * Independent of the CUDA kernels' flux form on purpose: the viscous block is written with the
 * reference's own coefficient table (preconditioner.py:127-179 for u-rows, :242-295 for v-rows), the
 * gradient / divergence from :203-238, the drag from :124-125, the mass term from :325-329, the system
 * assembly from :331-341, the preconditioner structure from solve.py:257-277, relaxation from
 * solve.py:149-159 and a textbook right-preconditioned FGMRES with the call shape of solve.py:285.
 * The sub-solver definition (rediscretised V(nu1,nu2) multigrid, Chebyshev acceleration) mirrors
 * oracle/mpbp_oracle.py operation for operation.  Pinned against the numpy oracle (itself pinned against
 * the reference's golden vectors) in tests/test_c_oracle.py.
 *
 * Layout: vectors [u_n | v_n | u_s | v_s | p], each n x n row-major (preconditioner.py:100-106).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_MAX_LEVELS 16

typedef struct {
  int n;
  double h;
  int mass_analytic;     /* 1: c*thn evaluated analytically at the faces (level 0 of the reference problem) */
  double* theta;         /* n*n cell-centred theta_n */
  double* mass_u;        /* n*n theta_n at u faces (mass term) */
  double* mass_v;        /* n*n theta_n at v faces */
  double* node;          /* n*n corner average of theta_n at the top-left corner of cell (r,c) (:112, :195) */
  int *im, *ip;          /* periodic index tables: im[i] = (i-1) mod n, ip[i] = (i+1) mod n */
  /* multigrid work vectors */
  double *bF, *xF, *tF, *rF; /* 4N */
  double *bP, *xP, *tP, *rP; /* N */
} oc_level;

typedef struct {
  int nlev;
  oc_level lev[OC_MAX_LEVELS];
  double xi, eta_n, eta_s, c, d_u, d_p, d_div;
  /* sub-solver */
  int kind;            /* 0 jacobi, 1 mg */
  int F_cycles, P_cycles, F_sweeps, P_sweeps, nu1, nu2, cheb, project;
  double omega, lmin, lmax;
  double *Finv, *Pinv; /* dense (pseudo-)inverses on the coarsest level, row-major */
  int mF, mP;
  /* level-0 scratch */
  double *w, *g, *t2, *rinF, *zF, *dvF, *rhs, *xa, *xb, *rinP, *zP, *dvP;
} oc_ctx;

static const double OC_PI = 3.141592653589793;

static inline int wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }
#define TH(r, c) th[(size_t)wrap((r), n) * n + wrap((c), n)]
#define AT(f, r, c) (f)[(size_t)wrap((r), n) * n + wrap((c), n)]

static double thn(double y, double x) { return 0.25 * sin(2 * OC_PI * x) * sin(2 * OC_PI * y) + 0.5; } /* preconditioner.py:9-11 */

/* ---- one phase of the viscous operator with the reference's coefficient table ---- */
/* s = 0: theta_n, s = 1: theta_s = 1 - theta_n (preconditioner.py:74-81) */
static inline double ph(double t, int s) { return s ? 1.0 - t : t; }

/* coefficient-table form, preconditioner.py:127-179 (u rows) and :242-295 (v rows); indices pre-wrapped */
typedef struct { int r, c, rm, rp, cm, cp; } oc_idx;
#define IX(rr, cc) ((size_t)(rr) * n + (cc))
static inline double L_u(const oc_level* L, int s, const double* u, const double* v, oc_idx i) {
  const int n = L->n;
  const double* th = L->theta;
  const double tE = ph(th[IX(i.r, i.c)], s), tW = ph(th[IX(i.r, i.cm)], s);
  const double nN = ph(L->node[IX(i.r, i.c)], s), nS = ph(L->node[IX(i.rp, i.c)], s);
  return tW * u[IX(i.r, i.cm)] + tE * u[IX(i.r, i.cp)] + nN * u[IX(i.rm, i.c)] + nS * u[IX(i.rp, i.c)] -
         (tE + tW + nN + nS) * u[IX(i.r, i.c)] + (nN - tE) * v[IX(i.r, i.c)] + (tW - nN) * v[IX(i.r, i.cm)] +
         (nS - tW) * v[IX(i.rp, i.cm)] + (tE - nS) * v[IX(i.rp, i.c)];
}
static inline double L_v(const oc_level* L, int s, const double* u, const double* v, oc_idx i) {
  const int n = L->n;
  const double* th = L->theta;
  const double tC = ph(th[IX(i.r, i.c)], s), tN = ph(th[IX(i.rm, i.c)], s);
  const double nL = ph(L->node[IX(i.r, i.c)], s), nR = ph(L->node[IX(i.r, i.cp)], s);
  return nL * v[IX(i.r, i.cm)] + nR * v[IX(i.r, i.cp)] + tN * v[IX(i.rm, i.c)] + tC * v[IX(i.rp, i.c)] -
         (tN + tC + nL + nR) * v[IX(i.r, i.c)] + (nL - tC) * u[IX(i.r, i.c)] + (tC - nR) * u[IX(i.r, i.cp)] +
         (tN - nL) * u[IX(i.rm, i.c)] + (nR - tN) * u[IX(i.rm, i.cp)];
}
/* The same two rows evaluated differences-first (stress-divergence form): algebraically identical to the table above,
 * but the rounding error is relative to the DIFFERENCES of neighbouring values, not to the values -- on smooth fields
 * orders of magnitude smaller.  Selected with oc_set_form(1); used to measure how much of a residual history is an
 * artefact of the evaluation order (tests/golden/make_large.py), never as the reference definition.
 *   Q[r,c] = th[r,c]   ((u[r,c+1]-u[r,c]) + (v[r+1,c]-v[r,c]))     T[r,c] = node[r,c] ((u[r-1,c]-u[r,c]) + (v[r,c]-v[r,c-1]))
 *   (L u)_u = (Q[r,c]-Q[r,c-1]) + (T[r,c]-T[r+1,c])                (L u)_v = (T[r,c+1]-T[r,c]) + (Q[r,c]-Q[r-1,c])      */
static int g_form = 0;
void oc_set_form(int f) { g_form = f; }
static inline double L_u_flux(const oc_level* L, int s, const double* u, const double* v, oc_idx i) {
  const int n = L->n;
  const double* th = L->theta;
  const double Qc = ph(th[IX(i.r, i.c)], s) * ((u[IX(i.r, i.cp)] - u[IX(i.r, i.c)]) + (v[IX(i.rp, i.c)] - v[IX(i.r, i.c)]));
  const double Qw = ph(th[IX(i.r, i.cm)], s) * ((u[IX(i.r, i.c)] - u[IX(i.r, i.cm)]) + (v[IX(i.rp, i.cm)] - v[IX(i.r, i.cm)]));
  const double Tc = ph(L->node[IX(i.r, i.c)], s) * ((u[IX(i.rm, i.c)] - u[IX(i.r, i.c)]) + (v[IX(i.r, i.c)] - v[IX(i.r, i.cm)]));
  const double Ts = ph(L->node[IX(i.rp, i.c)], s) * ((u[IX(i.r, i.c)] - u[IX(i.rp, i.c)]) + (v[IX(i.rp, i.c)] - v[IX(i.rp, i.cm)]));
  return (Qc - Qw) + (Tc - Ts);
}
static inline double L_v_flux(const oc_level* L, int s, const double* u, const double* v, oc_idx i) {
  const int n = L->n;
  const double* th = L->theta;
  const double Te = ph(L->node[IX(i.r, i.cp)], s) * ((u[IX(i.rm, i.cp)] - u[IX(i.r, i.cp)]) + (v[IX(i.r, i.cp)] - v[IX(i.r, i.c)]));
  const double Tc = ph(L->node[IX(i.r, i.c)], s) * ((u[IX(i.rm, i.c)] - u[IX(i.r, i.c)]) + (v[IX(i.r, i.c)] - v[IX(i.r, i.cm)]));
  const double Qc = ph(th[IX(i.r, i.c)], s) * ((u[IX(i.r, i.cp)] - u[IX(i.r, i.c)]) + (v[IX(i.rp, i.c)] - v[IX(i.r, i.c)]));
  const double Qn = ph(th[IX(i.rm, i.c)], s) * ((u[IX(i.rm, i.cp)] - u[IX(i.rm, i.c)]) + (v[IX(i.r, i.c)] - v[IX(i.rm, i.c)]));
  return (Te - Tc) + (Qc - Qn);
}
static inline double L_u_diag(const oc_level* L, int s, oc_idx i) {
  const int n = L->n;
  return -(ph(L->theta[IX(i.r, i.c)], s) + ph(L->theta[IX(i.r, i.cm)], s) + ph(L->node[IX(i.r, i.c)], s) +
           ph(L->node[IX(i.rp, i.c)], s));
}
static inline double L_v_diag(const oc_level* L, int s, oc_idx i) {
  const int n = L->n;
  return -(ph(L->theta[IX(i.rm, i.c)], s) + ph(L->theta[IX(i.r, i.c)], s) + ph(L->node[IX(i.r, i.c)], s) +
           ph(L->node[IX(i.r, i.cp)], s));
}

/* y = F x (mode 0), b - F x (mode 1), x + omega (b - F x)/diag (mode 2); with_p adds G p and the p row of A */
static void stokes_op(const oc_ctx* C, const oc_level* L, int mode, int with_p, const double* x, const double* b,
                      double* y, double omega) {
  const int n = L->n;
  const size_t N = (size_t)n * n;
  const double* th = L->theta;
  const double ih = 1.0 / L->h, ih2 = ih * ih;
  const double* un = x;
  const double* vn = x + N;
  const double* us = x + 2 * N;
  const double* vs = x + 3 * N;
  const double* p = with_p ? x + 4 * N : NULL;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) {
      const size_t k = (size_t)r * n + c;
      const oc_idx ix = {r, c, L->im[r], L->ip[r], L->im[c], L->ip[c]};
      const double fu = 0.5 * (th[k] + th[IX(r, ix.cm)]); /* :114 */
      const double fv = 0.5 * (th[k] + th[IX(ix.rm, c)]); /* :120 */
      const double Xu = C->xi * fu * (1.0 - fu);          /* :124 */
      const double Xv = C->xi * fv * (1.0 - fv);          /* :125 */
      const double mu = L->mass_u[k], mv = L->mass_v[k];  /* :325-326 */
      const double d = C->d_u;
      /* F = XI_block + d_u * blockdiag(eta_n L_n, eta_s L_s), :331-337 */
      double y_un, y_vn, y_us, y_vs;
      if (g_form == 0) {
        y_un = (C->c * mu - d * Xu) * un[k] + d * Xu * us[k] + d * C->eta_n * ih2 * L_u(L, 0, un, vn, ix);
        y_vn = (C->c * mv - d * Xv) * vn[k] + d * Xv * vs[k] + d * C->eta_n * ih2 * L_v(L, 0, un, vn, ix);
        y_us = (C->c * (1.0 - mu) - d * Xu) * us[k] + d * Xu * un[k] + d * C->eta_s * ih2 * L_u(L, 1, us, vs, ix);
        y_vs = (C->c * (1.0 - mv) - d * Xv) * vs[k] + d * Xv * vn[k] + d * C->eta_s * ih2 * L_v(L, 1, us, vs, ix);
      } else {
        const double du = un[k] - us[k], dv = vn[k] - vs[k];
        y_un = C->c * mu * un[k] - d * Xu * du + d * C->eta_n * ih2 * L_u_flux(L, 0, un, vn, ix);
        y_vn = C->c * mv * vn[k] - d * Xv * dv + d * C->eta_n * ih2 * L_v_flux(L, 0, un, vn, ix);
        y_us = C->c * (1.0 - mu) * us[k] + d * Xu * du + d * C->eta_s * ih2 * L_u_flux(L, 1, us, vs, ix);
        y_vs = C->c * (1.0 - mv) * vs[k] + d * Xv * dv + d * C->eta_s * ih2 * L_v_flux(L, 1, us, vs, ix);
      }
      if (with_p) {
        const double gx = C->d_p * ih * (p[k] - p[IX(r, ix.cm)]);  /* :204-210 */
        const double gy = C->d_p * ih * (p[IX(ix.rm, c)] - p[k]);  /* :213-219 */
        y_un += fu * gx;
        y_us += (1.0 - fu) * gx;
        y_vn += fv * gy;
        y_vs += (1.0 - fv) * gy;
      }
      if (mode == 1) {
        y_un = b[k] - y_un;
        y_vn = b[k + N] - y_vn;
        y_us = b[k + 2 * N] - y_us;
        y_vs = b[k + 3 * N] - y_vs;
      } else if (mode == 2) {
        const double dun = C->c * mu - d * Xu + d * C->eta_n * ih2 * L_u_diag(L, 0, ix);
        const double dvn = C->c * mv - d * Xv + d * C->eta_n * ih2 * L_v_diag(L, 0, ix);
        const double dus = C->c * (1.0 - mu) - d * Xu + d * C->eta_s * ih2 * L_u_diag(L, 1, ix);
        const double dvs = C->c * (1.0 - mv) - d * Xv + d * C->eta_s * ih2 * L_v_diag(L, 1, ix);
        y_un = un[k] + omega * (b[k] - y_un) / dun;           /* solve.py:158 (damped) */
        y_vn = vn[k] + omega * (b[k + N] - y_vn) / dvn;
        y_us = us[k] + omega * (b[k + 2 * N] - y_us) / dus;
        y_vs = vs[k] + omega * (b[k + 3 * N] - y_vs) / dvs;
      }
      y[k] = y_un;
      y[k + N] = y_vn;
      y[k + 2 * N] = y_us;
      y[k + 3 * N] = y_vs;
      if (with_p) {
        /* d_div * (D_n [u_n v_n] + D_s [u_s v_s]), :221-238, :311-312 */
        const double fuE = 0.5 * (th[IX(r, ix.cp)] + th[k]);
        const double fvS = 0.5 * (th[IX(ix.rp, c)] + th[k]);
        const double dn = ih * (fuE * un[IX(r, ix.cp)] - fu * un[k]) + ih * (fv * vn[k] - fvS * vn[IX(ix.rp, c)]);
        const double ds = ih * ((1.0 - fuE) * us[IX(r, ix.cp)] - (1.0 - fu) * us[k]) +
                          ih * ((1.0 - fv) * vs[k] - (1.0 - fvS) * vs[IX(ix.rp, c)]);
        y[k + 4 * N] = C->d_div * (dn + ds);
      }
    }
  }
}

/* r = scale * D w + add (D un-negated as returned at preconditioner.py:349) */
static void div_op(const oc_level* L, const double* w, const double* add, double* out, double scale) {
  const int n = L->n;
  const size_t N = (size_t)n * n;
  const double* th = L->theta;
  const double ih = 1.0 / L->h;
  const double *un = w, *vn = w + N, *us = w + 2 * N, *vs = w + 3 * N;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      const size_t k = (size_t)r * n + c;
      const double fu = 0.5 * (TH(r, c) + TH(r, c - 1)), fuE = 0.5 * (TH(r, c + 1) + TH(r, c));
      const double fv = 0.5 * (TH(r, c) + TH(r - 1, c)), fvS = 0.5 * (TH(r + 1, c) + TH(r, c));
      const double dn = ih * (fuE * AT(un, r, c + 1) - fu * un[k]) + ih * (fv * vn[k] - fvS * AT(vn, r + 1, c));
      const double ds = ih * ((1.0 - fuE) * AT(us, r, c + 1) - (1.0 - fu) * us[k]) +
                        ih * ((1.0 - fv) * vs[k] - (1.0 - fvS) * AT(vs, r + 1, c));
      out[k] = scale * (dn + ds) + (add ? add[k] : 0.0);
    }
}

/* y = G p = d_p [G_n; G_s] p (preconditioner.py:203-219, :313) */
static void grad_op(const oc_ctx* C, const oc_level* L, const double* p, double* y) {
  const int n = L->n;
  const size_t N = (size_t)n * n;
  const double* th = L->theta;
  const double ih = 1.0 / L->h;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      const size_t k = (size_t)r * n + c;
      const double fu = 0.5 * (TH(r, c) + TH(r, c - 1)), fv = 0.5 * (TH(r, c) + TH(r - 1, c));
      const double gx = C->d_p * ih * (p[k] - AT(p, r, c - 1));
      const double gy = C->d_p * ih * (AT(p, r - 1, c) - p[k]);
      y[k] = fu * gx;
      y[k + N] = fv * gy;
      y[k + 2 * N] = (1.0 - fu) * gx;
      y[k + 3 * N] = (1.0 - fv) * gy;
    }
}

/* Gt_G = (-D) G (solve.py:246-247) as the 5-point operator it is; modes as stokes_op, 3: omega b / diag */
static void poisson_op(const oc_ctx* C, const oc_level* L, int mode, const double* p, const double* b, double* y,
                       double omega) {
  const int n = L->n;
  const double* th = L->theta;
  const double ih2 = 1.0 / (L->h * L->h);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      const size_t k = (size_t)r * n + c;
      const double fu = 0.5 * (TH(r, c) + TH(r, c - 1)), fuE = 0.5 * (TH(r, c + 1) + TH(r, c));
      const double fv = 0.5 * (TH(r, c) + TH(r - 1, c)), fvS = 0.5 * (TH(r + 1, c) + TH(r, c));
      const double wu = fu * fu + (1 - fu) * (1 - fu), wuE = fuE * fuE + (1 - fuE) * (1 - fuE);
      const double wv = fv * fv + (1 - fv) * (1 - fv), wvS = fvS * fvS + (1 - fvS) * (1 - fvS);
      double out = 0.0;
      if (mode != 3)
        out = -C->d_p * ih2 * (wuE * (AT(p, r, c + 1) - p[k]) - wu * (p[k] - AT(p, r, c - 1)) +
                               wv * (AT(p, r - 1, c) - p[k]) - wvS * (p[k] - AT(p, r + 1, c)));
      const double dg = C->d_p * ih2 * (wuE + wu + wv + wvS);
      if (mode == 1) out = b[k] - out;
      else if (mode == 2) out = p[k] + omega * (b[k] - out) / dg;
      else if (mode == 3) out = omega * b[k] / dg;
      y[k] = out;
    }
}

/* ---- grid transfers (mirror oracle/mpbp_oracle.py restrict_u/v, prolong_u/v, restrict_cell, prolong_cell) ---- */
static void restrict_F(const double* f, double* yc, int nf) {
  const int n = nf, nc = nf / 2;
  const size_t Nf = (size_t)nf * nf, Nc = (size_t)nc * nc;
#pragma omp parallel for schedule(static)
  for (int R = 0; R < nc; ++R)
    for (int Cc = 0; Cc < nc; ++Cc) {
      for (int s = 0; s < 2; ++s) {
        const double* u = f + (size_t)(2 * s) * Nf;
        const double* v = f + (size_t)(2 * s + 1) * Nf;
        const int c0 = 2 * Cc, ra = 2 * R, rb = 2 * R + 1;
        const double um = 0.5 * (AT(u, ra, c0 - 1) + AT(u, rb, c0 - 1)), u0 = 0.5 * (AT(u, ra, c0) + AT(u, rb, c0)),
                     up = 0.5 * (AT(u, ra, c0 + 1) + AT(u, rb, c0 + 1));
        yc[(size_t)(2 * s) * Nc + (size_t)R * nc + Cc] = 0.25 * um + 0.5 * u0 + 0.25 * up;
        const double wm = 0.5 * (AT(v, ra - 1, c0) + AT(v, ra - 1, c0 + 1)), w0 = 0.5 * (AT(v, ra, c0) + AT(v, ra, c0 + 1)),
                     wp = 0.5 * (AT(v, rb, c0) + AT(v, rb, c0 + 1));
        yc[(size_t)(2 * s + 1) * Nc + (size_t)R * nc + Cc] = 0.25 * wm + 0.5 * w0 + 0.25 * wp;
      }
    }
}
static void prolong_add_F(const double* xc, double* xf, int nf) {
  const int nc = nf / 2;
  const size_t Nf = (size_t)nf * nf, Nc = (size_t)nc * nc;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nf; ++r)
    for (int c = 0; c < nf; ++c) {
      const int R = r >> 1, Cc = c >> 1, Cp = (Cc + 1) % nc, Rp = (R + 1) % nc;
      for (int s = 0; s < 2; ++s) {
        const double* uc = xc + (size_t)(2 * s) * Nc;
        const double* vc = xc + (size_t)(2 * s + 1) * Nc;
        const double eu = (c & 1) ? 0.5 * (uc[(size_t)R * nc + Cc] + uc[(size_t)R * nc + Cp]) : uc[(size_t)R * nc + Cc];
        const double ev = (r & 1) ? 0.5 * (vc[(size_t)R * nc + Cc] + vc[(size_t)Rp * nc + Cc]) : vc[(size_t)R * nc + Cc];
        xf[(size_t)(2 * s) * Nf + (size_t)r * nf + c] += eu;
        xf[(size_t)(2 * s + 1) * Nf + (size_t)r * nf + c] += ev;
      }
    }
}
static void restrict_P(const double* f, double* yc, int nf) {
  const int nc = nf / 2;
#pragma omp parallel for schedule(static)
  for (int R = 0; R < nc; ++R)
    for (int Cc = 0; Cc < nc; ++Cc) {
      const double* a = f + (size_t)(2 * R) * nf + 2 * Cc;
      yc[(size_t)R * nc + Cc] = 0.25 * (a[0] + a[1] + a[nf] + a[nf + 1]);
    }
}
static void prolong_add_P(const double* xc, double* xf, int nf) {
  const int nc = nf / 2;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nf; ++r)
    for (int c = 0; c < nf; ++c) xf[(size_t)r * nf + c] += xc[(size_t)(r >> 1) * nc + (c >> 1)];
}

/* ---- small helpers ---- */
static void vcopy(const double* x, double* y, size_t len) { memcpy(y, x, len * sizeof(double)); }
static void axpby(double a, const double* x, double b, const double* y, double* z, size_t len) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)len; ++i) z[i] = a * x[i] + b * y[i];
}
static double vdot(const double* x, const double* y, size_t len) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (long long i = 0; i < (long long)len; ++i) s += x[i] * y[i];
  return s;
}
static void dense_mv(const double* M, const double* x, double* y, int m) {
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int k = 0; k < m; ++k) s += M[(size_t)i * m + k] * x[k];
    y[i] = s;
  }
}
static int invert_dense(double* A, int m) { /* Gauss-Jordan with partial pivoting, in place */
  double* I = (double*)calloc((size_t)m * m, sizeof(double));
  for (int i = 0; i < m; ++i) I[(size_t)i * m + i] = 1.0;
  for (int col = 0; col < m; ++col) {
    int piv = col;
    double best = fabs(A[(size_t)col * m + col]);
    for (int r = col + 1; r < m; ++r)
      if (fabs(A[(size_t)r * m + col]) > best) best = fabs(A[(size_t)r * m + col]), piv = r;
    if (best == 0.0) { free(I); return 1; }
    if (piv != col)
      for (int k = 0; k < m; ++k) {
        double t = A[(size_t)piv * m + k]; A[(size_t)piv * m + k] = A[(size_t)col * m + k]; A[(size_t)col * m + k] = t;
        t = I[(size_t)piv * m + k]; I[(size_t)piv * m + k] = I[(size_t)col * m + k]; I[(size_t)col * m + k] = t;
      }
    const double d = 1.0 / A[(size_t)col * m + col];
    for (int k = 0; k < m; ++k) A[(size_t)col * m + k] *= d, I[(size_t)col * m + k] *= d;
    for (int r = 0; r < m; ++r) {
      if (r == col) continue;
      const double f = A[(size_t)r * m + col];
      if (f == 0.0) continue;
      for (int k = 0; k < m; ++k) A[(size_t)r * m + k] -= f * A[(size_t)col * m + k], I[(size_t)r * m + k] -= f * I[(size_t)col * m + k];
    }
  }
  memcpy(A, I, (size_t)m * m * sizeof(double));
  free(I);
  return 0;
}

/* ---- multigrid ---- */
static void vcycle(oc_ctx* C, int l, int isF, const double* b, double* x) {
  oc_level* L = &C->lev[l];
  const size_t len = (size_t)(isF ? 4 : 1) * L->n * L->n;
  if (l == C->nlev - 1) {
    dense_mv(isF ? C->Finv : C->Pinv, b, x, isF ? C->mF : C->mP);
    return;
  }
  double* t = isF ? L->tF : L->tP;
  double* r = isF ? L->rF : L->rP;
  /* x = omega b / diag: a Jacobi sweep from x = 0 */
  if (isF) {
    memset(x, 0, len * sizeof(double));
    stokes_op(C, L, 2, 0, x, b, t, C->omega);
    vcopy(t, x, len);
  } else {
    poisson_op(C, L, 3, NULL, b, x, C->omega);
  }
  for (int s = 1; s < C->nu1; ++s) {
    if (isF) stokes_op(C, L, 2, 0, x, b, t, C->omega); else poisson_op(C, L, 2, x, b, t, C->omega);
    vcopy(t, x, len);
  }
  if (isF) stokes_op(C, L, 1, 0, x, b, r, 0.0); else poisson_op(C, L, 1, x, b, r, 0.0);
  oc_level* Lc = &C->lev[l + 1];
  if (isF) restrict_F(r, Lc->bF, L->n); else restrict_P(r, Lc->bP, L->n);
  vcycle(C, l + 1, isF, isF ? Lc->bF : Lc->bP, isF ? Lc->xF : Lc->xP);
  if (isF) prolong_add_F(Lc->xF, x, L->n); else prolong_add_P(Lc->xP, x, L->n);
  for (int s = 0; s < C->nu2; ++s) {
    if (isF) stokes_op(C, L, 2, 0, x, b, t, C->omega); else poisson_op(C, L, 2, x, b, t, C->omega);
    vcopy(t, x, len);
  }
}

static void sub_solve(oc_ctx* C, int isF, const double* b, double* x) {
  oc_level* L = &C->lev[0];
  const size_t len = (size_t)(isF ? 4 : 1) * L->n * L->n;
  double* t = isF ? L->tF : L->tP;
  if (C->kind == 0) {
    const int sweeps = isF ? C->F_sweeps : C->P_sweeps;
    memset(x, 0, len * sizeof(double));
    for (int s = 0; s < sweeps; ++s) {
      if (isF) stokes_op(C, L, 2, 0, x, b, t, C->omega); else poisson_op(C, L, 2, x, b, t, C->omega);
      vcopy(t, x, len);
    }
  } else {
    const int cycles = isF ? C->F_cycles : C->P_cycles;
    double* rin = isF ? C->rinF : C->rinP;
    double* z = isF ? C->zF : C->zP;
    double* dv = isF ? C->dvF : C->dvP;
    if (!C->cheb) {
      vcycle(C, 0, isF, b, x);
      for (int k = 1; k < cycles; ++k) {
        if (isF) stokes_op(C, L, 1, 0, x, b, rin, 0.0); else poisson_op(C, L, 1, x, b, rin, 0.0);
        vcycle(C, 0, isF, rin, z);
        axpby(1.0, x, 1.0, z, x, len);
      }
    } else {
      const double th = 0.5 * (C->lmax + C->lmin), de = 0.5 * (C->lmax - C->lmin), sig = th / de;
      double rho_k = 1.0 / sig;
      vcycle(C, 0, isF, b, z);
      axpby(1.0 / th, z, 0.0, z, dv, len);
      vcopy(dv, x, len);
      for (int k = 1; k < cycles; ++k) {
        if (isF) stokes_op(C, L, 1, 0, x, b, rin, 0.0); else poisson_op(C, L, 1, x, b, rin, 0.0);
        vcycle(C, 0, isF, rin, z);
        const double rho_n = 1.0 / (2.0 * sig - rho_k);
        axpby(rho_n * rho_k, dv, 2.0 * rho_n / de, z, dv, len);
        axpby(1.0, x, 1.0, dv, x, len);
        rho_k = rho_n;
      }
    }
  }
  if (!isF && C->project) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (long long i = 0; i < (long long)len; ++i) s += x[i];
    const double m = s / (double)len;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)len; ++i) x[i] -= m;
  }
}

/* ---- exported API ---- */
void oc_destroy(oc_ctx* C);

oc_ctx* oc_create(int n, double xi, double eta_n, double eta_s, double c, double d_u, double d_p, double d_div,
                  const double* theta /* NULL = analytic */, int kind, int F_cycles, int P_cycles, int F_sweeps,
                  int P_sweeps, double omega, int nu1, int nu2, int n_coarse, int cheb, double lmin, double lmax,
                  int project) {
  oc_ctx* C = (oc_ctx*)calloc(1, sizeof(oc_ctx));
  C->xi = xi; C->eta_n = eta_n; C->eta_s = eta_s; C->c = c; C->d_u = d_u; C->d_p = d_p; C->d_div = d_div;
  C->kind = kind; C->F_cycles = F_cycles; C->P_cycles = P_cycles; C->F_sweeps = F_sweeps; C->P_sweeps = P_sweeps;
  C->omega = omega; C->nu1 = nu1; C->nu2 = nu2; C->cheb = cheb; C->lmin = lmin; C->lmax = lmax; C->project = project;
  int cur = n, l = 0;
  while (1) {
    oc_level* L = &C->lev[l];
    const size_t N = (size_t)cur * cur;
    L->n = cur;
    L->h = 1.0 / cur;
    L->theta = (double*)malloc(N * sizeof(double));
    L->mass_u = (double*)malloc(N * sizeof(double));
    L->mass_v = (double*)malloc(N * sizeof(double));
    if (l == 0) {
      for (int r = 0; r < cur; ++r)
        for (int cc = 0; cc < cur; ++cc)
          L->theta[(size_t)r * cur + cc] = theta ? theta[(size_t)r * cur + cc] : thn(-(r + 0.5) * L->h, (cc + 0.5) * L->h);
    } else {
      const oc_level* Lf = &C->lev[l - 1];
      const int nf = Lf->n;
      for (int R = 0; R < cur; ++R)
        for (int Cc = 0; Cc < cur; ++Cc) {
          const double* a = Lf->theta + (size_t)(2 * R) * nf + 2 * Cc;
          L->theta[(size_t)R * cur + Cc] = 0.25 * (a[0] + a[1] + a[nf] + a[nf + 1]);
        }
    }
    L->node = (double*)malloc(N * sizeof(double));
    L->im = (int*)malloc(cur * sizeof(int));
    L->ip = (int*)malloc(cur * sizeof(int));
    for (int i = 0; i < cur; ++i) L->im[i] = (i + cur - 1) % cur, L->ip[i] = (i + 1) % cur;
    for (int r = 0; r < cur; ++r)
      for (int cc = 0; cc < cur; ++cc) {
        const double* t_ = L->theta;
        const int rm = L->im[r], cm = L->im[cc];
        L->node[(size_t)r * cur + cc] = 0.25 * (t_[(size_t)r * cur + cm] + t_[(size_t)rm * cur + cm] +
                                                t_[(size_t)rm * cur + cc] + t_[(size_t)r * cur + cc]);
      }
    L->mass_analytic = (l == 0 && !theta);
    {
      const int nn = cur;
      const double* th = L->theta;
      const int n = nn;
      for (int r = 0; r < nn; ++r)
        for (int cc = 0; cc < nn; ++cc) {
          const size_t k = (size_t)r * nn + cc;
          if (L->mass_analytic) {
            L->mass_u[k] = thn(-(r + 0.5) * L->h, cc * L->h);     /* :325 */
            L->mass_v[k] = thn(-r * L->h, (cc + 0.5) * L->h);     /* :326 */
          } else {
            L->mass_u[k] = 0.5 * (TH(r, cc) + TH(r, cc - 1));
            L->mass_v[k] = 0.5 * (TH(r, cc) + TH(r - 1, cc));
          }
        }
    }
    L->tF = (double*)malloc(4 * N * sizeof(double)); L->rF = (double*)malloc(4 * N * sizeof(double));
    L->bF = (double*)malloc(4 * N * sizeof(double)); L->xF = (double*)malloc(4 * N * sizeof(double));
    L->tP = (double*)malloc(N * sizeof(double)); L->rP = (double*)malloc(N * sizeof(double));
    L->bP = (double*)malloc(N * sizeof(double)); L->xP = (double*)malloc(N * sizeof(double));
    l++;
    if (cur <= n_coarse || (cur % 2) || l >= OC_MAX_LEVELS) break;
    cur /= 2;
  }
  C->nlev = l;
  const size_t N0 = (size_t)n * n;
  double** v4[] = {&C->w, &C->g, &C->t2, &C->rinF, &C->zF, &C->dvF};
  for (int i = 0; i < 6; ++i) *v4[i] = (double*)malloc(4 * N0 * sizeof(double));
  double** v1[] = {&C->rhs, &C->xa, &C->xb, &C->rinP, &C->zP, &C->dvP};
  for (int i = 0; i < 6; ++i) *v1[i] = (double*)malloc(N0 * sizeof(double));
  /* coarsest-level dense inverses from the operator applied to unit vectors */
  oc_level* Lc = &C->lev[C->nlev - 1];
  const int nc = Lc->n;
  C->mF = 4 * nc * nc;
  C->mP = nc * nc;
  C->Finv = (double*)malloc((size_t)C->mF * C->mF * sizeof(double));
  C->Pinv = (double*)malloc((size_t)C->mP * C->mP * sizeof(double));
  {
    double* e = (double*)calloc(C->mF, sizeof(double));
    double* col = (double*)malloc(C->mF * sizeof(double));
    for (int j = 0; j < C->mF; ++j) {
      e[j] = 1.0;
      stokes_op(C, Lc, 0, 0, e, NULL, col, 0.0);
      e[j] = 0.0;
      for (int i = 0; i < C->mF; ++i) C->Finv[(size_t)i * C->mF + j] = col[i];
    }
    for (int j = 0; j < C->mP; ++j) {
      e[j] = 1.0;
      poisson_op(C, Lc, 0, e, NULL, col, 0.0);
      e[j] = 0.0;
      for (int i = 0; i < C->mP; ++i) C->Pinv[(size_t)i * C->mP + j] = col[i];
    }
    free(e);
    free(col);
    const double ee = 1.0 / C->mP; /* pseudo-inverse of the singular operator: (M + e e^T)^-1 - e e^T */
    for (size_t i = 0; i < (size_t)C->mP * C->mP; ++i) C->Pinv[i] += ee;
    if (invert_dense(C->Finv, C->mF) || invert_dense(C->Pinv, C->mP)) { oc_destroy(C); return NULL; }
    for (size_t i = 0; i < (size_t)C->mP * C->mP; ++i) C->Pinv[i] -= ee;
  }
  return C;
}

void oc_destroy(oc_ctx* C) {
  if (!C) return;
  for (int l = 0; l < C->nlev; ++l) {
    oc_level* L = &C->lev[l];
    free(L->theta); free(L->mass_u); free(L->mass_v); free(L->node); free(L->im); free(L->ip);
    free(L->tF); free(L->rF); free(L->bF); free(L->xF); free(L->tP); free(L->rP); free(L->bP); free(L->xP);
  }
  free(C->w); free(C->g); free(C->t2); free(C->rinF); free(C->zF); free(C->dvF);
  free(C->rhs); free(C->xa); free(C->xb); free(C->rinP); free(C->zP); free(C->dvP);
  free(C->Finv); free(C->Pinv);
  free(C);
}

int oc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline legs set the thread count explicitly */
void oc_set_num_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}

void oc_apply_A(oc_ctx* C, const double* x, double* y) { stokes_op(C, &C->lev[0], 0, 1, x, NULL, y, 0.0); }
void oc_apply_F(oc_ctx* C, const double* x, double* y) { stokes_op(C, &C->lev[0], 0, 0, x, NULL, y, 0.0); }
void oc_apply_G(oc_ctx* C, const double* p, double* y) { grad_op(C, &C->lev[0], p, y); }
void oc_apply_D(oc_ctx* C, const double* w, double* y) { div_op(&C->lev[0], w, NULL, y, 1.0); }
void oc_apply_GtG(oc_ctx* C, const double* p, double* y) { poisson_op(C, &C->lev[0], 0, p, NULL, y, 0.0); }
void oc_vcycle(oc_ctx* C, int isF, const double* b, double* x) { vcycle(C, 0, isF, b, x); }
void oc_solve(oc_ctx* C, int isF, const double* b, double* x) { sub_solve(C, isF, b, x); }

/* approx_schur_op, solve.py:257-277 */
void oc_precond(oc_ctx* C, const double* v, double* z) {
  const size_t N = (size_t)C->lev[0].n * C->lev[0].n;
  sub_solve(C, 1, v, C->w);                       /* :258 */
  div_op(&C->lev[0], C->w, v + 4 * N, C->rhs, 1.0); /* :259 */
  sub_solve(C, 0, C->rhs, C->xa);                 /* :265 */
  grad_op(C, &C->lev[0], C->xa, C->g);            /* :267  Gt_F_G x_a = -D F G x_a */
  stokes_op(C, &C->lev[0], 0, 0, C->g, NULL, C->t2, 0.0);
  div_op(&C->lev[0], C->t2, NULL, C->xb, -1.0);
  sub_solve(C, 0, C->xb, z + 4 * N);              /* :271 */
  grad_op(C, &C->lev[0], z + 4 * N, C->g);        /* :273 */
  sub_solve(C, 1, C->g, C->t2);                   /* :274 */
  axpby(1.0, C->w, -1.0, C->t2, z, 4 * N);        /* :275-276 */
}

/* right-preconditioned flexible GMRES (call shape of pyamg.krylov.fgmres at solve.py:285); returns iterations */
/* Conditioning probe for the residual-history tests: when amp > 0, the result of every A.x and M.v inside
 * oc_fgmres (and b itself) is multiplied elementwise by (1 + amp*u), u ~ U(-sqrt3, sqrt3) (unit variance), a
 * deterministic hash of (seed, call, index).  amp = 1.2e-16 models "another correct fp64 implementation". */
static double g_noise_amp = 0.0;
static unsigned long long g_noise_seed = 0, g_noise_call = 0;
void oc_set_noise(double amp, unsigned long long seed) { g_noise_amp = amp; g_noise_seed = seed; g_noise_call = 0; }
static void add_noise(double* y, size_t len) {
  if (g_noise_amp <= 0.0) return;
  const unsigned long long base = (g_noise_seed * 0x9E3779B97F4A7C15ull) ^ ((++g_noise_call) * 0xBF58476D1CE4E5B9ull);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < len; ++i) {
    unsigned long long z = base + (unsigned long long)i * 0x94D049BB133111EBull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double u = ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 3.4641016151377544;
    y[i] *= 1.0 + g_noise_amp * u;
  }
}

int oc_fgmres(oc_ctx* C, const double* b_in, double* x, double tol, int restart, int maxiter, int use_pc, double* hist,
              int* info) {
  const size_t len = 5 * (size_t)C->lev[0].n * C->lev[0].n;
  const int m = restart;
  const int mb = m < maxiter ? m : maxiter; /* bases actually touched (a bounded sample never needs all m) */
  double* V = (double*)malloc((size_t)(mb + 1) * len * sizeof(double));
  double* Z = (double*)malloc((size_t)mb * len * sizeof(double));
  double* w = (double*)malloc(len * sizeof(double));
  double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
  double* cs = (double*)calloc(m, sizeof(double));
  double* sn = (double*)calloc(m, sizeof(double));
  double* g = (double*)calloc(m + 1, sizeof(double));
  double* yv = (double*)calloc(m, sizeof(double));
  double* bcopy = NULL;
  const double* b = b_in;
  if (g_noise_amp > 0.0) {
    bcopy = (double*)malloc(len * sizeof(double));
    memcpy(bcopy, b_in, len * sizeof(double));
    add_noise(bcopy, len);
    b = bcopy;
  }
  double bn = sqrt(vdot(b, b, len));
  if (bn == 0.0) bn = 1.0;
  memset(x, 0, len * sizeof(double));
  int it = 0;
  *info = maxiter;
  while (it < maxiter) {
    oc_apply_A(C, x, w);
    axpby(1.0, b, -1.0, w, w, len);
    const double beta = sqrt(vdot(w, w, len));
    if (beta < tol * bn) { *info = 0; break; }
    axpby(1.0 / beta, w, 0.0, w, V, len);
    memset(g, 0, (m + 1) * sizeof(double));
    g[0] = beta;
    int jdone = 0;
    double res = beta;
    for (int j = 0; j < m; ++j) {
      double* zj = Z + (size_t)j * len;
      if (use_pc) { oc_precond(C, V + (size_t)j * len, zj); add_noise(zj, len); } else vcopy(V + (size_t)j * len, zj, len);
      oc_apply_A(C, zj, w);
      add_noise(w, len);
      for (int i = 0; i <= j; ++i) {
        const double hij = vdot(V + (size_t)i * len, w, len);
        H[(size_t)i * m + j] = hij;
        axpby(1.0, w, -hij, V + (size_t)i * len, w, len);
      }
      const double hn = sqrt(vdot(w, w, len));
      H[(size_t)(j + 1) * m + j] = hn;
      if (hn != 0.0) axpby(1.0 / hn, w, 0.0, w, V + (size_t)(j + 1) * len, len);
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = t;
      }
      const double den = hypot(H[(size_t)j * m + j], H[(size_t)(j + 1) * m + j]);
      cs[j] = H[(size_t)j * m + j] / den;
      sn[j] = H[(size_t)(j + 1) * m + j] / den;
      H[(size_t)j * m + j] = den;
      H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      it++;
      jdone = j + 1;
      res = fabs(g[j + 1]);
      if (hist) hist[it - 1] = res / bn;
      if (res < tol * bn || it >= maxiter) break;
    }
    for (int i = jdone - 1; i >= 0; --i) {
      double s = g[i];
      for (int k = i + 1; k < jdone; ++k) s -= H[(size_t)i * m + k] * yv[k];
      yv[i] = s / H[(size_t)i * m + i];
    }
    for (int k = 0; k < jdone; ++k) axpby(1.0, x, yv[k], Z + (size_t)k * len, x, len);
    if (res < tol * bn) { *info = 0; break; }
  }
  free(V); free(Z); free(w); free(H); free(cs); free(sn); free(g); free(yv); free(bcopy);
  return it;
}
