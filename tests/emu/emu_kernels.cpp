// Host build of the CUDA stencil kernels on the SIMT-on-CPU shim (tests/emu/cuda_emu.h): logic checks of
// the kernels against the oracle on the GPU-less build container.  TEST INFRASTRUCTURE ONLY.
#define MPBP_EMU 1
#include <chrono>
#include "../../mp-block-preconditioners_b200/csrc/stencil.cuh"
#include "../../mp-block-preconditioners_b200/csrc/coarse.cuh"
#include "../../mp-block-preconditioners_b200/csrc/stokes.cuh"
#include "../../mp-block-preconditioners_b200/csrc/cell.cuh"
#include "../../mp-block-preconditioners_b200/csrc/poisson.cuh"

using namespace mpbp;

namespace {
struct Tables {
  std::vector<double> sxf, sxc, syf, syc;
};
Phys make_phys(int n, double xi, double eta_n, double eta_s, double c, double d_u, double d_p, double d_div,
               int mass_mode, Tables& t) {
  const double PI = 3.141592653589793, h = 1.0 / n;
  t.sxf.resize(n); t.sxc.resize(n); t.syf.resize(n); t.syc.resize(n);
  for (int i = 0; i < n; ++i) {
    t.sxf[i] = std::sin(2 * PI * (i * h));
    t.sxc[i] = std::sin(2 * PI * ((i + 0.5) * h));
    t.syf[i] = std::sin(2 * PI * (-i * h));
    t.syc[i] = std::sin(2 * PI * (-(i + 0.5) * h));
  }
  Phys ph{};
  ph.xi = xi; ph.c = c; ph.d_u = d_u;
  ph.kap_n = d_u * eta_n / (h * h);
  ph.kap_s = d_u * eta_s / (h * h);
  ph.dp_h = d_p / h; ph.ddiv_h = d_div / h; ph.inv_h = 1.0 / h; ph.dp_h2 = d_p / (h * h);
  ph.mass_mode = mass_mode;
  ph.sxf = t.sxf.data(); ph.sxc = t.sxc.data(); ph.syf = t.syf.data(); ph.syc = t.syc.data();
  return ph;
}
VecIn whole_grid_view(const double* x, int n) {
  VecIn v{};
  v.x = x;
  v.fs = v.hs = (size_t)n * n;
  v.top = x + (size_t)(n - 1) * n;
  v.bot = x;
  return v;
}
dim3 sgrid(int n, int rs) { return dim3((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (n + rs - 1) / rs); }
}  // namespace

extern "C" {

// params: xi, eta_n, eta_s, c, d_u, d_p, d_div ; th_pad: (n+2) x n padded theta
void emu_stokes(int mode, int with_p, int n, const double* prm, int mass_mode, const double* th_pad, const double* x,
                const double* b, double* y, int rs, int pf, double omega) {
  Tables t;
  const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], mass_mode, t);
  const Geo g{n, n, 0, rs, pf};
  const VecIn in = whole_grid_view(x, n);
  StokesArgs a{};
  a.xin = in; a.th = th_pad; a.b = b; a.y = y; a.g = g; a.ph = ph; a.omega = omega;
  emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] {
    if (with_p) k_stokes_x<0, 0, true, 0, false, 0>(a);
    else if (mode == 0) k_stokes_x<0, 0, false, 0, false, 0>(a);
    else if (mode == 1) k_stokes_x<0, 1, false, 0, false, 0>(a);
    else k_stokes_x<0, 2, false, 0, false, 0>(a);
  });
}

void emu_stokes_fused(int variant, int n, const double* prm, int mass_mode, const double* th_pad, const double* x,
                      const double* b, const double* wd, const double* ec, double* y, int rs, int pf, double omega) {
  Tables t;
  const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], mass_mode, t);
  const Geo g{n, n, 0, rs, pf};
  const VecIn in = whole_grid_view(x, n);
  StokesArgs a{};
  a.xin = in; a.th = th_pad; a.b = b; a.y = y; a.g = g; a.ph = ph; a.omega = omega;
  if (wd) a.wd = whole_grid_view(wd, n);
  if (ec) a.cin = whole_grid_view(ec, n / 2);
  a.nc = n / 2;
  a.rows_c = n / 2;
  emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] {
    if (variant == 0) k_stokes_x<1, 2, false, 0, false, 0>(a);
    else k_stokes_x<2, 2, false, 0, false, 0>(a);
  });
}

// the unified marching kernel of csrc/stokes.cuh on a whole (periodic) grid.  out: y (EP 0, 4 or 5 fields), nothing
// (EP 1: d / xk are updated in place), or the coarse rhs (EP 2)
void emu_stokes_x(int in, int mode, int with_p, int ep, int n, const double* prm, int mass_mode, const double* th_pad,
                  const double* x, const double* b, const double* wd, const double* ec, double* d, double* xk,
                  const double* cheb, const int* flags, double* out, int rs, int pf, double omega, int re) {
  Tables t;
  StokesArgs a{};
  a.ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], mass_mode, t);
  a.g = Geo{n, n, 0, rs, pf, re};
  a.xin = whole_grid_view(x, n);
  a.th = th_pad;
  a.b = b;
  a.y = out;
  a.omega = omega;
  if (wd) a.wd = whole_grid_view(wd, n);
  if (ec) a.cin = whole_grid_view(ec, n / 2);
  a.nc = n / 2;
  a.rows_c = n / 2;
  if (cheb) a.ce = ChebEp{cheb[0], cheb[1], d, xk, flags[0], flags[1], flags[2]};
  a.bc = out;
  const int wc = (ep == 2) ? WarpTile<2>::cols : WarpTile<0>::cols;
  const dim3 grid((n + wc * kBlockWarps - 1) / (wc * kBlockWarps), strip_count(a.g));
  const int key = in * 1000 + mode * 100 + with_p * 10 + ep;
  emu::launch(grid, dim3(kBlockThreads), [&] {
    switch (key) {
      case 10: k_stokes_x<0, 0, true, 0, false, 0>(a); break;
      case 0: k_stokes_x<0, 0, false, 0, false, 0>(a); break;
      case 100: k_stokes_x<0, 1, false, 0, false, 0>(a); break;
      case 200: k_stokes_x<0, 2, false, 0, false, 0>(a); break;
      case 1200: k_stokes_x<1, 2, false, 0, false, 0>(a); break;
      case 2200: k_stokes_x<2, 2, false, 0, false, 0>(a); break;
      case 201: k_stokes_x<0, 2, false, 1, false, 0>(a); break;
      case 2201: k_stokes_x<2, 2, false, 1, false, 0>(a); break;
      case 102: k_stokes_x<0, 1, false, 2, false, 0>(a); break;
      default: break;
    }
  });
}

void emu_jacobi0_F(int n, const double* prm, int mass_mode, const double* th_pad, const double* b, double* y, int rs,
                   double omega) {
  Tables t;
  const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], mass_mode, t);
  const Geo g{n, n, 0, rs, 0};
  emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] { k_jacobi0_F(th_pad, b, y, (size_t)n * n, g, ph, omega); });
}

void emu_poisson(int mode, int n, const double* prm, const double* th_pad, const double* p, const double* b, double* y,
                 int rs, double omega) {
  Tables t;
  const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 0, t);
  const Geo g{n, n, 0, rs, 0};
  const VecIn in = whole_grid_view(p ? p : b, n);
  emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] {
    switch (mode) {
      case 0: k_poisson<0>(in, th_pad, b, y, g, ph, omega); break;
      case 1: k_poisson<1>(in, th_pad, b, y, g, ph, omega); break;
      case 2: k_poisson<2>(in, th_pad, b, y, g, ph, omega); break;
      default: k_poisson<3>(in, th_pad, b, y, g, ph, omega); break;
    }
  });
}

void emu_div_grad(int what, int n, const double* prm, const double* th_pad, const double* in_, const double* add,
                  double* out, int rs, double scale) {
  // what 0: out = scale * D w + add (k_div) ; 1: out = G p (k_grad)
  Tables t;
  Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 0, t);
  const Geo g{n, n, 0, rs, 0};
  const VecIn in = whole_grid_view(in_, n);
  if (what == 0) {
    ph.inv_h *= scale;
    emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] { k_div(in, th_pad, add, out, g, ph); });
  } else {
    emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] { k_grad<false>(in, th_pad, out, (size_t)n * n, g, ph); });
  }
}

void emu_transfer(int what, int nf, const double* in_, double* out) {
  // what: 0 restrict_F (in: fine 4 fields -> out: coarse), 1 prolong_add_F (in: coarse, out: fine, accumulates),
  //       2 restrict_P, 3 prolong_add_P
  const int nc = nf / 2;
  if (what == 0) {
    const VecIn v = whole_grid_view(in_, nf);
    emu::launch(dim3((nc + 127) / 128, nc), dim3(128), [&] { k_restrict_F<false>(v, out, nf, nf); });
  } else if (what == 1) {
    const VecIn v = whole_grid_view(in_, nc);
    emu::launch(dim3((nf + 127) / 128, nf), dim3(128), [&] { k_prolong_add_F(v, out, nf, nf); });
  } else if (what == 2) {
    emu::launch(dim3((nc + 127) / 128, nc), dim3(128), [&] { k_restrict_P(in_, out, nf, nf); });
  } else {
    emu::launch(dim3((nf + 127) / 128, nf), dim3(128), [&] { k_prolong_add_P<false>(in_, out, nf, nf); });
  }
}

// persistent coarse V-cycle (csrc/coarse.cuh) on the hierarchy n -> n_coarse built here the way plan.cu builds it
// (theta averaged 2x2, padded; mass term analytic on the first level only when mass_mode is set)
void emu_coarse_vcycle(int isF, int n, const double* prm, int mass_mode, const double* theta, int n_coarse,
                       double omega, int nu1, int nu2, const double* Minv_t, int m, const double* b, double* x,
                       int threads) {
  std::vector<Tables> tabs(kCoarseMaxLevels);
  std::vector<std::vector<double>> th_pad, work;
  CoarseArgs a{};
  std::vector<double> th(theta, theta + (size_t)n * n);
  int cur = n, l = 0;
  const int nf = isF ? 4 : 1;
  while (true) {
    if (l > 0) {
      const int f = 2 * cur;
      std::vector<double> tc((size_t)cur * cur);
      for (int R = 0; R < cur; ++R)
        for (int C = 0; C < cur; ++C) {
          const double* q = &th[(size_t)(2 * R) * f + 2 * C];
          tc[(size_t)R * cur + C] = 0.25 * (q[0] + q[f] + q[1] + q[f + 1]);
        }
      th.swap(tc);
    }
    th_pad.emplace_back((size_t)(cur + 2) * cur);
    for (int r = -1; r <= cur; ++r)
      std::memcpy(&th_pad.back()[(size_t)(r + 1) * cur], &th[(size_t)((r + cur) % cur) * cur], cur * sizeof(double));
    CoarseLevel& L = a.lev[l];
    L.n = cur;
    L.ph = make_phys(cur, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], (l == 0) ? mass_mode : 0, tabs[l]);
    for (int k = 0; k < 4; ++k) work.emplace_back((size_t)nf * cur * cur, 0.0);
    l++;
    if (cur <= n_coarse || (cur % 2) || l >= kCoarseMaxLevels) break;
    cur /= 2;
  }
  a.nlev = l;
  for (int i = 0; i < l; ++i) {
    a.lev[i].th = th_pad[i].data();
    a.lev[i].b = work[4 * i].data();
    a.lev[i].x = work[4 * i + 1].data();
    a.lev[i].t = work[4 * i + 2].data();
    a.lev[i].r = work[4 * i + 3].data();
  }
  a.Minv_t = Minv_t;
  a.m = m;
  a.b_in = b;
  a.x_out = x;
  a.omega = omega;
  a.nu1 = nu1;
  a.nu2 = nu2;
  emu::launch(dim3(1), dim3(threads), [&] {
    if (isF) k_coarse_vcycle<true>(a);
    else k_coarse_vcycle<false>(a);
  });
}

// P ranks of the slab decomposition emulated in ONE process: every "rank" owns a slab of x, a comm buffer with
// the layout of plan.cu (flags + [slot][dir][5][n] halo areas) and a device-resident exchange counter.  First all
// ranks run k_halo_push into their ring neighbours' buffers (so every flag is already set and nothing has to
// spin), then every rank runs the consumer stencil kernel on its slab with the peer-pushed halo rows.  `rounds`
// repeats push+apply to exercise the alternating slots.  y receives the assembled global result of A.x.
void emu_slab_apply_A(int P, int rounds, int n, const double* prm, const double* theta, const double* x, double* y,
                      int rs) {
  const int rows = n / P;
  const size_t fs = (size_t)rows * n, area = (size_t)5 * n;
  const size_t comm_bytes = comm_halo_bytes(area);
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_bytes, 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<unsigned int> counter(P, 0u);
  std::vector<std::vector<double>> xs(P, std::vector<double>(5 * fs)), ys(P, std::vector<double>(5 * fs)), thp(P);
  std::vector<std::vector<double>> land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  for (int g = 0; g < P; ++g) {
    for (int k = 0; k < 5; ++k)
      std::memcpy(&xs[g][k * fs], x + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
  }
  for (int it = 0; it < rounds; ++it) {
    for (int g = 0; g < P; ++g) {
      const int prev = (g + P - 1) % P, next = (g + 1) % P;
      emu::launch(dim3((5 * n + 255) / 256), dim3(256), [&] {
        k_halo_push(xs[g].data(), 5, fs, rows, n, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                    &counter[g]);
      });
    }
    for (int g = 0; g < P; ++g) {
      const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 1, tabs[g]);
      const Geo geo{n, rows, g * rows, rs, 3};
      VecIn in{};
      in.x = xs[g].data();
      in.fs = fs;
      in.hs = n;
      in.dseq = &dseq[g];
      in.comm = comm[g].data();
      in.area = area;
      in.land = land[g].data(); in.top = in.land; in.bot = in.land + 5 * n;
      StokesArgs a{};
      a.xin = in; a.th = thp[g].data(); a.y = ys[g].data(); a.g = geo; a.ph = ph;
      emu::launch(dim3((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (rows + rs - 1) / rs),
                  dim3(kBlockThreads), [&] { k_stokes_x<0, 0, true, 0, false, 0>(a); });
    }
  }
  for (int g = 0; g < P; ++g)
    for (int k = 0; k < 5; ++k)
      std::memcpy(y + (size_t)k * n * n + (size_t)g * rows * n, &ys[g][k * fs], fs * sizeof(double));
}

// Fused halo push chain on P emulated ranks: x0 --k_halo_push--> sweep (fused push) -> sweep (fused push) ->
// residual.  Every kernel is run for all ranks before the next one starts, so all flags are set when read.
// out receives the assembled global residual b - F x2 with x_{k+1} = x_k + omega (b - F x_k)/diag.
void emu_slab_fused_push_chain(int P, int n, const double* prm, const double* theta, const double* x0, const double* b,
                               double* out, int rs, double omega) {
  const int rows = n / P;
  const size_t fs = (size_t)rows * n, area = (size_t)5 * n;
  const size_t comm_bytes = comm_halo_bytes(area);
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_bytes, 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<std::vector<unsigned int>> counter(P, std::vector<unsigned int>(8, 0u));
  std::vector<std::vector<double>> xa(P, std::vector<double>(4 * fs)), xb(P, std::vector<double>(4 * fs)),
      bs(P, std::vector<double>(4 * fs)), thp(P);
  std::vector<std::vector<double>> land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  for (int g = 0; g < P; ++g) {
    for (int k = 0; k < 4; ++k) {
      std::memcpy(&xa[g][k * fs], x0 + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
      std::memcpy(&bs[g][k * fs], b + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
    }
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
  }
  const dim3 grid((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (rows + rs - 1) / rs);
  auto view = [&](int g, const double* x) {
    VecIn in{};
    in.x = x;
    in.fs = fs;
    in.hs = n;
    in.dseq = &dseq[g];
    in.comm = comm[g].data();
    in.area = area;
    in.land = land[g].data(); in.top = in.land; in.bot = in.land + 5 * n;
    return in;
  };
  // exchange 1: classic push kernel of x0
  for (int g = 0; g < P; ++g) {
    const int prev = (g + P - 1) % P, next = (g + 1) % P;
    emu::launch(dim3((4 * n + 255) / 256), dim3(256), [&] {
      k_halo_push(xa[g].data(), 4, fs, rows, n, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                  &counter[g][0]);
    });
  }
  // two sweeps with fused pushes (ping-pong xa -> xb -> xa), then the residual of xa into xb
  for (int step = 0; step < 3; ++step) {
    for (int g = 0; g < P; ++g) {
      const int prev = (g + P - 1) % P, next = (g + 1) % P;
      const Phys ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 1, tabs[g]);
      const Geo geo{n, rows, g * rows, rs, 3};
      const double* src = (step == 1) ? xb[g].data() : xa[g].data();
      double* dst = (step == 1) ? xa[g].data() : xb[g].data();
      StokesArgs a{};
      a.xin = view(g, src);
      a.th = thp[g].data(); a.b = bs[g].data(); a.y = dst; a.g = geo; a.ph = ph; a.omega = omega;
      a.po = PushOut{comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g], &counter[g][4]};
      emu::launch(grid, dim3(kBlockThreads), [&] {
        if (step < 2) k_stokes_x<0, 2, false, 0, true, 0>(a);
        else k_stokes_x<0, 1, false, 0, false, 0>(a);
      });
    }
  }
  for (int g = 0; g < P; ++g)
    for (int k = 0; k < 4; ++k)
      std::memcpy(out + (size_t)k * n * n + (size_t)g * rows * n, &xb[g][k * fs], fs * sizeof(double));
}

// The fused V-cycle kernels on DISTRIBUTED levels, P ranks emulated in one process (every kernel runs on all ranks
// before the next starts, so all flags are set when read):
//   push(b) -> pre-smoothing pair [IN 1: live halo of b, static halo of wd, fused push of x2]
//           -> residual [consumes the pushed rows, STASHES them, fused push of r]
//           -> push(e_c on the coarse slabs)
//           -> prolongation + sweep [IN 2: x2 with its stashed halo rows, live halo of e_c, fused push of x3]
//           -> sweep [consumes the pushed rows of x3]
// out_x receives the assembled x4, out_r the assembled residual b - F x2.
void emu_slab_fused_vcycle_chain(int P, int n, const double* prm, const double* theta, const double* b,
                                 const double* wd, const double* ec, double* out_x, double* out_r, int rs, double omega) {
  const int rows = n / P, nc = n / 2, rows_c = rows / 2;
  const size_t fs = (size_t)rows * n, fsc = (size_t)rows_c * nc, area = (size_t)5 * n;
  const size_t comm_bytes = comm_halo_bytes(area);
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_bytes, 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<std::vector<unsigned int>> counter(P, std::vector<unsigned int>(8, 0u));
  auto slab = [&](const double* g, int g_n, int g_rows, int rank) {
    const size_t f = (size_t)g_rows * g_n;
    std::vector<double> v(4 * f);
    for (int k = 0; k < 4; ++k)
      std::memcpy(&v[k * f], g + (size_t)k * g_n * g_n + (size_t)rank * g_rows * g_n, f * sizeof(double));
    return v;
  };
  std::vector<std::vector<double>> bs(P), ws(P), es(P), x2(P), rr(P), x3(P), x4(P), thp(P), wdh(P), stash(P);
  std::vector<std::vector<double>> land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  for (int g = 0; g < P; ++g) {
    bs[g] = slab(b, n, rows, g);
    ws[g] = slab(wd, n, rows, g);
    es[g] = slab(ec, nc, rows_c, g);
    x2[g].assign(4 * fs, 0.0); rr[g].assign(4 * fs, 0.0); x3[g].assign(4 * fs, 0.0); x4[g].assign(4 * fs, 0.0);
    stash[g].assign(2 * 5 * n, 0.0);
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
    wdh[g].resize(2 * 4 * n);  // the neighbours' boundary rows of wd (fetched once at plan creation in plan.cu)
    for (int k = 0; k < 4; ++k) {
      const int rt = ((g * rows - 1) % n + n) % n, rb = ((g + 1) * rows) % n;
      std::memcpy(&wdh[g][k * n], wd + (size_t)k * n * n + (size_t)rt * n, n * sizeof(double));
      std::memcpy(&wdh[g][(4 + k) * n], wd + (size_t)k * n * n + (size_t)rb * n, n * sizeof(double));
    }
  }
  const dim3 grid((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (rows + rs - 1) / rs);
  auto live = [&](int g, const double* x, size_t f, int hs) {
    VecIn in{};
    in.x = x; in.fs = f; in.hs = hs; in.dseq = &dseq[g]; in.comm = comm[g].data(); in.area = area;
    in.land = land[g].data(); in.top = in.land; in.bot = in.land + 5 * hs;  // landing buffer = the view's halo rows
    return in;
  };
  auto base = [&](int g) {
    StokesArgs a{};
    a.th = thp[g].data();
    a.ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 1, tabs[g]);
    a.g = Geo{n, rows, g * rows, rs, 3};
    a.omega = omega;
    const int prev = (g + P - 1) % P, next = (g + 1) % P;
    a.po = PushOut{comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g], &counter[g][4]};
    return a;
  };
  auto push = [&](std::vector<std::vector<double>>& xs, size_t f, int r_, int n_) {
    for (int g = 0; g < P; ++g) {
      const int prev = (g + P - 1) % P, next = (g + 1) % P;
      emu::launch(dim3((4 * n_ + 255) / 256), dim3(256), [&] {
        k_halo_push(xs[g].data(), 4, f, r_, n_, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                    &counter[g][0]);
      });
    }
  };
  push(bs, fs, rows, n);
  for (int g = 0; g < P; ++g) {  // pre-smoothing pair
    StokesArgs a = base(g);
    a.xin = live(g, bs[g].data(), fs, n);
    a.wd.x = ws[g].data(); a.wd.fs = fs; a.wd.hs = n; a.wd.top = wdh[g].data(); a.wd.bot = wdh[g].data() + 4 * n;
    a.y = x2[g].data();
    emu::launch(grid, dim3(kBlockThreads), [&] { k_stokes_x<1, 2, false, 0, true, 0>(a); });
  }
  for (int g = 0; g < P; ++g) {  // residual, stashing the halo rows of x2
    StokesArgs a = base(g);
    a.xin = live(g, x2[g].data(), fs, n);
    a.b = bs[g].data();
    a.y = rr[g].data();
    a.xin.land = stash[g].data();  // land x2's halo rows in the static stash instead
    a.xin.top = a.xin.land; a.xin.bot = a.xin.land + 5 * n;
    emu::launch(grid, dim3(kBlockThreads), [&] { k_stokes_x<0, 1, false, 0, true, 0>(a); });
  }
  push(es, fsc, rows_c, nc);
  for (int g = 0; g < P; ++g) {  // prolongation + first post-sweep
    StokesArgs a = base(g);
    a.xin.x = x2[g].data(); a.xin.fs = fs; a.xin.hs = n; a.xin.top = stash[g].data(); a.xin.bot = stash[g].data() + 5 * n;
    a.cin = live(g, es[g].data(), fsc, nc);
    a.nc = nc; a.rows_c = rows_c;
    a.b = bs[g].data();
    a.y = x3[g].data();
    emu::launch(grid, dim3(kBlockThreads), [&] { k_stokes_x<2, 2, false, 0, true, 0>(a); });
  }
  for (int g = 0; g < P; ++g) {  // second post-sweep
    StokesArgs a = base(g);
    a.xin = live(g, x3[g].data(), fs, n);
    a.b = bs[g].data();
    a.y = x4[g].data();
    emu::launch(grid, dim3(kBlockThreads), [&] { k_stokes_x<0, 2, false, 0, false, 0>(a); });
  }
  for (int g = 0; g < P; ++g)
    for (int k = 0; k < 4; ++k) {
      std::memcpy(out_x + (size_t)k * n * n + (size_t)g * rows * n, &x4[g][k * fs], fs * sizeof(double));
      std::memcpy(out_r + (size_t)k * n * n + (size_t)g * rows * n, &rr[g][k * fs], fs * sizeof(double));
    }
}

// Fused residual + restriction on the slabs of a DISTRIBUTED level (EP 2 with PUSH), P ranks emulated: x's halo rows
// are pushed first (all ranks), then every rank runs the kernel CONCURRENTLY (one host thread per rank): a slab's first
// strip waits inside the kernel for the previous rank's last-row half-sums.  out_bc: assembled coarse rhs R (b - F x);
// out_rows [P][2][4][nc]: the coarse rows each rank RECEIVED in its comm buffer (row -1 from the previous rank, row
// rows_c from the next), i.e. what the next level's pre-smoother will read as halos.
void emu_slab_residual_restrict(int P, int n, const double* prm, const double* theta, const double* x, const double* b,
                                double* out_bc, double* out_rows, int rs) {
  const int rows = n / P, nc = n / 2, rows_c = rows / 2;
  const size_t fs = (size_t)rows * n, fsc = (size_t)rows_c * nc, area = (size_t)5 * n;
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_halo_bytes(area), 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<std::vector<unsigned int>> counter(P, std::vector<unsigned int>(8, 0u));
  std::vector<std::vector<double>> xs(P), bs(P), bc(P, std::vector<double>(4 * fsc, 0.0)), thp(P);
  std::vector<std::vector<double>> land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  for (int g = 0; g < P; ++g) {
    xs[g].resize(4 * fs);
    bs[g].resize(4 * fs);
    for (int k = 0; k < 4; ++k) {
      std::memcpy(&xs[g][k * fs], x + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
      std::memcpy(&bs[g][k * fs], b + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
    }
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
  }
  for (int g = 0; g < P; ++g) {
    const int prev = (g + P - 1) % P, next = (g + 1) % P;
    emu::launch(dim3((4 * n + 255) / 256), dim3(256), [&] {
      k_halo_push(xs[g].data(), 4, fs, rows, n, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                  &counter[g][0]);
    });
  }
  std::vector<std::thread> ranks;
  std::vector<StokesArgs> args(P);
  for (int g = 0; g < P; ++g) {
    const int prev = (g + P - 1) % P, next = (g + 1) % P;
    StokesArgs& a = args[g];
    a = StokesArgs{};
    a.th = thp[g].data();
    a.ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 1, tabs[g]);
    a.g = Geo{n, rows, g * rows, rs, 3};
    a.xin.x = xs[g].data(); a.xin.fs = fs; a.xin.hs = n; a.xin.dseq = &dseq[g]; a.xin.comm = comm[g].data();
    a.xin.area = area; a.xin.land = land[g].data(); a.xin.top = a.xin.land; a.xin.bot = a.xin.land + 5 * n;
    a.b = bs[g].data();
    a.bc = bc[g].data();
    a.nc = nc;
    a.rows_c = rows_c;
    a.po = PushOut{comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g], &counter[g][4]};
  }
  const dim3 grid((n + WarpTile<2>::cols * kBlockWarps - 1) / (WarpTile<2>::cols * kBlockWarps), (rows + rs - 1) / rs);
  for (int g = 0; g < P; ++g)
    ranks.emplace_back([&, g] { emu::launch_concurrent(grid, dim3(kBlockThreads), [&, g] { k_stokes_x<0, 1, false, 2, true, 0>(args[g]); }); });
  for (auto& t : ranks) t.join();
  for (int g = 0; g < P; ++g) {
    for (int k = 0; k < 4; ++k)
      std::memcpy(out_bc + (size_t)k * nc * nc + (size_t)g * rows_c * nc, &bc[g][k * fsc], fsc * sizeof(double));
    const int slot = (int)(dseq[g] % (unsigned long long)kHaloSlots);
    for (int dir = 0; dir < 2; ++dir) {
      const LLElem* e = comm_halo(comm[g].data(), area, slot, dir);
      for (int i = 0; i < 4 * nc; ++i)
        out_rows[((size_t)g * 2 + dir) * 4 * nc + i] = (e[i].tag == dseq[g]) ? e[i].v : std::nan("");
    }
  }
}

// csrc/cell.cuh: the cell-parallel kernels of the small whole-grid levels.  in: 0 sweep, 1 pre-smoothing pair,
// 2 prolongation + sweep, 3 residual + restriction (out: 4 * (n/2)^2 values).
void emu_cell(int in, int n, const double* prm, const double* th_pad, const double* x, const double* b, const double* wd,
              const double* ec, double* out, double omega) {
  Tables t;
  CellArgs a{};
  a.th = th_pad;
  a.ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 0, t);
  a.n = n;
  a.x = x; a.b = b; a.wd = wd; a.ec = ec; a.y = out; a.bc = out; a.omega = omega;
  const dim3 block(kCellBX * kCellBY);
  auto grid = [](int nx, int ny) { return dim3((nx + kCellBX - 1) / kCellBX, (ny + kCellBY - 1) / kCellBY); };
  if (in == 3) emu::launch(grid(n / 2, n / 2), block, [&] { k_cell_rr(a); });
  else if (in == 2) emu::launch(grid(n, n), block, [&] { k_cell_sweep<2>(a); });
  else if (in == 1) emu::launch(grid(n, n), block, [&] { k_cell_sweep<1>(a); });
  else emu::launch(grid(n, n), block, [&] { k_cell_sweep<0>(a); });
}

// csrc/poisson.cuh: fused pressure-Poisson kernels on a whole-grid level.  in: 1 pre-smoothing pair, 3 residual +
// restriction (out: (n/2)^2 values), 2 prolongation + sweep, 4 prolongation + last sweep with the Chebyshev epilogue
// (d, xk updated in place; cheb = [ca, cb], flags = [read_d, read_x, write_d]).
void emu_poisson_f(int in, int n, const double* prm, const double* th_pad, const double* x, const double* b,
                   const double* wd, const double* ec, double* d, double* xk, const double* cheb, const int* flags,
                   double* out, int rs, double omega) {
  Tables t;
  PoissonFArgs a{};
  a.x = x; a.b = b; a.wd = wd; a.ec = ec; a.th = th_pad; a.y = out; a.bc = out;
  a.g = Geo{n, n, 0, rs, 0};
  a.ph = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 0, t);
  a.omega = omega;
  if (in == 4) a.ce = ChebEp{cheb[0], cheb[1], d, xk, flags[0], flags[1], flags[2]};
  emu::launch(sgrid(n, rs), dim3(kBlockThreads), [&] {
    switch (in) {
      case 1: k_poisson_f<1, 2, 0>(a); break;
      case 3: k_poisson_f<0, 1, 2>(a); break;
      case 2: k_poisson_f<2, 2, 0>(a); break;
      default: k_poisson_f<2, 2, 1>(a); break;
    }
  });
}

// Stress test of the LL halo protocol under SKEW (slot reuse with >= 3 ranks): every emulated rank runs its
// whole kernel sequence -- k_halo_push(x0), `sweeps` Jacobi sweeps with fused pushes, a final residual -- in its own
// host thread, free-running, with pseudo-random delays of up to a few milliseconds between its kernels.  Nothing but
// the protocol (sequence tags, three slots, producer credit) keeps a fast rank from overwriting rows a slow neighbour
// has not read yet.  out receives the assembled residual b - F x_sweeps.
void emu_slab_push_chain_skewed(int P, int n, const double* prm, const double* theta, const double* x0, const double* b,
                                double* out, int rs, double omega, int sweeps, unsigned seed) {
  const int rows = n / P;
  const size_t fs = (size_t)rows * n, area = (size_t)5 * n;
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_halo_bytes(area), 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<std::vector<unsigned int>> counter(P, std::vector<unsigned int>(8, 0u));
  std::vector<std::vector<double>> xa(P, std::vector<double>(4 * fs)), xb(P, std::vector<double>(4 * fs)),
      bs(P, std::vector<double>(4 * fs)), thp(P), land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  std::vector<Phys> phys(P);
  for (int g = 0; g < P; ++g) {
    for (int k = 0; k < 4; ++k) {
      std::memcpy(&xa[g][k * fs], x0 + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
      std::memcpy(&bs[g][k * fs], b + (size_t)k * n * n + (size_t)g * rows * n, fs * sizeof(double));
    }
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
    phys[g] = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 1, tabs[g]);
  }
  const dim3 grid((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (rows + rs - 1) / rs);
  std::vector<std::thread> ranks;
  for (int g = 0; g < P; ++g)
    ranks.emplace_back([&, g] {
      unsigned state = seed * 2654435761u + 97u * (unsigned)g + 1u;
      auto nap = [&] {
        state = state * 1664525u + 1013904223u;
        std::this_thread::sleep_for(std::chrono::microseconds((state >> 16) % 3000));
      };
      const int prev = (g + P - 1) % P, next = (g + 1) % P;
      nap();
      emu::launch(dim3((4 * n + 255) / 256), dim3(256), [&] {
        k_halo_push(xa[g].data(), 4, fs, rows, n, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                    &counter[g][0]);
      });
      double* src = xa[g].data();
      double* dst = xb[g].data();
      for (int step = 0; step <= sweeps; ++step) {
        nap();
        StokesArgs a{};
        a.xin.x = src; a.xin.fs = fs; a.xin.hs = n; a.xin.dseq = &dseq[g]; a.xin.comm = comm[g].data(); a.xin.area = area;
        a.xin.land = land[g].data(); a.xin.top = a.xin.land; a.xin.bot = a.xin.land + 5 * n;
        a.th = thp[g].data(); a.b = bs[g].data(); a.y = dst; a.g = Geo{n, rows, g * rows, rs, 3}; a.ph = phys[g];
        a.omega = omega;
        a.po = PushOut{comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g], &counter[g][4]};
        const bool last = step == sweeps;
        emu::launch(grid, dim3(kBlockThreads), [&] {
          if (!last) k_stokes_x<0, 2, false, 0, true, 0>(a);
          else k_stokes_x<0, 1, false, 0, false, 0>(a);
        });
        std::swap(src, dst);
      }
      // after the swap `src` holds the residual
      for (int k = 0; k < 4; ++k)
        std::memcpy(out + (size_t)k * n * n + (size_t)g * rows * n, src + k * fs, fs * sizeof(double));
    });
  for (auto& t : ranks) t.join();
}

// The pressure-Poisson kernels on the slabs of a distributed level, P free-running emulated ranks (one host thread each,
// random delays): k_halo_push(p0), `sweeps` GtG Jacobi sweeps with fused pushes (k_poisson<2, false, true>: edge strips
// fetch the neighbours' rows -- the FETCH instantiation -- and push their own), then the residual k_poisson<1>.
void emu_slab_poisson_chain(int P, int n, const double* prm, const double* theta, const double* p0, const double* b,
                            double* out, int rs, double omega, int sweeps, unsigned seed) {
  const int rows = n / P;
  const size_t fs = (size_t)rows * n, area = (size_t)5 * n;
  std::vector<std::vector<char>> comm(P, std::vector<char>(comm_halo_bytes(area), 0));
  std::vector<unsigned long long> dseq(P, 0ull);
  std::vector<std::vector<unsigned int>> counter(P, std::vector<unsigned int>(8, 0u));
  std::vector<std::vector<double>> xa(P, std::vector<double>(fs)), xb(P, std::vector<double>(fs)), bs(P, std::vector<double>(fs)),
      thp(P), land(P, std::vector<double>(2 * 5 * n, 0.0));
  std::vector<Tables> tabs(P);
  std::vector<Phys> phys(P);
  for (int g = 0; g < P; ++g) {
    std::memcpy(xa[g].data(), p0 + (size_t)g * rows * n, fs * sizeof(double));
    std::memcpy(bs[g].data(), b + (size_t)g * rows * n, fs * sizeof(double));
    thp[g].resize((size_t)(rows + 2) * n);
    for (int r = -1; r <= rows; ++r)
      std::memcpy(&thp[g][(size_t)(r + 1) * n], theta + (size_t)(((g * rows + r) % n + n) % n) * n, n * sizeof(double));
    phys[g] = make_phys(n, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], 0, tabs[g]);
  }
  const dim3 grid((n + kWarpCols * kBlockWarps - 1) / (kWarpCols * kBlockWarps), (rows + rs - 1) / rs);
  std::vector<std::thread> ranks;
  for (int g = 0; g < P; ++g)
    ranks.emplace_back([&, g] {
      unsigned state = seed * 2654435761u + 131u * (unsigned)g + 7u;
      auto nap = [&] {
        state = state * 1664525u + 1013904223u;
        std::this_thread::sleep_for(std::chrono::microseconds((state >> 16) % 2000));
      };
      const int prev = (g + P - 1) % P, next = (g + 1) % P;
      nap();
      emu::launch(dim3((n + 255) / 256), dim3(256), [&] {
        k_halo_push(xa[g].data(), 1, fs, rows, n, comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g],
                    &counter[g][0]);
      });
      double* src = xa[g].data();
      double* dst = xb[g].data();
      for (int step = 0; step <= sweeps; ++step) {
        nap();
        VecIn in{};
        in.x = src; in.fs = fs; in.hs = n; in.dseq = &dseq[g]; in.comm = comm[g].data(); in.area = area;
        in.land = land[g].data(); in.top = in.land; in.bot = in.land + 5 * n;
        const Geo geo{n, rows, g * rows, rs, 0};
        const PushOut po{comm[prev].data(), comm[next].data(), comm[g].data(), area, &dseq[g], &counter[g][4]};
        const bool last = step == sweeps;
        emu::launch(grid, dim3(kBlockThreads), [&] {
          if (!last) k_poisson<2, false, true>(in, thp[g].data(), bs[g].data(), dst, geo, phys[g], omega, ChebEp{}, po);
          else k_poisson<1>(in, thp[g].data(), bs[g].data(), dst, geo, phys[g], omega);
        });
        std::swap(src, dst);
      }
      std::memcpy(out + (size_t)g * rows * n, src, fs * sizeof(double));
    });
  for (auto& t : ranks) t.join();
}

}  // extern "C"
