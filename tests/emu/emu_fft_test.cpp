// Host-emulated run of the product's FFT kernels against a naive long-double DFT.
// TEST TOOL: compiled with g++ -DMRL_EMU; validates index math / barriers without a GPU.
#define MRL_EMU 1
#include "../../marlin_b200/csrc/mrl_passes_slab.cuh"
#include "../../marlin_b200/csrc/mrl_mech_tma.cuh"

#include <complex>
#include <random>

using namespace mrl;
typedef std::complex<long double> lc;
static const long double PI = 3.141592653589793238462643383279502884L;
static int g_fail = 0;

static std::vector<cx<double>> make_tw(int n) {
  std::vector<cx<double>> tw(n);
  for (int k = 0; k < n; ++k) {
    long double a = -2 * PI * k / n;
    tw[k] = mk<double>((double)cosl(a), (double)sinl(a));
  }
  return tw;
}
static std::vector<lc> dft(const std::vector<lc> &x, int sign) {
  int n = x.size();
  std::vector<lc> y(n);
  for (int k = 0; k < n; ++k) {
    lc acc = 0;
    for (int j = 0; j < n; ++j) acc += x[j] * std::polar(1.0L, sign * 2 * PI * ((long long)j * k % n) / n);
    y[k] = acc;
  }
  return y;
}
static FFTPlanDev make_plan(int n) {
  FFTPlanDev p;
  p.n = n;
  p.nstages = 0;
  int m = n;
  const int pref[] = {8, 4, 2, 3, 5};
  for (int r : pref)
    while (m % r == 0 && m > 1) {
      p.radix[p.nstages++] = r;
      m /= r;
    }
  for (int f = 7; m > 1; f += 2)
    while (m % f == 0) {
      p.radix[p.nstages++] = f;
      m /= f;
    }
  if (n == 1) { p.radix[0] = 1; p.nstages = 0; }
  return p;
}
static void report(const char *name, double err, double tol) {
  printf("%-44s err=%.3e %s\n", name, err, err < tol ? "ok" : "FAIL");
  if (!(err < tol)) g_fail++;
}

// ---- strided pass on a [nouter][n][ncols] array, forward or inverse
template <class C, int TK> static void run_strided_fast(StridedIO<double> io, const cx<double> *tw) {
  size_t smem = (size_t)(C::N * TK + C::N) * sizeof(cx<double>);
  emu::launch(dim3(3), dim3(TK * C::TP), smem, [=] { k_strided_fast<double, C, TK>(io, tw); });
}
static void run_strided_gen(StridedIO<double> io, const cx<double> *tw, FFTPlanDev plan) {
  size_t smem = (size_t)(2 * plan.n * 8) * sizeof(cx<double>);
  emu::launch(dim3(3), dim3(256), smem, [=] { k_strided_gen<double, 8>(io, tw, plan); });
}
template <class RUN> static void test_strided(const char *name, int n, int ncols, int nouter, int inverse, RUN run) {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1, 1);
  size_t total = (size_t)nouter * n * ncols;
  std::vector<cx<double>> a(total), out(total);
  for (auto &v : a) v = mk<double>(U(rng), U(rng));
  auto tw = make_tw(n);
  StridedIO<double> io{};
  io.in[0] = a.data();
  io.out[0] = out.data();
  io.nfields = 1;
  io.n = n;
  io.ncols = ncols;
  io.nouter = nouter;
  io.pitch = ncols;
  io.outer_stride = (long long)n * ncols;
  io.ncb = (ncols + 7) / 8;
  io.scale = inverse ? 1.0 / n : 1.0;
  io.inverse = inverse;
  run(io, tw.data());
  double err = 0;
  for (int o = 0; o < nouter; ++o)
    for (int c = 0; c < ncols; ++c) {
      std::vector<lc> x(n);
      for (int j = 0; j < n; ++j) {
        auto v = a[(size_t)o * n * ncols + (size_t)j * ncols + c];
        x[j] = lc(v.x, v.y);
      }
      auto y = dft(x, inverse ? +1 : -1);
      for (int j = 0; j < n; ++j) {
        auto v = out[(size_t)o * n * ncols + (size_t)j * ncols + c];
        lc ref = y[j] * (long double)io.scale;
        err = std::max(err, (double)std::abs(lc(v.x, v.y) - ref));
      }
    }
  report(name, err, 1e-12 * n);
}

// ---- r2c / c2r over rows
template <class RUNF, class RUNI> static void test_real(const char *name, int n, int nrows, RUNF runf, RUNI runi) {
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(-1, 1);
  int nc = n / 2 + 1;
  std::vector<double> a((size_t)nrows * n), back((size_t)nrows * n);
  for (auto &v : a) v = U(rng);
  std::vector<cx<double>> spec((size_t)nrows * nc);
  auto tw = make_tw(n);
  long long npen = (nrows + 1) / 2;
  ZLoadPairs<double> ld{a.data(), nrows, n};
  ZStorePairs<double> st{spec.data(), nrows, nc};
  runf(ld, st, tw.data(), npen);
  double err = 0;
  for (int r = 0; r < nrows; ++r) {
    std::vector<lc> x(n);
    for (int j = 0; j < n; ++j) x[j] = a[(size_t)r * n + j];
    auto y = dft(x, -1);
    for (int k = 0; k < nc; ++k) {
      auto v = spec[(size_t)r * nc + k];
      err = std::max(err, (double)std::abs(lc(v.x, v.y) - y[k]));
    }
  }
  char nm[128];
  snprintf(nm, sizeof nm, "%s r2c", name);
  report(nm, err, 1e-12 * n);
  // pollute imaginary parts of DC / Nyquist: must be ignored by c2r
  for (int r = 0; r < nrows; ++r) {
    spec[(size_t)r * nc].y = 0.37;
    if (n % 2 == 0) spec[(size_t)r * nc + n / 2].y = -0.21;
  }
  ZInvLoadPairs<double> ldi{spec.data(), nrows, n, nc};
  ZInvStorePairs<double> sti{back.data(), nrows, n, 1.0 / n};
  runi(ldi, sti, tw.data(), npen);
  err = 0;
  for (size_t i = 0; i < a.size(); ++i) err = std::max(err, std::fabs(a[i] - back[i]));
  snprintf(nm, sizeof nm, "%s c2r roundtrip", name);
  report(nm, err, 1e-12 * n);
}

template <class C, int PPB> struct RealFast {
  static void f(ZLoadPairs<double> ld, ZStorePairs<double> st, const cx<double> *tw, long long np) {
    size_t smem = (size_t)((C::N + C::N / 8 + 1) * PPB + C::N) * sizeof(cx<double>);
    emu::launch(dim3(2), dim3(PPB * C::TP), smem,
                [=] { k_zfwd_fast<double, C, PPB, ZLoadPairs<double>, ZStorePairs<double>>(ld, st, tw, np); });
  }
  static void i(ZInvLoadPairs<double> ld, ZInvStorePairs<double> st, const cx<double> *tw, long long np) {
    size_t smem = (size_t)((C::N + C::N / 8 + 1) * PPB + C::N) * sizeof(cx<double>);
    emu::launch(dim3(2), dim3(PPB * C::TP), smem,
                [=] { k_zinv_fast<double, C, PPB, ZInvLoadPairs<double>, ZInvStorePairs<double>>(ld, st, tw, np); });
  }
};
static void real_gen_f(int n, ZLoadPairs<double> ld, ZStorePairs<double> st, const cx<double> *tw, long long np) {
  FFTPlanDev plan = make_plan(n);
  size_t smem = (size_t)(2 * n * 4) * sizeof(cx<double>);
  emu::launch(dim3(2), dim3(256), smem,
              [=] { k_zfwd_gen<double, 4, ZLoadPairs<double>, ZStorePairs<double>>(ld, st, tw, plan, np); });
}
static void real_gen_i(int n, ZInvLoadPairs<double> ld, ZInvStorePairs<double> st, const cx<double> *tw, long long np) {
  FFTPlanDev plan = make_plan(n);
  size_t smem = (size_t)(2 * n * 4) * sizeof(cx<double>);
  emu::launch(dim3(2), dim3(256), smem,
              [=] { k_zinv_gen<double, 4, ZInvLoadPairs<double>, ZInvStorePairs<double>>(ld, st, tw, plan, np); });
}

// ---- fused pass vs composition of strided fwd + update + strided inv (computed with dft())
template <class RUN> static void test_fused(const char *name, int n, int ny, int nzc, int kmode, RUN run, int nold = 1) {
  std::mt19937_64 rng(5);
  std::uniform_real_distribution<double> U(-1, 1);
  int ncols = (kmode == MRL_KMODE_2D) ? ny : ny * nzc;
  size_t total = (size_t)n * ncols;
  std::vector<cx<double>> C(total), G(total), Uo(total), Nout(total), Nold0(total);
  for (auto &v : C) v = mk<double>(U(rng), U(rng));
  for (auto &v : G) v = mk<double>(U(rng), U(rng));
  for (auto &v : Nold0) v = mk<double>(U(rng), U(rng));
  std::vector<double> kx(n), ky(ny), kz(nzc);
  for (auto &v : kx) v = U(rng);
  for (auto &v : ky) v = U(rng);
  for (auto &v : kz) v = U(rng);
  auto tw = make_tw(n);
  FusedIO<double> io{C.data(), G.data(), Uo.data(), n, ncols, 1, ncols, 0, (ncols + 7) / 8, 1.0 / n};
  SpectralUpdate<double> up{};
  up.kx = kx.data(); up.ky = ky.data(); up.kz = kz.data();
  up.kmode = kmode; up.nzc = nzc; up.x0 = 0;
  up.closed_M = 1; up.closed_L = 1; up.has_L = 1;
  up.Mfac = 0.2; up.Lfac = -0.001; up.dt = 0.01;
  up.b0 = 1.5 * up.dt; up.nold = nold; up.bold[0] = -0.5 * up.dt; up.Nold[0] = Nold0.data();
  up.Nout = Nout.data();
  run(io, up, tw.data());
  double err = 0, errN = 0;
  for (int c = 0; c < ncols; ++c) {
    std::vector<lc> xc(n), xg(n);
    for (int j = 0; j < n; ++j) {
      xc[j] = lc(C[(size_t)j * ncols + c].x, C[(size_t)j * ncols + c].y);
      xg[j] = lc(G[(size_t)j * ncols + c].x, G[(size_t)j * ncols + c].y);
    }
    auto yc = dft(xc, -1), yg = dft(xg, -1);
    std::vector<lc> u(n);
    for (int j = 0; j < n; ++j) {
      long double a = kx[j], b = (kmode == MRL_KMODE_2D) ? ky[c] : ky[c / nzc], d = (kmode == MRL_KMODE_2D) ? 0 : kz[c % nzc];
      long double kk = a * a + b * b + d * d;
      lc N = (-kk * 0.2L) * yg[j];
      auto no = Nold0[(size_t)j * ncols + c];
      u[j] = (yc[j] + (long double)up.b0 * N + (nold ? (long double)up.bold[0] : 0.0L) * lc(no.x, no.y)) /
             (1.0L - (long double)up.dt * (kk * kk * -0.001L));
      auto nn = Nout[(size_t)j * ncols + c];
      errN = std::max(errN, (double)std::abs(lc(nn.x, nn.y) - N));
    }
    auto r = dft(u, +1);
    for (int j = 0; j < n; ++j) {
      auto v = Uo[(size_t)j * ncols + c];
      err = std::max(err, (double)std::abs(lc(v.x, v.y) - r[j] / (long double)n));
    }
  }
  char nm[128];
  snprintf(nm, sizeof nm, "%s u", name);
  report(nm, err, 1e-12 * n);
  snprintf(nm, sizeof nm, "%s N", name);
  report(nm, errN, 1e-12 * n);
}


// ======================================================================== TMA-pipelined kernels
static TensorMap emu_map(const void *base, int esize, long long d0, long long d1, long long d2, long long s1, long long s2,
                         int b0, int b1) {
  TensorMap m;
  m.base = (const unsigned char *)base;
  m.esize = esize;
  m.dim[0] = d0; m.dim[1] = d1; m.dim[2] = d2; m.dim[3] = 1; m.dim[4] = 1;
  m.stride[0] = esize; m.stride[1] = s1; m.stride[2] = s2; m.stride[3] = 0; m.stride[4] = 0;
  m.box[0] = b0; m.box[1] = b1; m.box[2] = 1; m.box[3] = 1; m.box[4] = 1;
  return m;
}

template <class C, int TK, int NG, int NS> static void run_strided_tma(StridedIO<double> io0, const cx<double> *tw, int grid) {
  StridedTmaIO<double> io;
  io.out = io0.out[0]; io.out1 = io0.out[0]; io.nouter_f = io0.nouter; io.nvalid = io0.ncols; io.peer_tab = nullptr; io.peer_rows = 0; io.peer_field = 0; io.peer_off = 0;
  io.n = io0.n; io.ncols = io0.ncols; io.nouter = io0.nouter;
  io.pitch = io0.pitch; io.outer_stride = io0.outer_stride;
  io.ncb = (io0.ncols + TK - 1) / TK;
  io.scale = io0.scale; io.inverse = io0.inverse;
  const long long rowb = io.pitch * 16;
  TensorMap tm = emu_map(io0.in[0], 8, 2LL * io.ncols, io.n, io.nouter, rowb, rowb * io.n, 2 * TK, C::N < 256 ? C::N : 256);
  size_t smem = (size_t)(NS * C::N * TK) * 16 + NS * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_strided_tma<double, C, TK, NG, NS>(tm, io, tw); }, 64 * 1024);
}

template <class C, int TK, int NG> static void run_fused_tma(FusedIO<double> io0, SpectralUpdate<double> up0, const cx<double> *tw, int grid) {
  FusedTmaIO<double> io;
  io.outU = io0.outU; io.n = io0.n; io.ncols = io0.ncols; io.ncb = (io0.ncols + TK - 1) / TK; io.pitch = io0.pitch; io.scale = io0.scale;
  io.slab = 0; io.nouter = 1; io.nouter_full = 1; io.nyl = 0; io.peer_tab = nullptr; io.peer_x0 = 0;
  SpectralUpdate2<double> up{};
  up.kx = up0.kx; up.ky = up0.ky; up.kz = up0.kz; up.kmode = up0.kmode; up.nzc = up0.nzc; up.nzv = up0.nzc; up.x0 = up0.x0;
  up.closed_M = up0.closed_M; up.closed_L = up0.closed_L; up.has_L = up0.has_L; up.Mfac = up0.Mfac; up.Lfac = up0.Lfac;
  up.dt = up0.dt; up.b0 = up0.b0; up.nold = up0.nold; up.bold0 = up0.bold[0]; up.Nout = up0.Nout;
  const long long rowb = io.pitch * 16;
  const int boxr = C::N < 256 ? C::N : 256;
  TensorMap tmC = emu_map(io0.inC, 8, 2LL * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
  TensorMap tmG = emu_map(io0.inG, 8, 2LL * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
  TensorMap tmO = emu_map(up0.nold ? (const void *)up0.Nold[0] : (const void *)io0.inC, 8, 2LL * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
  size_t smem = (size_t)(NG * 3 * C::N * TK) * 16 + NG * 3 * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_fused_tma<double, C, TK, NG>(tmC, tmG, tmO, io, up, tw); }, 64 * 1024);
}

// P1 (c + i F(c)) and P5 against the naive DFT
// `chunks` > 0: the rows are a slab [nx][nyl] and the pass runs once per y-chunk with a RowMap (the
// multi-GPU forward phase overlaps such chunk passes with the x passes of the previous chunk)
template <class C, int PPB, int NG, int NS> static void test_zfwd_tma(const char *name, int nrows, int grid, int nyl = 0, int ych = 0) {
  constexpr int n = C::N, nc = n / 2 + 1, NP = n + n / 8 + 1, ncp = (nc + 7) & ~7;
  std::mt19937_64 rng(13);
  std::uniform_real_distribution<double> U(0, 1);
  std::vector<double> c((size_t)nrows * n), mu((size_t)nrows * n, -7.0);
  for (auto &v : c) v = U(rng);
  std::vector<cx<double>> oc((size_t)nrows * ncp), og((size_t)nrows * ncp);
  auto tw = make_tw(n);
  DoubleWellDeriv<double> f{0.1, 0.0, 1.0};
  size_t smem = (size_t)NG * NS * PPB * n * 8 + (size_t)(NG * PPB * NP) * 16 + NG * NS * 8 + 128;
  const double *cp = c.data();
  double *mp = mu.data();
  cx<double> *ocp = oc.data(), *ogp = og.data();
  const cx<double> *twp = tw.data();
  if (!ych) {
    emu::launch(dim3(grid), dim3(NG * PPB * C::TP), smem,
                [=] { k_zfwd_tma<double, C, PPB, NG, NS, DoubleWellDeriv<double>>(cp, mp, ocp, ogp, nrows, ncp, f, twp, RowMap{0, 0, 0}); }, 64 * 1024);
  } else {
    const int nx = nrows / nyl;
    for (int y0 = 0; y0 < nyl; y0 += ych) {
      const int w = std::min(ych, nyl - y0);
      const long long rows = (long long)nx * w;
      const RowMap rm{w, nyl, y0};
      emu::launch(dim3(grid), dim3(NG * PPB * C::TP), smem,
                  [=] { k_zfwd_tma<double, C, PPB, NG, NS, DoubleWellDeriv<double>>(cp, mp, ocp, ogp, rows, ncp, f, twp, rm); }, 64 * 1024);
    }
  }
  double err = 0;
  for (int r = 0; r < nrows; ++r) {
    std::vector<lc> x(n), g(n);
    for (int j = 0; j < n; ++j) {
      double v = c[(size_t)r * n + j];
      x[j] = v;
      g[j] = f(v);
      err = std::max(err, std::fabs(mu[(size_t)r * n + j] - f(v)));
    }
    auto yx = dft(x, -1), yg = dft(g, -1);
    for (int k = 0; k < nc; ++k) {
      auto a = oc[(size_t)r * ncp + k], b = og[(size_t)r * ncp + k];
      err = std::max(err, (double)std::abs(lc(a.x, a.y) - yx[k]));
      err = std::max(err, (double)std::abs(lc(b.x, b.y) - yg[k]));
    }
  }
  report(name, err, 1e-12 * n);
}

template <class C, int PPB, int NG, int NS> static void test_zinv_tma(const char *name, int nrows, int grid) {
  constexpr int n = C::N, nc = n / 2 + 1, NP = n + n / 8 + 1, ncp = (nc + 7) & ~7, ncmax = (nc + 15) & ~15;
  std::mt19937_64 rng(17);
  std::uniform_real_distribution<double> U(-1, 1);
  std::vector<double> a((size_t)nrows * n), back((size_t)nrows * n);
  for (auto &v : a) v = U(rng);
  std::vector<cx<double>> spec((size_t)nrows * ncp);
  for (int r = 0; r < nrows; ++r) {
    std::vector<lc> x(n);
    for (int j = 0; j < n; ++j) x[j] = a[(size_t)r * n + j];
    auto y = dft(x, -1);
    for (int k = 0; k < nc; ++k) spec[(size_t)r * ncp + k] = mk<double>((double)y[k].real(), (double)y[k].imag());
    spec[(size_t)r * ncp].y = 0.37;  // must be ignored
    spec[(size_t)r * ncp + n / 2].y = -0.21;
  }
  auto tw = make_tw(n);
  size_t smem = (size_t)(NG * NS * 2 * PPB * ncmax + NG * PPB * NP) * 16 + NG * NS * 8 + 128;
  const cx<double> *sp = spec.data(), *twp = tw.data();
  double *bp = back.data();
  emu::launch(dim3(grid), dim3(NG * PPB * C::TP), smem, [=] { k_zinv_tma<double, C, PPB, NG, NS>(sp, ncp, bp, nrows, 1.0 / n, twp, nullptr, nullptr); },
              64 * 1024);
  double err = 0;
  for (size_t i = 0; i < a.size(); ++i) err = std::max(err, std::fabs(a[i] - back[i]));
  report(name, err, 1e-12 * n);
  // the same pass with the inner product sum(result * w) fused into the store (mechanics: p.Ap)
  std::vector<double> w(a.size()), back2(a.size()), partials(grid, 0.0);
  for (auto &v : w) v = U(rng);
  const double *wp = w.data();
  double *b2 = back2.data(), *pp = partials.data();
  emu::launch(dim3(grid), dim3(NG * PPB * C::TP), smem, [=] { k_zinv_tma<double, C, PPB, NG, NS, true>(sp, ncp, b2, nrows, 1.0 / n, twp, wp, pp); },
              64 * 1024);
  double want = 0, got = 0, e2 = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    want += back[i] * w[i];
    e2 = std::max(e2, std::fabs(back[i] - back2[i]));
  }
  for (double v : partials) got += v;
  report((std::string(name) + " + dot").c_str(), std::max(e2, std::fabs(want - got)), 1e-12 * n);
}

// mechanics: first pass with the tangent K4(F) : p (and the direction update p = r + beta p) fused into its load
template <class C, int NG> static void test_mech_tangent(const char *name, int nrows, int grid, bool update, bool staged = false) {
  constexpr int n = C::N, nc = n / 2 + 1, NP = n + n / 8 + 1, ncp = (nc + 7) & ~7;
  const long long nv = (long long)nrows * n;
  std::mt19937_64 rng(23);
  std::uniform_real_distribution<double> U(-1, 1);
  std::vector<double> F(9 * nv), P(9 * nv), R(9 * nv), K(nv), MU(nv), scal(8, 0.0);
  for (long long v = 0; v < nv; ++v) {
    for (int c = 0; c < 9; ++c) {
      F[c * nv + v] = (c / 3 == c % 3 ? 1.0 : 0.0) + 0.1 * U(rng);
      P[c * nv + v] = U(rng);
      R[c * nv + v] = U(rng);
    }
    K[v] = 0.8 + 0.1 * U(rng);
    MU[v] = 0.4 + 0.1 * U(rng);
  }
  scal[4] = 0.37;
  // expected: direction update, tangent, r2c per row
  std::vector<double> pn = P, tmp(9 * nv);
  for (long long v = 0; v < nv; ++v) {
    MD<double, 3> Fm, X, Rm;
    for (int c = 0; c < 9; ++c) {
      if (update) pn[c * nv + v] = R[c * nv + v] + scal[4] * P[c * nv + v];
      Fm.a[c / 3][c % 3] = F[c * nv + v];
      X.a[c / 3][c % 3] = pn[c * nv + v];
    }
    mech_point<double, 3>(1, Fm, K[v], MU[v], X, Rm);
    for (int c = 0; c < 9; ++c) tmp[c * nv + v] = Rm.a[c / 3][c % 3];
  }
  std::vector<cx<double>> out((size_t)9 * nrows * ncp, mk<double>(0.0, 0.0));
  auto tw = make_tw(n);
  MechTangentIO<double> io;
  io.F = F.data();
  io.K = K.data();
  io.mu = MU.data();
  io.p = P.data();
  io.r = update ? R.data() : nullptr;
  io.scal = scal.data();
  io.n = nv;
  io.nrows = nrows;
  io.out = out.data();
  io.ncp = ncp;
  const cx<double> *twp = tw.data();
  size_t smem = (size_t)NG * 9 * 2 * n * 8 + (size_t)NG * 9 * NP * 16 + 128;
  if (staged) {
    smem = (size_t)(2 * 29 + 18) * n * 8 + (size_t)9 * NP * 16 + 16 + 128;
    emu::launch(dim3(grid), dim3(9 * C::TP), smem, [=] { k_mech_tangent_zfwd_tma<double, C>(io, twp); }, 64 * 1024);
  } else
  emu::launch(dim3(grid), dim3(NG * 9 * C::TP), smem, [=] { k_mech_tangent_zfwd<double, C, NG>(io, twp); }, 64 * 1024);
  double err = 0;
  for (size_t i = 0; i < pn.size(); ++i) err = std::max(err, std::fabs(pn[i] - P[i]));
  for (int c = 0; c < 9; ++c)
    for (int r = 0; r < nrows; ++r) {
      std::vector<lc> x(n);
      for (int j = 0; j < n; ++j) x[j] = tmp[c * nv + (long long)r * n + j];
      auto y = dft(x, -1);
      for (int k = 0; k < nc; ++k) {
        auto a = out[((size_t)c * nrows + r) * ncp + k];
        err = std::max(err, (double)std::abs(lc(a.x, a.y) - y[k]));
      }
    }
  report(name, err, 1e-11 * n);
}

// fused pass in the multi-GPU slab layout: data staged as [P][nxl][nyl][nzc], transform along y
template <class C, int TK, int NG>
static void test_fused_slab(const char *name, int P, int nxl, int nzc, int x0, int grid, bool peer = false, int xchunks = 1) {
  constexpr int ny = C::N;
  const int nyl = ny / P;
  std::mt19937_64 rng(23);
  std::uniform_real_distribution<double> U(-1, 1);
  const size_t total = (size_t)nxl * ny * nzc;
  // logical [x][y][kz] arrays and their staged copies
  std::vector<cx<double>> Cl(total), Gl(total), Ol(total), Cs(total), Gs(total), Os(total), Us(total), Ns(total);
  for (auto &v : Cl) v = mk<double>(U(rng), U(rng));
  for (auto &v : Gl) v = mk<double>(U(rng), U(rng));
  for (auto &v : Ol) v = mk<double>(U(rng), U(rng));
  auto sidx = [&](int x, int y, int kz) { return (((size_t)(y / nyl) * nxl + x) * nyl + (y % nyl)) * nzc + kz; };
  for (int x = 0; x < nxl; ++x)
    for (int y = 0; y < ny; ++y)
      for (int kz = 0; kz < nzc; ++kz) {
        size_t l = ((size_t)x * ny + y) * nzc + kz;
        Cs[sidx(x, y, kz)] = Cl[l]; Gs[sidx(x, y, kz)] = Gl[l]; Os[sidx(x, y, kz)] = Ol[l];
      }
  std::vector<double> kx(x0 + nxl + 3), ky(ny), kz(nzc);
  for (auto &v : kx) v = U(rng);
  for (auto &v : ky) v = U(rng);
  for (auto &v : kz) v = U(rng);
  auto tw = make_tw(ny);
  FusedTmaIO<double> io;
  io.outU = Us.data(); io.n = ny; io.ncols = nzc; io.ncb = (nzc + TK - 1) / TK; io.pitch = nzc; io.scale = 1.0 / ny;
  io.slab = 1; io.nouter = nxl; io.nouter_full = nxl; io.nyl = nyl; io.peer_tab = nullptr; io.peer_x0 = 0;
  io.kzb_major = 0; io.nx = 0; io.rank = 0; io.nranks = P; io.flag_wait = nullptr; io.flag_expect = 0; io.flag_tab = nullptr;
  SpectralUpdate2<double> up{};
  up.kx = kx.data(); up.ky = ky.data(); up.kz = kz.data(); up.kmode = MRL_KMODE_3D_SLAB; up.nzc = nzc; up.nzv = nzc; up.x0 = x0;
  up.closed_M = 1; up.closed_L = 1; up.has_L = 1; up.Mfac = 0.2; up.Lfac = -0.001; up.dt = 0.01;
  up.b0 = 1.5 * up.dt; up.nold = 1; up.bold0 = -0.5 * up.dt; up.Nout = Ns.data();
  auto mk4 = [&](const void *base) {
    TensorMap m;
    m.base = (const unsigned char *)base; m.esize = 8;
    m.dim[0] = 2LL * nzc; m.dim[1] = nyl; m.dim[2] = nxl; m.dim[3] = P;
    m.stride[0] = 8; m.stride[1] = 16LL * nzc; m.stride[2] = 16LL * nzc * nyl; m.stride[3] = 16LL * nzc * nyl * nxl;
    m.box[0] = 2 * TK; m.box[1] = nyl; m.box[2] = 1; m.box[3] = P;
    m.dim[4] = 1; m.stride[4] = 0; m.box[4] = 1;
    return m;
  };
  TensorMap tmC = mk4(Cs.data()), tmG = mk4(Gs.data()), tmO = mk4(Os.data());
  const cx<double> *twp = tw.data();
  // peer mode: the result rows go to P separate "rank" arrays [nxtot][nyl][nzc] at x = peer_x0 + o
  const int nxtot = x0 + nxl + 1;
  std::vector<std::vector<cx<double>>> peerbuf(P, std::vector<cx<double>>((size_t)nxtot * nyl * nzc));
  std::vector<unsigned long long> tab(P);
  for (int q = 0; q < P; ++q) tab[q] = (unsigned long long)peerbuf[q].data();
  if (peer) { io.peer_tab = tab.data(); io.peer_x0 = x0; }
  size_t smem = (size_t)(NG * 3 * C::N * TK) * 16 + NG * 3 * 8 + 128;
  if (xchunks > 1) {
    // the pass in x sub-ranges (copy-engine exchange mode): every base shifted by o0 layers, the rank blocks keep their distance
    const int xch = nxl / xchunks;
    for (int i = 0; i < xchunks; ++i) {
      const size_t sh = (size_t)i * xch * nyl * nzc;
      FusedTmaIO<double> ioc = io;
      SpectralUpdate2<double> upc = up;
      ioc.outU = io.outU + sh; ioc.nouter = xch; ioc.nouter_full = nxl;
      upc.x0 = up.x0 + i * xch; upc.Nout = up.Nout + sh;
      auto sub = [&](TensorMap m) { m.base += sh * 16; m.dim[2] = xch; return m; };
      TensorMap c = sub(tmC), g = sub(tmG), o = sub(tmO);
      emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_fused_tma<double, C, TK, NG, 1>(c, g, o, ioc, upc, twp); }, 64 * 1024);
    }
  } else
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_fused_tma<double, C, TK, NG, 1>(tmC, tmG, tmO, io, up, twp); }, 64 * 1024);
  if (peer)
    for (int x = 0; x < nxl; ++x)
      for (int y = 0; y < ny; ++y)
        for (int kz = 0; kz < nzc; ++kz)
          Us[sidx(x, y, kz)] = peerbuf[y / nyl][((size_t)(x0 + x) * nyl + (y % nyl)) * nzc + kz];
  double err = 0, errN = 0;
  for (int x = 0; x < nxl; ++x)
    for (int q = 0; q < nzc; ++q) {
      std::vector<lc> xc(ny), xg(ny);
      for (int y = 0; y < ny; ++y) {
        auto a = Cl[((size_t)x * ny + y) * nzc + q], b = Gl[((size_t)x * ny + y) * nzc + q];
        xc[y] = lc(a.x, a.y); xg[y] = lc(b.x, b.y);
      }
      auto yc = dft(xc, -1), yg = dft(xg, -1);
      std::vector<lc> u(ny);
      for (int y = 0; y < ny; ++y) {
        long double a = kx[x0 + x], b = ky[y], d = kz[q];
        long double kk = a * a + b * b + d * d;
        lc N = (-kk * 0.2L) * yg[y];
        auto no = Ol[((size_t)x * ny + y) * nzc + q];
        u[y] = (yc[y] + (long double)up.b0 * N + (long double)up.bold0 * lc(no.x, no.y)) / (1.0L - (long double)up.dt * (kk * kk * -0.001L));
        auto nn = Ns[sidx(x, y, q)];
        errN = std::max(errN, (double)std::abs(lc(nn.x, nn.y) - N));
      }
      auto r = dft(u, +1);
      for (int y = 0; y < ny; ++y) {
        auto v = Us[sidx(x, y, q)];
        err = std::max(err, (double)std::abs(lc(v.x, v.y) - r[y] / (long double)ny));
      }
    }
  char nm[128];
  snprintf(nm, sizeof nm, "%s u", name);
  report(nm, err, 1e-12 * ny);
  snprintf(nm, sizeof nm, "%s N", name);
  report(nm, errN, 1e-12 * ny);
}

// strided pass with the result rows scattered to P "ranks" (fused forward all-to-all)
template <class C, int TK, int NG, int NS> static void test_strided_peer(const char *name, int P, int ncols) {
  constexpr int nx = C::N;
  const int nxl = nx / P, me = 1;
  std::mt19937_64 rng(29);
  std::uniform_real_distribution<double> U(-1, 1);
  const size_t field = (size_t)nx * ncols;
  std::vector<cx<double>> in(2 * field);
  for (auto &v : in) v = mk<double>(U(rng), U(rng));
  // every rank's staging: [2 fields][P sources][nxl][ncols]
  std::vector<std::vector<cx<double>>> stage(P, std::vector<cx<double>>(2 * field));
  std::vector<unsigned long long> tab(P);
  for (int q = 0; q < P; ++q) tab[q] = (unsigned long long)stage[q].data();
  auto tw = make_tw(nx);
  StridedTmaIO<double> io;
  io.out = io.out1 = nullptr; io.nouter_f = 1;
  io.n = nx; io.ncols = ncols; io.nouter = 2; io.nvalid = ncols;
  io.pitch = ncols; io.outer_stride = (long long)nx * ncols;
  io.ncb = (ncols + TK - 1) / TK; io.scale = 1.0; io.inverse = 0;
  io.peer_tab = tab.data(); io.peer_rows = nxl; io.peer_field = field; io.peer_off = (long long)me * nxl * ncols;
  const long long rowb = (long long)ncols * 16;
  TensorMap tm = emu_map(in.data(), 8, 2LL * ncols, nx, 2, rowb, rowb * nx, 2 * TK, nx < 256 ? nx : 256);
  size_t smem = (size_t)(NS * nx * TK) * 16 + NS * 8 + 128;
  const cx<double> *twp = tw.data();
  emu::launch(dim3(3), dim3(NG * TK * C::TP), smem, [=] { k_strided_tma<double, C, TK, NG, NS>(tm, io, twp); }, 64 * 1024);
  double err = 0;
  for (int f = 0; f < 2; ++f)
    for (int c = 0; c < ncols; ++c) {
      std::vector<lc> x(nx);
      for (int j = 0; j < nx; ++j) x[j] = lc(in[f * field + (size_t)j * ncols + c].x, in[f * field + (size_t)j * ncols + c].y);
      auto y = dft(x, -1);
      for (int j = 0; j < nx; ++j) {
        auto v = stage[j / nxl][f * field + ((size_t)me * nxl + j % nxl) * ncols + c];
        err = std::max(err, (double)std::abs(lc(v.x, v.y) - y[j]));
      }
    }
  report(name, err, 1e-12 * nx);
}

// mechanics: x-forward + Green projection + x-inverse in one pass vs dft() composition
template <class C, int TK, int NG> static void test_mech_fused(const char *name, int ny, int nzv, int ncp, int grid) {
  constexpr int n = C::N;
  std::mt19937_64 rng(23);
  std::uniform_real_distribution<double> U(-1, 1);
  const int ncols = ny * ncp;
  const size_t field = (size_t)n * ncols;
  std::vector<cx<double>> A(9 * field), A0;
  for (auto &v : A) v = mk<double>(U(rng), U(rng));
  A0 = A;
  std::vector<double> kx(n), ky(ny), kz(ncp);
  for (auto &v : kx) v = U(rng);
  for (auto &v : ky) v = U(rng);
  for (auto &v : kz) v = U(rng);
  kx[0] = ky[0] = kz[0] = 0.0;  // q = 0 -> projection 0
  auto tw = make_tw(n);
  MechFusedTmaIO<double> io;
  io.out = A.data(); io.n = n; io.ncols = ncols; io.ncb = (ncols + TK - 1) / TK; io.pitch = ncols; io.field = (long long)field;
  io.kx = kx.data(); io.ky = ky.data(); io.kz = kz.data(); io.ncp = ncp; io.nzv = nzv; io.scale = 1.0;
  const long long rowb = (long long)ncols * 16;
  TensorMap tm = emu_map(A.data(), 8, 2LL * ncols, n, 9, rowb, rowb * n, 2 * TK, n < 256 ? n : 256);
  const cx<double> *twp = tw.data();
  size_t smem = (size_t)(NG * 3 * n * TK) * 16 + NG * 3 * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_mech_fused_tma<double, C, TK, NG>(tm, io, twp); }, 64 * 1024);
  double err = 0, pad = 0;
  for (int c = 0; c < ncols; ++c) {
    const bool valid = (c % ncp) < nzv;
    for (int i = 0; i < 3; ++i) {
      std::vector<lc> y[3];
      for (int l = 0; l < 3; ++l) {
        std::vector<lc> x(n);
        for (int j = 0; j < n; ++j) x[j] = lc(A0[(3 * i + l) * field + (size_t)j * ncols + c].x, A0[(3 * i + l) * field + (size_t)j * ncols + c].y);
        y[l] = dft(x, -1);
      }
      for (int jj = 0; jj < 3; ++jj) {
        std::vector<lc> u(n);
        for (int j = 0; j < n; ++j) {
          long double q[3] = {kx[j], ky[c / ncp], kz[c % ncp]};
          long double Q = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
          lc sacc = y[0][j] * q[0] + y[1][j] * q[1] + y[2][j] * q[2];
          u[j] = Q == 0 ? lc(0, 0) : sacc * (q[jj] / Q);
        }
        auto r = dft(u, +1);
        for (int j = 0; j < n; ++j) {
          auto v = A[(3 * i + jj) * field + (size_t)j * ncols + c];
          auto v0 = A0[(3 * i + jj) * field + (size_t)j * ncols + c];
          if (valid) err = std::max(err, (double)std::abs(lc(v.x, v.y) - r[j]));
          else pad = std::max(pad, std::abs(v.x - v0.x) + std::abs(v.y - v0.y));  // padding columns untouched
        }
      }
    }
  }
  report(name, err, 1e-11 * n);
  char nm[160];
  snprintf(nm, sizeof nm, "%s (padding untouched)", name);
  report(nm, pad, 1e-300);
}

// fused pass on PADDED work spectra (row pitch ncp > nzv) with the mobility / linear operator read from
// caller buffers in their natural, unpadded layout [n][ny][nzv] (kmode 3D) or [n][nzv] (kmode 2D)
template <class C, int TK, int NG> static void test_fused_padded_buffers(const char *name, int kmode, int ny, int nzv, int ncp, int grid) {
  constexpr int n = C::N;
  std::mt19937_64 rng(31);
  std::uniform_real_distribution<double> U(-1, 1);
  const int nyy = kmode == MRL_KMODE_2D ? 1 : ny;
  const int ncols = nyy * ncp;
  const size_t total = (size_t)n * ncols, mtotal = (size_t)n * nyy * nzv;
  std::vector<cx<double>> Cc(total), G(total), Uo(total), Nout(total), Nold0(total);
  for (size_t i = 0; i < total; ++i) {
    const bool valid = (int)(i % ncp) < nzv;
    Cc[i] = valid ? mk<double>(U(rng), U(rng)) : mk<double>(0, 0);
    G[i] = valid ? mk<double>(U(rng), U(rng)) : mk<double>(0, 0);
    Nold0[i] = valid ? mk<double>(U(rng), U(rng)) : mk<double>(0, 0);
  }
  std::vector<double> Mb(mtotal), Lb(mtotal), kx(n, 0.0), ky(std::max(nyy, ncp), 0.0), kz(ncp, 0.0);
  for (auto &v : Mb) v = U(rng);
  for (auto &v : Lb) v = -std::fabs(U(rng));
  auto tw = make_tw(n);
  FusedTmaIO<double> io;
  io.outU = Uo.data(); io.n = n; io.ncols = ncols; io.ncb = (ncols + TK - 1) / TK; io.pitch = ncols; io.scale = 1.0 / n;
  io.slab = 0; io.nouter = 1; io.nouter_full = 1; io.nyl = 0; io.peer_tab = nullptr; io.peer_x0 = 0;
  SpectralUpdate2<double> up{};
  up.kx = kx.data(); up.ky = ky.data(); up.kz = kz.data(); up.kmode = kmode; up.nzc = ncp; up.nzv = nzv; up.x0 = 0;
  up.closed_M = 0; up.closed_L = 0; up.has_L = 1; up.Mbuf = Mb.data(); up.Lbuf = Lb.data();
  up.dt = 0.01; up.b0 = 1.5 * up.dt; up.nold = 1; up.bold0 = -0.5 * up.dt; up.Nout = Nout.data();
  const long long rowb = (long long)ncols * 16;
  const int boxr = n < 256 ? n : 256;
  TensorMap tmC = emu_map(Cc.data(), 8, 2LL * ncols, n, 1, rowb, rowb * n, 2 * TK, boxr);
  TensorMap tmG = emu_map(G.data(), 8, 2LL * ncols, n, 1, rowb, rowb * n, 2 * TK, boxr);
  TensorMap tmO = emu_map(Nold0.data(), 8, 2LL * ncols, n, 1, rowb, rowb * n, 2 * TK, boxr);
  const cx<double> *twp = tw.data();
  size_t smem = (size_t)(NG * 3 * n * TK) * 16 + NG * 3 * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_fused_tma<double, C, TK, NG>(tmC, tmG, tmO, io, up, twp); }, 64 * 1024);
  double err = 0;
  for (int c = 0; c < ncols; ++c) {
    if (c % ncp >= nzv) continue;
    std::vector<lc> xc(n), xg(n);
    for (int j = 0; j < n; ++j) {
      xc[j] = lc(Cc[(size_t)j * ncols + c].x, Cc[(size_t)j * ncols + c].y);
      xg[j] = lc(G[(size_t)j * ncols + c].x, G[(size_t)j * ncols + c].y);
    }
    auto yc = dft(xc, -1), yg = dft(xg, -1);
    std::vector<lc> u(n);
    for (int j = 0; j < n; ++j) {
      const size_t m = ((size_t)j * nyy + c / ncp) * nzv + c % ncp;
      lc N = (long double)Mb[m] * yg[j];
      auto no = Nold0[(size_t)j * ncols + c];
      u[j] = (yc[j] + (long double)up.b0 * N + (long double)up.bold0 * lc(no.x, no.y)) / (1.0L - (long double)up.dt * (long double)Lb[m]);
      auto nn = Nout[(size_t)j * ncols + c];
      err = std::max(err, (double)std::abs(lc(nn.x, nn.y) - N));
    }
    auto r = dft(u, +1);
    for (int j = 0; j < n; ++j) {
      auto v = Uo[(size_t)j * ncols + c];
      err = std::max(err, (double)std::abs(lc(v.x, v.y) - r[j] / (long double)n));
    }
  }
  report(name, err, 1e-12 * n);
}

// ---- multi-GPU slab passes with bulk peer stores into the blocked staging layouts (mrl_passes_slab.cuh)
// forward x pass of `me`: R_d[f][me][kzb][yl][xl][W] = DFT_x(in)[f][d*nxl + xl][yl][kzb*W + w]
template <class C, int TK, int NG, int NS> static void test_slab_xfwd(const char *name, int P, int nyl, int kb, int kzb_major, int chunks, int grid) {
  constexpr int nx = C::N;
  const int nxl = nx / P, me = 1 % P, ncp = kb * TK;
  std::mt19937_64 rng(31);
  std::uniform_real_distribution<double> U(-1, 1);
  const size_t field = (size_t)nx * nyl * ncp;
  std::vector<cx<double>> in(2 * field);
  for (auto &v : in) v = mk<double>(U(rng), U(rng));
  std::vector<std::vector<cx<double>>> R(P, std::vector<cx<double>>(2 * field));
  std::vector<std::vector<unsigned long long>> cnt(P, std::vector<unsigned long long>((size_t)P * kb, 0));
  std::vector<unsigned long long> tab(P), ftab(P);
  for (int q = 0; q < P; ++q) { tab[q] = (unsigned long long)R[q].data(); ftab[q] = (unsigned long long)cnt[q].data(); }
  auto tw = make_tw(nx);
  const cx<double> *twp = tw.data();
  const long long rowb = (long long)nyl * ncp * 16;
  TensorMap tm = emu_map(in.data(), 8, 2LL * nyl * ncp, nx, 2, rowb, rowb * nx, 2 * TK, nx < 256 ? nx : 256);
  size_t smem = (size_t)((NS + NG) * nx * TK) * 16 + NS * 8 + 128;
  const int ych = nyl / chunks;
  for (int ci = 0; ci < chunks; ++ci) {
    SlabXIO<double> io{};
    io.n = nx; io.nyl = nyl; io.kb = kb; io.nf = 2; io.nranks = P; io.rank = me; io.nxl = nxl; io.y0 = ci * ych; io.ych = ych;
    io.kzb_major = kzb_major; io.scale = 1.0; io.peer_tab = tab.data(); io.field = (long long)field; io.flag_tab = ftab.data();
    emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_slab_xfwd<double, C, TK, NG, NS>(tm, io, twp); }, 64 * 1024);
  }
  double err = 0;
  for (int f = 0; f < 2; ++f)
    for (int yl = 0; yl < nyl; ++yl)
      for (int c = 0; c < ncp; ++c) {
        std::vector<lc> x(nx);
        for (int j = 0; j < nx; ++j) {
          auto v = in[f * field + ((size_t)j * nyl + yl) * ncp + c];
          x[j] = lc(v.x, v.y);
        }
        auto y = dft(x, -1);
        for (int j = 0; j < nx; ++j) {
          const int d = j / nxl, xl = j % nxl, kzb = c / TK, w = c % TK;
          auto v = R[d][f * field + ((((size_t)me * kb + kzb) * nyl + yl) * nxl + xl) * TK + w];
          err = std::max(err, (double)std::abs(lc(v.x, v.y) - y[j]));
        }
      }
  report(name, err, 1e-12 * nx);
  // every destination has seen 2 * nyl arrivals from `me` for every column block
  bool ok = true;
  for (int d = 0; d < P; ++d)
    for (int q = 0; q < P; ++q)
      for (int k = 0; k < kb; ++k) ok = ok && cnt[d][(size_t)q * kb + k] == (q == me ? 2ull * nyl : 0ull);
  char nm[160];
  snprintf(nm, sizeof nm, "%s arrival counters", name);
  report(nm, ok ? 0.0 : 1.0, 0.5);
}

// inverse x pass: S = [kb][nx][nyl][W] -> natural [nx][nyl][ncp]
template <class C, int TK, int NG, int NS> static void test_slab_xinv(const char *name, int nyl, int kb, int kzb_major, int grid) {
  constexpr int nx = C::N;
  const int ncp = kb * TK;
  std::mt19937_64 rng(37);
  std::uniform_real_distribution<double> U(-1, 1);
  const size_t field = (size_t)nx * nyl * ncp;
  std::vector<cx<double>> S(field), out(field);
  for (auto &v : S) v = mk<double>(U(rng), U(rng));
  std::vector<unsigned long long> cnt((size_t)2 * kb, 5);
  auto tw = make_tw(nx);
  const cx<double> *twp = tw.data();
  TensorMap tm;
  tm.base = (const unsigned char *)S.data(); tm.esize = 8;
  tm.dim[0] = 2 * TK; tm.dim[1] = nyl; tm.dim[2] = nx; tm.dim[3] = kb; tm.dim[4] = 1;
  tm.stride[0] = 8; tm.stride[1] = 16LL * TK; tm.stride[2] = 16LL * TK * nyl; tm.stride[3] = 16LL * TK * nyl * nx; tm.stride[4] = 0;
  tm.box[0] = 2 * TK; tm.box[1] = 1; tm.box[2] = nx < 256 ? nx : 256; tm.box[3] = 1; tm.box[4] = 1;
  SlabXIO<double> io{};
  io.n = nx; io.nyl = nyl; io.kb = kb; io.nf = 1; io.nranks = 2; io.rank = 0; io.nxl = nx / 2; io.kzb_major = kzb_major; io.scale = 0.5;
  io.out = out.data(); io.out_pitch = (long long)nyl * ncp; io.flag_wait = cnt.data(); io.flag_expect = 5;
  size_t smem = (size_t)(NS * nx * TK) * 16 + NS * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_slab_xinv<double, C, TK, NG, NS>(tm, io, twp); }, 64 * 1024);
  double err = 0;
  for (int yl = 0; yl < nyl; ++yl)
    for (int c = 0; c < ncp; ++c) {
      std::vector<lc> x(nx);
      for (int j = 0; j < nx; ++j) {
        auto v = S[(((size_t)(c / TK) * nx + j) * nyl + yl) * TK + c % TK];
        x[j] = lc(v.x, v.y);
      }
      auto y = dft(x, +1);
      for (int j = 0; j < nx; ++j) {
        auto v = out[((size_t)j * nyl + yl) * ncp + c];
        err = std::max(err, (double)std::abs(lc(v.x, v.y) - y[j] * 0.5L));
      }
    }
  report(name, err, 1e-12 * nx);
}

// fused y pass on the blocked staging: inputs R = [P][kb][nyl][nxl][W], old term / new term in the staged layout
// [P][nxl][nyl][ncp], result rows into the P return stagings S_d = [kb][nxtot][nyl][W] at x = x0 + o
template <class C, int TK, int NG> static void test_fused_slab2(const char *name, int P, int nxl, int kb, int x0, int kzb_major, int grid, int nold = 1) {
  constexpr int ny = C::N;
  const int nyl = ny / P, ncp = kb * TK, me = P - 1;
  std::mt19937_64 rng(41);
  std::uniform_real_distribution<double> U(-1, 1);
  const size_t total = (size_t)nxl * ny * ncp;
  std::vector<cx<double>> Cl(total), Gl(total), Ol(total), Cb(total), Gb(total), Os(total), Ns(total);
  for (auto &v : Cl) v = mk<double>(U(rng), U(rng));
  for (auto &v : Gl) v = mk<double>(U(rng), U(rng));
  for (auto &v : Ol) v = mk<double>(U(rng), U(rng));
  auto sidx = [&](int x, int y, int kz) { return (((size_t)(y / nyl) * nxl + x) * nyl + (y % nyl)) * ncp + kz; };
  auto bidx = [&](int x, int y, int kz) { return ((((size_t)(y / nyl) * kb + kz / TK) * nyl + (y % nyl)) * nxl + x) * TK + kz % TK; };
  for (int x = 0; x < nxl; ++x)
    for (int y = 0; y < ny; ++y)
      for (int kz = 0; kz < ncp; ++kz) {
        size_t l = ((size_t)x * ny + y) * ncp + kz;
        Cb[bidx(x, y, kz)] = Cl[l]; Gb[bidx(x, y, kz)] = Gl[l]; Os[sidx(x, y, kz)] = Ol[l];
      }
  std::vector<double> kx(x0 + nxl + 3), ky(ny), kz(ncp);
  for (auto &v : kx) v = U(rng);
  for (auto &v : ky) v = U(rng);
  for (auto &v : kz) v = U(rng);
  auto tw = make_tw(ny);
  const int nxtot = x0 + nxl + 1;
  std::vector<std::vector<cx<double>>> Sd(P, std::vector<cx<double>>((size_t)kb * nxtot * nyl * TK));
  std::vector<std::vector<unsigned long long>> cnt2(P, std::vector<unsigned long long>((size_t)P * kb, 0));
  std::vector<unsigned long long> tab(P), ftab(P), cnt1((size_t)P * kb, 7);
  for (int q = 0; q < P; ++q) { tab[q] = (unsigned long long)Sd[q].data(); ftab[q] = (unsigned long long)cnt2[q].data(); }
  FusedTmaIO<double> io;
  io.outU = nullptr; io.n = ny; io.ncols = ncp; io.ncb = kb; io.pitch = ncp; io.scale = 1.0 / ny;
  io.slab = 2; io.nouter = nxl; io.nouter_full = nxl; io.nyl = nyl; io.peer_tab = tab.data(); io.peer_x0 = x0;
  io.kzb_major = kzb_major; io.nx = nxtot; io.rank = me; io.nranks = P; io.flag_wait = cnt1.data(); io.flag_expect = 7; io.flag_tab = ftab.data();
  SpectralUpdate2<double> up{};
  up.kx = kx.data(); up.ky = ky.data(); up.kz = kz.data(); up.kmode = MRL_KMODE_3D_SLAB; up.nzc = ncp; up.nzv = ncp; up.x0 = x0;
  up.closed_M = 1; up.closed_L = 1; up.has_L = 1; up.Mfac = 0.2; up.Lfac = -0.001; up.dt = 0.01;
  up.b0 = 1.5 * up.dt; up.nold = nold; up.bold0 = -0.5 * up.dt; up.Nout = Ns.data();
  auto mk5 = [&](const void *base) {
    TensorMap m;
    m.base = (const unsigned char *)base; m.esize = 8;
    m.dim[0] = 2 * TK; m.dim[1] = nxl; m.dim[2] = nyl; m.dim[3] = kb; m.dim[4] = P;
    m.stride[0] = 8; m.stride[1] = 16LL * TK; m.stride[2] = 16LL * TK * nxl; m.stride[3] = 16LL * TK * nxl * nyl; m.stride[4] = 16LL * TK * nxl * nyl * kb;
    m.box[0] = 2 * TK; m.box[1] = 1; m.box[2] = nyl; m.box[3] = 1; m.box[4] = P;
    return m;
  };
  TensorMap tmO;
  tmO.base = (const unsigned char *)Os.data(); tmO.esize = 8;
  tmO.dim[0] = 2LL * ncp; tmO.dim[1] = nyl; tmO.dim[2] = nxl; tmO.dim[3] = P; tmO.dim[4] = 1;
  tmO.stride[0] = 8; tmO.stride[1] = 16LL * ncp; tmO.stride[2] = 16LL * ncp * nyl; tmO.stride[3] = 16LL * ncp * nyl * nxl; tmO.stride[4] = 0;
  tmO.box[0] = 2 * TK; tmO.box[1] = nyl; tmO.box[2] = 1; tmO.box[3] = P; tmO.box[4] = 1;
  TensorMap tmC = mk5(Cb.data()), tmG = mk5(Gb.data());
  const cx<double> *twp = tw.data();
  size_t smem = (size_t)(NG * 3 * C::N * TK) * 16 + NG * 3 * 8 + 128;
  emu::launch(dim3(grid), dim3(NG * TK * C::TP), smem, [=] { k_fused_tma<double, C, TK, NG, 2>(tmC, tmG, tmO, io, up, twp); }, 64 * 1024);
  double err = 0, errN = 0;
  for (int x = 0; x < nxl; ++x)
    for (int q = 0; q < ncp; ++q) {
      std::vector<lc> xc(ny), xg(ny);
      for (int y = 0; y < ny; ++y) {
        auto a = Cl[((size_t)x * ny + y) * ncp + q], b = Gl[((size_t)x * ny + y) * ncp + q];
        xc[y] = lc(a.x, a.y); xg[y] = lc(b.x, b.y);
      }
      auto yc = dft(xc, -1), yg = dft(xg, -1);
      std::vector<lc> u(ny);
      for (int y = 0; y < ny; ++y) {
        long double a = kx[x0 + x], b = ky[y], d = kz[q];
        long double kk = a * a + b * b + d * d;
        lc N = (-kk * 0.2L) * yg[y];
        auto no = Ol[((size_t)x * ny + y) * ncp + q];
        u[y] = (yc[y] + (long double)up.b0 * N + (nold ? (long double)up.bold0 * lc(no.x, no.y) : lc(0, 0))) / (1.0L - (long double)up.dt * (kk * kk * -0.001L));
        auto nn = Ns[sidx(x, y, q)];
        errN = std::max(errN, (double)std::abs(lc(nn.x, nn.y) - N));
      }
      auto r = dft(u, +1);
      for (int y = 0; y < ny; ++y) {
        auto v = Sd[y / nyl][(((size_t)(q / TK) * nxtot + x0 + x) * nyl + (y % nyl)) * TK + q % TK];
        err = std::max(err, (double)std::abs(lc(v.x, v.y) - r[y] / (long double)ny));
      }
    }
  char nm[160];
  snprintf(nm, sizeof nm, "%s u", name);
  report(nm, err, 1e-12 * ny);
  snprintf(nm, sizeof nm, "%s N", name);
  report(nm, errN, 1e-12 * ny);
  bool ok = true;
  for (int d = 0; d < P; ++d)
    for (int q = 0; q < P; ++q)
      for (int k = 0; k < kb; ++k) ok = ok && cnt2[d][(size_t)q * kb + k] == (q == me ? (unsigned long long)nxl : 0ull);
  snprintf(nm, sizeof nm, "%s arrival counters", name);
  report(nm, ok ? 0.0 : 1.0, 0.5);
}

static void tma_tests() {
  test_fused_padded_buffers<FFTCfg<64, 8, 8, 8>, 8, 2>("fused tma 64 3D padded pitch, M/L from buffers", MRL_KMODE_3D, 3, 5, 8, 2);
  test_fused_padded_buffers<FFTCfg<64, 8, 8, 8>, 8, 1>("fused tma 64 2D padded pitch, M/L from buffers", MRL_KMODE_2D, 1, 33, 40, 1);
  test_mech_fused<FFTCfg<64, 8, 8, 8>, 8, 2>("mech fused tma 64 TK8 NG2 ny3 nzv5 ncp8 g2", 3, 5, 8, 2);
  test_mech_fused<FFTCfg<64, 8, 8, 8>, 4, 1>("mech fused tma 64 TK4 NG1 ny2 nzv3 ncp3 g1", 2, 3, 3, 1);
  test_mech_fused<FFTCfg<512, 64, 8, 8, 8>, 4, 1>("mech fused tma 512 TK4 NG1 (2 boxes) g3", 1, 3, 4, 3);
  test_strided("strided tma 64 TK8 NG2 NS3", 64, 19, 3, 0, [](auto io, auto tw) { run_strided_tma<FFTCfg<64, 8, 8, 8>, 8, 2, 3>(io, tw, 2); });
  test_strided("strided tma 64 TK8 NG2 NS3 inv g5", 64, 19, 3, 1, [](auto io, auto tw) { run_strided_tma<FFTCfg<64, 8, 8, 8>, 8, 2, 3>(io, tw, 5); });
  test_strided("strided tma 64 TK4 NG4 NS6", 64, 9, 2, 0, [](auto io, auto tw) { run_strided_tma<FFTCfg<64, 8, 8, 8>, 4, 4, 6>(io, tw, 1); });
  test_strided("strided tma 128 TK8 NG1 NS3", 128, 8, 2, 0, [](auto io, auto tw) { run_strided_tma<FFTCfg<128, 16, 8, 4, 4>, 8, 1, 3>(io, tw, 3); });
  test_strided("strided tma 512 TK4 NG2 NS3 (2 boxes)", 512, 5, 1, 1, [](auto io, auto tw) { run_strided_tma<FFTCfg<512, 64, 8, 8, 8>, 4, 2, 3>(io, tw, 1); });
  test_fused("fused tma 64 3D NG2", 64, 3, 5, MRL_KMODE_3D, [](auto io, auto up, auto tw) { run_fused_tma<FFTCfg<64, 8, 8, 8>, 8, 2>(io, up, tw, 1); });
  test_fused("fused tma 64 2D NG1 g2", 64, 21, 1, MRL_KMODE_2D, [](auto io, auto up, auto tw) { run_fused_tma<FFTCfg<64, 8, 8, 8>, 8, 1>(io, up, tw, 2); });
  test_fused("fused tma 64 3D TK4 NG2 nold=0", 64, 3, 5, MRL_KMODE_3D, [](auto io, auto up, auto tw) {
    up.nold = 0;
    run_fused_tma<FFTCfg<64, 8, 8, 8>, 4, 2>(io, up, tw, 1);
  }, 0);
  test_fused_slab<FFTCfg<64, 8, 8, 8>, 8, 2>("fused tma slab 64 P4 nxl3 nzc5", 4, 3, 5, 2, 2);
  test_fused_slab<FFTCfg<64, 8, 8, 8>, 8, 1>("fused tma slab 64 P2 nxl2 nzc9", 2, 2, 9, 0, 1);
  test_fused_slab<FFTCfg<64, 8, 8, 8>, 8, 2>("fused tma slab 64 P4 peer stores", 4, 3, 5, 6, 2, true);
  test_fused_slab<FFTCfg<64, 8, 8, 8>, 8, 2>("fused tma slab 64 P4 nxl6 in 3 x-chunks", 4, 6, 5, 2, 2, false, 3);
  test_strided_peer<FFTCfg<64, 8, 8, 8>, 8, 2, 3>("strided tma 64 peer scatter P4", 4, 11);
  test_slab_xfwd<FFTCfg<64, 8, 8, 8>, 8, 2, 3>("slab xfwd bulk 64 P4 nyl3 kb2 y-major", 4, 3, 2, 0, 1, 2);
  test_slab_xfwd<FFTCfg<64, 8, 8, 8>, 8, 1, 2>("slab xfwd bulk 64 P2 nyl4 kb3 kzb-major 2 chunks", 2, 4, 3, 1, 2, 3);
  test_slab_xfwd<FFTCfg<512, 64, 8, 8, 8>, 4, 1, 2>("slab xfwd bulk 512 P8 nyl1 kb1 (2 boxes)", 8, 1, 1, 1, 1, 1);
  test_slab_xinv<FFTCfg<64, 8, 8, 8>, 8, 2, 3>("slab xinv 64 nyl3 kb2 kzb-major", 3, 2, 1, 2);
  test_slab_xinv<FFTCfg<64, 8, 8, 8>, 4, 1, 3>("slab xinv 64 nyl2 kb3 y-major", 2, 3, 0, 1);
  test_fused_slab2<FFTCfg<64, 8, 8, 8>, 8, 2>("fused tma slab2 64 P4 nxl3 kb2 kzb-major", 4, 3, 2, 5, 1, 2);
  test_fused_slab2<FFTCfg<64, 8, 8, 8>, 8, 1>("fused tma slab2 64 P2 nxl2 kb1 x-major nold=0", 2, 2, 1, 0, 0, 1, 0);
  test_zfwd_tma<FFTCfg<64, 8, 8, 8>, 4, 3, 2>("zfwd tma 64 PPB4 NG3 NS2 rows=21", 21, 2);
  test_zfwd_tma<FFTCfg<64, 8, 8, 8>, 4, 2, 2>("zfwd tma 64 PPB4 y-chunks of 4 in a [5][12] slab", 60, 2, 12, 4);
  test_zfwd_tma<FFTCfg<64, 8, 8, 8>, 4, 1, 2>("zfwd tma 64 PPB4 y-chunks of 8 (last 4) in a [3][12] slab", 36, 1, 12, 8);
  test_zfwd_tma<FFTCfg<64, 8, 8, 8>, 8, 2, 3>("zfwd tma 64 PPB8 NG2 NS3 rows=50", 50, 1);
  test_zfwd_tma<FFTCfg<512, 64, 8, 8, 8>, 1, 2, 2>("zfwd tma 512 PPB1 NG2 NS2 rows=5 (pair_map)", 5, 1);
  test_zfwd_tma<FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 1, 2>("zfwd tma 1024 PPB1 NG1 NS2 rows=2 (pair_map)", 2, 1);
  test_zfwd_tma<FFTCfg<128, 16, 8, 4, 4>, 4, 2, 2>("zfwd tma 128 PPB4 NG2 NS2 rows=9", 9, 2);
  test_zinv_tma<FFTCfg<64, 8, 8, 8>, 2, 2, 2>("zinv tma 64 PPB2 NG2 NS2 rows=7", 7, 1);
  test_zinv_tma<FFTCfg<64, 8, 8, 8>, 1, 4, 3>("zinv tma 64 PPB1 NG4 NS3 rows=40", 40, 2);
  test_zinv_tma<FFTCfg<512, 64, 8, 8, 8>, 2, 2, 2>("zinv tma 512 PPB2 NG2 NS2 rows=9", 9, 1);
  test_mech_tangent<FFTCfg<256, 32, 8, 8, 4>, 2>("mech tangent + zfwd 256 NG2 rows=10 update", 10, 2, true);
  test_mech_tangent<FFTCfg<256, 32, 8, 8, 4>, 2>("mech tangent + zfwd 256 NG2 rows=6", 6, 1, false);
  test_mech_tangent<FFTCfg<512, 64, 8, 8, 8>, 1>("mech tangent + zfwd 512 NG1 rows=4 update", 4, 1, true);
  test_mech_tangent<FFTCfg<256, 32, 8, 8, 4>, 1>("mech tangent + zfwd 256 bulk-staged rows=14 update g2", 14, 2, true, true);
  test_mech_tangent<FFTCfg<256, 32, 8, 8, 4>, 1>("mech tangent + zfwd 256 bulk-staged rows=6 g1", 6, 1, false, true);
}

int main() {
  tma_tests();
  // butterflies via single-stage configs and multi-stage register FFTs
  test_strided("strided fast 8 (R8)", 8, 11, 2, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<8, 1, 8>, 8>(io, tw); });
  test_strided("strided fast 64 (8,8)", 64, 9, 2, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<64, 8, 8, 8>, 8>(io, tw); });
  test_strided("strided fast 32 (8,4) inv", 32, 17, 1, 1, [](auto io, auto tw) { run_strided_fast<FFTCfg<32, 4, 8, 4>, 8>(io, tw); });
  test_strided("strided fast 128 (8,4,4)", 128, 8, 1, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<128, 16, 8, 4, 4>, 8>(io, tw); });
  test_strided("strided fast 512 (8,8,8)", 512, 5, 1, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<512, 64, 8, 8, 8>, 8>(io, tw); });
  test_strided("strided fast 256 (8,8,4) inv", 256, 3, 1, 1, [](auto io, auto tw) { run_strided_fast<FFTCfg<256, 32, 8, 8, 4>, 8>(io, tw); });
  test_strided("strided fast 1024 (8,8,4,4)", 1024, 2, 1, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<1024, 128, 8, 8, 4, 4>, 8>(io, tw); });
  test_strided("strided fast 16 (4,4)", 16, 8, 3, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<16, 4, 4, 4>, 8>(io, tw); });
  test_strided("strided fast 200 (8,5,5)", 200, 4, 1, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<200, 5, 8, 5, 5>, 8>(io, tw); });
  test_strided("strided fast 24 (3,8)", 24, 4, 1, 0, [](auto io, auto tw) { run_strided_fast<FFTCfg<24, 1, 3, 8>, 8>(io, tw); });
  for (int n : {1, 2, 3, 5, 7, 9, 10, 11, 12, 13, 20, 30, 49, 100, 150, 200})
    for (int inv = 0; inv < 2; ++inv) {
      char nm[64];
      snprintf(nm, sizeof nm, "strided gen n=%d inv=%d", n, inv);
      FFTPlanDev plan = make_plan(n);
      test_strided(nm, n, 11, 2, inv, [plan](auto io, auto tw) { run_strided_gen(io, tw, plan); });
    }
  test_real("real fast 512 rows=5", 512, 5, RealFast<FFTCfg<512, 64, 8, 8, 8>, 4>::f, RealFast<FFTCfg<512, 64, 8, 8, 8>, 4>::i);
  test_real("real fast 64 rows=8", 64, 8, RealFast<FFTCfg<64, 8, 8, 8>, 4>::f, RealFast<FFTCfg<64, 8, 8, 8>, 4>::i);
  test_real("real fast 16 rows=3", 16, 3, RealFast<FFTCfg<16, 4, 4, 4>, 2>::f, RealFast<FFTCfg<16, 4, 4, 4>, 2>::i);
  for (int n : {1, 2, 3, 8, 9, 10, 11, 12, 13, 20, 150})
    for (int rows : {1, 6, 7}) {
      char nm[64];
      snprintf(nm, sizeof nm, "real gen n=%d rows=%d", n, rows);
      test_real(nm, n, rows, [n](auto a, auto b, auto c, auto d) { real_gen_f(n, a, b, c, d); },
                [n](auto a, auto b, auto c, auto d) { real_gen_i(n, a, b, c, d); });
    }
  test_fused("fused fast 64 3D", 64, 3, 5, MRL_KMODE_3D, [](auto io, auto up, auto tw) {
    typedef FFTCfg<64, 8, 8, 8> C;
    emu::launch(dim3(2), dim3(8 * C::TP), (size_t)(C::N * 8 + C::N) * sizeof(cx<double>),
                [=] { k_fused_fast<double, C, 8>(io, up, tw); });
  });
  test_fused("fused gen 20 2D", 20, 11, 1, MRL_KMODE_2D, [](auto io, auto up, auto tw) {
    FFTPlanDev plan = make_plan(20);
    emu::launch(dim3(2), dim3(256), (size_t)(3 * 20 * 8) * sizeof(cx<double>),
                [=] { k_fused_gen<double, 8>(io, up, tw, plan); });
  });
  test_fused("fused gen 15 3D", 15, 4, 3, MRL_KMODE_3D, [](auto io, auto up, auto tw) {
    FFTPlanDev plan = make_plan(15);
    emu::launch(dim3(2), dim3(256), (size_t)(3 * 15 * 8) * sizeof(cx<double>),
                [=] { k_fused_gen<double, 8>(io, up, tw, plan); });
  });
  printf("%s (%d failures)\n", g_fail ? "EMU TESTS FAILED" : "EMU TESTS PASSED", g_fail);
  return g_fail ? 1 : 0;
}
