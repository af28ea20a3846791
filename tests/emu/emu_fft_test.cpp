// Host-emulated run of the product's FFT kernels against a naive long-double DFT.
// TEST TOOL: compiled with g++ -DMRL_EMU; validates index math / barriers without a GPU.
#define MRL_EMU 1
#include "../../marlin_b200/csrc/mrl_passes.cuh"

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
template <class RUN> static void test_fused(const char *name, int n, int ny, int nzc, int kmode, RUN run) {
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
  up.b0 = 1.5 * up.dt; up.nold = 1; up.bold[0] = -0.5 * up.dt; up.Nold[0] = Nold0.data();
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
      u[j] = (yc[j] + (long double)up.b0 * N + (long double)up.bold[0] * lc(no.x, no.y)) /
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

int main() {
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
