// Expression-specialised first pass of the fused split plan: the last-axis r2c of
// (c + i F(c, other fields...)) with the user's ParsedCompute expression F compiled INTO the pass.
// The pass kernels are the same templates the built-in double-well path uses (k_zfwd_tma /
// k_zfwd_fast / k_zfwd_gen); NVRTC instantiates them for sm_100a with a functor generated from
// the expression AST, so an arbitrary free-energy derivative costs no extra HBM round trip
// (the reference runs it as separate ATen kernels: src/tensor_computes/ParsedCompute.C:184-265).
#include <cuda_runtime.h>

#include <sstream>
#include <utility>

#include "mrl_expr_internal.h"
#include "mrl_internal.h"
#include "mrl_embedded_headers.inc"

using namespace mrl;

namespace {
template <class T> struct HostF {
  const void *in[16];
  T t;
};
template <class T> struct HostLoadFused {
  const void *c;
  void *mu_out;
  int n;
  HostF<T> f;
};
struct HostStoreTwo {
  void *outA, *outB;
  int nc;
};

struct TmaCfg {
  int n, tp, r0, r1, r2, r3, ppb, ng, ns;
};
// same table as launch_zfwd_nonlin_tma (k_tma.cu)
bool tma_cfg(int n, bool f64, TmaCfg &c) {
  switch (n) {
    case 128: c = {128, 16, 8, 4, 4, 1, 8, 4, 4}; return true;
    case 256: c = {256, 32, 8, 8, 4, 1, 4, 4, 4}; return true;
    case 512: c = f64 ? TmaCfg{512, 64, 8, 8, 8, 1, 2, 5, 3} : TmaCfg{512, 64, 8, 8, 8, 1, 2, 4, 4}; return true;
    case 1024: c = {1024, 128, 8, 8, 4, 4, 1, 4, 3}; return true;
    default: return false;
  }
}
bool fast_cfg(int n, TmaCfg &c) {
  switch (n) {
#define X(N, TP, R0, R1, R2, R3) \
  case N: c = {N, TP, R0, R1, R2, R3, (256 / TP) < 1 ? 1 : (256 / TP), 0, 0}; return true;
    MRL_FAST_SIZES(X)
#undef X
    default: return false;
  }
}
std::string cfg_name(const TmaCfg &c) {
  std::ostringstream s;
  s << "mrl::FFTCfg<" << c.n << ", " << c.tp << ", " << c.r0 << ", " << c.r1 << ", " << c.r2 << ", " << c.r3 << ">";
  return s.str();
}
enum { ZK_TMA = 1, ZK_FAST = 2, ZK_GEN = 3 };
}  // namespace

// Source of the functor: F(v_staged, p) with the other inputs read from global memory at p.
static int functor_source(const mrl_expr *e, int precision, int staged_var, std::string &out) {
  const mrlx::ExprProgram &pr = e->pr;
  if (e->result_type == 2) return mrl_fail(MRL_ERR_UNSUPPORTED, "fused nonlinearity must be real valued");
  if (e->space == 2) return mrl_fail(MRL_ERR_UNSUPPORTED, "fused nonlinearity must live in real space");
  std::string pre, res;
  int rt;
  std::set<std::string> used;
  try {
    mrlx::generate_body(pr, pre, res, rt, used);
  } catch (const std::exception &ex) {
    return mrl_fail(MRL_ERR_PARSE, "%s", ex.what());
  }
  std::ostringstream s;
  s << "namespace ux {\ntypedef " << (precision == MRL_F64 ? "double" : "float") << " T;\n";
  s << mrlx_expr_prelude();
  s << "struct F {\n  const void *in[16];\n  T t;\n  __device__ __forceinline__ T operator()(T staged, long long p) const {\n";
  for (size_t v = 0; v < pr.vars.size(); ++v) {
    if (!used.count(pr.vars[v])) continue;
    const int lay = pr.layouts[v];
    if (lay != MRL_VAR_REAL && lay != MRL_VAR_SCALAR)
      return mrl_fail(MRL_ERR_UNSUPPORTED, "fused nonlinearity: input '%s' is not a real-space real field", pr.vars[v].c_str());
    if ((int)v == staged_var)
      s << "      const T v_" << v << " = staged;\n";
    else
      s << "      const T v_" << v << " = ((const T *)in[" << v << "])[" << (lay == MRL_VAR_SCALAR ? "0" : "p") << "];\n";
  }
  for (const char *sym : {"x", "y", "z", "kx", "ky", "kz", "k2"})
    if (used.count(sym)) return mrl_fail(MRL_ERR_UNSUPPORTED, "fused nonlinearity: coordinate symbol '%s' is not supported", sym);
  if (used.count("t")) s << "      const T e_t = t;\n";
  s << pre << "      return " << (rt == 0 ? "T(" + res + ")" : res) << ";\n  }\n};\n}  // namespace ux\n";
  out = s.str();
  return MRL_OK;
}

static int build_zfwd(mrl_expr *e, int n, int staged_var, bool want_tma, bool padded) {
  mrl_context *ctx = e->ctx;
  const bool f64 = ctx->precision == MRL_F64;
  const std::string T = f64 ? "double" : "float";
  TmaCfg c{};
  int kind;
  std::string name;
  std::ostringstream nm;
  if (want_tma && tma_cfg(n, f64, c)) {
    kind = ZK_TMA;
    nm << "mrl::k_zfwd_tma<" << T << ", " << cfg_name(c) << ", " << c.ppb << ", " << c.ng << ", " << c.ns << ", ux::F>";
    const int NP = c.n + (c.n >> 3) + 1;
    const size_t esz = f64 ? 8 : 4;
    e->zfwd_smem = (unsigned)((size_t)c.ng * c.ns * c.ppb * c.n * esz + (size_t)(c.ng * c.ppb * NP) * 2 * esz + c.ng * c.ns * 8 + 128);
    e->zfwd_block = c.ng * c.ppb * c.tp;
    e->zfwd_ppb = c.ppb * c.ng;
  } else if (padded) {
    return mrl_fail(MRL_ERR_UNSUPPORTED, "padded work spectra need the TMA first pass");
  } else if (fast_cfg(n, c)) {
    kind = ZK_FAST;
    nm << "mrl::k_zfwd_fast<" << T << ", " << cfg_name(c) << ", " << c.ppb << ", mrl::ZLoadFused<" << T << ", ux::F>, mrl::ZStoreTwo<" << T << ">>";
    const int NP = c.n + (c.n >> 3) + 1;
    e->zfwd_smem = (unsigned)((size_t)(NP * c.ppb + c.n) * 2 * (f64 ? 8 : 4));
    e->zfwd_block = c.ppb * c.tp;
    e->zfwd_ppb = c.ppb;
  } else {
    kind = ZK_GEN;
    const int tk = f64 ? gen_tk<double>(n, 2) : gen_tk<float>(n, 2);
    if (tk < 1) return mrl_fail(MRL_ERR_UNSUPPORTED, "axis of %d points does not fit in shared memory", n);
    nm << "mrl::k_zfwd_gen<" << T << ", " << tk << ", mrl::ZLoadFused<" << T << ", ux::F>, mrl::ZStoreTwo<" << T << ">>";
    e->zfwd_smem = (unsigned)((size_t)2 * n * tk * 2 * (f64 ? 8 : 4));
    e->zfwd_block = 256;
    e->zfwd_ppb = tk;
  }
  std::string fsrc;
  int rc = functor_source(e, ctx->precision, staged_var, fsrc);
  if (rc) return rc;
  std::string src = "#include \"mrl_passes_tma.cuh\"\n" + fsrc;
  std::vector<std::pair<std::string, std::string>> headers;
  for (const auto &h : kEmbeddedHeaders) headers.push_back({h.first, h.second});
  std::vector<char> cubin;
  std::vector<std::string> low;
  std::string log;
  rc = mrlx_nvrtc_compile(src, headers, {nm.str()}, cubin, low, log);
  if (rc) return rc;
  if (low.empty() || low[0].empty()) return mrl_fail(MRL_ERR_CUDA, "NVRTC did not return a lowered name for %s", nm.str().c_str());
  mrlx_module_unload(e->zfwd_module);
  e->zfwd_module = nullptr;
  if ((rc = mrlx_module_load(cubin, low[0], &e->zfwd_module, &e->zfwd_fn))) return rc;
  e->zfwd_n = n;
  e->zfwd_kind = kind;
  e->zfwd_staged = staged_var;
  e->zfwd_source = src;
  return MRL_OK;
}

template <class T>
static int launch_typed(mrl_context *ctx, mrl_expr *e, const void *const *inputs, double t, const void *c, void *g_out, void *outC,
                        void *outG, long long rows, int n, int ncp, const int *rowmap, void *stream_or_null, int sm_count) {
  cudaStream_t stream = stream_or_null ? (cudaStream_t)stream_or_null : ctx->stream;
  if (sm_count <= 0) sm_count = ctx->sm_count;
  HostF<T> f;
  memset(&f, 0, sizeof f);
  for (size_t v = 0; v < e->pr.vars.size() && v < 16; ++v) f.in[v] = inputs ? inputs[v] : nullptr;
  f.t = (T)t;
  const void *tw;
  int rc = ctx->twiddles(n, &tw);
  if (rc) return rc;
  ctx->launches++;
  if (e->zfwd_kind == ZK_TMA) {
    long long nrows = rows;
    const long long nwork = (rows + e->zfwd_ppb - 1) / e->zfwd_ppb;
    const unsigned grid = (unsigned)(nwork < sm_count ? nwork : sm_count);
    mrl::RowMap rm{0, 0, 0};
    if (rowmap) rm = mrl::RowMap{rowmap[0], rowmap[1], rowmap[2]};
    void *params[] = {(void *)&c, &g_out, &outC, &outG, &nrows, &ncp, &f, (void *)&tw, &rm};
    return mrlx_launch(e->zfwd_fn, grid, e->zfwd_block, e->zfwd_smem, stream, params);
  }
  if (rowmap && rowmap[0]) return mrl_fail(MRL_ERR_UNSUPPORTED, "row-mapped first pass needs the TMA kernel");
  if (ncp != n / 2 + 1) return mrl_fail(MRL_ERR_UNSUPPORTED, "padded work spectra need the TMA first pass");
  HostLoadFused<T> ld;
  memset(&ld, 0, sizeof ld);
  ld.c = c;
  ld.mu_out = g_out;
  ld.n = n;
  ld.f = f;
  HostStoreTwo st{outC, outG, n / 2 + 1};
  long long npencils = rows;
  const long long nblk = (rows + e->zfwd_ppb - 1) / e->zfwd_ppb;
  const long long cap = (long long)ctx->sm_count * 4;
  const unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
  if (e->zfwd_kind == ZK_FAST) {
    void *params[] = {&ld, &st, (void *)&tw, &npencils};
    return mrlx_launch(e->zfwd_fn, grid, e->zfwd_block, e->zfwd_smem, stream, params);
  }
  FFTPlanDev plan = make_fft_plan(n);
  void *params[] = {&ld, &st, (void *)&tw, &plan, &npencils};
  return mrlx_launch(e->zfwd_fn, grid, e->zfwd_block, e->zfwd_smem, stream, params);
}

int mrl_expr_launch_zfwd(mrl_context *ctx, void *expr, int staged_var, const void *const *inputs, double t, const void *c,
                         void *g_out, void *outC, void *outG, long long rows, int n, int ncp) {
  return mrl_expr_launch_zfwd_rows(ctx, expr, staged_var, inputs, t, c, g_out, outC, outG, rows, n, ncp, nullptr, nullptr, 0);
}

int mrl_expr_launch_zfwd_rows(mrl_context *ctx, void *expr, int staged_var, const void *const *inputs, double t, const void *c, void *g_out,
                              void *outC, void *outG, long long rows, int n, int ncp, const int *rowmap, void *stream, int sm_count) {
  mrl_expr *e = (mrl_expr *)expr;
  if (!e || e->ctx != ctx) return mrl_fail(MRL_ERR_INVALID, "expression belongs to another context");
  const bool padded = ncp != n / 2 + 1;
  const size_t esz = ctx->precision == MRL_F64 ? 8 : 4;
  const bool aligned = !((unsigned long long)c & 15ull) && ((size_t)n * esz) % 16 == 0;
  const bool want_tma = tma_enabled() && aligned;
  if (!e->zfwd_fn || e->zfwd_n != n || e->zfwd_staged != staged_var || (e->zfwd_kind == ZK_TMA) != (want_tma && (n == 128 || n == 256 || n == 512 || n == 1024))) {
    int rc = build_zfwd(e, n, staged_var, want_tma, padded);
    if (rc) return rc;
  }
  return ctx->precision == MRL_F64 ? launch_typed<double>(ctx, e, inputs, t, c, g_out, outC, outG, rows, n, ncp, rowmap, stream, sm_count)
                                   : launch_typed<float>(ctx, e, inputs, t, c, g_out, outC, outG, rows, n, ncp, rowmap, stream, sm_count);
}

// Dry run for the CPU test-suite: generate + compile the specialised pass without a device.
extern "C" int mrl_expr_check_fused(const mrl_expr_desc *d, int precision, int n, int staged_var) {
  mrl_context fake;
  fake.precision = precision;
  fake.dim = 3;
  mrl_expr e;
  e.ctx = &fake;
  int rc = mrlx_fill_program(e.pr, d);
  if (rc) return rc;
  std::string pre, res;
  std::set<std::string> used;
  try {
    mrlx::generate_body(e.pr, pre, res, e.result_type, used);
  } catch (const std::exception &ex) {
    return mrl_fail(MRL_ERR_PARSE, "%s", ex.what());
  }
  e.space = 1;
  const bool f64 = precision == MRL_F64;
  const std::string T = f64 ? "double" : "float";
  std::string fsrc;
  if ((rc = functor_source(&e, precision, staged_var, fsrc))) return rc;
  TmaCfg c{};
  std::vector<std::string> names;
  if (tma_cfg(n, f64, c)) {
    std::ostringstream nm;
    nm << "mrl::k_zfwd_tma<" << T << ", " << cfg_name(c) << ", " << c.ppb << ", " << c.ng << ", " << c.ns << ", ux::F>";
    names.push_back(nm.str());
  }
  if (fast_cfg(n, c)) {
    std::ostringstream nm;
    nm << "mrl::k_zfwd_fast<" << T << ", " << cfg_name(c) << ", " << c.ppb << ", mrl::ZLoadFused<" << T << ", ux::F>, mrl::ZStoreTwo<" << T << ">>";
    names.push_back(nm.str());
  } else {
    const int tk = f64 ? gen_tk<double>(n, 2) : gen_tk<float>(n, 2);
    std::ostringstream nm;
    nm << "mrl::k_zfwd_gen<" << T << ", " << tk << ", mrl::ZLoadFused<" << T << ", ux::F>, mrl::ZStoreTwo<" << T << ">>";
    names.push_back(nm.str());
  }
  std::vector<std::pair<std::string, std::string>> headers;
  for (const auto &h : kEmbeddedHeaders) headers.push_back({h.first, h.second});
  std::vector<char> cubin;
  std::vector<std::string> low;
  std::string log;
  rc = mrlx_nvrtc_compile("#include \"mrl_passes_tma.cuh\"\n" + fsrc, headers, names, cubin, low, log);
  e.ctx = nullptr;
  return rc;
}
