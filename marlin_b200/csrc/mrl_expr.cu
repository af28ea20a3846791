// Runtime expressions on the device: ParsedCompute (src/tensor_computes/ParsedCompute.C:50-265) /
// ParsedJITTensor (src/utils/ParsedJITTensor.C:22-156).  The reference lowers the simplified AST to
// a TorchScript graph (one ATen kernel per node, fused by NNC at best); here the AST is lowered to
// ONE CUDA kernel (typed: bool / real / complex values in registers), compiled for sm_100a with
// NVRTC at init() time and launched through the driver entry points of the CUDA runtime.
// No interpreter, no CPU evaluation of fields: without NVRTC + a device this fails loudly.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <sstream>

#include "mrl_expr_ast.h"
#include "mrl_expr_internal.h"
#include "mrl_internal.h"

using namespace mrlx;

// ------------------------------------------------------------------------------ NVRTC (dlopen)
namespace {
struct Nvrtc {
  void *h = nullptr;
  int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(void *, int, const char *const *) = nullptr;
  int (*GetProgramLogSize)(void *, size_t *) = nullptr;
  int (*GetProgramLog)(void *, char *) = nullptr;
  int (*GetCUBINSize)(void *, size_t *) = nullptr;
  int (*GetCUBIN)(void *, char *) = nullptr;
  int (*DestroyProgram)(void **) = nullptr;
  int (*AddNameExpression)(void *, const char *) = nullptr;
  int (*GetLoweredName)(void *, const char *, const char **) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string where;
};
Nvrtc *nvrtc() {
  static Nvrtc lib;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> cand;
    if (const char *e = getenv("MRL_NVRTC_PATH")) cand.push_back(e);
    cand.insert(cand.end(), {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"});
    for (const std::string &c : cand) {
      lib.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (lib.h) {
        lib.where = c;
        break;
      }
    }
    if (!lib.h) return;
#define SYM(n) *(void **)(&lib.n) = dlsym(lib.h, "nvrtc" #n)
    SYM(CreateProgram);
    SYM(CompileProgram);
    SYM(GetProgramLogSize);
    SYM(GetProgramLog);
    SYM(GetCUBINSize);
    SYM(GetCUBIN);
    SYM(DestroyProgram);
    SYM(AddNameExpression);
    SYM(GetLoweredName);
    SYM(GetErrorString);
#undef SYM
    if (!lib.CreateProgram || !lib.CompileProgram || !lib.GetCUBIN) {
      dlclose(lib.h);
      lib.h = nullptr;
    }
  });
  return lib.h ? &lib : nullptr;
}

struct Driver {
  CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                           void **) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
  bool ok = false;
};
Driver *driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char *name, void **p) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *p;
    };
    d.ok = get("cuModuleLoadData", (void **)&d.ModuleLoadData) && get("cuModuleUnload", (void **)&d.ModuleUnload) &&
           get("cuModuleGetFunction", (void **)&d.ModuleGetFunction) && get("cuLaunchKernel", (void **)&d.LaunchKernel) &&
           get("cuFuncSetAttribute", (void **)&d.FuncSetAttribute) && get("cuGetErrorString", (void **)&d.GetErrorString);
  });
  return d.ok ? &d : nullptr;
}
}  // namespace

int mrlx_nvrtc_compile(const std::string &src, const std::vector<std::pair<std::string, std::string>> &headers,
                       const std::vector<std::string> &name_exprs, std::vector<char> &cubin, std::vector<std::string> &lowered,
                       std::string &log) {
  Nvrtc *rt = nvrtc();
  if (!rt) return mrl_fail(MRL_ERR_UNSUPPORTED, "libnvrtc.so.12 not found (set MRL_NVRTC_PATH); there is no interpreter fallback");
  std::vector<const char *> hsrc, hname;
  for (auto &h : headers) {
    hname.push_back(h.first.c_str());
    hsrc.push_back(h.second.c_str());
  }
  void *prog = nullptr;
  int rc = rt->CreateProgram(&prog, src.c_str(), "mrl_expr.cu", (int)headers.size(), hsrc.data(), hname.data());
  if (rc) return mrl_fail(MRL_ERR_CUDA, "nvrtcCreateProgram failed (%d)", rc);
  for (const std::string &n : name_exprs) rt->AddNameExpression(prog, n.c_str());
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
  rc = rt->CompileProgram(prog, 4, opts);
  size_t ls = 0;
  rt->GetProgramLogSize(prog, &ls);
  log.assign(ls, '\0');
  if (ls) rt->GetProgramLog(prog, &log[0]);
  if (rc) {
    rt->DestroyProgram(&prog);
    return mrl_fail(MRL_ERR_CUDA, "NVRTC compilation failed (%s):\n%.800s", rt->GetErrorString ? rt->GetErrorString(rc) : "?", log.c_str());
  }
  size_t cs = 0;
  rt->GetCUBINSize(prog, &cs);
  cubin.assign(cs, 0);
  rt->GetCUBIN(prog, cubin.data());
  lowered.clear();
  for (const std::string &n : name_exprs) {
    const char *ln = nullptr;
    rt->GetLoweredName(prog, n.c_str(), &ln);
    lowered.push_back(ln ? ln : "");
  }
  rt->DestroyProgram(&prog);
  return MRL_OK;
}

int mrlx_module_load(const std::vector<char> &cubin, const std::string &fn, void **module, void **function) {
  Driver *d = driver();
  if (!d) return mrl_fail(MRL_ERR_CUDA, "CUDA driver entry points unavailable");
  cudaFree(0);  // make sure the primary context exists and is current
  CUmodule m;
  CUresult r = d->ModuleLoadData(&m, cubin.data());
  const char *es = "?";
  if (r != CUDA_SUCCESS) {
    d->GetErrorString(r, &es);
    return mrl_fail(MRL_ERR_CUDA, "cuModuleLoadData failed: %s", es);
  }
  CUfunction f;
  r = d->ModuleGetFunction(&f, m, fn.c_str());
  if (r != CUDA_SUCCESS) {
    d->GetErrorString(r, &es);
    d->ModuleUnload(m);
    return mrl_fail(MRL_ERR_CUDA, "cuModuleGetFunction(%s) failed: %s", fn.c_str(), es);
  }
  *module = m;
  *function = f;
  return MRL_OK;
}
int mrlx_module_get(void *module, const std::string &fn, void **function) {
  Driver *d = driver();
  CUfunction f;
  CUresult r = d->ModuleGetFunction(&f, (CUmodule)module, fn.c_str());
  if (r != CUDA_SUCCESS) return mrl_fail(MRL_ERR_CUDA, "cuModuleGetFunction(%s) failed", fn.c_str());
  *function = f;
  return MRL_OK;
}
void mrlx_module_unload(void *module) {
  if (module && driver()) driver()->ModuleUnload((CUmodule)module);
}
int mrlx_launch(void *function, unsigned grid, unsigned block, unsigned smem, cudaStream_t stream, void **params) {
  Driver *d = driver();
  if (smem > 48 * 1024) {
    CUresult a = d->FuncSetAttribute((CUfunction)function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
    if (a != CUDA_SUCCESS) return mrl_fail(MRL_ERR_CUDA, "cuFuncSetAttribute(max dynamic smem %u) failed", smem);
  }
  CUresult r = d->LaunchKernel((CUfunction)function, grid, 1, 1, block, 1, 1, smem, (CUstream)stream, params, nullptr);
  if (r != CUDA_SUCCESS) {
    const char *es = "?";
    d->GetErrorString(r, &es);
    return mrl_fail(MRL_ERR_CUDA, "cuLaunchKernel failed: %s", es);
  }
  return MRL_OK;
}

// ------------------------------------------------------------------------------ code generation
namespace {
enum Ty { TB = 0, TR = 1, TC = 2 };  // bool, real, complex

std::string lit(double v) {
  char buf[64];
  if (std::isnan(v)) return "mrl_nan()";
  if (std::isinf(v)) return v > 0 ? "mrl_inf()" : "(-mrl_inf())";
  snprintf(buf, sizeof buf, "T(%.17g)", v);
  return buf;
}

struct Gen {
  const mrlx::ExprProgram &pr;
  std::map<std::string, std::pair<std::string, Ty>> scope;  // symbol -> (C identifier, type)
  std::set<std::string> used_inputs;                        // input / extra symbols referenced
  std::ostringstream pre;
  int tmp = 0;

  explicit Gen(const mrlx::ExprProgram &p) : pr(p) {}

  static std::string real(const std::pair<std::string, Ty> &v) { return v.second == TB ? "T(" + v.first + ")" : v.first; }
  static std::string boolean(const std::pair<std::string, Ty> &v) {
    if (v.second == TC) throw std::runtime_error("complex value used as a condition");
    return v.second == TB ? v.first : "(" + v.first + " != T(0))";
  }
  static std::string cplx(const std::pair<std::string, Ty> &v) { return v.second == TC ? v.first : "mrl_cx(" + real(v) + ")"; }

  std::pair<std::string, Ty> symbol(const std::string &n) {
    auto it = scope.find(n);
    if (it != scope.end()) return it->second;
    auto c = pr.constants.find(n);
    if (c != pr.constants.end()) return {lit(c->second), TR};
    if (pr.extra) {
      if (n == "pi") return {lit(3.14159265358979323846), TR};
      if (n == "e") return {lit(2.71828182845904523536), TR};
      if (n == "i") return {"mrl_i()", TC};
      static const char *ex[] = {"x", "y", "z", "kx", "ky", "kz", "k2", "t"};
      for (const char *s : ex)
        if (n == s) {
          used_inputs.insert(n);
          return {std::string("e_") + s, TR};
        }
    }
    for (size_t v = 0; v < pr.vars.size(); ++v)
      if (pr.vars[v] == n) {
        used_inputs.insert(n);
        const int lay = pr.layouts[v];
        return {"v_" + std::to_string(v), (lay == MRL_VAR_RECIP_COMPLEX || lay == MRL_VAR_REAL_COMPLEX) ? TC : TR};
      }
    throw std::runtime_error("Variable '" + n + "' not found in variable list");
  }

  std::pair<std::string, Ty> gen(const P &e) {
    switch (e->k) {
      case Kind::Num: return {lit(e->v), TR};
      case Kind::Var:
      case Kind::Const: return symbol(e->s);
      case Kind::Bin: {
        const std::string &op = e->s;
        if (op == "^") return power(e);
        auto l = gen(e->a[0]), r = gen(e->a[1]);
        if (op == "%") {
          if (l.second == TC || r.second == TC) throw std::runtime_error("'%' is not defined for complex values");
          return {"mrl_rem(" + real(l) + ", " + real(r) + ")", TR};
        }
        if (l.second == TC || r.second == TC) {
          const std::string a = l.second == TC ? l.first : real(l), b = r.second == TC ? r.first : real(r);
          return {"(" + a + " " + op + " " + b + ")", TC};
        }
        if (l.second == TB && r.second == TB) {
          // ATen arithmetic on two bool tensors stays bool: add = or, mul = and, sub is an error
          if (op == "+") return {"(" + l.first + " || " + r.first + ")", TB};
          if (op == "*") return {"(" + l.first + " && " + r.first + ")", TB};
          if (op == "-") throw std::runtime_error("Subtraction, the `-` operator, with two bool tensors is not supported");
        }
        return {"(" + real(l) + " " + op + " " + real(r) + ")", TR};
      }
      case Kind::Un: {
        auto x = gen(e->a[0]);
        if (e->s == "-") return x.second == TC ? std::make_pair("(-" + x.first + ")", TC) : std::make_pair("(-" + real(x) + ")", TR);
        return {"(!" + boolean(x) + ")", TB};
      }
      case Kind::Cmp: {
        auto l = gen(e->a[0]), r = gen(e->a[1]);
        if (l.second == TC || r.second == TC) {
          if (e->s != "==" && e->s != "!=") throw std::runtime_error("ordering comparison of complex values");
          return {"(" + std::string(e->s == "!=" ? "!" : "") + "mrl_eq(" + cplx(l) + ", " + cplx(r) + "))", TB};
        }
        return {"(" + real(l) + " " + e->s + " " + real(r) + ")", TB};
      }
      case Kind::Log: {
        auto l = gen(e->a[0]), r = gen(e->a[1]);
        return {"(" + boolean(l) + (e->s == "&" ? " && " : " || ") + boolean(r) + ")", TB};
      }
      case Kind::Call: return function(e);
      case Kind::Let: {
        auto saved = scope;
        for (size_t i = 0; i < e->names.size(); ++i) {
          auto v = gen(e->a[i]);
          const std::string id = "l" + std::to_string(tmp++) + "_" + e->names[i];
          pre << "      const " << (v.second == TC ? "cx" : v.second == TB ? "bool" : "T") << " " << id << " = " << v.first << ";\n";
          scope[e->names[i]] = {id, v.second};
        }
        auto body = gen(e->a.back());
        // materialise the body before the bindings go out of scope (identifiers stay valid in C)
        scope = saved;
        return body;
      }
    }
    throw std::runtime_error("internal: unknown node");
  }

  std::pair<std::string, Ty> power(const P &e) {
    auto l = gen(e->a[0]);
    const P &re = e->a[1];
    if (re->k == Kind::Num && std::floor(re->v) == re->v && std::fabs(re->v) <= 64) {
      const int n = (int)re->v;
      const std::string base = l.second == TC ? l.first : real(l);
      return {"mrl_ipow<" + std::to_string(std::abs(n)) + ", " + (n < 0 ? "true" : "false") + ">(" + base + ")", l.second == TC ? TC : TR};
    }
    auto r = gen(re);
    if (l.second == TC || r.second == TC) throw std::runtime_error("complex power with a non-integer exponent is not supported");
    if (re->k == Kind::Num && re->v == 0.5) return {"sqrt(" + real(l) + ")", TR};
    return {"pow(" + real(l) + ", " + real(r) + ")", TR};
  }

  std::pair<std::string, Ty> function(const P &e) {
    const std::string &f = e->s;
    std::vector<std::pair<std::string, Ty>> v;
    for (const P &x : e->a) v.push_back(gen(x));
    static const std::set<std::string> one = {"sin",  "cos",   "tan",  "sinh", "cosh", "tanh",  "asin", "acos", "atan", "asinh", "acosh", "atanh",
                                              "exp",  "exp2",  "log",  "log10", "log2", "sqrt",  "rsqrt", "abs", "ceil", "floor", "round", "trunc"};
    if (v.size() == 1 && one.count(f)) {
      if (v[0].second == TC) {
        if (f == "exp") return {"mrl_cexp(" + v[0].first + ")", TC};
        if (f == "abs") return {"mrl_cabs(" + v[0].first + ")", TR};
        throw std::runtime_error("function '" + f + "' is not supported for complex arguments");
      }
      const std::string a = real(v[0]);
      if (f == "abs") return {"fabs(" + a + ")", TR};
      if (f == "round") return {"rint(" + a + ")", TR};  // aten::round: half to even
      if (f == "rsqrt") return {"(T(1) / sqrt(" + a + "))", TR};
      return {f + "(" + a + ")", TR};
    }
    if (v.size() == 2) {
      if (v[0].second == TC || v[1].second == TC) throw std::runtime_error("function '" + f + "' is not supported for complex arguments");
      const std::string a = real(v[0]), b = real(v[1]);
      if (f == "min") return {"mrl_min(" + a + ", " + b + ")", TR};
      if (f == "max") return {"mrl_max(" + a + ", " + b + ")", TR};
      if (f == "atan2" || f == "hypot" || f == "pow") return {f + "(" + a + ", " + b + ")", TR};
    }
    if (f == "if" && v.size() == 3) {
      if (v[1].second == TC || v[2].second == TC) return {"(" + boolean(v[0]) + " ? " + cplx(v[1]) + " : " + cplx(v[2]) + ")", TC};
      if (v[1].second == TB && v[2].second == TB) return {"(" + boolean(v[0]) + " ? " + v[1].first + " : " + v[2].first + ")", TB};
      return {"(" + boolean(v[0]) + " ? " + real(v[1]) + " : " + real(v[2]) + ")", TR};
    }
    throw std::runtime_error("Unknown or unsupported function: " + f);
  }
};

const char *kPrelude = R"SRC(
struct cx { T x, y; };
__device__ __forceinline__ cx mrl_cx(T a) { cx r; r.x = a; r.y = T(0); return r; }
__device__ __forceinline__ cx mrl_i() { cx r; r.x = T(0); r.y = T(1); return r; }
__device__ __forceinline__ cx operator+(cx a, cx b) { cx r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
__device__ __forceinline__ cx operator+(cx a, T b) { cx r; r.x = a.x + b; r.y = a.y; return r; }
__device__ __forceinline__ cx operator+(T a, cx b) { cx r; r.x = a + b.x; r.y = b.y; return r; }
__device__ __forceinline__ cx operator-(cx a, cx b) { cx r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
__device__ __forceinline__ cx operator-(cx a, T b) { cx r; r.x = a.x - b; r.y = a.y; return r; }
__device__ __forceinline__ cx operator-(T a, cx b) { cx r; r.x = a - b.x; r.y = -b.y; return r; }
__device__ __forceinline__ cx operator-(cx a) { cx r; r.x = -a.x; r.y = -a.y; return r; }
__device__ __forceinline__ cx operator*(cx a, cx b) { cx r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
__device__ __forceinline__ cx operator*(cx a, T b) { cx r; r.x = a.x * b; r.y = a.y * b; return r; }
__device__ __forceinline__ cx operator*(T a, cx b) { cx r; r.x = a * b.x; r.y = a * b.y; return r; }
__device__ __forceinline__ cx operator/(cx a, T b) { cx r; r.x = a.x / b; r.y = a.y / b; return r; }
__device__ __forceinline__ cx operator/(cx a, cx b) {
  // Smith's algorithm, as c10::complex division
  cx r;
  if (fabs(b.x) >= fabs(b.y)) {
    const T q = b.y / b.x, d = b.x + b.y * q;
    r.x = (a.x + a.y * q) / d; r.y = (a.y - a.x * q) / d;
  } else {
    const T q = b.x / b.y, d = b.x * q + b.y;
    r.x = (a.x * q + a.y) / d; r.y = (a.y * q - a.x) / d;
  }
  return r;
}
__device__ __forceinline__ cx operator/(T a, cx b) { return mrl_cx(a) / b; }
__device__ __forceinline__ bool mrl_eq(cx a, cx b) { return a.x == b.x && a.y == b.y; }
__device__ __forceinline__ cx mrl_cexp(cx a) { const T m = exp(a.x); T s, c; sincos(a.y, &s, &c); cx r; r.x = m * c; r.y = m * s; return r; }
__device__ __forceinline__ T mrl_cabs(cx a) { return hypot(a.x, a.y); }
__device__ __forceinline__ T mrl_nan() { return T(__longlong_as_double(0x7ff8000000000000LL)); }
__device__ __forceinline__ T mrl_inf() { return T(__longlong_as_double(0x7ff0000000000000LL)); }
__device__ __forceinline__ T mrl_min(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ T mrl_max(T a, T b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ T mrl_rem(T a, T b) { T r = fmod(a, b); if (r != T(0) && ((r < T(0)) != (b < T(0)))) r += b; return r; }
__device__ __forceinline__ T mrl_one(T) { return T(1); }
__device__ __forceinline__ cx mrl_one(cx) { return mrl_cx(T(1)); }
template <int N, bool INV, class V> __device__ __forceinline__ V mrl_ipow(V x) {
  V r = mrl_one(x);
  V b = x;
#pragma unroll
  for (int n = N; n > 0; n >>= 1) {
    if (n & 1) r = r * b;
    if (n > 1) b = b * b;
  }
  if (INV) return mrl_one(x) / r;
  return r;
}
)SRC";

}  // namespace

const char *mrlx_expr_prelude() { return kPrelude; }

namespace mrlx {

// Builds the AST of a program (parse, derivatives in order, simplify), ParsedCompute.C:126-181.
void build_ast(ExprProgram &pr, const std::string &expression, const std::vector<std::string> &derivatives) {
  std::set<std::string> cn;
  for (auto &c : pr.constants) cn.insert(c.first);
  if (pr.extra) cn.insert({"pi", "e", "i"});
  P ast = parse(expression, cn);
  for (const std::string &d : derivatives) {
    bool listed = false;
    for (const std::string &v : pr.vars) listed = listed || v == d;
    if (!listed) throw std::runtime_error("Derivative w.r.t `" + d + "` was requested, but it is not listed in `inputs`.");
    ast = differentiate(ast, d);
  }
  pr.ast = simplify(ast);
}

// Emits the body of one evaluation: loads of the used inputs + bindings + the result expression.
// `load_var(v)` gives the C expression that reads input v at the current point.
void generate_body(const ExprProgram &pr, std::string &loads_and_bindings, std::string &result, int &result_type,
                   std::set<std::string> &used) {
  Gen g(pr);
  auto r = g.gen(pr.ast);
  loads_and_bindings = g.pre.str();
  result = r.first;
  result_type = r.second;
  used = g.used_inputs;
}

}  // namespace mrlx

// The generic pointwise kernel: out[p] = f(in_0[p], ..., x, y, z, kx, ky, kz, k2, t) over a real
// or reciprocal shaped index space (d0, d1, d2), d2 fastest.  Inputs broadcast through per-input
// strides (full field: (d1*d2, d2, 1); one value: (0, 0, 0)).
static std::string generic_source(const ExprProgram &pr, int precision, int &result_type, int &space, bool &needs_coords) {
  std::string pre, res;
  std::set<std::string> used;
  generate_body(pr, pre, res, result_type, used);
  bool real_space = false, recip_space = false;
  needs_coords = false;
  std::ostringstream ld;
  for (size_t v = 0; v < pr.vars.size(); ++v) {
    if (!used.count(pr.vars[v])) continue;
    const int lay = pr.layouts[v];
    const bool c = lay == MRL_VAR_RECIP_COMPLEX || lay == MRL_VAR_REAL_COMPLEX;
    if (lay == MRL_VAR_REAL || lay == MRL_VAR_REAL_COMPLEX) real_space = true;
    if (lay == MRL_VAR_RECIP_REAL || lay == MRL_VAR_RECIP_COMPLEX) recip_space = true;
    ld << "      const " << (c ? "cx" : "T") << " v_" << v << " = ((const " << (c ? "cx" : "T") << " *)a.in[" << v << "])["
       << (lay == MRL_VAR_SCALAR ? "0" : "p") << "];\n";
  }
  static const char *ax[] = {"x", "y", "z"};
  for (int d = 0; d < 3; ++d) {
    if (used.count(ax[d])) {
      real_space = true;
      needs_coords = true;
      ld << "      const T e_" << ax[d] << " = ((const T *)a.axis[" << d << "])[c" << d << "];\n";
    }
    const std::string k = std::string("k") + ax[d];
    if (used.count(k) || used.count("k2")) {
      recip_space = true;
      needs_coords = true;
      ld << "      const T e_" << k << " = ((const T *)a.kaxis[" << d << "])[c" << d << "];\n";
    }
  }
  if (used.count("k2")) ld << "      const T e_k2 = e_kx * e_kx + e_ky * e_ky + e_kz * e_kz;\n";
  if (used.count("t")) ld << "      const T e_t = T(a.t);\n";
  if (real_space && recip_space) throw std::runtime_error("expression mixes real-space and reciprocal-space operands");
  space = recip_space ? 2 : real_space ? 1 : 0;
  if (pr.expand == MRL_EXPAND_REAL) {
    if (recip_space) throw std::runtime_error("expand = REAL on a reciprocal-space expression");
    space = 1;
  } else if (pr.expand == MRL_EXPAND_RECIPROCAL) {
    if (real_space) throw std::runtime_error("expand = RECIPROCAL on a real-space expression");
    space = 2;
  }
  std::ostringstream s;
  s << "typedef " << (precision == MRL_F64 ? "double" : "float") << " T;\n" << kPrelude;
  s << "struct Args { const void *in[16]; void *out; const void *axis[3]; const void *kaxis[3]; long long total; long long d[3]; "
       "int cmap[3]; double t; };\n";
  s << "template <class I> __device__ __forceinline__ void mrl_body(const Args &a) {\n"
       "  const I total = (I)a.total, d2 = (I)a.d[2], d1 = (I)a.d[1];\n"
       "  const I step = (I)gridDim.x * (I)blockDim.x;\n"
       "  for (I p = (I)blockIdx.x * (I)blockDim.x + (I)threadIdx.x; p < total; p += step) {\n";
  if (needs_coords)
    s << "      const I q = p / d2;\n"
         "      const int g2 = (int)(p - q * d2);\n"
         "      const I q0 = q / d1;\n"
         "      const int g1 = (int)(q - q0 * d1), g0 = (int)q0;\n"
         "      const int gs[4] = {g0, g1, g2, 0};\n"
         "      const int c0 = gs[a.cmap[0]], c1 = gs[a.cmap[1]], c2 = gs[a.cmap[2]];\n"
         "      (void)c0; (void)c1; (void)c2;\n";
  s << ld.str() << pre;
  const char *ot = result_type == TC ? "cx" : "T";
  s << "      ((" << ot << " *)a.out)[p] = " << (result_type == TB ? "T(" + res + ")" : res) << ";\n  }\n}\n";
  s << "extern \"C\" __global__ void __launch_bounds__(256) mrl_expr_u32(const Args a) { mrl_body<unsigned int>(a); }\n";
  s << "extern \"C\" __global__ void __launch_bounds__(256) mrl_expr_u64(const Args a) { mrl_body<unsigned long long>(a); }\n";
  return s.str();
}

struct HostArgs {
  const void *in[16];
  void *out;
  const void *axis[3];
  const void *kaxis[3];
  long long total;
  long long d[3];
  int cmap[3];
  double t;
};

// ------------------------------------------------------------------------------ C ABI
int mrlx_fill_program(ExprProgram &pr, const mrl_expr_desc *d) {
  if (!d || !d->expression) return mrl_fail(MRL_ERR_INVALID, "mrl_expr: null description");
  if (d->nvars < 0 || d->nvars > 16) return mrl_fail(MRL_ERR_INVALID, "mrl_expr: at most 16 inputs");
  pr.extra = d->extra_symbols != 0;
  pr.expand = d->expand;
  for (int i = 0; i < d->nvars; ++i) {
    pr.vars.push_back(d->var_names[i]);
    pr.layouts.push_back(d->var_layouts ? d->var_layouts[i] : MRL_VAR_REAL);
  }
  for (int i = 0; i < d->nconstants; ++i) pr.constants[d->constant_names[i]] = d->constant_values[i];
  std::vector<std::string> der;
  for (int i = 0; i < d->nderivatives; ++i) der.push_back(d->derivatives[i]);
  try {
    build_ast(pr, d->expression, der);
  } catch (const ParseError &e) {
    return mrl_fail(MRL_ERR_PARSE, "%s", e.what());
  } catch (const std::exception &e) {
    return mrl_fail(MRL_ERR_PARSE, "%s", e.what());
  }
  return MRL_OK;
}

static int copy_out(const std::string &s, char *buf, size_t cap) {
  if (!buf || cap == 0) return mrl_fail(MRL_ERR_INVALID, "null output buffer");
  if (s.size() + 1 > cap) return mrl_fail(MRL_ERR_INVALID, "output buffer too small (%zu bytes needed)", s.size() + 1);
  memcpy(buf, s.c_str(), s.size() + 1);
  return MRL_OK;
}

extern "C" int mrl_expr_simplified(const mrl_expr_desc *d, char *buf, size_t cap) {
  ExprProgram pr;
  int rc = mrlx_fill_program(pr, d);
  if (rc) return rc;
  return copy_out(to_string(pr.ast), buf, cap);
}

extern "C" int mrl_expr_constant(const char *expression, int nconst, const char *const *names, const double *values, double *out) {
  if (!expression || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_expr_constant: bad arguments");
  try {
    std::map<std::string, double> env;
    for (int i = 0; i < nconst; ++i) env[names[i]] = values[i];
    env["pi"] = 3.14159265358979323846;
    env["e"] = 2.71828182845904523536;
    *out = eval_scalar(simplify(parse(expression, {})), env);
  } catch (const std::exception &e) {
    return mrl_fail(MRL_ERR_PARSE, "%s", e.what());
  }
  return MRL_OK;
}

extern "C" int mrl_expr_check(const mrl_expr_desc *d, int precision, char *src_buf, size_t cap) {
  ExprProgram pr;
  int rc = mrlx_fill_program(pr, d);
  if (rc) return rc;
  int rt, space;
  bool nc;
  std::string src;
  try {
    src = generic_source(pr, precision, rt, space, nc);
  } catch (const std::exception &e) {
    return mrl_fail(MRL_ERR_PARSE, "%s", e.what());
  }
  if (src_buf && (rc = copy_out(src, src_buf, cap))) return rc;
  std::vector<char> cubin;
  std::vector<std::string> low;
  std::string log;
  return mrlx_nvrtc_compile(src, {}, {}, cubin, low, log);
}

extern "C" int mrl_expr_compile(mrl_context *ctx, const mrl_expr_desc *d, mrl_expr **out) {
  if (!ctx || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_expr_compile: bad arguments");
  if (!ctx->dim) return mrl_fail(MRL_ERR_INVALID, "mrl_expr_compile: set the domain first");
  auto *e = new mrl_expr();
  e->ctx = ctx;
  int rc = mrlx_fill_program(e->pr, d);
  if (rc) {
    delete e;
    return rc;
  }
  try {
    e->source = generic_source(e->pr, ctx->precision, e->result_type, e->space, e->needs_coords);
  } catch (const std::exception &ex) {
    delete e;
    return mrl_fail(MRL_ERR_PARSE, "%s", ex.what());
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) {
    delete e;
    return mrl_fail(MRL_ERR_CUDA, "cudaSetDevice failed");
  }
  std::vector<char> cubin;
  std::vector<std::string> low;
  std::string log;
  rc = mrlx_nvrtc_compile(e->source, {}, {}, cubin, low, log);
  if (!rc) rc = mrlx_module_load(cubin, "mrl_expr_u32", &e->module, &e->fn32);
  if (!rc) rc = mrlx_module_get(e->module, "mrl_expr_u64", &e->fn64);
  if (rc) {
    mrl_expr_destroy(e);
    return rc;
  }
  *out = e;
  return MRL_OK;
}

extern "C" int mrl_expr_destroy(mrl_expr *e) {
  if (!e) return MRL_OK;
  if (e->ctx) mrl_quiesce(e->ctx);
  mrlx_module_unload(e->module);
  mrlx_module_unload(e->zfwd_module);
  delete e;
  return MRL_OK;
}

extern "C" int mrl_expr_result(const mrl_expr *e, int *space, int *is_complex) {
  if (!e) return mrl_fail(MRL_ERR_INVALID, "null expression");
  if (space) *space = e->space;
  if (is_complex) *is_complex = e->result_type == TC;
  return MRL_OK;
}

extern "C" int mrl_expr_string(const mrl_expr *e, char *buf, size_t cap) {
  if (!e) return mrl_fail(MRL_ERR_INVALID, "null expression");
  return copy_out(to_string(e->pr.ast), buf, cap);
}

extern "C" int mrl_expr_eval(mrl_expr *e, const void *const *inputs, double t, void *out) {
  if (!e || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_expr_eval: bad arguments");
  mrl_context *ctx = e->ctx;
  HostArgs a;
  memset(&a, 0, sizeof a);
  for (size_t v = 0; v < e->pr.vars.size(); ++v) {
    a.in[v] = inputs ? inputs[v] : nullptr;
  }
  a.out = out;
  const int dim = ctx->dim;
  const int *shape = e->space == 2 ? ctx->nr : ctx->n;
  // index space (d0, d1, d2) = the `dim` used axes right-aligned; cmap[axis] = which of (g0,g1,g2)
  // is the coordinate along spatial axis `axis` (3 = the constant 0 for unused axes)
  for (int k = 0; k < 3; ++k) a.d[k] = 1;
  for (int ax = 0; ax < 3; ++ax) a.cmap[ax] = 3;
  for (int ax = 0; ax < dim; ++ax) {
    a.d[3 - dim + ax] = shape[ax];
    a.cmap[ax] = 3 - dim + ax;
  }
  a.total = e->space == 0 ? 1 : a.d[0] * a.d[1] * a.d[2];
  for (int ax = 0; ax < 3; ++ax) {
    a.axis[ax] = ctx->axis_dev[ax];
    a.kaxis[ax] = ctx->kaxis_dev[ax];
  }
  a.t = t;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return mrl_fail(MRL_ERR_CUDA, "cudaSetDevice failed");
  long long blocks = (a.total + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  void *params[] = {&a};
  ctx->launches++;
  return mrlx_launch(a.total < (1LL << 32) ? e->fn32 : e->fn64, (unsigned)blocks, 256, 0, ctx->stream, params);
}
