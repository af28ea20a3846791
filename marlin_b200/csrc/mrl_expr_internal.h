// Internal state of a compiled expression (mrl_expr) and the NVRTC / driver helpers shared by the
// generic pointwise kernel and the expression-specialised first FFT pass.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/marlin_b200.h"
#include "mrl_expr_ast.h"

struct mrl_context;

namespace mrlx {
struct ExprProgram {
  std::vector<std::string> vars;   // `inputs`, in order
  std::vector<int> layouts;        // mrl_var_layout per input
  std::map<std::string, double> constants;
  bool extra = false;              // extra_symbols: x y z kx ky kz k2 t pi e i
  int expand = 0;
  P ast;                           // parsed, differentiated, simplified
};
void build_ast(ExprProgram &pr, const std::string &expression, const std::vector<std::string> &derivatives);
// result_type: 0 bool, 1 real, 2 complex
void generate_body(const ExprProgram &pr, std::string &bindings, std::string &result, int &result_type, std::set<std::string> &used);
}  // namespace mrlx

struct mrl_expr {
  mrl_context *ctx = nullptr;
  mrlx::ExprProgram pr;
  std::string source;
  int result_type = 1, space = 1;
  bool needs_coords = false;
  void *module = nullptr, *fn32 = nullptr, *fn64 = nullptr;
  // expression-specialised first pass of the fused split plan (built on first use)
  void *zfwd_module = nullptr, *zfwd_fn = nullptr;
  int zfwd_n = 0, zfwd_block = 0, zfwd_ppb = 1, zfwd_staged = -2;
  unsigned zfwd_smem = 0;
  int zfwd_kind = 0;
  std::string zfwd_source;
};

int mrlx_fill_program(mrlx::ExprProgram &pr, const mrl_expr_desc *d);
const char *mrlx_expr_prelude();
int mrlx_nvrtc_compile(const std::string &src, const std::vector<std::pair<std::string, std::string>> &headers,
                       const std::vector<std::string> &name_exprs, std::vector<char> &cubin, std::vector<std::string> &lowered,
                       std::string &log);
int mrlx_module_load(const std::vector<char> &cubin, const std::string &fn, void **module, void **function);
int mrlx_module_get(void *module, const std::string &fn, void **function);
void mrlx_module_unload(void *module);
int mrlx_launch(void *function, unsigned grid, unsigned block, unsigned smem, cudaStream_t stream, void **params);
