// The operators of the spectral path as MOOSE-style objects on top of the C ABI.
//   ConstantTensor / ConstantReciprocalTensor   src/tensor_computes/ConstantTensor.C:17-53
//   RandomTensor                                src/tensor_computes/RandomTensor.C:17-55
//   ForwardFFT / InverseFFT                     src/tensor_computes/PerformFFT.C:17-40
//   ReciprocalLaplacianFactor                   src/tensor_computes/ReciprocalLaplacianFactor.C:14-33
//   ReciprocalLaplacianSquareFactor             src/tensor_computes/ReciprocalLaplacianSquareFactor.C:14-34
//   ParsedCompute                               src/tensor_computes/ParsedCompute.C:20-265
//   FFTGradient / FFTGradientSquare             src/tensor_computes/FFTGradient.C:15-40, FFTGradientSquare.C:15-48
//   FFTSemiImplicit                             src/tensor_timeintegrators/FFTSemiImplicit.C:16-62
#include "TensorComputes.h"

#include <cmath>
#include <cstring>

using marlin::Space;
using marlin::Tensor;

// ------------------------------------------------------------------------------------ ExprKernel
ExprKernel::~ExprKernel() { reset(); }
void ExprKernel::reset() {
  if (_expr) mrl_expr_destroy(_expr);
  _expr = nullptr;
}

int ExprKernel::layoutOf(const Tensor &t) {
  if (t.space() == Space::SCALAR) return MRL_VAR_SCALAR;
  if (t.space() == Space::REAL) return t.is_complex() ? MRL_VAR_REAL_COMPLEX : MRL_VAR_REAL;
  return t.is_complex() ? MRL_VAR_RECIP_COMPLEX : MRL_VAR_RECIP_REAL;
}

void ExprKernel::configure(const std::string &expression, std::vector<std::string> inputs, std::vector<std::string> derivatives,
                           std::vector<std::string> constant_names, std::vector<double> constant_values, bool extra_symbols, int expand) {
  reset();
  _expression = expression;
  _inputs = std::move(inputs);
  _derivatives = std::move(derivatives);
  _constant_names = std::move(constant_names);
  _constant_values = std::move(constant_values);
  _extra = extra_symbols;
  _expand = expand;
  _layouts.clear();
}

void ExprKernel::fillDesc(mrl_expr_desc &d, std::vector<const char *> &in, std::vector<const char *> &der, std::vector<const char *> &cn,
                          const std::vector<int> &layouts) const {
  in.clear();
  der.clear();
  cn.clear();
  for (const auto &s : _inputs) in.push_back(s.c_str());
  for (const auto &s : _derivatives) der.push_back(s.c_str());
  for (const auto &s : _constant_names) cn.push_back(s.c_str());
  std::memset(&d, 0, sizeof d);
  d.expression = _expression.c_str();
  d.nvars = (int)in.size();
  d.var_names = in.data();
  d.var_layouts = layouts.empty() ? nullptr : layouts.data();
  d.nderivatives = (int)der.size();
  d.derivatives = der.data();
  d.nconstants = (int)cn.size();
  d.constant_names = cn.data();
  d.constant_values = _constant_values.data();
  d.extra_symbols = _extra ? 1 : 0;
  d.expand = _expand;
}

std::string ExprKernel::simplified() const {
  mrl_expr_desc d;
  std::vector<const char *> in, der, cn;
  fillDesc(d, in, der, cn, {});
  std::vector<char> buf(1 << 16);
  if (mrl_expr_simplified(&d, buf.data(), buf.size()) != MRL_OK) ::mooseError(mrl_last_error());
  return buf.data();
}

Tensor ExprKernel::eval(const DomainAction &domain, const std::vector<const Tensor *> &inputs, double t) {
  std::vector<int> layouts;
  for (const Tensor *in : inputs) {
    if (!in->defined()) ::mooseError("an input tensor of the expression '", _expression, "' is not defined yet");
    if (in->ncomp() != 1) ::mooseError("expressions act on scalar fields; got a tensor with ", in->ncomp(), " components");
    layouts.push_back(layoutOf(*in));
  }
  if (!_expr || layouts != _layouts) {
    reset();
    mrl_expr_desc d;
    std::vector<const char *> in, der, cn;
    fillDesc(d, in, der, cn, layouts);
    if (mrl_expr_compile(domain.context(), &d, &_expr) != MRL_OK) ::mooseError(mrl_last_error());
    _layouts = layouts;
    domain.check(mrl_expr_result(_expr, &_space, &_is_complex), "mrl_expr_result");
  }
  Tensor out = domain.empty(_space == 0 ? Space::SCALAR : (_space == 1 ? Space::REAL : Space::RECIPROCAL), _is_complex != 0, 1);
  std::vector<const void *> ptrs;
  for (const Tensor *in : inputs) ptrs.push_back(in->data_ptr());
  domain.check(mrl_expr_eval(_expr, ptrs.data(), t, out.data_ptr()), "mrl_expr_eval");
  return out;
}

// -------------------------------------------------------------------------------- ConstantTensor
registerMooseObject("MarlinApp", ConstantTensor);
registerMooseObject("MarlinApp", ConstantReciprocalTensor);

template <bool reciprocal>
InputParameters ConstantTensorTempl<reciprocal>::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addParam<MarlinConstantName>("imaginary", "0.0", "Imaginary part of the constant value.");
  if (reciprocal)
    params.addClassDescription("Constant tensor in reciprocal space.");
  else {
    params.addClassDescription("Constant tensor in real space.");
    params.suppressParameter<MarlinConstantName>("imaginary");
  }
  params.addParam<MarlinConstantName>("real", "0.0", "Real part of the constant value.");
  params.addParam<bool>("full", false, "Construct a full tensor will all entries");
  return params;
}

template <bool reciprocal>
ConstantTensorTempl<reciprocal>::ConstantTensorTempl(const InputParameters &parameters) : TensorOperator<>(parameters) {}

template <bool reciprocal>
void ConstantTensorTempl<reciprocal>::computeBuffer() {
  // the reference builds a one-element tensor expanded to the grid shape; here the field is filled
  const Real re = getConstant("real");
  const Real im = reciprocal ? getConstant("imaginary") : 0.0;
  const Space sp = reciprocal ? Space::RECIPROCAL : Space::REAL;
  Tensor t = _domain.empty(sp, reciprocal, 1);
  const size_t n = size_t(t.count());
  if (re == 0.0 && im == 0.0) {
    checkC(mrl_memset(_domain.context(), t.data_ptr(), 0, t.nbytes()), "mrl_memset");
  } else {
    std::vector<double> host(reciprocal ? 2 * n : n);
    if (reciprocal)
      for (size_t i = 0; i < n; ++i) {
        host[2 * i] = re;
        host[2 * i + 1] = im;
      }
    else
      std::fill(host.begin(), host.end(), re);
    t = _domain.fromHost(host, sp, reciprocal, 1);
  }
  _u = t;
}
template class ConstantTensorTempl<false>;
template class ConstantTensorTempl<true>;

// ---------------------------------------------------------------------------------- RandomTensor
registerMooseObject("MarlinApp", RandomTensor);

namespace {
// ATen's CPU generator (at::mt19937 + uniform_real_distribution): bit-for-bit torch::rand on the host.
struct TorchMT19937 {
  uint32_t mt[624];
  int idx = 624;
  void seed(uint64_t s) {
    mt[0] = uint32_t(s & 0xffffffffu);
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + uint32_t(i);
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};
TorchMT19937 &globalGenerator() {
  static TorchMT19937 g = [] {
    TorchMT19937 x;
    x.seed(67280421310721ull);  // c10::detail::default_rng_seed_val
    return x;
  }();
  return g;
}
}  // namespace

void marlinTorchRand(std::vector<double> &out, size_t n, bool single, double min, double max, const int *seed) {
  auto &g = globalGenerator();
  if (seed) g.seed(uint64_t(int64_t(*seed)));
  out.resize(n);
  if (single) {
    const float fmin = float(min), fr = float(max - min);
    for (size_t i = 0; i < n; ++i) {
      const float r = float(g.next() & ((1u << 24) - 1)) * (1.0f / float(1u << 24));
      volatile float scaled = r * fr;  // two roundings, like the separate mul and add in ATen
      out[i] = double(float(scaled + fmin));
    }
  } else {
    const double range = max - min;
    for (size_t i = 0; i < n; ++i) {
      const uint64_t hi = g.next(), lo = g.next();
      const uint64_t x = ((hi << 32) | lo) & ((1ull << 53) - 1);
      const double r = double(x) * (1.0 / double(1ull << 53));
      volatile double scaled = r * range;
      out[i] = scaled + min;
    }
  }
}

InputParameters RandomTensor::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Uniform random IC with values between `min` and `max`.");
  params.addRequiredParam<Real>("min", "Minimum value.");
  params.addRequiredParam<Real>("max", "Maximum value.");
  params.addParam<int>("seed", "Random number seed.");
  params.addParam<bool>("generate_on_cpu", true, "To ensure reproducibility across devices it is recommended to generate random tensors on the CPU.");
  return params;
}

RandomTensor::RandomTensor(const InputParameters &parameters) : TensorOperator<>(parameters) {
  if (!getParam<bool>("generate_on_cpu"))
    mooseWarning("generate_on_cpu = false: the device generator of libTorch is not reproduced; generating on the host (bit-identical to torch::rand on the CPU).");
}

void RandomTensor::computeBuffer() {
  std::vector<double> host;
  int seed = 0;
  const bool has_seed = isParamValid("seed");
  if (has_seed) seed = getParam<int>("seed");
  marlinTorchRand(host, size_t(_domain.getNumberOfLocalCells()), _domain.single(), getParam<Real>("min"), getParam<Real>("max"), has_seed ? &seed : nullptr);
  _u = _domain.fromHost(host, Space::REAL, false, 1);
}

// ------------------------------------------------------------------------------------ PerformFFT
registerMooseObject("MarlinApp", ForwardFFT);
registerMooseObject("MarlinApp", InverseFFT);

template <bool forward>
InputParameters PerformFFTTempl<forward>::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("PerformFFT object.");
  params.addParam<TensorInputBufferName>("input", "Input buffer name");
  return params;
}
template <bool forward>
PerformFFTTempl<forward>::PerformFFTTempl(const InputParameters &parameters) : TensorOperator<>(parameters), _input(getInputBuffer("input")) {}
template <bool forward>
void PerformFFTTempl<forward>::computeBuffer() {
  if (forward)
    _u = _domain.fft(_input);
  else
    _u = _domain.ifft(_input);
}
template class PerformFFTTempl<true>;
template class PerformFFTTempl<false>;

// ------------------------------------------------------------------------ ReciprocalLaplacian*Factor
registerMooseObject("MarlinApp", ReciprocalLaplacianFactor);
registerMooseObject("MarlinApp", ReciprocalLaplacianSquareFactor);

template <int kind>
InputParameters ReciprocalLaplacianFactorTempl<kind>::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription(kind == MRL_KFACTOR_LAPLACIAN ? "Reciprocal space Laplacian IC." : "Reciprocal space square Laplacian IC.");
  params.addParam<Real>("factor", 1.0, "Prefactor");
  return params;
}
template <int kind>
ReciprocalLaplacianFactorTempl<kind>::ReciprocalLaplacianFactorTempl(const InputParameters &parameters)
  : TensorOperator<>(parameters), _factor(getParam<Real>("factor")) {}
template <int kind>
void ReciprocalLaplacianFactorTempl<kind>::computeBuffer() {
  Tensor t = _domain.empty(Space::RECIPROCAL, false, 1);
  checkC(mrl_kfactor(_domain.context(), kind, _factor, t.data_ptr()), "mrl_kfactor");
  _u = t;
}
template class ReciprocalLaplacianFactorTempl<MRL_KFACTOR_LAPLACIAN>;
template class ReciprocalLaplacianFactorTempl<MRL_KFACTOR_LAPLACIAN_SQUARE>;

// --------------------------------------------------------------------------------- ParsedCompute
registerMooseObject("MarlinApp", ParsedCompute);

InputParameters ParsedCompute::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("ParsedCompute object.");
  params.addRequiredParam<std::string>("expression", "Parsed expression");
  params.addParam<std::vector<TensorInputBufferName>>("inputs", {}, "Buffer names used in the expression");
  params.addParam<std::vector<TensorInputBufferName>>("derivatives", {}, "List of inputs to take the derivative w.r.t. (or none)");
  params.addParam<bool>("enable_fpoptimizer", true, "Use algebraic optimizer");
  params.addParam<bool>("extra_symbols", false,
                        "Provide i (imaginary unit), kx,ky,kz (reciprocal space frequency), k2 (square of the k-vector), x,y,z "
                        "(real space coordinates), time t, pi, and e.");
  params.addParam<std::vector<std::string>>("constant_names", {}, "Vector of constants used in the parsed function");
  params.addParam<std::vector<std::string>>("constant_expressions", {}, "Vector of values for the constants in constant_names (can be an FParser expression)");
  params.addParam<MooseEnum>("expand", MooseEnum("REAL RECIPROCAL NONE", "NONE"), "Expand the tensor to full size.");
  params.addParam<bool>("is_integer", false, "Turn the function result into an integer tensor");
  return params;
}

ParsedCompute::ParsedCompute(const InputParameters &parameters) : TensorOperator<>(parameters), _extra_symbols(getParam<bool>("extra_symbols")) {
  const auto expression = getParam<std::string>("expression");
  const auto names = getParam<std::vector<TensorInputBufferName>>("inputs");
  const auto derivatives = getParam<std::vector<TensorInputBufferName>>("derivatives");
  if (getParam<bool>("is_integer")) paramError("is_integer", "integer tensors are not on the spectral path and are not supported by marlin_b200");

  // derivatives must name inputs (ParsedCompute.C:150-157)
  for (const auto &d : derivatives) {
    bool ok = false;
    for (const auto &n : names) ok = ok || n == d;
    if (!ok) paramError("derivatives", "Derivative w.r.t `", d, "` was requested, but it is not listed in `inputs`.");
  }
  // constants: each expression may use the previously evaluated ones (ParsedCompute.C:104-123)
  const auto cnames = getParam<std::vector<std::string>>("constant_names");
  const auto cexprs = getParam<std::vector<std::string>>("constant_expressions");
  if (cnames.size() != cexprs.size()) paramError("constant_names", "Must have the same number of entries as 'constant_expressions'.");
  std::vector<double> cvals;
  for (size_t i = 0; i < cnames.size(); ++i) {
    std::vector<const char *> nm;
    for (size_t j = 0; j < i; ++j) nm.push_back(cnames[j].c_str());
    double v = 0;
    if (mrl_expr_constant(cexprs[i].c_str(), (int)i, nm.data(), cvals.data(), &v) != MRL_OK)
      paramError("constant_expressions", "Invalid constant expression '", cexprs[i], "': ", mrl_last_error());
    cvals.push_back(v);
  }
  for (const auto &n : names) _params.push_back(&getInputBufferByName(n));

  const auto expand = getParam<MooseEnum>("expand");
  const int ex = expand == "REAL" ? MRL_EXPAND_REAL : (expand == "RECIPROCAL" ? MRL_EXPAND_RECIPROCAL : MRL_EXPAND_NONE);
  _kernel.configure(expression, names, derivatives, cnames, cvals, _extra_symbols, ex);
  // parse now so that syntax errors surface at construction, like the reference
  try {
    _kernel.simplified();
  } catch (const MooseException &e) {
    paramError("expression", "Invalid function\n", expression, "\nin ParsedCompute.\n", e.what());
  }
}

void ParsedCompute::computeBuffer() { _u = _kernel.eval(_domain, _params, _time); }

// ----------------------------------------------------------------------------------- FFTGradient
registerMooseObject("MarlinApp", FFTGradient);

InputParameters FFTGradient::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Tensor gradient.");
  params.addRequiredParam<TensorInputBufferName>("input", "Input buffer name");
  params.addParam<bool>("input_is_reciprocal", false, "Input buffer is already in reciprocal space");
  params.addRequiredParam<MooseEnum>("direction", MooseEnum("X=0 Y=1 Z=2"), "Which axis to take the gradient along.");
  return params;
}

static const char *const kAxisName[3] = {"kx", "ky", "kz"};

FFTGradient::FFTGradient(const InputParameters &parameters)
  : TensorOperator<>(parameters),
    _input(getInputBuffer("input")),
    _input_is_reciprocal(getParam<bool>("input_is_reciprocal")),
    _direction(int(getParam<MooseEnum>("direction"))) {
  // ifft(ubar * k_d * i): the multiplication is one generated kernel
  _kernel.configure(std::string("ubar*") + kAxisName[_direction] + "*i", {"ubar"}, {}, {}, {}, true, MRL_EXPAND_NONE);
}

void FFTGradient::computeBuffer() {
  Tensor ubar = _input_is_reciprocal ? _input : _domain.fft(_input);
  _u = _domain.ifft(_kernel.eval(_domain, {&ubar}, _time));
}

// ----------------------------------------------------------------------------- FFTGradientSquare
registerMooseObject("MarlinApp", FFTGradientSquare);

InputParameters FFTGradientSquare::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Tensor gradient.");
  params.addRequiredParam<TensorInputBufferName>("input", "Input buffer name");
  params.addParam<bool>("input_is_reciprocal", false, "Input buffer is already in reciprocal space");
  params.addParam<Real>("factor", 1.0, "Prefactor to the gradient square");
  return params;
}

FFTGradientSquare::FFTGradientSquare(const InputParameters &parameters)
  : TensorOperator<>(parameters), _input(getInputBuffer("input")), _input_is_reciprocal(getParam<bool>("input_is_reciprocal")), _factor(getParam<Real>("factor")) {
  for (unsigned int d = 0; d < _dim; ++d) _grad[d].configure(std::string("ubar*") + kAxisName[d] + "*i", {"ubar"}, {}, {}, {}, true, MRL_EXPAND_NONE);
  // (gx^2 [+ gy^2 [+ gz^2]]) [* factor], summed in the reference's order
  std::string e = "gx*gx";
  std::vector<std::string> in = {"gx"};
  if (_dim > 1) {
    e = e + " + gy*gy";
    in.push_back("gy");
  }
  if (_dim > 2) {
    e = e + " + gz*gz";
    in.push_back("gz");
  }
  if (_factor != 1.0) e = "(" + e + ")*factor";
  _square.configure(e, in, {}, {"factor"}, {_factor}, false, MRL_EXPAND_NONE);
}

void FFTGradientSquare::computeBuffer() {
  Tensor ubar = _input_is_reciprocal ? _input : _domain.fft(_input);
  std::vector<Tensor> g(_dim);
  std::vector<const Tensor *> gp;
  for (unsigned int d = 0; d < _dim; ++d) {
    g[d] = _domain.ifft(_grad[d].eval(_domain, {&ubar}, _time));
    gp.push_back(&g[d]);
  }
  _u = _square.eval(_domain, gp, _time);
}

// ------------------------------------------------------------------------------- FFTSemiImplicit
registerMooseObject("MarlinApp", FFTSemiImplicit);

InputParameters FFTSemiImplicit::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Semi-implicit time integrator.");
  params.addRequiredParam<TensorInputBufferName>("reciprocal_buffer", "Buffer with the reciprocal of the integrated buffer");
  params.addRequiredParam<TensorInputBufferName>("linear_reciprocal", "Buffer with the reciprocal of the linear prefactor (e.g. kappa*k^2)");
  params.addRequiredParam<TensorInputBufferName>("nonlinear_reciprocal", "Buffer with the reciprocal of the non-linear contribution");
  params.addParam<unsigned int>("history_size", 1, "How many old states to use (determines time integration order).");
  return params;
}

FFTSemiImplicit::FFTSemiImplicit(const InputParameters &parameters)
  : TensorOperator<>(parameters),
    _history_size(getParam<unsigned int>("history_size")),
    _sub_dt(_tensor_problem.subDt()),
    _reciprocal_buffer(getInputBuffer("reciprocal_buffer")),
    _linear_reciprocal(getInputBuffer("linear_reciprocal")),
    _non_linear_reciprocal(getInputBuffer("nonlinear_reciprocal")),
    _old_reciprocal_buffer(_tensor_problem.getBufferOld(getParam<TensorInputBufferName>("reciprocal_buffer"), _history_size)),
    _old_non_linear_reciprocal(_tensor_problem.getBufferOld(getParam<TensorInputBufferName>("nonlinear_reciprocal"), _history_size)) {}

void FFTSemiImplicit::computeBuffer() {
  // FFTSemiImplicit.C:43-62: the sub step is whatever the TensorSolver wrote into TensorProblem::subDt()
  const Real dt = _sub_dt;
  const auto n_old = std::min(_old_reciprocal_buffer.size(), _old_non_linear_reciprocal.size());
  Tensor ubar = _domain.empty(Space::RECIPROCAL, true, 1);
  if (n_old == 0) {
    const double beta[1] = {1.0};
    checkC(mrl_ab_update(_domain.context(), ubar.data_ptr(), _reciprocal_buffer.data_ptr(), _non_linear_reciprocal.data_ptr(), _linear_reciprocal.data_ptr(), dt,
                         beta, 0, nullptr),
           "mrl_ab_update");
  } else {
    const double beta[2] = {1.5, -0.5};
    const void *old[1] = {_old_non_linear_reciprocal[0].data_ptr()};
    checkC(mrl_ab_update(_domain.context(), ubar.data_ptr(), _reciprocal_buffer.data_ptr(), _non_linear_reciprocal.data_ptr(), _linear_reciprocal.data_ptr(), dt,
                         beta, 1, old),
           "mrl_ab_update");
  }
  _u = _domain.ifft(ubar);
}

// --------------------------------------------------------------------------- SwiftHohenbergLinear
registerMooseObject("MarlinApp", SwiftHohenbergLinear);

InputParameters SwiftHohenbergLinear::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Reciprocal space linear term in the semi-implicit time integration of the Swift-Hohenberg equation IC.");
  params.addParam<Real>("r", -0.5, "Phase field crystal parameter r");
  params.addParam<Real>("alpha", 1.0, "Regularization factor <=1");
  return params;
}
SwiftHohenbergLinear::SwiftHohenbergLinear(const InputParameters &parameters) : TensorOperator<>(parameters) {
  // src/tensor_computes/SwiftHohenbergLinear.C:33-36
  _kernel.configure("r - alpha*alpha*(1 - k2)*(1 - k2)", {}, {}, {"r", "alpha"}, {getParam<Real>("r"), getParam<Real>("alpha")}, true, MRL_EXPAND_RECIPROCAL);
}
void SwiftHohenbergLinear::computeBuffer() { _u = _kernel.eval(_domain, {}, _time); }

// ------------------------------------------------------------------------- SmoothRectangleCompute
registerMooseObject("MarlinApp", SmoothRectangleCompute);

InputParameters SmoothRectangleCompute::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Interpolate a value between the inside and outside of a rectangle smoothly.");
  params.addRequiredParam<Real>("x1", "The x coordinate of the lower left-hand corner of the box.");
  params.addRequiredParam<Real>("x2", "The x coordinate of the upper right-hand corner of the box.");
  params.addRequiredParam<Real>("y1", "The y coordinate of the lower left-hand corner of the box.");
  params.addRequiredParam<Real>("y2", "The y coordinate of the upper right-hand corner of the box.");
  params.addParam<Real>("z1", 0, "The z coordinate of the lower left-hand corner of the box.");
  params.addParam<Real>("z2", 0, "The z coordinate of the upper right-hand corner of the box.");
  params.addParam<MooseEnum>("profile", MooseEnum("COS TANH"), "Functional dependence for the interface profile");
  params.addParam<Real>("int_width", 0, "The width of the diffuse interface. Set to 0 for sharp interface.");
  params.addParam<Real>("inside", 1, "The value inside the rectangle.");
  params.addParam<Real>("outside", 0, "The value outside the rectangle.");
  return params;
}

// src/tensor_computes/SmoothRectangleCompute.C:60-131 as ONE generated kernel over the cell-centre
// axes: per used axis the distance d to the nearest face, h = [d inside] (sharp), 0.5 + 0.5 sin(pi
// clamp(d, -w/2, w/2) / w) (COS) or 0.5 + 0.5 tanh(4 d / w) (TANH); u = H inside + (1 - H) outside
// with H the product over the axes.  The reference's fill values for the unused axes (w/2, 10 w,
// 5 w) give a factor of exactly one, so those factors are left out.
std::string SmoothRectangleCompute::expression(unsigned int dim, Real w, const std::string &profile) {
  static const char *ax[3] = {"x", "y", "z"};
  std::string expr;
  if (w <= 0.0) {
    std::string cond;
    for (unsigned int d = 0; d < dim; ++d)
      cond += std::string(d ? " & " : "") + ax[d] + " >= " + ax[d] + "1 & " + ax[d] + " <= " + ax[d] + "2";
    return "if(" + cond + ", vin, vout)";
  }
  if (profile != "COS" && profile != "TANH") return "vout";  // no profile chosen: the reference leaves the indicator at zero
  std::string prod;
  for (unsigned int d = 0; d < dim; ++d) {
    const std::string a = ax[d], dist = "dist_" + a, h = "blend_" + a;
    expr += dist + " := min(" + a + " - " + a + "1, " + a + "2 - " + a + "); ";
    if (profile == "COS")
      expr += h + " := 0.5 + 0.5*sin(pi*max(-w2, min(w2, " + dist + "))/w); ";
    else
      expr += h + " := 0.5 + 0.5*tanh(4*" + dist + "/w); ";
    prod += (d ? "*" : "") + h;
  }
  return expr + "blend := " + prod + "; blend*vin + (1 - blend)*vout";
}

SmoothRectangleCompute::SmoothRectangleCompute(const InputParameters &parameters) : TensorOperator<>(parameters) {
  const Real w = getParam<Real>("int_width");
  if (w < 0.0) mooseError("Interface width must be a non-negative real number.");
  const std::string profile = isParamValid("profile") ? std::string(getParam<MooseEnum>("profile")) : std::string();
  _kernel.configure(expression(_dim, w, profile), {}, {},
                    {"x1", "x2", "y1", "y2", "z1", "z2", "w", "w2", "vin", "vout", "pi"},
                    {getParam<Real>("x1"), getParam<Real>("x2"), getParam<Real>("y1"), getParam<Real>("y2"), getParam<Real>("z1"), getParam<Real>("z2"), w,
                     w / 2.0, getParam<Real>("inside"), getParam<Real>("outside"), M_PI},
                    true, MRL_EXPAND_REAL);
}
void SmoothRectangleCompute::computeBuffer() { _u = _kernel.eval(_domain, {}, _time); }

// ---------------------------------------------------------------------------- MooseFunctionTensor
registerMooseObject("MarlinApp", MooseFunctionTensor);

InputParameters MooseFunctionTensor::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Map a MooseFunction to a tensor.");
  params.addRequiredParam<FunctionName>("function", "Function to map.");
  return params;
}
MooseFunctionTensor::MooseFunctionTensor(const InputParameters &parameters) : TensorOperator<>(parameters), _function(getParam<FunctionName>("function")) {}

// The reference samples Function::value at the points i*dx + dx/2 on the host
// (src/tensor_computes/MooseFunctionTensor.C:31-72).  Here the ParsedFunction expression (FParser
// grammar: `:=` bindings, if(), ^, elementary functions, pi - a subset of the Marlin grammar) is
// compiled into a device kernel over the cell-centre axes; symbols bound to other functions are
// evaluated first and enter as input fields.
Tensor MooseFunctionTensor::evaluate(const std::string &function, int depth) {
  if (depth > 16) mooseError("[Functions]: symbol_values of '", function, "' are nested too deep (cyclic?)");
  const auto *f = _tensor_problem.getFunction(function);
  if (!f) paramError("function", "no ParsedFunction named '", function, "' in [Functions]");
  if (f->symbol_names.size() != f->symbol_values.size()) mooseError("[Functions/", function, "]: symbol_names and symbol_values differ in length");
  // sample points i*dx + dx/2: the cell-centre axes shifted by the domain minimum (the reference
  // ignores the minimum, MooseFunctionTensor.C:45-66), materialised as coordinate fields
  if (_coords.empty()) {
    static const char *axis[3] = {"x", "y", "z"};
    for (unsigned int d = 0; d < 3; ++d) {
      ExprKernel k;
      k.configure(std::string(axis[d]) + " - shift", {}, {}, {"shift"}, {d < _dim ? _domain.getDomainMin()[d] : 0.0}, true, MRL_EXPAND_REAL);
      _coords.push_back(k.eval(_domain, {}, 0.0));
    }
  }
  std::vector<std::string> inputs = {"x", "y", "z"}, cnames = {"pi", "e", "t"};
  std::vector<double> cvalues = {M_PI, M_E, _tensor_problem.time()};
  std::vector<Tensor> held;
  for (std::size_t i = 0; i < f->symbol_names.size(); ++i) {
    if (_tensor_problem.getFunction(f->symbol_values[i])) {
      inputs.push_back(f->symbol_names[i]);
      held.push_back(evaluate(f->symbol_values[i], depth + 1));
    } else {
      cnames.push_back(f->symbol_names[i]);
      cvalues.push_back(shim_detail::Conv<double>::from(f->symbol_values[i], "Functions/" + function + "/symbol_values"));
    }
  }
  ExprKernel k;
  k.configure(f->expression, inputs, {}, cnames, cvalues, false, MRL_EXPAND_REAL);
  std::vector<const Tensor *> in = {&_coords[0], &_coords[1], &_coords[2]};
  for (const auto &t : held) in.push_back(&t);
  return k.eval(_domain, in, 0.0);
}

void MooseFunctionTensor::computeBuffer() {
  _u = evaluate(_function, 0);
  _coords.clear();
}

// -------------------------------------------------------------------------- ReciprocalMatDiffusion
registerMooseObject("MarlinApp", ReciprocalMatDiffusion);

InputParameters ReciprocalMatDiffusion::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Calculates the divergence of flux for a variable mobility in reciprocal space.");
  params.addRequiredParam<TensorInputBufferName>("chemical_potential", "Chemical potential buffer name");
  params.addRequiredParam<TensorInputBufferName>("mobility", "Mobility buffer name");
  params.addParam<TensorInputBufferName>("psi", "Variable to impose Neuamnn BC.");
  params.addParam<bool>("always_update_psi", false, "Set to true if the BC changes .");
  return params;
}

ReciprocalMatDiffusion::ReciprocalMatDiffusion(const InputParameters &parameters)
  : TensorOperator<>(parameters),
    _chem_pot(getInputBuffer("chemical_potential")),
    _M(getInputBuffer("mobility")),
    _psi(getInputBuffer("psi")),
    _always_update_psi(getParam<bool>("always_update_psi")) {
  // src/tensor_computes/ReciprocalMatDiffusion.C:43-66, each line one generated kernel
  for (int d = 0; d < 3; ++d) _grad[d].configure(std::string("ubar*") + kAxisName[d] + "*i", {"ubar"}, {}, {}, {}, true, MRL_EXPAND_NONE);
  _by_psi.configure("if(psi > 0, g/psi, 0)", {"psi", "g"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  _flux.configure("if(psi > 0, M, 0)*g", {"M", "psi", "g"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  _dot[0].configure("gx*Jx", {"gx", "Jx"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  _dot[1].configure("gx*Jx + gy*Jy", {"gx", "Jx", "gy", "Jy"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  _dot[2].configure("gx*Jx + gy*Jy + gz*Jz", {"gx", "Jx", "gy", "Jy", "gz", "Jz"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  _div[0].configure("i*(kx*Jx) + nf", {"Jx", "nf"}, {}, {}, {}, true, MRL_EXPAND_NONE);
  _div[1].configure("i*(kx*Jx + ky*Jy) + nf", {"Jx", "Jy", "nf"}, {}, {}, {}, true, MRL_EXPAND_NONE);
  _div[2].configure("i*(kx*Jx + ky*Jy + kz*Jz) + nf", {"Jx", "Jy", "Jz", "nf"}, {}, {}, {}, true, MRL_EXPAND_NONE);
}

void ReciprocalMatDiffusion::computeBuffer() {
  const unsigned int D = _dim;  // the components along unused dimensions vanish (their k axis is {0})
  if (_update_psi || _always_update_psi) {
    const Tensor psibar = _domain.fft(_psi);
    for (unsigned int d = 0; d < D; ++d) {
      const Tensor g = _domain.ifft(_grad[d].eval(_domain, {&psibar}, _time));
      _grad_psi_by_psi[d] = _by_psi.eval(_domain, {&_psi, &g}, _time);
    }
    _update_psi = false;
  }
  const Tensor mubar = _domain.fft(_chem_pot);
  Tensor J[3], Jbar[3];
  for (unsigned int d = 0; d < D; ++d) {
    const Tensor g = _domain.ifft(_grad[d].eval(_domain, {&mubar}, _time));
    J[d] = _flux.eval(_domain, {&_M, &_psi, &g}, _time);
    Jbar[d] = _domain.fft(J[d]);
  }
  std::vector<const Tensor *> in;
  for (unsigned int d = 0; d < D; ++d) {
    in.push_back(&_grad_psi_by_psi[d]);
    in.push_back(&J[d]);
  }
  const Tensor nf = _domain.fft(_dot[D - 1].eval(_domain, in, _time));
  in.clear();
  for (unsigned int d = 0; d < D; ++d) in.push_back(&Jbar[d]);
  in.push_back(&nf);
  _u = _div[D - 1].eval(_domain, in, _time);
}

// ----------------------------------------------------------------------------- ReciprocalAllenCahn
registerMooseObject("MarlinApp", ReciprocalAllenCahn);

InputParameters ReciprocalAllenCahn::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Calculates the Allen-Cahn bulk driving force masked using psi.");
  params.addRequiredParam<TensorInputBufferName>("dF_chem_deta", "Driving force buffer name");
  params.addRequiredParam<TensorInputBufferName>("L", "Allen-Cahn mobility buffer name");
  params.addRequiredParam<TensorInputBufferName>("psi", "Variable to impose Neumann BC.");
  params.addParam<bool>("always_update_psi", false, "Set to true if the BC changes .");
  return params;
}

ReciprocalAllenCahn::ReciprocalAllenCahn(const InputParameters &parameters)
  : TensorOperator<>(parameters), _dF_chem_deta(getInputBuffer("dF_chem_deta")), _L(getInputBuffer("L")), _psi(getInputBuffer("psi")) {
  // src/tensor_computes/ReciprocalAllenCahn.C:39-50 (psi itself is re-read every time: same values
  // as the cached threshold unless always_update_psi semantics are needed, which re-reading covers)
  _rate.configure("if(psi > 0, -1*L*dF, 0)", {"psi", "L", "dF"}, {}, {}, {}, false, MRL_EXPAND_NONE);
}

void ReciprocalAllenCahn::computeBuffer() { _u = _domain.fft(_rate.eval(_domain, {&_psi, &_L, &_dF_chem_deta}, _time)); }

// -------------------------------------------------------------------------------- DeAliasingTensor
registerMooseObject("MarlinApp", DeAliasingTensor);

InputParameters DeAliasingTensor::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Create a de-aliasing filter.");
  params.addRequiredParam<MooseEnum>("method", MooseEnum("SHARP HOULI"), "Prefactor");
  params.addParam<Real>("p", 16, "Hou-Li filter exponent");
  params.addParam<Real>("alpha", 36, "Hou-Li filter pre-factor");
  return params;
}
DeAliasingTensor::DeAliasingTensor(const InputParameters &parameters)
  : TensorOperator<>(parameters), _houli(getParam<MooseEnum>("method") == "HOULI"), _p(getParam<Real>("p")), _alpha(getParam<Real>("alpha")) {}

// src/tensor_computes/DeAliasingTensor.C:37-62; the maximum frequencies come from the host copies of
// the reciprocal axes (unused dimensions: the one-element axis {0})
void DeAliasingTensor::computeBuffer() {
  double kmax[3] = {0, 0, 0};
  for (unsigned int d = 0; d < _dim; ++d)
    for (double v : _domain.getReciprocalAxis(d)) kmax[d] = std::max(kmax[d], std::fabs(v));
  if (_houli)
    _kernel.configure("exp(-alpha*((abs(kx)/a)^p + (abs(ky)/b)^p + (abs(kz)/c)^p))", {}, {}, {"alpha", "p", "a", "b", "c"},
                      {_alpha, _p, kmax[0] ? kmax[0] : 1.0, kmax[1] ? kmax[1] : 1.0, kmax[2] ? kmax[2] : 1.0}, true, MRL_EXPAND_RECIPROCAL);
  else
    _kernel.configure("if((abs(kx) > a) | (abs(ky) > b) | (abs(kz) > c), 0, 1)", {}, {}, {"a", "b", "c"},
                      {2 * kmax[0] / 3, 2 * kmax[1] / 3, 2 * kmax[2] / 3}, true, MRL_EXPAND_RECIPROCAL);
  _u = _kernel.eval(_domain, {}, _time);
}
