// Postprocessors on tensor buffers (reference citations in include/TensorPostprocessor.h).
#include "TensorPostprocessor.h"

#include <algorithm>
#include <cmath>

#include "TensorComputes.h"

using marlin::Space;
using marlin::Tensor;

InputParameters TensorPostprocessor::validParams() {
  InputParameters params = MooseObject::validParams();
  params.registerBase("Postprocessor");
  params.addClassDescription("A normal Postprocessor acting on a Tensor buffer.");
  params.addRequiredParam<TensorInputBufferName>("buffer", "The buffer this compute is operating on");
  params.addParam<std::string>("execute_on", "TIMESTEP_END", "When to execute (INITIAL, TIMESTEP_BEGIN, TIMESTEP_END, FINAL)");
  params.addParam<std::vector<std::string>>("outputs", "Accepted for input compatibility");
  params.addPrivateParam<TensorProblem *>("_tensor_problem", nullptr);
  return params;
}

TensorPostprocessor::TensorPostprocessor(const InputParameters &parameters)
  : MooseObject(parameters),
    _tensor_problem(*getCheckedPointerParam<TensorProblem>("_tensor_problem")),
    _domain(_tensor_problem.domain()),
    _buffer_name(getParam<TensorInputBufferName>("buffer")),
    _buffer_base(_tensor_problem.getBufferBase(_buffer_name)),
    _u(_buffer_base.getRawTensor()),
    _execute_on(parseExecFlags(getParam<std::string>("execute_on"), _path + "/execute_on")) {}

namespace {

// gatherSum / gatherMin / gatherMax of MOOSE's postprocessors, over the process group of the domain
void gather(const DomainAction &domain, Real &v, Comm::Op op) { domain.comm().allreduce(&v, 1, op); }

class TensorAveragePostprocessor : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Compute the average value over a buffer.");
    return params;
  }
  using TensorPostprocessor::TensorPostprocessor;
  void execute() override {
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    _sum = _domain.sum(_u);
    _numel = Real(_u.numel());
  }
  void finalize() override {
    // src/postprocessors/TensorAveragePostprocessor.C:41-48
    gather(_domain, _sum, Comm::SUM);
    gather(_domain, _numel, Comm::SUM);
    _average = _sum / _numel;
  }
  Real getValue() const override { return _average; }

protected:
  Real _sum = 0, _numel = 0, _average = 0;
};

class TensorIntegralPostprocessor : public TensorAveragePostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorAveragePostprocessor::validParams();
    params.addClassDescription("Compute the integral over a buffer");
    return params;
  }
  using TensorAveragePostprocessor::TensorAveragePostprocessor;
  void finalize() override {
    TensorAveragePostprocessor::finalize();
    Real volume = 1.0;
    for (unsigned int d = 0; d < _domain.getDim(); ++d) volume *= _domain.getDomainMax()[d] - _domain.getDomainMin()[d];
    _integral = _average * volume;
  }
  Real getValue() const override { return _integral; }

protected:
  Real _integral = 0;
};

class TensorExtremeValuePostprocessor : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Find extreme values in the Tensor buffer");
    params.addRequiredParam<MooseEnum>("value_type", MooseEnum("MIN MAX"), "Extreme value type");
    return params;
  }
  explicit TensorExtremeValuePostprocessor(const InputParameters &p) : TensorPostprocessor(p), _is_min(getParam<MooseEnum>("value_type") == "MIN") {}
  void execute() override {
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    _value = _domain.reduce(_is_min ? MRL_MIN : MRL_MAX, _u);
  }
  void finalize() override { gather(_domain, _value, _is_min ? Comm::MIN : Comm::MAX); }  // TensorExtremeValuePostprocessor.C:38-44
  Real getValue() const override { return _value; }

protected:
  const bool _is_min;
  Real _value = 0;
};

class TensorIntegralChangePostprocessor : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Compute the integral of the absolute change of a buffer over a time step");
    return params;
  }
  explicit TensorIntegralChangePostprocessor(const InputParameters &p) : TensorPostprocessor(p), _u_old(_tensor_problem.getBufferOld(_buffer_name, 1)) {
    _diff.configure("abs(u - u_old)", {"u", "u_old"}, {}, {}, {}, false, MRL_EXPAND_NONE);
    _abs.configure("abs(u)", {"u"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  }
  void execute() override {
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    Tensor d = !_u_old.empty() && _u_old[0].defined() ? _diff.eval(_domain, {&_u, &_u_old[0]}, 0.0) : _abs.eval(_domain, {&_u}, 0.0);
    _integral = _domain.sum(d);
    for (unsigned int dd = 0; dd < _domain.getDim(); ++dd) _integral *= _domain.getGridSpacing()[dd];
  }
  void finalize() override { gather(_domain, _integral, Comm::SUM); }
  Real getValue() const override { return _integral; }

protected:
  const std::vector<Tensor> &_u_old;
  ExprKernel _diff, _abs;
  Real _integral = 0;
};

class SemiImplicitCriticalTimeStep : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Compute the critical timestep given the reciprocal space representation of the linear operator in a semi-implicit time integrator.");
    params.addParam<Real>("c", 1.0, "Courant number (CFL factor)");
    return params;
  }
  explicit SemiImplicitCriticalTimeStep(const InputParameters &p) : TensorPostprocessor(p) {
    _norm.configure("L*L", {"L"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  }
  void execute() override {
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    if (_u.is_complex()) mooseError("the linear operator buffer is expected to be real");
    const Real max_norm_k = std::sqrt(_domain.reduce(MRL_MAX, _norm.eval(_domain, {&_u}, 0.0)));
    _critical_dt = max_norm_k > 0.0 ? 1.0 / max_norm_k : 1e30;
  }
  void finalize() override { gather(_domain, _critical_dt, Comm::MIN); }  // SemiImplicitCriticalTimeStep.C:39
  Real getValue() const override { return _critical_dt; }

protected:
  ExprKernel _norm;
  Real _critical_dt = 0;
};

// src/postprocessors/TensorInterfaceVelocityPostprocessor.C:41-65: max over the grid of |du/dt / grad u|
class TensorInterfaceVelocityPostprocessor : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Compute the integral over a buffer");
    params.addParam<Real>("gradient_threshold", 1e-3, "Ignore cells with a gradient component magnitude below this threshold.");
    return params;
  }
  explicit TensorInterfaceVelocityPostprocessor(const InputParameters &p) : TensorPostprocessor(p), _u_old(_tensor_problem.getBufferOld(_buffer_name, 1)) {
    static const char *const k[3] = {"kx", "ky", "kz"};
    for (int d = 0; d < 3; ++d) _grad[d].configure(std::string("ubar*") + k[d] + "*i", {"ubar"}, {}, {}, {}, true, MRL_EXPAND_NONE);
    // `t` carries dt; the threshold is the literal of the reference code (:55), not the parameter
    _first.configure("v := if(abs(g) > 1e-3, ((u - uo)/t)/g, 0); v*v", {"u", "uo", "g"}, {}, {}, {}, true, MRL_EXPAND_NONE);
    _next.configure("v := if(abs(g) > 1e-3, ((u - uo)/t)/g, 0); acc + v*v", {"u", "uo", "g", "acc"}, {}, {}, {}, true, MRL_EXPAND_NONE);
  }
  void execute() override {
    if (_u_old.empty() || !_u_old[0].defined()) {
      _velocity = 0.0;
      return;
    }
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    const Tensor ubar = _domain.fft(_u);
    const Real dt = _tensor_problem.dt();
    Tensor vsq;
    for (unsigned int d = 0; d < _domain.getDim(); ++d) {
      const Tensor g = _domain.ifft(_grad[d].eval(_domain, {&ubar}, 0.0));
      vsq = d == 0 ? _first.eval(_domain, {&_u, &_u_old[0], &g}, dt) : _next.eval(_domain, {&_u, &_u_old[0], &g, &vsq}, dt);
    }
    _velocity = std::sqrt(_domain.reduce(MRL_MAX, vsq));
  }
  void finalize() override { gather(_domain, _velocity, Comm::MAX); }
  Real getValue() const override { return _velocity; }

protected:
  const std::vector<Tensor> &_u_old;
  ExprKernel _grad[3], _first, _next;
  Real _velocity = 0;
};

// src/vectorpostprocessors/TensorHistogram.C:31-84 ([VectorPostprocessors]): counts of the buffer values in `bins`
// equal bins between min and max.  The reference evaluates at::native::histogramdd on the CPU copy of the buffer
// (there is no CUDA histogramdd); so does this class: half-open bins [e_i, e_i+1), the last one closed, edges
// = linspace(min, max, bins + 1) with ATen's symmetric evaluation.
class TensorHistogram : public TensorVectorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Compute a histogram of the given tensor.");
    params.addRequiredParam<Real>("min", "Lower bound of the histogram.");
    params.addRequiredParam<Real>("max", "Upper bound of the histogram.");
    params.addRequiredParam<std::size_t>("bins", "Number of histogram bins.");
    return params;
  }
  explicit TensorHistogram(const InputParameters &p)
    : TensorVectorPostprocessor(p), _min(getParam<Real>("min")), _max(getParam<Real>("max")), _bins(getParam<std::size_t>("bins")) {
    if (_bins == 0) paramError("bins", "bins>0");
    if (_min > _max) paramError("min", "max must be greater than min");
    const std::size_t steps = _bins + 1;
    const Real step = (_max - _min) / Real(steps - 1);
    _edges.resize(steps);
    for (std::size_t i = 0; i < steps; ++i) _edges[i] = i < steps / 2 ? _min + step * Real(i) : _max - step * Real(steps - 1 - i);
    auto &bin = _vectors["bin"];
    bin.resize(_bins);
    _vectors["count"].assign(_bins, 0.0);
    const Real w = (_max - _min) / Real(_bins);
    for (std::size_t i = 0; i < _bins; ++i) bin[i] = _min + w / 2.0 + w * Real(i);
  }
  void execute() override {
    if (!_u.defined()) mooseError("buffer '", _buffer_name, "' is not defined");
    const std::vector<double> host = _domain.toHost(_u);
    auto &count = _vectors["count"];
    count.assign(_bins, 0.0);
    for (double v : host) {
      if (!(v >= _edges.front() && v <= _edges.back())) continue;
      std::size_t pos = std::size_t(std::upper_bound(_edges.begin(), _edges.end(), v) - _edges.begin());
      pos = pos == 0 ? 0 : pos - 1;
      if (pos >= _bins) pos = _bins - 1;  // v == max belongs to the last bin
      count[pos] += 1.0;
    }
  }
  void finalize() override {
    auto &count = _vectors["count"];
    _domain.comm().allreduce(count.data(), count.size(), Comm::SUM);
  }

protected:
  const Real _min, _max;
  const std::size_t _bins;
  std::vector<Real> _edges;
};

// src/postprocessors/ReciprocalIntegral.C:31-52: Re(ubar[0,0,0]) / #cells * volume
class ReciprocalIntegral : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Extract the zero k-vector value (corresponding to the integral).");
    return params;
  }
  using TensorPostprocessor::TensorPostprocessor;
  void execute() override {
    if (!_u.defined() || !_u.is_complex()) mooseError("buffer '", _buffer_name, "' must be a defined reciprocal-space (complex) buffer");
    double z[2] = {0, 0};
    if (_domain.realBytes() == 8) {
      _domain.check(mrl_download(_domain.context(), z, _u.data_ptr(), sizeof z), "mrl_download");
      _domain.synchronize();
    } else {
      float zf[2] = {0, 0};
      _domain.check(mrl_download(_domain.context(), zf, _u.data_ptr(), sizeof zf), "mrl_download");
      _domain.synchronize();
      z[0] = zf[0];
    }
    _integral = z[0] / Real(_domain.getNumberOfCells()) * _domain.getVolume();
    if (_domain.rank() != 0) _integral = 0.0;  // the zero wavevector is in rank 0's x slab
  }
  void finalize() override { gather(_domain, _integral, Comm::SUM); }
  Real getValue() const override { return _integral; }

protected:
  Real _integral = 0;
};

// src/postprocessors/ComputeGroupExecutionCount.C: computeBuffer() calls issued to a ComputeGroup
class ComputeGroupExecutionCount : public TensorPostprocessor {
public:
  static InputParameters validParams() {
    InputParameters params = TensorPostprocessor::validParams();
    params.addClassDescription("Return the number of computeBuffer() calls issued to the given compute group object.");
    params.addParam<TensorComputeName>("compute_group", "root", "ComputeGroup TensorCompute object to get execution count from.");
    // not a buffer postprocessor in the reference (GeneralPostprocessor): the base parameter is unused
    params.addParam<TensorInputBufferName>("buffer", "_compute_group_execution_count_unused", "unused");
    return params;
  }
  explicit ComputeGroupExecutionCount(const InputParameters &p) : TensorPostprocessor(p), _group_name(getParam<TensorComputeName>("compute_group")) {}
  void execute() override {}
  Real getValue() const override {
    for (const auto &cmp : _tensor_problem.getComputes())
      if (cmp->name() == _group_name) {
        const auto *g = dynamic_cast<const ComputeGroup *>(cmp.get());
        if (!g) mooseError("'", _group_name, "' is not a ComputeGroup");
        return Real(g->computeCount());
      }
    mooseError("compute_group '", _group_name, "' not found");
    return 0;
  }

protected:
  const std::string _group_name;
};

}  // namespace

registerMooseObject("MarlinApp", TensorHistogram);
registerMooseObject("MarlinApp", TensorInterfaceVelocityPostprocessor);
registerMooseObject("MarlinApp", ReciprocalIntegral);
registerMooseObject("MarlinApp", ComputeGroupExecutionCount);
registerMooseObject("MarlinApp", TensorAveragePostprocessor);
registerMooseObject("MarlinApp", TensorIntegralPostprocessor);
registerMooseObject("MarlinApp", TensorExtremeValuePostprocessor);
registerMooseObject("MarlinApp", TensorIntegralChangePostprocessor);
registerMooseObject("MarlinApp", SemiImplicitCriticalTimeStep);
