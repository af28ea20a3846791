// TensorOperatorBase / TensorOperator<T> / ComputeGroup - the operator plug-in interface.
// Same virtuals, parameter names and buffer-request protocol as the reference
// (include/tensor_computes/TensorOperatorBase.h:36-63,89-99; include/tensor_computes/TensorOperator.h:20-44;
//  src/tensor_computes/ComputeGroup.C:38-146), with marlin::Tensor in place of torch::Tensor.
#pragma once
#include <set>
#include <string>
#include <vector>

#include "TensorProblem.h"

class TensorOperatorBase : public MooseObject, public DependencyResolverInterface {
public:
  static InputParameters validParams();
  explicit TensorOperatorBase(const InputParameters &parameters);

  const std::set<std::string> &getRequestedItems() override { return _requested_buffers; }
  const std::set<std::string> &getSuppliedItems() override { return _supplied_buffers; }

  /// Helper to recursively update dependencies for grouped operators
  virtual void updateDependencies() {}
  /// perform the computation
  virtual void computeBuffer() = 0;
  /// perform the computation in real space (halo exchanging computes; not on the spectral path)
  virtual void realSpaceComputeBuffer();
  /// called after all objects have been constructed (before dependency resolution)
  virtual void init() {}
  /// called after all objects have been constructed (after dependency resolution)
  virtual void check() {}
  /// called if the simulation cell dimensions change
  virtual void gridChanged() {}
  /// the reference's JIT tracer is replaced by hand-fused kernels; kept for interface parity
  virtual bool supportsJIT() const { return true; }

  const marlin::Tensor &getInputBuffer(const std::string &param, unsigned int ghost_layers = 0);
  const marlin::Tensor &getInputBufferByName(const TensorInputBufferName &buffer_name, unsigned int ghost_layers = 0);
  marlin::Tensor &getOutputBuffer(const std::string &param);
  marlin::Tensor &getOutputBufferByName(const TensorOutputBufferName &buffer_name);
  TensorOperatorBase &getCompute(const std::string &param_name);
  TensorBufferBase &getBufferBase(const TensorInputBufferName &buffer_name) { return _tensor_problem.getBufferBase(buffer_name); }
  Real getConstant(const std::string &param) const { return _tensor_problem.getConstant(getParam<std::string>(param), _path + "/" + param); }

  std::set<std::string> _requested_buffers;
  std::set<std::string> _supplied_buffers;

  TensorProblem &_tensor_problem;
  const DomainAction &_domain;
  /// substep time
  const Real &_time;
  /// spatial dimension
  const unsigned int _dim;

protected:
  void checkC(int rc, const char *what) const;
};

template <typename T = marlin::Tensor>
class TensorOperator : public TensorOperatorBase {
public:
  static InputParameters validParams() {
    InputParameters params = TensorOperatorBase::validParams();
    params.addRequiredParam<TensorOutputBufferName>("buffer", "The buffer this compute is writing to");
    params.addClassDescription("TensorOperator object.");
    return params;
  }
  explicit TensorOperator(const InputParameters &parameters) : TensorOperatorBase(parameters), _u(getOutputBuffer("buffer")) {}

protected:
  /// output buffer
  T &_u;
};

class ComputeGroup : public TensorOperatorBase {
public:
  static InputParameters validParams();
  explicit ComputeGroup(const InputParameters &parameters);
  void init() override;
  void computeBuffer() override;
  void updateDependencies() override;
  const std::vector<std::shared_ptr<TensorOperatorBase>> &getComputes() const { return _computes; }
  unsigned int computeCount() const { return _compute_count; }
  // used by the problem to wrap all [Solve] computes when the solver names no root_compute
  void setComputes(std::vector<std::shared_ptr<TensorOperatorBase>> computes) { _computes = std::move(computes); }

protected:
  std::vector<std::shared_ptr<TensorOperatorBase>> _computes;
  bool _visited = false;
  unsigned int _compute_count = 0;
  std::vector<std::vector<std::string>> _checked_tensors;
};
