// [Postprocessors] acting on tensor buffers: scalar reductions computed on the device.
// Host mirror of src/postprocessors/TensorPostprocessor.C:14-34 (base, `buffer` parameter),
// TensorAveragePostprocessor.C:33-53, TensorIntegralPostprocessor.C:29-45,
// TensorExtremeValuePostprocessor.C:29-50, TensorIntegralChangePostprocessor.C:24-58,
// SemiImplicitCriticalTimeStep.C:30-50.  The reference syncs with `.cpu().item()`; here each value
// is one device reduction (mrl_reduce) whose result is read back.
#pragma once
#include "TensorProblem.h"

class TensorPostprocessor : public MooseObject {
public:
  static InputParameters validParams();
  explicit TensorPostprocessor(const InputParameters &parameters);
  virtual void initialize() {}
  virtual void execute() = 0;
  virtual void finalize() {}
  virtual Real getValue() const = 0;
  int executeOn() const { return _execute_on; }
  const std::string &bufferName() const { return _buffer_name; }

protected:
  TensorProblem &_tensor_problem;
  const DomainAction &_domain;
  const std::string _buffer_name;
  TensorBufferBase &_buffer_base;
  const marlin::Tensor &_u;
  int _execute_on;
};

// [VectorPostprocessors] acting on tensor buffers (src/vectorpostprocessors/TensorVectorPostprocessor.C): named
// vectors instead of one value; the driver writes them to <file_base>_<name>_<step>.csv like MOOSE's CSV output.
class TensorVectorPostprocessor : public TensorPostprocessor {
public:
  using TensorPostprocessor::TensorPostprocessor;
  Real getValue() const override { return 0.0; }
  const std::map<std::string, std::vector<Real>> &vectors() const { return _vectors; }

protected:
  std::map<std::string, std::vector<Real>> _vectors;
};
