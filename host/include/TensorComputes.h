// Operator classes of the spectral path (see host/src/TensorComputes.C for reference citations).
#pragma once
#include <array>

#include "TensorOperatorBase.h"

// A runtime expression lowered to one CUDA kernel through mrl_expr_* (ParsedJITTensor's role,
// src/utils/ParsedJITTensor.C:22-156).  Compiled on first use, once the layouts of the input
// tensors are known; recompiled if they change.
class ExprKernel {
public:
  ExprKernel() = default;
  ~ExprKernel();
  ExprKernel(const ExprKernel &) = delete;
  ExprKernel &operator=(const ExprKernel &) = delete;
  void configure(const std::string &expression, std::vector<std::string> inputs, std::vector<std::string> derivatives, std::vector<std::string> constant_names,
                 std::vector<double> constant_values, bool extra_symbols, int expand);
  marlin::Tensor eval(const DomainAction &domain, const std::vector<const marlin::Tensor *> &inputs, double t);
  std::string simplified() const;  // string form after derivatives + simplification (host only)
  mrl_expr *handle() const { return _expr; }
  const std::string &expression() const { return _expression; }
  const std::vector<std::string> &inputs() const { return _inputs; }
  const std::vector<std::string> &derivatives() const { return _derivatives; }
  const std::vector<std::string> &constantNames() const { return _constant_names; }
  const std::vector<double> &constantValues() const { return _constant_values; }
  bool extraSymbols() const { return _extra; }
  static int layoutOf(const marlin::Tensor &t);

private:
  void reset();
  void fillDesc(mrl_expr_desc &d, std::vector<const char *> &in, std::vector<const char *> &der, std::vector<const char *> &cn, const std::vector<int> &layouts) const;
  std::string _expression;
  std::vector<std::string> _inputs, _derivatives, _constant_names;
  std::vector<double> _constant_values;
  bool _extra = false;
  int _expand = 0;
  std::vector<int> _layouts;
  mrl_expr *_expr = nullptr;
  int _space = 1, _is_complex = 0;
};

template <bool reciprocal>
class ConstantTensorTempl : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ConstantTensorTempl(const InputParameters &parameters);
  void computeBuffer() override;
};
using ConstantTensor = ConstantTensorTempl<false>;
using ConstantReciprocalTensor = ConstantTensorTempl<true>;

// torch::manual_seed(seed); torch::rand(n, CPU) * (max - min) + min, bit for bit (ATen mt19937)
void marlinTorchRand(std::vector<double> &out, size_t n, bool single, double min, double max, const int *seed);

class RandomTensor : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit RandomTensor(const InputParameters &parameters);
  void computeBuffer() override;
};

template <bool forward>
class PerformFFTTempl : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit PerformFFTTempl(const InputParameters &parameters);
  void computeBuffer() override;
  const marlin::Tensor &_input;
};
using ForwardFFT = PerformFFTTempl<true>;
using InverseFFT = PerformFFTTempl<false>;

template <int kind>
class ReciprocalLaplacianFactorTempl : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ReciprocalLaplacianFactorTempl(const InputParameters &parameters);
  void computeBuffer() override;
  Real factor() const { return _factor; }

protected:
  const Real _factor;
};
using ReciprocalLaplacianFactor = ReciprocalLaplacianFactorTempl<MRL_KFACTOR_LAPLACIAN>;
using ReciprocalLaplacianSquareFactor = ReciprocalLaplacianFactorTempl<MRL_KFACTOR_LAPLACIAN_SQUARE>;

class ParsedCompute : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ParsedCompute(const InputParameters &parameters);
  void computeBuffer() override;
  ExprKernel &kernel() { return _kernel; }
  const std::vector<const marlin::Tensor *> &inputTensors() const { return _params; }

protected:
  const bool _extra_symbols;
  ExprKernel _kernel;
  std::vector<const marlin::Tensor *> _params;
};

class FFTGradient : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit FFTGradient(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_input;
  const bool _input_is_reciprocal;
  const int _direction;
  ExprKernel _kernel;
};

class FFTGradientSquare : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit FFTGradientSquare(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_input;
  const bool _input_is_reciprocal;
  const Real _factor;
  std::array<ExprKernel, 3> _grad;
  ExprKernel _square;
};

class FFTSemiImplicit : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit FFTSemiImplicit(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const unsigned int _history_size;
  const Real &_sub_dt;
  const marlin::Tensor &_reciprocal_buffer;
  const marlin::Tensor &_linear_reciprocal;
  const marlin::Tensor &_non_linear_reciprocal;
  const std::vector<marlin::Tensor> &_old_reciprocal_buffer;
  const std::vector<marlin::Tensor> &_old_non_linear_reciprocal;
};

// src/tensor_computes/SwiftHohenbergLinear.C: r - alpha^2 (1 - k^2)^2
class SwiftHohenbergLinear : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit SwiftHohenbergLinear(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  ExprKernel _kernel;
};

// src/tensor_computes/SmoothRectangleCompute.C: box indicator blended between `inside` and `outside`
// (sharp, half-sine or tanh profile of the distance to the nearest face)
class SmoothRectangleCompute : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit SmoothRectangleCompute(const InputParameters &parameters);
  void computeBuffer() override;
  // the generated-kernel expression over x, y, z and the constants x1..z2, w, w2 = w/2, vin, vout, pi
  static std::string expression(unsigned int dim, Real w, const std::string &profile);

protected:
  ExprKernel _kernel;
};

// src/tensor_computes/MooseFunctionTensor.C: a MOOSE Function ([Functions], type ParsedFunction)
// sampled at the cell centres
class MooseFunctionTensor : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit MooseFunctionTensor(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  marlin::Tensor evaluate(const std::string &function, int depth);
  const std::string _function;
  std::vector<marlin::Tensor> _coords;
};

// src/tensor_computes/ReciprocalMatDiffusion.C: i k . fft(M grad mu) + fft(grad(psi)/psi . J)
class ReciprocalMatDiffusion : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ReciprocalMatDiffusion(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_chem_pot, &_M, &_psi;
  bool _update_psi = true;
  const bool _always_update_psi;
  marlin::Tensor _grad_psi_by_psi[3];
  ExprKernel _grad[3], _by_psi, _flux, _dot[3], _div[3];
};

// src/tensor_computes/ReciprocalAllenCahn.C: fft(where(psi > 0, -L dF/deta, 0))
class ReciprocalAllenCahn : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ReciprocalAllenCahn(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_dF_chem_deta, &_L, &_psi;
  ExprKernel _rate;
};

// src/tensor_computes/DeAliasingTensor.C: 2/3-rule (SHARP) or Hou-Li exponential de-aliasing filter
class DeAliasingTensor : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit DeAliasingTensor(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const bool _houli;
  const Real _p, _alpha;
  ExprKernel _kernel;
};
