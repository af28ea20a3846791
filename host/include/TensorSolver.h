// TensorSolver family: substep loop, split-operator variables, the concrete integrators.
// Host mirror of include/tensor_solver/{TensorSolver,SplitOperatorBase,ExplicitSolverBase,
// AdamsBashforthMoulton,ForwardEulerSolver,ETDRK4Solver}.h and their sources
// (src/tensor_solver/TensorSolver.C:15-110, SplitOperatorBase.C:14-64, ExplicitSolverBase.C,
//  AdamsBashforthMoulton.C:20-178, ForwardEulerSolver.C:29-38, ETDRK4Solver.C:29-115).
#pragma once
#include "TensorComputes.h"
#include <array>

#include "TensorOperatorBase.h"

class TensorSolver : public TensorOperatorBase {
public:
  static InputParameters validParams();
  explicit TensorSolver(const InputParameters &parameters);
  void computeBuffer() override;
  void updateDependencies() override;

protected:
  virtual void substep() = 0;
  const std::vector<marlin::Tensor> &getBufferOld(const std::string &param, unsigned int max_states);
  const std::vector<marlin::Tensor> &getBufferOldByName(const TensorInputBufferName &buffer_name, unsigned int max_states);
  void forwardBuffers();

  unsigned int _substeps;
  unsigned int _substep = 0;
  Real &_sub_dt;
  Real &_sub_time;
  const Real &_dt;
  const Real &_dt_old;
  const Real &_problem_time_old;
  std::shared_ptr<TensorOperatorBase> _compute;
  std::vector<std::pair<marlin::Tensor *, const marlin::Tensor *>> _forwarded_buffers;
};

class SplitOperatorBase : public TensorSolver {
public:
  static InputParameters validParams();
  explicit SplitOperatorBase(const InputParameters &parameters);

  struct Variable {
    marlin::Tensor &_buffer;
    const marlin::Tensor &_reciprocal_buffer;
    const marlin::Tensor *_linear_reciprocal;
    const marlin::Tensor &_nonlinear_reciprocal;
    const std::vector<marlin::Tensor> &_old_nonlinear_reciprocal;
    // names, used by the fused-plan pattern matcher
    std::string _buffer_name, _reciprocal_name, _linear_name, _nonlinear_name;
  };

protected:
  void getVariables(unsigned int history_size);
  std::vector<Variable> _variables;
};

class ExplicitSolverBase : public TensorSolver {
public:
  static InputParameters validParams();
  explicit ExplicitSolverBase(const InputParameters &parameters);
  struct Variable {
    marlin::Tensor &_buffer;
    const marlin::Tensor &_reciprocal_buffer;
    const marlin::Tensor &_time_derivative_reciprocal;
  };

protected:
  std::vector<Variable> _variables;
};

class ForwardEulerSolver : public ExplicitSolverBase {
public:
  static InputParameters validParams();
  explicit ForwardEulerSolver(const InputParameters &parameters);

protected:
  void substep() override;
};

struct mrl_split_plan;
struct mrl_slab_plan;
struct mrl_expr;

class AdamsBashforthMoulton : public SplitOperatorBase {
public:
  static InputParameters validParams();
  explicit AdamsBashforthMoulton(const InputParameters &parameters);
  ~AdamsBashforthMoulton() override;
  static constexpr std::size_t max_order = 5;
  void check() override;
  void computeBuffer() override;
  bool fused() const { return !_plans.empty(); }

protected:
  void substep() override;
  void fusedSubstep();
  void tryBuildFusedPlans();
  void decideFusion();
  bool _fusion_decided = false;

  std::size_t _predictor_order;
  std::size_t _corrector_order;
  std::size_t _corrector_steps;
  const bool _allow_fusion;
  const bool _allow_batching;

  // fused five-pass plans (one per solver variable) when the root compute matches the canonical
  // split-operator pattern; empty = generic operator-by-operator path
  struct FusedVariable {
    mrl_split_plan *plan = nullptr;
    mrl_slab_plan *slab = nullptr;  // [Domain] parallel_mode = FFT_SLAB: the multi-GPU plan with the exchanges fused into the passes
    mrl_expr *expr = nullptr;
    int stored = 0;  // old nonlinear terms currently held by the plan's ring
    std::string g_name;  // real-space nonlinearity buffer to materialise (observed), or empty
  };
  std::vector<FusedVariable> _plans;
  int _fused_last_step = -1;
  std::string _fusion_note;
};

// src/tensor_solver/AdamsBashforthMoultonCoupled.C: Adams-Bashforth-Moulton with a dense linear
// operator (off-diagonal reciprocal-space couplings), one N x N solve per wavevector.
class AdamsBashforthMoultonCoupled : public SplitOperatorBase {
public:
  static InputParameters validParams();
  explicit AdamsBashforthMoultonCoupled(const InputParameters &parameters);
  static constexpr std::size_t max_order = 5;

protected:
  void substep() override;
  void solveAndInvert(const std::vector<marlin::Tensor> &rhs);

  std::size_t _predictor_order;
  std::size_t _corrector_order;
  std::size_t _corrector_steps;
  const bool _assume_symmetric;
  std::vector<std::pair<unsigned int, unsigned int>> _L_offdiag_indices;
  std::vector<TensorInputBufferName> _L_offdiag_names;
  std::vector<const marlin::Tensor *> _L_offdiag_buffer;
};

// include/tensor_solver/IterativeTensorSolverInterface.h: what TensorSolveIterationAdaptiveDT reads
// include/tensor_predictor/TensorPredictor.h, src/tensor_predictor/TensorPredictor.C:14-39: forward-predicts a solver
// output buffer from its old states ([TensorSolver/Predictors/*])
class TensorPredictor : public MooseObject {
public:
  static InputParameters validParams();
  explicit TensorPredictor(const InputParameters &parameters);
  virtual void computeBuffer() = 0;
  virtual void gridChanged() {}

protected:
  TensorProblem &_tensor_problem;
  const DomainAction &_domain;
  const TensorOutputBufferName _u_name;
  marlin::Tensor &_u;
  const std::vector<marlin::Tensor> &_u_old;
};

// src/tensor_predictor/LinearTensorPredictor.C:11-38: u += scale * (u_old[0] - u_old[1])
class LinearTensorPredictor : public TensorPredictor {
public:
  static InputParameters validParams();
  explicit LinearTensorPredictor(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const Real _scale;
  ExprKernel _extrapolate;
};

class IterativeTensorSolverInterface {
public:
  virtual ~IterativeTensorSolverInterface() = default;
  const unsigned int &getIterations() const { return _iterations; }
  bool isConverged() const { return _is_converged; }
  // include/tensor_solver/IterativeTensorSolverInterface.h:27-32.  The reference's AddTensorPredictorAction builds the
  // predictor objects but leaves `solver.addPredictor()` commented out (src/actions/AddTensorPredictorAction.C:41), so
  // its applyPredictors() loops over an empty list; the stand-alone driver registers them only when asked to
  // ([TensorSolver] apply_predictors = true, an option of this build).
  void addPredictor(std::shared_ptr<TensorPredictor> p) { _predictors.push_back(std::move(p)); }
  void applyPredictors() {
    for (const auto &pred : _predictors) pred->computeBuffer();
  }

protected:
  unsigned int _iterations = 0;
  bool _is_converged = true;
  std::vector<std::shared_ptr<TensorPredictor>> _predictors;
};

// src/tensor_solver/SecantSolver.C: implicit Euler, secant iteration per wavevector
class SecantSolver : public SplitOperatorBase, public IterativeTensorSolverInterface {
public:
  static InputParameters validParams();
  explicit SecantSolver(const InputParameters &parameters);

protected:
  void substep() override;
  const unsigned int _max_iterations;
  const Real _relative_tolerance, _absolute_tolerance;
  const bool _verbose;
  const Real _damping, _dt_epsilon;
  // generated pointwise kernels, with / without a linear operator
  ExprKernel _r0[2], _start[2], _res[2], _update;
  Real complexNorm(const marlin::Tensor &t) const;
};

// src/tensor_solver/BroydenSolver.C: implicit Euler, Broyden update of a per-wavevector inverse Jacobian
class BroydenSolver : public SplitOperatorBase, public IterativeTensorSolverInterface {
public:
  static InputParameters validParams();
  explicit BroydenSolver(const InputParameters &parameters);

protected:
  void substep() override;
  Real stackedNorm(const std::vector<marlin::Tensor> &R) const;
  const unsigned int _max_iterations;
  const Real _relative_tolerance, _absolute_tolerance;
  const bool _verbose;
  const Real _eye_factor;
  marlin::Tensor _M;  // [n*n][reciprocal points] complex, component major; built at the first substep
  ExprKernel _res0[2], _res[2];
};

class ETDRK4Solver : public SplitOperatorBase {
public:
  static InputParameters validParams();
  explicit ETDRK4Solver(const InputParameters &parameters);
  ~ETDRK4Solver() override;

protected:
  void substep() override;
  mrl_expr *_e_stage_half = nullptr, *_e_stage_full = nullptr, *_e_final = nullptr;
  mrl_expr *kernel(mrl_expr *&slot, const char *expression, const std::vector<std::string> &names, const std::vector<int> &layouts);
};
