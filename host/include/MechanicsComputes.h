// de Geus FFT mechanics operator classes (see host/src/MechanicsComputes.C for reference citations).
#pragma once
#include "TensorComputes.h"

class RankTwoIdentity : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit RankTwoIdentity(const InputParameters &parameters);
  void computeBuffer() override;
};

class MacroscopicShearTensor : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit MacroscopicShearTensor(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_tF;
};

class PhaseMechanicsTest : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit PhaseMechanicsTest(const InputParameters &parameters);
  void computeBuffer() override;
};

// owns an mrl_mech_plan bound to one pair of (K, mu) fields
class MechPlanHolder {
public:
  ~MechPlanHolder();
  mrl_mech_plan *get(const DomainAction &domain, const mrl_mech_desc &desc, const marlin::Tensor &K, const marlin::Tensor &mu);
  void reset();

private:
  mrl_mech_plan *_plan = nullptr;
  const void *_K = nullptr, *_mu = nullptr;
};

class HyperElasticIsotropic : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit HyperElasticIsotropic(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_tF;
  const marlin::Tensor &_tmu;
  const marlin::Tensor &_tK;
  MechPlanHolder _plan;
};

class FFTMechanics : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit FFTMechanics(const InputParameters &parameters);
  void computeBuffer() override;
  void check() override;
  const mrl_mech_stats &stats() const { return _stats; }

protected:
  const marlin::Tensor &_tK;
  const marlin::Tensor &_tmu;
  const marlin::Tensor &_tF;
  const marlin::Tensor &_tP;
  TensorOperatorBase &_constitutive_model;
  const marlin::Tensor *const _applied_macroscopic_strain;
  const bool _verbose;
  mrl_mech_desc _desc;
  mrl_mech_stats _stats;
  MechPlanHolder _plan;
};

// src/tensor_computes/ComputeVonMisesStress.C
class ComputeVonMisesStress : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ComputeVonMisesStress(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_stress;
};

// src/tensor_computes/ComputeDisplacements.C
class ComputeDisplacements : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit ComputeDisplacements(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_deformation_gradient_tensor;
};

// src/tensor_computes/FFTQuasistaticElasticity.C: per-wavevector 3x3 solve for the displacements (3-D)
class FFTQuasistaticElasticity : public TensorOperatorBase {
public:
  static InputParameters validParams();
  explicit FFTQuasistaticElasticity(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const Real _mu, _lambda, _e0;
  const marlin::Tensor &_cbar;
  std::vector<marlin::Tensor *> _displacements;
  ExprKernel _coef[6], _rhs[3];  // L = I - A entries (xx yy zz xy xz yz), right-hand sides
  marlin::Tensor _L[6];
};

// src/tensor_computes/FFTElasticChemicalPotential.C
class FFTElasticChemicalPotential : public TensorOperator<> {
public:
  static InputParameters validParams();
  explicit FFTElasticChemicalPotential(const InputParameters &parameters);
  void computeBuffer() override;

protected:
  const marlin::Tensor &_cbar;
  std::vector<const marlin::Tensor *> _displacements;
  ExprKernel _kernel;
};
