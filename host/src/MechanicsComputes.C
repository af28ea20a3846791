// de Geus finite-strain FFT mechanics as MOOSE-style objects on top of mrl_mech_*.
//   FFTMechanics              src/tensor_computes/FFTMechanics.C:20-163
//   HyperElasticIsotropic     src/tensor_computes/HyperElasticIsotropic.C:14-52
//   RankTwoIdentity           src/tensor_computes/RankTwoIdentity.C:14-33
//   MacroscopicShearTensor    test/src/tensor_computes/MacroscopicShearTensor.C:15-41 (test fixture)
//   PhaseMechanicsTest        test/src/tensor_computes/PhaseMechanicsTest.C:15-50     (test fixture)
// Rank-two fields are dim x dim and stored component major on the device ([D*D][nx][ny][nz],
// c = D i + j; see include/marlin_b200.h); Ghat4, C4 and the tangent K4 are never materialised.
#include "MechanicsComputes.h"

#include <cstring>

using marlin::Space;
using marlin::Tensor;

// ------------------------------------------------------------------------------- RankTwoIdentity
registerMooseObject("MarlinApp", RankTwoIdentity);

InputParameters RankTwoIdentity::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Second order identity tensor field.");
  return params;
}
RankTwoIdentity::RankTwoIdentity(const InputParameters &parameters) : TensorOperator<>(parameters) {}

void RankTwoIdentity::computeBuffer() {
  if (_dim != 3 && _dim != 2) mooseError("the CUDA mechanics path is 2-D or 3-D");
  const size_t n = size_t(_domain.getNumberOfLocalCells());
  const int D = (int)_dim, nc = D * D;
  std::vector<double> host(nc * n, 0.0);
  for (int i = 0; i < D; ++i) std::fill(host.begin() + (D * i + i) * n, host.begin() + (D * i + i + 1) * n, 1.0);
  _u = _domain.fromHost(host, Space::REAL, false, nc);
}

// ------------------------------------------------------------------------ MacroscopicShearTensor
registerMooseObject("MarlinApp", MacroscopicShearTensor);

InputParameters MacroscopicShearTensor::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Macroscopic shear deformation gradient increment (test object).");
  params.addParam<TensorInputBufferName>("F", "F", "Deformation gradient tensor.");
  return params;
}
MacroscopicShearTensor::MacroscopicShearTensor(const InputParameters &parameters) : TensorOperator<>(parameters), _tF(getInputBuffer("F")) {}

void MacroscopicShearTensor::computeBuffer() {
  const int D = (int)_dim, nc = D * D;
  if (!_tF.defined() || _tF.ncomp() != nc) mooseError("F must be an initialised rank-two field");
  // I + t e0 e1^T - <F>, one device reduction per component (DomainAction::average, src/actions/DomainAction.C:1570-1574)
  std::vector<double> applied(nc);
  const size_t stride = size_t(_tF.count()) * _domain.realBytes();
  for (int c = 0; c < nc; ++c) {
    double s = 0;
    checkC(mrl_reduce(_domain.context(), MRL_SUM, static_cast<const char *>(_tF.data_ptr()) + c * stride, _tF.count(), &s), "mrl_reduce");
    _domain.comm().allreduce(&s, 1, Comm::SUM);
    applied[c] = ((c / D == c % D) ? 1.0 : 0.0) - s / Real(_domain.getNumberOfCells());
  }
  applied[1] += _time;
  _u = _domain.fromHost(applied, Space::SCALAR, false, nc);
}

// ---------------------------------------------------------------------------- PhaseMechanicsTest
registerMooseObject("MarlinApp", PhaseMechanicsTest);

InputParameters PhaseMechanicsTest::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Inclusion indicator of the de Geus example (test object).");
  return params;
}
PhaseMechanicsTest::PhaseMechanicsTest(const InputParameters &parameters) : TensorOperator<>(parameters) {}

void PhaseMechanicsTest::computeBuffer() {
  const auto &n = _domain.getGridSize();
  const int64_t s = _dim == 2 ? 30 : 9;
  if (_dim != 2 && _dim != 3) mooseError("Unsupported problem dimension");
  // this rank's part [b, e) of the grid (the whole grid in serial runs)
  std::array<int64_t, 3> b, e;
  _domain.getLocalBounds(_domain.rank(), b, e);
  const int64_t ly = e[1] - b[1], lz = _dim == 3 ? e[2] - b[2] : 1;
  std::vector<double> host(size_t(_domain.getNumberOfLocalCells()), 0.0);
  // phase[-s:, :s, -s:] = 1 (python slicing semantics: clipped to the grid)
  for (int64_t i = std::max<int64_t>(0, n[0] - s); i < n[0]; ++i)
    for (int64_t j = std::max<int64_t>(0, b[1]); j < std::min<int64_t>(std::min<int64_t>(s, n[1]), e[1]); ++j) {
      if (_dim == 2)
        host[size_t(i * ly + (j - b[1]))] = 1.0;
      else
        for (int64_t k = std::max<int64_t>(std::max<int64_t>(0, n[2] - s), b[2]); k < std::min<int64_t>(n[2], e[2]); ++k)
          host[size_t((i * ly + (j - b[1])) * lz + (k - b[2]))] = 1.0;
    }
  _u = _domain.fromHost(host, Space::REAL, false, 1);
}

// ------------------------------------------------------------------------------------ MechPlanHolder
MechPlanHolder::~MechPlanHolder() { reset(); }
void MechPlanHolder::reset() {
  if (_plan) mrl_mech_plan_destroy(_plan);
  _plan = nullptr;
}
mrl_mech_plan *MechPlanHolder::get(const DomainAction &domain, const mrl_mech_desc &desc, const Tensor &K, const Tensor &mu) {
  if (!K.defined() || !mu.defined()) ::mooseError("mechanics: the K and mu fields must be initialised");
  if (_plan && _K == K.data_ptr() && _mu == mu.data_ptr()) return _plan;
  reset();
  if (mrl_mech_plan_create(domain.context(), &desc, K.data_ptr(), mu.data_ptr(), &_plan) != MRL_OK) ::mooseError("marlin_b200: ", mrl_last_error());
  if (domain.dist()) {
    // FFT_SLAB / FFT_PENCIL: transforms through the decomposed path, inner products summed over the ranks
    auto sum = [](void *user, double *v, int count) -> int {
      try {
        static_cast<Comm *>(user)->allreduce(v, size_t(count), Comm::SUM);
      } catch (const std::exception &) {
        return 1;
      }
      return 0;
    };
    if (mrl_mech_plan_set_dist(_plan, domain.dist(), sum, &domain.comm()) != MRL_OK) ::mooseError("marlin_b200: ", mrl_last_error());
  }
  _K = K.data_ptr();
  _mu = mu.data_ptr();
  return _plan;
}

// ------------------------------------------------------------------------- HyperElasticIsotropic
registerMooseObject("MarlinApp", HyperElasticIsotropic);

InputParameters HyperElasticIsotropic::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Hyperelastic isotropic constitutive model.");
  params.addRequiredParam<TensorInputBufferName>("F", "Deformation gradient tensor");
  params.addRequiredParam<TensorInputBufferName>("mu", "Shear modulus");
  params.addRequiredParam<TensorInputBufferName>("K", "Bulk modulus");
  params.addParam<TensorOutputBufferName>("tangent_operator", "dstressdstrain", "Stiffness tensor");
  return params;
}

HyperElasticIsotropic::HyperElasticIsotropic(const InputParameters &parameters)
  : TensorOperator<>(parameters), _tF(getInputBuffer("F")), _tmu(getInputBuffer("mu")), _tK(getInputBuffer("K")) {
  // the tangent is applied in closed form from (F, K, mu) inside the CG operator; the buffer name is
  // still announced so that dependency resolution sees the same graph as the reference
  _supplied_buffers.insert(getParam<TensorOutputBufferName>("tangent_operator"));
}

void HyperElasticIsotropic::computeBuffer() {
  mrl_mech_desc d;
  std::memset(&d, 0, sizeof d);
  d.l_tol = 1e-2;
  d.nl_rel_tol = 1e-5;
  d.nl_abs_tol = 1e-8;
  d.nl_max_its = 100;
  mrl_mech_plan *plan = _plan.get(_domain, d, _tK, _tmu);
  Tensor P = _domain.empty(Space::REAL, false, int(_dim * _dim));
  checkC(mrl_mech_constitutive(plan, _tF.data_ptr(), P.data_ptr()), "mrl_mech_constitutive");
  _u = P;
}

// ---------------------------------------------------------------------------------- FFTMechanics
registerMooseObject("MarlinApp", FFTMechanics);

InputParameters FFTMechanics::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("deGeus variational mechanics solve. Updates the coupled buffer holding the deformation gradient tensor.");
  params.addRequiredParam<TensorInputBufferName>("K", "Bulk modulus");
  params.addParam<TensorInputBufferName>("mu", "Shear modulus");
  params.addParam<Real>("l_tol", 1e-2, "Linear congugate gradient solve tolerance");
  params.addParam<unsigned int>("l_max_its", "Maximum number of congugate gradient solve iterations");
  params.addParam<Real>("nl_rel_tol", 1e-5, "Nonlinear solve absolute tolerance");
  params.addParam<Real>("nl_abs_tol", 1e-8, "Nonlinear solve relative tolerance");
  params.addParam<unsigned int>("nl_max_its", 100, "Maximum number of nonlinear solve iterations");
  params.addParam<TensorInputBufferName>("stress", "stress", "Computed stress");
  params.addParam<TensorInputBufferName>("tangent_operator", "dstressdstrain", "Tangent operator");
  params.addRequiredParam<TensorComputeName>("constitutive_model", "Tensor compute for the constitutive model (computes stress from displacement gradeint tensor)");
  params.addParam<TensorInputBufferName>("applied_macroscopic_strain", "Applied macroscopic strain");
  params.addParam<TensorInputBufferName>("F", "F", "Deformation gradient tensor.");
  params.addParam<bool>("verbose", false, "Print non-linear residuals.");
  return params;
}

FFTMechanics::FFTMechanics(const InputParameters &parameters)
  : TensorOperator<>(parameters),
    _tK(getInputBuffer("K")),
    _tmu(getInputBuffer("mu")),
    _tF(getInputBuffer("F")),
    _tP(getInputBuffer("stress")),
    _constitutive_model(getCompute("constitutive_model")),
    _applied_macroscopic_strain(isParamValid("applied_macroscopic_strain") ? &getInputBuffer("applied_macroscopic_strain") : nullptr),
    _verbose(getParam<bool>("verbose")) {
  getInputBuffer("tangent_operator");
  std::memset(&_desc, 0, sizeof _desc);
  _desc.l_tol = getParam<Real>("l_tol");
  _desc.l_max_its = isParamValid("l_max_its") ? (int64_t)getParam<unsigned int>("l_max_its") : 0;  // <= 0: number of cells
  _desc.nl_rel_tol = getParam<Real>("nl_rel_tol");
  _desc.nl_abs_tol = getParam<Real>("nl_abs_tol");
  _desc.nl_max_its = (int)getParam<unsigned int>("nl_max_its");
  if (!dynamic_cast<HyperElasticIsotropic *>(&_constitutive_model))
    paramError("constitutive_model", "the CUDA mechanics path evaluates the HyperElasticIsotropic model in closed form; '", _constitutive_model.type(),
               "' is not supported.");
}

void FFTMechanics::check() {
  const auto stress_name = getParam<TensorOutputBufferName>("stress");
  if (!_constitutive_model.getSuppliedItems().count(stress_name)) paramError("constitutive_model", "does not provide stress tensor '", stress_name, "'.");
  // The solve evaluates the HyperElasticIsotropic law in closed form on its own iterate with THIS object's K / mu instead of
  // calling _constitutive_model.computeBuffer() (FFTMechanics.C:114-116, :139-141).  That is the same computation only if the
  // model block reads the iterate (this object's output buffer, e.g. F = Fnew in mech3d.i) and the same K / mu buffers.
  const auto iterate = getParam<TensorOutputBufferName>("buffer");
  const auto model_F = _constitutive_model.getParam<TensorInputBufferName>("F");
  if (model_F != iterate)
    paramError("constitutive_model", "'", _constitutive_model.name(), "' evaluates the stress from '", model_F, "' but FFTMechanics iterates on '", iterate,
               "'; the CUDA mechanics path needs the model to read the iterate.");
  for (const char *name : {"K", "mu"}) {
    const auto mine = getParam<TensorInputBufferName>(name), theirs = _constitutive_model.getParam<TensorInputBufferName>(name);
    if (mine != theirs)
      paramError(name, "the constitutive model '", _constitutive_model.name(), "' reads '", theirs, "' for ", name, " but FFTMechanics reads '", mine,
                 "'; the CUDA mechanics path needs both blocks to name the same buffers.");
  }
}

void FFTMechanics::computeBuffer() {
  mrl_mech_plan *plan = _plan.get(_domain, _desc, _tK, _tmu);
  // _u = _tF (+ applied strain + Newton increments): the solve updates its F argument in place
  Tensor F = _domain.clone(_tF);
  const int nc = int(_dim * _dim);
  if (_tF.ncomp() != nc) mooseError("F must be a ", _dim, "x", _dim, " tensor field");
  Tensor P = _domain.empty(Space::REAL, false, nc);
  std::vector<double> applied;
  if (_applied_macroscopic_strain) {
    if (_applied_macroscopic_strain->numel() != nc) mooseError("applied_macroscopic_strain must be a ", _dim, "x", _dim, " tensor");
    applied = _domain.toHost(*_applied_macroscopic_strain);
  }
  std::memset(&_stats, 0, sizeof _stats);
  const int rc = mrl_mech_solve(plan, F.data_ptr(), applied.empty() ? nullptr : applied.data(), P.data_ptr(), &_stats);
  if (rc != MRL_OK) {
    const std::string why = mrl_last_error();
    if (why.find("nonlinear") != std::string::npos) paramError("nl_max_its", "Exceeded the maximum number of nonlinear iterations without converging.");
    mooseError("marlin_b200: ", why);
  }
  if (_verbose) std::cerr << "|R|=" << _stats.final_anorm << "\t|R/R0|=" << _stats.final_rnorm << '\n';
  _u = F;
  // the stress of the final state is what the constitutive model's last evaluation leaves behind
  _tensor_problem.getBuffer(getParam<TensorOutputBufferName>("stress")) = P;
}

// ------------------------------------------------------------------------- ComputeVonMisesStress
registerMooseObject("MarlinApp", ComputeVonMisesStress);

InputParameters ComputeVonMisesStress::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Compute vonMises stress.");
  params.addParam<TensorInputBufferName>("stress", "stress", "Stress tensor.");
  return params;
}
ComputeVonMisesStress::ComputeVonMisesStress(const InputParameters &parameters) : TensorOperator<>(parameters), _stress(getInputBuffer("stress")) {}

void ComputeVonMisesStress::computeBuffer() {
  if (!_stress.defined()) return;  // ComputeVonMisesStress.C:33-34
  if (_dim != 2 && _dim != 3) mooseError("Unsupported problem dimension ", _dim);
  if (_stress.ncomp() != int(_dim * _dim)) mooseError("stress must be a ", _dim, "x", _dim, " tensor field");
  Tensor out = _domain.empty(Space::REAL, false, 1);
  checkC(mrl_von_mises(_domain.context(), _stress.data_ptr(), out.data_ptr()), "mrl_von_mises");
  _u = out;
}

// -------------------------------------------------------------------------- ComputeDisplacements
registerMooseObject("MarlinApp", ComputeDisplacements);

InputParameters ComputeDisplacements::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("Compute updated displacements from the deformation gradient tensor.");
  params.addRequiredParam<TensorInputBufferName>("F", "Deformation gradient tensor.");
  return params;
}
ComputeDisplacements::ComputeDisplacements(const InputParameters &parameters)
  : TensorOperator<>(parameters), _deformation_gradient_tensor(getInputBuffer("F")) {}

void ComputeDisplacements::computeBuffer() {
  const Tensor &F = _deformation_gradient_tensor;
  if (!F.defined()) return;  // ComputeDisplacements.C:56-58
  if (_dim != 2 && _dim != 3) mooseError("Unsupported problem dimension");
  if (F.ncomp() != int(_dim * _dim)) mooseError("Value dimensions of the deformation gradient tensor to not match the problem dimension");
  // (n+1)^dim nodal field, component major
  Tensor out = _domain.empty(Space::NODAL, false, int(_dim));
  checkC(mrl_displacements(_domain.context(), F.data_ptr(), out.data_ptr()), "mrl_displacements");
  _u = out;
}

// ---------------------------------------------------------------------- FFTQuasistaticElasticity
registerMooseObject("MarlinApp", FFTQuasistaticElasticity);

InputParameters FFTQuasistaticElasticity::validParams() {
  InputParameters params = TensorOperatorBase::validParams();
  params.addClassDescription("FFT based monolithic homogeneous quasistatic elasticity solve.");
  params.addParam<std::vector<TensorOutputBufferName>>("displacements", "Displacements");
  params.addParam<TensorInputBufferName>("cbar", "FFT of concentration buffer");
  params.addRequiredParam<Real>("mu", "Lame mu");
  params.addRequiredParam<Real>("lambda", "Lame lambda");
  params.addRequiredParam<Real>("e0", "volumetric eigenstrain");
  return params;
}

FFTQuasistaticElasticity::FFTQuasistaticElasticity(const InputParameters &parameters)
  : TensorOperatorBase(parameters), _mu(getParam<Real>("mu")), _lambda(getParam<Real>("lambda")), _e0(getParam<Real>("e0")), _cbar(getInputBuffer("cbar")) {
  for (const auto &name : getParam<std::vector<TensorOutputBufferName>>("displacements")) _displacements.push_back(&getOutputBufferByName(name));
  if (_domain.getDim() != _displacements.size()) paramError("displacements", "Need one displacement variable per mesh dimension");
  if (_dim != 3) paramError("displacements", "FFTQuasistaticElasticity is written for three dimensions");
  // FFTQuasistaticElasticity.C:62-91.  With k_d = 2 pi i * (reciprocal axis) every matrix entry
  // A_ab = coef * k_a k_b is REAL (-(2 pi)^2 coef axis_a axis_b); the diagonal is 1 at k = 0.  The solve
  // goes through mrl_coupled_solve, which solves (I - dt L) x = b: L = I - A with dt = 1.
  const std::vector<std::string> cn = {"mu", "lambda", "e0"};
  const std::vector<double> cv = {_mu, _lambda, _e0};
  const char *diag[3] = {"(2*mu + lambda)*kx*kx + mu*ky*ky + mu*kz*kz", "(2*mu + lambda)*ky*ky + mu*kx*kx + mu*kz*kz",
                         "(2*mu + lambda)*kz*kz + mu*kx*kx + mu*ky*ky"};
  const char *off[3] = {"kx*ky", "kx*kz", "ky*kz"};
  const char *axis[3] = {"kx", "ky", "kz"};
  for (int a = 0; a < 3; ++a) {
    _coef[a].configure(std::string("if(k2 == 0, 0, 1 + 4*pi*pi*(") + diag[a] + "))", {}, {}, cn, cv, true, MRL_EXPAND_RECIPROCAL);
    _coef[3 + a].configure(std::string("4*pi*pi*(lambda + mu)*") + off[a], {}, {}, cn, cv, true, MRL_EXPAND_RECIPROCAL);
    _rhs[a].configure(std::string("if(k2 == 0, 0, (2*pi*i*") + axis[a] + ")*(2*e0*cbar*(3*lambda + mu)))", {"cbar"}, {}, cn, cv, true, MRL_EXPAND_NONE);
  }
}

void FFTQuasistaticElasticity::computeBuffer() {
  if (!_L[0].defined())
    for (int a = 0; a < 6; ++a) _L[a] = _coef[a].eval(_domain, {}, _time);
  Tensor b[3], x[3];
  for (int a = 0; a < 3; ++a) {
    b[a] = _rhs[a].eval(_domain, {&_cbar}, _time);
    x[a] = _domain.empty(Space::RECIPROCAL, true, 1);
  }
  // symmetric matrix: xx xy xz / xy yy yz / xz yz zz
  const void *L[9] = {_L[0].data_ptr(), _L[3].data_ptr(), _L[4].data_ptr(), _L[3].data_ptr(), _L[1].data_ptr(),
                      _L[5].data_ptr(), _L[4].data_ptr(), _L[5].data_ptr(), _L[2].data_ptr()};
  const void *rhs[3] = {b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr()};
  void *out[3] = {x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr()};
  checkC(mrl_coupled_solve(_domain.context(), 3, L, rhs, out, 1.0, 0), "mrl_coupled_solve");
  for (int a = 0; a < 3; ++a) *_displacements[a] = _domain.ifft(x[a]);
}

// ------------------------------------------------------------------- FFTElasticChemicalPotential
registerMooseObject("MarlinApp", FFTElasticChemicalPotential);

InputParameters FFTElasticChemicalPotential::validParams() {
  InputParameters params = TensorOperator<>::validParams();
  params.addClassDescription("FFT based elastic strain energy chemical potential solve.");
  params.addParam<std::vector<TensorInputBufferName>>("displacements", "Displacements");
  params.addParam<TensorInputBufferName>("cbar", "FFT of concentration buffer");
  params.addRequiredParam<Real>("mu", "Lame mu");
  params.addRequiredParam<Real>("lambda", "Lame lambda");
  params.addRequiredParam<Real>("e0", "volumetric eigenstrain");
  return params;
}

FFTElasticChemicalPotential::FFTElasticChemicalPotential(const InputParameters &parameters) : TensorOperator<>(parameters), _cbar(getInputBuffer("cbar")) {
  for (const auto &name : getParam<std::vector<TensorInputBufferName>>("displacements")) _displacements.push_back(&getInputBufferByName(name));
  if (_domain.getDim() != _displacements.size()) paramError("displacements", "Need one displacement variable per mesh dimension");
  if (_dim != 3) paramError("displacements", "FFTElasticChemicalPotential is written for three dimensions");
  // FFTElasticChemicalPotential.C:49-60
  _kernel.configure("-e0*(e0*(9*lambda*cbar + mu*6*cbar) - (2*mu + 3*lambda)*(2*pi*i)*(kx*ux + ky*uy + kz*uz))", {"cbar", "ux", "uy", "uz"}, {},
                    {"mu", "lambda", "e0"}, {getParam<Real>("mu"), getParam<Real>("lambda"), getParam<Real>("e0")}, true, MRL_EXPAND_NONE);
}

void FFTElasticChemicalPotential::computeBuffer() {
  const Tensor ux = _domain.fft(*_displacements[0]), uy = _domain.fft(*_displacements[1]), uz = _domain.fft(*_displacements[2]);
  _u = _kernel.eval(_domain, {&_cbar, &ux, &uy, &uz}, _time);
}
