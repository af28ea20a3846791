// TensorSolver, SplitOperatorBase, ExplicitSolverBase, ForwardEulerSolver, AdamsBashforthMoulton
// (+ alias SemiImplicitSolver), ETDRK4Solver.  Reference citations in include/TensorSolver.h.
#include "TensorSolver.h"

#include <cmath>
#include <cstring>

#include "TensorComputes.h"

using marlin::Space;
using marlin::Tensor;

// ========================================================================================== TensorSolver
InputParameters TensorSolver::validParams() {
  InputParameters params = TensorOperatorBase::validParams();
  params.addClassDescription("TensorSolver object.");
  params.registerBase("TensorSolver");
  params.addParam<TensorComputeName>("root_compute",
                                     "Primary compute object that updates the buffers. This is usually a ComputeGroup object. A ComputeGroup "
                                     "encompassing all computes will be generated automatically if the user does not provide this parameter.");
  params.addParam<unsigned int>("substeps", 1, "Solver substeps per time step.");
  params.addParam<std::vector<TensorOutputBufferName>>("forward_buffer", {},
                                                       "These buffers are updated with the corresponding buffers from forward_buffer_new. No integration is "
                                                       "performed. Buffer forwarding is used only to resolve cyclic dependencies.");
  params.addParam<std::vector<TensorInputBufferName>>("forward_buffer_new", {}, "New values to update `forward_buffer` with.");
  return params;
}

TensorSolver::TensorSolver(const InputParameters &parameters)
  : TensorOperatorBase(parameters),
    _substeps(getParam<unsigned int>("substeps")),
    _sub_dt(_tensor_problem.subDt()),
    _sub_time(_tensor_problem.subTime()),
    _dt(_tensor_problem.dt()),
    _dt_old(_tensor_problem.dtOld()),
    _problem_time_old(_tensor_problem.timeOld()) {
  for (const auto &[forward_buffer, forward_buffer_new] : getParam<TensorOutputBufferName, TensorInputBufferName>("forward_buffer", "forward_buffer_new"))
    _forwarded_buffers.emplace_back(&getOutputBufferByName(forward_buffer), &getInputBufferByName(forward_buffer_new));
}

const std::vector<Tensor> &TensorSolver::getBufferOld(const std::string &param, unsigned int max_states) {
  return getBufferOldByName(getParam<TensorInputBufferName>(param), max_states);
}
const std::vector<Tensor> &TensorSolver::getBufferOldByName(const TensorInputBufferName &buffer_name, unsigned int max_states) {
  return _tensor_problem.getBufferOld(buffer_name, max_states);
}

void TensorSolver::updateDependencies() {
  const auto root_name = getParam<TensorComputeName>("root_compute");
  for (const auto &cmp : _tensor_problem.getComputes())
    if (cmp->name() == root_name) {
      _compute = cmp;
      _compute->updateDependencies();
      return;
    }
  paramError("root_compute", "Compute object not found.");
}

void TensorSolver::forwardBuffers() {
  for (const auto &[forward_buffer, forward_buffer_new] : _forwarded_buffers) *forward_buffer = *forward_buffer_new;
}

// TensorSolver.C:93-110.  `_sub_time = _time` there assigns the problem's sub-time reference to
// itself (both names alias TensorProblem::_sub_time), so the sub-time continues from the value
// TensorProblem::execute set at TIMESTEP_BEGIN (timeOld).
void TensorSolver::computeBuffer() {
  _sub_dt = _dt / _substeps;
  for (_substep = 0; _substep < _substeps; _substep++) {
    substep();
    if (_substep < _substeps - 1) _tensor_problem.advanceState();
    _sub_time += _sub_dt;
  }
}

// ===================================================================================== SplitOperatorBase
InputParameters SplitOperatorBase::validParams() {
  InputParameters params = TensorSolver::validParams();
  params.addClassDescription("Base class for non-linear/linear operator splits.");
  params.addRequiredParam<std::vector<TensorOutputBufferName>>("buffer", "The buffer this solver is writing to");
  params.addRequiredParam<std::vector<TensorInputBufferName>>("reciprocal_buffer", "Buffer with the reciprocal of the integrated buffer");
  params.addRequiredParam<std::vector<TensorInputBufferName>>(
      "linear_reciprocal",
      "Buffer with the reciprocal of the linear prefactor (e.g. kappa*k^2). Either one buffer per nonlinear_reciprocal, or no buffer names, or `0` to skip "
      "linear reciprocal buffers for a given variable.");
  params.addRequiredParam<std::vector<TensorInputBufferName>>("nonlinear_reciprocal", "Buffer with the reciprocal of the non-linear contribution");
  return params;
}

SplitOperatorBase::SplitOperatorBase(const InputParameters &parameters) : TensorSolver(parameters) {}

void SplitOperatorBase::getVariables(unsigned int history_size) {
  auto buffers = getParam<std::vector<TensorOutputBufferName>>("buffer");
  auto reciprocal_buffers = getParam<std::vector<TensorInputBufferName>>("reciprocal_buffer");
  auto linear_reciprocals = getParam<std::vector<TensorInputBufferName>>("linear_reciprocal");
  auto nonlinear_reciprocals = getParam<std::vector<TensorInputBufferName>>("nonlinear_reciprocal");
  const auto n = buffers.size();
  if (linear_reciprocals.empty()) linear_reciprocals.assign(n, "0");
  if (reciprocal_buffers.size() != n || linear_reciprocals.size() != n || nonlinear_reciprocals.size() != n)
    paramError("buffer", "Must have the same number of entries as 'reciprocal_buffer', 'linear_reciprocal' and 'nonlinear_reciprocal'.");
  for (std::size_t i = 0; i < n; ++i)
    _variables.push_back(Variable{getOutputBufferByName(buffers[i]), getInputBufferByName(reciprocal_buffers[i]),
                                  linear_reciprocals[i] == "0" ? nullptr : &getInputBufferByName(linear_reciprocals[i]),
                                  getInputBufferByName(nonlinear_reciprocals[i]), getBufferOldByName(nonlinear_reciprocals[i], history_size), buffers[i],
                                  reciprocal_buffers[i], linear_reciprocals[i], nonlinear_reciprocals[i]});
}

// ==================================================================================== ExplicitSolverBase
InputParameters ExplicitSolverBase::validParams() {
  InputParameters params = TensorSolver::validParams();
  params.addClassDescription("Base class for explicit time integrators.");
  params.addParam<std::vector<TensorOutputBufferName>>("buffer", {}, "The buffer this solver is writing to");
  params.addParam<std::vector<TensorInputBufferName>>("reciprocal_buffer", {}, "Buffer with the reciprocal of the integrated buffer");
  params.addParam<std::vector<TensorInputBufferName>>("time_derivative_reciprocal", {},
                                                      "Buffer with the reciprocal of the time derivative (e.g. the divergence of the flux)");
  return params;
}

ExplicitSolverBase::ExplicitSolverBase(const InputParameters &parameters) : TensorSolver(parameters) {
  const auto buffers = getParam<std::vector<TensorOutputBufferName>>("buffer");
  const auto reciprocal_buffers = getParam<std::vector<TensorInputBufferName>>("reciprocal_buffer");
  const auto tdr = getParam<std::vector<TensorInputBufferName>>("time_derivative_reciprocal");
  const auto n = buffers.size();
  if (reciprocal_buffers.size() != n || tdr.size() != n)
    paramError("buffer", "Must have the same number of entries as 'reciprocal_buffer' and 'time_derivative_reciprocal'.");
  for (std::size_t i = 0; i < n; ++i)
    _variables.push_back(Variable{getOutputBufferByName(buffers[i]), getInputBufferByName(reciprocal_buffers[i]), getInputBufferByName(tdr[i])});
}

// ==================================================================================== ForwardEulerSolver
registerMooseObject("MarlinApp", ForwardEulerSolver);

InputParameters ForwardEulerSolver::validParams() {
  InputParameters params = ExplicitSolverBase::validParams();
  params.addClassDescription("Explicit forward Euler time integration solver.");
  return params;
}
ForwardEulerSolver::ForwardEulerSolver(const InputParameters &parameters) : ExplicitSolverBase(parameters) {}

void ForwardEulerSolver::substep() {
  _compute->computeBuffer();
  forwardBuffers();
  // u = ifft(ubar + dt * udot_bar): the AB update with beta = {1} and no linear operator
  const double beta[1] = {1.0};
  for (auto &[u, reciprocal_buffer, time_derivative_reciprocal] : _variables) {
    Tensor ubar = _domain.empty(Space::RECIPROCAL, true, 1);
    checkC(mrl_ab_update(_domain.context(), ubar.data_ptr(), reciprocal_buffer.data_ptr(), time_derivative_reciprocal.data_ptr(), nullptr, _sub_dt, beta, 0,
                         nullptr),
           "mrl_ab_update");
    u = _domain.ifft(ubar);
  }
}

// ================================================================================= AdamsBashforthMoulton
registerMooseObject("MarlinApp", AdamsBashforthMoulton);
registerMooseObjectRenamed("MarlinApp", SemiImplicitSolver, "10/01/2025 00:01", AdamsBashforthMoulton);

namespace {
// src/tensor_solver/AdamsBashforthMoulton.C:67-73.  The AB5 leading coefficient is 190/720 in the
// reference (quirk Q3 of SURVEY.md 8a); reproduced as coded.
constexpr double AB_BETA[5][5] = {
    {1.0, 0.0, 0.0, 0.0, 0.0},
    {3.0 / 2.0, -1.0 / 2.0, 0.0, 0.0, 0.0},
    {23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0, 0.0, 0.0},
    {55.0 / 24.0, -59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0, 0.0},
    {190.0 / 720.0, -2774.0 / 720.0, 2616.0 / 720.0, -1274.0 / 720.0, 251.0 / 720.0},
};
// :108-114
constexpr double AM_ALPHA[5][5] = {
    {1.0, 0.0, 0.0, 0.0, 0.0},
    {0.5, 0.5, 0.0, 0.0, 0.0},
    {5.0 / 12.0, 8.0 / 12.0, -1.0 / 12.0, 0.0, 0.0},
    {9.0 / 24.0, 19.0 / 24.0, -5.0 / 24.0, 1.0 / 24.0, 0.0},
    {251.0 / 720.0, 646.0 / 720.0, -264.0 / 720.0, 106.0 / 720.0, -19.0 / 720.0},
};
}  // namespace

InputParameters AdamsBashforthMoulton::validParams() {
  InputParameters params = SplitOperatorBase::validParams();
  params.addClassDescription("Adams-Bashforth-Moulton semi-implicit/explicit time integration solver with optional implicit corrector.");
  params.addParam<unsigned int>("substeps", 1, "semi-implicit substeps per time step.");
  params.addRangeCheckedParam<std::size_t>("predictor_order", 2, "predictor_order > 0 & predictor_order <= 5", "Order of the Adams-Bashforth predictor.");
  params.addRangeCheckedParam<std::size_t>("corrector_order", 2, "corrector_order > 0 & corrector_order <= 5", "Order of the Adams-Moulton corrector.");
  params.addParam<std::size_t>("corrector_steps", 0, "Number the Adams-Moulton corrector steps to take (one is usually sufficient).");
  // not a reference parameter: lets a user (or a test) force the operator-by-operator path
  params.addParam<bool>("batch_substeps", true, "Run the steady-state part of a step's substep loop as one batched call (CUDA graph replay) when the "
                                                "fused plan is active (marlin_b200 extension; results and postprocessor histories are identical).");
  params.addParam<bool>("fuse", true, "Replace the root compute and the update by the fused five-pass CUDA plan when the root compute has the canonical "
                                      "split-operator structure (marlin_b200 extension; results are identical).");
  return params;
}

AdamsBashforthMoulton::AdamsBashforthMoulton(const InputParameters &parameters)
  : SplitOperatorBase(parameters),
    _predictor_order(getParam<std::size_t>("predictor_order") - 1),
    _corrector_order(getParam<std::size_t>("corrector_order") - 1),
    _corrector_steps(getParam<std::size_t>("corrector_steps")),
    _allow_fusion(getParam<bool>("fuse")),
    _allow_batching(getParam<bool>("batch_substeps")) {
  const auto history = std::max(_predictor_order, _corrector_order);
  getVariables(history);
}

AdamsBashforthMoulton::~AdamsBashforthMoulton() {
  bool slab = false;
  for (auto &p : _plans) slab = slab || p.slab;
  if (slab) {
    // the peers may still be writing into this rank's staging buffers
    mrl_synchronize(_domain.context());
    try {
      _domain.comm().barrier();
    } catch (const std::exception &) {
    }
  }
  for (auto &p : _plans) {
    if (p.plan) mrl_split_plan_destroy(p.plan);
    if (p.slab) mrl_slab_plan_destroy(p.slab);
    if (p.expr) mrl_expr_destroy(p.expr);
  }
}

void AdamsBashforthMoulton::check() {}

// called at the first substep: the initial conditions have run, so every IC-time buffer the plans
// capture (mobility, linear operator, the variables themselves) is defined
void AdamsBashforthMoulton::decideFusion() {
  _fusion_decided = true;
  if (_allow_fusion) tryBuildFusedPlans();
  if (_tensor_problem.debugOutput())
    mooseInfo("AdamsBashforthMoulton '", name(), "': ", fused() ? (_plans[0].slab ? "fused slab-decomposed plan" : "fused five-pass plan") : "operator-by-operator path",
              _fusion_note.empty() ? "" : " (", _fusion_note,
              _fusion_note.empty() ? "" : ")");
}

// Recognises, per solver variable, the structure
//     g = ParsedCompute(F(real-space buffers))          [Solve]
//     gbar = ForwardFFT(g), cbar = ForwardFFT(c)        [Solve]
//     N = ParsedCompute('M*gbar') with M from a ReciprocalLaplacianFactor IC  -or-  N = gbar
//     L = ReciprocalLaplacian(Square)Factor IC or any IC-time real reciprocal buffer or none
// and replaces it by mrl_split_plan (include/marlin_b200.h).  Anything else keeps the generic path.
void AdamsBashforthMoulton::tryBuildFusedPlans() {
  auto note = [&](const std::string &why) { _fusion_note = "not fused: " + why; };
  if (_corrector_steps) return note("corrector steps need the separate operators");
  if (!_forwarded_buffers.empty()) return note("forward_buffer in use");
  auto *group = dynamic_cast<ComputeGroup *>(_compute.get());
  if (!group) return note("root compute is not a ComputeGroup");

  // flatten nested groups
  std::vector<TensorOperatorBase *> ops;
  std::function<void(ComputeGroup *)> flatten = [&](ComputeGroup *g) {
    for (const auto &c : g->getComputes()) {
      if (auto *sub = dynamic_cast<ComputeGroup *>(c.get()))
        flatten(sub);
      else if (std::find(ops.begin(), ops.end(), c.get()) == ops.end())
        ops.push_back(c.get());
    }
  };
  flatten(group);
  auto supplier = [&](const std::string &buf) -> TensorOperatorBase * {
    for (auto *o : ops)
      if (o->getSuppliedItems().count(buf)) return o;
    return nullptr;
  };
  auto ic_supplier = [&](const std::string &buf) -> TensorOperatorBase * {
    for (const auto &o : _tensor_problem.getICs())
      if (o->getSuppliedItems().count(buf)) return o.get();
    return nullptr;
  };
  const auto observed = _tensor_problem.observedBuffers();
  std::set<std::string> variable_names;
  for (const auto &v : _variables) variable_names.insert(v._buffer_name);

  struct Match {
    ParsedCompute *G = nullptr;
    int M_mode = 2;
    double M_factor = 0;
    std::string M_name;
    int has_L = 0, L_closed = 0;
    double L_factor = 0;
    std::string g_name;
    int var_index = -1;
  };
  std::vector<Match> matches;
  std::set<TensorOperatorBase *> used;
  for (const auto &v : _variables) {
    Match m;
    auto *cfft = dynamic_cast<ForwardFFT *>(supplier(v._reciprocal_name));
    if (!cfft || cfft->getRequestedItems().count(v._buffer_name) == 0) return note("'" + v._reciprocal_name + "' is not the ForwardFFT of '" + v._buffer_name + "'");
    used.insert(cfft);
    if (observed.count(v._reciprocal_name)) return note("'" + v._reciprocal_name + "' is read by a postprocessor");
    TensorOperatorBase *nsup = supplier(v._nonlinear_name);
    if (observed.count(v._nonlinear_name)) return note("'" + v._nonlinear_name + "' is read by a postprocessor");
    ForwardFFT *gfft = dynamic_cast<ForwardFFT *>(nsup);
    if (!gfft) {
      auto *prod = dynamic_cast<ParsedCompute *>(nsup);
      if (!prod) return note("'" + v._nonlinear_name + "' is neither a ForwardFFT nor a ParsedCompute product");
      used.insert(prod);
      const auto &in = prod->kernel().inputs();
      if (in.size() != 2 || !prod->kernel().derivatives().empty()) return note("nonlinear term is not a two-factor product");
      // toString() of a product is "(a * b)": compare without blanks and the enclosing parentheses
      std::string s;
      for (char ch : prod->kernel().simplified())
        if (ch != ' ') s += ch;
      if (s.size() > 2 && s.front() == '(' && s.back() == ')' && s.find('(', 1) == std::string::npos) s = s.substr(1, s.size() - 2);
      int which = -1;
      if (s == in[0] + "*" + in[1]) which = 0;
      if (s == in[1] + "*" + in[0]) which = 1;
      if (which < 0) return note("nonlinear term '" + s + "' is not a plain product of its inputs");
      // one factor is the transformed nonlinearity, the other an IC-time mobility
      for (int a = 0; a < 2 && !gfft; ++a) {
        auto *f = dynamic_cast<ForwardFFT *>(supplier(in[a]));
        if (f && !supplier(in[1 - a])) {
          gfft = f;
          m.M_name = in[1 - a];
        }
      }
      if (!gfft) return note("no ForwardFFT factor in the nonlinear product");
      if (observed.count(*gfft->getSuppliedItems().begin())) return note("'" + *gfft->getSuppliedItems().begin() + "' is read by a postprocessor");
      if (auto *lap = dynamic_cast<ReciprocalLaplacianFactor *>(ic_supplier(m.M_name))) {
        m.M_mode = 1;
        m.M_factor = lap->factor();
      } else {
        m.M_mode = 0;
      }
    }
    used.insert(gfft);
    m.g_name = *gfft->getRequestedItems().begin();
    m.G = dynamic_cast<ParsedCompute *>(supplier(m.g_name));
    if (!m.G) return note("'" + m.g_name + "' is not produced by a ParsedCompute in the root compute");
    used.insert(m.G);
    if (m.G->kernel().extraSymbols()) return note("the nonlinearity uses extra_symbols");
    const auto &gin = m.G->kernel().inputs();
    if (gin.size() > 16) return note("too many inputs");
    for (size_t i = 0; i < gin.size(); ++i) {
      if (gin[i] == v._buffer_name) m.var_index = (int)i;
      if (supplier(gin[i])) return note("nonlinearity input '" + gin[i] + "' is computed inside the root compute");
    }
    if (v._linear_reciprocal) {
      m.has_L = 1;
      if (supplier(v._linear_name)) return note("linear operator is recomputed every substep");
      if (auto *l2 = dynamic_cast<ReciprocalLaplacianSquareFactor *>(ic_supplier(v._linear_name))) {
        m.L_closed = 1;
        m.L_factor = l2->factor();
      }
    }
    matches.push_back(m);
  }
  for (auto *o : ops)
    if (!used.count(o)) return note("compute '" + o->name() + "' is not part of the split-operator pattern");

  // build: inputs must be defined (ICs have run) and live in real space
  std::vector<FusedVariable> plans;
  auto fail = [&](const std::string &why) {
    for (auto &p : plans) {
      if (p.plan) mrl_split_plan_destroy(p.plan);
      if (p.slab) mrl_slab_plan_destroy(p.slab);
      if (p.expr) mrl_expr_destroy(p.expr);
    }
    note(why);
  };
  const bool parallel = _domain.isParallelFFT();
  for (std::size_t k = 0; k < _variables.size(); ++k) {
    const auto &v = _variables[k];
    const auto &m = matches[k];
    FusedVariable fv;
    const auto &K = m.G->kernel();
    std::vector<const char *> in, der, cn;
    for (const auto &s : K.inputs()) in.push_back(s.c_str());
    for (const auto &s : K.derivatives()) der.push_back(s.c_str());
    for (const auto &s : K.constantNames()) cn.push_back(s.c_str());
    std::vector<int> layouts(in.size(), MRL_VAR_REAL);
    mrl_expr_desc d;
    std::memset(&d, 0, sizeof d);
    d.expression = K.expression().c_str();
    d.nvars = (int)in.size();
    d.var_names = in.data();
    d.var_layouts = layouts.data();
    d.nderivatives = (int)der.size();
    d.derivatives = der.data();
    d.nconstants = (int)cn.size();
    d.constant_names = cn.data();
    d.constant_values = K.constantValues().data();
    if (mrl_expr_compile(_domain.context(), &d, &fv.expr) != MRL_OK) return fail(std::string("expression: ") + mrl_last_error());
    mrl_split_desc sd;
    std::memset(&sd, 0, sizeof sd);
    sd.nonlin_kind = MRL_NONLIN_EXPR;
    sd.nonlin_expr = fv.expr;
    sd.M_closed_form = m.M_mode;
    sd.M_factor = m.M_factor;
    if (m.M_mode == 0) {
      const Tensor &M = _tensor_problem.getBuffer(m.M_name);
      if (!M.defined() || M.space() != Space::RECIPROCAL || M.is_complex()) {
        mrl_expr_destroy(fv.expr);
        return fail("mobility '" + m.M_name + "' is not an initialised real reciprocal-space buffer");
      }
      sd.M_real_dev = M.data_ptr();
    }
    sd.has_L = m.has_L;
    sd.L_closed_form = m.L_closed;
    sd.L_factor = m.L_factor;
    if (m.has_L && !m.L_closed) {
      const Tensor &L = *v._linear_reciprocal;
      if (!L.defined() || L.space() != Space::RECIPROCAL || L.is_complex()) {
        mrl_expr_destroy(fv.expr);
        return fail("linear operator '" + v._linear_name + "' is not an initialised real reciprocal-space buffer");
      }
      sd.L_real_dev = L.data_ptr();
    }
    sd.history = (int)_predictor_order;
    sd.nonlin_var = m.var_index;
    for (size_t i = 0; i < K.inputs().size(); ++i) {
      const Tensor &t = _tensor_problem.getBuffer(K.inputs()[i]);
      if (!t.defined() || t.space() != Space::REAL || t.is_complex() || t.ncomp() != 1) {
        mrl_expr_destroy(fv.expr);
        return fail("nonlinearity input '" + K.inputs()[i] + "' is not an initialised real field");
      }
      sd.nonlin_inputs_dev[i] = t.data_ptr();
    }
    if (observed.count(m.g_name)) {
      // keep the real-space nonlinearity materialised for postprocessors / outputs
      Tensor &g = _tensor_problem.getBuffer(m.g_name);
      if (!g.defined()) g = _domain.zeros(Space::REAL, false, 1);
      sd.g_out_real_dev = g.data_ptr();
      fv.g_name = m.g_name;
    }
    if (parallel) {
      // every rank takes the same decision: a plan any rank cannot build is dropped by all
      std::string why;
      if (mrl_slab_plan_create_peer(_domain.context(), &sd, &fv.slab) != MRL_OK) {
        why = mrl_last_error();
        fv.slab = nullptr;
      }
      double ok = fv.slab ? 1.0 : 0.0;
      _domain.comm().allreduce(&ok, 1, Comm::MIN);
      if (ok == 0.0) {
        if (fv.slab) mrl_slab_plan_destroy(fv.slab);
        mrl_expr_destroy(fv.expr);
        return fail("slab plan: " + (why.empty() ? std::string("another rank could not build it") : why));
      }
      // the peers' staging buffers (CUDA IPC), where the reference posts its MPI messages
      unsigned char mine[128];
      checkC(mrl_slab_ipc_export(fv.slab, mine), "mrl_slab_ipc_export");
      std::vector<unsigned char> all(sizeof mine * _domain.nRanks());
      _domain.comm().allgather(mine, sizeof mine, all.data());
      checkC(mrl_slab_ipc_import(fv.slab, all.data()), "mrl_slab_ipc_import");
      _domain.comm().barrier();
    } else if (mrl_split_plan_create(_domain.context(), &sd, &fv.plan) != MRL_OK) {
      const std::string why = mrl_last_error();
      mrl_expr_destroy(fv.expr);
      return fail("plan: " + why);
    }
    plans.push_back(fv);
  }
  _plans = plans;
  _fusion_note.clear();
  // the plans keep the history of the nonlinear terms; follow the problem's advanceState
  _tensor_problem.addAdvanceStateHook([this]() {
    for (auto &p : _plans) {
      if (p.slab)
        checkC(mrl_slab_advance_state(p.slab, &p.stored), "mrl_slab_advance_state");
      else
        checkC(mrl_split_advance_state(p.plan, &p.stored), "mrl_split_advance_state");
    }
  });
}

// TensorSolver::computeBuffer (TensorSolver.C:93-110) with the steady-state middle of the substep loop
// handed to mrl_split_substeps.  The first substeps (until the predictor runs at full order with a full
// ring) and the last maxOldStates()+1 substeps are issued one by one with TensorProblem::advanceState in
// between, so every old state a postprocessor or output reads at the end of the step is exactly what the
// plain loop leaves behind.
void AdamsBashforthMoulton::computeBuffer() {
  if (!_fusion_decided) decideFusion();
  const std::size_t tail = _tensor_problem.maxOldStates() + 1;
  const std::size_t P = _predictor_order;
  if (!_allow_batching || !fused() || _plans[0].slab || _variables.size() != 1 || _tensor_problem.timeStep() <= 1 || _substeps < tail + P + 4 * (P + 1) + 1)
    return TensorSolver::computeBuffer();
  _sub_dt = _dt / _substeps;
  const bool dt_changed = (_dt != _dt_old);
  _substep = 0;
  auto single = [&]() {
    substep();
    if (_substep < _substeps - 1) _tensor_problem.advanceState();
    _sub_time += _sub_dt;
    ++_substep;
  };
  while (_substep + tail < _substeps && !((!dt_changed || _substep >= P) && (std::size_t)_plans[0].stored >= P)) single();
  const long nbatch = (long)_substeps - (long)tail - (long)_substep;
  if (nbatch >= (long)(4 * (P + 1))) {
    Tensor &u = _variables[0]._buffer;
    if (u.use_count() > 1) {  // as in fusedSubstep: the plan updates the block in place
      Tensor copy = _domain.clone(u);
      Tensor::swapBlocks(u, copy);
      u = copy;
    }
    checkC(mrl_split_set_time(_plans[0].plan, _sub_time), "mrl_split_set_time");
    checkC(mrl_split_substeps(_plans[0].plan, u.data_ptr(), _sub_dt, AB_BETA[P], (int)P, (int)nbatch), "mrl_split_substeps");
    for (long i = 0; i < nbatch; ++i) _sub_time += _sub_dt;  // same summation order as the loop
    _substep += (unsigned int)nbatch;
  }
  while (_substep < _substeps) single();
}

void AdamsBashforthMoulton::fusedSubstep() {
  const bool dt_changed = (_dt != _dt_old);
  // every variable's nonlinearity is evaluated from the OLD fields before any variable is updated
  for (std::size_t k = 0; k < _variables.size(); ++k) {
    Tensor &u = _variables[k]._buffer;
    if (u.use_count() > 1) {
      // an old state (or another holder) still refers to this block and the plan updates it in
      // place: give the other holders a copy, keep the device pointer of the variable stable
      // (the plans captured it as an expression input)
      Tensor copy = _domain.clone(u);
      Tensor::swapBlocks(u, copy);
      u = copy;
    }
  }
  if (_plans[0].slab) {
    // z r2c + x forward with the rows pushed into the owners' HBM | y pass + update, result rows pushed back | x inverse + z c2r
    for (std::size_t k = 0; k < _variables.size(); ++k) checkC(mrl_slab_forward(_plans[k].slab, _variables[k]._buffer.data_ptr()), "mrl_slab_forward");
    for (std::size_t k = 0; k < _variables.size(); ++k) checkC(mrl_slab_barrier(_plans[k].slab), "mrl_slab_barrier");
    for (std::size_t k = 0; k < _variables.size(); ++k) {
      const std::size_t n_old = (std::size_t)_plans[k].stored;
      const auto order = std::min(_substep < _predictor_order && dt_changed ? std::size_t(0) : n_old, _predictor_order);
      checkC(mrl_slab_update(_plans[k].slab, _sub_dt, AB_BETA[order], (int)order), "mrl_slab_update");
    }
    for (std::size_t k = 0; k < _variables.size(); ++k) checkC(mrl_slab_barrier(_plans[k].slab), "mrl_slab_barrier");
    for (std::size_t k = 0; k < _variables.size(); ++k) checkC(mrl_slab_inverse(_plans[k].slab, _variables[k]._buffer.data_ptr()), "mrl_slab_inverse");
    return;
  }
  for (std::size_t k = 0; k < _variables.size(); ++k) {
    checkC(mrl_split_set_time(_plans[k].plan, _sub_time), "mrl_split_set_time");
    checkC(mrl_split_forward(_plans[k].plan, _variables[k]._buffer.data_ptr()), "mrl_split_forward");
  }
  for (std::size_t k = 0; k < _variables.size(); ++k) {
    const std::size_t n_old = (std::size_t)_plans[k].stored;
    const auto order = std::min(_substep < _predictor_order && dt_changed ? std::size_t(0) : n_old, _predictor_order);
    checkC(mrl_split_finish(_plans[k].plan, _variables[k]._buffer.data_ptr(), _sub_dt, AB_BETA[order], (int)order), "mrl_split_finish");
  }
}

void AdamsBashforthMoulton::substep() {
  if (!_fusion_decided) decideFusion();
  if (fused()) return fusedSubstep();

  _compute->computeBuffer();
  forwardBuffers();
  const bool dt_changed = (_dt != _dt_old);
  auto update = [&](const Tensor &base, const Tensor &N, const Tensor *L, const double *coef, const std::vector<const void *> &old) {
    Tensor ubar = _domain.empty(Space::RECIPROCAL, true, 1);
    checkC(mrl_ab_update(_domain.context(), ubar.data_ptr(), base.data_ptr(), N.data_ptr(), L ? L->data_ptr() : nullptr, _sub_dt, coef, (int)old.size(),
                         old.empty() ? nullptr : old.data()),
           "mrl_ab_update");
    return _domain.ifft(ubar);
  };

  // Adams-Bashforth predictor on all variables
  for (auto &v : _variables) {
    const auto n_old = v._old_nonlinear_reciprocal.size();
    const auto order = std::min(_substep < _predictor_order && dt_changed ? std::size_t(0) : n_old, _predictor_order);
    std::vector<const void *> old;
    for (std::size_t i = 0; i < order; ++i) old.push_back(v._old_nonlinear_reciprocal[i].data_ptr());
    v._buffer = update(v._reciprocal_buffer, v._nonlinear_reciprocal, v._linear_reciprocal, AB_BETA[order], old);
  }

  // Adams-Moulton corrector
  if (_corrector_steps) {
    _sub_time += _sub_dt;
    std::vector<Tensor> ubar_n(_variables.size()), N_n;
    for (std::size_t k = 0; k < _variables.size(); ++k) ubar_n[k] = _variables[k]._reciprocal_buffer;
    if (_corrector_order > 0) {
      N_n.resize(_variables.size());
      for (std::size_t k = 0; k < _variables.size(); ++k) N_n[k] = _variables[k]._nonlinear_reciprocal;
    }
    for (std::size_t j = 0; j < _corrector_steps; ++j) {
      // re-evaluate the solve compute with the predicted variable values
      _compute->computeBuffer();
      forwardBuffers();
      for (std::size_t k = 0; k < _variables.size(); ++k) {
        auto &v = _variables[k];
        const auto n_old = v._old_nonlinear_reciprocal.size();
        const auto order = std::min(_substep < _corrector_order && dt_changed ? std::size_t(1) : n_old + 1, _corrector_order);
        if (order == 0) continue;  // corrector_order = 1 is a no-op (quirk Q4)
        std::vector<const void *> old = {N_n[k].data_ptr()};
        for (std::size_t i = 0; i + 1 < order; ++i) old.push_back(v._old_nonlinear_reciprocal[i].data_ptr());
        v._buffer = update(ubar_n[k], v._nonlinear_reciprocal, v._linear_reciprocal, AM_ALPHA[order], old);
      }
    }
    _sub_time -= _sub_dt;
  }
}

// ========================================================================= AdamsBashforthMoultonCoupled
registerMooseObject("MarlinApp", AdamsBashforthMoultonCoupled);

InputParameters AdamsBashforthMoultonCoupled::validParams() {
  InputParameters params = SplitOperatorBase::validParams();
  params.addClassDescription("Coupled Adams-Bashforth-Moulton solver with dense linear operator and batched solve in reciprocal space.");
  params.addParam<unsigned int>("substeps", 1, "semi-implicit substeps per time step.");
  params.addRangeCheckedParam<std::size_t>("predictor_order", 2, "predictor_order > 0 & predictor_order <= 5", "Order of the Adams-Bashforth predictor.");
  params.addRangeCheckedParam<std::size_t>("corrector_order", 2, "corrector_order > 0 & corrector_order <= 5", "Order of the Adams-Moulton corrector.");
  params.addParam<std::size_t>("corrector_steps", 0, "Number of Adams-Moulton corrector steps (0 disables the corrector).");
  params.addParam<std::vector<unsigned int>>("linear_offdiag_rows", {}, "Row indices for L_ij.");
  params.addParam<std::vector<unsigned int>>("linear_offdiag_cols", {}, "Column indices for L_ij.");
  params.addParam<std::vector<TensorInputBufferName>>("linear_offdiag", {}, "Off-diagonal linear operator buffers.");
  params.addParam<bool>("assume_symmetric", false, "Mirror off-diagonal entries (i,j) into (j,i) if not explicitly provided.");
  return params;
}

AdamsBashforthMoultonCoupled::AdamsBashforthMoultonCoupled(const InputParameters &parameters)
  : SplitOperatorBase(parameters),
    _predictor_order(getParam<std::size_t>("predictor_order") - 1),
    _corrector_order(getParam<std::size_t>("corrector_order") - 1),
    _corrector_steps(getParam<std::size_t>("corrector_steps")),
    _assume_symmetric(getParam<bool>("assume_symmetric")),
    _L_offdiag_indices(getParam<unsigned int, unsigned int>("linear_offdiag_rows", "linear_offdiag_cols")),
    _L_offdiag_names(getParam<std::vector<TensorInputBufferName>>("linear_offdiag")) {
  getVariables(std::max(_predictor_order, _corrector_order));
  if (_L_offdiag_indices.size() != _L_offdiag_names.size())
    paramError("linear_offdiag", "'linear_offdiag_rows', 'linear_offdiag_cols', and 'linear_offdiag' must all have the same length.");
  const auto N = _variables.size();
  for (const auto &[i, j] : _L_offdiag_indices) {
    if (i >= N) paramError("linear_offdiag_rows", "Off-diagonal indices out of range.");
    if (j >= N) paramError("linear_offdiag_cols", "Off-diagonal indices out of range.");
  }
  if (N > 6) paramError("buffer", "The CUDA per-wavevector solve supports at most 6 coupled variables.");
  for (const auto &name : _L_offdiag_names) _L_offdiag_buffer.push_back(&getInputBufferByName(name));
}

// :131-171 (and :225-257 for the corrector): assemble L, solve (I - dt L) ubar = rhs per
// wavevector, inverse transform.  As coded in the reference:
//  * L is assembled as stack(stack(cols) per row, -1), so the matrix that reaches linalg_solve
//    is the TRANSPOSE of the L_ij table: equation a, unknown b  <-  table entry (b, a);
//  * the right-hand side is cast to the dtype of the first variable's linear operator (real):
//    only its real part is used (gold coupled_*.csv pin this).
void AdamsBashforthMoultonCoupled::solveAndInvert(const std::vector<Tensor> &rhs) {
  const auto N = _variables.size();
  if (!_variables[0]._linear_reciprocal) paramError("linear_reciprocal", "The first variable needs a linear operator buffer (the reference dereferences it).");
  std::vector<const Tensor *> table(N * N, nullptr);
  for (std::size_t i = 0; i < N; ++i) table[i * N + i] = _variables[i]._linear_reciprocal;
  for (std::size_t k = 0; k < _L_offdiag_buffer.size(); ++k) {
    const auto &[i, j] = _L_offdiag_indices[k];
    table[i * N + j] = _L_offdiag_buffer[k];
  }
  if (_assume_symmetric)
    for (std::size_t k = 0; k < _L_offdiag_buffer.size(); ++k) {
      const auto &[i, j] = _L_offdiag_indices[k];
      if (i != j && !table[j * N + i]) table[j * N + i] = _L_offdiag_buffer[k];
    }
  std::vector<const void *> L(N * N, nullptr), b(N);
  std::vector<void *> out(N);
  std::vector<Tensor> ubar(N);
  for (std::size_t a = 0; a < N; ++a)
    for (std::size_t c = 0; c < N; ++c)
      if (const Tensor *t = table[c * N + a]) {
        if (!t->defined()) mooseError("AdamsBashforthMoultonCoupled: a linear operator buffer is not defined");
        if (t->is_complex()) mooseError("AdamsBashforthMoultonCoupled: linear operator buffers are expected to be real");
        L[a * N + c] = t->data_ptr();
      }
  for (std::size_t i = 0; i < N; ++i) {
    ubar[i] = _domain.empty(Space::RECIPROCAL, true, 1);
    b[i] = rhs[i].data_ptr();
    out[i] = ubar[i].data_ptr();
  }
  const bool drop_imag = !_variables[0]._linear_reciprocal->is_complex();
  checkC(mrl_coupled_solve(_domain.context(), (int)N, L.data(), b.data(), out.data(), _sub_dt, drop_imag ? 1 : 0), "mrl_coupled_solve");
  for (std::size_t i = 0; i < N; ++i) _variables[i]._buffer = _domain.ifft(ubar[i]);
}

void AdamsBashforthMoultonCoupled::substep() {
  _compute->computeBuffer();
  forwardBuffers();
  const bool dt_changed = (_dt != _dt_old);
  const auto N = _variables.size();
  if (N == 0) return;
  // right-hand side cbar + dt * sum_i coef_i N_i: the AB update without a linear operator
  auto combine = [&](const Tensor &base, const Tensor &Nl, const double *coef, const std::vector<const void *> &old) {
    Tensor r = _domain.empty(Space::RECIPROCAL, true, 1);
    checkC(mrl_ab_update(_domain.context(), r.data_ptr(), base.data_ptr(), Nl.data_ptr(), nullptr, _sub_dt, coef, (int)old.size(),
                         old.empty() ? nullptr : old.data()),
           "mrl_ab_update");
    return r;
  };
  std::vector<Tensor> rhs(N);
  for (std::size_t i = 0; i < N; ++i) {
    auto &v = _variables[i];
    const auto n_old = v._old_nonlinear_reciprocal.size();
    const auto order = std::min(_substep < _predictor_order && dt_changed ? std::size_t(0) : n_old, _predictor_order);
    std::vector<const void *> old;
    for (std::size_t j = 0; j < order; ++j) old.push_back(v._old_nonlinear_reciprocal[j].data_ptr());
    rhs[i] = combine(v._reciprocal_buffer, v._nonlinear_reciprocal, AB_BETA[order], old);
  }
  solveAndInvert(rhs);
  // :177 advances the sub-time here, in addition to TensorSolver::computeBuffer (as coded)
  _sub_time += _sub_dt;

  if (_corrector_steps) {
    std::vector<Tensor> ubar_n(N), N_n(N);
    for (std::size_t i = 0; i < N; ++i) ubar_n[i] = _variables[i]._reciprocal_buffer;
    if (_corrector_order > 0)
      for (std::size_t i = 0; i < N; ++i) N_n[i] = _variables[i]._nonlinear_reciprocal;
    for (std::size_t jc = 0; jc < _corrector_steps; ++jc) {
      _compute->computeBuffer();
      forwardBuffers();
      for (std::size_t i = 0; i < N; ++i) {
        auto &v = _variables[i];
        const auto n_old = v._old_nonlinear_reciprocal.size();
        const auto order = std::min(_substep < _corrector_order && dt_changed ? std::size_t(1) : n_old + 1, _corrector_order);
        if (order == 0) {  // :213-217: the solve still runs on the old spectrum
          rhs[i] = ubar_n[i];
          continue;
        }
        std::vector<const void *> old = {N_n[i].data_ptr()};
        for (std::size_t j = 0; j + 1 < order; ++j) old.push_back(v._old_nonlinear_reciprocal[j].data_ptr());
        rhs[i] = combine(ubar_n[i], v._nonlinear_reciprocal, AM_ALPHA[order], old);
      }
      solveAndInvert(rhs);
    }
  }
}

// ========================================================================================== SecantSolver
registerMooseObject("MarlinApp", SecantSolver);

InputParameters SecantSolver::validParams() {
  InputParameters params = SplitOperatorBase::validParams();
  params.addClassDescription("Implicit secant solver time integration.");
  params.addParam<unsigned int>("substeps", 1, "secant solver substeps per time step.");
  params.addParam<unsigned int>("max_iterations", 30, "Maximum number of secant solver iteration.");
  params.addParam<Real>("relative_tolerance", 1e-9, "Convergence tolerance.");
  params.addParam<Real>("absolute_tolerance", 1e-9, "Convergence tolerance.");
  params.addParam<Real>("damping", 1.0, "Damping factor for the update step.");
  params.addParam<Real>("dt_epsilon", 1e-4, "Semi-implicit stable timestep to bootstrap secant solve.");
  params.addParam<bool>("verbose", false, "Show convergence history.");
  return params;
}

SecantSolver::SecantSolver(const InputParameters &parameters)
  : SplitOperatorBase(parameters),
    _max_iterations(getParam<unsigned int>("max_iterations")),
    _relative_tolerance(getParam<Real>("relative_tolerance")),
    _absolute_tolerance(getParam<Real>("absolute_tolerance")),
    _verbose(getParam<bool>("verbose")),
    _damping(getParam<Real>("damping")),
    _dt_epsilon(getParam<Real>("dt_epsilon")) {
  getVariables(0);  // no history required
  if (_variables.size() > 1)
    mooseWarning("The secant solver only work well for uncoupled variables. Use the BroydenSolver for solves with multiple coupled variables.");
  // SecantSolver.C:71-75,85-88 (bootstrap), :122-126 (residual), :129-139 (secant update); `t` is the sub step
  const int E = MRL_EXPAND_NONE;
  _r0[1].configure("(N + L*u)*t", {"N", "L", "u"}, {}, {}, {}, true, E);
  _r0[0].configure("N*t", {"N"}, {}, {}, {}, true, E);
  _start[1].configure("(u + eps*N)/(1 - eps*L)", {"N", "L", "u"}, {}, {"eps"}, {_dt_epsilon}, false, E);
  _start[0].configure("u + eps*N", {"N", "u"}, {}, {"eps"}, {_dt_epsilon}, false, E);
  _res[1].configure("(N + L*u)*t + uold - u", {"N", "L", "u", "uold"}, {}, {}, {}, true, E);
  _res[0].configure("N*t + uold - u", {"N", "u", "uold"}, {}, {}, {}, true, E);
  _update.configure(_damping == 1.0 ? "dy := R - Rp; u + if(dy != 0, -R*(u - up)/dy, 0)" : "dy := R - Rp; u + if(dy != 0, -R*(u - up)/dy, 0)*damping",
                    {"u", "up", "R", "Rp"}, {}, {"damping"}, {_damping}, false, E);
}

// torch::norm of a complex tensor: sqrt(sum |z|^2)
Real SecantSolver::complexNorm(const Tensor &t) const {
  double s = 0;
  checkC(mrl_reduce(_domain.context(), MRL_SUMSQ, t.data_ptr(), t.numel() * (t.is_complex() ? 2 : 1), &s), "mrl_reduce");
  _domain.comm().allreduce(&s, 1, Comm::SUM);  // every rank takes the same convergence decision
  return std::sqrt(s);
}

void SecantSolver::substep() {
  const auto n = _variables.size();
  std::vector<Tensor> u_old(n), Rprev(n), uprev(n);
  std::vector<Real> R0norm(n);
  if (_verbose) std::cerr << "Substep " << _substep << "\n";

  // initial guess computed using semi-implicit Euler
  _compute->computeBuffer();
  forwardBuffers();
  for (std::size_t i = 0; i < n; ++i) {
    auto &v = _variables[i];
    const Tensor &u = v._reciprocal_buffer, &N = v._nonlinear_reciprocal;
    const Tensor *L = v._linear_reciprocal;
    Rprev[i] = L ? _r0[1].eval(_domain, {&N, L, &u}, _sub_dt) : _r0[0].eval(_domain, {&N}, _sub_dt);
    uprev[i] = u;
    R0norm[i] = complexNorm(Rprev[i]);
    u_old[i] = u;
    v._buffer = _domain.ifft(L ? _start[1].eval(_domain, {&N, L, &u}, 0.0) : _start[0].eval(_domain, {&N, &u}, 0.0));
    if (_verbose) std::cerr << "|R0|=" << R0norm[i] << std::endl;
  }

  // forward predict (on solver outputs), SecantSolver.C:99-100
  applyPredictors();

  bool all_converged = false;
  for (_iterations = 0; _iterations < _max_iterations; ++_iterations) {
    _compute->computeBuffer();
    forwardBuffers();
    all_converged = true;
    for (std::size_t i = 0; i < n; ++i) {
      auto &v = _variables[i];
      const Tensor u = v._reciprocal_buffer;  // keep this state alive: it becomes uprev
      const Tensor &N = v._nonlinear_reciprocal;
      const Tensor *L = v._linear_reciprocal;
      Tensor R = L ? _res[1].eval(_domain, {&N, L, &u, &u_old[i]}, _sub_dt) : _res[0].eval(_domain, {&N, &u, &u_old[i]}, _sub_dt);
      Tensor unew = _update.eval(_domain, {&u, &uprev[i], &R, &Rprev[i]}, 0.0);
      uprev[i] = u;
      Rprev[i] = R;
      v._buffer = _domain.ifft(unew);
      const Real Rnorm = complexNorm(R);
      if (_verbose) std::cerr << _iterations << " |R|=" << Rnorm << std::endl;
      if (std::isnan(Rnorm)) {
        all_converged = false;
        _iterations = _max_iterations;
        std::cerr << "NaN detected, aborting solve.\n";
        break;
      }
      all_converged = all_converged && (Rnorm < _absolute_tolerance || Rnorm / R0norm[i] < _relative_tolerance);
    }
    if (all_converged) {
      _is_converged = true;
      break;
    }
  }
  if (!all_converged) {
    std::cerr << "Solve not converged.\n";
    for (std::size_t i = 0; i < n; ++i) _variables[i]._buffer = _domain.ifft(u_old[i]);
    _is_converged = false;
  }
}

// ========================================================================================== BroydenSolver
registerMooseObject("MarlinApp", BroydenSolver);

InputParameters BroydenSolver::validParams() {
  InputParameters params = SplitOperatorBase::validParams();
  params.addClassDescription("Implicit secant solver time integration.");
  params.addParam<unsigned int>("substeps", 1, "secant solver substeps per time step.");
  params.addParam<unsigned int>("max_iterations", 5, "Maximum number of secant solver iteration.");
  params.addParam<Real>("relative_tolerance", 1e-9, "Convergence tolerance.");
  params.addParam<Real>("absolute_tolerance", 1e-9, "Convergence tolerance.");
  params.addParam<Real>("damping", 1.0, "Damping factor for the update step.");
  params.addParam<Real>("initial_jacobian_guess", 1.0, "Factor for the initial inverse jacobian guess.");
  params.addParam<Real>("dt_epsilon", 1e-4, "Semi-implicit stable timestep to bootstrap secant solve.");
  params.addParam<bool>("verbose", false, "Show convergence history.");
  return params;
}

BroydenSolver::BroydenSolver(const InputParameters &parameters)
  : SplitOperatorBase(parameters),
    _max_iterations(getParam<unsigned int>("max_iterations")),
    _relative_tolerance(getParam<Real>("relative_tolerance")),
    _absolute_tolerance(getParam<Real>("absolute_tolerance")),
    _verbose(getParam<bool>("verbose")),
    _eye_factor(getParam<Real>("initial_jacobian_guess")) {
  getVariables(0);
  if (_variables.size() > 6) paramError("buffer", "The CUDA per-wavevector Broyden update supports at most 6 coupled variables.");
  // BroydenSolver.C:97 (first residual, u = u_old) and :137 (later residuals); `t` is the sub step
  const int E = MRL_EXPAND_NONE;
  _res0[1].configure("(N + L*u)*t", {"N", "L", "u"}, {}, {}, {}, true, E);
  _res0[0].configure("N*t", {"N"}, {}, {}, {}, true, E);
  _res[1].configure("(N + L*u)*t + uold - u", {"N", "L", "u", "uold"}, {}, {}, {}, true, E);
  _res[0].configure("N*t + uold - u", {"N", "u", "uold"}, {}, {}, {}, true, E);
}

// torch::norm of the stacked residual: sqrt(sum_i sum |R_i|^2)
Real BroydenSolver::stackedNorm(const std::vector<Tensor> &R) const {
  double total = 0;
  for (const Tensor &t : R) {
    double s = 0;
    checkC(mrl_reduce(_domain.context(), MRL_SUMSQ, t.data_ptr(), t.numel() * (t.is_complex() ? 2 : 1), &s), "mrl_reduce");
    total += s;
  }
  _domain.comm().allreduce(&total, 1, Comm::SUM);
  return std::sqrt(total);
}

// BroydenSolver.C:69-176, as coded: the step is u + 0.5 sk (`damping` and `dt_epsilon` are read but unused)
void BroydenSolver::substep() {
  const auto n = _variables.size();
  if (!_M.defined()) {
    // eye(n) * initial_jacobian_guess at every wavevector (:50-63)
    const size_t pts = size_t(_domain.getNumberOfReciprocalCells());
    std::vector<double> host(2 * n * n * pts, 0.0);
    for (std::size_t i = 0; i < n; ++i)
      for (size_t q = 0; q < pts; ++q) host[2 * ((i * n + i) * pts + q)] = _eye_factor;
    _M = _domain.fromHost(host, Space::RECIPROCAL, true, int(n * n));
  }
  _compute->computeBuffer();
  forwardBuffers();

  std::vector<Tensor> u_old(n), u(n), R(n), Rnew(n), sk(n), unew(n);
  auto residual = [&](bool first, std::vector<Tensor> &out) {
    for (std::size_t i = 0; i < n; ++i) {
      auto &v = _variables[i];
      u[i] = v._reciprocal_buffer;
      const Tensor &N = v._nonlinear_reciprocal;
      const Tensor *L = v._linear_reciprocal;
      if (first)
        out[i] = L ? _res0[1].eval(_domain, {&N, L, &u[i]}, _sub_dt) : _res0[0].eval(_domain, {&N}, _sub_dt);
      else
        out[i] = L ? _res[1].eval(_domain, {&N, L, &u[i], &u_old[i]}, _sub_dt) : _res[0].eval(_domain, {&N, &u[i], &u_old[i]}, _sub_dt);
    }
  };
  for (std::size_t i = 0; i < n; ++i) u_old[i] = _variables[i]._reciprocal_buffer;
  residual(true, R);
  const Real R0norm = stackedNorm(R);

  std::vector<const void *> pa(n), pb(n), pc(n);
  std::vector<void *> po0(n), po1(n);
  for (_iterations = 0; _iterations < _max_iterations; ++_iterations) {
    const Real Rnorm = stackedNorm(R);
    if (std::isnan(Rnorm)) mooseError("NAN!");
    if (_iterations > 4 && Rnorm * 10.0 / _iterations > R0norm) mooseWarning("Diverging residual ", Rnorm, " ", Rnorm * 10.0 / _iterations, ' ', R0norm);
    if (Rnorm < _absolute_tolerance || Rnorm / R0norm < _relative_tolerance) {
      if (_verbose) std::cout << "Broyden solve converged after " << _iterations << " iterations. |R|=" << Rnorm << " |R|/|R0|=" << Rnorm / R0norm << '\n';
      _is_converged = true;
      return;
    } else if (_verbose)
      std::cout << _iterations << " |R|=" << Rnorm << std::endl;

    for (std::size_t i = 0; i < n; ++i) {
      sk[i] = _domain.empty(Space::RECIPROCAL, true, 1);
      unew[i] = _domain.empty(Space::RECIPROCAL, true, 1);
      pa[i] = R[i].data_ptr();
      pb[i] = u[i].data_ptr();
      po0[i] = sk[i].data_ptr();
      po1[i] = unew[i].data_ptr();
    }
    checkC(mrl_broyden_step(_domain.context(), (int)n, _M.data_ptr(), pa.data(), pb.data(), po0.data(), po1.data()), "mrl_broyden_step");
    for (std::size_t i = 0; i < n; ++i) _variables[i]._buffer = _domain.ifft(unew[i]);

    _compute->computeBuffer();
    forwardBuffers();
    residual(false, Rnew);
    for (std::size_t i = 0; i < n; ++i) {
      pa[i] = sk[i].data_ptr();
      pb[i] = R[i].data_ptr();
      pc[i] = Rnew[i].data_ptr();
    }
    checkC(mrl_broyden_update(_domain.context(), (int)n, _M.data_ptr(), pa.data(), pb.data(), pc.data()), "mrl_broyden_update");
    R = Rnew;
  }
  std::cerr << "Broyden solve did not converge within the maximum number of iterations.\n";
  _is_converged = false;
}

// ========================================================================================== ETDRK4Solver
registerMooseObject("MarlinApp", ETDRK4Solver);

InputParameters ETDRK4Solver::validParams() {
  InputParameters params = SplitOperatorBase::validParams();
  params.addClassDescription("Fourth-order exponential time differencing solver.");
  return params;
}

ETDRK4Solver::ETDRK4Solver(const InputParameters &parameters) : SplitOperatorBase(parameters) { getVariables(0); }

ETDRK4Solver::~ETDRK4Solver() {
  for (auto *e : {_e_stage_half, _e_stage_full, _e_final})
    if (e) mrl_expr_destroy(e);
}

mrl_expr *ETDRK4Solver::kernel(mrl_expr *&slot, const char *expression, const std::vector<std::string> &names, const std::vector<int> &layouts) {
  if (slot) return slot;
  std::vector<const char *> in;
  for (const auto &s : names) in.push_back(s.c_str());
  mrl_expr_desc d;
  std::memset(&d, 0, sizeof d);
  d.expression = expression;
  d.nvars = (int)in.size();
  d.var_names = in.data();
  d.var_layouts = layouts.data();
  d.extra_symbols = 1;  // `t` carries the sub step dt
  checkC(mrl_expr_compile(_domain.context(), &d, &slot), "mrl_expr_compile (ETDRK4 stage kernel)");
  return slot;
}

// ETDRK4Solver.C:29-115, as coded there (not textbook Cox-Matthews): the stage and final
// combinations are generated pointwise kernels; `t` inside the expressions is the sub step.
void ETDRK4Solver::substep() {
  _compute->computeBuffer();
  forwardBuffers();
  const std::size_t n = _variables.size();
  const int LR = MRL_VAR_RECIP_REAL, LC = MRL_VAR_RECIP_COMPLEX;

  std::vector<Tensor> ubar_n(n), linear(n), N1(n), N2(n), N3(n), N4(n);
  for (std::size_t i = 0; i < n; ++i) {
    ubar_n[i] = _variables[i]._reciprocal_buffer;
    N1[i] = _variables[i]._nonlinear_reciprocal;
    linear[i] = _variables[i]._linear_reciprocal ? *_variables[i]._linear_reciprocal : _domain.zeros(Space::RECIPROCAL, false, 1);
    if (linear[i].is_complex()) mooseError("ETDRK4Solver expects a real linear_reciprocal buffer");
  }
  auto evaluate_nonlinear = [&](const std::vector<Tensor> &ubar_stage, std::vector<Tensor> &out) {
    for (std::size_t i = 0; i < n; ++i) _variables[i]._buffer = _domain.ifft(ubar_stage[i]);
    _compute->computeBuffer();
    forwardBuffers();
    for (std::size_t i = 0; i < n; ++i) out[i] = _variables[i]._nonlinear_reciprocal;
  };
  auto run = [&](mrl_expr *e, std::initializer_list<const Tensor *> in) {
    Tensor out = _domain.empty(Space::RECIPROCAL, true, 1);
    std::vector<const void *> p;
    for (const Tensor *t : in) p.push_back(t->data_ptr());
    checkC(mrl_expr_eval(e, p.data(), _sub_dt, out.data_ptr()), "mrl_expr_eval");
    return out;
  };

  mrl_expr *half = kernel(_e_stage_half, "exp(L*t/2)*un + 0.5*t*N", {"L", "un", "N"}, {LR, LC, LC});
  mrl_expr *full = kernel(_e_stage_full, "exp(L*t)*un + t*N", {"L", "un", "N"}, {LR, LC, LC});
  mrl_expr *fin = kernel(_e_final,
                         "Ldt := L*t; E := exp(Ldt); den := Ldt*Ldt*Ldt;"
                         "p1 := if(Ldt == 0, t, t*(-4 - 3*Ldt + E*(4 - Ldt))/den);"
                         "p2 := if(Ldt == 0, t*t/2, t*(2 + Ldt + E*(-2 + Ldt))/den);"
                         "p3 := if(Ldt == 0, t*t/6, t*(-4 - 3*Ldt - Ldt*Ldt + E*(4 - Ldt))/den);"  // dt^2/6 as coded, ETDRK4Solver.C:89
                         "E*un + p1*N1 + 2*p2*(N2 + N3) + p3*N4",
                         {"L", "un", "N1", "N2", "N3", "N4"}, {LR, LC, LC, LC, LC, LC});

  std::vector<Tensor> stage(n);
  for (std::size_t i = 0; i < n; ++i) stage[i] = run(half, {&linear[i], &ubar_n[i], &N1[i]});
  evaluate_nonlinear(stage, N2);
  for (std::size_t i = 0; i < n; ++i) stage[i] = run(half, {&linear[i], &ubar_n[i], &N2[i]});
  evaluate_nonlinear(stage, N3);
  for (std::size_t i = 0; i < n; ++i) stage[i] = run(full, {&linear[i], &ubar_n[i], &N3[i]});
  evaluate_nonlinear(stage, N4);
  for (std::size_t i = 0; i < n; ++i) _variables[i]._buffer = _domain.ifft(run(fin, {&linear[i], &ubar_n[i], &N1[i], &N2[i], &N3[i], &N4[i]}));
}

// ================================================================================== TensorPredictor
InputParameters TensorPredictor::validParams() {
  InputParameters params = MooseObject::validParams();
  params.registerBase("TensorPredictor");
  params.addPrivateParam<TensorProblem *>("_tensor_problem", nullptr);
  params.addPrivateParam<const DomainAction *>("_domain", nullptr);
  params.addClassDescription("TensorPredictor object.");
  params.addRequiredParam<TensorOutputBufferName>("buffer", "The buffer this compute is forward predicting");
  params.addParam<unsigned int>("history_size", 1, "How many old states to use (determines time integration order).");
  return params;
}

TensorPredictor::TensorPredictor(const InputParameters &parameters)
  : MooseObject(parameters),
    _tensor_problem(*getCheckedPointerParam<TensorProblem>("_tensor_problem")),
    _domain(_tensor_problem.domain()),
    _u_name(getParam<TensorOutputBufferName>("buffer")),
    _u(_tensor_problem.getBuffer(_u_name)),
    _u_old(_tensor_problem.getBufferOld(_u_name, getParam<unsigned int>("history_size"))) {}

registerMooseObject("MarlinApp", LinearTensorPredictor);

InputParameters LinearTensorPredictor::validParams() {
  InputParameters params = TensorPredictor::validParams();
  params.addParam<Real>("scale", 1.0, "The scale factor for the predictor (can range from 0 to 1)");
  params.set<unsigned int>("history_size", 2);
  return params;
}

LinearTensorPredictor::LinearTensorPredictor(const InputParameters &parameters) : TensorPredictor(parameters), _scale(getParam<Real>("scale")) {
  // one generated kernel for u + (uo0 - uo1) [* scale] (LinearTensorPredictor.C:31-36)
  if (_scale == 1.0)
    _extrapolate.configure("u + (uo0 - uo1)", {"u", "uo0", "uo1"}, {}, {}, {}, false, MRL_EXPAND_NONE);
  else
    _extrapolate.configure("u + (uo0 - uo1) * scale", {"u", "uo0", "uo1"}, {}, {"scale"}, {_scale}, false, MRL_EXPAND_NONE);
}

void LinearTensorPredictor::computeBuffer() {
  if (_u_old.size() > 1 && _u_old[0].defined() && _u_old[1].defined() && _u.defined())
    _u = _extrapolate.eval(_domain, {&_u, &_u_old[0], &_u_old[1]}, 0.0);
}
