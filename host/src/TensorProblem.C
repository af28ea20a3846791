#include "TensorProblem.h"

#include <cstdlib>

#include "TensorOperatorBase.h"
#include "TensorPostprocessor.h"
#include "TensorOutput.h"
#include "TensorSolver.h"

using marlin::Tensor;

registerMooseObject("MarlinApp", TensorProblem);

InputParameters TensorProblem::validParams() {
  InputParameters params = MooseObject::validParams();
  params.addClassDescription("A normal Problem object that adds the ability to perform spectral solves.");
  params.registerBase("Problem");
  params.addParam<bool>("print_debug_output", false, "Show Tensor specific debug outputs");
  params.addParam<unsigned int>("spectral_solve_substeps", 1, "How many substeps to divide the spectral solve for each MOOSE timestep into.");
  params.addParam<std::vector<std::string>>("scalar_constant_names", "Scalar constant names");
  params.addParam<std::vector<Real>>("scalar_constant_values", "Scalar constant values");
  // FEProblem parameters that appear in Marlin inputs and have no meaning without a mesh
  params.addParam<bool>("solve", true, "unused (FEProblem)");
  params.addParam<bool>("kernel_coverage_check", false, "unused (FEProblem)");
  params.addParam<bool>("skip_nl_system_check", true, "unused (FEProblem)");
  params.addPrivateParam<const DomainAction *>("_domain", nullptr);
  return params;
}

TensorProblem::TensorProblem(const InputParameters &parameters)
  : MooseObject(parameters), _domain(*getCheckedPointerParam<const DomainAction>("_domain")), _debug(getParam<bool>("print_debug_output")) {
  for (const auto &[name, value] : getParam<std::string, Real>("scalar_constant_names", "scalar_constant_values")) declareConstant(name, value);
}

void TensorProblem::waitForOutputs() {
  for (auto &out : _outputs) out->waitForCompletion();
}

TensorProblem::~TensorProblem() {
  // the output threads read the buffers' CPU copies (TensorProblem.C:66-72)
  try {
    waitForOutputs();
  } catch (const std::exception &e) {
    std::cerr << "marlin_b200: output failed: " << e.what() << "\n";
  }
  _outputs.clear();
  // operators and buffers hold device memory that must go back to the pool before the domain dies
  _solver.reset();
  _postprocessors.clear();
  _pps.clear();
  _computes.clear();
  _ics.clear();
  _tensor_buffer.clear();
}

Tensor &TensorProblem::getBuffer(const std::string &buffer_name) {
  auto it = _tensor_buffer.find(buffer_name);
  if (it == _tensor_buffer.end()) {
    if (_debug) mooseInfo("Automatically adding tensor '", buffer_name, "'");
    it = _tensor_buffer.emplace(buffer_name, std::make_shared<TensorBuffer<Tensor>>(buffer_name)).first;
  }
  return it->second->getTensor();
}

const std::vector<Tensor> &TensorProblem::getBufferOld(const std::string &buffer_name, unsigned int max_states) {
  getBuffer(buffer_name);
  return _tensor_buffer.at(buffer_name)->getOldTensor(max_states);
}

TensorBufferBase &TensorProblem::getBufferBase(const std::string &buffer_name) {
  getBuffer(buffer_name);
  return *_tensor_buffer.at(buffer_name);
}

Real TensorProblem::getConstant(const std::string &name_or_number, const std::string &what) const {
  auto it = _constants.find(name_or_number);
  if (it != _constants.end()) return it->second;
  const char *b = name_or_number.c_str();
  char *e = nullptr;
  const double v = std::strtod(b, &e);
  if (e == b || *e) {
    // reported at EXEC_INITIAL, all at once, like the reference (src/problems/TensorProblem.C:157-165)
    _fetched_constants.insert(name_or_number);
    return 0.0;
  }
  return v;
}

void TensorProblem::setSolver(std::shared_ptr<TensorSolver> solver) { _solver = std::move(solver); }

std::set<std::string> TensorProblem::observedBuffers() const {
  std::set<std::string> out = _extra_observed;
  for (const auto &pp : _postprocessors) out.insert(pp->bufferName());
  for (const auto &pp : _pps)
    for (const auto &n : pp->getRequestedItems()) out.insert(n);
  return out;
}

void TensorProblem::gridChanged() {
  for (auto &op : _ics) op->gridChanged();
  for (auto &op : _computes) op->gridChanged();
  for (auto &op : _pps) op->gridChanged();
}

// TensorProblem::init, src/problems/TensorProblem.C:75-151
void TensorProblem::init() {
  gridChanged();
  for (auto &initializer : _ics) initializer->init();
  for (auto &cmp : _computes) cmp->init();
  for (auto &pp : _pps) pp->init();

  if (_solver)
    _solver->updateDependencies();
  else
    DependencyResolverInterface::sort(_computes);
  DependencyResolverInterface::sort(_ics);
  DependencyResolverInterface::sort(_pps);

  if (_debug) {
    std::cerr << "Compute object execution order:\n";
    for (auto &cmp : _computes) {
      std::cerr << "  " << cmp->name() << '\n';
      for (const auto &ri : cmp->getRequestedItems()) std::cerr << "    <- " << ri << '\n';
      for (const auto &si : cmp->getSuppliedItems()) std::cerr << "    -> " << si << '\n';
    }
  }
  for (auto &out : _outputs) out->init();
  for (auto &cmp : _computes) cmp->check();
  if (_solver) static_cast<TensorOperatorBase *>(_solver.get())->check();
}

// TensorProblem::execute, src/problems/TensorProblem.C:154-197
void TensorProblem::execute(ExecFlagType exec_type) {
  // executeTensorOutputs (src/problems/TensorProblem.C:219-248): postprocess computes, CPU copies of the
  // buffers the outputs asked for, then the outputs scheduled for this flag
  auto run_pps = [&]() {
    for (auto &pp : _pps) pp->computeBuffer();
    // wait for the threads still writing the previous frame, then refresh the CPU copies (the device
    // synchronisation point) and start this frame's outputs in their threads
    for (auto &out : _outputs) out->waitForCompletion();
    _output_time = _time;
    for (auto &kv : _tensor_buffer) kv.second->makeCPUCopy(_domain);
    for (auto &out : _outputs)
      if (out->shouldRun(exec_type)) out->startOutput();
  };
  if (exec_type == EXEC_INITIAL) {
    if (!_fetched_constants.empty()) {
      std::string names;
      for (const auto &n : _fetched_constants) names += (names.empty() ? "" : ", ") + n;
      ::mooseError(_fetched_constants.size() == 1 ? "Constant " : "Constants ", names, _fetched_constants.size() == 1 ? " was" : " were",
                   " requested but never declared.");
    }
    _sub_time = _time;
    for (auto &ic : _ics) ic->computeBuffer();
    run_pps();
  }
  if (exec_type == EXEC_TIMESTEP_BEGIN) {
    _sub_time = _time_old;
    if (_solver)
      static_cast<TensorOperatorBase *>(_solver.get())->computeBuffer();
    else
      for (auto &cmp : _computes) cmp->computeBuffer();
  }
  if (exec_type == EXEC_TIMESTEP_END) run_pps();

  // MOOSE postprocessors scheduled for this flag (FEProblem::execute)
  for (auto &pp : _postprocessors)
    if (pp->executeOn() & exec_type) {
      pp->initialize();
      pp->execute();
      pp->finalize();
    }
}

// TensorProblem::advanceState, src/problems/TensorProblem.C:451-472.  No history is recorded while
// timeStep() <= 1 (quirk Q1 of SURVEY.md 8a).
void TensorProblem::advanceState() {
  if (_t_step <= 1) return;
  std::size_t total_max = 0;
  for (auto &pair : _tensor_buffer) total_max = std::max(total_max, pair.second->advanceState());
  if (_old_dt.size() < total_max) _old_dt.push_back(0.0);
  if (!_old_dt.empty()) {
    for (std::size_t i = _old_dt.size() - 1; i > 0; --i) _old_dt[i] = _old_dt[i - 1];
    _old_dt[0] = _dt;
  }
  for (auto &hook : _advance_hooks) hook();
}

// ================================================================================ TensorOperatorBase
InputParameters TensorOperatorBase::validParams() {
  InputParameters params = MooseObject::validParams();
  params.registerBase("TensorOperator");
  params.addPrivateParam<TensorProblem *>("_tensor_problem", nullptr);
  params.addPrivateParam<const DomainAction *>("_domain", nullptr);
  params.addClassDescription("TensorOperatorBase object.");
  return params;
}

TensorOperatorBase::TensorOperatorBase(const InputParameters &parameters)
  : MooseObject(parameters),
    _tensor_problem(*getCheckedPointerParam<TensorProblem>("_tensor_problem")),
    _domain(*getCheckedPointerParam<const DomainAction>("_domain")),
    _time(_tensor_problem.subTime()),
    _dim(_domain.getDim()) {}

void TensorOperatorBase::realSpaceComputeBuffer() { mooseError("This compute does not support real space operations."); }

const Tensor &TensorOperatorBase::getInputBuffer(const std::string &param, unsigned int ghost_layers) {
  return getInputBufferByName(getParam<TensorInputBufferName>(param), ghost_layers);
}
const Tensor &TensorOperatorBase::getInputBufferByName(const TensorInputBufferName &buffer_name, unsigned int) {
  _requested_buffers.insert(buffer_name);
  return _tensor_problem.getBuffer(buffer_name);
}
Tensor &TensorOperatorBase::getOutputBuffer(const std::string &param) { return getOutputBufferByName(getParam<TensorOutputBufferName>(param)); }
Tensor &TensorOperatorBase::getOutputBufferByName(const TensorOutputBufferName &buffer_name) {
  _supplied_buffers.insert(buffer_name);
  return _tensor_problem.getBuffer(buffer_name);
}
TensorOperatorBase &TensorOperatorBase::getCompute(const std::string &param_name) {
  const auto name = getParam<TensorComputeName>(param_name);
  for (const auto &cmp : _tensor_problem.getComputes())
    if (cmp->name() == name) return *cmp;
  paramError(param_name, "Compute not found.");
}
void TensorOperatorBase::checkC(int rc, const char *what) const {
  if (rc != MRL_OK) mooseError("marlin_b200: ", what, " failed: ", mrl_last_error());
}

// ========================================================================================= ComputeGroup
registerMooseObject("MarlinApp", ComputeGroup);

InputParameters ComputeGroup::validParams() {
  InputParameters params = TensorOperatorBase::validParams();
  params.addClassDescription("Group of operators with internal dependency resolution.");
  params.addParam<std::vector<TensorComputeName>>("computes", {}, "List of grouped tensor computes.");
  params.addParam<bool>("enable_jit", false, "Accepted for input compatibility; fusion is done by hand-written kernels, not by tracing.");
  return params;
}

ComputeGroup::ComputeGroup(const InputParameters &parameters) : TensorOperatorBase(parameters) {}

void ComputeGroup::init() {
  const auto computes = getParam<std::vector<TensorComputeName>>("computes");
  std::set<TensorComputeName> requested(computes.begin(), computes.end());
  // the list may address ICs, solve computes or postprocess computes, depending on where the group lives
  const auto &lists = {&_tensor_problem.getComputes(), &_tensor_problem.getICs(), &_tensor_problem.getPostprocessComputes()};
  for (const auto *list : lists) {
    bool own_list = false;
    for (const auto &cmp : *list)
      if (cmp.get() == this) own_list = true;
    if (!own_list) continue;
    for (const auto &cmp : *list)
      if (requested.count(cmp->name()) && cmp.get() != this) _computes.push_back(cmp);
  }
  if (_computes.size() != requested.size()) {
    for (const auto &n : requested) {
      bool found = false;
      for (const auto &c : _computes) found = found || c->name() == n;
      if (!found) paramError("computes", "Compute '", n, "' not found.");
    }
  }
}

void ComputeGroup::computeBuffer() {
  for (std::size_t i = 0; i < _computes.size(); ++i) {
    if (_domain.debug())
      for (const auto &buffer_name : _checked_tensors[i])
        if (!_tensor_problem.getRawBuffer(buffer_name).defined())
          mooseError("The tensor '", buffer_name, "' requested by '", _computes[i]->name(), "' is not defined yet. Initialize it first.");
    const auto &cmp = _computes[i];
    try {
      cmp->computeBuffer();
    } catch (const MooseException &) {
      throw;
    } catch (const std::exception &e) {
      cmp->mooseError("Exception: ", e.what());
    }
  }
  _compute_count++;
}

void ComputeGroup::updateDependencies() {
  if (!_visited)
    _visited = true;
  else
    paramError("computes", "Compute is using itself, creating an unresolvable dependency.");
  for (const auto &cmp : _computes) cmp->updateDependencies();
  DependencyResolverInterface::sort(_computes);

  std::set<std::string> in, out;
  _checked_tensors.clear();
  for (const auto &cmp : _computes) {
    const auto &cin = cmp->getRequestedItems();
    const auto &cout = cmp->getSuppliedItems();
    in.insert(cin.begin(), cin.end());
    out.insert(cout.begin(), cout.end());
    _checked_tensors.emplace_back(cin.begin(), cin.end());
  }
  _requested_buffers.clear();
  _supplied_buffers.clear();
  for (const auto &n : in)
    if (!out.count(n)) _requested_buffers.insert(n);
  for (const auto &n : out)
    if (!in.count(n)) _supplied_buffers.insert(n);
}
