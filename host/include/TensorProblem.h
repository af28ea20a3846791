// TensorProblem / TensorBuffer: buffer ownership, history, time bookkeeping and the execution
// order of one MOOSE time step.  Host mirror of
//   include/problems/TensorProblem.h:208,289-356 (buffer map, auto-creation, getBuffer / getBufferOld)
//   src/problems/TensorProblem.C:75-151 (init), :154-197 (execute), :199-216 (ICs),
//   :219-251 (outputs + postprocess computes), :451-472 (advanceState, quirk Q1)
//   include/tensor_buffers/TensorBuffer.h:64-116 (history ring of handles)
// The FEProblem / Transient parts of MOOSE that TensorProblem relies on (time, dt, step counter,
// postprocessor table) are carried by this class too, since MOOSE is not linked here.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "DomainAction.h"
#include "MarlinTensor.h"
#include "moose_shim.h"

class TensorOperatorBase;
class TensorSolver;
class TensorPostprocessor;
class TensorOutput;

// ---- buffers ------------------------------------------------------------------------------------
class TensorBufferBase {
public:
  explicit TensorBufferBase(const std::string &name) : _name(name) {}
  virtual ~TensorBufferBase() = default;
  const std::string &name() const { return _name; }
  virtual std::size_t advanceState() = 0;
  virtual std::size_t maxStates() const = 0;
  virtual void clearStates() = 0;
  virtual const marlin::Tensor &getRawTensor() const = 0;
  virtual void makeCPUCopy(const DomainAction &domain) = 0;
  virtual const std::vector<double> &getRawCPUTensor() = 0;

private:
  std::string _name;
};

template <typename T>
class TensorBuffer : public TensorBufferBase {
public:
  using TensorBufferBase::TensorBufferBase;
  std::size_t advanceState() override {
    if (_u_old.size() < _max_states) _u_old.resize(_u_old.size() + 1);
    if (!_u_old.empty()) {
      for (std::size_t i = _u_old.size() - 1; i > 0; --i) _u_old[i] = _u_old[i - 1];
      _u_old[0] = _u;
    }
    return _u_old.size();
  }
  void clearStates() override { _u_old.clear(); }
  std::size_t maxStates() const override { return _max_states; }
  T &getTensor() { return _u; }
  const std::vector<T> &getOldTensor(std::size_t states_requested) {
    _max_states = std::max(_max_states, states_requested);
    return _u_old;
  }
  const marlin::Tensor &getRawTensor() const override { return _u; }
  // lazily requested host copy (PlainTensorBuffer::makeCPUCopy, src/tensor_buffers/PlainTensorBuffer.C:38-52)
  void makeCPUCopy(const DomainAction &domain) override {
    if (_cpu_copy_requested && _u.defined()) _u_cpu = domain.toHost(_u);
  }
  const std::vector<double> &getRawCPUTensor() override {
    _cpu_copy_requested = true;
    return _u_cpu;
  }

protected:
  T _u;
  std::vector<double> _u_cpu;
  bool _cpu_copy_requested = false;
  std::vector<T> _u_old;
  std::size_t _max_states = 0;
};

// ---- problem ------------------------------------------------------------------------------------
class TensorProblem : public MooseObject {
public:
  static InputParameters validParams();
  explicit TensorProblem(const InputParameters &parameters);
  ~TensorProblem() override;

  const DomainAction &domain() const { return _domain; }

  // buffers (auto-created on first request)
  marlin::Tensor &getBuffer(const std::string &buffer_name);
  const std::vector<marlin::Tensor> &getBufferOld(const std::string &buffer_name, unsigned int max_states);
  TensorBufferBase &getBufferBase(const std::string &buffer_name);
  const marlin::Tensor &getRawBuffer(const std::string &buffer_name) { return getBufferBase(buffer_name).getRawTensor(); }
  const std::vector<double> &getRawCPUBuffer(const std::string &buffer_name) { return getBufferBase(buffer_name).getRawCPUTensor(); }
  bool hasBuffer(const std::string &buffer_name) const { return _tensor_buffer.count(buffer_name) != 0; }
  const std::map<std::string, std::shared_ptr<TensorBuffer<marlin::Tensor>>> &getBuffers() const { return _tensor_buffer; }

  // [Functions] of type ParsedFunction (MOOSE ParsedFunction: expression, symbol_names, symbol_values);
  // sampled by MooseFunctionTensor
  struct ParsedFunctionDesc {
    std::string expression;
    std::vector<std::string> symbol_names, symbol_values;
  };
  void addFunction(const std::string &name, ParsedFunctionDesc f) { _functions[name] = std::move(f); }
  const ParsedFunctionDesc *getFunction(const std::string &name) const {
    auto it = _functions.find(name);
    return it == _functions.end() ? nullptr : &it->second;
  }

  // scalar constants (MarlinConstantInterface: a name declared in [Problem] or a literal number)
  void declareConstant(const std::string &name, Real value) { _constants[name] = value; }
  Real getConstant(const std::string &name_or_number, const std::string &what) const;

  // object lists
  void addTensorIC(std::shared_ptr<TensorOperatorBase> op) { _ics.push_back(std::move(op)); }
  void addTensorCompute(std::shared_ptr<TensorOperatorBase> op) { _computes.push_back(std::move(op)); }
  void addTensorPostprocess(std::shared_ptr<TensorOperatorBase> op) { _pps.push_back(std::move(op)); }
  void setSolver(std::shared_ptr<TensorSolver> solver);
  void addPostprocessor(std::shared_ptr<TensorPostprocessor> pp) { _postprocessors.push_back(std::move(pp)); }
  void addTensorOutput(std::shared_ptr<TensorOutput> out) { _outputs.push_back(std::move(out)); }
  const std::vector<std::shared_ptr<TensorOperatorBase>> &getComputes() const { return _computes; }
  const std::vector<std::shared_ptr<TensorOperatorBase>> &getICs() const { return _ics; }
  const std::vector<std::shared_ptr<TensorOperatorBase>> &getPostprocessComputes() const { return _pps; }
  const std::vector<std::shared_ptr<TensorPostprocessor>> &getPostprocessors() const { return _postprocessors; }
  TensorSolver *getSolver() const { return _solver.get(); }

  // MOOSE problem protocol
  void init();
  void execute(ExecFlagType exec_type);
  void advanceState();
  void gridChanged();
  // objects that keep history outside the buffer table (fused solver plans) follow advanceState
  void addAdvanceStateHook(std::function<void()> hook) { _advance_hooks.push_back(std::move(hook)); }

private:
  std::vector<std::shared_ptr<TensorOutput>> _outputs;
  std::map<std::string, ParsedFunctionDesc> _functions;
  std::set<std::string> _extra_observed;

public:

  // time bookkeeping (FEProblemBase::time() etc.; owned here because MOOSE is not linked)
  Real &time() { return _time; }
  void waitForOutputs();  // joins the output threads (end of run; TensorProblem.C:66-72)
  const Real &outputTime() const { return _output_time; }  // time of the frame the output threads are writing
  Real &timeOld() { return _time_old; }
  Real &dt() { return _dt; }
  Real &dtOld() { return _dt_old; }
  int &timeStep() { return _t_step; }
  Real &subDt() { return _sub_dt; }
  Real &subTime() { return _sub_time; }
  bool debugOutput() const { return _debug; }

  // names of buffers read by postprocessors / outputs (a fused solver must keep those materialised)
  std::set<std::string> observedBuffers() const;
  // buffers an output object (or the driver's --dump) reads: they must stay materialised when a solver
  // fuses the computes that produce them
  void observeBuffer(const std::string &name) { _extra_observed.insert(name); }
  // deepest history any object asked for (getBufferOld): after that many + 1 substeps issued one by
  // one, every old state is what an all-single-substep run would hold
  std::size_t maxOldStates() const {
    std::size_t m = 0;
    for (const auto &pair : _tensor_buffer) m = std::max(m, pair.second->maxStates());
    return m;
  }

private:
  const DomainAction &_domain;
  const bool _debug;
  std::map<std::string, std::shared_ptr<TensorBuffer<marlin::Tensor>>> _tensor_buffer;
  std::map<std::string, Real> _constants;
  mutable std::set<std::string> _fetched_constants;  // requested by name, never declared
  std::vector<std::shared_ptr<TensorOperatorBase>> _ics, _computes, _pps;
  std::shared_ptr<TensorSolver> _solver;
  std::vector<std::shared_ptr<TensorPostprocessor>> _postprocessors;
  std::vector<Real> _old_dt;
  std::vector<std::function<void()>> _advance_hooks;
  Real _time = 0, _time_old = 0, _dt = 0, _dt_old = 0, _sub_dt = 0, _sub_time = 0, _output_time = 0;
  int _t_step = 0;
};
