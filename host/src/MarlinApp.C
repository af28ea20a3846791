// marlin_b200-opt: reads an unmodified Marlin input file and runs the spectral path on the GPU.
//
// Stands in for the pieces of MOOSE that sit around Marlin's objects and cannot be built here:
//   * syntax -> object wiring of MarlinApp::registerAll (src/base/MarlinApp.C:94-172):
//     [Domain], [TensorBuffers], [TensorComputes/{Initialize,Solve,Postprocess}] (five levels deep,
//     untyped sub-blocks become ComputeGroups: src/actions/AddTensorComputeAction.C:33-80),
//     [TensorSolver] (automatic root compute: src/actions/CreateTensorSolverAction.C:33-62),
//     [Problem], [Postprocessors], [GlobalParams]
//   * the Transient executioner's step loop with ConstantDT / IterationAdaptiveDT
//     (moose/framework/src/executioners/TransientBase.C, src/timesteppers/IterationAdaptiveDT.C:243-300,
//      TimeStepper.C:103-135)
//   * the CSV postprocessor output ([Outputs] csv = true)
// Blocks that belong to the finite-element side of MOOSE (Mesh, Variables, AuxKernels, ...) and the
// XDMF tensor output are outside the spectral time-step path; they are reported and skipped.
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <chrono>
#include <iostream>
#include <sstream>

#include "MechanicsComputes.h"
#include "TensorComputes.h"
#include "TensorPostprocessor.h"
#include "TensorOutput.h"
#include "TensorSolver.h"
#include "hit.h"

namespace {

struct Options {
  std::string input;
  std::vector<std::string> overrides;
  bool check_only = false;
  bool list_objects = false;
  bool parse_only = false;  // print the resolved input tree (no device needed)
  std::string output_dir;
  std::vector<std::string> dump;
  std::string dump_dir = ".";
  bool quiet = false;
  std::string compute_device;  // --compute-device=<dev> (moose TestHarness; src/base/MarlinInit.C:12-18)
  bool timing = false;        // --timing: device-synchronised wall time of every step's solve, on stderr
  bool allow_unused = false;  // --allow-unused / -w: unused input-file parameters are warnings (MooseApp.C:1294-1297)
};

std::string dirName(const std::string &p) {
  const size_t s = p.rfind('/');
  return s == std::string::npos ? "." : p.substr(0, s);
}
std::string baseName(const std::string &p) {
  const size_t s = p.rfind('/');
  std::string b = s == std::string::npos ? p : p.substr(s + 1);
  const size_t d = b.rfind('.');
  return d == std::string::npos ? b : b.substr(0, d);
}

class MarlinApp {
public:
  explicit MarlinApp(Options opt) : _opt(std::move(opt)) {}
  int run();

private:
  // fill an object's parameters from its block (+ GlobalParams), reject unknown keys
  InputParameters fill(const std::string &type, const hit::Node &block, const std::string &name, const std::set<std::string> &also_allowed = {});
  void addComputes(const hit::Node &parent, int task, int depth);
  void buildObjects();
  void transient();
  void writeCSVRow(bool header);
  void dumpBuffers();

  Options _opt;
  std::unique_ptr<hit::Node> _root;
  std::unique_ptr<DomainAction> _domain;
  std::shared_ptr<TensorProblem> _problem;
  std::vector<std::string> _skipped;
  std::ofstream _csv;
  std::vector<std::shared_ptr<TensorPostprocessor>> _csv_pps;
  std::vector<std::shared_ptr<TensorVectorPostprocessor>> _vpps;  // [VectorPostprocessors]
  std::vector<std::shared_ptr<TensorPredictor>> _predictors;      // [TensorSolver/Predictors/*]
  void runVectorPostprocessors(int flag, bool csv, const std::string &file_base);
};

InputParameters MarlinApp::fill(const std::string &type, const hit::Node &block, const std::string &name, const std::set<std::string> &also_allowed) {
  InputParameters p = type == "DomainAction" ? DomainAction::validParams() : Factory::instance().getValidParams(type);
  p.set<std::string>("_object_name", name);
  p.set<std::string>("_type", type);
  p.set<std::string>("_object_path", block.fullpath());
  if (const hit::Node *gp = _root->find("GlobalParams"))
    for (const hit::Node *f : gp->fields())
      if (p.have(f->name) && !p.entries().at(f->name).is_private) p.setFromInput(f->name, f->value);
  for (const hit::Node *f : block.fields()) {
    if (f->name == "type" || f->name == "active" || f->name == "inactive") {
      if (f->name == "type") p.setFromInput("type", f->value);
      continue;
    }
    if (!p.have(f->name) || p.entries().at(f->name).is_private) {
      if (also_allowed.count(f->name)) continue;
      // MOOSE's default is to error on unused parameters (MooseApp.C:477 ERROR_UNUSED, Builder.C:361-399)
      if (_opt.allow_unused) {
        std::cerr << _opt.input << ":" << f->line << ": warning: unused parameter '" << block.fullpath() << "/" << f->name << "' (object type "
                  << type << ")\n";
        continue;
      }
      mooseError(_opt.input, ":", f->line, ": unused parameter '", block.fullpath(), "/", f->name, "' (object type ", type,
                 ")\n\nAppend --allow-unused (or -w) on the command line to ignore unused parameters.");
    }
    p.setFromInput(f->name, f->value);
  }
  p.setPointer("_domain", _domain.get());
  p.setPointer("_tensor_problem", _problem.get());
  p.check(block.fullpath());
  return p;
}

// task: 0 = Initialize, 1 = Solve, 2 = Postprocess
void MarlinApp::addComputes(const hit::Node &parent, int task, int depth) {
  if (depth > 5) mooseError(parent.fullpath(), ": TensorComputes blocks may nest at most five levels deep");
  for (hit::Node *blk : parent.sections()) {
    const hit::Node *tf = blk->field("type");
    std::string type = tf ? tf->value : "ComputeGroup";
    if (!Factory::instance().isRegistered(type)) mooseError(_opt.input, ":", blk->line, ": A '", type, "' is not a registered object (block ", blk->fullpath(), ")");
    // sub-blocks first?  No: MOOSE acts on the blocks in input order, parents before children.
    InputParameters p = fill(type, *blk, blk->name);
    if (type == "ComputeGroup") {
      // automatically populate `computes` with the sub-blocks
      auto computes = p.get<std::vector<TensorComputeName>>("computes", blk->fullpath());
      std::set<TensorComputeName> s(computes.begin(), computes.end());
      for (hit::Node *child : blk->sections()) s.insert(child->name);
      p.set<std::vector<TensorComputeName>>("computes", std::vector<TensorComputeName>(s.begin(), s.end()));
    }
    auto obj = std::dynamic_pointer_cast<TensorOperatorBase>(Factory::instance().create(type, p));
    if (!obj) mooseError(blk->fullpath(), ": '", type, "' is not a TensorOperator");
    if (task == 0) _problem->addTensorIC(obj);
    if (task == 1) _problem->addTensorCompute(obj);
    if (task == 2) _problem->addTensorPostprocess(obj);
    addComputes(*blk, task, depth + 1);
  }
}

void MarlinApp::buildObjects() {
  static const std::set<std::string> known = {"Domain", "GlobalParams", "TensorBuffers", "TensorComputes", "TensorSolver", "Problem", "Postprocessors", "VectorPostprocessors", "Executioner", "Outputs", "Functions", "TensorOutputs"};
  for (hit::Node *s : _root->sections())
    if (!known.count(s->name)) _skipped.push_back(s->name);

  const hit::Node *dom = _root->find("Domain");
  if (!dom || !dom->is_section) mooseError(_opt.input, ": missing [Domain] block");
  {
    InputParameters dp = fill("DomainAction", *dom, "Domain");
    if (!_opt.compute_device.empty()) dp.set<std::string>("_cli_compute_device", _opt.compute_device);
    _domain = std::make_unique<DomainAction>(dp);
  }

  // [Problem]
  {
    hit::Node empty;
    const hit::Node *pb = _root->find("Problem");
    std::string type = "TensorProblem";
    if (pb && pb->field("type")) type = pb->field("type")->value;
    if (type != "TensorProblem") mooseError("Tensor objects are only supported if the problem class is set to `TensorProblem` (got '", type, "')");
    empty.name = "Problem";
    _problem = std::dynamic_pointer_cast<TensorProblem>(Factory::instance().create(type, fill(type, pb ? *pb : empty, "Problem")));
  }
  // [Functions]: ParsedFunction objects, sampled by MooseFunctionTensor
  if (const hit::Node *fb = _root->find("Functions"))
    for (hit::Node *b : fb->sections()) {
      const hit::Node *tf = b->field("type");
      if (!tf || tf->value != "ParsedFunction") {
        _skipped.push_back(b->fullpath() + (tf ? " (type " + tf->value + ")" : ""));
        continue;
      }
      TensorProblem::ParsedFunctionDesc f;
      const hit::Node *e = b->field("expression");
      if (!e) e = b->field("value");
      if (!e) mooseError(b->fullpath(), ": missing 'expression'");
      f.expression = e->value;
      auto list = [&](const char *key) {
        std::vector<std::string> out;
        if (const hit::Node *n = b->field(key)) out = shim_detail::Conv<std::vector<std::string>>::from(n->value, b->fullpath() + "/" + key);
        return out;
      };
      f.symbol_names = list("symbol_names");
      f.symbol_values = list("symbol_values");
      _problem->addFunction(b->name, f);
    }
  // [TensorBuffers]
  if (const hit::Node *tb = _root->find("TensorBuffers"))
    for (hit::Node *b : tb->sections()) {
      const hit::Node *tf = b->field("type");
      if (tf && tf->value != "PlainTensorBuffer") mooseError(b->fullpath(), ": buffer type '", tf->value, "' is not on the spectral path (only PlainTensorBuffer)");
      _problem->getBuffer(b->name);
    }
  // [TensorComputes]
  if (const hit::Node *tc = _root->find("TensorComputes")) {
    for (hit::Node *s : tc->sections()) {
      if (s->name == "Initialize")
        addComputes(*s, 0, 1);
      else if (s->name == "Solve")
        addComputes(*s, 1, 1);
      else if (s->name == "Postprocess")
        addComputes(*s, 2, 1);
      else if (s->name == "Boundary")
        _skipped.push_back("TensorComputes/Boundary");
      else
        mooseError(s->fullpath(), ": unknown TensorComputes section");
    }
  }
  // [TensorSolver]
  if (const hit::Node *ts = _root->find("TensorSolver")) {
    const hit::Node *tf = ts->field("type");
    if (!tf) mooseError("[TensorSolver]: missing 'type'");
    if (!Factory::instance().isRegistered(tf->value)) mooseError(_opt.input, ":", ts->line, ": A '", tf->value, "' is not a registered object (block TensorSolver)");
    InputParameters p = fill(tf->value, *ts, "TensorSolver", {"apply_predictors"});
    if (!p.isParamValid("root_compute")) {
      // CreateTensorSolverAction.C:43-60
      std::vector<TensorComputeName> names;
      for (const auto &cmp : _problem->getComputes()) names.push_back(cmp->name());
      InputParameters gp = Factory::instance().getValidParams("ComputeGroup");
      gp.set<std::string>("_object_name", "automatic_root_compute");
      gp.set<std::string>("_type", "ComputeGroup");
      gp.set<std::string>("_object_path", "TensorComputes/Solve/automatic_root_compute");
      gp.set<std::vector<TensorComputeName>>("computes", names);
      gp.setPointer("_domain", _domain.get());
      gp.setPointer("_tensor_problem", _problem.get());
      _problem->addTensorCompute(std::dynamic_pointer_cast<TensorOperatorBase>(Factory::instance().create("ComputeGroup", gp)));
      p.set<TensorComputeName>("root_compute", "automatic_root_compute");
    }
    auto solver = std::dynamic_pointer_cast<TensorSolver>(Factory::instance().create(tf->value, p));
    if (!solver) mooseError("[TensorSolver]: '", tf->value, "' is not a TensorSolver");
    _problem->setSolver(solver);
    // [TensorSolver/Predictors/*] (AddTensorPredictorAction, src/actions/AddTensorPredictorAction.C:29-42: needs an
    // iterative solver; the objects are built - which registers the old states they read - and, as in the reference,
    // handed to the solver only on request)
    for (hit::Node *s : ts->sections()) {
      if (s->name != "Predictors") {
        _skipped.push_back("TensorSolver/" + s->name);
        continue;
      }
      auto *iterative = dynamic_cast<IterativeTensorSolverInterface *>(solver.get());
      if (!iterative) mooseError("[TensorSolver/Predictors]: the solver '", tf->value, "' is not an iterative tensor solver");
      bool apply = false;
      if (const hit::Node *f = ts->field("apply_predictors")) apply = shim_detail::Conv<bool>::from(f->value, "TensorSolver/apply_predictors");
      for (hit::Node *pb : s->sections()) {
        const hit::Node *pt = pb->field("type");
        if (!pt) mooseError(pb->fullpath(), ": missing 'type'");
        if (!Factory::instance().isRegistered(pt->value)) mooseError(_opt.input, ":", pb->line, ": A '", pt->value, "' is not a registered object (block ", pb->fullpath(), ")");
        auto pred = std::dynamic_pointer_cast<TensorPredictor>(Factory::instance().create(pt->value, fill(pt->value, *pb, pb->name)));
        if (!pred) mooseError(pb->fullpath(), ": '", pt->value, "' is not a TensorPredictor");
        if (apply) iterative->addPredictor(pred);
        _predictors.push_back(pred);
      }
    }
  }
  // [TensorOutputs] (AddTensorOutputAction): XDMFTensorOutput; other types are reported and skipped
  if (const hit::Node *to = _root->find("TensorOutputs"))
    for (hit::Node *b : to->sections()) {
      const hit::Node *tf = b->field("type");
      if (!tf) mooseError(b->fullpath(), ": missing 'type'");
      if (!Factory::instance().isRegistered(tf->value)) {
        _skipped.push_back(b->fullpath() + " (type " + tf->value + ")");
        continue;
      }
      InputParameters p = fill(tf->value, *b, b->name);
      // MooseApp::getOutputFileBase: <input dir>/<input base>_<object name> unless file_base is given
      std::string base = dirName(_opt.input) + "/" + baseName(_opt.input);  // e.g. cahnhilliard.i -> cahnhilliard.xmf
      if (p.isParamValid("file_base")) base = p.get<std::string>("file_base", b->fullpath());
      if (!_opt.output_dir.empty()) base = _opt.output_dir + "/" + base.substr(base.rfind('/') + 1);
      p.set<std::string>("file_base", base);
      auto out = std::dynamic_pointer_cast<TensorOutput>(Factory::instance().create(tf->value, p));
      if (!out) mooseError(b->fullpath(), ": '", tf->value, "' is not a TensorOutput");
      _problem->addTensorOutput(out);
    }
  // [VectorPostprocessors] on tensor buffers (TensorHistogram)
  if (const hit::Node *vb = _root->find("VectorPostprocessors"))
    for (hit::Node *b : vb->sections()) {
      const hit::Node *tf = b->field("type");
      if (!tf) mooseError(b->fullpath(), ": missing 'type'");
      if (!Factory::instance().isRegistered(tf->value)) {
        _skipped.push_back(b->fullpath() + " (type " + tf->value + ")");
        continue;
      }
      auto vpp = std::dynamic_pointer_cast<TensorVectorPostprocessor>(Factory::instance().create(tf->value, fill(tf->value, *b, b->name)));
      if (!vpp) mooseError(b->fullpath(), ": '", tf->value, "' is not a VectorPostprocessor");
      _vpps.push_back(vpp);
    }
  // [Postprocessors]
  if (const hit::Node *pps = _root->find("Postprocessors"))
    for (hit::Node *b : pps->sections()) {
      const hit::Node *tf = b->field("type");
      if (!tf) mooseError(b->fullpath(), ": missing 'type'");
      if (!Factory::instance().isRegistered(tf->value)) {
        _skipped.push_back(b->fullpath() + " (type " + tf->value + ")");
        continue;
      }
      auto pp = std::dynamic_pointer_cast<TensorPostprocessor>(Factory::instance().create(tf->value, fill(tf->value, *b, b->name)));
      if (!pp) mooseError(b->fullpath(), ": '", tf->value, "' is not a Postprocessor");
      _problem->addPostprocessor(pp);
    }
}

void MarlinApp::writeCSVRow(bool header) {
  if (!_csv.is_open()) return;  // rank 0 only in parallel runs
  if (header) {
    _csv << "time";
    for (const auto &pp : _csv_pps) _csv << "," << pp->name();
    _csv << "\n";
    return;
  }
  _csv << std::setprecision(14) << _problem->time();
  for (const auto &pp : _csv_pps) _csv << "," << std::setprecision(14) << pp->getValue();
  _csv << "\n";
  _csv.flush();
}

// VectorPostprocessors scheduled for `flag`; MOOSE's CSV output writes one file per object and time step:
// <file_base>_<name>_<step, 4 digits>.csv with the vectors as columns in name order
void MarlinApp::runVectorPostprocessors(int flag, bool csv, const std::string &file_base) {
  for (auto &vpp : _vpps) {
    if (!(vpp->executeOn() & flag)) continue;
    vpp->initialize();
    vpp->execute();
    vpp->finalize();
    if (!csv || _domain->rank() != 0) continue;
    char tag[16];
    std::snprintf(tag, sizeof tag, "%04d", _problem->timeStep());
    std::ofstream f(file_base + "_" + vpp->name() + "_" + tag + ".csv");
    if (!f) mooseError("cannot write the CSV file of VectorPostprocessor '", vpp->name(), "'");
    const auto &vecs = vpp->vectors();
    std::size_t rows = 0;
    bool first = true;
    for (const auto &kv : vecs) {
      f << (first ? "" : ",") << kv.first;
      rows = std::max(rows, kv.second.size());
      first = false;
    }
    f << "\n";
    for (std::size_t r = 0; r < rows; ++r) {
      first = true;
      for (const auto &kv : vecs) {
        f << (first ? "" : ",");
        if (r < kv.second.size()) f << std::setprecision(14) << kv.second[r];
        first = false;
      }
      f << "\n";
    }
  }
}

void MarlinApp::dumpBuffers() {
  for (const auto &name : _opt.dump) {
    if (!_problem->hasBuffer(name)) mooseError("--dump: no buffer named '", name, "'");
    const marlin::Tensor &t = _problem->getRawBuffer(name);
    if (!t.defined()) mooseError("--dump: buffer '", name, "' is not defined");
    const std::vector<double> host = _domain->toHost(t);
    // little-endian float64, C order; complex tensors interleaved; rank-two fields component major
    // parallel runs: one file per rank holding its part (real space: [nx][ny_local](,[nz]))
    char tag[32] = "";
    if (_domain->nRanks() > 1) std::snprintf(tag, sizeof tag, ".rank%04u", _domain->rank());
    std::ofstream f(_opt.dump_dir + "/" + name + tag + ".f64", std::ios::binary);
    f.write(reinterpret_cast<const char *>(host.data()), std::streamsize(host.size() * sizeof(double)));
  }
}

void MarlinApp::transient() {
  hit::Node empty;
  const hit::Node *ex = _root->find("Executioner");
  if (!ex) ex = &empty;
  auto num = [&](const hit::Node *blk, const char *key, double dflt) {
    const hit::Node *f = blk ? blk->field(key) : nullptr;
    return f ? shim_detail::Conv<double>::from(f->value, blk->fullpath() + "/" + key) : dflt;
  };
  if (const hit::Node *t = ex->field("type"))
    if (t->value != "Transient") mooseError("[Executioner]: only type = Transient drives a TensorProblem (got '", t->value, "')");
  const double start_time = num(ex, "start_time", 0.0);
  const double end_time = num(ex, "end_time", 1e30);
  const double dtmax = num(ex, "dtmax", 1e30);
  const double dtmin = num(ex, "dtmin", 0.0);
  const long num_steps = (long)num(ex, "num_steps", 4294967295.0);
  double dt0 = num(ex, "dt", 1.0);
  double growth = 1.0, cutback = 0.5;
  bool iteration_feedback = false;
  unsigned int min_iterations = 0, max_iterations = 4294967295u;
  const hit::Node *ts = ex->find("TimeStepper");
  if (ts && ts->is_section) {
    const hit::Node *t = ts->field("type");
    const std::string type = t ? t->value : "ConstantDT";
    dt0 = num(ts, "dt", dt0);
    if (type == "IterationAdaptiveDT" || type == "TensorSolveIterationAdaptiveDT")
      growth = num(ts, "growth_factor", 2.0);
    else if (type != "ConstantDT")
      mooseError("[Executioner/TimeStepper]: time stepper '", type, "' is not supported by the stand-alone driver");
    if (type == "TensorSolveIterationAdaptiveDT") {
      // src/timesteppers/TensorSolveIterationAdaptiveDT.C:161-174
      iteration_feedback = dynamic_cast<IterativeTensorSolverInterface *>(_problem->getSolver()) != nullptr;
      if (!iteration_feedback) mooseError("TensorSolveIterationAdaptiveDT needs an iterative tensor solver (SecantSolver)");
      if (ts->field("min_iterations")) min_iterations = (unsigned int)num(ts, "min_iterations", 0);
      if (ts->field("max_iterations")) max_iterations = (unsigned int)num(ts, "max_iterations", 0);
      cutback = num(ts, "cutback_factor", 0.5);
    }
  }

  // [Outputs]
  const hit::Node *out = _root->find("Outputs");
  bool csv = false;
  std::string file_base = dirName(_opt.input) + "/" + baseName(_opt.input) + "_out";
  int out_on = EXEC_INITIAL | EXEC_TIMESTEP_END;
  if (out) {
    if (const hit::Node *f = out->field("csv")) csv = shim_detail::Conv<bool>::from(f->value, "Outputs/csv");
    if (const hit::Node *f = out->field("file_base")) file_base = f->value[0] == '/' ? f->value : dirName(_opt.input) + "/" + f->value;
    if (const hit::Node *f = out->field("execute_on")) out_on = parseExecFlags(f->value, "Outputs/execute_on");
    for (hit::Node *s : out->sections()) {
      const hit::Node *t = s->field("type");
      if (t && t->value == "CSV") {
        csv = true;
        if (const hit::Node *f = s->field("file_base")) file_base = f->value[0] == '/' ? f->value : dirName(_opt.input) + "/" + f->value;
        if (const hit::Node *f = s->field("execute_on")) out_on = parseExecFlags(f->value, s->fullpath() + "/execute_on");
      } else {
        _skipped.push_back(s->fullpath());
      }
    }
  }
  if (!_opt.output_dir.empty()) file_base = _opt.output_dir + "/" + file_base.substr(file_base.rfind('/') + 1);

  if (_domain->rank() != 0) _opt.quiet = true;  // one console
  if (!_skipped.empty() && !_opt.quiet) {
    std::cerr << "marlin_b200: blocks outside the spectral time-step path were skipped:";
    for (const auto &s : _skipped) std::cerr << " [" << s << "]";
    std::cerr << "\n";
  }

  for (const auto &name : _opt.dump) _problem->observeBuffer(name);
  // [TensorOutputs] are not written by this driver, but the buffers they name are part of the run's
  // observable state: keep them materialised
  if (const hit::Node *to = _root->find("TensorOutputs"))
    for (hit::Node *o : to->sections())
      if (const hit::Node *b = o->field("buffer"))
        for (const auto &name : shim_detail::Conv<std::vector<std::string>>::from(b->value, o->fullpath() + "/buffer")) _problem->observeBuffer(name);
  _problem->init();
  if (_opt.check_only) {
    std::cout << "Syntax OK\n";
    return;
  }

  _csv_pps = _problem->getPostprocessors();
  std::sort(_csv_pps.begin(), _csv_pps.end(), [](const auto &a, const auto &b) { return a->name() < b->name(); });
  if (csv && !_csv_pps.empty() && _domain->rank() == 0) {
    _csv.open(file_base + ".csv");
    if (!_csv) mooseError("cannot write '", file_base, ".csv'");
    writeCSVRow(true);
  }

  Real &time = _problem->time();
  Real &time_old = _problem->timeOld();
  Real &dt = _problem->dt();
  Real &dt_old = _problem->dtOld();
  int &t_step = _problem->timeStep();
  time = time_old = start_time;
  t_step = 0;
  _problem->execute(EXEC_INITIAL);
  runVectorPostprocessors(EXEC_INITIAL, csv && (out_on & EXEC_INITIAL), file_base);
  if (out_on & EXEC_INITIAL) writeCSVRow(false);

  double next_dt = dt0;
  while (t_step < num_steps && time < end_time - 1e-14 * std::max(1.0, std::fabs(end_time))) {
    // TimeStepper::constrainStep
    double step = next_dt;
    if (step > dtmax) step = dtmax;
    if (step < dtmin) step = dtmin;
    if (time + step > end_time) step = end_time - time;

    time_old = time;
    t_step += 1;
    _problem->advanceState();
    dt_old = t_step > 1 ? dt : step;
    dt = step;
    time = time_old + dt;
    std::chrono::steady_clock::time_point t0;
    if (_opt.timing) {
      _domain->synchronize();
      t0 = std::chrono::steady_clock::now();
    }
    _problem->execute(EXEC_TIMESTEP_BEGIN);
    if (_opt.timing) {
      _domain->synchronize();
      std::cerr << "marlin_b200: step " << t_step << " solve " << std::setprecision(9)
                << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() << " ms\n";
    }
    _problem->execute(EXEC_TIMESTEP_END);
    runVectorPostprocessors(EXEC_TIMESTEP_END, csv && (out_on & EXEC_TIMESTEP_END), file_base);
    if (out_on & EXEC_TIMESTEP_END) writeCSVRow(false);
    if (!_opt.quiet) std::cerr << "Time Step " << t_step << ", time = " << std::setprecision(8) << time << ", dt = " << dt << "\n";
    if (iteration_feedback) {
      const auto *it = dynamic_cast<IterativeTensorSolverInterface *>(_problem->getSolver());
      if (!it->isConverged()) {
        // TimeStepper::computeFailedDT: repeat the step with cutback_factor_at_failure (0.5)
        if (!_opt.quiet) std::cerr << "Solve failed, cutting timestep.\n";
        time = time_old;
        t_step -= 1;
        next_dt = dt * 0.5;
        if (next_dt < std::max(dtmin, 1e-14)) mooseError("Solve failed and timestep already at dtmin, cannot continue!");
        continue;
      }
      next_dt = dt;
      if (it->getIterations() < min_iterations) next_dt = dt * growth;
      else if (it->getIterations() > max_iterations) next_dt = dt * cutback;
    } else {
      next_dt = dt * growth;
    }
  }
  _problem->execute(EXEC_FINAL);
  if ((out_on & EXEC_FINAL) && !(out_on & EXEC_TIMESTEP_END)) writeCSVRow(false);
  _problem->waitForOutputs();
  _domain->synchronize();
  dumpBuffers();
  if (!_opt.quiet) {
    int64_t launches = 0;
    mrl_launch_count(_domain->context(), &launches);
    std::cerr << "marlin_b200: " << launches << " kernel launches, " << _domain->pool().allocations() << " device allocations ("
              << _domain->pool().bytesAllocated() / double(1 << 20) << " MiB)\n";
  }
}

int MarlinApp::run() {
  if (_opt.list_objects) {
    for (const auto &n : Factory::instance().registeredNames()) std::cout << n << "\n";
    return 0;
  }
  std::ifstream in(_opt.input);
  if (!in) mooseError("Unable to open file \"", _opt.input, "\".");
  std::stringstream ss;
  ss << in.rdbuf();
  try {
    _root = hit::parse(ss.str(), _opt.input, _opt.overrides);
  } catch (const std::exception &e) {
    mooseError(e.what());
  }
  if (_opt.parse_only) {
    // one `path = value` line per field, after ${...} expansion, overrides and active/inactive filtering
    std::function<void(const hit::Node &)> walk = [&](const hit::Node &n) {
      for (const hit::Node *f : n.fields()) std::cout << f->fullpath() << " = " << f->value << "\n";
      for (const hit::Node *s : n.sections()) {
        std::cout << "[" << s->fullpath() << "]\n";
        walk(*s);
      }
    };
    walk(*_root);
    return 0;
  }
  buildObjects();
  transient();
  // tear down in dependency order: objects -> buffers -> pool -> context
  _csv_pps.clear();
  _predictors.clear();
  _problem.reset();
  _domain.reset();
  return 0;
}

}  // namespace

// --xdmf-selftest DIR: exercises the XDMF writer on synthetic host data (no device needed; used by the
// CPU test-suite): a 3 x 2 grid, NODE / CELL / OVERSIZED_NODAL fields, two frames, with and without transpose
static int xdmfSelfTest(const std::string &dir) {
  const std::array<int64_t, 3> n = {3, 2, 1};
  const std::array<double, 3> dx = {0.5, 0.25, 1.0}, mn = {0.0, -1.0, 0.0};
  std::vector<double> c(6), mu(6), disp(2 * 12);
  for (int i = 0; i < 6; ++i) {
    c[i] = 10 + i;
    mu[i] = 20 + i;
  }
  for (int i = 0; i < 24; ++i) disp[i] = 100 + i;
  for (int tr = 0; tr < 2; ++tr) {
    XDMFWriter w(2, n, dx, mn, tr != 0, dir + (tr ? "/selftest_t" : "/selftest"));
    for (int f = 0; f < 2; ++f) {
      for (int i = 0; i < 6; ++i) c[i] = 10 + i + 100 * f;
      w.addFrame(0.001 * 3 * f, {{"c", XDMFWriter::Mode::NODE, 1, c.data()},
                                 {"disp", XDMFWriter::Mode::OVERSIZED_NODAL, 2, disp.data()},
                                 {"mu", XDMFWriter::Mode::CELL, 1, mu.data()}});
    }
  }
  // a 4 x 5 grid split along y over two ranks (2 + 3 rows): each rank writes its part, rank 0 the document
  {
    const std::array<int64_t, 3> ng = {4, 5, 1};
    std::vector<XDMFWriter::Bounds> bounds = {{{0, 0, 0}, {4, 2, 1}}, {{0, 2, 0}, {4, 5, 1}}};
    for (unsigned int r = 0; r < 2; ++r) {
      XDMFWriter w(2, ng, dx, mn, true, dir + "/selftest_par", r, bounds);
      const int64_t nyl = bounds[r].second[1] - bounds[r].first[1];
      std::vector<double> part(size_t(4 * nyl));
      for (int64_t i = 0; i < 4; ++i)
        for (int64_t j = 0; j < nyl; ++j) part[size_t(i * nyl + j)] = 100 * i + (bounds[r].first[1] + j);
      for (int f = 0; f < 2; ++f) w.addFrame(0.5 * f, {{"c", XDMFWriter::Mode::CELL, 1, part.data()}});
    }
  }
  // the same fields as `selftest`, stored as HDF5 datasets; and a file with enough datasets for a two-level group B-tree
  {
    XDMFWriter w(2, n, dx, mn, true, dir + "/selftest_h5", 0, {}, true);
    for (int f = 0; f < 2; ++f) {
      for (int i = 0; i < 6; ++i) c[i] = 10 + i + 100 * f;
      w.addFrame(0.001 * 3 * f, {{"c", XDMFWriter::Mode::NODE, 1, c.data()},
                                 {"disp", XDMFWriter::Mode::OVERSIZED_NODAL, 2, disp.data()},
                                 {"mu", XDMFWriter::Mode::CELL, 1, mu.data()}});
    }
    H5LiteFile many(dir + "/selftest_many.h5");
    std::vector<float> v(24);
    for (int k = 0; k < 300; ++k) {
      for (int i = 0; i < 24; ++i) v[i] = float(k) + 0.5f * float(i);
      many.addDataset("field_" + std::to_string(k) + ".0", {2, 3, 4}, 4, v.data());
      if (k % 50 == 0) many.flush();
    }
  }
  return 0;
}

// --comm-selftest: exercises the process-group rendezvous of the stand-alone driver (host/shim/comm.h) without a device:
// every rank prints what allgather / allreduce / barrier return (used by the CPU test-suite with several processes)
static int commSelfTest() {
  Comm &c = Comm::world();
  const int r = c.rank(), n = c.size();
  std::vector<int32_t> all(size_t(3) * n);
  const int32_t mine[3] = {r, 10 * r, 7};
  c.allgather(mine, sizeof mine, all.data());
  double v[3] = {double(r + 1), double(r + 1), double(r + 1)};
  c.allreduce(&v[0], 1, Comm::SUM);
  c.allreduce(&v[1], 1, Comm::MIN);
  c.allreduce(&v[2], 1, Comm::MAX);
  std::vector<double> big(1000, 0.001 * (r + 1));
  c.allreduce(big.data(), big.size(), Comm::SUM);
  c.barrier();
  std::cout << "rank " << r << " of " << n << " local " << c.localRank() << " gathered";
  for (int32_t x : all) std::cout << ' ' << x;
  std::cout << " sum " << v[0] << " min " << v[1] << " max " << v[2] << " big " << std::setprecision(12) << big[999] << "\n";
  return 0;
}

int main(int argc, char **argv) {
  Options opt;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto need = [&](const char *what) -> std::string {
      if (i + 1 >= argc) {
        std::cerr << "missing value after " << what << "\n";
        std::exit(2);
      }
      return argv[++i];
    };
    if (a == "-i")
      opt.input = need("-i");
    else if (a == "--check-input" || a == "--check")
      opt.check_only = true;
    else if (a == "--list-objects")
      opt.list_objects = true;
    else if (a == "--parse-only")
      opt.parse_only = true;
    else if (a == "--output-dir")
      opt.output_dir = need("--output-dir");
    else if (a == "--dump") {
      std::stringstream s(need("--dump"));
      std::string w;
      while (std::getline(s, w, ',')) opt.dump.push_back(w);
    } else if (a == "--dump-dir")
      opt.dump_dir = need("--dump-dir");
    else if (a == "--xdmf-selftest")
      return xdmfSelfTest(need("--xdmf-selftest"));
    else if (a == "--comm-selftest") {
      try {
        return commSelfTest();
      } catch (const std::exception &e) {
        std::cerr << "\n*** ERROR ***\n" << e.what() << "\n";
        return 1;
      }
    }
    else if (a == "--smooth-rectangle-expr") {  // DIM WIDTH PROFILE: the kernel expression (CPU test-suite)
      const unsigned int dim = std::stoul(need("--smooth-rectangle-expr"));
      const double w = std::stod(need("--smooth-rectangle-expr"));
      std::cout << SmoothRectangleCompute::expression(dim, w, need("--smooth-rectangle-expr")) << "\n";
      return 0;
    } else if (a == "--quiet")
      opt.quiet = true;
    else if (a == "--timing")
      opt.timing = true;
    else if (a == "--n-threads" || a == "--color")
      need(a.c_str());
    else if (a.rfind("--compute-device=", 0) == 0)
      opt.compute_device = a.substr(17);
    else if (a.rfind("--n-threads=", 0) == 0)
      continue;
    else if (a.find('=') != std::string::npos)
      opt.overrides.push_back(a);
    else if (a == "--allow-unused" || a == "-w")
      opt.allow_unused = true;
    else if (a == "--error" || a == "--error-unused" || a == "-e")
      continue;
    else {
      std::cerr << "unknown argument '" << a << "'\nusage: marlin_b200-opt -i input.i [Block/param=value ...] [--check-input] [--output-dir DIR] [--dump buf1,buf2 --dump-dir DIR]\n";
      return 2;
    }
  }
  if (opt.input.empty() && !opt.list_objects) {
    std::cerr << "usage: marlin_b200-opt -i input.i [Block/param=value ...] [--check-input] [--list-objects]\n";
    return 2;
  }
  try {
    MarlinApp app(opt);
    return app.run();
  } catch (const std::exception &e) {
    std::cerr << "\n*** ERROR ***\n" << e.what() << "\n";
    return 1;
  }
}
