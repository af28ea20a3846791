// TensorOutput / XDMFTensorOutput (reference citations in include/TensorOutput.h).
#include "TensorOutput.h"

#include <algorithm>
#include <cstdio>
#include <fstream>

InputParameters TensorOutput::validParams() {
  InputParameters params = MooseObject::validParams();
  params.addRequiredParam<std::vector<TensorInputBufferName>>("buffer", "The buffers to output");
  params.addParam<std::string>("file_base", "The desired solution output name without an extension.");
  params.addParam<std::string>("execute_on", "INITIAL TIMESTEP_END", "When to write (INITIAL, TIMESTEP_END)");
  params.registerBase("TensorOutput");
  params.addPrivateParam<TensorProblem *>("_tensor_problem", nullptr);
  params.addPrivateParam<std::string>("_default_file_base", "");
  params.addClassDescription("TensorOutput object.");
  return params;
}

TensorOutput::TensorOutput(const InputParameters &parameters)
  : MooseObject(parameters),
    _tensor_problem(*getCheckedPointerParam<TensorProblem>("_tensor_problem")),
    _domain(_tensor_problem.domain()),
    _time(_tensor_problem.outputTime()),
    _file_base(isParamValid("file_base") ? getParam<std::string>("file_base") : getParam<std::string>("_default_file_base")),
    _execute_on(parseExecFlags(getParam<std::string>("execute_on"), _path + "/execute_on")) {
  auto names = getParam<std::vector<TensorInputBufferName>>("buffer");
  std::sort(names.begin(), names.end());
  names.erase(std::unique(names.begin(), names.end()), names.end());
  for (const auto &name : names) {
    TensorBufferBase &b = _tensor_problem.getBufferBase(name);
    _out_buffers.push_back(Source{name, &b, &b.getRawCPUTensor()});
  }
}

TensorOutput::~TensorOutput() {
  try {
    waitForCompletion();
  } catch (const std::exception &e) {
    std::cerr << "marlin_b200: output '" << name() << "' failed: " << e.what() << "\n";
  }
}

void TensorOutput::startOutput() {
  prepareForOutput();
  if (_output_thread.joinable())
    mooseError("Output thread is already running. Must call waitForCompletion() first. This is a code error.");
  _output_thread = std::thread([this]() {
    try {
      output();
    } catch (...) {
      _thread_error = std::current_exception();
    }
  });
}

void TensorOutput::waitForCompletion() {
  if (_output_thread.joinable()) _output_thread.join();
  if (_thread_error) {
    std::exception_ptr e = _thread_error;
    _thread_error = nullptr;
    std::rethrow_exception(e);
  }
}

// ================================================================================================= XDMFWriter
namespace {
std::string join(const std::vector<int64_t> &v) {
  std::ostringstream os;
  for (std::size_t i = 0; i < v.size(); ++i) os << (i ? " " : "") << v[i];
  return os.str();
}
std::string joinReal(const std::vector<double> &v) {  // Moose::stringify(vector<Real>, " "): operator<< of an ostringstream
  std::ostringstream os;
  for (std::size_t i = 0; i < v.size(); ++i) os << (i ? " " : "") << v[i];
  return os.str();
}
std::string g17(double v) {  // pugixml's attribute = double
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  return buf;
}
}  // namespace

XDMFWriter::XDMFWriter(unsigned int dim, const std::array<int64_t, 3> &n, const std::array<double, 3> &dx, const std::array<double, 3> &min,
                       bool transpose, std::string file_base, unsigned int rank, std::vector<Bounds> bounds, bool enable_hdf5)
  : _dim(dim), _n(n), _dx(dx), _min(min), _rank(rank), _bounds(std::move(bounds)), _transpose(transpose), _file_base(std::move(file_base)) {
  if (dim != 2 && dim != 3) ::mooseError("XDMFTensorOutput: Unsupported tensor dimension");
  if (enable_hdf5) {
    try {
      _h5 = std::make_unique<H5LiteFile>(hdf5FileName(_rank));  // H5Fcreate(..., H5F_ACC_TRUNC, ...), :206-216
    } catch (const std::exception &e) {
      ::mooseError(e.what());
    }
  }
  if (parallel()) {
    if (_rank >= _bounds.size()) ::mooseError("XDMFWriter: rank ", _rank, " outside the ", _bounds.size(), " parts");
    for (unsigned int d = 0; d < 3; ++d) _n[d] = d < dim ? _bounds[_rank].second[d] - _bounds[_rank].first[d] : 1;
  }
  std::vector<int64_t> cells, nodes;
  std::vector<double> origin, dgrid;
  for (unsigned int i = 0; i < dim; ++i) {
    const unsigned int j = transpose ? dim - i - 1 : i;  // mappedAxis (:671-674)
    cells.push_back(n[j]);
    nodes.push_back(n[j] + 1);
    dgrid.push_back(dx[j]);
    origin.push_back(min[j]);
  }
  _cell_dims = join(cells);
  _node_dims = join(nodes);
  static const char *dxyz[] = {"DX", "DY", "DZ"};
  std::string geometry = "ORIGIN_";
  for (unsigned int i = 0; i < dim; ++i) geometry += dxyz[i];
  const std::string sdim = std::to_string(dim);
  std::ostringstream h;
  h << "<?xml version=\"1.0\"?>\n"
    << "<Xdmf xmlns:xi=\"http://www.w3.org/2003/XInclude\" Version=\"2.2\">\n"
    << "\t<Domain>\n"
    << "\t\t<Topology TopologyType=\"" << sdim << "DCoRectMesh\" Dimensions=\"" << _node_dims << "\" />\n"
    << "\t\t<Geometry Type=\"" << geometry << "\">\n"
    << "\t\t\t<DataItem Format=\"XML\" Dimensions=\"" << sdim << "\">" << joinReal(origin) << "</DataItem>\n"
    << "\t\t\t<DataItem Format=\"XML\" Dimensions=\"" << sdim << "\">" << joinReal(dgrid) << "</DataItem>\n"
    << "\t\t</Geometry>\n"
    << "\t\t<Grid Name=\"TimeSeries\" GridType=\"Collection\" CollectionType=\"Temporal\">\n";
  _head = h.str();
}

std::string XDMFWriter::xml() const {
  if (_frames.empty()) {
    std::string h = _head;
    h.replace(h.rfind(">\n"), 2, " />\n");  // an empty element, as pugixml prints it
    return h + "\t</Domain>\n</Xdmf>\n";
  }
  return _head + _frames + "\t\t</Grid>\n\t</Domain>\n</Xdmf>\n";
}

std::vector<std::string> XDMFWriter::attributeNames(const std::string &buffer_name, int64_t num_fields) {
  static const char *xyz[] = {"x", "y", "z"};
  std::vector<std::string> names;
  for (int64_t i = 0; i < num_fields; ++i) {
    std::string name = buffer_name;
    if (num_fields > 1) name += "_" + (num_fields <= 3 ? std::string(xyz[i]) : std::to_string(i));
    names.push_back(name);
  }
  return names;
}

// one component on the output grid: periodic continuation for NODE (extendTensor), then the x<->y (2-D) or
// x<->z (3-D) transpose
std::vector<double> XDMFWriter::arrange(const Field &f, int component, std::vector<uint64_t> *dims) const {
  std::array<int64_t, 3> in = {1, 1, 1}, ext = {1, 1, 1};
  for (unsigned int d = 0; d < _dim; ++d) {
    in[d] = f.mode == Mode::OVERSIZED_NODAL ? _n[d] + 1 : _n[d];
    ext[d] = f.mode == Mode::CELL ? _n[d] : _n[d] + 1;
  }
  const int64_t count_in = in[0] * in[1] * in[2];
  const double *src = f.data + (size_t)component * count_in;
  std::array<int64_t, 3> out = ext;
  if (_transpose) std::swap(out[0], out[_dim - 1]);
  if (dims) dims->assign(out.begin(), out.begin() + _dim);
  std::vector<double> dst((size_t)(ext[0] * ext[1] * ext[2]));
  for (int64_t i = 0; i < ext[0]; ++i)
    for (int64_t j = 0; j < ext[1]; ++j)
      for (int64_t k = 0; k < ext[2]; ++k) {
        const int64_t si = i % in[0], sj = j % in[1], sk = k % in[2];  // index n wraps to 0 (NODE); identity otherwise
        const double v = src[(si * in[1] + sj) * in[2] + sk];
        std::array<int64_t, 3> o = {i, j, k};
        if (_transpose) std::swap(o[0], o[_dim - 1]);
        dst[(size_t)((o[0] * out[1] + o[1]) * out[2] + o[2])] = v;
      }
  return dst;
}

std::string XDMFWriter::dataItem(const std::string &dims, const std::string &dataset, unsigned int rank) const {
  if (_h5) return "<DataItem DataType=\"Float\" Dimensions=\"" + dims + "\" Format=\"HDF\">" + hdf5FileName(rank) + ":/" + dataset + "</DataItem>\n";
  return "<DataItem DataType=\"Float\" Dimensions=\"" + dims + "\" Format=\"Binary\" Endian=\"Little\" Precision=\"8\">" + binaryFileName(dataset, rank) +
         "</DataItem>\n";
}

std::string XDMFWriter::rankTag(unsigned int rank) const {
  if (!parallel()) return "";
  char buf[32];
  std::snprintf(buf, sizeof buf, ".rank%04u", rank);
  return buf;
}

// writeLocalData (:266-355), then the frame's XML by rank 0
void XDMFWriter::addFrame(double time, const std::vector<Field> &fields) {
  for (const Field &f : fields) {
    if (parallel() && f.mode != Mode::CELL) ::mooseError("XDMFTensorOutput currently supports only CELL output mode in parallel.");
    const auto names = attributeNames(f.name, f.ncomp);
    for (int c = 0; c < f.ncomp; ++c) {
      const std::string setname = names[c] + "." + std::to_string(_frame);
      std::vector<uint64_t> dims;
      const std::vector<double> data = arrange(f, c, &dims);
      if (_h5) {
        try {
          _h5->addDataset(setname, dims, 8, data.data());  // addDataToHDF5 (:572-651): one deflate-9 chunk
        } catch (const std::exception &e) {
          ::mooseError(e.what());
        }
        continue;
      }
      const std::string fname = binaryFileName(setname, _rank);
      std::ofstream file(fname, std::ios::out | std::ios::binary);
      if (!file) ::mooseError("XDMFTensorOutput: cannot write '", fname, "'");
      file.write(reinterpret_cast<const char *>(data.data()), std::streamsize(data.size() * sizeof(double)));
    }
  }
  if (_h5) _h5->flush();  // H5Fflush (:257-260)
  if (!parallel() || _rank == 0) {
    _frames += parallel() ? parallelFrame(time, fields) : serialFrame(time, fields);
    std::ofstream x(_file_base + ".xmf");
    if (!x) ::mooseError("XDMFTensorOutput: cannot write '", _file_base, ".xmf'");
    x << xml();
  }
  _frame++;
}

// writeParallelXMF (:429-527): a spatial collection with one uniform sub-grid per rank
std::string XDMFWriter::parallelFrame(double time, const std::vector<Field> &fields) const {
  static const char *dxyz[] = {"DX", "DY", "DZ"};
  std::string geometry = "ORIGIN_";
  for (unsigned int i = 0; i < _dim; ++i) geometry += dxyz[i];
  const std::string sdim = std::to_string(_dim);
  std::vector<double> spacing;
  for (unsigned int i = 0; i < _dim; ++i) spacing.push_back(_dx[_transpose ? _dim - i - 1 : i]);
  std::ostringstream g;
  g << "\t\t\t<Grid Name=\"T" << _frame << "\" GridType=\"Collection\" CollectionType=\"Spatial\">\n"
    << "\t\t\t\t<Time Value=\"" << g17(time) << "\" />\n";
  for (unsigned int r = 0; r < _bounds.size(); ++r) {
    std::vector<int64_t> cells, nodes;
    std::vector<double> origin;
    for (unsigned int i = 0; i < _dim; ++i) {
      const unsigned int j = _transpose ? _dim - i - 1 : i;
      cells.push_back(_bounds[r].second[j] - _bounds[r].first[j]);
      nodes.push_back(cells.back() + 1);
      origin.push_back(_min[j] + _bounds[r].first[j] * _dx[j]);
    }
    g << "\t\t\t\t<Grid Name=\"Rank" << r << "\" GridType=\"Uniform\">\n"
      << "\t\t\t\t\t<Topology TopologyType=\"" << sdim << "DCoRectMesh\" Dimensions=\"" << join(nodes) << "\" />\n"
      << "\t\t\t\t\t<Geometry Type=\"" << geometry << "\">\n"
      << "\t\t\t\t\t\t<DataItem Format=\"XML\" Dimensions=\"" << sdim << "\">" << joinReal(origin) << "</DataItem>\n"
      << "\t\t\t\t\t\t<DataItem Format=\"XML\" Dimensions=\"" << sdim << "\">" << joinReal(spacing) << "</DataItem>\n"
      << "\t\t\t\t\t</Geometry>\n";
    for (const Field &f : fields) {
      const auto names = attributeNames(f.name, f.ncomp);
      for (int c = 0; c < f.ncomp; ++c)
        g << "\t\t\t\t\t<Attribute Name=\"" << names[c] << "\" Center=\"Cell\">\n"
          << "\t\t\t\t\t\t" << dataItem(join(cells), names[c] + "." + std::to_string(_frame), r)
          << "\t\t\t\t\t</Attribute>\n";
    }
    g << "\t\t\t\t</Grid>\n";
  }
  g << "\t\t\t</Grid>\n";
  return g.str();
}

// writeSerialXMF (:358-426)
std::string XDMFWriter::serialFrame(double time, const std::vector<Field> &fields) const {
  std::ostringstream g;
  g << "\t\t\t<Grid Name=\"T" << _frame << "\" GridType=\"Uniform\">\n"
    << "\t\t\t\t<Time Value=\"" << g17(time) << "\" />\n"
    << "\t\t\t\t<xi:include xpointer=\"xpointer(//Xdmf/Domain/Topology)\" />\n"
    << "\t\t\t\t<xi:include xpointer=\"xpointer(//Xdmf/Domain/Geometry)\" />\n";
  for (const Field &f : fields) {
    const bool is_cell = f.mode == Mode::CELL;
    const auto names = attributeNames(f.name, f.ncomp);
    for (int c = 0; c < f.ncomp; ++c) {
      const std::string dataset = names[c] + "." + std::to_string(_frame);
      g << "\t\t\t\t<Attribute Name=\"" << names[c] << "\" Center=\"" << (is_cell ? "Cell" : "Node") << "\">\n"
        << "\t\t\t\t\t" << dataItem(is_cell ? _cell_dims : _node_dims, dataset, 0)
        << "\t\t\t\t</Attribute>\n";
    }
  }
  g << "\t\t\t</Grid>\n";
  return g.str();
}

// ============================================================================================ XDMFTensorOutput
registerMooseObject("MarlinApp", XDMFTensorOutput);

InputParameters XDMFTensorOutput::validParams() {
  InputParameters params = TensorOutput::validParams();
  params.addClassDescription("Output a tensor in XDMF format.");
  params.addParam<bool>("enable_hdf5", false, "Use HDF5 for binary data storage.");
  params.addParam<std::vector<std::string>>("output_mode", {}, "Output as cell or node data (CELL NODE OVERSIZED_NODAL), one entry per buffer");
  params.addParam<bool>("transpose", true,
                        "The Paraview XDMF reader swaps x-y (x-z in 3d), so we transpose the tensors before we output to make the data look right in Paraview.");
  return params;
}

XDMFTensorOutput::XDMFTensorOutput(const InputParameters &parameters)
  : TensorOutput(parameters), _transpose(getParam<bool>("transpose")), _enable_hdf5(getParam<bool>("enable_hdf5")) {
  auto modes = getParam<std::vector<std::string>>("output_mode");
  const auto names = getParam<std::vector<TensorInputBufferName>>("buffer");
  if (modes.empty())
    for (const auto &s : _out_buffers) _output_mode[s.name] = XDMFWriter::Mode::CELL;
  else if (modes.size() != names.size())
    paramError("output_mode", "Specify one output mode per buffer.", modes.size(), " != ", names.size());
  else
    for (std::size_t i = 0; i < names.size(); ++i) {
      std::string m = modes[i];
      std::transform(m.begin(), m.end(), m.begin(), ::toupper);
      if (m == "CELL") _output_mode[names[i]] = XDMFWriter::Mode::CELL;
      else if (m == "NODE") _output_mode[names[i]] = XDMFWriter::Mode::NODE;
      else if (m == "OVERSIZED_NODAL") _output_mode[names[i]] = XDMFWriter::Mode::OVERSIZED_NODAL;
      else paramError("output_mode", "Invalid option \"", modes[i], "\" (CELL NODE OVERSIZED_NODAL)");
    }
  if (_domain.nRanks() > 1)
    for (const auto &m : _output_mode)
      if (m.second != XDMFWriter::Mode::CELL) mooseError("XDMFTensorOutput currently supports only CELL output mode in parallel.");
}

void XDMFTensorOutput::init() {
  std::array<double, 3> dx = {1, 1, 1}, mn = {0, 0, 0};
  for (unsigned int d = 0; d < _domain.getDim(); ++d) {
    dx[d] = _domain.getGridSpacing()[d];
    mn[d] = _domain.getDomainMin()[d];
  }
  std::vector<XDMFWriter::Bounds> bounds;
  if (_domain.nRanks() > 1) {
    bounds.resize(_domain.nRanks());
    for (unsigned int r = 0; r < _domain.nRanks(); ++r) _domain.getLocalBounds(r, bounds[r].first, bounds[r].second);
  }
  _writer = std::make_unique<XDMFWriter>(_domain.getDim(), _domain.getGridSize(), dx, mn, _transpose, _file_base, _domain.rank(), bounds, _enable_hdf5);
}

// the buffers' metadata is read here, on the main thread; the output thread only touches the CPU copies
void XDMFTensorOutput::prepareForOutput() {
  if (!_writer) init();
  _fields.clear();
  int64_t cells = _domain.getNumberOfLocalCells(), nodes = 1;
  for (unsigned int d = 0; d < _domain.getDim(); ++d) nodes *= _domain.getShape()[d] + 1;
  for (const auto &s : _out_buffers) {
    const marlin::Tensor &t = s.buffer->getRawTensor();
    if (!t.defined()) continue;  // :270-271
    if (t.is_complex() || t.space() == marlin::Space::RECIPROCAL) mooseError("XDMFTensorOutput: buffer '", s.name, "' is a reciprocal-space buffer");
    const XDMFWriter::Mode mode = _output_mode.at(s.name);
    const int64_t expect = (mode == XDMFWriter::Mode::OVERSIZED_NODAL ? nodes : cells) * t.ncomp();
    if ((int64_t)s.cpu->size() != expect)
      mooseError("XDMFTensorOutput: buffer '", s.name, "' has ", s.cpu->size(), " values, output mode expects ", expect);
    _fields.push_back(XDMFWriter::Field{s.name, mode, t.ncomp(), s.cpu->data()});
  }
}

void XDMFTensorOutput::output() { _writer->addFrame(_time, _fields); }
