// [TensorOutputs]: objects that write tensor buffers to disk at INITIAL / TIMESTEP_END.
//   TensorOutput       src/tensor_outputs/TensorOutput.C:16-81 (buffer, file_base, execute_on; output() runs in a
//                      std::thread on the CPU copies TensorProblem made at the step's synchronisation point, :66-81,
//                      with the output time snapshotted so the next step cannot race it)
//   XDMFTensorOutput   src/tensor_outputs/XDMFTensorOutput.C:29-761 (XMF skeleton :118-221, writeLocalData :266-355,
//                      writeSerialXMF :358-426, writeParallelXMF :429-527, extendTensor :529-553,
//                      buildAttributeNames :654-668, rankTag / hdf5FileName / binaryFileName :737-761).  Data either
//                      as raw little-endian binary files (<file_base>[.rankNNNN].<name>.<frame>.bin) or, with
//                      enable_hdf5 = true, as deflate-compressed one-chunk datasets of <file_base>[.rankNNNN].h5
//                      (addDataToHDF5 :572-651; written by host/shim/h5lite - no libhdf5 in this build); one set
//                      of data files per rank and a spatial collection written by rank 0 in parallel runs.
#pragma once
#include <sstream>
#include <thread>
#include <utility>

#include "TensorProblem.h"
#include "h5lite.h"

class TensorOutput : public MooseObject {
public:
  static InputParameters validParams();
  explicit TensorOutput(const InputParameters &parameters);
  ~TensorOutput() override;  // joins the output thread; an error it left behind is reported, not thrown
  virtual void init() {}
  bool shouldRun(ExecFlagType flag) const { return (_execute_on & flag) != 0; }
  // TensorOutput.C:66-81: output() in a dedicated thread; an exception it throws is re-thrown by waitForCompletion
  void startOutput();
  void waitForCompletion();

protected:
  virtual void prepareForOutput() {}  // snapshot of light-weight metadata before the thread starts
  virtual void output() = 0;
  TensorProblem &_tensor_problem;
  const DomainAction &_domain;
  const Real &_time;  // TensorProblem::outputTime(): not advanced while an output is running
  const std::string _file_base;
  const int _execute_on;
  // name -> (buffer, CPU copy made by TensorProblem::execute before the outputs run)
  struct Source {
    std::string name;
    TensorBufferBase *buffer;
    const std::vector<double> *cpu;
  };
  std::vector<Source> _out_buffers;  // in std::map (name) order like the reference

private:
  std::thread _output_thread;
  std::exception_ptr _thread_error;
};

// The XMF document and the data files, independent of the device (unit-tested on the host).
class XDMFWriter {
public:
  enum class Mode { CELL, NODE, OVERSIZED_NODAL };
  struct Field {
    std::string name;
    Mode mode;
    int ncomp;                   // trailing value dimensions flattened; data is component major
    const double *data;          // [ncomp][cells] (CELL, NODE) or [ncomp][nodes] (OVERSIZED_NODAL)
  };
  using Bounds = std::pair<std::array<int64_t, 3>, std::array<int64_t, 3>>;  // [begin, end) of a rank's real-space part
  // n: global grid.  bounds: one entry per rank in parallel runs (empty: serial); `rank` writes its own part's data
  // files, rank 0 the XMF document.
  XDMFWriter(unsigned int dim, const std::array<int64_t, 3> &n, const std::array<double, 3> &dx, const std::array<double, 3> &min, bool transpose,
             std::string file_base, unsigned int rank = 0, std::vector<Bounds> bounds = {}, bool enable_hdf5 = false);
  void addFrame(double time, const std::vector<Field> &fields);  // writes the .bin files and <file_base>.xmf
  std::string xml() const;
  static std::vector<std::string> attributeNames(const std::string &buffer_name, int64_t num_fields);

private:
  bool parallel() const { return !_bounds.empty(); }
  std::string rankTag(unsigned int rank) const;
  std::string binaryFileName(const std::string &setname, unsigned int rank) const { return _file_base + rankTag(rank) + "." + setname + ".bin"; }
  std::string hdf5FileName(unsigned int rank) const { return _file_base + rankTag(rank) + ".h5"; }
  std::string dataItem(const std::string &dims, const std::string &dataset, unsigned int rank) const;  // the <DataItem> of one dataset
  std::vector<double> arrange(const Field &f, int component, std::vector<uint64_t> *dims = nullptr) const;  // extend (NODE) + transpose, one component
  std::string serialFrame(double time, const std::vector<Field> &fields) const;
  std::string parallelFrame(double time, const std::vector<Field> &fields) const;
  unsigned int _dim;
  std::array<int64_t, 3> _n;         // the part this process writes (the whole grid in serial runs)
  std::array<double, 3> _dx, _min;
  unsigned int _rank;
  std::vector<Bounds> _bounds;
  bool _transpose;
  std::string _file_base, _head, _frames;
  std::string _cell_dims, _node_dims;
  unsigned int _frame = 0;
  std::unique_ptr<H5LiteFile> _h5;  // enable_hdf5: this process's data file
};

class XDMFTensorOutput : public TensorOutput {
public:
  static InputParameters validParams();
  explicit XDMFTensorOutput(const InputParameters &parameters);
  void init() override;

protected:
  void prepareForOutput() override;
  void output() override;
  std::vector<XDMFWriter::Field> _fields;
  std::map<std::string, XDMFWriter::Mode> _output_mode;
  const bool _transpose;
  const bool _enable_hdf5;
  std::unique_ptr<XDMFWriter> _writer;
};
