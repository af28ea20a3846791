// [TensorOutputs]: objects that write tensor buffers to disk at INITIAL / TIMESTEP_END.
//   TensorOutput       src/tensor_outputs/TensorOutput.C:16-81 (buffer, file_base, execute_on; the reference runs
//                      output() in a std::thread on the CPU copies made by TensorProblem - here the copies are
//                      made at the same synchronisation point and written before the next step starts)
//   XDMFTensorOutput   src/tensor_outputs/XDMFTensorOutput.C:29-761 (XMF skeleton :118-221, writeLocalData :266-355,
//                      writeSerialXMF :358-426, extendTensor :529-553, buildAttributeNames :654-668,
//                      binaryFileName :758-761).  Serial layout, raw little-endian binary data files
//                      (<file_base>.<name>.<frame>.bin, the reference's non-HDF5 format); HDF5 cannot be written
//                      in this build (no libhdf5): enable_hdf5 = true is accepted and falls back to binary.
#pragma once
#include <sstream>

#include "TensorProblem.h"

class TensorOutput : public MooseObject {
public:
  static InputParameters validParams();
  explicit TensorOutput(const InputParameters &parameters);
  virtual void init() {}
  virtual void output() = 0;
  bool shouldRun(ExecFlagType flag) const { return (_execute_on & flag) != 0; }

protected:
  TensorProblem &_tensor_problem;
  const DomainAction &_domain;
  const std::string _file_base;
  const int _execute_on;
  // name -> (buffer, CPU copy made by TensorProblem::execute before the outputs run)
  struct Source {
    std::string name;
    TensorBufferBase *buffer;
    const std::vector<double> *cpu;
  };
  std::vector<Source> _out_buffers;  // in std::map (name) order like the reference
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
  XDMFWriter(unsigned int dim, const std::array<int64_t, 3> &n, const std::array<double, 3> &dx, const std::array<double, 3> &min, bool transpose,
             std::string file_base);
  void addFrame(double time, const std::vector<Field> &fields);  // writes the .bin files and <file_base>.xmf
  std::string xml() const;
  static std::vector<std::string> attributeNames(const std::string &buffer_name, int64_t num_fields);

private:
  std::string binaryFileName(const std::string &setname) const { return _file_base + "." + setname + ".bin"; }
  std::vector<double> arrange(const Field &f, int component) const;  // extend (NODE) + transpose, one component
  unsigned int _dim;
  std::array<int64_t, 3> _n;
  bool _transpose;
  std::string _file_base, _head, _frames;
  std::string _cell_dims, _node_dims;
  unsigned int _frame = 0;
};

class XDMFTensorOutput : public TensorOutput {
public:
  static InputParameters validParams();
  explicit XDMFTensorOutput(const InputParameters &parameters);
  void init() override;
  void output() override;

protected:
  std::map<std::string, XDMFWriter::Mode> _output_mode;
  const bool _transpose;
  std::unique_ptr<XDMFWriter> _writer;
};
