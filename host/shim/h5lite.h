// A writer for the subset of HDF5 that XDMFTensorOutput's data files need: one flat group of N-dimensional float / double
// datasets, each stored as ONE deflate-compressed chunk - what the reference produces through libhdf5's H5Pset_chunk(dims) +
// H5Pset_deflate(9) + H5Dwrite (src/tensor_outputs/XDMFTensorOutput.C:572-651).  libhdf5 is not available in this build
// environment, so the file structures are written directly, following the HDF5 File Format Specification (superblock
// version 0, version 1 object headers, symbol-table groups with a version 1 B-tree and a local heap, version 3 chunked layout
// with a version 1 chunk B-tree, version 1 filter pipeline) with the same versions and constants libhdf5 1.8+ chose for the
// reference's gold files (group leaf K = 4, internal K = 16, chunk B-tree K = 32).  tests/h5lite.py reads both.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

class H5LiteFile {
public:
  // creates / truncates `path` (H5Fcreate(..., H5F_ACC_TRUNC, ...))
  explicit H5LiteFile(const std::string &path);
  ~H5LiteFile();
  H5LiteFile(const H5LiteFile &) = delete;
  H5LiteFile &operator=(const H5LiteFile &) = delete;

  // one dataset at the root: element size 4 (float) or 8 (double), C order; throws if the name exists
  void addDataset(const std::string &name, const std::vector<uint64_t> &dims, int elem_size, const void *data);
  // makes the file self-consistent on disk (H5Fflush): group structures, superblock, end-of-file address
  void flush();
  size_t numDatasets() const { return _entries.size(); }

private:
  struct Entry {
    std::string name;
    uint64_t header = 0;   // address of the dataset's object header
    uint64_t heap_off = 0; // offset of the name in the local heap (set by flush)
  };
  void put(uint64_t addr, const std::vector<uint8_t> &bytes);
  std::string _path;
  std::FILE *_f = nullptr;
  uint64_t _data_end = 96;  // datasets are appended here; the group metadata is rewritten behind them on every flush
  std::vector<Entry> _entries;
};
