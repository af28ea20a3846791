// Device-resident tensor handle used by the host objects in place of torch::Tensor.
//
// The reference passes torch::Tensor handles around (reference counted, rebindable: `_u = expr`
// replaces the handle, history vectors keep the old ones alive - include/tensor_buffers/TensorBuffer.h:64-79).
// marlin::Tensor keeps exactly those semantics for a block of HBM owned through the C ABI
// (mrl_malloc / mrl_free): copying a Tensor copies the handle, the memory returns to the domain's
// pool when the last handle goes away.  A tensor knows which space it lives in, because the C ABI
// needs that to pick shapes (real [nx][ny][nz], reciprocal [nx][ny][nz/2+1]).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>

namespace marlin {

enum class Space { SCALAR = 0, REAL = 1, RECIPROCAL = 2, NODAL = 3 };  // NODAL: (n+1)^dim oversized nodal fields (ComputeDisplacements)

class TensorPool;

struct TensorStorage {
  TensorPool *pool = nullptr;
  void *dev = nullptr;
  size_t bytes = 0;
  ~TensorStorage();
};

class Tensor {
public:
  Tensor() = default;
  Tensor(std::shared_ptr<TensorStorage> st, Space space, bool is_complex, int ncomp, int64_t count)
    : _st(std::move(st)), _space(space), _complex(is_complex), _ncomp(ncomp), _count(count) {}
  bool defined() const { return (bool)_st; }
  void *data_ptr() const { return _st ? _st->dev : nullptr; }
  Space space() const { return _space; }
  bool is_complex() const { return _complex; }
  int ncomp() const { return _ncomp; }          // trailing value dimensions (1, 9 for rank two), component major
  int64_t count() const { return _count; }      // grid points (per component)
  int64_t numel() const { return _count * _ncomp; }
  size_t nbytes() const { return _st ? _st->bytes : 0; }
  bool is_same(const Tensor &o) const { return _st == o._st; }
  long use_count() const { return _st.use_count(); }
  // Exchange the device blocks behind two equally sized tensors without touching the handles:
  // every holder of `a` then sees b's block and vice versa.  Used for copy-on-write that must keep
  // the device pointer of one particular holder stable.
  static void swapBlocks(Tensor &a, Tensor &b) {
    void *t = a._st->dev;
    a._st->dev = b._st->dev;
    b._st->dev = t;
  }

private:
  std::shared_ptr<TensorStorage> _st;
  Space _space = Space::REAL;
  bool _complex = false;
  int _ncomp = 1;
  int64_t _count = 0;
};

}  // namespace marlin
