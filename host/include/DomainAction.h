// [Domain] block: grid contract, device context, FFT entry points.
// Host mirror of DomainAction (reference include/actions/DomainAction.h:31-120,
// src/actions/DomainAction.C:24-338 parameters + gridChanged, :854-867 / :1054-1066 fft / ifft,
// :1559-1574 sum / average).  All numerics go through the C ABI (include/marlin_b200.h).
// parallel_mode = FFT_SLAB (:511-566 partitionSlabs, :870-1019 fftSlab / ifftSlab) and FFT_PENCIL (:569-742
// partitionPencils, :1022-1047, :1106-1404 fftPencil / ifftPencil and their exchange stages): one process per GPU, the
// processes find each other through Comm (host/shim/comm.h, torchrun-style environment); the domain of the
// context is then the rank's part (mrl_domain_set_dist) and fft / ifft are mrl_dist_rfftn / mrl_dist_irfftn.
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>

#include "MarlinTensor.h"
#include "comm.h"
#include "marlin_b200.h"
#include "moose_shim.h"

namespace marlin {

// size-bucketed free list in front of mrl_malloc / mrl_free: every operator result is a fresh
// tensor (reference semantics), so blocks are recycled instead of going back to the driver.
class TensorPool {
public:
  explicit TensorPool(mrl_context *ctx) : _ctx(ctx) {}
  ~TensorPool();
  std::shared_ptr<TensorStorage> get(size_t bytes);
  void put(void *dev, size_t bytes);
  size_t bytesAllocated() const { return _allocated; }
  size_t allocations() const { return _n_alloc; }

private:
  mrl_context *_ctx;
  std::multimap<size_t, void *> _free;
  size_t _allocated = 0, _n_alloc = 0;
};

}  // namespace marlin

class DomainAction : public MooseObject {
public:
  static InputParameters validParams();
  explicit DomainAction(const InputParameters &parameters);
  ~DomainAction() override;

  enum class ParallelMode { NONE, REAL_SPACE, FFT_SLAB, FFT_PENCIL };

  void gridChanged();

  unsigned int getDim() const { return _dim; }
  const std::array<int64_t, 3> &getGridSize() const { return _n_global; }
  const std::array<int64_t, 3> &getShape() const { return _shape; }
  const std::array<int64_t, 3> &getReciprocalShape() const { return _reciprocal_shape; }
  const std::array<Real, 3> &getDomainMin() const { return _min_global; }
  const std::array<Real, 3> &getDomainMax() const { return _max_global; }
  const std::array<Real, 3> &getGridSpacing() const { return _grid_spacing; }
  Real getVolume() const { return _volume; }
  // global cell count (DomainAction.C:1577-1580); tensors hold the local part
  int64_t getNumberOfCells() const { return _n_global[0] * _n_global[1] * _n_global[2]; }
  int64_t getNumberOfLocalCells() const { return _shape[0] * _shape[1] * _shape[2]; }
  int64_t getNumberOfReciprocalCells() const { return _reciprocal_shape[0] * _reciprocal_shape[1] * _reciprocal_shape[2]; }
  // host copies of the axes (cell centres / 2 pi fftfreq), bit-identical to the device's
  const std::vector<double> &getAxis(unsigned int d) const { return _axis[d]; }
  const std::vector<double> &getReciprocalAxis(unsigned int d) const { return _raxis[d]; }
  bool isParallelFFT() const { return _parallel_mode == ParallelMode::FFT_SLAB || _parallel_mode == ParallelMode::FFT_PENCIL; }
  // the process group (a world of one process in serial runs)
  Comm &comm() const { return _comm; }
  unsigned int rank() const { return (unsigned int)_comm.rank(); }
  unsigned int nRanks() const { return (unsigned int)_comm.size(); }
  // [begin, end) of rank's real-space part along every axis (DomainAction.C:1543-1556)
  void getLocalBounds(unsigned int rank, std::array<int64_t, 3> &begin, std::array<int64_t, 3> &end) const;
  bool debug() const { return _debug; }
  bool single() const { return _single; }
  size_t realBytes() const { return _single ? 4 : 8; }

  mrl_context *context() const { return _ctx; }
  mrl_dist *dist() const { return _dist; }  // the decomposed transforms (nullptr in serial mode)
  void check(int rc, const char *what) const;

  // ---- tensors -----------------------------------------------------------------------------
  marlin::Tensor empty(marlin::Space space, bool is_complex, int ncomp = 1) const;
  marlin::Tensor zeros(marlin::Space space, bool is_complex, int ncomp = 1) const;
  marlin::Tensor fromHost(const std::vector<double> &values, marlin::Space space, bool is_complex, int ncomp = 1) const;
  std::vector<double> toHost(const marlin::Tensor &t) const;  // synchronous; converts float -> double
  marlin::Tensor clone(const marlin::Tensor &t) const;

  // rfftn / irfftn over the spatial dims; value dims are batch (component major here)
  marlin::Tensor fft(const marlin::Tensor &t) const;
  marlin::Tensor ifft(const marlin::Tensor &t) const;

  // reductions of real tensors (sum over all entries, extreme values)
  Real sum(const marlin::Tensor &t) const;
  Real reduce(int op, const marlin::Tensor &t) const;
  Real average(const marlin::Tensor &t) const { return sum(t) / Real(t.numel()); }
  void synchronize() const;

  marlin::TensorPool &pool() const { return *_pool; }

private:
  const unsigned int _dim;
  std::array<int64_t, 3> _n_global;
  std::array<Real, 3> _min_global, _max_global, _grid_spacing;
  std::array<int64_t, 3> _shape, _reciprocal_shape;
  Real _volume = 0;
  std::array<std::vector<double>, 3> _axis, _raxis;
  const ParallelMode _parallel_mode;
  bool _single = false;
  const bool _debug;
  int _device = 0;
  Comm &_comm;
  std::vector<double> _local_weights;                      // one per rank (device_weights by local rank, DomainAction.C:176-189)
  mrl_context *_ctx = nullptr;
  mrl_dist *_dist = nullptr;
  std::unique_ptr<marlin::TensorPool> _pool;
};
