// Shared launch helpers for the kernel translation units.
#pragma once
#include <mutex>
#include <unordered_map>

#include "mrl_launch.h"

namespace mrl {

// Opt in to the dynamic shared memory a kernel needs and find how many CTAs fit on one SM;
// grids are sized to sm_count * resident CTAs (persistent grid-stride loops over tiles).
inline cudaError_t kernel_prep(const void *fn, int block, size_t smem, int *ctas_per_sm) {
  struct Info {
    size_t max_smem = 0;
    std::unordered_map<size_t, int> occ;
  };
  static std::mutex mu;
  static std::unordered_map<const void *, Info> cache;
  std::lock_guard<std::mutex> g(mu);
  Info &m = cache[fn];
  if (smem > m.max_smem) {  // the attribute is sticky per function: only ever raise it
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    m.max_smem = smem;
  }
  auto it = m.occ.find(smem * 4096 + block);
  if (it != m.occ.end()) {
    *ctas_per_sm = it->second;
    return cudaSuccess;
  }
  int nb = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, block, smem);
  if (e != cudaSuccess) return e;
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  m.occ[smem * 4096 + block] = nb;
  *ctas_per_sm = nb;
  return cudaSuccess;
}

inline int grid_for(long long nwork, const LaunchCtx &lc, int ctas_per_sm) {
  long long cap = (long long)lc.sm_count * ctas_per_sm;
  long long g = nwork < cap ? nwork : cap;
  return (int)(g < 1 ? 1 : g);
}

// 128 bytes of interleaved columns per shared-memory row
// (capped so that a CTA never exceeds 1024 threads)
template <class T, class C> struct TileK {
  static constexpr int full = 128 / (int)sizeof(cx<T>);
  static constexpr int value = (full * C::TP > 1024) ? (1024 / C::TP) : full;
};

}  // namespace mrl
