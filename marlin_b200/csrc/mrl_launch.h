// Host-side launch interface between the C-ABI implementation (mrl_api.cpp) and the kernel
// translation units (k_*.cu).  Internal; the public boundary is include/marlin_b200.h.
#pragma once
#include <cuda_runtime.h>

#include "mrl_passes.cuh"

namespace mrl {

struct LaunchCtx {
  cudaStream_t stream;
  int sm_count;
};

// X(N, TP, R0, R1, R2, R3): sizes with a register-resident FFT configuration
#define MRL_FAST_SIZES(X)  \
  X(16, 4, 4, 4, 1, 1)     \
  X(32, 4, 8, 4, 1, 1)     \
  X(64, 8, 8, 8, 1, 1)     \
  X(128, 16, 8, 4, 4, 1)   \
  X(256, 32, 8, 8, 4, 1)   \
  X(512, 64, 8, 8, 8, 1)   \
  X(1024, 128, 8, 8, 4, 4)

inline bool has_fast_cfg(int n) {
  switch (n) {
#define X(N, TP, R0, R1, R2, R3) case N:
    MRL_FAST_SIZES(X)
#undef X
    return true;
    default: return false;
  }
}

// Largest number of interleaved pencils (TK) the generic kernel can hold in shared memory.
template <class T> inline int gen_tk(int n, int nbuf) {
  const size_t budget = 200 * 1024;
  for (int tk = 8; tk >= 1; tk >>= 1)
    if ((size_t)nbuf * n * tk * sizeof(cx<T>) <= budget) return tk;
  return 0;
}

// Same, but no more pencils per CTA than leaves every SM a tile: small 2-D grids (PFHub 200^2) would
// otherwise run on a dozen CTAs.  work = pencils (columns x slices) of the pass.
template <class T> inline int gen_tk_for(int n, int nbuf, long long work, int sm_count) {
  int tk = gen_tk<T>(n, nbuf);
  while (tk > 1 && (work + tk - 1) / tk < sm_count) tk >>= 1;
  return tk;
}

// Real-space nonlinearity selector for the fused first pass.
struct NonlinDesc {
  int kind;          // 0: double-well derivative 2A(c-a)(b-c)(a+b-2c)
  double p[4];       // A, a, b
};

template <class T> cudaError_t launch_strided(const LaunchCtx &lc, const StridedIO<T> &io, const cx<T> *tw, const FFTPlanDev &plan);
template <class T>
cudaError_t launch_zfwd_pairs(const LaunchCtx &lc, const T *in, cx<T> *out, long long nrows, int n, const cx<T> *tw,
                              const FFTPlanDev &plan);
template <class T>
cudaError_t launch_zinv_pairs(const LaunchCtx &lc, const cx<T> *in, T *out, long long nrows, int n, T scale,
                              const cx<T> *tw, const FFTPlanDev &plan);
template <class T>
cudaError_t launch_zfwd_nonlin(const LaunchCtx &lc, const T *c, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows,
                               int n, const NonlinDesc &nl, const cx<T> *tw, const FFTPlanDev &plan);
template <class T>
cudaError_t launch_fused(const LaunchCtx &lc, const FusedIO<T> &io, const SpectralUpdate<T> &up, const cx<T> *tw,
                         const FFTPlanDev &plan);
// TMA-pipelined versions (k_tma.cu); return cudaErrorNotSupported when the size / layout has no
// pipelined configuration, in which case the caller uses the kernels above.
template <class T> cudaError_t launch_strided_tma(const LaunchCtx &lc, const StridedIO<T> &io, const cx<T> *tw, int n);
template <class T>
cudaError_t launch_fused_tma(const LaunchCtx &lc, const FusedIO<T> &io, const SpectralUpdate<T> &up, const cx<T> *tw, int n);
template <class T>
cudaError_t launch_zfwd_nonlin_tma(const LaunchCtx &lc, const T *c, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows, int n,
                                   int ncp, const NonlinDesc &nl, const cx<T> *tw, const RowMap &rm = RowMap{0, 0, 0});
// optional inner product fused into the store of the last-axis c2r pass: sum(result * with) as one partial per CTA
// (partials[0 .. *count), capacity entries available); with == nullptr: none
template <class T> struct ZinvDot {
  const T *with;
  double *partials;
  int capacity;
  int *count;
};
template <class T>
cudaError_t launch_zinv_pairs_tma(const LaunchCtx &lc, const cx<T> *in, int ncp, T *out, long long nrows, int n, T scale,
                                  const cx<T> *tw, const ZinvDot<T> *dot = nullptr);
template <class T>
cudaError_t launch_zfwd_pairs_tma(const LaunchCtx &lc, const T *in, cx<T> *out, long long nrows, int n, int ncp, const cx<T> *tw);
// first pass of G(K4 : p) with the tangent (and the CG direction update) fused into its load; 3-D, n = 256 or 512
template <class T> cudaError_t launch_mech_tangent_zfwd(const LaunchCtx &lc, const MechTangentIO<T> &io, const cx<T> *tw, int n);
template <class T>
cudaError_t launch_mech_fused_tma(const LaunchCtx &lc, cx<T> *spec, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzv,
                                  int ncp, const cx<T> *tw);
// slab x passes with bulk peer stores into the blocked staging layouts (mrl_passes_slab.cuh)
template <class T> struct SlabXIO;
template <class T> cudaError_t launch_slab_xfwd(const LaunchCtx &lc, const cx<T> *in, const SlabXIO<T> &io, const cx<T> *tw, int n);
template <class T> cudaError_t launch_slab_xinv(const LaunchCtx &lc, const cx<T> *S, const SlabXIO<T> &io, const cx<T> *tw, int n);
template <class T> int fused_tma_tk(int n);  // column-block width of the fused pass (0: no TMA configuration)
bool tma_enabled();
template <class T>
cudaError_t launch_kfactor(const LaunchCtx &lc, T *out, const T *kx, const T *ky, const T *kz, int n0, int n1, int n2,
                           int kind, T factor);
template <class T>
cudaError_t launch_ab_update(const LaunchCtx &lc, cx<T> *ubar, const cx<T> *cbar, const cx<T> *N, const T *L, T dt, T b0,
                             int nold, const cx<T> *const *old, const T *bold, long long total);
template <class T> cudaError_t launch_mul_rc(const LaunchCtx &lc, cx<T> *out, const T *a, const cx<T> *b, long long total);
template <class T>
cudaError_t launch_nonlin(const LaunchCtx &lc, T *out, const T *in, const NonlinDesc &nl, long long total);

}  // namespace mrl
