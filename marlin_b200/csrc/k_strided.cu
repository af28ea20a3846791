// Strided complex FFT pass launchers (forward / inverse along a non-contiguous axis).
#include "k_common.cuh"

namespace mrl {

template <class T, class C> static cudaError_t strided_fast(const LaunchCtx &lc, const StridedIO<T> &io0, const cx<T> *tw) {
  constexpr int TK = TileK<T, C>::value;
  StridedIO<T> io = io0;
  io.ncb = (io.ncols + TK - 1) / TK;
  const int block = TK * C::TP;
  const size_t smem = (size_t)(C::N * TK + C::N) * sizeof(cx<T>);
  auto k = k_strided_fast<T, C, TK>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, block, smem, &per_sm);
  if (e != cudaSuccess) return e;
  k<<<grid_for(io.ntiles(), lc, per_sm), block, smem, lc.stream>>>(io, tw);
  return cudaGetLastError();
}

template <class T, int TK>
static cudaError_t strided_gen(const LaunchCtx &lc, const StridedIO<T> &io0, const cx<T> *tw, const FFTPlanDev &plan) {
  StridedIO<T> io = io0;
  io.ncb = (io.ncols + TK - 1) / TK;
  const size_t smem = (size_t)2 * plan.n * TK * sizeof(cx<T>);
  auto k = k_strided_gen<T, TK>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, 256, smem, &per_sm);
  if (e != cudaSuccess) return e;
  k<<<grid_for(io.ntiles(), lc, per_sm), 256, smem, lc.stream>>>(io, tw, plan);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_strided(const LaunchCtx &lc, const StridedIO<T> &io, const cx<T> *tw, const FFTPlanDev &plan) {
  switch (plan.n) {
#define X(N, TP, R0, R1, R2, R3) \
  case N: return strided_fast<T, FFTCfg<N, TP, R0, R1, R2, R3>>(lc, io, tw);
    MRL_FAST_SIZES(X)
#undef X
    default: break;
  }
  switch (gen_tk_for<T>(plan.n, 2, (long long)io.nfields * io.nouter * io.ncols, lc.sm_count)) {
    case 8: return strided_gen<T, 8>(lc, io, tw, plan);
    case 4: return strided_gen<T, 4>(lc, io, tw, plan);
    case 2: return strided_gen<T, 2>(lc, io, tw, plan);
    case 1: return strided_gen<T, 1>(lc, io, tw, plan);
    default: return cudaErrorInvalidValue;
  }
}

template cudaError_t launch_strided<double>(const LaunchCtx &, const StridedIO<double> &, const cx<double> *, const FFTPlanDev &);
template cudaError_t launch_strided<float>(const LaunchCtx &, const StridedIO<float> &, const cx<float> *, const FFTPlanDev &);

}  // namespace mrl
