// Fused forward-FFT + semi-implicit k-space update + inverse-FFT pass, and small pointwise kernels.
#include "k_common.cuh"

namespace mrl {

template <class T, class C>
static cudaError_t fused_fast(const LaunchCtx &lc, const FusedIO<T> &io0, const SpectralUpdate<T> &up, const cx<T> *tw) {
  constexpr int TK = TileK<T, C>::value;
  FusedIO<T> io = io0;
  io.ncb = (io.ncols + TK - 1) / TK;
  const int block = TK * C::TP;
  const size_t smem = (size_t)(C::N * TK + C::N) * sizeof(cx<T>);
  auto k = k_fused_fast<T, C, TK>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, block, smem, &per_sm);
  if (e != cudaSuccess) return e;
  k<<<grid_for((long long)io.nouter * io.ncb, lc, per_sm), block, smem, lc.stream>>>(io, up, tw);
  return cudaGetLastError();
}
template <class T, int TK>
static cudaError_t fused_gen(const LaunchCtx &lc, const FusedIO<T> &io0, const SpectralUpdate<T> &up, const cx<T> *tw,
                             const FFTPlanDev &plan) {
  FusedIO<T> io = io0;
  io.ncb = (io.ncols + TK - 1) / TK;
  const size_t smem = (size_t)3 * plan.n * TK * sizeof(cx<T>);
  auto k = k_fused_gen<T, TK>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, 256, smem, &per_sm);
  if (e != cudaSuccess) return e;
  k<<<grid_for((long long)io.nouter * io.ncb, lc, per_sm), 256, smem, lc.stream>>>(io, up, tw, plan);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_fused(const LaunchCtx &lc, const FusedIO<T> &io, const SpectralUpdate<T> &up, const cx<T> *tw,
                         const FFTPlanDev &plan) {
  switch (plan.n) {
#define X(N, TP, R0, R1, R2, R3) \
  case N: return fused_fast<T, FFTCfg<N, TP, R0, R1, R2, R3>>(lc, io, up, tw);
    MRL_FAST_SIZES(X)
#undef X
    default: break;
  }
  switch (gen_tk_for<T>(plan.n, 3, (long long)(io.nouter > 0 ? io.nouter : 1) * io.ncols, lc.sm_count)) {
    case 8: return fused_gen<T, 8>(lc, io, up, tw, plan);
    case 4: return fused_gen<T, 4>(lc, io, up, tw, plan);
    case 2: return fused_gen<T, 2>(lc, io, up, tw, plan);
    case 1: return fused_gen<T, 1>(lc, io, up, tw, plan);
    default: return cudaErrorInvalidValue;
  }
}

static inline int ew_grid(long long total, const LaunchCtx &lc) {
  long long g = (total + 255) / 256;
  long long cap = (long long)lc.sm_count * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

template <class T>
cudaError_t launch_kfactor(const LaunchCtx &lc, T *out, const T *kx, const T *ky, const T *kz, int n0, int n1, int n2,
                           int kind, T factor) {
  const long long total = (long long)n0 * n1 * n2;
  k_kfactor<T><<<ew_grid(total, lc), 256, 0, lc.stream>>>(out, kx, ky, kz, n0, n1, n2, kind, factor);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_ab_update(const LaunchCtx &lc, cx<T> *ubar, const cx<T> *cbar, const cx<T> *N, const T *L, T dt, T b0,
                             int nold, const cx<T> *const *old, const T *bold, long long total) {
  const cx<T> *o[4] = {nullptr, nullptr, nullptr, nullptr};
  T c[4] = {0, 0, 0, 0};
  for (int i = 0; i < nold && i < 4; ++i) {
    o[i] = old[i];
    c[i] = bold[i];
  }
  k_ab_update<T><<<ew_grid(total, lc), 256, 0, lc.stream>>>(ubar, cbar, N, L, dt, b0, nold, o[0], o[1], o[2], o[3], c[0],
                                                             c[1], c[2], c[3], total);
  return cudaGetLastError();
}

template <class T> cudaError_t launch_mul_rc(const LaunchCtx &lc, cx<T> *out, const T *a, const cx<T> *b, long long total) {
  k_mul_rc<T><<<ew_grid(total, lc), 256, 0, lc.stream>>>(out, a, b, total);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_nonlin(const LaunchCtx &lc, T *out, const T *in, const NonlinDesc &nl, long long total) {
  if (nl.kind != 0) return cudaErrorInvalidValue;
  DoubleWellDeriv<T> f{(T)nl.p[0], (T)nl.p[1], (T)nl.p[2]};
  k_pointwise1<T, DoubleWellDeriv<T>><<<ew_grid(total, lc), 256, 0, lc.stream>>>(out, in, f, total);
  return cudaGetLastError();
}

#define INST(T)                                                                                                        \
  template cudaError_t launch_fused<T>(const LaunchCtx &, const FusedIO<T> &, const SpectralUpdate<T> &, const cx<T> *, \
                                       const FFTPlanDev &);                                                            \
  template cudaError_t launch_kfactor<T>(const LaunchCtx &, T *, const T *, const T *, const T *, int, int, int, int, T); \
  template cudaError_t launch_ab_update<T>(const LaunchCtx &, cx<T> *, const cx<T> *, const cx<T> *, const T *, T, T,  \
                                           int, const cx<T> *const *, const T *, long long);                           \
  template cudaError_t launch_mul_rc<T>(const LaunchCtx &, cx<T> *, const T *, const cx<T> *, long long);              \
  template cudaError_t launch_nonlin<T>(const LaunchCtx &, T *, const T *, const NonlinDesc &, long long);
INST(double)
INST(float)

}  // namespace mrl
