// Last-axis real FFT pass launchers: r2c (row pairs, or variable + nonlinearity) and c2r.
#include "k_common.cuh"

namespace mrl {

template <class T, class C> struct ZCfg {
  // pencils per CTA: aim for 256 threads
  static constexpr int PPB = (256 / C::TP) < 1 ? 1 : (256 / C::TP);
  static constexpr int NP = C::N + (C::N >> 3) + 1;
  static constexpr size_t smem = (size_t)(NP * PPB + C::N) * sizeof(cx<T>);
};

template <class T, class C, class LD, class ST>
static cudaError_t zfwd_fast(const LaunchCtx &lc, const LD &ld, const ST &st, const cx<T> *tw, long long npencils) {
  typedef ZCfg<T, C> Z;
  auto k = k_zfwd_fast<T, C, Z::PPB, LD, ST>;
  int per_sm = 0;
  const int block = Z::PPB * C::TP;
  cudaError_t e = kernel_prep((const void *)k, block, Z::smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nblk = (npencils + Z::PPB - 1) / Z::PPB;
  k<<<grid_for(nblk, lc, per_sm), block, Z::smem, lc.stream>>>(ld, st, tw, npencils);
  return cudaGetLastError();
}
template <class T, int TK, class LD, class ST>
static cudaError_t zfwd_gen(const LaunchCtx &lc, const LD &ld, const ST &st, const cx<T> *tw, const FFTPlanDev &plan,
                            long long npencils) {
  auto k = k_zfwd_gen<T, TK, LD, ST>;
  const size_t smem = (size_t)2 * plan.n * TK * sizeof(cx<T>);
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, 256, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nblk = (npencils + TK - 1) / TK;
  k<<<grid_for(nblk, lc, per_sm), 256, smem, lc.stream>>>(ld, st, tw, plan, npencils);
  return cudaGetLastError();
}
template <class T, class LD, class ST>
static cudaError_t zfwd_any(const LaunchCtx &lc, const LD &ld, const ST &st, const cx<T> *tw, const FFTPlanDev &plan,
                            long long npencils) {
  switch (plan.n) {
#define X(N, TP, R0, R1, R2, R3) \
  case N: return zfwd_fast<T, FFTCfg<N, TP, R0, R1, R2, R3>, LD, ST>(lc, ld, st, tw, npencils);
    MRL_FAST_SIZES(X)
#undef X
    default: break;
  }
  switch (gen_tk_for<T>(plan.n, 2, npencils, lc.sm_count)) {
    case 8: return zfwd_gen<T, 8, LD, ST>(lc, ld, st, tw, plan, npencils);
    case 4: return zfwd_gen<T, 4, LD, ST>(lc, ld, st, tw, plan, npencils);
    case 2: return zfwd_gen<T, 2, LD, ST>(lc, ld, st, tw, plan, npencils);
    case 1: return zfwd_gen<T, 1, LD, ST>(lc, ld, st, tw, plan, npencils);
    default: return cudaErrorInvalidValue;
  }
}

template <class T, class C, class LD, class ST>
static cudaError_t zinv_fast(const LaunchCtx &lc, const LD &ld, const ST &st, const cx<T> *tw, long long npencils) {
  typedef ZCfg<T, C> Z;
  auto k = k_zinv_fast<T, C, Z::PPB, LD, ST>;
  int per_sm = 0;
  const int block = Z::PPB * C::TP;
  cudaError_t e = kernel_prep((const void *)k, block, Z::smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nblk = (npencils + Z::PPB - 1) / Z::PPB;
  k<<<grid_for(nblk, lc, per_sm), block, Z::smem, lc.stream>>>(ld, st, tw, npencils);
  return cudaGetLastError();
}
template <class T, int TK, class LD, class ST>
static cudaError_t zinv_gen(const LaunchCtx &lc, const LD &ld, const ST &st, const cx<T> *tw, const FFTPlanDev &plan,
                            long long npencils) {
  auto k = k_zinv_gen<T, TK, LD, ST>;
  const size_t smem = (size_t)2 * plan.n * TK * sizeof(cx<T>);
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, 256, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nblk = (npencils + TK - 1) / TK;
  k<<<grid_for(nblk, lc, per_sm), 256, smem, lc.stream>>>(ld, st, tw, plan, npencils);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_zfwd_pairs(const LaunchCtx &lc, const T *in, cx<T> *out, long long nrows, int n, const cx<T> *tw,
                              const FFTPlanDev &plan) {
  ZLoadPairs<T> ld{in, nrows, n};
  ZStorePairs<T> st{out, nrows, n / 2 + 1};
  return zfwd_any<T>(lc, ld, st, tw, plan, (nrows + 1) / 2);
}

template <class T>
cudaError_t launch_zinv_pairs(const LaunchCtx &lc, const cx<T> *in, T *out, long long nrows, int n, T scale,
                              const cx<T> *tw, const FFTPlanDev &plan) {
  ZInvLoadPairs<T> ld{in, nrows, n, n / 2 + 1};
  ZInvStorePairs<T> st{out, nrows, n, scale};
  const long long npencils = (nrows + 1) / 2;
  switch (plan.n) {
#define X(N, TP, R0, R1, R2, R3) \
  case N: return zinv_fast<T, FFTCfg<N, TP, R0, R1, R2, R3>>(lc, ld, st, tw, npencils);
    MRL_FAST_SIZES(X)
#undef X
    default: break;
  }
  switch (gen_tk_for<T>(plan.n, 2, npencils, lc.sm_count)) {
    case 8: return zinv_gen<T, 8>(lc, ld, st, tw, plan, npencils);
    case 4: return zinv_gen<T, 4>(lc, ld, st, tw, plan, npencils);
    case 2: return zinv_gen<T, 2>(lc, ld, st, tw, plan, npencils);
    case 1: return zinv_gen<T, 1>(lc, ld, st, tw, plan, npencils);
    default: return cudaErrorInvalidValue;
  }
}

template <class T>
cudaError_t launch_zfwd_nonlin(const LaunchCtx &lc, const T *c, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows,
                               int n, const NonlinDesc &nl, const cx<T> *tw, const FFTPlanDev &plan) {
  if (nl.kind != 0) return cudaErrorInvalidValue;
  typedef DoubleWellDeriv<T> F;
  ZLoadFused<T, F> ld{c, mu_out, n, F{(T)nl.p[0], (T)nl.p[1], (T)nl.p[2]}};
  ZStoreTwo<T> st{outC, outG, n / 2 + 1};
  return zfwd_any<T>(lc, ld, st, tw, plan, nrows);
}

#define INST(T)                                                                                                      \
  template cudaError_t launch_zfwd_pairs<T>(const LaunchCtx &, const T *, cx<T> *, long long, int, const cx<T> *,   \
                                            const FFTPlanDev &);                                                     \
  template cudaError_t launch_zinv_pairs<T>(const LaunchCtx &, const cx<T> *, T *, long long, int, T, const cx<T> *, \
                                            const FFTPlanDev &);                                                     \
  template cudaError_t launch_zfwd_nonlin<T>(const LaunchCtx &, const T *, T *, cx<T> *, cx<T> *, long long, int,    \
                                             const NonlinDesc &, const cx<T> *, const FFTPlanDev &);
INST(double)
INST(float)

}  // namespace mrl
