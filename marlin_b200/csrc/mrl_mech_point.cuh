// marlin_b200 - per-voxel constitutive law / tangent action of the de Geus mechanics in closed form (see k_mech.cu
// for the algebra and the reference citations); shared by the pointwise kernels (k_mech.cu) and the first FFT pass
// with the tangent fused into its load (mrl_passes_tma.cuh).
#pragma once

namespace mrl {

template <class T, int D> struct MD {
  T a[D][D];
};

template <class T, int D> __device__ __forceinline__ void second_pk(const MD<T, D> &F, T K, T mu, MD<T, D> &S) {
  T E[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < D; ++k) s += F.a[k][i] * F.a[k][j];
      E[i][j] = T(0.5) * (s - (i == j ? T(1) : T(0)));
    }
  T tr = T(0);
#pragma unroll
  for (int i = 0; i < D; ++i) tr += E[i][i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) S.a[i][j] = T(2) * mu * E[i][j] + (i == j ? (K - T(2) * mu / T(3)) * tr : T(0));
}

// mode 0: R = P = F S.   mode 1-3: R = K4 : X.
template <class T, int D> __device__ __forceinline__ void mech_point(int mode, const MD<T, D> &Fm, T K, T mu, const MD<T, D> &X, MD<T, D> &R) {
  MD<T, D> S;
  second_pk(Fm, K, mu, S);
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) s += Fm.a[i][k] * S.a[k][j];
        R.a[i][j] = s;
      }
  } else {
    T W[D][D];
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
      for (int l = 0; l < D; ++l) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) s += Fm.a[k][p] * X.a[k][l];
        W[p][l] = s;
      }
    T tr = T(0);
#pragma unroll
    for (int i = 0; i < D; ++i) tr += W[i][i];
    T Tm[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) Tm[i][j] = mu * (W[i][j] + W[j][i]) + (i == j ? (K - T(2) * mu / T(3)) * tr : T(0));
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) s += X.a[i][k] * S.a[k][j];
#pragma unroll
        for (int k = 0; k < D; ++k) s += Fm.a[i][k] * Tm[k][j];
        R.a[i][j] = s;
      }
  }
}

}  // namespace mrl
