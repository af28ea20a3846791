// Per-wavevector dense solve of the coupled semi-implicit update,
//     (I - dt L(k)) ubar(k) = rhs(k),   L(k) real NV x NV,  rhs complex,
// behind AdamsBashforthMoultonCoupled (src/tensor_solver/AdamsBashforthMoultonCoupled.C:131-171:
// the reference stacks L into [grid..., N, N] and calls at::linalg_solve, i.e. batched LU with
// partial pivoting).  One thread per wavevector; matrix and right-hand sides live in registers;
// LU with row pivoting on |a| like getrf, so the pivot order equals LAPACK's.  HBM-bound:
// reads NV^2 real + NV complex, writes NV complex per point.
#include "k_common.cuh"
#include "mrl_internal.h"

namespace mrl {

template <class T, int NV> struct CoupledArgs {
  const T *L[NV * NV];  // L[r*NV + c] multiplies unknown c in equation r; nullptr = 0
  const cx<T> *rhs[NV];
  cx<T> *out[NV];
};

template <class T, int NV>
__global__ void __launch_bounds__(256) k_coupled_solve(CoupledArgs<T, NV> a, T dt, int drop_imag, long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    T A[NV][NV];
    cx<T> b[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) {
#pragma unroll
      for (int c = 0; c < NV; ++c) A[r][c] = (r == c ? T(1) : T(0)) - dt * (a.L[r * NV + c] ? a.L[r * NV + c][p] : T(0));
      b[r] = a.rhs[r][p];
      if (drop_imag) b[r].y = T(0);
    }
    // LU with partial pivoting, applied to the right-hand side on the fly
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int piv = k;
      T best = fabs(A[k][k]);
#pragma unroll
      for (int r = k + 1; r < NV; ++r) {
        const T v = fabs(A[r][k]);
        if (v > best) { best = v; piv = r; }
      }
#pragma unroll
      for (int r = k + 1; r < NV; ++r) {
        if (r == piv) {  // swap rows k and piv (compile-time indices keep A in registers)
#pragma unroll
          for (int c = 0; c < NV; ++c) { const T t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; }
          const cx<T> t = b[k]; b[k] = b[r]; b[r] = t;
        }
      }
      const T inv = T(1) / A[k][k];
#pragma unroll
      for (int r = k + 1; r < NV; ++r) {
        const T f = A[r][k] * inv;
#pragma unroll
        for (int c = k + 1; c < NV; ++c) A[r][c] -= f * A[k][c];
        b[r].x -= f * b[k].x;
        b[r].y -= f * b[k].y;
      }
    }
#pragma unroll
    for (int k = NV - 1; k >= 0; --k) {
      cx<T> s = b[k];
#pragma unroll
      for (int c = k + 1; c < NV; ++c) { s.x -= A[k][c] * b[c].x; s.y -= A[k][c] * b[c].y; }
      const T inv = T(1) / A[k][k];
      b[k] = mk<T>(s.x * inv, s.y * inv);
    }
#pragma unroll
    for (int r = 0; r < NV; ++r) a.out[r][p] = b[r];
  }
}

template <class T, int NV>
static cudaError_t coupled_go(const LaunchCtx &lc, const void *const *L, const void *const *rhs, void *const *out, double dt, int drop_imag,
                              long long total) {
  CoupledArgs<T, NV> a;
  for (int i = 0; i < NV * NV; ++i) a.L[i] = (const T *)L[i];
  for (int i = 0; i < NV; ++i) {
    a.rhs[i] = (const cx<T> *)rhs[i];
    a.out[i] = (cx<T> *)out[i];
  }
  long long g = (total + 255) / 256;
  const long long cap = (long long)lc.sm_count * 8;
  k_coupled_solve<T, NV><<<(int)(g < cap ? g : cap), 256, 0, lc.stream>>>(a, (T)dt, drop_imag, total);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_coupled_solve(const LaunchCtx &lc, int nvar, const void *const *L, const void *const *rhs, void *const *out, double dt,
                                 int drop_imag, long long total) {
  switch (nvar) {
    case 1: return coupled_go<T, 1>(lc, L, rhs, out, dt, drop_imag, total);
    case 2: return coupled_go<T, 2>(lc, L, rhs, out, dt, drop_imag, total);
    case 3: return coupled_go<T, 3>(lc, L, rhs, out, dt, drop_imag, total);
    case 4: return coupled_go<T, 4>(lc, L, rhs, out, dt, drop_imag, total);
    case 5: return coupled_go<T, 5>(lc, L, rhs, out, dt, drop_imag, total);
    case 6: return coupled_go<T, 6>(lc, L, rhs, out, dt, drop_imag, total);
    default: return cudaErrorNotSupported;
  }
}
template cudaError_t launch_coupled_solve<double>(const LaunchCtx &, int, const void *const *, const void *const *, void *const *, double, int,
                                                  long long);
template cudaError_t launch_coupled_solve<float>(const LaunchCtx &, int, const void *const *, const void *const *, void *const *, double, int,
                                                 long long);

}  // namespace mrl
