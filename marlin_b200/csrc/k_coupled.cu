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

// ------------------------------------------------------------------------------------------ Broyden
// BroydenSolver (src/tensor_solver/BroydenSolver.C:118-168): per wavevector an NV x NV complex inverse
// Jacobian estimate M (component-major [NV*NV][points], kept between calls).
//   step  : sk = -M R;  unew = u + 0.5 sk                                            (:118-128)
//   update: yk = Rnew - R;  d = sk^T yk (no conjugation);  M += (sk - M yk) sk^T / d  if |d| > 1e-12 (:141-157)
template <class T, int NV> struct BroydenArgs {
  cx<T> *M;
  const cx<T> *a[NV];  // step: R      update: sk
  const cx<T> *b[NV];  // step: u      update: R
  const cx<T> *c[NV];  //              update: Rnew
  cx<T> *o0[NV];       // step: sk
  cx<T> *o1[NV];       // step: unew
};
template <class T> __device__ __forceinline__ cx<T> cmul(cx<T> x, cx<T> y) { return mk<T>(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x); }

template <class T, int NV> __global__ void __launch_bounds__(256) k_broyden_step(BroydenArgs<T, NV> g, long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    cx<T> R[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) R[j] = g.a[j][p];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx<T> s = mk<T>(T(0), T(0));
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const cx<T> m = cmul(g.M[(long long)(i * NV + j) * total + p], R[j]);
        s.x += m.x;
        s.y += m.y;
      }
      s = mk<T>(-s.x, -s.y);
      g.o0[i][p] = s;
      const cx<T> u = g.b[i][p];
      g.o1[i][p] = mk<T>(u.x + s.x * T(0.5), u.y + s.y * T(0.5));
    }
  }
}

template <class T, int NV> __global__ void __launch_bounds__(256) k_broyden_update(BroydenArgs<T, NV> g, long long total) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    cx<T> sk[NV], yk[NV];
    cx<T> d = mk<T>(T(0), T(0));
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      sk[j] = g.a[j][p];
      const cx<T> r = g.b[j][p], rn = g.c[j][p];
      yk[j] = mk<T>(rn.x - r.x, rn.y - r.y);
      const cx<T> m = cmul(sk[j], yk[j]);
      d.x += m.x;
      d.y += m.y;
    }
    if (!(hypot(d.x, d.y) > T(1e-12))) continue;  // torch::where(abs(denom) > 1e-12, ..., 0)
    // 1/d (Smith's algorithm is what c10::complex division uses; the quotient below follows it)
    cx<T> M[NV][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < NV; ++j) M[i][j] = g.M[(long long)(i * NV + j) * total + p];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx<T> my = mk<T>(T(0), T(0));
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const cx<T> m = cmul(M[i][j], yk[j]);
        my.x += m.x;
        my.y += m.y;
      }
      const cx<T> w = mk<T>(sk[i].x - my.x, sk[i].y - my.y);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const cx<T> num = cmul(w, sk[j]);
        cx<T> q;
        if (fabs(d.x) >= fabs(d.y)) {
          const T r = d.y / d.x, den = d.x + d.y * r;
          q = mk<T>((num.x + num.y * r) / den, (num.y - num.x * r) / den);
        } else {
          const T r = d.x / d.y, den = d.x * r + d.y;
          q = mk<T>((num.x * r + num.y) / den, (num.y * r - num.x) / den);
        }
        g.M[(long long)(i * NV + j) * total + p] = mk<T>(M[i][j].x + q.x, M[i][j].y + q.y);
      }
    }
  }
}

template <class T, int NV>
static cudaError_t broyden_go(const LaunchCtx &lc, int update, void *M, const void *const *a, const void *const *b, const void *const *c,
                              void *const *o0, void *const *o1, long long total) {
  BroydenArgs<T, NV> g;
  g.M = (cx<T> *)M;
  for (int i = 0; i < NV; ++i) {
    g.a[i] = (const cx<T> *)a[i];
    g.b[i] = (const cx<T> *)b[i];
    g.c[i] = c ? (const cx<T> *)c[i] : nullptr;
    g.o0[i] = o0 ? (cx<T> *)o0[i] : nullptr;
    g.o1[i] = o1 ? (cx<T> *)o1[i] : nullptr;
  }
  long long gr = (total + 255) / 256;
  const long long cap = (long long)lc.sm_count * 8;
  const int grid = (int)(gr < cap ? gr : cap);
  if (update) k_broyden_update<T, NV><<<grid, 256, 0, lc.stream>>>(g, total);
  else k_broyden_step<T, NV><<<grid, 256, 0, lc.stream>>>(g, total);
  return cudaGetLastError();
}
template <class T>
cudaError_t launch_broyden(const LaunchCtx &lc, int nvar, int update, void *M, const void *const *a, const void *const *b,
                           const void *const *c, void *const *o0, void *const *o1, long long total) {
  switch (nvar) {
    case 1: return broyden_go<T, 1>(lc, update, M, a, b, c, o0, o1, total);
    case 2: return broyden_go<T, 2>(lc, update, M, a, b, c, o0, o1, total);
    case 3: return broyden_go<T, 3>(lc, update, M, a, b, c, o0, o1, total);
    case 4: return broyden_go<T, 4>(lc, update, M, a, b, c, o0, o1, total);
    case 5: return broyden_go<T, 5>(lc, update, M, a, b, c, o0, o1, total);
    case 6: return broyden_go<T, 6>(lc, update, M, a, b, c, o0, o1, total);
    default: return cudaErrorNotSupported;
  }
}
template cudaError_t launch_broyden<double>(const LaunchCtx &, int, int, void *, const void *const *, const void *const *, const void *const *,
                                            void *const *, void *const *, long long);
template cudaError_t launch_broyden<float>(const LaunchCtx &, int, int, void *, const void *const *, const void *const *, const void *const *,
                                           void *const *, void *const *, long long);

}  // namespace mrl
