// Small layout kernels of the generic slab-decomposed transforms (mrl_dist.cu).
#include "k_common.cuh"
#include "mrl_launch.h"

namespace mrl {

template <class T> __global__ void k_real_to_complex(const T *in, cx<T> *out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = mk<T>(in[i], T(0));
}
template <class T> __global__ void k_complex_real_scale(const cx<T> *in, T *out, long long n, T scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i].x * scale;
}
// 2-D inverse, one-sided form: a real field is 2 Re(H) with H built from the half spectrum alone,
//   H(x, y) = sum_kx sum_{ky = 0}^{n/2} w(ky) F(kx, ky) e^{i (kx x + ky y)},  w = 1/2 at ky = 0 and at the Nyquist index, else 1
// (the entries with ky > n/2 are conj F(-kx, -ky), which another rank holds).  Here the factor 2 is folded in:
// out[row][ky] = in[row][ky] * (1 or 2) for ky <= n/2, 0 above; rows of n/2+1 in, n out.  The real part taken at
// the end ignores Im of the DC / Nyquist columns exactly like a c2r transform does.
template <class T> __global__ void k_expand_half(const cx<T> *in, cx<T> *out, long long rows, int n) {
  const int nc = n / 2 + 1;
  const long long total = rows * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int k = (int)(i - r * n);
    cx<T> v = mk<T>(T(0), T(0));
    if (k < nc) {
      v = in[r * nc + k];
      const bool edge = k == 0 || (2 * k == n);
      if (!edge) v = mk<T>(v.x * T(2), v.y * T(2));
    }
    out[i] = v;
  }
}

template <class T>
__global__ void k_copy3d(cx<T> *dst, long long d0, long long d1, const cx<T> *src, long long s0, long long s1, long long n0, long long n1, long long w) {
  const long long total = n0 * n1 * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / w, k = i - r * w, i0 = r / n1, i1 = r - i0 * n1;
    dst[i0 * d0 + i1 * d1 + k] = src[i0 * s0 + i1 * s1 + k];
  }
}
template <class T> __global__ void k_hermitian_rows(cx<T> *a, int n, long long ncols) {
  const int first = n / 2 + 1;
  const long long total = (long long)(n - first) * ncols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ncols, c = i - r * ncols;
    const int k = first + (int)r;
    const cx<T> v = a[(long long)(n - k) * ncols + c];
    a[(long long)k * ncols + c] = mk<T>(v.x, -v.y);
  }
}

static int grid_for(const LaunchCtx &lc, long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)lc.sm_count * 16;
  return (int)(b < 1 ? 1 : b > cap ? cap : b);
}
template <class T> cudaError_t launch_real_to_complex(const LaunchCtx &lc, const T *in, cx<T> *out, long long n) {
  k_real_to_complex<T><<<grid_for(lc, n), 256, 0, lc.stream>>>(in, out, n);
  return cudaGetLastError();
}
template <class T> cudaError_t launch_complex_real_scale(const LaunchCtx &lc, const cx<T> *in, T *out, long long n, T scale) {
  k_complex_real_scale<T><<<grid_for(lc, n), 256, 0, lc.stream>>>(in, out, n, scale);
  return cudaGetLastError();
}
template <class T> cudaError_t launch_expand_half(const LaunchCtx &lc, const cx<T> *in, cx<T> *out, long long rows, int n) {
  k_expand_half<T><<<grid_for(lc, rows * n), 256, 0, lc.stream>>>(in, out, rows, n);
  return cudaGetLastError();
}
template <class T>
cudaError_t launch_copy3d(const LaunchCtx &lc, cx<T> *dst, long long d0, long long d1, const cx<T> *src, long long s0, long long s1, long long n0,
                          long long n1, long long w) {
  if (n0 * n1 * w == 0) return cudaSuccess;
  k_copy3d<T><<<grid_for(lc, n0 * n1 * w), 256, 0, lc.stream>>>(dst, d0, d1, src, s0, s1, n0, n1, w);
  return cudaGetLastError();
}
template <class T> cudaError_t launch_hermitian_rows(const LaunchCtx &lc, cx<T> *a, int n, long long ncols) {
  const long long total = (long long)(n - n / 2 - 1) * ncols;
  if (total <= 0) return cudaSuccess;
  k_hermitian_rows<T><<<grid_for(lc, total), 256, 0, lc.stream>>>(a, n, ncols);
  return cudaGetLastError();
}
#define INST(T)                                                                                                    \
  template cudaError_t launch_real_to_complex<T>(const LaunchCtx &, const T *, cx<T> *, long long);                  \
  template cudaError_t launch_complex_real_scale<T>(const LaunchCtx &, const cx<T> *, T *, long long, T);            \
  template cudaError_t launch_expand_half<T>(const LaunchCtx &, const cx<T> *, cx<T> *, long long, int);                                          \
  template cudaError_t launch_copy3d<T>(const LaunchCtx &, cx<T> *, long long, long long, const cx<T> *, long long, long long, long long, long long, \
                                        long long);                                                                                              \
  template cudaError_t launch_hermitian_rows<T>(const LaunchCtx &, cx<T> *, int, long long);
INST(double)
INST(float)

}  // namespace mrl
