// de Geus finite-strain FFT mechanics: per-voxel constitutive law / tangent action, the
// Green-operator projection per wavevector, and the fused vector updates of the CG solver.
//
// Reference: FFTMechanics (src/tensor_computes/FFTMechanics.C:48-163), HyperElasticIsotropic
// (src/tensor_computes/HyperElasticIsotropic.C:24-52), conjugateGradientSolve
// (include/utils/MarlinUtils.h:57-131), einsum helpers (src/utils/MarlinUtils.C:147-187).
// The reference materialises C4, K4 (81 values per voxel) and Ghat4 (81 complex values per
// wavevector) and contracts them with einsum; here the same contractions are written out in
// closed form and evaluated in registers, so nothing but F, K, mu and the iterate is read:
//   S      = K tr(E) I + 2 mu (E - tr(E)/3 I),  E = (F^T F - I)/2          (C4 : E)
//   P      = F S                                                           (dot22(F, S))
//   K4 : x = x S + F T(F^T x),  T(W) = K tr(W) I + 2 mu (sym W - tr(W)/3 I) (trans2(ddot42(K4, trans2 x)))
//   Ghat:A = (A q) q^T / |q|^2   (0 at q = 0)                              (ddot42(Ghat4, A))
// Fields are component-major: [D*D][nx][ny][nz] real, [D*D][nx][ny][ncp] complex (c = D i + j), with
// D = 2 or 3 the problem dimension (the reference's tensors are dim x dim, FFTMechanics.C:50-58; the
// literal 1/3 of the deviatoric split stays 1/3 in 2-D, HyperElasticIsotropic.C:45).
#include "k_common.cuh"
#include "mrl_internal.h"
#include "mrl_mech_point.cuh"

namespace mrl {

static inline int ew_grid_fwd(long long total, const LaunchCtx &lc) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)lc.sm_count * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

template <class T, int W> struct VecW {};
template <> struct VecW<double, 2> { typedef double2 type; };
template <> struct VecW<float, 2> { typedef float2 type; };
template <> struct VecW<double, 1> { typedef double type; };
template <> struct VecW<float, 1> { typedef float type; };
template <class T, int W> struct Lanes {
  T v[W];
};
template <class T, int W> __device__ __forceinline__ Lanes<T, W> ldw(const T *p) {
  typedef typename VecW<T, W>::type V;
  union { V vec; Lanes<T, W> l; } u;
  u.vec = *reinterpret_cast<const V *>(p);
  return u.l;
}
template <class T, int W> __device__ __forceinline__ void stw(T *p, const Lanes<T, W> &l) {
  typedef typename VecW<T, W>::type V;
  union { V vec; Lanes<T, W> l; } u;
  u.l = l;
  *reinterpret_cast<V *>(p) = u.vec;
}

// mode 0: out = P = F S.   mode 1: out = K4 : x (x from memory).   mode 2: same with a spatially
// constant x (9 values in xc) - the applied macroscopic strain.   mode 3: the CG direction update
// fused in: x <- r + beta x (beta = scal[SC_BETA], x written back through xw), then out = K4 : x.
// W consecutive voxels per thread (W = 2: 128-bit accesses in fp64); n must be a multiple of W.
template <class T, int MODE, int W, int D>
__global__ void __launch_bounds__(256) k_mech_pointwise(const T *F, const T *Kf, const T *muf, const T *x, MD<T, D> xc, T *out, long long n,
                                                        T scale, const T *r, T *xw, const double *scal) {
  constexpr int NC = D * D;
  const double beta = MODE == 3 ? scal[4 /* SC_BETA */] : 0.0;
  const long long nw = n / W;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nw; q += (long long)gridDim.x * blockDim.x) {
    const long long v = q * W;
    Lanes<T, W> f[NC], xv[NC], rv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) f[c] = ldw<T, W>(F + c * n + v);
    if (MODE == 1 || MODE == 3) {
#pragma unroll
      for (int c = 0; c < NC; ++c) xv[c] = ldw<T, W>(x + c * n + v);
    }
    if (MODE == 3) {
#pragma unroll
      for (int c = 0; c < NC; ++c) rv[c] = ldw<T, W>(r + c * n + v);
    }
    const Lanes<T, W> Kv = ldw<T, W>(Kf + v), muv = ldw<T, W>(muf + v);
    Lanes<T, W> o[NC];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      MD<T, D> Fm, X = xc, R;
#pragma unroll
      for (int c = 0; c < NC; ++c) Fm.a[c / D][c % D] = f[c].v[w];
      if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < NC; ++c) X.a[c / D][c % D] = xv[c].v[w];
      } else if (MODE == 3) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const T pn = (T)((double)rv[c].v[w] + beta * (double)xv[c].v[w]);  // same arithmetic as VOP_XPBY
          xv[c].v[w] = pn;
          X.a[c / D][c % D] = pn;
        }
      }
      mech_point<T, D>(MODE, Fm, Kv.v[w], muv.v[w], X, R);
#pragma unroll
      for (int c = 0; c < NC; ++c) o[c].v[w] = R.a[c / D][c % D] * scale;
    }
    if (MODE == 3) {
#pragma unroll
      for (int c = 0; c < NC; ++c) stw<T, W>(xw + c * n + v, xv[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) stw<T, W>(out + c * n + v, o[c]);
  }
}

// In place: A_ij <- (sum_k A_ik q_k) q_j / |q|^2 for every wavevector (0 at q = 0).
// D = 3: [9][n0][n1][ncp], q = (kx[ix], ky[iy], kz[iz]).  D = 2: [4][n0][ncp] (n1 = 1), q = (kx[ix], ky[iz])
// with ky the half-spectrum axis passed in `kz`.
template <class T, int D>
__global__ void __launch_bounds__(256) k_mech_project(cx<T> *A, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp) {
  const long long plane = (long long)n1 * ncp, field = (long long)n0 * plane;
  const long long total = (long long)n0 * n1 * nzc;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(w % nzc);
    const int iy = (int)((w / nzc) % n1);
    const int ix = (int)(w / ((long long)nzc * n1));
    const long long off = (long long)ix * plane + (long long)iy * ncp + iz;
    T q[D];
    q[0] = kx[ix];
    if (D == 3) q[1] = ky[iy];
    q[D - 1] = kz[iz];
    T Q = T(0);
#pragma unroll
    for (int d = 0; d < D; ++d) Q += q[d] * q[d];
    const T inv = Q == T(0) ? T(0) : T(1) / Q;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      T vx = T(0), vy = T(0);
#pragma unroll
      for (int l = 0; l < D; ++l) {
        const cx<T> a = A[(D * i + l) * field + off];
        vx += a.x * q[l];
        vy += a.y * q[l];
      }
      vx *= inv;
      vy *= inv;
#pragma unroll
      for (int j = 0; j < D; ++j) A[(D * i + j) * field + off] = mk<T>(vx * q[j], vy * q[j]);
    }
  }
}

// ---------------------------------------------------------------------------- vector algebra
// Scalars of the CG iteration live in a small device array so that no host round trip sits
// between the reductions and the updates that use them.
enum { SC_RZ = 0, SC_PAP = 1, SC_ALPHA = 2, SC_RES2 = 3, SC_BETA = 4, SC_TMP = 5, SC_COUNT = 8 };
enum { VOP_DOT = 0, VOP_CG_XR = 1, VOP_XPBY = 2, VOP_AXPY = 3, VOP_SUB = 4, VOP_COPY = 5 };
enum { FIN_STORE = 0, FIN_ALPHA = 1, FIN_RES = 2 };

template <class T> __device__ __forceinline__ double block_sum(double acc) {
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = acc;
  __syncthreads();
  double r = 0;
  if (warp == 0) {
    r = sm[lane & 7];
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

// VOP_DOT   : partial sums of a.b
// VOP_CG_XR : x += alpha p (a = p, y = x);  r -= alpha Ap (b = Ap, z = r);  partial sums of r.r
// VOP_XPBY  : y = a + beta y           (p = r + beta p)
// VOP_AXPY  : y += s a                 (s = host scalar)
// VOP_SUB   : z = a - b                (r = b - A x)
// VOP_COPY  : y = a
template <class T> struct Vec2 {};
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

template <class T> __device__ __forceinline__ double vec_elem(int op, double alpha, T a, T b, T &y, T &z) {
  if (op == VOP_DOT) return (double)a * (double)b;
  if (op == VOP_CG_XR) {
    y = (T)((double)y + alpha * (double)a);
    const double r = (double)z - alpha * (double)b;
    z = (T)r;
    return r * r;
  }
  if (op == VOP_XPBY) y = (T)((double)a + alpha * (double)y);
  else if (op == VOP_AXPY) y = (T)((double)y + alpha * (double)a);
  else if (op == VOP_SUB) z = (T)((double)a - (double)b);
  else y = a;
  return 0.0;
}

// OP is a template parameter so that each operation loads and stores only its own operands;
// two elements per thread and iteration (128-bit accesses for fp64), all loads of an iteration
// issued before the first use.
template <class T, int OP>
__global__ void __launch_bounds__(256) k_vec(const T *a, const T *b, T *y, T *z, const double *scal, double s, long long n,
                                             double *partials) {
  typedef typename Vec2<T>::type V;
  constexpr bool RA = true, RB = (OP == VOP_DOT || OP == VOP_CG_XR || OP == VOP_SUB);
  constexpr bool RY = (OP == VOP_CG_XR || OP == VOP_XPBY || OP == VOP_AXPY), RZ = (OP == VOP_CG_XR);
  constexpr bool WY = (OP == VOP_CG_XR || OP == VOP_XPBY || OP == VOP_AXPY || OP == VOP_COPY), WZ = (OP == VOP_CG_XR || OP == VOP_SUB);
  double acc = 0;
  const double alpha = (OP == VOP_CG_XR) ? scal[SC_ALPHA] : (OP == VOP_XPBY ? scal[SC_BETA] : s);
  const long long n2 = n >> 1;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n2; i += nth) {
    V va = RA ? reinterpret_cast<const V *>(a)[i] : V{0, 0};
    V vb = RB ? reinterpret_cast<const V *>(b)[i] : V{0, 0};
    V vy = RY ? reinterpret_cast<const V *>(y)[i] : V{0, 0};
    V vz = RZ ? reinterpret_cast<const V *>(z)[i] : V{0, 0};
    acc += vec_elem<T>(OP, alpha, va.x, vb.x, vy.x, vz.x);
    acc += vec_elem<T>(OP, alpha, va.y, vb.y, vy.y, vz.y);
    if (WY) reinterpret_cast<V *>(y)[i] = vy;
    if (WZ) reinterpret_cast<V *>(z)[i] = vz;
  }
  if ((n & 1) && tid == 0) {
    const long long i = n - 1;
    T ea = a[i], eb = RB ? b[i] : T(0), ey = RY ? y[i] : T(0), ez = RZ ? z[i] : T(0);
    acc += vec_elem<T>(OP, alpha, ea, eb, ey, ez);
    if (WY) y[i] = ey;
    if (WZ) z[i] = ez;
  }
  if (OP == VOP_DOT || OP == VOP_CG_XR) {
    const double r = block_sum<T>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
  }
}

// single block: finish a reduction and derive the dependent scalars
__global__ void __launch_bounds__(256) k_vec_final(int fin, int slot, const double *partials, int nblk, double *scal) {
  double acc = 0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) acc += partials[i];
  const double r = block_sum<double>(acc);
  if (threadIdx.x == 0) {
    if (fin == FIN_STORE) {
      scal[slot] = r;
    } else if (fin == FIN_ALPHA) {  // r = p.Ap
      scal[SC_PAP] = r;
      scal[SC_ALPHA] = scal[SC_RZ] / r;
    } else {  // r = new r.r
      scal[SC_RES2] = r;
      scal[SC_BETA] = r / scal[SC_RZ];
      scal[SC_RZ] = r;
    }
  }
}

// y[c][v] += s[c] (D*D constants): F + applied macroscopic strain
template <class T, int D> __global__ void __launch_bounds__(256) k_add_const9(T *y, MD<T, D> s, long long n) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < D * D; ++c) y[c * n + v] += s.a[c / D][c % D];
  }
}

// von Mises stress of a component-major stress field (ComputeVonMisesStress.C:37-63, both forms as coded)
template <class T, int D> __global__ void __launch_bounds__(256) k_von_mises(const T *s, T *out, long long n) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    auto S = [&](int i, int j) { return s[(D * i + j) * n + v]; };
    T t;
    if (D == 3) {
      const T a = S(0, 0) - S(1, 1), b = S(1, 1) - S(2, 2), c = S(2, 2) - S(0, 0);
      const T xy = S(0, 1), yz = S(1, 2), zx = S(2, 0);
      t = a * a + b * b + c * c + T(6) * (xy * xy + yz * yz + zx * zx);
    } else {
      const T a = S(0, 0) - S(1, 1), xy = S(0, 1);
      t = a * a + T(6) * (xy * xy);
    }
    out[v] = sqrt(T(0.5) * t);
  }
}
template <class T> cudaError_t launch_von_mises(const LaunchCtx &lc, int dim, const T *s, T *out, long long n) {
  if (dim == 3) k_von_mises<T, 3><<<ew_grid_fwd(n, lc), 256, 0, lc.stream>>>(s, out, n);
  else k_von_mises<T, 2><<<ew_grid_fwd(n, lc), 256, 0, lc.stream>>>(s, out, n);
  return cudaGetLastError();
}

// ComputeDisplacements.C:74-78: per wavevector u_i = sum_j Hbar_ij q_j (-i) / |q|^2 (0 at q = 0), written in
// place into component i of the spectra (row i = components D i .. D i + D - 1 is read first).
template <class T, int D>
__global__ void __launch_bounds__(256) k_disp_contract(cx<T> *H, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp) {
  const long long plane = (long long)n1 * ncp, field = (long long)n0 * plane;
  const long long total = (long long)n0 * n1 * nzc;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(w % nzc);
    const int iy = (int)((w / nzc) % n1);
    const int ix = (int)(w / ((long long)nzc * n1));
    const long long off = (long long)ix * plane + (long long)iy * ncp + iz;
    T q[D];
    q[0] = kx[ix];
    if (D == 3) q[1] = ky[iy];
    q[D - 1] = kz[iz];
    T Q = T(0);
#pragma unroll
    for (int d = 0; d < D; ++d) Q += q[d] * q[d];
    cx<T> u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      T sx = T(0), sy = T(0);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const cx<T> h = H[(D * i + j) * field + off];
        sx += h.x * q[j];
        sy += h.y * q[j];
      }
      // (sx + i sy) * (-i) = sy - i sx
      u[i] = Q == T(0) ? mk<T>(T(0), T(0)) : mk<T>(sy / Q, -sx / Q);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) H[i * field + off] = u[i];
  }
}

// ComputeDisplacements.C:84-106: out_i(node) = interpolate(u_aff_i + u_per_i) on the (n+1)^D nodal grid,
// (bi/tri)linear with align_corners = true (source coordinate = node * (n-1)/n);
// u_aff_i(cell) = sum_j A_ij X_j(cell), A = <F> - I, X the cell-centre axes.
template <class T, int D> struct DispArgs {
  const T *uper;   // [D][n0*n1*n2]
  T *out;          // [D][(n0+1)(n1+1)(n2+1)]
  const T *ax[3];  // cell-centre axes
  int n[3];
  T A[D][D];
};
template <class T, int D> __global__ void __launch_bounds__(256) k_disp_nodal(DispArgs<T, D> a) {
  const int m0 = a.n[0] + 1, m1 = D >= 2 ? a.n[1] + 1 : 1, m2 = D == 3 ? a.n[2] + 1 : 1;
  const long long nodes = (long long)m0 * m1 * m2, cells = (long long)a.n[0] * (D >= 2 ? a.n[1] : 1) * (D == 3 ? a.n[2] : 1);
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < nodes; w += (long long)gridDim.x * blockDim.x) {
    int p[3] = {0, 0, 0};
    long long r = w;
    for (int d = D - 1; d >= 0; --d) {
      const int m = a.n[d] + 1;
      p[d] = (int)(r % m);
      r /= m;
    }
    int i0[3] = {0, 0, 0}, i1[3] = {0, 0, 0};
    double l1[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) {
      const double src = a.n[d] > 0 ? (double)p[d] * ((double)(a.n[d] - 1) / (double)a.n[d]) : 0.0;
      i0[d] = (int)src;
      i1[d] = i0[d] + (i0[d] < a.n[d] - 1 ? 1 : 0);
      l1[d] = src - i0[d];
    }
    double acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.0;
    for (int corner = 0; corner < (1 << D); ++corner) {
      double wgt = 1.0;
      int c[3] = {0, 0, 0};
      for (int d = 0; d < D; ++d) {
        const int hi = (corner >> d) & 1;
        c[d] = hi ? i1[d] : i0[d];
        wgt *= hi ? l1[d] : 1.0 - l1[d];
      }
      long long cell = c[0];
      if (D >= 2) cell = cell * a.n[1] + c[1];
      if (D == 3) cell = cell * a.n[2] + c[2];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double v = (double)a.uper[i * cells + cell];
#pragma unroll
        for (int j = 0; j < D; ++j) v += (double)a.A[i][j] * (double)a.ax[j][c[j]];
        acc[i] += wgt * v;
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) a.out[i * nodes + w] = (T)acc[i];
  }
}

template <class T>
cudaError_t launch_disp_contract(const LaunchCtx &lc, int dim, cx<T> *H, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp) {
  const int grid = ew_grid_fwd((long long)n0 * n1 * nzc, lc);
  if (dim == 3) k_disp_contract<T, 3><<<grid, 256, 0, lc.stream>>>(H, kx, ky, kz, n0, n1, nzc, ncp);
  else k_disp_contract<T, 2><<<grid, 256, 0, lc.stream>>>(H, kx, ky, kz, n0, n1, nzc, ncp);
  return cudaGetLastError();
}
template <class T, int D>
static cudaError_t disp_nodal_go(const LaunchCtx &lc, const T *uper, T *out, const T *const *ax, const int *n, const double *A) {
  DispArgs<T, D> a;
  a.uper = uper;
  a.out = out;
  long long nodes = 1;
  for (int d = 0; d < 3; ++d) {
    a.ax[d] = ax[d];
    a.n[d] = d < D ? n[d] : 1;
    if (d < D) nodes *= n[d] + 1;
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) a.A[i][j] = (T)A[D * i + j];
  k_disp_nodal<T, D><<<ew_grid_fwd(nodes, lc), 256, 0, lc.stream>>>(a);
  return cudaGetLastError();
}
template <class T>
cudaError_t launch_disp_nodal(const LaunchCtx &lc, int dim, const T *uper, T *out, const T *const *ax, const int *n, const double *A) {
  return dim == 3 ? disp_nodal_go<T, 3>(lc, uper, out, ax, n, A) : disp_nodal_go<T, 2>(lc, uper, out, ax, n, A);
}

// [n][ncomp] (reference layout, components fastest) <-> [ncomp][n]
template <class T> __global__ void __launch_bounds__(256) k_components(const T *in, T *out, long long n, int ncomp, int to_soa) {
  const long long total = n * ncomp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / ncomp;
    const int c = (int)(i - v * ncomp);
    if (to_soa) out[c * n + v] = in[i];
    else out[i] = in[c * n + v];
  }
}

static inline int ew_grid(long long total, const LaunchCtx &lc) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)lc.sm_count * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

template <class T, int W, int D>
static void mech_pointwise_go(const LaunchCtx &lc, int mode, const T *F, const T *K, const T *mu, const T *x, const double *xc, T *out,
                              long long n, T scale, const T *r, T *xw, const double *scal) {
  MD<T, D> X;
  for (int c = 0; c < D * D; ++c) X.a[c / D][c % D] = xc ? (T)xc[c] : T(0);
  const int grid = ew_grid(n / W, lc);
  switch (mode) {
    case 0: k_mech_pointwise<T, 0, W, D><<<grid, 256, 0, lc.stream>>>(F, K, mu, x, X, out, n, scale, r, xw, scal); break;
    case 1: k_mech_pointwise<T, 1, W, D><<<grid, 256, 0, lc.stream>>>(F, K, mu, x, X, out, n, scale, r, xw, scal); break;
    case 2: k_mech_pointwise<T, 2, W, D><<<grid, 256, 0, lc.stream>>>(F, K, mu, x, X, out, n, scale, r, xw, scal); break;
    default: k_mech_pointwise<T, 3, W, D><<<grid, 256, 0, lc.stream>>>(F, K, mu, x, X, out, n, scale, r, xw, scal); break;
  }
}
// dim = 2 or 3: tensors are dim x dim, xc holds dim*dim values
template <class T>
cudaError_t launch_mech_pointwise(const LaunchCtx &lc, int dim, int mode, const T *F, const T *K, const T *mu, const T *x, const double *xc,
                                  T *out, long long n, double scale, const T *r, T *xw, const double *scal) {
  bool wide = (n % 2) == 0;
  for (const void *q : {(const void *)F, (const void *)K, (const void *)mu, (const void *)x, (const void *)out, (const void *)r, (const void *)xw})
    wide = wide && (((unsigned long long)q & (2 * sizeof(T) - 1)) == 0);
  if (dim == 3) {
    if (wide) mech_pointwise_go<T, 2, 3>(lc, mode, F, K, mu, x, xc, out, n, (T)scale, r, xw, scal);
    else mech_pointwise_go<T, 1, 3>(lc, mode, F, K, mu, x, xc, out, n, (T)scale, r, xw, scal);
  } else {
    if (wide) mech_pointwise_go<T, 2, 2>(lc, mode, F, K, mu, x, xc, out, n, (T)scale, r, xw, scal);
    else mech_pointwise_go<T, 1, 2>(lc, mode, F, K, mu, x, xc, out, n, (T)scale, r, xw, scal);
  }
  return cudaGetLastError();
}
template <class T>
cudaError_t launch_mech_project(const LaunchCtx &lc, int dim, cx<T> *A, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp) {
  const int grid = ew_grid((long long)n0 * n1 * nzc, lc);
  if (dim == 3) k_mech_project<T, 3><<<grid, 256, 0, lc.stream>>>(A, kx, ky, kz, n0, n1, nzc, ncp);
  else k_mech_project<T, 2><<<grid, 256, 0, lc.stream>>>(A, kx, ky, kz, n0, n1, nzc, ncp);
  return cudaGetLastError();
}
template <class T>
cudaError_t launch_vec(const LaunchCtx &lc, int op, const T *a, const T *b, T *y, T *z, double *scal, double s, long long n, int fin, int slot,
                       double *partials, int nblk) {
  for (const void *q : {(const void *)a, (const void *)b, (const void *)y, (const void *)z})
    if ((unsigned long long)q & 15ull) return cudaErrorMisalignedAddress;
  switch (op) {
    case VOP_DOT: k_vec<T, VOP_DOT><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
    case VOP_CG_XR: k_vec<T, VOP_CG_XR><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
    case VOP_XPBY: k_vec<T, VOP_XPBY><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
    case VOP_AXPY: k_vec<T, VOP_AXPY><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
    case VOP_SUB: k_vec<T, VOP_SUB><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
    default: k_vec<T, VOP_COPY><<<nblk, 256, 0, lc.stream>>>(a, b, y, z, scal, s, n, partials); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (op == VOP_DOT || op == VOP_CG_XR) {
    k_vec_final<<<1, 256, 0, lc.stream>>>(fin, slot, partials, nblk, scal);
    e = cudaGetLastError();
  }
  return e;
}
// finish a reduction whose partial sums another kernel left behind (the c2r pass with the fused inner product)
template <class T> cudaError_t launch_vec_final(const LaunchCtx &lc, int fin, int slot, const double *partials, int nblk, double *scal) {
  k_vec_final<<<1, 256, 0, lc.stream>>>(fin, slot, partials, nblk, scal);
  return cudaGetLastError();
}
template <class T> cudaError_t launch_add_const9(const LaunchCtx &lc, int dim, T *y, const double *s, long long n) {
  if (dim == 3) {
    MD<T, 3> S;
    for (int c = 0; c < 9; ++c) S.a[c / 3][c % 3] = (T)s[c];
    k_add_const9<T, 3><<<ew_grid(n, lc), 256, 0, lc.stream>>>(y, S, n);
  } else {
    MD<T, 2> S;
    for (int c = 0; c < 4; ++c) S.a[c / 2][c % 2] = (T)s[c];
    k_add_const9<T, 2><<<ew_grid(n, lc), 256, 0, lc.stream>>>(y, S, n);
  }
  return cudaGetLastError();
}
template <class T> cudaError_t launch_components(const LaunchCtx &lc, const T *in, T *out, long long n, int ncomp, int to_soa) {
  k_components<T><<<ew_grid(n * ncomp, lc), 256, 0, lc.stream>>>(in, out, n, ncomp, to_soa);
  return cudaGetLastError();
}

#define INST(T)                                                                                                                   \
  template cudaError_t launch_mech_pointwise<T>(const LaunchCtx &, int, int, const T *, const T *, const T *, const T *, const double *, \
                                                T *, long long, double, const T *, T *, const double *);                          \
  template cudaError_t launch_vec_final<T>(const LaunchCtx &, int, int, const double *, int, double *);                            \
  template cudaError_t launch_mech_project<T>(const LaunchCtx &, int, cx<T> *, const T *, const T *, const T *, int, int, int, int); \
  template cudaError_t launch_vec<T>(const LaunchCtx &, int, const T *, const T *, T *, T *, double *, double, long long, int, int, \
                                     double *, int);                                                                              \
  template cudaError_t launch_add_const9<T>(const LaunchCtx &, int, T *, const double *, long long);                                    \
  template cudaError_t launch_components<T>(const LaunchCtx &, const T *, T *, long long, int, int);                              \
  template cudaError_t launch_von_mises<T>(const LaunchCtx &, int, const T *, T *, long long);                                    \
  template cudaError_t launch_disp_contract<T>(const LaunchCtx &, int, cx<T> *, const T *, const T *, const T *, int, int, int, int); \
  template cudaError_t launch_disp_nodal<T>(const LaunchCtx &, int, const T *, T *, const T *const *, const int *, const double *);
INST(double)
INST(float)

}  // namespace mrl
