// marlin_b200 - FFT pass kernels (one HBM round trip per axis) and the fused k-space update.
//
// A 3-D real transform is three passes:  last-axis real pass (r2c / c2r, two real rows per
// complex pencil)  +  one strided complex pass per remaining axis.  The semi-implicit
// Cahn-Hilliard substep (reference: src/tensor_solver/AdamsBashforthMoulton.C:60-101 with the
// root compute of examples/cahn_hilliard/cahnhilliard2.i) is five passes:
//   P1  z r2c of (c + i f'(c))            reads c, writes c^ and mu^ partial spectra
//   P2  y forward on both fields          (skipped in 2-D)
//   P3  x forward on both, AB/semi-implicit update in registers, x inverse on u^   (fused)
//   P4  y inverse                         (skipped in 2-D)
//   P5  z c2r                             writes the new c
// Each kernel exists in two flavours sharing the same load/store functors:
//   *_fast : RegFFT (registers + one smem exchange per stage) for sizes with a FFTCfg,
//   *_gen  : smem_fft (runtime mixed radix, any N).
#pragma once
#include "mrl_fft.cuh"

namespace mrl {

// Row remap for a pass over a y-chunk of a slab [nx][nyl][..]: the pass enumerates rows r' of the chunk
// [nx][ych] and works on row (r' / ych) * nyl + y0 + r' % ych of the full arrays.  ych = 0: identity.
// (PPB must divide ych so that the rows of a tile stay consecutive.)
struct RowMap {
  int ych, nyl, y0;
  MRL_DI long long operator()(long long r) const { return ych ? (r / ych) * nyl + y0 + r % ych : r; }
};


// Arguments of the mechanics' first pass with the tangent fused into its load (k_mech_tangent_zfwd, mrl_mech_tma.cuh)
template <class T> struct MechTangentIO {
  const T *F, *K, *mu;  // [9][n], [n], [n]
  T *p;                 // direction [9][n]: read; replaced by r + beta p when r != nullptr
  const T *r;           // residual [9][n] or nullptr
  const double *scal;   // device scalars of the CG recurrence (beta at index 4)
  long long n;          // voxels
  long long nrows;      // rows of the last axis per component (n / N, even)
  cx<T> *out;           // [9][nrows][ncp] half spectra
  int ncp;
};

// ======================================================================== strided c2c pass
// Data is [nfields][nouter][n][ncols] complex, transform along n (stride `pitch`).
template <class T> struct StridedIO {
  const cx<T> *in[4];
  cx<T> *out[4];
  int nfields;
  int n, ncols, nouter;
  long long pitch, outer_stride;
  int ncb;  // column blocks (of TK) per outer slice
  T scale;  // multiplies the result on store
  int inverse;
  // optional separate output layout (TMA kernel only; 0 = same as the input layout)
  long long out_pitch, out_outer_stride;
  int nvalid;  // TMA kernel only: columns >= nvalid are padding (0 = all ncols valid)
  // TMA kernel only, multi-GPU slab: rows of the result are scattered straight into the peers'
  // receive staging (peer_tab[s] = base of rank s's buffer, device array of nranks pointers):
  // row r of field f goes to peer r / peer_rows at  f*peer_field + peer_off + (r % peer_rows)*pitch
  const unsigned long long *peer_tab;
  int peer_rows;
  long long peer_field, peer_off;

  MRL_HD int ntiles() const { return nfields * nouter * ncb; }
};

template <class T, class C, int TK>
__global__ void __launch_bounds__(TK *C::TP) k_strided_fast(StridedIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E;
  MRL_DYN_SMEM(smem_raw);
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *tw = buf + N * TK;
  const int tid = threadIdx.x;
  for (int i = tid; i < N; i += TK * TP) tw[i] = tw_g[i];
  __syncthreads();
  const int col = tid % TK, t = tid / TK;
  const SmTile<T, TK> sm{buf, col};
  const int ntiles = io.ntiles();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int f = tile / (io.nouter * io.ncb);
    const int rem = tile - f * (io.nouter * io.ncb);
    const int o = rem / io.ncb, c = (rem - o * io.ncb) * TK + col;
    const bool ok = c < io.ncols;
    const long long off = (long long)o * io.outer_stride + c;
    const cx<T> *src = io.in[f] + off;
    cx<T> *dst = io.out[f] + off;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      v[e] = ok ? src[(long long)(t + TP * e) * io.pitch] : mk<T>(T(0), T(0));
      if (io.inverse) v[e].y = -v[e].y;
    }
    RegFFT<T, C>::run(v, t, sm, tw);
    if (ok) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        cx<T> r = mk<T>(v[e].x * io.scale, (io.inverse ? -v[e].y : v[e].y) * io.scale);
        dst[(long long)(t + TP * e) * io.pitch] = r;
      }
    }
  }
}

template <class T, int TK> __global__ void __launch_bounds__(256) k_strided_gen(StridedIO<T> io, const cx<T> *tw, FFTPlanDev plan) {
  MRL_DYN_SMEM(smem_raw);
  const int n = plan.n;
  cx<T> *A = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *B = A + n * TK;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ntiles = io.ntiles();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int f = tile / (io.nouter * io.ncb);
    const int rem = tile - f * (io.nouter * io.ncb);
    const int o = rem / io.ncb, c0 = (rem - o * io.ncb) * TK;
    const long long off = (long long)o * io.outer_stride + c0;
    const cx<T> *src = io.in[f] + off;
    cx<T> *dst = io.out[f] + off;
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      cx<T> v = (c0 + c < io.ncols) ? src[(long long)j * io.pitch + c] : mk<T>(T(0), T(0));
      if (io.inverse) v.y = -v.y;
      A[w] = v;
    }
    __syncthreads();
    const cx<T> *R = smem_fft<T, TK>(A, B, tw, plan, tid, nt);
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      if (c0 + c < io.ncols) {
        cx<T> v = R[w];
        dst[(long long)j * io.pitch + c] = mk<T>(v.x * io.scale, (io.inverse ? -v.y : v.y) * io.scale);
      }
    }
    __syncthreads();
  }
}

// ======================================================================== last-axis r2c pass
// Load functor:  cx<T> ld(p, j)            -> (a_j, b_j) of complex pencil p
// Store functor: void st(p, k, A_k, B_k)   -> half spectra, k in [0, n/2]
template <class T> MRL_DI void r2c_separate(cx<T> z, cx<T> w, cx<T> &A, cx<T> &B) {
  // z = Z_k, w = Z_{n-k}:  A = (z + conj w)/2,  B = (z - conj w)/(2i)
  A = mk<T>(T(0.5) * (z.x + w.x), T(0.5) * (z.y - w.y));
  B = mk<T>(T(0.5) * (z.y + w.y), T(-0.5) * (z.x - w.x));
}

// Generic pairing of adjacent rows of a [nrows][n] real array (any number of fields laid out
// back to back): pencil p = rows 2p, 2p+1.
template <class T> struct ZLoadPairs {
  const T *in;
  long long nrows;
  int n;
  MRL_DI cx<T> ld(long long p, int j) const {
    const long long r0 = 2 * p;
    const T a = in[r0 * n + j];
    const T b = (r0 + 1 < nrows) ? in[(r0 + 1) * n + j] : T(0);
    return mk<T>(a, b);
  }
};
template <class T> struct ZStorePairs {
  cx<T> *out;
  long long nrows;
  int nc;  // n/2+1
  MRL_DI void st(long long p, int k, cx<T> A, cx<T> B) const {
    const long long r0 = 2 * p;
    out[r0 * nc + k] = A;
    if (r0 + 1 < nrows) out[(r0 + 1) * nc + k] = B;
  }
};
// Cahn-Hilliard P1: a = c, b = F(c) (the real-space nonlinearity), one pencil per row.
template <class T, class F> struct ZLoadFused {
  const T *c;
  T *mu_out;  // optional copy of the nonlinearity in real space (nullptr = skip)
  int n;
  F f;
  MRL_DI cx<T> ld(long long p, int j) const {
    const T a = c[p * n + j];
    const T b = f(a, p * n + j);
    if (mu_out) mu_out[p * n + j] = b;
    return mk<T>(a, b);
  }
};
template <class T> struct ZStoreTwo {
  cx<T> *outA, *outB;
  int nc;
  MRL_DI void st(long long p, int k, cx<T> A, cx<T> B) const {
    outA[p * nc + k] = A;
    outB[p * nc + k] = B;
  }
};

template <class T, class C, int PPB, class LD, class ST>
__global__ void __launch_bounds__(PPB *C::TP) k_zfwd_fast(LD ld, ST st, const cx<T> *tw_g, long long npencils) {
  constexpr int N = C::N, TP = C::TP, E = C::E;
  constexpr int NP = N + (N >> 3) + 1;
  MRL_DYN_SMEM(smem_raw);
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *tw = buf + NP * PPB;
  const int tid = threadIdx.x;
  for (int i = tid; i < N; i += PPB * TP) tw[i] = tw_g[i];
  __syncthreads();
  const int t = tid % TP, pl = tid / TP;
  const SmPencil<T> sm{buf + pl * NP};
  const long long nblk = (npencils + PPB - 1) / PPB;
  for (long long pb = blockIdx.x; pb < nblk; pb += gridDim.x) {
    const long long p = pb * PPB + pl;
    const bool ok = p < npencils;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = ok ? ld.ld(p, t + TP * e) : mk<T>(T(0), T(0));
    RegFFT<T, C>::run(v, t, sm, tw);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) sm.st(t + TP * e, v[e]);
    __syncthreads();
    if (ok) {
      for (int k = t; k <= N / 2; k += TP) {
        cx<T> A, B;
        r2c_separate(sm.ld(k), sm.ld(k == 0 ? 0 : N - k), A, B);
        st.st(p, k, A, B);
      }
    }
    __syncthreads();
  }
}

template <class T, int TK, class LD, class ST>
__global__ void __launch_bounds__(256) k_zfwd_gen(LD ld, ST st, const cx<T> *tw, FFTPlanDev plan, long long npencils) {
  MRL_DYN_SMEM(smem_raw);
  const int n = plan.n;
  cx<T> *A = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *B = A + n * TK;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long nblk = (npencils + TK - 1) / TK;
  for (long long pb = blockIdx.x; pb < nblk; pb += gridDim.x) {
    for (int w = tid; w < n * TK; w += nt) {
      const int c = w / n, j = w - c * n;
      const long long p = pb * TK + c;
      A[j * TK + c] = (p < npencils) ? ld.ld(p, j) : mk<T>(T(0), T(0));
    }
    __syncthreads();
    const cx<T> *R = smem_fft<T, TK>(A, B, tw, plan, tid, nt);
    const int nc = n / 2 + 1;
    for (int w = tid; w < nc * TK; w += nt) {
      const int c = w / nc, k = w - c * nc;
      const long long p = pb * TK + c;
      if (p < npencils) {
        cx<T> Ak, Bk;
        r2c_separate(R[k * TK + c], R[(k == 0 ? 0 : n - k) * TK + c], Ak, Bk);
        st.st(p, k, Ak, Bk);
      }
    }
    __syncthreads();
  }
}

// ======================================================================== last-axis c2r pass
// Load functor:  cx<T> ld(p, idx)  -> conj(Z_idx), Z = X + iY rebuilt from the two half spectra
// Store functor: void st(p, j, a_j, b_j)
template <class T> MRL_DI cx<T> c2r_merge_conj(cx<T> X, cx<T> Y, int k, int n, bool mirrored) {
  // DC and (even n) Nyquist bins: imaginary parts are ignored, like pocketfft/MKL/cuFFT c2r
  if (k == 0 || 2 * k == n) {
    X.y = T(0);
    Y.y = T(0);
  }
  // Z_k = X + iY ; Z_{n-k} = conj(X) + i conj(Y); return the complex conjugate of it
  return mirrored ? mk<T>(X.x + Y.y, X.y - Y.x) : mk<T>(X.x - Y.y, -(X.y + Y.x));
}
template <class T> struct ZInvLoadPairs {
  const cx<T> *in;
  long long nrows;
  int n, nc;
  MRL_DI cx<T> ld(long long p, int idx) const {
    const bool mir = idx > n / 2;
    const int k = mir ? n - idx : idx;
    const long long r0 = 2 * p;
    const cx<T> X = in[r0 * nc + k];
    const cx<T> Y = (r0 + 1 < nrows) ? in[(r0 + 1) * nc + k] : mk<T>(T(0), T(0));
    return c2r_merge_conj(X, Y, k, n, mir);
  }
};
template <class T> struct ZInvStorePairs {
  T *out;
  long long nrows;
  int n;
  T scale;
  MRL_DI void st(long long p, int j, T a, T b) const {
    const long long r0 = 2 * p;
    out[r0 * n + j] = a * scale;
    if (r0 + 1 < nrows) out[(r0 + 1) * n + j] = b * scale;
  }
};

template <class T, class C, int PPB, class LD, class ST>
__global__ void __launch_bounds__(PPB *C::TP) k_zinv_fast(LD ld, ST st, const cx<T> *tw_g, long long npencils) {
  constexpr int N = C::N, TP = C::TP, E = C::E;
  constexpr int NP = N + (N >> 3) + 1;
  MRL_DYN_SMEM(smem_raw);
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *tw = buf + NP * PPB;
  const int tid = threadIdx.x;
  for (int i = tid; i < N; i += PPB * TP) tw[i] = tw_g[i];
  __syncthreads();
  const int t = tid % TP, pl = tid / TP;
  const SmPencil<T> sm{buf + pl * NP};
  const long long nblk = (npencils + PPB - 1) / PPB;
  for (long long pb = blockIdx.x; pb < nblk; pb += gridDim.x) {
    const long long p = pb * PPB + pl;
    const bool ok = p < npencils;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = ok ? ld.ld(p, t + TP * e) : mk<T>(T(0), T(0));
    RegFFT<T, C>::run(v, t, sm, tw);
    if (ok) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e) st.st(p, t + TP * e, v[e].x, -v[e].y);
    }
  }
}

template <class T, int TK, class LD, class ST>
__global__ void __launch_bounds__(256) k_zinv_gen(LD ld, ST st, const cx<T> *tw, FFTPlanDev plan, long long npencils) {
  MRL_DYN_SMEM(smem_raw);
  const int n = plan.n;
  cx<T> *A = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *B = A + n * TK;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long nblk = (npencils + TK - 1) / TK;
  for (long long pb = blockIdx.x; pb < nblk; pb += gridDim.x) {
    for (int w = tid; w < n * TK; w += nt) {
      const int c = w / n, j = w - c * n;
      const long long p = pb * TK + c;
      A[j * TK + c] = (p < npencils) ? ld.ld(p, j) : mk<T>(T(0), T(0));
    }
    __syncthreads();
    const cx<T> *R = smem_fft<T, TK>(A, B, tw, plan, tid, nt);
    for (int w = tid; w < n * TK; w += nt) {
      const int c = w / n, j = w - c * n;
      const long long p = pb * TK + c;
      if (p < npencils) {
        const cx<T> r = R[j * TK + c];
        st.st(p, j, r.x, -r.y);
      }
    }
    __syncthreads();
  }
}

// ======================================================================== fused k-space update
// Semi-implicit Adams-Bashforth update of one variable, per wavevector
// (src/tensor_solver/AdamsBashforthMoulton.C:94-99):
//   N    = Mbar(k) * g^                      (ParsedCompute 'Mbar*mubar')
//   u^   = c^ + dt*beta0*N + sum_i dt*beta_{i+1}*N_old[i]
//   u^  /= (1 - dt*L(k))                     (only when a linear term exists)
// Mbar and L either come from closed forms of k^2 computed from the three reciprocal axis
// vectors (ReciprocalLaplacianFactor.C:30: -k2*f, ReciprocalLaplacianSquareFactor.C:31:
// k2*k2*f) or are read from real full-size buffers.
enum { MRL_KMODE_2D = 0, MRL_KMODE_3D = 1, MRL_KMODE_3D_SLAB = 2 };
template <class T> struct SpectralUpdate {
  const T *kx, *ky, *kz;  // reciprocal axes (2 pi fftfreq / rfftfreq)
  int kmode;
  int nzc;        // row pitch of the last (halved) axis
  int nzv;        // valid entries of the last axis (0 = nzc; TMA kernel only)
  int x0;         // first global x index of this rank's slab (slab mode)
  int closed_M, closed_L, has_L;
  T Mfac, Lfac;
  const T *Mbuf, *Lbuf;   // full reciprocal-shape real buffers when not closed form
  T dt;
  T b0;           // dt*beta[order][0]
  int nold;
  T bold[4];      // dt*beta[order][i+1]
  const cx<T> *Nold[4];
  cx<T> *Nout;    // new nonlinear term (history for the next substeps); may be null

  // (o, j, col) -> k^2 and linear element offset `off` within a reciprocal-shape array
  MRL_DI T k2(int o, int j, int col) const {
    T a, b, c;
    if (kmode == MRL_KMODE_3D) {
      a = kx[j]; b = ky[col / nzc]; c = kz[col % nzc];
    } else if (kmode == MRL_KMODE_3D_SLAB) {
      a = kx[x0 + o]; b = ky[j]; c = kz[col];
    } else {
      a = kx[j]; b = ky[col]; c = T(0);
    }
    return a * a + b * b + c * c;
  }
  MRL_DI cx<T> apply(int o, int j, int col, long long off, cx<T> chat, cx<T> ghat) const {
    const T kk = k2(o, j, col);
    const T M = closed_M == 1 ? (-kk * Mfac) : closed_M == 2 ? T(1) : Mbuf[off];
    const cx<T> N = mk<T>(M * ghat.x, M * ghat.y);
    if (Nout) Nout[off] = N;
    cx<T> u = mk<T>(chat.x + b0 * N.x, chat.y + b0 * N.y);
    for (int i = 0; i < nold; ++i) {
      const cx<T> No = Nold[i][off];
      u.x += bold[i] * No.x;
      u.y += bold[i] * No.y;
    }
    if (has_L) {
      const T L = closed_L ? (kk * kk * Lfac) : Lbuf[off];
      const T den = T(1) - dt * L;
      u.x /= den;
      u.y /= den;
    }
    return u;
  }
};

// Fused pass: forward FFT of the two partially transformed fields along the last remaining
// axis, k-space update, inverse FFT of the updated field along the same axis.
template <class T> struct FusedIO {
  const cx<T> *inC, *inG;  // partial spectra of the variable and of the nonlinearity
  cx<T> *outU;             // partial spectrum of the updated variable (may alias inC)
  int n, ncols, nouter;
  long long pitch, outer_stride;
  int ncb;
  T scale;
  // multi-GPU slab layout (TMA kernel only): data staged as [nranks][nouter][nyl][ncols]
  int slab, nyl, nranks;
  // slab + peer stores: row y of the result goes to rank y / nyl, into its [nx][nyl][ncols]
  // array at x = peer_x0 + o (peer_tab: device array of nranks base pointers; null = staged store)
  const unsigned long long *peer_tab;
  int peer_x0;
  // slab == 2: blocked staging + bulk peer stores + optional arrival counters (mrl_passes_slab.cuh)
  int kzb_major, nx, rank;
  const unsigned long long *flag_wait;
  unsigned long long flag_expect;
  const unsigned long long *flag_tab;
  const void *ring_old;  // slab == 2: base the tensor map of the newest old nonlinear term is built on
  // slab == 1: the pass covers the x sub-range [o0, o0 + nouter) of arrays whose x extent is nouter_full (0: the whole
  // array); the caller shifts every base pointer by o0 * nyl * pitch elements
  int nouter_full;
};

template <class T, class C, int TK>
__global__ void __launch_bounds__(TK *C::TP) k_fused_fast(FusedIO<T> io, SpectralUpdate<T> up, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E;
  MRL_DYN_SMEM(smem_raw);
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *tw = buf + N * TK;
  const int tid = threadIdx.x;
  for (int i = tid; i < N; i += TK * TP) tw[i] = tw_g[i];
  __syncthreads();
  const int col = tid % TK, t = tid / TK;
  const SmTile<T, TK> sm{buf, col};
  const int ntiles = io.nouter * io.ncb;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int o = tile / io.ncb, c = (tile - o * io.ncb) * TK + col;
    const bool ok = c < io.ncols;
    const long long off = (long long)o * io.outer_stride + c;
    cx<T> a[E], g[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const long long q = off + (long long)(t + TP * e) * io.pitch;
      a[e] = ok ? io.inC[q] : mk<T>(T(0), T(0));
      g[e] = ok ? io.inG[q] : mk<T>(T(0), T(0));
    }
    RegFFT<T, C>::run(a, t, sm, tw);
    RegFFT<T, C>::run(g, t, sm, tw);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const int j = t + TP * e;
      if (ok) a[e] = conj(up.apply(o, j, c, off + (long long)j * io.pitch, a[e], g[e]));
    }
    RegFFT<T, C>::run(a, t, sm, tw);
    if (ok) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e)
        io.outU[off + (long long)(t + TP * e) * io.pitch] = mk<T>(a[e].x * io.scale, -a[e].y * io.scale);
    }
  }
}

template <class T, int TK>
__global__ void __launch_bounds__(256) k_fused_gen(FusedIO<T> io, SpectralUpdate<T> up, const cx<T> *tw, FFTPlanDev plan) {
  MRL_DYN_SMEM(smem_raw);
  const int n = plan.n;
  cx<T> *A = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *B = A + n * TK;
  cx<T> *G = B + n * TK;  // transformed nonlinearity
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ntiles = io.nouter * io.ncb;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int o = tile / io.ncb, c0 = (tile - o * io.ncb) * TK;
    const long long off = (long long)o * io.outer_stride + c0;
    // nonlinearity first, parked in G
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      A[w] = (c0 + c < io.ncols) ? io.inG[off + (long long)j * io.pitch + c] : mk<T>(T(0), T(0));
    }
    __syncthreads();
    const cx<T> *R = smem_fft<T, TK>(A, B, tw, plan, tid, nt);
    for (int w = tid; w < n * TK; w += nt) G[w] = R[w];
    __syncthreads();
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      A[w] = (c0 + c < io.ncols) ? io.inC[off + (long long)j * io.pitch + c] : mk<T>(T(0), T(0));
    }
    __syncthreads();
    R = smem_fft<T, TK>(A, B, tw, plan, tid, nt);
    cx<T> *D = (R == A) ? B : A;
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      cx<T> u = mk<T>(T(0), T(0));
      if (c0 + c < io.ncols) u = conj(up.apply(o, j, c0 + c, off + (long long)j * io.pitch + c, R[w], G[w]));
      D[w] = u;
    }
    __syncthreads();
    cx<T> *D2 = (D == A) ? B : A;
    R = smem_fft<T, TK>(D, D2, tw, plan, tid, nt);
    for (int w = tid; w < n * TK; w += nt) {
      const int j = w / TK, c = w % TK;
      if (c0 + c < io.ncols) {
        const cx<T> r = R[w];
        io.outU[off + (long long)j * io.pitch + c] = mk<T>(r.x * io.scale, -r.y * io.scale);
      }
    }
    __syncthreads();
  }
}

// ======================================================================== small pointwise kernels
#if defined(MRL_EMU)
template <class T> MRL_DI T mul_rn(T a, T b) { return a * b; }
template <class T> MRL_DI T add_rn(T a, T b) { return a + b; }
#else
MRL_DI double mul_rn(double a, double b) { return __dmul_rn(a, b); }
MRL_DI double add_rn(double a, double b) { return __dadd_rn(a, b); }
MRL_DI float mul_rn(float a, float b) { return __fmul_rn(a, b); }
MRL_DI float add_rn(float a, float b) { return __fadd_rn(a, b); }
#endif
// ReciprocalLaplacianFactor (kind 0: -k2*f) / ReciprocalLaplacianSquareFactor (kind 1: k2*k2*f)
template <class T>
__global__ void k_kfactor(T *out, const T *kx, const T *ky, const T *kz, int n0, int n1, int n2, int kind, T factor) {
  const long long total = (long long)n0 * n1 * n2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(i % n2);
    const int iy = (int)((i / n2) % n1);
    const int ix = (int)(i / ((long long)n1 * n2));
    // rounded exactly like the reference's separate libTorch ops (no FMA contraction):
    // k2 = (kx*kx + ky*ky) + kz*kz ; -k2*f ; (k2*k2)*f
    const T a = kx[ix], b = ky[iy], c = kz[iz];
    const T kk = add_rn(add_rn(mul_rn(a, a), mul_rn(b, b)), mul_rn(c, c));
    out[i] = kind == 0 ? mul_rn(-kk, factor) : mul_rn(mul_rn(kk, kk), factor);
  }
}

// Un-fused semi-implicit update on full spectra (generic path when the planner cannot fuse).
template <class T>
__global__ void k_ab_update(cx<T> *ubar, const cx<T> *cbar, const cx<T> *N, const T *L, T dt, T b0, int nold,
                            const cx<T> *o0, const cx<T> *o1, const cx<T> *o2, const cx<T> *o3, T c0, T c1, T c2,
                            T c3, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    cx<T> u = cbar[i];
    const cx<T> n = N[i];
    u.x += b0 * n.x;
    u.y += b0 * n.y;
    if (nold > 0) { u.x += c0 * o0[i].x; u.y += c0 * o0[i].y; }
    if (nold > 1) { u.x += c1 * o1[i].x; u.y += c1 * o1[i].y; }
    if (nold > 2) { u.x += c2 * o2[i].x; u.y += c2 * o2[i].y; }
    if (nold > 3) { u.x += c3 * o3[i].x; u.y += c3 * o3[i].y; }
    if (L) {
      const T den = T(1) - dt * L[i];
      u.x /= den;
      u.y /= den;
    }
    ubar[i] = u;
  }
}

// out = a * b with a real, b complex (e.g. Mbar * mubar)
template <class T> __global__ void k_mul_rc(cx<T> *out, const T *a, const cx<T> *b, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = mk<T>(a[i] * b[i].x, a[i] * b[i].y);
}

// Built-in real-space nonlinearity used until/unless a runtime expression is compiled:
// d/dc [ A (c-a)^2 (b-c)^2 ]  (examples/cahn_hilliard/cahnhilliard2.i: A=0.1,a=0,b=1;
// benchmarks/01_spinodal_decomposition/1a_solver.i: A=5,a=0.3,b=0.7)
template <class T> struct DoubleWellDeriv {
  T A, a, b;
  MRL_DI T operator()(T c, long long = 0) const {
    const T p = c - a, q = b - c;
    return T(2) * A * p * q * (q - p);
  }
};

template <class T, class F> __global__ void k_pointwise1(T *out, const T *in, F f, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = f(in[i]);
}

}  // namespace mrl
