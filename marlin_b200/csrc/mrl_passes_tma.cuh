// marlin_b200 - TMA-pipelined FFT passes for the large power-of-two sizes (sm_100a).
//
// Same arithmetic as the kernels in mrl_passes.cuh (RegFFT + the same load/store algebra), but
// every HBM read is an asynchronous bulk copy (cp.async.bulk / cp.async.bulk.tensor) into a
// shared-memory slot that is armed as soon as the previous occupant of the slot has been
// consumed, so the next tile is always in flight while the current one is being transformed.
// A CTA is split into NG independent groups (named barriers) working on different tiles, so
// one group's shared-memory exchange overlaps another group's FP64 butterflies; one CTA per SM.
// Results leave the SM as plain coalesced 128-byte-row stores.
//
//   k_strided_tma  - complex pass along a strided axis (P2 / P4 of the Cahn-Hilliard substep)
//   k_fused_tma    - P3: forward on two fields + semi-implicit update + inverse
//   k_zfwd_tma     - P1: last-axis r2c of (c + i F(c))
//   k_zinv_tma     - P5: last-axis c2r of row pairs
#pragma once
#include "mrl_passes.cuh"
#include "mrl_tma.cuh"

namespace mrl {

MRL_DI unsigned char *align128(unsigned char *p) {
  return (unsigned char *)(((unsigned long long)p + 127ull) & ~127ull);
}

// ======================================================================== strided pass
// Input through a 3-D tensor map over the real view [nouter][n][2*ncols] of the complex array
// (box = TK complex columns x min(n,256) rows); output by direct stores.
template <class T> struct StridedTmaIO {
  cx<T> *out;
  int n, ncols, nouter;  // nouter counts (field, outer) slices
  long long pitch, outer_stride;
  int ncb;
  T scale;
  int inverse;
};

template <class T, class C, int TK, int NG, int NS>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_strided_tma(const MRL_GRID_CONSTANT TensorMap tm, StridedTmaIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;  // complex elements per slot
  MRL_DYN_SMEM(smem_raw);
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  cx<T> *tw = slots + (size_t)NS * TILE;
  uint64_t *full = reinterpret_cast<uint64_t *>(tw + N);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  const int ntiles = io.nouter * io.ncb;
  const int nloc = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  auto issue = [&](int j) {  // one thread: arm slot j % NS with CTA-local tile j
    const int tile = blockIdx.x + j * gridDim.x;
    const int o = tile / io.ncb, cb = tile - o * io.ncb;
    const int s = j % NS;
    mbar_expect_tx(&full[s], (uint32_t)(TILE * sizeof(cx<T>)));
    MRL_UNROLL
    for (int b = 0; b < NBOX; ++b)
      tma_load_3d(slots + (size_t)s * TILE + b * BOXR * TK, &tm, &full[s], cb * TK * 2, b * BOXR, o);
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  for (int i = tid; i < N; i += NG * GT) tw[i] = tw_g[i];
  __syncthreads();
  if (tid == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  for (int j = g; j < nloc; j += NG) {
    const int s = j % NS;
    mbar_wait(&full[s], (uint32_t)((j / NS) & 1));
    const SmTile<T, TK> sm{slots + (size_t)s * TILE, col};
    const int tile = blockIdx.x + j * gridDim.x;
    const int o = tile / io.ncb, c = (tile - o * io.ncb) * TK + col;
    const bool ok = c < io.ncols;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      v[e] = sm.ld(t + TP * e);
      if (io.inverse) v[e].y = -v[e].y;
    }
    bar.sync();  // all inputs are in registers before the exchange overwrites the slot
    RegFFT<T, C>::run(v, t, sm, tw, bar, [&] {
      if (gt == 0 && j + NS < nloc) issue(j + NS);
    });
    if (ok) {
      cx<T> *dst = io.out + (long long)o * io.outer_stride + c;
      MRL_UNROLL
      for (int e = 0; e < E; ++e)
        dst[(long long)(t + TP * e) * io.pitch] = mk<T>(v[e].x * io.scale, (io.inverse ? -v[e].y : v[e].y) * io.scale);
    }
  }
}

// ======================================================================== fused P3
// Per group three dedicated slots: G (nonlinearity spectrum), C (variable spectrum), O (newest
// old nonlinear term, when the Adams-Bashforth order needs one).  Further old terms (order >= 3)
// are read with plain loads.
template <class T> struct FusedTmaIO {
  cx<T> *outU;
  int n, ncols, ncb;
  long long pitch;
  T scale;
};

template <class T> struct SpectralUpdate2 {
  const T *kx, *ky, *kz;
  int kmode, nzc, x0;
  int closed_M, closed_L, has_L;
  T Mfac, Lfac;
  const T *Mbuf, *Lbuf;
  T dt, b0;
  int nold;
  T bold0, bold1, bold2, bold3;
  const cx<T> *Nold1, *Nold2, *Nold3;  // old terms beyond the newest (which arrives through its slot)
  cx<T> *Nout;

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
  // chat, ghat: transformed variable / nonlinearity; nold0: newest old nonlinear term
  MRL_DI cx<T> apply(int o, int j, int col, long long off, cx<T> chat, cx<T> ghat, cx<T> nold0) const {
    const T kk = k2(o, j, col);
    const T M = closed_M ? (-kk * Mfac) : Mbuf[off];
    const cx<T> N = mk<T>(M * ghat.x, M * ghat.y);
    if (Nout) Nout[off] = N;
    cx<T> u = mk<T>(chat.x + b0 * N.x, chat.y + b0 * N.y);
    if (nold > 0) { u.x += bold0 * nold0.x; u.y += bold0 * nold0.y; }
    if (nold > 1) { const cx<T> q = Nold1[off]; u.x += bold1 * q.x; u.y += bold1 * q.y; }
    if (nold > 2) { const cx<T> q = Nold2[off]; u.x += bold2 * q.x; u.y += bold2 * q.y; }
    if (nold > 3) { const cx<T> q = Nold3[off]; u.x += bold3 * q.x; u.y += bold3 * q.y; }
    if (has_L) {
      const T L = closed_L ? (kk * kk * Lfac) : Lbuf[off];
      const T den = T(1) - dt * L;
      u.x /= den;
      u.y /= den;
    }
    return u;
  }
};

template <class T, class C, int TK, int NG>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_fused_tma(const MRL_GRID_CONSTANT TensorMap tmC, const MRL_GRID_CONSTANT TensorMap tmG,
                const MRL_GRID_CONSTANT TensorMap tmO, FusedTmaIO<T> io, SpectralUpdate2<T> up, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;
  MRL_DYN_SMEM(smem_raw);
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  cx<T> *tw = slots + (size_t)NG * 3 * TILE;
  uint64_t *full = reinterpret_cast<uint64_t *>(tw + N);  // [NG][3]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  const int ntiles = io.ncb;
  const int stride = gridDim.x * NG;
  const int first = blockIdx.x * NG + g;  // tiles of this group: first + j*stride
  const int nloc = (first < ntiles) ? (ntiles - first + stride - 1) / stride : 0;
  const bool use_old = up.nold > 0;
  cx<T> *sG = slots + (size_t)(g * 3 + 0) * TILE, *sC = slots + (size_t)(g * 3 + 1) * TILE, *sO = slots + (size_t)(g * 3 + 2) * TILE;
  uint64_t *bG = &full[g * 3 + 0], *bC = &full[g * 3 + 1], *bO = &full[g * 3 + 2];

  auto issue = [&](const TensorMap *tm, cx<T> *dst, uint64_t *b, int j) {
    const int cb = first + j * stride;
    mbar_expect_tx(b, (uint32_t)(TILE * sizeof(cx<T>)));
    MRL_UNROLL
    for (int q = 0; q < NBOX; ++q) tma_load_3d(dst + q * BOXR * TK, tm, b, cb * TK * 2, q * BOXR, 0);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * 3; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  for (int i = tid; i < N; i += NG * GT) tw[i] = tw_g[i];
  __syncthreads();
  if (gt == 0 && nloc > 0) {
    issue(&tmG, sG, bG, 0);
    issue(&tmC, sC, bC, 0);
    if (use_old) issue(&tmO, sO, bO, 0);
  }

  const GroupBarrier bar{1 + g, GT};
  for (int j = 0; j < nloc; ++j) {
    const uint32_t par = (uint32_t)(j & 1);
    const int c = (first + j * stride) * TK + col;
    const bool ok = c < io.ncols;
    const bool more = j + 1 < nloc;
    cx<T> a[E], gh[E];
    // ---- nonlinearity: forward transform, result stays in registers
    mbar_wait(bG, par);
    {
      const SmTile<T, TK> sm{sG, col};
      MRL_UNROLL
      for (int e = 0; e < E; ++e) gh[e] = sm.ld(t + TP * e);
      bar.sync();
      RegFFT<T, C>::run(gh, t, sm, tw, bar, [&] {
        if (gt == 0 && more) issue(&tmG, sG, bG, j + 1);
      });
    }
    // ---- variable: forward transform
    mbar_wait(bC, par);
    const SmTile<T, TK> smc{sC, col};
    MRL_UNROLL
    for (int e = 0; e < E; ++e) a[e] = smc.ld(t + TP * e);
    bar.sync();
    RegFFT<T, C>::run(a, t, smc, tw, bar);
    // ---- k-space update
    if (use_old) mbar_wait(bO, par);
    {
      const SmTile<T, TK> smo{sO, col};
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int jj = t + TP * e;
        const cx<T> no = use_old ? smo.ld(jj) : mk<T>(T(0), T(0));
        if (ok) a[e] = conj(up.apply(0, jj, c, (long long)jj * io.pitch + c, a[e], gh[e], no));
      }
    }
    if (use_old) {
      bar.sync_release();
      if (gt == 0 && more) issue(&tmO, sO, bO, j + 1);
    }
    // ---- inverse transform of the updated variable (exchange through the C slot)
    RegFFT<T, C>::run(a, t, smc, tw, bar, [&] {
      if (gt == 0 && more) issue(&tmC, sC, bC, j + 1);
    });
    if (ok) {
      cx<T> *dst = io.outU + c;
      MRL_UNROLL
      for (int e = 0; e < E; ++e) dst[(long long)(t + TP * e) * io.pitch] = mk<T>(a[e].x * io.scale, -a[e].y * io.scale);
    }
  }
}

// ======================================================================== P1: z r2c of (c + i F(c))
// A tile is PPB consecutive real rows (one contiguous bulk copy).  Each group owns NS input
// slots and one padded complex exchange buffer per pencil.
template <class T, class C, int PPB, int NG, int NS, class F>
__global__ void __launch_bounds__(NG *PPB *C::TP, 1)
    k_zfwd_tma(const T *cin, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows, F f, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  constexpr int NC = N / 2 + 1;
  MRL_DYN_SMEM(smem_raw);
  unsigned char *base = align128(smem_raw);
  T *slots = reinterpret_cast<T *>(base);                                       // [NG][NS][PPB*N] real
  cx<T> *xbuf = reinterpret_cast<cx<T> *>(slots + (size_t)NG * NS * PPB * N);   // [NG][PPB][NP]
  cx<T> *tw = xbuf + (size_t)NG * PPB * NP;
  uint64_t *full = reinterpret_cast<uint64_t *>(tw + N);                        // [NG][NS]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int t = gt % TP, pl = gt / TP;
  const long long ntiles = (nrows + PPB - 1) / PPB;
  const long long stride = (long long)gridDim.x * NG;
  const long long first = (long long)blockIdx.x * NG + g;
  const int nloc = (first < ntiles) ? (int)((ntiles - first + stride - 1) / stride) : 0;
  T *gs = slots + (size_t)g * NS * PPB * N;
  uint64_t *gb = full + g * NS;

  auto issue = [&](int j) {
    const long long row0 = (first + j * stride) * PPB;
    const long long left = nrows - row0;
    const int rows = left < PPB ? (int)left : PPB;
    const int s = j % NS;
    const uint32_t bytes = (uint32_t)(rows * N * sizeof(T));
    mbar_expect_tx(&gb[s], bytes);
    bulk_load_1d(gs + (size_t)s * PPB * N, cin + row0 * N, bytes, &gb[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  for (int i = tid; i < N; i += NG * GT) tw[i] = tw_g[i];
  __syncthreads();
  if (gt == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  for (int j = 0; j < nloc; ++j) {
    const int s = j % NS;
    const long long p = (first + j * stride) * PPB + pl;
    const bool ok = p < nrows;
    mbar_wait(&gb[s], (uint32_t)((j / NS) & 1));
    const T *src = gs + (size_t)s * PPB * N + pl * N;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const T a = ok ? src[t + TP * e] : T(0);
      const T b = f(a);
      if (mu_out && ok) mu_out[p * N + t + TP * e] = b;
      v[e] = mk<T>(a, b);
    }
    bar.sync_release();  // slot consumed; also orders the previous tile's separation reads
    if (gt == 0 && j + NS < nloc) issue(j + NS);
    RegFFT<T, C>::run(v, t, sm, tw, bar);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) sm.st(t + TP * e, v[e]);
    bar.sync();
    if (ok) {
      for (int k = t; k <= N / 2; k += TP) {
        cx<T> A, B;
        r2c_separate(sm.ld(k), sm.ld(k == 0 ? 0 : N - k), A, B);
        outC[p * NC + k] = A;
        outG[p * NC + k] = B;
      }
    }
  }
}

// ======================================================================== P5: z c2r of row pairs
// A tile is PPB pencils = 2*PPB consecutive half-spectrum rows (one contiguous bulk copy).
template <class T, class C, int PPB, int NG, int NS>
__global__ void __launch_bounds__(NG *PPB *C::TP, 1)
    k_zinv_tma(const cx<T> *in, T *out, long long nrows, T scale, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  constexpr int NC = N / 2 + 1;
  constexpr int SLOT = 2 * PPB * NC;  // complex elements per slot
  MRL_DYN_SMEM(smem_raw);
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));  // [NG][NS][SLOT]
  cx<T> *xbuf = slots + (size_t)NG * NS * SLOT;                  // [NG][PPB][NP]
  cx<T> *tw = xbuf + (size_t)NG * PPB * NP;
  uint64_t *full = reinterpret_cast<uint64_t *>(tw + N);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int t = gt % TP, pl = gt / TP;
  const long long npencils = (nrows + 1) / 2;
  const long long ntiles = (npencils + PPB - 1) / PPB;
  const long long stride = (long long)gridDim.x * NG;
  const long long first = (long long)blockIdx.x * NG + g;
  const int nloc = (first < ntiles) ? (int)((ntiles - first + stride - 1) / stride) : 0;
  cx<T> *gs = slots + (size_t)g * NS * SLOT;
  uint64_t *gb = full + g * NS;

  auto issue = [&](int j) {
    const long long row0 = (first + j * stride) * PPB * 2;
    const long long left = nrows - row0;
    const int rows = left < 2 * PPB ? (int)left : 2 * PPB;
    const int s = j % NS;
    const uint32_t bytes = (uint32_t)(rows * NC * sizeof(cx<T>));
    mbar_expect_tx(&gb[s], bytes);
    bulk_load_1d(gs + (size_t)s * SLOT, in + row0 * NC, bytes, &gb[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  for (int i = tid; i < N; i += NG * GT) tw[i] = tw_g[i];
  __syncthreads();
  if (gt == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  for (int j = 0; j < nloc; ++j) {
    const int s = j % NS;
    const long long p = (first + j * stride) * PPB + pl;
    const long long r0 = 2 * p;
    const bool ok = p < npencils, ok2 = r0 + 1 < nrows;
    mbar_wait(&gb[s], (uint32_t)((j / NS) & 1));
    const cx<T> *X = gs + (size_t)s * SLOT + (size_t)(2 * pl) * NC, *Y = X + NC;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const int idx = t + TP * e;
      const bool mir = idx > N / 2;
      const int k = mir ? N - idx : idx;
      const cx<T> x = ok ? X[k] : mk<T>(T(0), T(0));
      const cx<T> y = ok2 ? Y[k] : mk<T>(T(0), T(0));
      v[e] = c2r_merge_conj(x, y, k, N, mir);
    }
    bar.sync_release();
    if (gt == 0 && j + NS < nloc) issue(j + NS);
    RegFFT<T, C>::run(v, t, sm, tw, bar);
    if (ok) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int jj = t + TP * e;
        out[r0 * N + jj] = v[e].x * scale;
        if (ok2) out[(r0 + 1) * N + jj] = -v[e].y * scale;
      }
    }
  }
}

}  // namespace mrl
