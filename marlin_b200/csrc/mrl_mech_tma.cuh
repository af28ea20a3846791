// marlin_b200 - mechanics: first FFT pass with the tangent action fused into its load.
//
// One CG iteration of FFTMechanics (src/tensor_computes/FFTMechanics.C:107-112, conjugateGradientSolve in
// include/utils/MarlinUtils.h:83-127) applies  Ap = G( K4(F) : p ).  The pointwise part (p <- r + beta p, tmp = K4(F) : p)
// and the last-axis r2c of tmp were two kernels with a 9-component field written and re-read between them (18 S_r of
// the 110 S_r + 72 S_c an iteration moves).  Here a group of 9 pencils (one per tensor component) handles one PAIR of
// voxel rows: all its threads first evaluate the tangent for the 2 N voxels of the pair straight from F, K, mu, r, p
// (coalesced loads, the updated direction written back), leave the nine products in a shared-memory stage, and then run
// the nine packed real transforms (row 2q + i row 2q+1) exactly like k_zfwd_pairs_tma.  tmp never exists in HBM.
// Two groups per CTA work on different row pairs, so one group's loads overlap the other's transforms.
#pragma once
#include "mrl_mech_point.cuh"
#include "mrl_passes_tma.cuh"

namespace mrl {

template <class T, class C, int NG>
__global__ void __launch_bounds__(NG * 9 * C::TP, 1) k_mech_tangent_zfwd(MechTangentIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, PPB = 9, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  static_assert(E % 2 == 0 && GT % 32 == 0 && GT >= N, "whole warps per group, even points per thread, one thread per z");
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  unsigned char *base = align128(smem_raw);
  T *stage = reinterpret_cast<T *>(base);                                      // [NG][9][2][N] real
  cx<T> *xbuf = reinterpret_cast<cx<T> *>(stage + (size_t)NG * PPB * 2 * N);  // [NG][9][NP]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int pl = gt / TP;
  int plane;
  const int t = pair_map<TP>(gt % TP, tid & 31, plane);
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
  const long long npairs = io.nrows / 2;
  const long long stride = (long long)gridDim.x * NG;
  const bool upd = io.r != nullptr;
  const double beta = upd ? io.scal[4 /* SC_BETA */] : 0.0;
  T *gs = stage + (size_t)g * PPB * 2 * N;
  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  const long long n = io.n;
  for (long long q = (long long)blockIdx.x * NG + g; q < npairs; q += stride) {
    // ---- tangent: thread gt < N takes z = gt of both rows of the pair
    if (gt < N) {
      // one row at a time (not unrolled): 29 loads in flight per thread, half the registers of the two-row form
#if !defined(MRL_EMU)
#pragma unroll 1
#endif
      for (int row = 0; row < 2; ++row) {
        const long long v = (2 * q + row) * N + gt;
        MD<T, 3> Fm, X, R;
        T rv[9];
        MRL_UNROLL
        for (int c = 0; c < 9; ++c) Fm.a[c / 3][c % 3] = io.F[c * n + v];
        MRL_UNROLL
        for (int c = 0; c < 9; ++c) X.a[c / 3][c % 3] = io.p[c * n + v];
        if (upd) {
          MRL_UNROLL
          for (int c = 0; c < 9; ++c) rv[c] = io.r[c * n + v];
        }
        const T Kv = io.K[v], muv = io.mu[v];
        if (upd) {
          MRL_UNROLL
          for (int c = 0; c < 9; ++c) {
            const T pn = (T)((double)rv[c] + beta * (double)X.a[c / 3][c % 3]);  // same arithmetic as k_mech_pointwise mode 3
            X.a[c / 3][c % 3] = pn;
            io.p[c * n + v] = pn;
          }
        }
        mech_point<T, 3>(1, Fm, Kv, muv, X, R);
        MRL_UNROLL
        for (int c = 0; c < 9; ++c) gs[(size_t)(c * 2 + row) * N + gt] = R.a[c / 3][c % 3];
      }
    }
    bar.sync();
    // ---- nine packed real transforms (the stage is free again once every thread has passed the first exchange barrier)
    const T *src = gs + (size_t)(pl * 2) * N;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = mk<T>(src[t + TP * e], src[N + t + TP * e]);
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, NoHook());
    cx<T> w[E / 2];
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) w[e] = shfl_cx(v[E - 1 - e], plane);
    if (t == 0) {
      w[0] = v[0];
      MRL_UNROLL
      for (int e = 1; e < E / 2; ++e) w[e] = v[E - e];
    }
    cx<T> *oa = io.out + ((long long)pl * io.nrows + 2 * q) * io.ncp + t, *ob = oa + io.ncp;
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) {
      cx<T> A, B;
      r2c_separate(v[e], w[e], A, B);
      oa[TP * e] = A;
      ob[TP * e] = B;
    }
    if (t == 0) {
      cx<T> A, B;
      r2c_separate(v[E / 2], v[E / 2], A, B);
      oa[N / 2] = A;
      ob[N / 2] = B;
    }
    if (nofft) bar.sync();  // development switch: no exchange barrier separates the stage's readers from its next writers
  }
}

// The same pass with the input rows staged by asynchronous bulk copies.  The load-to-register form above waits on its 29
// loads per voxel (ncu: 10 long-scoreboard stall cycles per issue, 5.3 TB/s); here one CTA = one group of nine pencils, the
// rows of r, p, F, K, mu of ONE voxel row (29 x N values) travel into one of two shared-memory stages with cp.async.bulk
// while the other stage is being consumed: row 2q+1 is in flight while the tangent of row 2q is evaluated, and both rows
// of the next pair while the nine transforms of this pair run.
template <class T, class C>
__global__ void __launch_bounds__(9 * C::TP, 1) k_mech_tangent_zfwd_tma(MechTangentIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, PPB = 9, GT = PPB * TP, NIN = 29;
  constexpr int NP = N + (N >> 3) + 1;
  static_assert(E % 2 == 0 && GT % 32 == 0 && GT >= N, "whole warps, even points per thread, one thread per z");
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  unsigned char *base = align128(smem_raw);
  T *in = reinterpret_cast<T *>(base);                                    // [2 stages][29][N]: r(9) p(9) F(9) K mu
  T *prod = in + (size_t)2 * NIN * N;                                     // [9][2][N]
  cx<T> *xbuf = reinterpret_cast<cx<T> *>(prod + (size_t)PPB * 2 * N);    // [9][NP]
  uint64_t *full = reinterpret_cast<uint64_t *>(xbuf + (size_t)PPB * NP);  // [2]
  const int gt = threadIdx.x;
  const int pl = gt / TP;
  int plane;
  const int t = pair_map<TP>(gt % TP, gt & 31, plane);
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
  const long long npairs = io.nrows / 2, n = io.n;
  const long long first = blockIdx.x, stride = gridDim.x;
  const int nloc = first < npairs ? (int)((npairs - first + stride - 1) / stride) : 0;
  const bool upd = io.r != nullptr;
  const double beta = upd ? io.scal[4 /* SC_BETA */] : 0.0;
  const GroupBarrier bar{1, GT};
  const SmPencil<T> sm{xbuf + (size_t)pl * NP};

  // stage s <- the 29 (20 without a direction update) rows of voxel row `row`
  auto issue = [&](int s, long long row) {
    T *dst = in + (size_t)s * NIN * N;
    const long long v0 = row * N;
    const uint32_t rb = (uint32_t)(N * sizeof(T));
    mbar_expect_tx(&full[s], rb * (upd ? NIN : NIN - 9));
    MRL_UNROLL
    for (int c = 0; c < 9; ++c) {
      if (upd) bulk_load_1d(dst + (size_t)c * N, io.r + c * n + v0, rb, &full[s]);
      bulk_load_1d(dst + (size_t)(9 + c) * N, io.p + c * n + v0, rb, &full[s]);
      bulk_load_1d(dst + (size_t)(18 + c) * N, io.F + c * n + v0, rb, &full[s]);
    }
    bulk_load_1d(dst + (size_t)27 * N, io.K + v0, rb, &full[s]);
    bulk_load_1d(dst + (size_t)28 * N, io.mu + v0, rb, &full[s]);
  };

  if (gt == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (gt == 0 && nloc > 0) {
    issue(0, 2 * first);
    issue(1, 2 * first + 1);
  }
  for (int i = 0; i < nloc; ++i) {
    const long long q = first + i * stride;
    const uint32_t par = (uint32_t)(i & 1);
    // ---- tangent, one voxel row per stage
    MRL_UNROLL
    for (int row = 0; row < 2; ++row) {
      mbar_wait(&full[row], par);
      if (gt < N) {
        const T *src = in + (size_t)row * NIN * N + gt;
        const long long v = (2 * q + row) * N + gt;
        MD<T, 3> Fm, X, R;
        MRL_UNROLL
        for (int c = 0; c < 9; ++c) {
          Fm.a[c / 3][c % 3] = src[(18 + c) * N];
          T pv = src[(9 + c) * N];
          if (upd) {
            pv = (T)((double)src[c * N] + beta * (double)pv);  // same arithmetic as k_mech_pointwise mode 3
            io.p[c * n + v] = pv;
          }
          X.a[c / 3][c % 3] = pv;
        }
        mech_point<T, 3>(1, Fm, src[27 * N], src[28 * N], X, R);
        MRL_UNROLL
        for (int c = 0; c < 9; ++c) prod[(size_t)(c * 2 + row) * N + gt] = R.a[c / 3][c % 3];
      }
      bar.sync_release();  // the stage has been read by everyone: re-arm it with the same row of the next pair
      if (gt == 0 && i + 1 < nloc) issue(row, 2 * (q + stride) + row);
    }
    // ---- nine packed real transforms of the products
    const T *src = prod + (size_t)(pl * 2) * N;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = mk<T>(src[t + TP * e], src[N + t + TP * e]);
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, NoHook());
    cx<T> w[E / 2];
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) w[e] = shfl_cx(v[E - 1 - e], plane);
    if (t == 0) {
      w[0] = v[0];
      MRL_UNROLL
      for (int e = 1; e < E / 2; ++e) w[e] = v[E - e];
    }
    cx<T> *oa = io.out + ((long long)pl * io.nrows + 2 * q) * io.ncp + t, *ob = oa + io.ncp;
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) {
      cx<T> A, B;
      r2c_separate(v[e], w[e], A, B);
      oa[TP * e] = A;
      ob[TP * e] = B;
    }
    if (t == 0) {
      cx<T> A, B;
      r2c_separate(v[E / 2], v[E / 2], A, B);
      oa[N / 2] = A;
      ob[N / 2] = B;
    }
    if (nofft) bar.sync();
  }
}

}  // namespace mrl
