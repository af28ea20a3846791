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

// Development switch (MRL_DEBUG_NOFFT=1): skip the butterflies and exchanges so that a pass
// degenerates to "bulk-load tile, store tile" - measures the memory-side ceiling of its access
// pattern.  Results are wrong by construction; never set outside tools/.
#if defined(MRL_EMU)
static int g_debug_nofft = 0;
#else
__device__ int g_debug_nofft = 0;
#endif
template <class T, class C, class V, class SM, class TW, class BAR, class HOOK>
MRL_DI void fft_or_skip(bool nofft, V &v, int t, const SM &sm, const TW &tw, const BAR &bar, const HOOK &hook) {
  // nofft: g_debug_nofft read ONCE per kernel into a register (a load + dependent branch per
  // transform showed up as 8 % of the stall samples of the fused passes)
  if (nofft) {
    bar.sync_release();
    hook();
  } else {
    RegFFT<T, C>::run_tw(v, t, sm, tw, bar, hook);
  }
}

// Round a shared-memory pointer up to 128 bytes with pointer arithmetic only, so the compiler
// keeps the shared address space (LDS/STS instead of generic LD/ST).
MRL_DI unsigned char *align128(unsigned char *p) { return p + ((128u - (smem_u32(p) & 127u)) & 127u); }

// Thread -> pencil-thread index t such that the threads holding t and (TP - t) % TP sit in the
// same warp (needed to split/merge Hermitian pairs with warp shuffles).  wl: linear index of the
// thread within its pencil's TP threads; returns t and the partner's lane.
template <int TP> MRL_DI int pair_map(int wl, int lane, int &partner_lane) {
  if constexpr (TP <= 32) {
    partner_lane = lane - wl + (TP - wl) % TP;
    return wl;
  } else {
    const int l = wl & 31, q = (wl >> 5) * 16 + (l & 15);
    partner_lane = q == 0 ? lane : (lane ^ 16);
    return l < 16 ? q : (q == 0 ? TP / 2 : TP - q);
  }
}

// ======================================================================== strided pass
// Input through a 3-D tensor map over the real view [nouter][n][2*ncols] of the complex array
// (box = TK complex columns x min(n,256) rows); output by direct stores.
template <class T> struct StridedTmaIO {
  cx<T> *out, *out1;     // output base of field 0 / field 1
  int nouter_f;          // outer slices per field
  int n, ncols, nouter;  // nouter counts (field, outer) slices
  int nvalid;            // columns >= nvalid of a slice are padding: transformed but never stored
  const unsigned long long *peer_tab;  // multi-GPU: scatter result rows to the peers (see StridedIO)
  int peer_rows;
  long long peer_field, peer_off;
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
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)NS * TILE);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
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
    const bool ok = c < io.nvalid;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      v[e] = sm.ld(t + TP * e);
      if (io.inverse) v[e].y = -v[e].y;
    }
    bar.sync();  // all inputs are in registers before the exchange overwrites the slot
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, [&] {
      if (gt == 0 && j + NS < nloc) issue(j + NS);
    });
    if (ok && io.peer_tab) {
      // fused all-to-all: each row lands directly in the HBM of the rank that owns its x block
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int r = t + TP * e, s = r / io.peer_rows;
        cx<T> *dst = reinterpret_cast<cx<T> *>(io.peer_tab[s]) + (long long)o * io.peer_field + io.peer_off +
                     (long long)(r - s * io.peer_rows) * io.pitch + c;
        *dst = mk<T>(v[e].x * io.scale, (io.inverse ? -v[e].y : v[e].y) * io.scale);
      }
    } else if (ok) {
      cx<T> *dst = (o < io.nouter_f ? io.out + (long long)o * io.outer_stride
                                    : io.out1 + (long long)(o - io.nouter_f) * io.outer_stride) + c;
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
// Plain layout: [n][ncols] (one outer slice).  Slab layout (multi-GPU, reciprocal space split
// along x): the transform axis y arrives as P blocks of nyl rows received from the P ranks,
// stored [P][nouter = nxl][nyl][ncols = nzc]; the 4-D tensor map delivers the rows of one
// (x, column block) in y order, and the result goes back to the same staged layout so that
// block s is contiguous for the return all-to-all.
template <class T> struct FusedTmaIO {
  cx<T> *outU;
  int n, ncols, ncb;
  long long pitch;
  T scale;
  int slab, nouter, nyl;
  int nouter_full;                     // x extent of the staged arrays (= nouter unless the pass covers an x sub-range)
  const unsigned long long *peer_tab;  // slab: store the result rows into the owning ranks' arrays
  int peer_x0;
  // slab == 2 (mrl_passes_slab.cuh): variable / nonlinearity tiles through 5-D maps over the blocked staging R, result rows
  // pushed with bulk copies into the peers' blocked return staging S; the ring of old nonlinear terms keeps the
  // [nranks][nouter][nyl][ncols] layout of slab == 1
  int kzb_major;                        // tile order: column block slowest
  int nx, rank, nranks;
  const unsigned long long *flag_wait;  // this rank's arrival counters [source][ncb] of the forward exchange (null: none)
  unsigned long long flag_expect;
  const unsigned long long *flag_tab;   // bases of the peers' arrival counters of the return exchange
  MRL_DI long long row_off(int o, int row) const {
    if (!slab) return (long long)row * pitch;
    const int s = row / nyl, yl = row - s * nyl;
    return (((long long)s * nouter_full + o) * nyl + yl) * pitch;
  }
};

// 1/d to within an ulp without the division slow path: hardware seed + two Newton steps.
// (1 - dt*L >= 1 for the dissipative linear operators this is used with; never 0, Inf or NaN.)
#if defined(MRL_EMU)
template <class T> MRL_DI T fast_rcp(T d) { return T(1) / d; }
#else
MRL_DI double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}
MRL_DI float fast_rcp(float d) { return __frcp_rn(d); }
#endif

template <class T> struct SpectralUpdate2 {
  const T *kx, *ky, *kz;
  int kmode, nzc, x0;
  int nzv;  // valid entries of the last axis (nzc may be a padded pitch)
  int closed_M, closed_L, has_L;
  T Mfac, Lfac;
  const T *Mbuf, *Lbuf;
  T dt, b0;
  int nold;
  T bold0, bold1, bold2, bold3;
  const cx<T> *Nold1, *Nold2, *Nold3;  // old terms beyond the newest (which arrives through its slot)
  cx<T> *Nout;

  // k^2 = rowterm(j) + colterm: the part that is fixed for a thread's column during one tile is
  // evaluated once per tile, the part that follows the transform axis once per element.
  MRL_DI T colterm(int o, int col) const {
    if (kmode == MRL_KMODE_3D) {
      const T b = ky[col / nzc], c = kz[col % nzc];
      return b * b + c * c;
    } else if (kmode == MRL_KMODE_3D_SLAB) {
      const T a = kx[x0 + o], c = kz[col];
      return a * a + c * c;
    }
    const T b = ky[col];
    return b * b;
  }
  MRL_DI const T *rowaxis() const { return kmode == MRL_KMODE_3D_SLAB ? ky : kx; }
  // chat, ghat: transformed variable / nonlinearity; nold0: newest old nonlinear term.
  // off: element offset in the work / ring arrays (row pitch nzc, possibly padded); moff: offset of the
  // same wavevector in the caller's mobility / linear-operator buffers (natural layout, row length nzv)
  MRL_DI cx<T> apply(T kr, T kcol, long long off, long long moff, cx<T> chat, cx<T> ghat, cx<T> nold0) const {
    const T kk = kr * kr + kcol;
    const T M = closed_M == 1 ? (-kk * Mfac) : closed_M == 2 ? T(1) : Mbuf[moff];
    const cx<T> N = mk<T>(M * ghat.x, M * ghat.y);
    if (Nout) Nout[off] = N;
    cx<T> u = mk<T>(chat.x + b0 * N.x, chat.y + b0 * N.y);
    if (nold > 0) { u.x += bold0 * nold0.x; u.y += bold0 * nold0.y; }
    if (nold > 1) { const cx<T> q = Nold1[off]; u.x += bold1 * q.x; u.y += bold1 * q.y; }
    if (nold > 2) { const cx<T> q = Nold2[off]; u.x += bold2 * q.x; u.y += bold2 * q.y; }
    if (nold > 3) { const cx<T> q = Nold3[off]; u.x += bold3 * q.x; u.y += bold3 * q.y; }
    if (has_L) {
      const T L = closed_L ? (kk * kk * Lfac) : Lbuf[moff];
      const T r = fast_rcp(T(1) - dt * L);
      u.x *= r;
      u.y *= r;
    }
    return u;
  }
};

// SLAB: 0 = one outer slice in the plain layout, 1 = slab staging with per-thread stores, 2 = blocked staging with bulk
// peer stores (compile-time so that the single-GPU instantiation carries none of the exchange code)
template <class T, class C, int TK, int NG, int SLAB = 0>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_fused_tma(const MRL_GRID_CONSTANT TensorMap tmC, const MRL_GRID_CONSTANT TensorMap tmG,
                const MRL_GRID_CONSTANT TensorMap tmO, FusedTmaIO<T> io, SpectralUpdate2<T> up, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)NG * 3 * TILE);  // [NG][3]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  TwRegs<T, C, true> twr;
  twr.init(tw_g, t);
  const int ntiles = io.nouter * io.ncb;
  const int stride = gridDim.x * NG;
  const int first = blockIdx.x * NG + g;  // tiles of this group: first + j*stride
  const int nloc = (first < ntiles) ? (ntiles - first + stride - 1) / stride : 0;
  const bool use_old = up.nold > 0;
  cx<T> *sG = slots + (size_t)(g * 3 + 0) * TILE, *sC = slots + (size_t)(g * 3 + 1) * TILE, *sO = slots + (size_t)(g * 3 + 2) * TILE;
  uint64_t *bG = &full[g * 3 + 0], *bC = &full[g * 3 + 1], *bO = &full[g * 3 + 2];

  auto decode = [&](int tile, int &o, int &cb) {
    if (SLAB == 2 && io.kzb_major) {
      cb = tile / io.nouter;
      o = tile - cb * io.nouter;
    } else {
      o = tile / io.ncb;
      cb = tile - o * io.ncb;
    }
  };
  // blocked: the tile comes from the blocked staging R (5-D map); waitf: first load of a tile, wait for its arrivals
  auto issue = [&](const TensorMap *tm, cx<T> *dst, uint64_t *b, int j, bool blocked = false, bool waitf = false) {
    int o, cb;
    decode(first + j * stride, o, cb);
    if constexpr (SLAB == 2) {
      if (waitf && io.flag_wait) {
        for (int q = 0; q < io.nranks; ++q) wait_counter(io.flag_wait + (size_t)q * io.ncb + cb, io.flag_expect);
        fence_proxy_async();
      }
    }
    mbar_expect_tx(b, (uint32_t)(TILE * sizeof(cx<T>)));
    if (SLAB == 2 && blocked) {
      tma_load_5d(dst, tm, b, 0, o, 0, cb, 0);
    } else if (SLAB != 0) {
      tma_load_4d(dst, tm, b, cb * TK * 2, 0, o, 0);
    } else {
      MRL_UNROLL
      for (int q = 0; q < NBOX; ++q) tma_load_3d(dst + q * BOXR * TK, tm, b, cb * TK * 2, q * BOXR, 0);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < NG * 3; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  constexpr bool blk = SLAB == 2;
  if (gt == 0 && nloc > 0) {
    issue(&tmG, sG, bG, 0, blk, true);
    issue(&tmC, sC, bC, 0, blk);
    if (use_old) issue(&tmO, sO, bO, 0);
  }
  auto signal = [&](int cb) {  // this thread's bulk writes of a tile of column block cb have completed
    fence_proxy_async_global();
    for (int i = 0; i < io.nranks; ++i) {
      const int d = (io.rank + 1 + i) % io.nranks;
      red_release_sys_add(reinterpret_cast<unsigned long long *>(io.flag_tab[d]) + (size_t)io.rank * io.ncb + cb, 1ull);
    }
  };
  int prev_cb = -1;

  const GroupBarrier bar{1 + g, GT};
  for (int j = 0; j < nloc; ++j) {
    const uint32_t par = (uint32_t)(j & 1);
    int o, cbt;
    decode(first + j * stride, o, cbt);
    const int c = cbt * TK + col;
    const bool ok = c < io.ncols && (up.kmode == MRL_KMODE_2D || (c % up.nzc) < up.nzv);
    const bool more = j + 1 < nloc;
    cx<T> a[E], gh[E];
    // ---- nonlinearity: forward transform, result stays in registers
    mbar_wait(bG, par);
    {
      const SmTile<T, TK> sm{sG, col};
      MRL_UNROLL
      for (int e = 0; e < E; ++e) gh[e] = sm.ld(t + TP * e);
      bar.sync();
      fft_or_skip<T, C>(nofft, gh, t, sm, twr, bar, [&] {
        if (gt == 0 && more) issue(&tmG, sG, bG, j + 1, blk, true);
      });
    }
    // ---- variable: forward transform; its slot is re-armed as soon as the transform has left it
    mbar_wait(bC, par);
    {
      const SmTile<T, TK> smc{sC, col};
      MRL_UNROLL
      for (int e = 0; e < E; ++e) a[e] = smc.ld(t + TP * e);
      bar.sync();
      fft_or_skip<T, C>(nofft, a, t, smc, twr, bar, [&] {
        if (gt == 0 && more) issue(&tmC, sC, bC, j + 1, blk);
      });
    }
    // ---- k-space update (newest old nonlinear term from its slot)
    if (use_old) mbar_wait(bO, par);
    const SmTile<T, TK> smo{sO, col};
    if (ok) {
      const T kcol = up.colterm(o, c);
      const T *kr = up.rowaxis();
      // the caller's M / L buffers are not padded: [n][ncols / nzc][nzv]
      const long long mrow = (long long)(io.ncols / up.nzc) * up.nzv, mcol = (long long)(c / up.nzc) * up.nzv + (c % up.nzc);
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int jj = t + TP * e;
        const cx<T> no = use_old ? smo.ld(jj) : mk<T>(T(0), T(0));
        a[e] = conj(up.apply(kr[jj], kcol, io.row_off(o, jj) + c, jj * mrow + mcol, a[e], gh[e], no));
      }
    }
    bar.sync();  // every thread has taken its old-term values: the O slot becomes the exchange buffer
    // ---- inverse transform of the updated variable (exchange through the O slot)
    fft_or_skip<T, C>(nofft, a, t, smo, twr, bar, [&] {
      if (gt == 0 && more && use_old && !blk) issue(&tmO, sO, bO, j + 1);
    });
    if constexpr (blk) {
      // fused return all-to-all, bulk form: the O slot (its last reader has passed the exchange) stages the result
      // tile; the nyl rows that belong to rank d are contiguous there and in d's blocked staging S = [ncb][nx][nyl][W]
      MRL_UNROLL
      for (int e = 0; e < E; ++e) smo.st(t + TP * e, mk<T>(a[e].x * io.scale, -a[e].y * io.scale));
      bar.sync_release();
      if (gt == 0) {
        const size_t chunk = (size_t)io.nyl * TK;
        for (int i = 0; i < io.nranks; ++i) {
          const int d = (io.rank + 1 + i) % io.nranks;
          cx<T> *dst = reinterpret_cast<cx<T> *>(io.peer_tab[d]) + ((long long)cbt * io.nx + io.peer_x0 + o) * (long long)chunk;
          bulk_store_1d(dst, sO + (size_t)d * chunk, (uint32_t)(chunk * sizeof(cx<T>)));
        }
        bulk_commit();
        if (io.flag_tab) {
          if (prev_cb >= 0) {
            bulk_wait_done<1>();
            signal(prev_cb);
          }
          prev_cb = cbt;
        }
        bulk_wait_read_all();  // the staging tile has been read: the slot may take the next old-term tile
        if (more && use_old) issue(&tmO, sO, bO, j + 1);
      }
    } else if (SLAB == 1 && ok && io.peer_tab) {
      // fused return all-to-all: row y belongs to rank y / nyl, at x = peer_x0 + o of its slab
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int y = t + TP * e, s = y / io.nyl;
        cx<T> *dst = reinterpret_cast<cx<T> *>(io.peer_tab[s]) + ((long long)(io.peer_x0 + o) * io.nyl + (y - s * io.nyl)) * io.pitch + c;
        *dst = mk<T>(a[e].x * io.scale, -a[e].y * io.scale);
      }
    } else if (ok) {
      cx<T> *dst = io.outU + c;
      MRL_UNROLL
      for (int e = 0; e < E; ++e) dst[io.row_off(o, t + TP * e)] = mk<T>(a[e].x * io.scale, -a[e].y * io.scale);
    }
  }
  if constexpr (blk) if (gt == 0) {
    bulk_wait_done<0>();  // the kernel's completion implies the peers hold the rows
    if (io.flag_tab && prev_cb >= 0) signal(prev_cb);
    fence_proxy_async_global();
  }
}

// ======================================================================== mechanics: fused Green projection
// de Geus projection  A_ij <- (sum_l A_il q_l) q_j / |q|^2  (FFTMechanics.C:48-85,104-105;
// Ghat4_ijlm = delta_im q_j q_l / |q|^2, zero at q = 0) fused between the forward and the inverse
// transform along the axis that is transformed last (x).  The contraction couples only the three
// components of one tensor row i, so a tile is (row i, TK columns): three forward transforms
// accumulate s = sum_l q_l A_il in registers, then three inverse transforms of q_j s / |q|^2 go
// back to the same locations.  9 S_c read + 9 S_c written instead of the 36 + 36 of
// x-forward / projection kernel / x-inverse.
// Layout [9][n][ncols] (ncols = n1 * ncp), one 3-D tensor map over all nine components; in place.
// Slot schedule per group: slots 0 and 1 are re-armed with the next tile right after their forward
// transform; slot 2 serves as exchange buffer of the three inverse transforms and is re-armed
// after the last one.
template <class T> struct MechFusedTmaIO {
  cx<T> *out;
  int n, ncols, ncb;
  long long pitch, field;
  const T *kx, *ky, *kz;
  int ncp, nzv;
  T scale;
};

template <class T, class C, int TK, int NG>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_mech_fused_tma(const MRL_GRID_CONSTANT TensorMap tm, MechFusedTmaIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)NG * 3 * TILE);  // [NG][3]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  TwRegs<T, C, true> twr;
  twr.init(tw_g, t);
  const int ntiles = 3 * io.ncb;
  const int stride = gridDim.x * NG;
  const int first = blockIdx.x * NG + g;
  const int nloc = (first < ntiles) ? (ntiles - first + stride - 1) / stride : 0;
  cx<T> *sl = slots + (size_t)(g * 3) * TILE;
  uint64_t *bl = &full[g * 3];

  auto issue = [&](int l, int j) {
    const int tile = first + j * stride;
    const int i = tile / io.ncb, cb = tile - i * io.ncb;
    mbar_expect_tx(&bl[l], (uint32_t)(TILE * sizeof(cx<T>)));
    MRL_UNROLL
    for (int q = 0; q < NBOX; ++q) tma_load_3d(sl + (size_t)l * TILE + q * BOXR * TK, &tm, &bl[l], cb * TK * 2, q * BOXR, 3 * i + l);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * 3; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (gt == 0 && nloc > 0) {
    issue(0, 0);
    issue(1, 0);
    issue(2, 0);
  }

  const GroupBarrier bar{1 + g, GT};
  const SmTile<T, TK> sm0{sl, col}, sm1{sl + TILE, col}, sm2{sl + 2 * (size_t)TILE, col};
  for (int j = 0; j < nloc; ++j) {
    const uint32_t par = (uint32_t)(j & 1);
    const int tile = first + j * stride;
    const int i = tile / io.ncb;
    const int c = (tile - i * io.ncb) * TK + col;
    const bool ok = c < io.ncols && (c % io.ncp) < io.nzv;
    const bool more = j + 1 < nloc;
    const T qy = ok ? io.ky[c / io.ncp] : T(0), qz = ok ? io.kz[c % io.ncp] : T(0);
    cx<T> a[E], s[E];
    // ---- forward transforms of A_i0, A_i1, A_i2; s accumulates sum_l q_l A_il
    mbar_wait(&bl[0], par);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) a[e] = sm0.ld(t + TP * e);
    bar.sync();
    fft_or_skip<T, C>(nofft, a, t, sm0, twr, bar, [&] {
      if (gt == 0 && more) issue(0, j + 1);
    });
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const T qx = io.kx[t + TP * e];
      s[e] = mk<T>(a[e].x * qx, a[e].y * qx);
    }
    mbar_wait(&bl[1], par);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) a[e] = sm1.ld(t + TP * e);
    bar.sync();
    fft_or_skip<T, C>(nofft, a, t, sm1, twr, bar, [&] {
      if (gt == 0 && more) issue(1, j + 1);
    });
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      s[e].x += a[e].x * qy;
      s[e].y += a[e].y * qy;
    }
    mbar_wait(&bl[2], par);
    MRL_UNROLL
    for (int e = 0; e < E; ++e) a[e] = sm2.ld(t + TP * e);
    bar.sync();
    fft_or_skip<T, C>(nofft, a, t, sm2, twr, bar, NoHook());
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const T qx = io.kx[t + TP * e];
      const T Q = qx * qx + qy * qy + qz * qz;
      const T inv = Q == T(0) ? T(0) : fast_rcp(Q);
      s[e].x = (s[e].x + a[e].x * qz) * inv;
      s[e].y = (s[e].y + a[e].y * qz) * inv;
    }
    // ---- inverse transforms of q_j s (conj . FFT . conj), exchange through slot 2
    MRL_UNROLL
    for (int jj = 0; jj < 3; ++jj) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const T q = jj == 0 ? io.kx[t + TP * e] : jj == 1 ? qy : qz;
        a[e] = mk<T>(s[e].x * q, -s[e].y * q);
      }
      fft_or_skip<T, C>(nofft, a, t, sm2, twr, bar, [&] {
        if (jj == 2 && gt == 0 && more) issue(2, j + 1);
      });
      if (ok) {
        cx<T> *dst = io.out + (long long)(3 * i + jj) * io.field + c;
        MRL_UNROLL
        for (int e = 0; e < E; ++e) dst[(long long)(t + TP * e) * io.pitch] = mk<T>(a[e].x * io.scale, -a[e].y * io.scale);
      }
    }
  }
}

// ======================================================================== P1: z r2c of (c + i F(c))
// A tile is PPB consecutive real rows (one contiguous bulk copy).  Each group owns NS input
// slots and one padded complex exchange buffer per pencil.  The Hermitian split
// C_k = (Z_k + conj Z_{N-k})/2, G_k = (Z_k - conj Z_{N-k})/2i needs Z_{N-k}, which lives in the
// thread holding pencil index TP - t: pair_map() puts both in one warp and the values move with
// warp shuffles, so the split costs no shared-memory traffic and no barrier.
template <class T> MRL_DI cx<T> shfl_cx(cx<T> v, int lane) {
  return mk<T>(__shfl_sync(0xffffffffu, v.x, lane), __shfl_sync(0xffffffffu, v.y, lane));
}

template <class T, class C, int PPB, int NG, int NS, class F>
__global__ void __launch_bounds__(NG *PPB *C::TP, 1)
    k_zfwd_tma(const T *cin, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows, int ncp, F f, const cx<T> *tw_g, RowMap rm) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  // ncp: row pitch of the half spectra (>= N/2+1; padded to 128 bytes inside the split plan)
  static_assert(E % 2 == 0, "Hermitian split by shuffles needs an even number of points per thread");
  static_assert(GT % 32 == 0, "a group must be whole warps (full-mask shuffles inside the group loop)");
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  unsigned char *base = align128(smem_raw);
  T *slots = reinterpret_cast<T *>(base);                                       // [NG][NS][PPB*N] real
  cx<T> *xbuf = reinterpret_cast<cx<T> *>(slots + (size_t)NG * NS * PPB * N);   // [NG][PPB][NP]
  uint64_t *full = reinterpret_cast<uint64_t *>(xbuf + (size_t)NG * PPB * NP);  // [NG][NS]
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int pl = gt / TP;
  int plane;
  const int t = pair_map<TP>(gt % TP, tid & 31, plane);
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
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
    bulk_load_1d(gs + (size_t)s * PPB * N, cin + rm(row0) * N, bytes, &gb[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (gt == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  for (int j = 0; j < nloc; ++j) {
    const int s = j % NS;
    const long long pc = (first + j * stride) * PPB + pl;  // row of the (chunk) enumeration
    const bool ok = pc < nrows;
    const long long p = rm(pc);                            // row of the full arrays
    mbar_wait(&gb[s], (uint32_t)((j / NS) & 1));
    const T *src = gs + (size_t)s * PPB * N + pl * N;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) {
      const T a = ok ? src[t + TP * e] : T(0);
      const T b = ok ? f(a, p * N + t + TP * e) : T(0);
      if (mu_out && ok) mu_out[p * N + t + TP * e] = b;
      v[e] = mk<T>(a, b);
    }
    bar.sync_release();  // slot consumed by the whole group: re-arm it
    if (gt == 0 && j + NS < nloc) issue(j + NS);
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, NoHook());
    // Hermitian split: partner element of (t, e) is (TP - t, E-1-e); thread 0 pairs (0, e) with (0, E-e)
    cx<T> w[E / 2];
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) w[e] = shfl_cx(v[E - 1 - e], plane);
    if (t == 0) {
      w[0] = v[0];
      MRL_UNROLL
      for (int e = 1; e < E / 2; ++e) w[e] = v[E - e];
    }
    if (ok) {
      cx<T> *oc = outC + p * ncp + t, *og = outG + p * ncp + t;
      MRL_UNROLL
      for (int e = 0; e < E / 2; ++e) {
        cx<T> A, B;
        r2c_separate(v[e], w[e], A, B);
        oc[TP * e] = A;
        og[TP * e] = B;
      }
      if (t == 0) {
        cx<T> A, B;
        r2c_separate(v[E / 2], v[E / 2], A, B);
        oc[N / 2] = A;
        og[N / 2] = B;
      }
    }
  }
}

// ======================================================================== z r2c of row pairs
// Plain batched real transform (mechanics: 9 components): pencil p = rows 2p, 2p+1 packed as
// a + i b; a tile is 2*PPB consecutive real rows (one bulk copy); the half spectra of the two
// rows go to rows 2p and 2p+1 of `out` (row pitch ncp).
template <class T, class C, int PPB, int NG, int NS>
__global__ void __launch_bounds__(NG *PPB *C::TP, 1)
    k_zfwd_pairs_tma(const T *in, cx<T> *out, long long nrows, int ncp, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  static_assert(E % 2 == 0 && GT % 32 == 0, "whole warps per group, even points per thread");
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  unsigned char *base = align128(smem_raw);
  T *slots = reinterpret_cast<T *>(base);                                          // [NG][NS][2*PPB*N] real
  cx<T> *xbuf = reinterpret_cast<cx<T> *>(slots + (size_t)NG * NS * 2 * PPB * N);  // [NG][PPB][NP]
  uint64_t *full = reinterpret_cast<uint64_t *>(xbuf + (size_t)NG * PPB * NP);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int pl = gt / TP;
  int plane;
  const int t = pair_map<TP>(gt % TP, tid & 31, plane);
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
  const long long npencils = (nrows + 1) / 2;
  const long long ntiles = (npencils + PPB - 1) / PPB;
  const long long stride = (long long)gridDim.x * NG;
  const long long first = (long long)blockIdx.x * NG + g;
  const int nloc = (first < ntiles) ? (int)((ntiles - first + stride - 1) / stride) : 0;
  T *gs = slots + (size_t)g * NS * 2 * PPB * N;
  uint64_t *gb = full + g * NS;

  auto issue = [&](int j) {
    const long long row0 = (first + j * stride) * PPB * 2;
    const long long left = nrows - row0;
    const int rows = left < 2 * PPB ? (int)left : 2 * PPB;
    const int s = j % NS;
    const uint32_t bytes = (uint32_t)(rows * N * sizeof(T));
    mbar_expect_tx(&gb[s], bytes);
    bulk_load_1d(gs + (size_t)s * 2 * PPB * N, in + row0 * N, bytes, &gb[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (gt == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  [[maybe_unused]] double dacc = 0.0;
  for (int j = 0; j < nloc; ++j) {
    const int s = j % NS;
    const long long p = (first + j * stride) * PPB + pl;
    const long long r0 = 2 * p;
    const bool ok = p < npencils, ok2 = r0 + 1 < nrows;
    mbar_wait(&gb[s], (uint32_t)((j / NS) & 1));
    const T *src = gs + (size_t)s * 2 * PPB * N + (size_t)(2 * pl) * N;
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = mk<T>(ok ? src[t + TP * e] : T(0), ok2 ? src[N + t + TP * e] : T(0));
    bar.sync_release();
    if (gt == 0 && j + NS < nloc) issue(j + NS);
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, NoHook());
    cx<T> w[E / 2];
    MRL_UNROLL
    for (int e = 0; e < E / 2; ++e) w[e] = shfl_cx(v[E - 1 - e], plane);
    if (t == 0) {
      w[0] = v[0];
      MRL_UNROLL
      for (int e = 1; e < E / 2; ++e) w[e] = v[E - e];
    }
    if (ok) {
      cx<T> *oa = out + r0 * ncp + t, *ob = oa + ncp;
      MRL_UNROLL
      for (int e = 0; e < E / 2; ++e) {
        cx<T> A, B;
        r2c_separate(v[e], w[e], A, B);
        oa[TP * e] = A;
        if (ok2) ob[TP * e] = B;
      }
      if (t == 0) {
        cx<T> A, B;
        r2c_separate(v[E / 2], v[E / 2], A, B);
        oa[N / 2] = A;
        if (ok2) ob[N / 2] = B;
      }
    }
  }
}

// ======================================================================== P5: z c2r of row pairs
// A tile is PPB pencils = 2*PPB consecutive half-spectrum rows (one contiguous bulk copy).
// DOT: the inner product of the result with `dotv` (same layout as `out`) rides in the store epilogue; every CTA leaves
// its partial sum in partials[blockIdx.x] (mechanics: p.Ap of the CG iteration without re-reading Ap).
template <class T, class C, int PPB, int NG, int NS, bool DOT = false>
__global__ void __launch_bounds__(NG *PPB *C::TP, 1)
    k_zinv_tma(const cx<T> *in, int ncp, T *out, long long nrows, T scale, const cx<T> *tw_g, const T *dotv, double *partials) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = PPB * TP;
  constexpr int NP = N + (N >> 3) + 1;
  constexpr int NCMAX = (N / 2 + 1 + 15) & ~15;  // largest padded row pitch (ncp <= NCMAX)
  constexpr int SLOT = 2 * PPB * NCMAX;          // complex elements per slot
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));  // [NG][NS][SLOT]
  cx<T> *xbuf = slots + (size_t)NG * NS * SLOT;                  // [NG][PPB][NP]
  uint64_t *full = reinterpret_cast<uint64_t *>(xbuf + (size_t)NG * PPB * NP);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int t = gt % TP, pl = gt / TP;
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
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
    const uint32_t bytes = (uint32_t)(rows * ncp * sizeof(cx<T>));
    mbar_expect_tx(&gb[s], bytes);
    bulk_load_1d(gs + (size_t)s * SLOT, in + row0 * ncp, bytes, &gb[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < NG * NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (gt == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  const SmPencil<T> sm{xbuf + (size_t)(g * PPB + pl) * NP};
  [[maybe_unused]] double dacc = 0.0;
  for (int j = 0; j < nloc; ++j) {
    const int s = j % NS;
    const long long p = (first + j * stride) * PPB + pl;
    const long long r0 = 2 * p;
    const bool ok = p < npencils, ok2 = r0 + 1 < nrows;
    mbar_wait(&gb[s], (uint32_t)((j / NS) & 1));
    const cx<T> *X = gs + (size_t)s * SLOT + (size_t)(2 * pl) * ncp, *Y = X + ncp;
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
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, NoHook());
    if (ok) {
      MRL_UNROLL
      for (int e = 0; e < E; ++e) {
        const int jj = t + TP * e;
        const T o0 = v[e].x * scale;
        out[r0 * N + jj] = o0;
        if (DOT) dacc += (double)o0 * (double)dotv[r0 * N + jj];
        if (ok2) {
          const T o1 = -v[e].y * scale;
          out[(r0 + 1) * N + jj] = o1;
          if (DOT) dacc += (double)o1 * (double)dotv[(r0 + 1) * N + jj];
        }
      }
    }
  }
  if (DOT) {
    // fixed-order reduction: lanes (butterfly), then the warps in index order
    __shared__ double red[32];
    const int lane = tid & 31, warp = tid >> 5;
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_sync(0xffffffffu, dacc, lane ^ o);
    if (lane == 0) red[warp] = dacc;
    __syncthreads();
    if (tid == 0) {
      double r = 0.0;
      for (int w = 0; w < (NG * PPB * TP + 31) / 32; ++w) r += red[w];
      partials[blockIdx.x] = r;
    }
  }
}

}  // namespace mrl
