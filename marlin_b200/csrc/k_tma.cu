// Launchers of the TMA-pipelined passes (mrl_passes_tma.cuh) + tensor-map construction.
#include <cstdlib>
#include <cstring>

#include "k_common.cuh"
#include "mrl_mech_tma.cuh"
#include "mrl_passes_slab.cuh"

namespace mrl {

// ------------------------------------------------------------------ tensor maps
// cuTensorMapEncodeTiled is fetched from the driver at run time (no link-time libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}
static CUtensorMapL2promotion l2_promotion() {
  static int v = env_int("MRL_L2PROMO", 128);
  return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}

// Real view [d2][d1][d0] (d0 fastest, in scalars of type T) of a complex array; strides in bytes.
template <class T>
static cudaError_t make_map3(CUtensorMap *tm, const void *base, unsigned long long d0, unsigned long long d1,
                             unsigned long long d2, unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  if (((unsigned long long)base & 15ull) || (s1 & 15ull) || (s2 & 15ull)) return cudaErrorNotSupported;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(tm, dt, 3, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Slab staging [d3 = ranks][d2 = x][d1 = y_local][d0]: one box = all ranks x all local rows of one x
template <class T>
static cudaError_t make_map4(CUtensorMap *tm, const void *base, unsigned long long d0, unsigned long long d1,
                             unsigned long long d2, unsigned long long d3, unsigned b0, unsigned long long d2_full = 0) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  // d2_full: the x extent of the array when the map covers an x sub-range of it (the rank blocks keep their distance)
  const unsigned long long s1 = d0 * sizeof(T), s2 = s1 * d1, s3 = s2 * (d2_full ? d2_full : d2);
  if (((unsigned long long)base & 15ull) || (s1 & 15ull) || d1 > 256 || d3 > 256) return cudaErrorNotSupported;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1, s2, s3};
  cuuint32_t box[4] = {b0, (cuuint32_t)d1, 1, (cuuint32_t)d3};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(tm, dt, 4, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// General tiled map of rank `rank` (<= 5) over scalars of type T: dims / box fastest first, strides in bytes for dims 1..
template <class T>
static cudaError_t make_mapn(CUtensorMap *tm, const void *base, int rank, const unsigned long long *dims, const unsigned long long *strides,
                             const unsigned *box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  if ((unsigned long long)base & 15ull) return cudaErrorNotSupported;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    if (box[i] > 256 || box[i] == 0) return cudaErrorNotSupported;
    if (i > 0) {
      st[i - 1] = strides[i - 1];
      if (strides[i - 1] & 15ull) return cudaErrorNotSupported;
    }
  }
  const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(tm, dt, rank, const_cast<void *>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

bool tma_enabled() {
  static int on = [] {
    const int dbg = env_int("MRL_DEBUG_NOFFT", 0);
    if (dbg) cudaMemcpyToSymbol(g_debug_nofft, &dbg, sizeof(int));
    return env_int("MRL_TMA", 1);
  }();
  return on != 0;
}

static constexpr size_t kSmemBudget = 225 * 1024;

// ------------------------------------------------------------------ strided
template <class T, class C, int TK, int NG, int NS>
static cudaError_t strided_tma_go(const LaunchCtx &lc, const StridedIO<T> &io0, const cx<T> *tw) {
  constexpr size_t smem = (size_t)(NS * C::N * TK) * sizeof(cx<T>) + NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "strided_tma: shared memory budget");
  static_assert(NG * TK * C::TP <= 1024, "strided_tma: block size");
  // input (field, outer) slices must be laid out back to back so that they form one tensor
  // dimension; the output may have its own pitch / slice stride / per-field base
  const long long slice = (long long)io0.n * io0.pitch;
  if (io0.nouter > 1 && io0.outer_stride != slice) return cudaErrorNotSupported;
  if (io0.nfields > 2) return cudaErrorNotSupported;
  for (int f = 1; f < io0.nfields; ++f)
    if (io0.in[f] != io0.in[0] + (long long)f * io0.nouter * slice) return cudaErrorNotSupported;
  StridedTmaIO<T> io;
  io.out = io0.out[0];
  io.out1 = io0.nfields > 1 ? io0.out[1] : io0.out[0];
  io.nouter_f = io0.nouter;
  io.n = io0.n;
  io.ncols = io0.ncols;
  io.nouter = io0.nouter * io0.nfields;
  io.nvalid = io0.nvalid ? io0.nvalid : io0.ncols;
  io.peer_tab = io0.peer_tab;
  io.peer_rows = io0.peer_rows;
  io.peer_field = io0.peer_field;
  io.peer_off = io0.peer_off;
  io.pitch = io0.out_pitch ? io0.out_pitch : io0.pitch;
  io.outer_stride = io0.out_pitch ? io0.out_outer_stride : slice;
  io.ncb = (io.ncols + TK - 1) / TK;
  io.scale = io0.scale;
  io.inverse = io0.inverse;
  CUtensorMap tm;
  const unsigned long long rowb = (unsigned long long)io0.pitch * sizeof(cx<T>);
  cudaError_t e = make_map3<T>(&tm, io0.in[0], 2ull * io.ncols, io.n, io.nouter, rowb, rowb * io.n, 2 * TK,
                               C::N < 256 ? C::N : 256);
  if (e != cudaSuccess) return e;
  auto k = k_strided_tma<T, C, TK, NG, NS>;
  int per_sm = 0;
  e = kernel_prep((const void *)k, NG * TK * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long ntiles = (long long)io.nouter * io.ncb;
  const int grid = (int)(ntiles < lc.sm_count ? ntiles : lc.sm_count);
  k<<<grid, NG * TK * C::TP, smem, lc.stream>>>(tm, io, tw);
  return cudaGetLastError();
}

template <class T> cudaError_t launch_strided_tma(const LaunchCtx &lc, const StridedIO<T> &io, const cx<T> *tw, int n) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  static int variant = env_int("MRL_STRIDED_V", 0);
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return strided_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 8>(lc, io, tw);
      case 256:
        if (variant == 1) return strided_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 1, 3>(lc, io, tw);
        return strided_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 8, 2, 6>(lc, io, tw);
      case 512:
        switch (variant) {
          case 1: return strided_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 2, 3>(lc, io, tw);
          default: return strided_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 1, 3>(lc, io, tw);
        }
      case 1024: return strided_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 4, 1, 3>(lc, io, tw);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return strided_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 16, 4, 8>(lc, io, tw);
      case 256: return strided_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 2, 6>(lc, io, tw);
      case 512: return strided_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 16, 1, 3>(lc, io, tw);
      case 1024: return strided_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 8, 1, 3>(lc, io, tw);
      default: return cudaErrorNotSupported;
    }
  }
}

// ------------------------------------------------------------------ fused P3
template <class T, class C, int TK, int NG, int SLAB>
static cudaError_t fused_tma_go1(const LaunchCtx &lc, const FusedIO<T> &io0, const SpectralUpdate<T> &up0, const cx<T> *tw) {
  constexpr size_t smem = (size_t)(NG * 3 * C::N * TK) * sizeof(cx<T>) + NG * 3 * 8 + 128;
  static_assert(smem <= kSmemBudget, "fused_tma: shared memory budget");
  static_assert(NG * TK * C::TP <= 1024, "fused_tma: block size");
  if (!io0.slab && io0.nouter != 1) return cudaErrorNotSupported;
  if (io0.slab && (io0.nyl * io0.nranks != io0.n || io0.pitch != io0.ncols)) return cudaErrorNotSupported;
  if (io0.slab == 2 && (io0.ncols % TK || io0.nouter > 256 || !io0.peer_tab)) return cudaErrorNotSupported;
  FusedTmaIO<T> io;
  io.outU = io0.outU;
  io.n = io0.n;
  io.ncols = io0.ncols;
  io.ncb = (io0.ncols + TK - 1) / TK;
  io.pitch = io0.pitch;
  io.scale = io0.scale;
  io.slab = io0.slab;
  io.nouter = io0.slab ? io0.nouter : 1;
  io.nouter_full = io0.slab == 1 && io0.nouter_full ? io0.nouter_full : io.nouter;
  io.nyl = io0.nyl;
  io.peer_tab = io0.slab ? io0.peer_tab : nullptr;
  io.peer_x0 = io0.peer_x0;
  io.kzb_major = io0.kzb_major;
  io.nx = io0.nx;
  io.rank = io0.rank;
  io.nranks = io0.nranks;
  io.flag_wait = io0.slab == 2 ? io0.flag_wait : nullptr;
  io.flag_expect = io0.flag_expect;
  io.flag_tab = io0.slab == 2 ? io0.flag_tab : nullptr;
  SpectralUpdate2<T> up;
  memset(&up, 0, sizeof up);
  up.kx = up0.kx; up.ky = up0.ky; up.kz = up0.kz;
  up.kmode = up0.kmode; up.nzc = up0.nzc; up.x0 = up0.x0;
  up.nzv = up0.nzv ? up0.nzv : up0.nzc;
  up.closed_M = up0.closed_M; up.closed_L = up0.closed_L; up.has_L = up0.has_L;
  up.Mfac = up0.Mfac; up.Lfac = up0.Lfac; up.Mbuf = up0.Mbuf; up.Lbuf = up0.Lbuf;
  up.dt = up0.dt; up.b0 = up0.b0; up.nold = up0.nold;
  up.bold0 = up0.bold[0]; up.bold1 = up0.bold[1]; up.bold2 = up0.bold[2]; up.bold3 = up0.bold[3];
  up.Nold1 = up0.Nold[1]; up.Nold2 = up0.Nold[2]; up.Nold3 = up0.Nold[3];
  up.Nout = up0.Nout;
  const unsigned long long rowb = (unsigned long long)io.pitch * sizeof(cx<T>);
  const unsigned boxr = C::N < 256 ? C::N : 256;
  CUtensorMap tmC, tmG, tmO;
  const void *oldp = up0.nold > 0 ? (const void *)up0.Nold[0] : (const void *)io0.inC;
  cudaError_t e;
  if (io.slab == 2) {
    // blocked staging R = [source][ncb][nyl][nxl][W]: one box = (W columns, one xl, all yl, one block, all sources)
    const unsigned long long W2 = 2ull * TK, es = sizeof(T);
    const unsigned long long dims[5] = {W2, (unsigned long long)io.nouter, (unsigned long long)io.nyl, (unsigned long long)io.ncb,
                                        (unsigned long long)io0.nranks};
    const unsigned long long st[4] = {W2 * es, W2 * es * io.nouter, W2 * es * io.nouter * io.nyl, W2 * es * io.nouter * io.nyl * io.ncb};
    const unsigned box[5] = {(unsigned)W2, 1u, (unsigned)io.nyl, 1u, (unsigned)io0.nranks};
    e = make_mapn<T>(&tmC, io0.inC, 5, dims, st, box);
    if (e == cudaSuccess) e = make_mapn<T>(&tmG, io0.inG, 5, dims, st, box);
    if (e == cudaSuccess) e = make_map4<T>(&tmO, up0.nold > 0 ? (const void *)up0.Nold[0] : io0.ring_old, 2ull * io.ncols, io.nyl, io.nouter, io0.nranks, 2 * TK);
  } else if (io.slab) {
    e = make_map4<T>(&tmC, io0.inC, 2ull * io.ncols, io.nyl, io.nouter, io0.nranks, 2 * TK, io0.nouter_full);
    if (e == cudaSuccess) e = make_map4<T>(&tmG, io0.inG, 2ull * io.ncols, io.nyl, io.nouter, io0.nranks, 2 * TK, io0.nouter_full);
    if (e == cudaSuccess) e = make_map4<T>(&tmO, oldp, 2ull * io.ncols, io.nyl, io.nouter, io0.nranks, 2 * TK, io0.nouter_full);
  } else {
    e = make_map3<T>(&tmC, io0.inC, 2ull * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
    if (e == cudaSuccess) e = make_map3<T>(&tmG, io0.inG, 2ull * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
    if (e == cudaSuccess) e = make_map3<T>(&tmO, oldp, 2ull * io.ncols, io.n, 1, rowb, rowb * io.n, 2 * TK, boxr);
  }
  if (e != cudaSuccess) return e;
  auto k = k_fused_tma<T, C, TK, NG, SLAB>;
  int per_sm = 0;
  e = kernel_prep((const void *)k, NG * TK * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nwork = ((long long)io.nouter * io.ncb + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, NG * TK * C::TP, smem, lc.stream>>>(tmC, tmG, tmO, io, up, tw);
  return cudaGetLastError();
}
template <class T, class C, int TK, int NG>
static cudaError_t fused_tma_go(const LaunchCtx &lc, const FusedIO<T> &io, const SpectralUpdate<T> &up, const cx<T> *tw) {
  switch (io.slab) {
    case 0: return fused_tma_go1<T, C, TK, NG, 0>(lc, io, up, tw);
    case 1: return fused_tma_go1<T, C, TK, NG, 1>(lc, io, up, tw);
    case 2: return fused_tma_go1<T, C, TK, NG, 2>(lc, io, up, tw);
    default: return cudaErrorNotSupported;
  }
}

template <class T>
cudaError_t launch_fused_tma(const LaunchCtx &lc, const FusedIO<T> &io, const SpectralUpdate<T> &up, const cx<T> *tw, int n) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  static int variant = env_int("MRL_FUSED_V", 0);
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return fused_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4>(lc, io, up, tw);
      case 256: return fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 8, 2>(lc, io, up, tw);
      case 512:
        switch (variant) {
          case 1: return fused_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 4, 2>(lc, io, up, tw);
          default: return fused_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 1>(lc, io, up, tw);
        }
      case 1024: return fused_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 4, 1>(lc, io, up, tw);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return fused_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 16, 2>(lc, io, up, tw);
      case 256: return fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 2>(lc, io, up, tw);
      case 512: return fused_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 2>(lc, io, up, tw);
      case 1024: return fused_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 8, 1>(lc, io, up, tw);
      default: return cudaErrorNotSupported;
    }
  }
}

// ------------------------------------------------------------------ mechanics: fused Green projection
template <class T, class C, int TK, int NG>
static cudaError_t mech_fused_tma_go(const LaunchCtx &lc, cx<T> *spec, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzv,
                                     int ncp, const cx<T> *tw) {
  constexpr size_t smem = (size_t)(NG * 3 * C::N * TK) * sizeof(cx<T>) + NG * 3 * 8 + 128;
  static_assert(smem <= kSmemBudget, "mech_fused_tma: shared memory budget");
  static_assert(NG * TK * C::TP <= 1024, "mech_fused_tma: block size");
  MechFusedTmaIO<T> io;
  io.out = spec;
  io.n = n0;
  io.ncols = n1 * ncp;
  io.ncb = (io.ncols + TK - 1) / TK;
  io.pitch = io.ncols;
  io.field = (long long)n0 * io.ncols;
  io.kx = kx; io.ky = ky; io.kz = kz;
  io.ncp = ncp;
  io.nzv = nzv;
  io.scale = T(1);
  CUtensorMap tm;
  const unsigned long long rowb = (unsigned long long)io.pitch * sizeof(cx<T>);
  cudaError_t e = make_map3<T>(&tm, spec, 2ull * io.ncols, io.n, 9, rowb, rowb * io.n, 2 * TK, C::N < 256 ? C::N : 256);
  if (e != cudaSuccess) return e;
  auto k = k_mech_fused_tma<T, C, TK, NG>;
  int per_sm = 0;
  e = kernel_prep((const void *)k, NG * TK * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nwork = (3ll * io.ncb + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, NG * TK * C::TP, smem, lc.stream>>>(tm, io, tw);
  return cudaGetLastError();
}

// In place on spec = [9][n0][n1][ncp]: x-forward, Green projection, x-inverse (unnormalised).
template <class T>
cudaError_t launch_mech_fused_tma(const LaunchCtx &lc, cx<T> *spec, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzv,
                                  int ncp, const cx<T> *tw) {
  static const int enabled = env_int("MRL_MECH_FUSED", 1), variant = env_int("MRL_MECH_V", 0);
  if (!tma_enabled() || !enabled) return cudaErrorNotSupported;
  if constexpr (sizeof(T) == 8) {
    switch (n0) {
      case 128: return mech_fused_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 256:
        if (variant == 1) return mech_fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
        if (variant == 2) return mech_fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 3>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
        return mech_fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 8, 2>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 512: return mech_fused_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 1>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 1024: return mech_fused_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 4, 1>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n0) {
      case 128: return mech_fused_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 16, 2>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 256: return mech_fused_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 2>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 512: return mech_fused_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 2>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      case 1024: return mech_fused_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 8, 1>(lc, spec, kx, ky, kz, n0, n1, nzv, ncp, tw);
      default: return cudaErrorNotSupported;
    }
  }
}

// ------------------------------------------------------------------ P1 / P5
template <class T, class C, int PPB, int NG, int NS, class F>
static cudaError_t zfwd_tma_go(const LaunchCtx &lc, const T *c, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows, int ncp,
                               const F &f, const cx<T> *tw, const RowMap &rm) {
  constexpr int NP = C::N + (C::N >> 3) + 1;
  constexpr size_t smem = (size_t)NG * NS * PPB * C::N * sizeof(T) + (size_t)(NG * PPB * NP) * sizeof(cx<T>) + NG * NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "zfwd_tma: shared memory budget");
  static_assert(NG * PPB * C::TP <= 1024 && NG <= 15, "zfwd_tma: block size");
  if (((unsigned long long)c & 15ull) || (C::N * sizeof(T)) % 16) return cudaErrorNotSupported;
  if (rm.ych && rm.ych % PPB) return cudaErrorNotSupported;
  auto k = k_zfwd_tma<T, C, PPB, NG, NS, F>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, NG * PPB * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nwork = ((nrows + PPB - 1) / PPB + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, NG * PPB * C::TP, smem, lc.stream>>>(c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_zfwd_nonlin_tma(const LaunchCtx &lc, const T *c, T *mu_out, cx<T> *outC, cx<T> *outG, long long nrows, int n,
                                   int ncp, const NonlinDesc &nl, const cx<T> *tw, const RowMap &rm) {
  if (!tma_enabled() || nl.kind != 0) return cudaErrorNotSupported;
  static int variant = env_int("MRL_ZFWD_V", 0);
  typedef DoubleWellDeriv<T> F;
  const F f{(T)nl.p[0], (T)nl.p[1], (T)nl.p[2]};
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return zfwd_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      case 256: return zfwd_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      case 512:
        switch (variant) {
          case 1: return zfwd_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 4, 2, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
          case 2: return zfwd_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 1, 8, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
          case 3: return zfwd_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
          default: return zfwd_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 5, 3, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
        }
      case 1024: return zfwd_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 4, 3, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return zfwd_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      case 256: return zfwd_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      case 512: return zfwd_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 4, 4, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      case 1024: return zfwd_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 4, 3, F>(lc, c, mu_out, outC, outG, nrows, ncp, f, tw, rm);
      default: return cudaErrorNotSupported;
    }
  }
}

template <class T, class C, int PPB, int NG, int NS>
static cudaError_t zfwd_pairs_tma_go(const LaunchCtx &lc, const T *in, cx<T> *out, long long nrows, int ncp, const cx<T> *tw) {
  constexpr int NP = C::N + (C::N >> 3) + 1;
  constexpr size_t smem = (size_t)NG * NS * 2 * PPB * C::N * sizeof(T) + (size_t)(NG * PPB * NP) * sizeof(cx<T>) + NG * NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "zfwd_pairs_tma: shared memory budget");
  static_assert(NG * PPB * C::TP <= 1024 && NG <= 15, "zfwd_pairs_tma: block size");
  if (((unsigned long long)in & 15ull) || (C::N * sizeof(T)) % 16) return cudaErrorNotSupported;
  auto k = k_zfwd_pairs_tma<T, C, PPB, NG, NS>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, NG * PPB * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long npencils = (nrows + 1) / 2;
  const long long nwork = ((npencils + PPB - 1) / PPB + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, NG * PPB * C::TP, smem, lc.stream>>>(in, out, nrows, ncp, tw);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_zfwd_pairs_tma(const LaunchCtx &lc, const T *in, cx<T> *out, long long nrows, int n, int ncp, const cx<T> *tw) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  switch (n) {
    case 128: return zfwd_pairs_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 2>(lc, in, out, nrows, ncp, tw);
    case 256: return zfwd_pairs_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4, 2>(lc, in, out, nrows, ncp, tw);
    case 512: return zfwd_pairs_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 4, 2>(lc, in, out, nrows, ncp, tw);
    case 1024: return zfwd_pairs_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 4, 2>(lc, in, out, nrows, ncp, tw);
    default: return cudaErrorNotSupported;
  }
}

template <class T, class C, int NG> static cudaError_t mech_tangent_go(const LaunchCtx &lc, const MechTangentIO<T> &io, const cx<T> *tw) {
  constexpr int NP = C::N + (C::N >> 3) + 1;
  constexpr size_t smem = (size_t)NG * 9 * 2 * C::N * sizeof(T) + (size_t)(NG * 9 * NP) * sizeof(cx<T>) + 128;
  static_assert(smem <= kSmemBudget, "mech_tangent_zfwd: shared memory budget");
  static_assert(NG * 9 * C::TP <= 1024 && NG <= 15, "mech_tangent_zfwd: block size");
  if (io.nrows % 2 || io.n != io.nrows * C::N) return cudaErrorNotSupported;
  auto k = k_mech_tangent_zfwd<T, C, NG>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, NG * 9 * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nwork = (io.nrows / 2 + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, NG * 9 * C::TP, smem, lc.stream>>>(io, tw);
  return cudaGetLastError();
}
// bulk-copy staged variant: one group per CTA, two input stages of 29 rows
template <class T, class C> static cudaError_t mech_tangent_tma_go(const LaunchCtx &lc, const MechTangentIO<T> &io, const cx<T> *tw) {
  constexpr int NP = C::N + (C::N >> 3) + 1;
  constexpr size_t smem = (size_t)(2 * 29 + 9 * 2) * C::N * sizeof(T) + (size_t)(9 * NP) * sizeof(cx<T>) + 2 * 8 + 128;
  static_assert(smem <= kSmemBudget, "mech_tangent_zfwd_tma: shared memory budget");
  if (io.nrows % 2 || io.n != io.nrows * C::N || (C::N * sizeof(T)) % 16) return cudaErrorNotSupported;
  for (const void *q : {(const void *)io.F, (const void *)io.K, (const void *)io.mu, (const void *)io.p, (const void *)io.r})
    if ((unsigned long long)q & 15ull) return cudaErrorNotSupported;
  if ((io.n * sizeof(T)) % 16) return cudaErrorNotSupported;
  auto k = k_mech_tangent_zfwd_tma<T, C>;
  int per_sm = 0;
  cudaError_t e = kernel_prep((const void *)k, 9 * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long nwork = io.nrows / 2;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  k<<<grid, 9 * C::TP, smem, lc.stream>>>(io, tw);
  return cudaGetLastError();
}
template <class T> cudaError_t launch_mech_tangent_zfwd(const LaunchCtx &lc, const MechTangentIO<T> &io, const cx<T> *tw, int n) {
  if (!tma_enabled() || env_int("MRL_MECH_TANGENT_FUSED", 1) == 0) return cudaErrorNotSupported;
  static int variant = env_int("MRL_MECH_TANGENT_V", 0);  // 0: bulk-copy staged inputs where they fit, 1: loads to registers
  switch (n) {
    case 256:
      if (variant == 0) {
        cudaError_t e = mech_tangent_tma_go<T, FFTCfg<256, 32, 8, 8, 4>>(lc, io, tw);
        if (e != cudaErrorNotSupported) return e;
      }
      return mech_tangent_go<T, FFTCfg<256, 32, 8, 8, 4>, 2>(lc, io, tw);
    case 512: return mech_tangent_go<T, FFTCfg<512, 64, 8, 8, 8>, 1>(lc, io, tw);
    default: return cudaErrorNotSupported;
  }
}

template <class T, class C, int PPB, int NG, int NS>
static cudaError_t zinv_tma_go(const LaunchCtx &lc, const cx<T> *in, int ncp, T *out, long long nrows, T scale, const cx<T> *tw,
                               const ZinvDot<T> *dot) {
  constexpr int NP = C::N + (C::N >> 3) + 1;
  constexpr int NC = (C::N / 2 + 1 + 15) & ~15;  // slot rows sized for the largest padded pitch
  if (ncp > NC || ncp < C::N / 2 + 1) return cudaErrorNotSupported;
  constexpr size_t smem = (size_t)(NG * NS * 2 * PPB * NC + NG * PPB * NP) * sizeof(cx<T>) + NG * NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "zinv_tma: shared memory budget");
  static_assert(NG * PPB * C::TP <= 1024 && NG <= 15, "zinv_tma: block size");
  if (((unsigned long long)in & 15ull) || (ncp * sizeof(cx<T>)) % 16) return cudaErrorNotSupported;
  const long long npencils = (nrows + 1) / 2;
  const long long nwork = ((npencils + PPB - 1) / PPB + NG - 1) / NG;
  const int grid = (int)(nwork < lc.sm_count ? nwork : lc.sm_count);
  int per_sm = 0;
  if (dot && dot->with) {
    if (grid > dot->capacity) return cudaErrorInvalidValue;
    auto k = k_zinv_tma<T, C, PPB, NG, NS, true>;
    cudaError_t e = kernel_prep((const void *)k, NG * PPB * C::TP, smem, &per_sm);
    if (e != cudaSuccess) return e;
    k<<<grid, NG * PPB * C::TP, smem, lc.stream>>>(in, ncp, out, nrows, scale, tw, dot->with, dot->partials);
    *dot->count = grid;
    return cudaGetLastError();
  }
  auto k = k_zinv_tma<T, C, PPB, NG, NS, false>;
  cudaError_t e = kernel_prep((const void *)k, NG * PPB * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  k<<<grid, NG * PPB * C::TP, smem, lc.stream>>>(in, ncp, out, nrows, scale, tw, nullptr, nullptr);
  return cudaGetLastError();
}

template <class T>
cudaError_t launch_zinv_pairs_tma(const LaunchCtx &lc, const cx<T> *in, int ncp, T *out, long long nrows, int n, T scale,
                                  const cx<T> *tw, const ZinvDot<T> *dot) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  static int variant = env_int("MRL_ZINV_V", 0);
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return zinv_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 3, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
      case 256: return zinv_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
      case 512:
        switch (variant) {  // measured (profiles/r1z_variants.txt): 4 pencils x 2 groups 0.337 ms, 2 x 4: 0.379 ms
          case 1: return zinv_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 4, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
          case 2: return zinv_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 1, 8, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
          case 3: return zinv_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 3, 3>(lc, in, ncp, out, nrows, scale, tw, dot);
          default: return zinv_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 4, 2, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
        }
      case 1024: return zinv_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 4, 2>(lc, in, ncp, out, nrows, scale, tw, dot);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return zinv_tma_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 3>(lc, in, ncp, out, nrows, scale, tw, dot);
      case 256: return zinv_tma_go<T, FFTCfg<256, 32, 8, 8, 4>, 4, 4, 3>(lc, in, ncp, out, nrows, scale, tw, dot);
      case 512: return zinv_tma_go<T, FFTCfg<512, 64, 8, 8, 8>, 2, 4, 3>(lc, in, ncp, out, nrows, scale, tw, dot);
      case 1024: return zinv_tma_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 1, 4, 3>(lc, in, ncp, out, nrows, scale, tw, dot);
      default: return cudaErrorNotSupported;
    }
  }
}

// ------------------------------------------------------------------ slab x passes with bulk peer stores
template <class T, class C, int TK, int NG, int NS>
static cudaError_t slab_xfwd_go(const LaunchCtx &lc, const cx<T> *in, const SlabXIO<T> &io, const cx<T> *tw) {
  constexpr size_t smem = (size_t)((NS + NG) * C::N * TK) * sizeof(cx<T>) + NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "slab_xfwd: shared memory budget");
  static_assert(NG * TK * C::TP <= 1024, "slab_xfwd: block size");
  if (io.n != C::N || io.nxl * io.nranks != io.n) return cudaErrorNotSupported;
  const int ncp = io.kb * TK;
  CUtensorMap tm;
  const unsigned long long rowb = (unsigned long long)io.nyl * ncp * sizeof(cx<T>);
  cudaError_t e = make_map3<T>(&tm, in, 2ull * io.nyl * ncp, io.n, io.nf, rowb, rowb * io.n, 2 * TK, C::N < 256 ? C::N : 256);
  if (e != cudaSuccess) return e;
  auto k = k_slab_xfwd<T, C, TK, NG, NS>;
  int per_sm = 0;
  e = kernel_prep((const void *)k, NG * TK * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long ntiles = (long long)io.nf * io.ych * io.kb;
  const int grid = (int)(ntiles < lc.sm_count ? ntiles : lc.sm_count);
  k<<<grid, NG * TK * C::TP, smem, lc.stream>>>(tm, io, tw);
  return cudaGetLastError();
}

// Forward x pass of io.nf fields of the natural slab `in` = [nf][nx][nyl][ncp]; the result rows go to the peers' blocked
// staging R (io.peer_tab) as bulk copies.  The column-block width of every size equals fused_tma_tk().
template <class T> cudaError_t launch_slab_xfwd(const LaunchCtx &lc, const cx<T> *in, const SlabXIO<T> &io, const cx<T> *tw, int n) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return slab_xfwd_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 6>(lc, in, io, tw);
      case 256: return slab_xfwd_go<T, FFTCfg<256, 32, 8, 8, 4>, 8, 2, 4>(lc, in, io, tw);
      case 512: return slab_xfwd_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 1, 2>(lc, in, io, tw);
      case 1024: return slab_xfwd_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 4, 1, 2>(lc, in, io, tw);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return slab_xfwd_go<T, FFTCfg<128, 16, 8, 4, 4>, 16, 2, 6>(lc, in, io, tw);
      case 256: return slab_xfwd_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 2, 4>(lc, in, io, tw);
      case 512: return slab_xfwd_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 2, 4>(lc, in, io, tw);
      case 1024: return slab_xfwd_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 8, 1, 2>(lc, in, io, tw);
      default: return cudaErrorNotSupported;
    }
  }
}

template <class T, class C, int TK, int NG, int NS>
static cudaError_t slab_xinv_go(const LaunchCtx &lc, const cx<T> *S, const SlabXIO<T> &io, const cx<T> *tw) {
  constexpr size_t smem = (size_t)(NS * C::N * TK) * sizeof(cx<T>) + NS * 8 + 128;
  static_assert(smem <= kSmemBudget, "slab_xinv: shared memory budget");
  static_assert(NG * TK * C::TP <= 1024, "slab_xinv: block size");
  if (io.n != C::N) return cudaErrorNotSupported;
  // S = [kb][nx][nyl][W]
  const unsigned long long W2 = 2ull * TK, es = sizeof(T);
  const unsigned long long dims[4] = {W2, (unsigned long long)io.nyl, (unsigned long long)io.n, (unsigned long long)io.kb};
  const unsigned long long st[3] = {W2 * es, W2 * es * io.nyl, W2 * es * io.nyl * io.n};
  const unsigned box[4] = {(unsigned)W2, 1u, (unsigned)(C::N < 256 ? C::N : 256), 1u};
  CUtensorMap tm;
  cudaError_t e = make_mapn<T>(&tm, S, 4, dims, st, box);
  if (e != cudaSuccess) return e;
  auto k = k_slab_xinv<T, C, TK, NG, NS>;
  int per_sm = 0;
  e = kernel_prep((const void *)k, NG * TK * C::TP, smem, &per_sm);
  if (e != cudaSuccess) return e;
  const long long ntiles = (long long)io.nyl * io.kb;
  const int grid = (int)(ntiles < lc.sm_count ? ntiles : lc.sm_count);
  k<<<grid, NG * TK * C::TP, smem, lc.stream>>>(tm, io, tw);
  return cudaGetLastError();
}

// Inverse x pass: blocked return staging S -> natural slab io.out = [nx][nyl][ncp]
template <class T> cudaError_t launch_slab_xinv(const LaunchCtx &lc, const cx<T> *S, const SlabXIO<T> &io, const cx<T> *tw, int n) {
  if (!tma_enabled()) return cudaErrorNotSupported;
  if constexpr (sizeof(T) == 8) {
    switch (n) {
      case 128: return slab_xinv_go<T, FFTCfg<128, 16, 8, 4, 4>, 8, 4, 8>(lc, S, io, tw);
      case 256: return slab_xinv_go<T, FFTCfg<256, 32, 8, 8, 4>, 8, 2, 6>(lc, S, io, tw);
      case 512: return slab_xinv_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 1, 3>(lc, S, io, tw);
      case 1024: return slab_xinv_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 4, 1, 3>(lc, S, io, tw);
      default: return cudaErrorNotSupported;
    }
  } else {
    switch (n) {
      case 128: return slab_xinv_go<T, FFTCfg<128, 16, 8, 4, 4>, 16, 4, 8>(lc, S, io, tw);
      case 256: return slab_xinv_go<T, FFTCfg<256, 32, 8, 8, 4>, 16, 2, 6>(lc, S, io, tw);
      case 512: return slab_xinv_go<T, FFTCfg<512, 64, 8, 8, 8>, 8, 2, 6>(lc, S, io, tw);
      case 1024: return slab_xinv_go<T, FFTCfg<1024, 128, 8, 8, 4, 4>, 8, 1, 3>(lc, S, io, tw);
      default: return cudaErrorNotSupported;
    }
  }
}

// column-block width (TK) of the fused pass for a transform length (the blocked layouts need the same width on all passes)
template <class T> int fused_tma_tk(int n) {
  if constexpr (sizeof(T) == 8) return n == 1024 ? 4 : (n == 128 || n == 256 || n == 512) ? 8 : 0;
  else return n == 128 || n == 256 ? 16 : (n == 512 || n == 1024) ? 8 : 0;
}

#define INST(T)                                                                                                          \
  template cudaError_t launch_mech_fused_tma<T>(const LaunchCtx &, cx<T> *, const T *, const T *, const T *, int, int, int, int, \
                                                const cx<T> *);                                                          \
  template cudaError_t launch_strided_tma<T>(const LaunchCtx &, const StridedIO<T> &, const cx<T> *, int);               \
  template cudaError_t launch_mech_tangent_zfwd<T>(const LaunchCtx &, const MechTangentIO<T> &, const cx<T> *, int);     \
  template cudaError_t launch_fused_tma<T>(const LaunchCtx &, const FusedIO<T> &, const SpectralUpdate<T> &, const cx<T> *, \
                                           int);                                                                         \
  template cudaError_t launch_zfwd_nonlin_tma<T>(const LaunchCtx &, const T *, T *, cx<T> *, cx<T> *, long long, int,    \
                                                 int, const NonlinDesc &, const cx<T> *, const RowMap &);                \
  template cudaError_t launch_zfwd_pairs_tma<T>(const LaunchCtx &, const T *, cx<T> *, long long, int, int, const cx<T> *);  \
  template cudaError_t launch_slab_xfwd<T>(const LaunchCtx &, const cx<T> *, const SlabXIO<T> &, const cx<T> *, int);        \
  template cudaError_t launch_slab_xinv<T>(const LaunchCtx &, const cx<T> *, const SlabXIO<T> &, const cx<T> *, int);         \
  template int fused_tma_tk<T>(int);                                                                                       \
  template cudaError_t launch_zinv_pairs_tma<T>(const LaunchCtx &, const cx<T> *, int, T *, long long, int, T, const cx<T> *, \
                                                const ZinvDot<T> *);
INST(double)
INST(float)

}  // namespace mrl
