// marlin_b200 - generic decomposed transforms behind DomainAction::fft / ifft:
//   slab    partitionSlabs / fftSlab / ifftSlab        src/actions/DomainAction.C:511-566, :870-938, :941-1019
//   pencil  partitionPencils / fftPencil / ifftPencil  src/actions/DomainAction.C:569-742, :1022-1047, :1106-1404
// for any grid size, 2-D and 3-D (slab), unequal parts.  Host logic only; the passes are the kernels of k_*.cu, the
// exchanges are strided peer-to-peer copies over NVLink into staging buffers shared through CUDA IPC, bracketed by the
// device-side barrier of k_reduce.cu.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/marlin_b200.h"
#include "mrl_internal.h"

using namespace mrl;

#define CK(call)                                                                                                    \
  do {                                                                                                              \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess) return mrl_fail(MRL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// One allocation per rank is shared with the peers (cudaMalloc + one IPC handle); the staging areas are sub-ranges:
//   slab    R  [nxl][ny][ncz]       forward exchange, reciprocal-side layout
//           B  [nx][nyl][ncz]       return exchange, real-side layout
//   pencil  R1 [nxc_l][ny][nzl]     stage 1 forward (x <-> y inside a z group)
//           R2 [nxc_l][ny2_l][nz]   stage 2 forward (y <-> z inside an x group)
//           B2 [nxc_l][ny][nzl]     stage 2 return
//           B1 [nx][nyl][nzl]       stage 1 return (the received half spectrum along x fills the first nx/2+1 rows)
//   flags  one 8-byte slot per rank (barrier)
enum { A_R = 0, A_B = 1, A_R2 = 2, A_B2 = 3, A_FLAGS = 4, A_COUNT = 5 };

struct mrl_dist {
  mrl_context *ctx = nullptr;
  void *shared = nullptr;
  long long off[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // byte offsets of the areas in `shared`
  void *W = nullptr;                             // private work array (the largest layout)
  std::vector<void *> opened;
  std::vector<char *> peer;                      // base of every rank's shared allocation
  std::vector<std::vector<long long>> peer_off;  // and its area offsets
  void *flag_tab = nullptr;                      // device array: every rank's flags
  unsigned long long epoch = 0;
  bool imported = false;
  long long ncz = 1;  // slab: complex entries per (x, y): nz/2+1 in 3-D, 1 in 2-D
  char *area(int s, int a) const { return peer[s] + peer_off[s][a]; }
};

template <class T> static int upload_axis(const std::vector<double> &h, void **dev) {
  std::vector<T> t(h.begin(), h.end());
  cudaFree(*dev);
  *dev = nullptr;
  CK(cudaMalloc(dev, (t.empty() ? 1 : t.size()) * sizeof(T)));
  CK(cudaMemcpy(*dev, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MRL_OK;
}
static int upload_axes(mrl_context *ctx, int d, bool reciprocal) {
  const std::vector<double> &h = reciprocal ? ctx->kaxis_h[d] : ctx->axis_h[d];
  void **dev = reciprocal ? &ctx->kaxis_dev[d] : &ctx->axis_dev[d];
  return ctx->precision == MRL_F64 ? upload_axis<double>(h, dev) : upload_axis<float>(h, dev);
}
static void begins(const std::vector<int64_t> &count, std::vector<int64_t> &begin) {
  begin.assign(count.size(), 0);
  for (size_t r = 1; r < count.size(); ++r) begin[r] = begin[r - 1] + count[r - 1];
}
template <class V> static V slice(const V &v, int64_t b, int64_t n) { return V(v.begin() + b, v.begin() + b + n); }

extern "C" int mrl_domain_set_dist(mrl_context *ctx, int dim, const int64_t *n, const double *mn, const double *mx, int rank, int nranks,
                                   const double *weights) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set_dist: bad arguments");
  // reference: "Dimension must be 2 or 3 for slab decomposition." (DomainAction.C:513-514)
  if (dim < 2 || dim > 3) return mrl_fail(MRL_ERR_INVALID, "Dimension must be 2 or 3 for slab decomposition.");
  int rc = mrl_domain_set(ctx, dim, n, mn, mx);
  if (rc) return rc;
  if (n[0] < nranks || n[1] < nranks)
    return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set_dist: every rank needs at least one layer (nx = %lld, ny = %lld, %d ranks)", (long long)n[0],
                    (long long)n[1], nranks);
  ctx->ycount.assign(nranks, 0);
  ctx->xcount.assign(nranks, 0);
  // x is partitioned along the reciprocal axis, y along the real-space axis (DomainAction.C:519-523)
  if ((rc = mrl_partition(n[0], nranks, weights, ctx->xcount.data()))) return rc;
  if ((rc = mrl_partition(n[1], nranks, weights, ctx->ycount.data()))) return rc;
  begins(ctx->xcount, ctx->xbegin);
  begins(ctx->ycount, ctx->ybegin);
  ctx->dist = true;
  ctx->pencil = false;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->nyl = (int)ctx->ycount[rank];
  ctx->nxl = (int)ctx->xcount[rank];
  ctx->y0 = (int)ctx->ybegin[rank];
  ctx->x0 = (int)ctx->xbegin[rank];
  // local shapes and axes: real [nx][nyl](,[nz]), reciprocal [nxl][ny](,[nz/2+1]) = an x-slice of the serial layout
  ctx->n[1] = ctx->nyl;
  ctx->nr[0] = ctx->nxl;
  ctx->axis_h[1] = slice(ctx->axis_h[1], ctx->y0, ctx->nyl);
  ctx->kaxis_h[0] = slice(ctx->kaxis_h[0], ctx->x0, ctx->nxl);
  if ((rc = upload_axes(ctx, 1, false))) return rc;
  return upload_axes(ctx, 0, true);
}

// the factorisation nranks = Py * Pz closest to a square that fits the domain (DomainAction.C:574-613)
extern "C" int mrl_pencil_factors(int nranks, const int64_t *n, int *py, int *pz) {
  if (nranks < 1 || !n || !py || !pz) return mrl_fail(MRL_ERR_INVALID, "mrl_pencil_factors: bad arguments");
  const int64_t nxc = n[0] / 2 + 1;
  int best_py = 0, best_pz = 0, best_cost = 0;
  bool found = false;
  auto consider = [&](int px, int pz_) {
    if (px < 2 || pz_ < 2 || px > n[1] || px > nxc || pz_ > n[2] || pz_ > n[1]) return;
    const int cost = std::abs(px - pz_);
    if (!found || cost < best_cost) {
      best_py = px;
      best_pz = pz_;
      best_cost = cost;
      found = true;
    }
  };
  const int max_div = std::max(2, (int)std::sqrt((double)nranks));
  for (int d = 2; d <= max_div; ++d)
    if (nranks % d == 0) {
      consider(d, nranks / d);
      consider(nranks / d, d);
    }
  if (!found)
    return mrl_fail(MRL_ERR_INVALID,
                    "FFT_PENCIL requires factoring the number of MPI ranks into two integers greater than one that fit the domain (ranks = %d). "
                    "Use FFT_SLAB or adjust the rank count.",
                    nranks);
  *py = best_py;
  *pz = best_pz;
  return MRL_OK;
}

// partitionPencils (DomainAction.C:569-742).  Rank r = pz * Py + py: real space [nx][ny / Py][nz / Pz] (y part py, z part
// pz); reciprocal space [(nx/2+1) / Py][ny / Pz][nz] (kx part py, ky part pz) with the half spectrum on x.
extern "C" int mrl_domain_set_pencil(mrl_context *ctx, int dim, const int64_t *n, const double *mn, const double *mx, int rank, int nranks) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set_pencil: bad arguments");
  if (dim < 3) return mrl_fail(MRL_ERR_INVALID, "Dimension must be 3 for pencil decomposition.");
  int rc = mrl_domain_set(ctx, dim, n, mn, mx);
  if (rc) return rc;
  const int64_t nxc = n[0] / 2 + 1;
  int best_py = 0, best_pz = 0;
  if ((rc = mrl_pencil_factors(nranks, n, &best_py, &best_pz))) return rc;
  const int Py = best_py, Pz = best_pz;
  ctx->ycount.assign(Py, 0);
  ctx->zcount.assign(Pz, 0);
  ctx->xcount.assign(Py, 0);
  ctx->y2count.assign(Pz, 0);
  if ((rc = mrl_partition(n[1], Py, nullptr, ctx->ycount.data()))) return rc;
  if ((rc = mrl_partition(n[2], Pz, nullptr, ctx->zcount.data()))) return rc;
  if ((rc = mrl_partition(nxc, Py, nullptr, ctx->xcount.data()))) return rc;
  if ((rc = mrl_partition(n[1], Pz, nullptr, ctx->y2count.data()))) return rc;
  begins(ctx->ycount, ctx->ybegin);
  begins(ctx->zcount, ctx->zbegin);
  begins(ctx->xcount, ctx->xbegin);
  begins(ctx->y2count, ctx->y2begin);
  ctx->dist = true;
  ctx->pencil = true;
  ctx->py = Py;
  ctx->pz = Pz;
  ctx->rank = rank;
  ctx->nranks = nranks;
  const int iy = rank % Py, iz = rank / Py;
  ctx->nyl = (int)ctx->ycount[iy];
  ctx->y0 = (int)ctx->ybegin[iy];
  ctx->nxl = (int)ctx->xcount[iy];
  ctx->x0 = (int)ctx->xbegin[iy];
  // real [nx][nyl][nzl]; reciprocal [nxc_l][ny2_l][nz]: rfftfreq on x, fftfreq on y and z (gridChanged :289-306)
  ctx->n[1] = ctx->nyl;
  ctx->n[2] = (int)ctx->zcount[iz];
  ctx->nr[0] = ctx->nxl;
  ctx->nr[1] = (int)ctx->y2count[iz];
  ctx->nr[2] = (int)n[2];
  ctx->axis_h[1] = slice(ctx->axis_h[1], ctx->y0, ctx->nyl);
  ctx->axis_h[2] = slice(ctx->axis_h[2], ctx->zbegin[iz], ctx->zcount[iz]);
  std::vector<double> kx(nxc), kz(n[2]);
  if ((rc = mrl_axis_values(n[0], mn[0], mx[0], 1, 1, kx.data()))) return rc;
  if ((rc = mrl_axis_values(n[2], mn[2], mx[2], 1, 0, kz.data()))) return rc;
  ctx->kaxis_h[0] = slice(kx, ctx->x0, ctx->nxl);
  ctx->kaxis_h[1] = slice(ctx->kaxis_h[1], ctx->y2begin[iz], ctx->y2count[iz]);
  ctx->kaxis_h[2] = kz;
  for (int d = 1; d < 3; ++d)
    if ((rc = upload_axes(ctx, d, false))) return rc;
  for (int d = 0; d < 3; ++d)
    if ((rc = upload_axes(ctx, d, true))) return rc;
  return MRL_OK;
}

extern "C" int mrl_dist_partition(const mrl_context *ctx, int64_t *ng, int64_t *yc, int64_t *yb, int64_t *xc, int64_t *xb) {
  if (!ctx || !ctx->dist) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_partition: the domain is not decomposed");
  if (ctx->pencil) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_partition: slab decomposition only (pencils: mrl_dist_bounds)");
  for (int d = 0; d < 3 && ng; ++d) ng[d] = ctx->gn[d];
  for (int r = 0; r < ctx->nranks; ++r) {
    if (yc) yc[r] = ctx->ycount[r];
    if (yb) yb[r] = ctx->ybegin[r];
    if (xc) xc[r] = ctx->xcount[r];
    if (xb) xb[r] = ctx->xbegin[r];
  }
  return MRL_OK;
}

extern "C" int mrl_dist_bounds(const mrl_context *ctx, int rank, int64_t *rb, int64_t *re, int64_t *kb, int64_t *ke) {
  if (!ctx || !ctx->dim) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_bounds: domain not set");
  if (rank < 0 || rank >= ctx->nranks) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_bounds: rank %d outside [0, %d)", rank, ctx->nranks);
  int64_t b[2][3] = {{0, 0, 0}, {0, 0, 0}}, e[2][3];
  for (int d = 0; d < 3; ++d) {
    e[0][d] = ctx->gn[d];
    e[1][d] = ctx->gn[d];
  }
  if (!ctx->dist) {
    e[1][ctx->dim - 1] = ctx->gn[ctx->dim - 1] / 2 + 1;
  } else if (!ctx->pencil) {
    b[0][1] = ctx->ybegin[rank];
    e[0][1] = ctx->ybegin[rank] + ctx->ycount[rank];
    b[1][0] = ctx->xbegin[rank];
    e[1][0] = ctx->xbegin[rank] + ctx->xcount[rank];
    e[1][ctx->dim - 1] = ctx->gn[ctx->dim - 1] / 2 + 1;
  } else {
    const int iy = rank % ctx->py, iz = rank / ctx->py;
    b[0][1] = ctx->ybegin[iy];
    e[0][1] = ctx->ybegin[iy] + ctx->ycount[iy];
    b[0][2] = ctx->zbegin[iz];
    e[0][2] = ctx->zbegin[iz] + ctx->zcount[iz];
    b[1][0] = ctx->xbegin[iy];
    e[1][0] = ctx->xbegin[iy] + ctx->xcount[iy];
    b[1][1] = ctx->y2begin[iz];
    e[1][1] = ctx->y2begin[iz] + ctx->y2count[iz];
  }
  for (int d = 0; d < 3; ++d) {
    if (rb) rb[d] = b[0][d];
    if (re) re[d] = e[0][d];
    if (kb) kb[d] = b[1][d];
    if (ke) ke[d] = e[1][d];
  }
  return MRL_OK;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" int mrl_dist_create(mrl_context *ctx, mrl_dist **out) {
  if (!ctx || !out || !ctx->dist) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_create: needs a context with mrl_domain_set_dist / mrl_domain_set_pencil");
  if (ctx->nranks > 32) return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_dist_create: at most 32 ranks");
  CK(cudaSetDevice(ctx->device));
  mrl_dist *d = new mrl_dist();
  d->ctx = ctx;
  const size_t esz = ctx->precision == MRL_F64 ? 16 : 8;
  size_t bytes[A_COUNT] = {0, 0, 0, 0, 256 * sizeof(unsigned long long)};
  size_t work = 0;
  if (!ctx->pencil) {
    d->ncz = ctx->dim == 3 ? ctx->gn[2] / 2 + 1 : 1;
    bytes[A_R] = (size_t)ctx->nxl * ctx->gn[1] * d->ncz * esz;
    bytes[A_B] = (size_t)ctx->gn[0] * ctx->nyl * d->ncz * esz;
    work = std::max(bytes[A_R], bytes[A_B]);
  } else {
    const size_t nzl = ctx->n[2], ny2 = ctx->nr[1];
    bytes[A_R] = bytes[A_B2] = (size_t)ctx->nxl * ctx->gn[1] * nzl * esz;
    bytes[A_R2] = (size_t)ctx->nxl * ny2 * ctx->gn[2] * esz;
    bytes[A_B] = (size_t)ctx->gn[0] * ctx->nyl * nzl * esz;
    work = std::max(bytes[A_B], bytes[A_R2]);
  }
  size_t total = 0;
  for (int a = 0; a < A_COUNT; ++a) {
    d->off[a] = (long long)total;
    total += align256(bytes[a]);
  }
  cudaError_t e = cudaMalloc(&d->shared, total);
  if (e == cudaSuccess) e = cudaMalloc(&d->W, work ? work : 256);
  if (e == cudaSuccess) e = cudaMemset((char *)d->shared + d->off[A_FLAGS], 0, bytes[A_FLAGS]);
  if (e != cudaSuccess) {
    mrl_dist_destroy(d);
    return mrl_fail(MRL_ERR_CUDA, "mrl_dist_create: allocation failed: %s", cudaGetErrorString(e));
  }
  if (ctx->nranks == 1) {  // nothing to import
    d->peer.assign(1, (char *)d->shared);
    d->peer_off.assign(1, std::vector<long long>(d->off, d->off + 8));
    unsigned long long f = (unsigned long long)d->shared + d->off[A_FLAGS];
    CK(cudaMalloc(&d->flag_tab, sizeof f));
    CK(cudaMemcpy(d->flag_tab, &f, sizeof f, cudaMemcpyHostToDevice));
    d->imported = true;
  }
  *out = d;
  return MRL_OK;
}

extern "C" int mrl_dist_destroy(mrl_dist *d) {
  if (!d) return MRL_OK;
  mrl_quiesce(d->ctx);
  for (void *q : d->opened) cudaIpcCloseMemHandle(q);
  cudaFree(d->shared);
  cudaFree(d->W);
  cudaFree(d->flag_tab);
  delete d;
  return MRL_OK;
}

// MRL_DIST_IPC_BYTES = the CUDA IPC handle of the shared allocation (64 bytes) + the 8 area offsets
extern "C" int mrl_dist_ipc_export(mrl_dist *d, void *handles) {
  if (!d || !handles) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_ipc_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64 && MRL_DIST_IPC_BYTES == 64 + 8 * sizeof(long long), "IPC record size");
  CK(cudaSetDevice(d->ctx->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, d->shared));
  memcpy(handles, &h, sizeof h);
  memcpy((char *)handles + sizeof h, d->off, sizeof d->off);
  return MRL_OK;
}

extern "C" int mrl_dist_ipc_import(mrl_dist *d, const void *all) {
  if (!d || !all) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_ipc_import: bad arguments");
  if (d->imported) return MRL_OK;
  mrl_context *ctx = d->ctx;
  CK(cudaSetDevice(ctx->device));
  const int P = ctx->nranks;
  std::vector<unsigned long long> ftab(P);
  d->peer.assign(P, nullptr);
  d->peer_off.assign(P, std::vector<long long>(8, 0));
  for (int s = 0; s < P; ++s) {
    const char *rec = (const char *)all + (size_t)s * MRL_DIST_IPC_BYTES;
    memcpy(d->peer_off[s].data(), rec + 64, 8 * sizeof(long long));
    if (s == ctx->rank) {
      d->peer[s] = (char *)d->shared;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, rec, sizeof h);
      void *p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      d->opened.push_back(p);
      d->peer[s] = (char *)p;
    }
    ftab[s] = (unsigned long long)d->area(s, A_FLAGS);
  }
  CK(cudaMalloc(&d->flag_tab, P * sizeof(unsigned long long)));
  CK(cudaMemcpy(d->flag_tab, ftab.data(), P * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  d->imported = true;
  return MRL_OK;
}

static int dist_barrier(mrl_dist *d) {
  mrl_context *ctx = d->ctx;
  if (ctx->nranks == 1) return MRL_OK;
  ctx->launches++;
  // host processes reach a transform with seconds of skew (NVRTC, initial conditions on the CPU): ~30 s before the trap
  CK(launch_slab_barrier(ctx->lc(), d->flag_tab, 0, ctx->rank, ctx->nranks, ++d->epoch, 60000000000ll));
  return MRL_OK;
}

// ------------------------------------------------------------------------------------------------------------- slab
template <class T> static int slab_forward_one(mrl_dist *d, const T *in, cx<T> *out) {
  mrl_context *ctx = d->ctx;
  const int dim = ctx->dim, P = ctx->nranks, me = ctx->rank;
  const long long nx = ctx->gn[0], ny = ctx->gn[1], nyl = ctx->nyl, nxl = ctx->nxl, ncz = d->ncz;
  const size_t esz = sizeof(cx<T>);
  cx<T> *A = (cx<T> *)d->W, *R = (cx<T> *)d->area(me, A_R);
  int rc;
  // local axes: z r2c (3-D) / real -> complex (2-D), then x  (fft2 over {0, 2} resp. fft over 0, DomainAction.C:878-879)
  if (dim == 3) {
    if ((rc = mrl_pass_zfwd(ctx, in, A, nx * nyl, ctx->gn[2]))) return rc;
  } else {
    ctx->launches++;
    CK(launch_real_to_complex<T>(ctx->lc(), in, A, nx * nyl));
  }
  if ((rc = mrl_pass_strided(ctx, A, A, (int)nx, nyl * ncz, 1, 0))) return rc;
  // exchange (the MPI_Isend / MPI_Recv loop of DomainAction.C:886-927): x-block s of the local slab goes to rank s,
  // landing at y = ybegin[me].. of its [nxl][ny][ncz] staging - the torch::cat along y (:935) is the address arithmetic
  if ((rc = dist_barrier(d))) return rc;  // every rank has consumed its staging of the previous transform
  for (int i = 0; i < P; ++i) {
    const int s = (me + i) % P;
    CK(cudaMemcpy2DAsync(d->area(s, A_R) + (size_t)ctx->ybegin[me] * ncz * esz, (size_t)ny * ncz * esz, A + (size_t)ctx->xbegin[s] * nyl * ncz,
                         (size_t)nyl * ncz * esz, (size_t)nyl * ncz * esz, (size_t)ctx->xcount[s], cudaMemcpyDefault, ctx->stream));
  }
  if ((rc = dist_barrier(d))) return rc;  // every block has landed
  // y (DomainAction.C:937)
  if (dim == 3) return mrl_pass_strided(ctx, R, out, (int)ny, ncz, nxl, 0);
  if ((rc = mrl_pass_strided(ctx, R, R, (int)ny, 1, nxl, 0))) return rc;
  // keep the half ky <= ny/2: the x-slice of the serial rfft2 layout
  const long long nyc = ny / 2 + 1;
  CK(cudaMemcpy2DAsync(out, (size_t)nyc * esz, R, (size_t)ny * esz, (size_t)nyc * esz, (size_t)nxl, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRL_OK;
}

template <class T> static int slab_inverse_one(mrl_dist *d, const cx<T> *in, T *out, double scale) {
  mrl_context *ctx = d->ctx;
  const int dim = ctx->dim, P = ctx->nranks, me = ctx->rank;
  const long long nx = ctx->gn[0], ny = ctx->gn[1], nyl = ctx->nyl, nxl = ctx->nxl, ncz = d->ncz;
  const size_t esz = sizeof(cx<T>);
  cx<T> *W = (cx<T> *)d->W, *B = (cx<T> *)d->area(me, A_B);
  const double N = (double)ctx->gn[0] * ctx->gn[1] * ctx->gn[2];
  int rc;
  // y inverse (DomainAction.C:951)
  if (dim == 3) {
    if ((rc = mrl_pass_strided(ctx, in, W, (int)ny, ncz, nxl, 1))) return rc;
  } else {
    ctx->launches++;
    CK(launch_expand_half<T>(ctx->lc(), in, W, nxl, (int)ny));
    if ((rc = mrl_pass_strided(ctx, W, W, (int)ny, 1, nxl, 1))) return rc;
  }
  // exchange back (DomainAction.C:961-1002): y-block s goes to rank s, landing at x = xbegin[me].. of its [nx][nyl_s][ncz]
  if ((rc = dist_barrier(d))) return rc;
  for (int i = 0; i < P; ++i) {
    const int s = (me + i) % P;
    const long long nys = ctx->ycount[s];
    CK(cudaMemcpy2DAsync(d->area(s, A_B) + (size_t)ctx->xbegin[me] * nys * ncz * esz, (size_t)nys * ncz * esz, W + (size_t)ctx->ybegin[s] * ncz,
                         (size_t)ny * ncz * esz, (size_t)nys * ncz * esz, (size_t)nxl, cudaMemcpyDefault, ctx->stream));
  }
  if ((rc = dist_barrier(d))) return rc;
  // x inverse, then z c2r (3-D) / real part (2-D), normalised by 1/N (:1013-1016)
  if ((rc = mrl_pass_strided(ctx, B, B, (int)nx, nyl * ncz, 1, 1))) return rc;
  if (dim == 3) return mrl_pass_zinv(ctx, B, out, nx * nyl, ctx->gn[2], scale / N);
  ctx->launches++;
  CK(launch_complex_real_scale<T>(ctx->lc(), (const cx<T> *)B, out, nx * nyl, (T)(scale / N)));
  return MRL_OK;
}

// ----------------------------------------------------------------------------------------------------------- pencil
// fftPencil (DomainAction.C:1022-1034): rfft along x, stage 1 (x <-> y inside the z group, pencilStage1Forward
// :1106-1178), fft along y, stage 2 (y <-> z inside the x group, pencilStage2Forward :1181-1256), fft along z.
// The x transform runs as a full complex pass on the real field; the half spectrum kx <= nx/2 is the first nx/2+1
// rows of the result, which is all that travels.
template <class T> static int pencil_forward_one(mrl_dist *d, const T *in, cx<T> *out) {
  mrl_context *ctx = d->ctx;
  const int me = ctx->rank, Py = ctx->py, Pz = ctx->pz, iy = me % Py, iz = me / Py, gbase = iz * Py;
  const long long nx = ctx->gn[0], ny = ctx->gn[1], nz = ctx->gn[2], nyl = ctx->nyl, nzl = ctx->n[2], nxl = ctx->nxl, ny2 = ctx->nr[1];
  const size_t esz = sizeof(cx<T>);
  cx<T> *A = (cx<T> *)d->W, *R1 = (cx<T> *)d->area(me, A_R), *R2 = (cx<T> *)d->area(me, A_R2);
  int rc;
  ctx->launches++;
  CK(launch_real_to_complex<T>(ctx->lc(), in, A, nx * nyl * nzl));
  if ((rc = mrl_pass_strided(ctx, A, A, (int)nx, nyl * nzl, 1, 0))) return rc;
  if ((rc = dist_barrier(d))) return rc;
  for (int i = 0; i < Py; ++i) {
    const int p = (iy + i) % Py, s = gbase + p;  // kx block p goes to the rank that owns it, at y = ybegin[iy]..
    CK(cudaMemcpy2DAsync(d->area(s, A_R) + (size_t)ctx->ybegin[iy] * nzl * esz, (size_t)ny * nzl * esz, A + (size_t)ctx->xbegin[p] * nyl * nzl,
                         (size_t)nyl * nzl * esz, (size_t)nyl * nzl * esz, (size_t)ctx->xcount[p], cudaMemcpyDefault, ctx->stream));
  }
  if ((rc = dist_barrier(d))) return rc;
  if ((rc = mrl_pass_strided(ctx, R1, R1, (int)ny, nzl, nxl, 0))) return rc;
  for (int i = 0; i < Pz; ++i) {
    const int q = (iz + i) % Pz, s = q * Py + iy;  // ky block q goes to the rank that owns it, at z = zbegin[iz]..
    const long long nyq = ctx->y2count[q];
    ctx->launches++;
    CK(launch_copy3d<T>(ctx->lc(), (cx<T> *)d->area(s, A_R2) + ctx->zbegin[iz], nyq * nz, nz, R1 + ctx->y2begin[q] * nzl, ny * nzl, nzl, nxl, nyq, nzl));
  }
  if ((rc = dist_barrier(d))) return rc;
  (void)ny2;
  return mrl_pass_strided(ctx, R2, out, (int)nz, 1, nxl * ny2, 0);
}

// ifftPencil (:1037-1047): the stages in reverse; the inverse real transform along x expands the half spectrum by
// the Hermitian symmetry of the partially transformed array, A[nx - k][y][z] = conj A[k][y][z], which is local.
template <class T> static int pencil_inverse_one(mrl_dist *d, const cx<T> *in, T *out, double scale) {
  mrl_context *ctx = d->ctx;
  const int me = ctx->rank, Py = ctx->py, Pz = ctx->pz, iy = me % Py, iz = me / Py, gbase = iz * Py;
  const long long nx = ctx->gn[0], ny = ctx->gn[1], nz = ctx->gn[2], nyl = ctx->nyl, nzl = ctx->n[2], nxl = ctx->nxl, ny2 = ctx->nr[1];
  const size_t esz = sizeof(cx<T>);
  cx<T> *W = (cx<T> *)d->W, *B2 = (cx<T> *)d->area(me, A_B2), *B1 = (cx<T> *)d->area(me, A_B);
  const double N = (double)nx * ny * nz;
  int rc;
  if ((rc = mrl_pass_strided(ctx, in, W, (int)nz, 1, nxl * ny2, 1))) return rc;
  if ((rc = dist_barrier(d))) return rc;
  for (int i = 0; i < Pz; ++i) {
    const int q = (iz + i) % Pz, s = q * Py + iy;  // z block q goes back to the rank that owns it, at y = y2begin[iz]..
    const long long nzq = ctx->zcount[q];
    ctx->launches++;
    CK(launch_copy3d<T>(ctx->lc(), (cx<T> *)d->area(s, A_B2) + ctx->y2begin[iz] * nzq, ny * nzq, nzq, W + ctx->zbegin[q], ny2 * nz, nz, nxl, ny2, nzq));
  }
  if ((rc = dist_barrier(d))) return rc;
  if ((rc = mrl_pass_strided(ctx, B2, B2, (int)ny, nzl, nxl, 1))) return rc;
  for (int i = 0; i < Py; ++i) {
    const int p = (iy + i) % Py, s = gbase + p;  // y block p goes back to the rank that owns it, at kx = xbegin[iy]..
    const long long nyp = ctx->ycount[p];
    CK(cudaMemcpy2DAsync(d->area(s, A_B) + (size_t)ctx->xbegin[iy] * nyp * nzl * esz, (size_t)nyp * nzl * esz, B2 + (size_t)ctx->ybegin[p] * nzl,
                         (size_t)ny * nzl * esz, (size_t)nyp * nzl * esz, (size_t)nxl, cudaMemcpyDefault, ctx->stream));
  }
  if ((rc = dist_barrier(d))) return rc;
  ctx->launches++;
  CK(launch_hermitian_rows<T>(ctx->lc(), B1, (int)nx, nyl * nzl));
  if ((rc = mrl_pass_strided(ctx, B1, B1, (int)nx, nyl * nzl, 1, 1))) return rc;
  ctx->launches++;
  CK(launch_complex_real_scale<T>(ctx->lc(), (const cx<T> *)B1, out, nx * nyl * nzl, (T)(scale / N)));
  return MRL_OK;
}

extern "C" int mrl_dist_rfftn(mrl_dist *d, const void *in, void *out, int batch) {
  if (!d || !in || !out || batch < 1) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_rfftn: bad arguments");
  if (!d->imported) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_rfftn: peers' buffers not imported (mrl_dist_ipc_import)");
  mrl_context *ctx = d->ctx;
  CK(cudaSetDevice(ctx->device));
  const long long rl = (long long)ctx->n[0] * ctx->n[1] * ctx->n[2], kl = (long long)ctx->nr[0] * ctx->nr[1] * ctx->nr[2];
  const bool f64 = ctx->precision == MRL_F64;
  for (int b = 0; b < batch; ++b) {
    int rc;
    if (ctx->pencil)
      rc = f64 ? pencil_forward_one<double>(d, (const double *)in + b * rl, (cx<double> *)out + b * kl)
               : pencil_forward_one<float>(d, (const float *)in + b * rl, (cx<float> *)out + b * kl);
    else
      rc = f64 ? slab_forward_one<double>(d, (const double *)in + b * rl, (cx<double> *)out + b * kl)
               : slab_forward_one<float>(d, (const float *)in + b * rl, (cx<float> *)out + b * kl);
    if (rc) return rc;
  }
  return MRL_OK;
}

extern "C" int mrl_dist_irfftn(mrl_dist *d, const void *in, void *out, int batch) { return mrl_dist_irfftn_scaled(d, in, out, batch, 1.0); }

int mrl_dist_irfftn_scaled(mrl_dist *d, const void *in, void *out, int batch, double scale) {
  if (!d || !in || !out || batch < 1) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_irfftn: bad arguments");
  if (!d->imported) return mrl_fail(MRL_ERR_INVALID, "mrl_dist_irfftn: peers' buffers not imported (mrl_dist_ipc_import)");
  mrl_context *ctx = d->ctx;
  CK(cudaSetDevice(ctx->device));
  const long long rl = (long long)ctx->n[0] * ctx->n[1] * ctx->n[2], kl = (long long)ctx->nr[0] * ctx->nr[1] * ctx->nr[2];
  const bool f64 = ctx->precision == MRL_F64;
  for (int b = 0; b < batch; ++b) {
    int rc;
    if (ctx->pencil)
      rc = f64 ? pencil_inverse_one<double>(d, (const cx<double> *)in + b * kl, (double *)out + b * rl, scale)
               : pencil_inverse_one<float>(d, (const cx<float> *)in + b * kl, (float *)out + b * rl, scale);
    else
      rc = f64 ? slab_inverse_one<double>(d, (const cx<double> *)in + b * kl, (double *)out + b * rl, scale)
               : slab_inverse_one<float>(d, (const cx<float> *)in + b * kl, (float *)out + b * rl, scale);
    if (rc) return rc;
  }
  return MRL_OK;
}
