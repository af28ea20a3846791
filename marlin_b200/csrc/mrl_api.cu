// marlin_b200 - implementation of the C ABI declared in include/marlin_b200.h.
// Host logic only; all device work is in k_*.cu (hand-written sm_100a kernels).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/marlin_b200.h"
#include "mrl_internal.h"
#include "mrl_passes_slab.cuh"

using namespace mrl;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
int mrl_fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
extern "C" const char *mrl_last_error(void) { return g_err.c_str(); }
extern "C" const char *mrl_version(void) { return "marlin_b200 0.1 (sm_100a)"; }
extern "C" int mrl_device_count(int *count) {
  if (!count) return mrl_fail(MRL_ERR_INVALID, "mrl_device_count: null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
  *count = n;
  return MRL_OK;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return mrl_fail(MRL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                      __FILE__, __LINE__);                                               \
  } while (0)
#define CKL(ctx, call) \
  do {                 \
    (ctx)->launches++; \
    CK(call);          \
  } while (0)

// ------------------------------------------------------------------------------ axes
// Bit-compatible with ATen: linspace (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp:
// symmetric start+step*i / end-step*(steps-1-i) halves) and fft_fftfreq / fft_rfftfreq
// (arange * (1/(n*d))), followed by the reference's `freq * 2.0 * pi`
// (src/actions/DomainAction.C:241-293).
extern "C" int mrl_axis_values(int64_t n, double min, double max, int reciprocal, int half, double *out) {
  if (n < 1 || !out || !(max > min)) return mrl_fail(MRL_ERR_INVALID, "mrl_axis_values: bad arguments");
  const double dx = (max - min) / (double)n;
  if (!reciprocal) {
    const double start = min + dx / 2.0, end = max - dx / 2.0;
    if (n == 1) {
      out[0] = start;
      return MRL_OK;
    }
    const double step = (end - start) / (double)(n - 1);
    const int64_t halfway = n / 2;
    // single rounding per value (ATen's vectorised kernel is compiled with FMA contraction)
    for (int64_t i = 0; i < n; ++i)
      out[i] = (i < halfway) ? std::fma(step, (double)i, start) : std::fma(-step, (double)(n - i - 1), end);
    return MRL_OK;
  }
  const double pi = 3.14159265358979323846;
  const double inv = 1.0 / ((double)n * dx);
  if (half) {
    for (int64_t k = 0; k <= n / 2; ++k) out[k] = ((double)k * inv) * 2.0 * pi;
  } else {
    for (int64_t k = 0; k < n; ++k) {
      const int64_t ks = (k < (n + 1) / 2) ? k : k - n;
      out[k] = ((double)ks * inv) * 2.0 * pi;
    }
  }
  return MRL_OK;
}

// ------------------------------------------------------------------------------ context
FFTPlanDev mrl::make_fft_plan(int n) {
  FFTPlanDev p;
  memset(&p, 0, sizeof p);
  p.n = n;
  int m = n;
  const int pref[] = {8, 4, 2, 3, 5};
  for (int r : pref)
    while (m > 1 && m % r == 0) {
      p.radix[p.nstages++] = r;
      m /= r;
    }
  for (int f = 7; m > 1; f += 2)
    while (m % f == 0) {
      p.radix[p.nstages++] = f;
      m /= f;
    }
  return p;
}

template <class T> static int upload_vec(mrl_context *ctx, const std::vector<double> &h, void **dev) {
  std::vector<T> t(h.begin(), h.end());
  CK(cudaMalloc(dev, t.size() * sizeof(T)));
  CK(cudaMemcpy(*dev, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MRL_OK;
}

int mrl_context::twiddles(int n, const void **out) {
  auto it = tw.find(n);
  if (it != tw.end()) {
    *out = it->second;
    return MRL_OK;
  }
  const long double PI = 3.141592653589793238462643383279502884L;
  std::vector<double> h(2 * (size_t)n);
  for (int k = 0; k < n; ++k) {
    const long double a = -2 * PI * (long double)k / (long double)n;
    h[2 * k] = (double)cosl(a);
    h[2 * k + 1] = (double)sinl(a);
  }
  void *d = nullptr;
  int rc = precision == MRL_F64 ? upload_vec<double>(this, h, &d) : upload_vec<float>(this, h, &d);
  if (rc) return rc;
  tw[n] = d;
  *out = d;
  return MRL_OK;
}

int mrl_context::scratch(size_t bytes, void **out) {
  if (bytes > scratch_bytes) {
    if (scratch_ptr) {
      CK(cudaStreamSynchronize(stream));
      CK(cudaFree(scratch_ptr));
      scratch_ptr = nullptr;
      scratch_bytes = 0;
    }
    CK(cudaMalloc(&scratch_ptr, bytes));
    scratch_bytes = bytes;
  }
  *out = scratch_ptr;
  return MRL_OK;
}

// Live contexts: handles derived from a context (plans, expressions) may outlive it when a host
// language finalises objects out of order; their destroy functions must not touch a freed context.
static std::mutex g_live_mu;
static std::set<const mrl_context *> g_live;
void mrl_quiesce(const mrl_context *ctx) {
  bool alive;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    alive = g_live.count(ctx) != 0;
  }
  if (alive) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  } else {
    cudaDeviceSynchronize();
  }
}

extern "C" int mrl_create(int device, int precision, mrl_context **out) {
  if (!out || (precision != MRL_F64 && precision != MRL_F32)) return mrl_fail(MRL_ERR_INVALID, "mrl_create: bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return mrl_fail(MRL_ERR_NO_DEVICE, "mrl_create: CUDA device %d not available (%d devices%s%s); marlin_b200 has no CPU path",
                    device, ndev, e != cudaSuccess ? ", " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return mrl_fail(MRL_ERR_NO_DEVICE, "mrl_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                    prop.major, prop.minor);
  mrl_context *c = new mrl_context();
  c->device = device;
  c->precision = precision;
  c->stream = 0;
  c->sm_count = prop.multiProcessorCount;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.insert(c);
  }
  *out = c;
  return MRL_OK;
}

extern "C" int mrl_destroy(mrl_context *ctx) {
  if (!ctx) return MRL_OK;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    if (!g_live.erase(ctx)) return mrl_fail(MRL_ERR_INVALID, "mrl_destroy: not a live context");
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->tw) cudaFree(kv.second);
  for (int d = 0; d < 3; ++d) {
    cudaFree(ctx->axis_dev[d]);
    cudaFree(ctx->kaxis_dev[d]);
  }
  cudaFree(ctx->scratch_ptr);
  cudaFree(ctx->reduce_dev);
  if (ctx->reduce_host) cudaFreeHost(ctx->reduce_host);
  if (ctx->owned_stream) cudaStreamDestroy(ctx->owned_stream);
  if (ctx->s_in) {
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->s_out);
    cudaStreamDestroy(ctx->s_in);
    cudaStreamDestroy(ctx->s_out);
    cudaEventDestroy(ctx->ev_in);
    cudaEventDestroy(ctx->ev_compute);
    for (auto &kv : ctx->dl_done) cudaEventDestroy(kv.second);
  }
  delete ctx;
  return MRL_OK;
}

extern "C" int mrl_set_stream(mrl_context *ctx, void *s) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  ctx->stream = (cudaStream_t)s;
  return MRL_OK;
}
extern "C" int mrl_own_stream(mrl_context *ctx) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->owned_stream) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamCreateWithFlags(&ctx->owned_stream, cudaStreamNonBlocking));
  }
  ctx->stream = ctx->owned_stream;
  return MRL_OK;
}
extern "C" int mrl_synchronize(mrl_context *ctx) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaStreamSynchronize(ctx->stream));
  return MRL_OK;
}
extern "C" int mrl_precision_of(const mrl_context *ctx) { return ctx ? ctx->precision : MRL_ERR_INVALID; }
extern "C" int mrl_launch_count(const mrl_context *ctx, int64_t *count) {
  if (!ctx || !count) return mrl_fail(MRL_ERR_INVALID, "null argument");
  *count = ctx->launches;
  return MRL_OK;
}

// ------------------------------------------------------------------------------ domain
extern "C" int mrl_domain_set(mrl_context *ctx, int dim, const int64_t *n, const double *mn, const double *mx) {
  if (!ctx || dim < 1 || dim > 3 || !n || !mn || !mx) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set: bad arguments");
  CK(cudaSetDevice(ctx->device));
  for (int d = 0; d < dim; ++d) {
    if (n[d] < 1 || n[d] > (1 << 24)) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set: bad grid size");
    // reference: "Max coordinate must be larger than the min coordinate in every dimension"
    if (!(mx[d] > mn[d]))
      return mrl_fail(MRL_ERR_INVALID, "Max coordinate must be larger than the min coordinate in every dimension");
  }
  ctx->dim = dim;
  ctx->dist = false;
  ctx->pencil = false;
  ctx->rank = 0;
  ctx->nranks = 1;
  ctx->nyl = ctx->nxl = ctx->y0 = ctx->x0 = 0;
  for (int d = 0; d < 3; ++d) {
    ctx->gn[d] = d < dim ? (int)n[d] : 1;
    ctx->n[d] = d < dim ? (int)n[d] : 1;
    ctx->min[d] = d < dim ? mn[d] : 0.0;
    ctx->max[d] = d < dim ? mx[d] : 1.0;
    const bool half = (d == dim - 1);
    ctx->nr[d] = d < dim ? (half ? ctx->n[d] / 2 + 1 : ctx->n[d]) : 1;
    ctx->axis_h[d].assign(1, 0.0);
    ctx->kaxis_h[d].assign(1, 0.0);
    if (d < dim) {
      ctx->axis_h[d].resize(ctx->n[d]);
      ctx->kaxis_h[d].resize(ctx->nr[d]);
      int rc = mrl_axis_values(ctx->n[d], mn[d], mx[d], 0, 0, ctx->axis_h[d].data());
      if (rc) return rc;
      rc = mrl_axis_values(ctx->n[d], mn[d], mx[d], 1, half ? 1 : 0, ctx->kaxis_h[d].data());
      if (rc) return rc;
    }
    cudaFree(ctx->axis_dev[d]);
    cudaFree(ctx->kaxis_dev[d]);
    ctx->axis_dev[d] = ctx->kaxis_dev[d] = nullptr;
    int rc;
    if (ctx->precision == MRL_F64) {
      if ((rc = upload_vec<double>(ctx, ctx->axis_h[d], &ctx->axis_dev[d]))) return rc;
      if ((rc = upload_vec<double>(ctx, ctx->kaxis_h[d], &ctx->kaxis_dev[d]))) return rc;
    } else {
      if ((rc = upload_vec<float>(ctx, ctx->axis_h[d], &ctx->axis_dev[d]))) return rc;
      if ((rc = upload_vec<float>(ctx, ctx->kaxis_h[d], &ctx->kaxis_dev[d]))) return rc;
    }
  }
  return MRL_OK;
}

extern "C" int mrl_domain_shape(const mrl_context *ctx, int64_t *rs, int64_t *ks) {
  if (!ctx || !ctx->dim) return mrl_fail(MRL_ERR_INVALID, "domain not set");
  for (int d = 0; d < 3; ++d) {
    if (rs) rs[d] = ctx->n[d];
    if (ks) ks[d] = ctx->nr[d];
  }
  return MRL_OK;
}

extern "C" int mrl_domain_axis(const mrl_context *ctx, int d, int reciprocal, double *out) {
  if (!ctx || !ctx->dim || d < 0 || d > 2 || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_axis: bad arguments");
  const std::vector<double> &v = reciprocal ? ctx->kaxis_h[d] : ctx->axis_h[d];
  memcpy(out, v.data(), v.size() * sizeof(double));
  return MRL_OK;
}

// ------------------------------------------------------------------------------ memory
extern "C" int mrl_malloc(mrl_context *ctx, size_t bytes, void **dev) {
  if (!ctx || !dev) return mrl_fail(MRL_ERR_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(dev, bytes ? bytes : 1));
  return MRL_OK;
}
extern "C" int mrl_free(mrl_context *ctx, void *dev) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaStreamSynchronize(ctx->stream));
  auto it = ctx->dl_done.find(dev);
  if (it != ctx->dl_done.end()) {
    CK(cudaEventSynchronize(it->second));
    cudaEventDestroy(it->second);
    ctx->dl_done.erase(it);
  }
  CK(cudaFree(dev));
  return MRL_OK;
}
extern "C" int mrl_memset(mrl_context *ctx, void *dev, int value, size_t bytes) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaMemsetAsync(dev, value, bytes, ctx->stream));
  return MRL_OK;
}
extern "C" int mrl_upload(mrl_context *ctx, void *dev, const void *host, size_t bytes) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return MRL_OK;
}
extern "C" int mrl_download(mrl_context *ctx, void *host, const void *dev, size_t bytes) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MRL_OK;
}
static int staged_init(mrl_context *ctx) {
  if (ctx->s_in) return MRL_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
  return MRL_OK;
}
extern "C" int mrl_upload_staged(mrl_context *ctx, void *dev, const void *host, size_t bytes) {
  if (!ctx || !dev || !host) return mrl_fail(MRL_ERR_INVALID, "mrl_upload_staged: bad arguments");
  int rc = staged_init(ctx);
  if (rc) return rc;
  auto it = ctx->dl_done.find(dev);
  if (it != ctx->dl_done.end()) CK(cudaStreamWaitEvent(ctx->s_in, it->second, 0));  // WAR against its last download
  CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->s_in));
  CK(cudaEventRecord(ctx->ev_in, ctx->s_in));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in, 0));
  return MRL_OK;
}
extern "C" int mrl_download_staged(mrl_context *ctx, void *host, const void *dev, size_t bytes) {
  if (!ctx || !dev || !host) return mrl_fail(MRL_ERR_INVALID, "mrl_download_staged: bad arguments");
  int rc = staged_init(ctx);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev_compute, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_compute, 0));
  CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->s_out));
  cudaEvent_t &ev = ctx->dl_done[dev];
  if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaEventRecord(ev, ctx->s_out));
  return MRL_OK;
}
extern "C" int mrl_staged_wait(mrl_context *ctx) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  if (!ctx->s_in) return MRL_OK;
  CK(cudaStreamSynchronize(ctx->s_in));
  CK(cudaStreamSynchronize(ctx->s_out));
  return MRL_OK;
}
extern "C" int mrl_copy(mrl_context *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return mrl_fail(MRL_ERR_INVALID, "null context");
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return MRL_OK;
}

// ------------------------------------------------------------------------------ FFT
// Complex pass along `axis` of batched [batch][n0][n1][ncp] spectra (ncp = pitch of the last axis;
// ncp > n_last/2+1 only with the TMA kernels, whose padding columns are carried along as zeros).
template <class T> static int strided_axis(mrl_context *ctx, const cx<T> *in, cx<T> *out, int nfields, long long field_stride,
                                           int axis, int batch, int inverse, int ncp = 0) {
  const int dim = ctx->dim;
  const int nc = ctx->nr[dim - 1];
  if (!ncp) ncp = nc;
  long long ncols = ncp;
  for (int b = axis + 1; b < dim - 1; ++b) ncols *= ctx->n[b];
  long long nouter = batch;
  for (int b = 0; b < axis; ++b) nouter *= ctx->n[b];
  StridedIO<T> io;
  memset(&io, 0, sizeof io);
  if (nfields > 4) return mrl_fail(MRL_ERR_INVALID, "too many fields");
  for (int f = 0; f < nfields; ++f) {
    io.in[f] = in + f * field_stride;
    io.out[f] = out + f * field_stride;
  }
  io.nfields = nfields;
  io.n = ctx->n[axis];
  io.ncols = (int)ncols;
  io.nouter = (int)nouter;
  io.pitch = ncols;
  io.outer_stride = (long long)io.n * ncols;
  io.scale = T(1);
  io.inverse = inverse;
  const void *tw;
  int rc = ctx->twiddles(io.n, &tw);
  if (rc) return rc;
  ctx->launches++;
  cudaError_t te = launch_strided_tma<T>(ctx->lc(), io, (const cx<T> *)tw, io.n);
  if (te == cudaErrorNotSupported && ncp == nc) te = launch_strided<T>(ctx->lc(), io, (const cx<T> *)tw, make_fft_plan(io.n));
  CK(te);
  return MRL_OK;
}

template <class T>
static cudaError_t zinv_dispatch(mrl_context *ctx, const cx<T> *in, T *out, long long rows, int n, T scale, const cx<T> *tw,
                                 int ncp = 0, const ZinvDot<T> *dot = nullptr) {
  if (!ncp) ncp = n / 2 + 1;
  if (dot) *dot->count = 0;  // stays 0 when the pass that runs cannot carry the inner product
  cudaError_t e = launch_zinv_pairs_tma<T>(ctx->lc(), in, ncp, out, rows, n, scale, tw, dot);
  if (e == cudaErrorNotSupported && ncp == n / 2 + 1) e = launch_zinv_pairs<T>(ctx->lc(), in, out, rows, n, scale, tw, make_fft_plan(n));
  return e;
}

// stop_axis: strided axes below it are left untransformed (mechanics: the x pass is fused with the
// Green projection)
template <class T> static int rfftn_impl(mrl_context *ctx, const T *in, cx<T> *out, int batch, int ncp = 0, int stop_axis = 0) {
  const int dim = ctx->dim, nl = ctx->n[dim - 1];
  const int nc = nl / 2 + 1;
  if (!ncp) ncp = nc;
  long long rows = batch;
  for (int d = 0; d < dim - 1; ++d) rows *= ctx->n[d];
  const void *tw;
  int rc = ctx->twiddles(nl, &tw);
  if (rc) return rc;
  ctx->launches++;
  cudaError_t e = launch_zfwd_pairs_tma<T>(ctx->lc(), in, out, rows, nl, ncp, (const cx<T> *)tw);
  if (e == cudaErrorNotSupported && ncp == nc) e = launch_zfwd_pairs<T>(ctx->lc(), in, out, rows, nl, (const cx<T> *)tw, make_fft_plan(nl));
  CK(e);
  for (int a = dim - 2; a >= stop_axis; --a)
    if ((rc = strided_axis<T>(ctx, out, out, 1, 0, a, batch, 0, ncp))) return rc;
  return MRL_OK;
}

// src == nullptr: transform `work` in place (it is destroyed); otherwise src is preserved and
// `work` receives the partially transformed spectra
template <class T>
static int irfftn_impl2(mrl_context *ctx, const cx<T> *src, cx<T> *work, T *out, int batch, int ncp, double scale, int start_axis = 0,
                        const ZinvDot<T> *dot = nullptr) {
  const int dim = ctx->dim, nl = ctx->n[dim - 1];
  long long rows = batch;
  for (int d = 0; d < dim - 1; ++d) rows *= ctx->n[d];
  const cx<T> *cur = src ? src : work;
  int rc;
  for (int a = start_axis; a <= dim - 2; ++a) {
    if ((rc = strided_axis<T>(ctx, cur, work, 1, 0, a, batch, 1, ncp))) return rc;
    cur = work;
  }
  const void *tw;
  if ((rc = ctx->twiddles(nl, &tw))) return rc;
  CKL(ctx, zinv_dispatch<T>(ctx, cur, out, rows, nl, (T)scale, (const cx<T> *)tw, ncp, dot));
  return MRL_OK;
}

template <class T> static int irfftn_impl(mrl_context *ctx, const cx<T> *in, T *out, int batch) {
  const int dim = ctx->dim;
  long long total = batch;
  for (int d = 0; d < dim; ++d) total *= ctx->nr[d];
  double N = 1;
  for (int d = 0; d < dim; ++d) N *= ctx->n[d];
  cx<T> *sc = nullptr;
  if (dim > 1) {
    void *s;
    int rc = ctx->scratch((size_t)total * sizeof(cx<T>), &s);
    if (rc) return rc;
    sc = (cx<T> *)s;
  }
  return irfftn_impl2<T>(ctx, in, sc, out, batch, 0, 1.0 / N);
}

// single passes on explicit shapes: see mrl_internal.h
template <class T> static int pass_strided_t(mrl_context *ctx, const cx<T> *in, cx<T> *out, int n, long long ncols, long long nouter, int inverse) {
  StridedIO<T> io;
  memset(&io, 0, sizeof io);
  io.in[0] = in;
  io.out[0] = out;
  io.nfields = 1;
  io.n = n;
  io.ncols = (int)ncols;
  io.nouter = (int)nouter;
  io.pitch = ncols;
  io.outer_stride = (long long)n * ncols;
  io.scale = T(1);
  io.inverse = inverse;
  const void *tw;
  int rc = ctx->twiddles(n, &tw);
  if (rc) return rc;
  ctx->launches++;
  cudaError_t te = launch_strided_tma<T>(ctx->lc(), io, (const cx<T> *)tw, n);
  if (te == cudaErrorNotSupported) te = launch_strided<T>(ctx->lc(), io, (const cx<T> *)tw, make_fft_plan(n));
  CK(te);
  return MRL_OK;
}
int mrl_pass_strided(mrl_context *ctx, const void *in, void *out, int n, long long ncols, long long nouter, int inverse) {
  return ctx->precision == MRL_F64 ? pass_strided_t<double>(ctx, (const cx<double> *)in, (cx<double> *)out, n, ncols, nouter, inverse)
                                   : pass_strided_t<float>(ctx, (const cx<float> *)in, (cx<float> *)out, n, ncols, nouter, inverse);
}
template <class T> static int pass_zfwd_t(mrl_context *ctx, const T *in, cx<T> *out, long long rows, int n) {
  const void *tw;
  int rc = ctx->twiddles(n, &tw);
  if (rc) return rc;
  ctx->launches++;
  cudaError_t e = launch_zfwd_pairs_tma<T>(ctx->lc(), in, out, rows, n, n / 2 + 1, (const cx<T> *)tw);
  if (e == cudaErrorNotSupported) e = launch_zfwd_pairs<T>(ctx->lc(), in, out, rows, n, (const cx<T> *)tw, make_fft_plan(n));
  CK(e);
  return MRL_OK;
}
int mrl_pass_zfwd(mrl_context *ctx, const void *in, void *out, long long rows, int n) {
  return ctx->precision == MRL_F64 ? pass_zfwd_t<double>(ctx, (const double *)in, (cx<double> *)out, rows, n)
                                   : pass_zfwd_t<float>(ctx, (const float *)in, (cx<float> *)out, rows, n);
}
int mrl_pass_zinv(mrl_context *ctx, const void *in, void *out, long long rows, int n, double scale) {
  const void *tw;
  int rc = ctx->twiddles(n, &tw);
  if (rc) return rc;
  if (ctx->precision == MRL_F64) CKL(ctx, zinv_dispatch<double>(ctx, (const cx<double> *)in, (double *)out, rows, n, scale, (const cx<double> *)tw));
  else CKL(ctx, zinv_dispatch<float>(ctx, (const cx<float> *)in, (float *)out, rows, n, (float)scale, (const cx<float> *)tw));
  return MRL_OK;
}

// Internal batched transforms on padded layouts (mechanics): see mrl_internal.h
int mrl_fftb_pitch(const mrl_context *ctx) {
  const int dim = ctx->dim, nc = ctx->nr[dim - 1];
  bool all = tma_enabled() && !getenv("MRL_NOPAD");
  for (int a = 0; a < dim; ++a) all = all && (ctx->n[a] == 128 || ctx->n[a] == 256 || ctx->n[a] == 512 || ctx->n[a] == 1024);
  if (!all) return nc;
  const int per128 = ctx->precision == MRL_F64 ? 8 : 16;
  return (nc + per128 - 1) / per128 * per128;
}
int mrl_fftb_forward(mrl_context *ctx, const void *in, void *out, int batch, int ncp, int first_axis) {
  return ctx->precision == MRL_F64 ? rfftn_impl<double>(ctx, (const double *)in, (cx<double> *)out, batch, ncp, first_axis)
                                   : rfftn_impl<float>(ctx, (const float *)in, (cx<float> *)out, batch, ncp, first_axis);
}
int mrl_fftb_strided(mrl_context *ctx, void *spec, int batch, int ncp, int axis, int inverse) {
  return ctx->precision == MRL_F64
             ? strided_axis<double>(ctx, (const cx<double> *)spec, (cx<double> *)spec, 1, 0, axis, batch, inverse, ncp)
             : strided_axis<float>(ctx, (const cx<float> *)spec, (cx<float> *)spec, 1, 0, axis, batch, inverse, ncp);
}
int mrl_fftb_inverse(mrl_context *ctx, void *work, void *out, int batch, int ncp, double scale, int first_axis, const void *dot_with,
                     double *dot_partials, int dot_capacity, int *dot_count) {
  if (dot_count) *dot_count = 0;
  if (ctx->precision == MRL_F64) {
    const ZinvDot<double> dot{(const double *)dot_with, dot_partials, dot_capacity, dot_count};
    return irfftn_impl2<double>(ctx, nullptr, (cx<double> *)work, (double *)out, batch, ncp, scale, first_axis, dot_with ? &dot : nullptr);
  }
  const ZinvDot<float> dot{(const float *)dot_with, dot_partials, dot_capacity, dot_count};
  return irfftn_impl2<float>(ctx, nullptr, (cx<float> *)work, (float *)out, batch, ncp, scale, first_axis, dot_with ? &dot : nullptr);
}

extern "C" int mrl_rfftn(mrl_context *ctx, const void *in, void *out, int batch) {
  if (!ctx || !ctx->dim || !in || !out || batch < 1) return mrl_fail(MRL_ERR_INVALID, "mrl_rfftn: bad arguments / domain not set");
  if (ctx->dist) return mrl_fail(MRL_ERR_INVALID, "mrl_rfftn: the domain is slab-decomposed (mrl_domain_set_dist): use mrl_dist_rfftn");
  CK(cudaSetDevice(ctx->device));
  return ctx->precision == MRL_F64 ? rfftn_impl<double>(ctx, (const double *)in, (cx<double> *)out, batch)
                                   : rfftn_impl<float>(ctx, (const float *)in, (cx<float> *)out, batch);
}
extern "C" int mrl_irfftn(mrl_context *ctx, const void *in, void *out, int batch) {
  if (!ctx || !ctx->dim || !in || !out || batch < 1) return mrl_fail(MRL_ERR_INVALID, "mrl_irfftn: bad arguments / domain not set");
  if (ctx->dist) return mrl_fail(MRL_ERR_INVALID, "mrl_irfftn: the domain is slab-decomposed (mrl_domain_set_dist): use mrl_dist_irfftn");
  CK(cudaSetDevice(ctx->device));
  return ctx->precision == MRL_F64 ? irfftn_impl<double>(ctx, (const cx<double> *)in, (double *)out, batch)
                                   : irfftn_impl<float>(ctx, (const cx<float> *)in, (float *)out, batch);
}

// ------------------------------------------------------------------------------ pointwise
extern "C" int mrl_kfactor(mrl_context *ctx, int kind, double factor, void *out) {
  if (!ctx || !ctx->dim || !out || kind < 0 || kind > 1) return mrl_fail(MRL_ERR_INVALID, "mrl_kfactor: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (ctx->precision == MRL_F64)
    CKL(ctx, launch_kfactor<double>(ctx->lc(), (double *)out, (const double *)ctx->kaxis_dev[0], (const double *)ctx->kaxis_dev[1],
                                    (const double *)ctx->kaxis_dev[2], ctx->nr[0], ctx->nr[1], ctx->nr[2], kind, factor));
  else
    CKL(ctx, launch_kfactor<float>(ctx->lc(), (float *)out, (const float *)ctx->kaxis_dev[0], (const float *)ctx->kaxis_dev[1],
                                   (const float *)ctx->kaxis_dev[2], ctx->nr[0], ctx->nr[1], ctx->nr[2], kind, (float)factor));
  return MRL_OK;
}

extern "C" int mrl_mul_real_complex(mrl_context *ctx, const void *a, const void *b, void *out) {
  if (!ctx || !ctx->dim || !a || !b || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_mul_real_complex: bad arguments");
  const long long total = ctx->rtotal();
  if (ctx->precision == MRL_F64)
    CKL(ctx, launch_mul_rc<double>(ctx->lc(), (cx<double> *)out, (const double *)a, (const cx<double> *)b, total));
  else
    CKL(ctx, launch_mul_rc<float>(ctx->lc(), (cx<float> *)out, (const float *)a, (const cx<float> *)b, total));
  return MRL_OK;
}

extern "C" int mrl_ab_update(mrl_context *ctx, void *ubar, const void *cbar, const void *N, const void *L, double dt,
                             const double *beta, int nold, const void *const *Nold) {
  if (!ctx || !ctx->dim || !ubar || !cbar || !N || !beta || nold < 0 || nold > 4 || (nold && !Nold))
    return mrl_fail(MRL_ERR_INVALID, "mrl_ab_update: bad arguments");
  const long long total = ctx->rtotal();
  if (ctx->precision == MRL_F64) {
    double bo[4] = {0, 0, 0, 0};
    for (int i = 0; i < nold; ++i) bo[i] = dt * beta[i + 1];
    CKL(ctx, launch_ab_update<double>(ctx->lc(), (cx<double> *)ubar, (const cx<double> *)cbar, (const cx<double> *)N,
                                      (const double *)L, dt, dt * beta[0], nold, (const cx<double> *const *)Nold, bo, total));
  } else {
    float bo[4] = {0, 0, 0, 0};
    for (int i = 0; i < nold; ++i) bo[i] = (float)(dt * beta[i + 1]);
    CKL(ctx, launch_ab_update<float>(ctx->lc(), (cx<float> *)ubar, (const cx<float> *)cbar, (const cx<float> *)N,
                                     (const float *)L, (float)dt, (float)(dt * beta[0]), nold, (const cx<float> *const *)Nold,
                                     bo, total));
  }
  return MRL_OK;
}

// ------------------------------------------------------------------------------ reductions
namespace mrl {
template <class T>
cudaError_t launch_coupled_solve(const LaunchCtx &lc, int nvar, const void *const *L, const void *const *rhs, void *const *out, double dt,
                                 int drop_imag, long long total);
}
extern "C" int mrl_coupled_solve(mrl_context *ctx, int nvar, const void *const *L, const void *const *rhs, void *const *out, double dt,
                                 int drop_imag) {
  if (!ctx || !ctx->dim || !L || !rhs || !out || nvar < 1) return mrl_fail(MRL_ERR_INVALID, "mrl_coupled_solve: bad arguments");
  if (nvar > 6) return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_coupled_solve: at most 6 coupled variables (got %d)", nvar);
  for (int i = 0; i < nvar; ++i)
    if (!rhs[i] || !out[i]) return mrl_fail(MRL_ERR_INVALID, "mrl_coupled_solve: null right-hand side / output %d", i);
  const long long total = ctx->rtotal();
  if (ctx->precision == MRL_F64) CKL(ctx, launch_coupled_solve<double>(ctx->lc(), nvar, L, rhs, out, dt, drop_imag, total));
  else CKL(ctx, launch_coupled_solve<float>(ctx->lc(), nvar, L, rhs, out, dt, drop_imag, total));
  return MRL_OK;
}

namespace mrl {
template <class T>
cudaError_t launch_broyden(const LaunchCtx &lc, int nvar, int update, void *M, const void *const *a, const void *const *b,
                           const void *const *c, void *const *o0, void *const *o1, long long total);
}
static int broyden_check(mrl_context *ctx, int nvar, const void *M, const char *who) {
  if (!ctx || !ctx->dim || !M || nvar < 1) return mrl_fail(MRL_ERR_INVALID, "%s: bad arguments", who);
  if (nvar > 6) return mrl_fail(MRL_ERR_UNSUPPORTED, "%s: at most 6 coupled variables (got %d)", who, nvar);
  return MRL_OK;
}
extern "C" int mrl_broyden_step(mrl_context *ctx, int nvar, const void *M, const void *const *R, const void *const *u, void *const *sk,
                                void *const *unew) {
  int rc = broyden_check(ctx, nvar, M, "mrl_broyden_step");
  if (rc) return rc;
  if (!R || !u || !sk || !unew) return mrl_fail(MRL_ERR_INVALID, "mrl_broyden_step: null argument");
  const long long total = ctx->rtotal();
  if (ctx->precision == MRL_F64) CKL(ctx, launch_broyden<double>(ctx->lc(), nvar, 0, const_cast<void *>(M), R, u, nullptr, sk, unew, total));
  else CKL(ctx, launch_broyden<float>(ctx->lc(), nvar, 0, const_cast<void *>(M), R, u, nullptr, sk, unew, total));
  return MRL_OK;
}
extern "C" int mrl_broyden_update(mrl_context *ctx, int nvar, void *M, const void *const *sk, const void *const *R, const void *const *Rnew) {
  int rc = broyden_check(ctx, nvar, M, "mrl_broyden_update");
  if (rc) return rc;
  if (!sk || !R || !Rnew) return mrl_fail(MRL_ERR_INVALID, "mrl_broyden_update: null argument");
  const long long total = ctx->rtotal();
  if (ctx->precision == MRL_F64) CKL(ctx, launch_broyden<double>(ctx->lc(), nvar, 1, M, sk, R, Rnew, nullptr, nullptr, total));
  else CKL(ctx, launch_broyden<float>(ctx->lc(), nvar, 1, M, sk, R, Rnew, nullptr, nullptr, total));
  return MRL_OK;
}

extern "C" int mrl_reduce(mrl_context *ctx, int op, const void *in, int64_t count, double *host_out) {
  if (!ctx || !in || !host_out || count < 1 || op < 0 || op > 3) return mrl_fail(MRL_ERR_INVALID, "mrl_reduce: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int nblk = ctx->sm_count * 4;
  if (!ctx->reduce_dev) {
    CK(cudaMalloc(&ctx->reduce_dev, nblk * sizeof(double)));
    CK(cudaMallocHost(&ctx->reduce_host, nblk * sizeof(double)));
  }
  ctx->launches++;
  cudaError_t e = ctx->precision == MRL_F64
                      ? launch_reduce<double>(ctx->lc(), op, (const double *)in, count, (double *)ctx->reduce_dev, nblk)
                      : launch_reduce<float>(ctx->lc(), op, (const float *)in, count, (double *)ctx->reduce_dev, nblk);
  CK(e);
  CK(cudaMemcpyAsync(ctx->reduce_host, ctx->reduce_dev, nblk * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const double *p = (const double *)ctx->reduce_host;
  double r = p[0];
  for (int i = 1; i < nblk; ++i) {
    if (op == MRL_MIN) r = p[i] < r ? p[i] : r;
    else if (op == MRL_MAX) r = p[i] > r ? p[i] : r;
    else r += p[i];
  }
  *host_out = r;
  return MRL_OK;
}

// ------------------------------------------------------------------------------ fused split plan
// Sizes that have a TMA-pipelined configuration in k_tma.cu
static bool tma_size(int n) { return n == 128 || n == 256 || n == 512 || n == 1024; }
extern "C" int mrl_split_plan_create(mrl_context *ctx, const mrl_split_desc *d, mrl_split_plan **out) {
  if (!ctx || !ctx->dim || !d || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_split_plan_create: bad arguments");
  if (ctx->dim < 2)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "fused split plan needs dim >= 2 (1-D problems use the un-fused operators)");
  if (ctx->dist)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "the single-GPU fused split plan does not run on a slab-decomposed domain (un-fused operators over "
                                         "mrl_dist_rfftn / mrl_dist_irfftn, or mrl_slab_plan_create_peer)");
  if (d->nonlin_kind != MRL_NONLIN_DOUBLE_WELL && d->nonlin_kind != MRL_NONLIN_EXPR)
    return mrl_fail(MRL_ERR_INVALID, "unknown nonlin_kind");
  if (d->nonlin_kind == MRL_NONLIN_EXPR && !d->nonlin_expr) return mrl_fail(MRL_ERR_INVALID, "nonlin_expr is NULL");
  if (!d->M_closed_form && !d->M_real_dev) return mrl_fail(MRL_ERR_INVALID, "M_real_dev is NULL");
  if (d->has_L && !d->L_closed_form && !d->L_real_dev) return mrl_fail(MRL_ERR_INVALID, "L_real_dev is NULL");
  if (d->history < 0 || d->history > 4) return mrl_fail(MRL_ERR_INVALID, "history must be in [0,4]");
  CK(cudaSetDevice(ctx->device));
  mrl_split_plan *p = new mrl_split_plan();
  p->ctx = ctx;
  p->desc = *d;
  const size_t esz = ctx->precision == MRL_F64 ? 16 : 8;
  // Work spectra are private to the plan, so their rows are padded to a multiple of 128 bytes
  // whenever every axis runs on the TMA kernels: aligned 128-byte row segments are what lets
  // the strided passes stream at full HBM rate (measured 4.4 -> 6.1 TB/s at 512^3).
  const int nc = ctx->nr[ctx->dim - 1];
  bool all_tma = tma_enabled() && !getenv("MRL_NOPAD");
  for (int a = 0; a < ctx->dim; ++a) all_tma = all_tma && tma_size(ctx->n[a]);
  const int per128 = (int)(128 / esz);
  p->ncp = all_tma ? (nc + per128 - 1) / per128 * per128 : nc;
  size_t rowsn = 1;
  for (int a = 0; a < ctx->dim - 1; ++a) rowsn *= ctx->n[a];
  const size_t sc = rowsn * p->ncp * esz;
  cudaError_t e = cudaMalloc(&p->A, 2 * sc);
  if (e == cudaSuccess) p->B = (char *)p->A + sc;
  if (e == cudaSuccess) e = cudaMemsetAsync(p->A, 0, 2 * sc, ctx->stream);  // padding columns stay zero
  for (int i = 0; e == cudaSuccess && d->history > 0 && i < d->history + 1; ++i) {
    void *q = nullptr;
    e = cudaMalloc(&q, sc);
    if (e == cudaSuccess) e = cudaMemsetAsync(q, 0, sc, ctx->stream);
    if (e == cudaSuccess) p->ring.push_back(q);
  }
  if (e != cudaSuccess) {
    mrl_split_plan_destroy(p);
    return mrl_fail(MRL_ERR_CUDA, "split plan allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return MRL_OK;
}

extern "C" int mrl_split_plan_destroy(mrl_split_plan *p) {
  if (!p) return MRL_OK;
  mrl_quiesce(p->ctx);
  if (p->graph) cudaGraphExecDestroy(p->graph);
  cudaFree(p->A);
  for (void *q : p->ring) cudaFree(q);
  delete p;
  return MRL_OK;
}

extern "C" int mrl_split_launches_per_substep(const mrl_split_plan *p) {
  if (!p) return MRL_ERR_INVALID;
  return p->ctx->dim == 3 ? 5 : 3;
}

extern "C" int mrl_split_clear_states(mrl_split_plan *p) {
  if (!p) return mrl_fail(MRL_ERR_INVALID, "null plan");
  p->stored = 0;
  return MRL_OK;
}

extern "C" int mrl_split_advance_state(mrl_split_plan *p, int *stored) {
  if (!p) return mrl_fail(MRL_ERR_INVALID, "null plan");
  const int H = p->desc.history;
  if (H > 0) {
    if (p->stored < H) p->stored++;
    p->cur = (p->cur + 1) % (H + 1);
  }
  if (stored) *stored = p->stored;
  return MRL_OK;
}

#define PASS_MARK()                                                    \
  do {                                                                 \
    if (ev) CK(cudaEventRecord(ev[nev++], ctx->stream));               \
  } while (0)
template <class T>
static int split_substep_impl(mrl_split_plan *p, T *c, double dt, const double *beta, int nold, cudaEvent_t *ev = nullptr,
                              int phases = 3) {
  mrl_context *ctx = p->ctx;
  int nev = 0;
  const mrl_split_desc &d = p->desc;
  const int dim = ctx->dim;
  const int nl = ctx->n[dim - 1], nc = ctx->nr[dim - 1];
  const int ncp = p->ncp;            // row pitch of the work spectra (nc, or nc padded to 128 bytes)
  const bool padded = ncp != nc;     // padded plans run on the TMA kernels only
  cx<T> *A = (cx<T> *)p->A, *B = (cx<T> *)p->B;
  long long rows = 1;
  for (int a = 0; a < dim - 1; ++a) rows *= ctx->n[a];
  const long long ftotal = rows * ncp;  // elements of one work spectrum
  const void *twl, *tw0;
  int rc;
  if ((rc = ctx->twiddles(nl, &twl))) return rc;
  if ((rc = ctx->twiddles(ctx->n[0], &tw0))) return rc;

  // strided pass along axis 1 of the [n0][n1][ncp] work arrays (3-D only)
  auto ypass = [&](int nfields, int inverse) -> int {
    StridedIO<T> sio;
    memset(&sio, 0, sizeof sio);
    for (int f = 0; f < nfields; ++f) {
      sio.in[f] = A + f * ftotal;
      sio.out[f] = A + f * ftotal;
    }
    sio.nfields = nfields;
    sio.n = ctx->n[1];
    sio.ncols = ncp;
    sio.nvalid = nc;
    sio.nouter = ctx->n[0];
    sio.pitch = ncp;
    sio.outer_stride = (long long)sio.n * ncp;
    sio.scale = T(1);
    sio.inverse = inverse;
    const void *tw;
    int r = ctx->twiddles(sio.n, &tw);
    if (r) return r;
    ctx->launches++;
    cudaError_t te = launch_strided_tma<T>(ctx->lc(), sio, (const cx<T> *)tw, sio.n);
    if (te == cudaErrorNotSupported && !padded) te = launch_strided<T>(ctx->lc(), sio, (const cx<T> *)tw, make_fft_plan(sio.n));
    CK(te);
    return MRL_OK;
  };

  PASS_MARK();
  if (phases & 1) {
  // P1: last-axis r2c of (c + i F(c))
  if (d.nonlin_kind == MRL_NONLIN_DOUBLE_WELL) {
    NonlinDesc nlz{0, {d.nonlin_params[0], d.nonlin_params[1], d.nonlin_params[2], 0}};
    ctx->launches++;
    cudaError_t te = launch_zfwd_nonlin_tma<T>(ctx->lc(), c, (T *)d.g_out_real_dev, A, B, rows, nl, ncp, nlz, (const cx<T> *)twl);
    if (te == cudaErrorNotSupported && !padded)
      te = launch_zfwd_nonlin<T>(ctx->lc(), c, (T *)d.g_out_real_dev, A, B, rows, nl, nlz, (const cx<T> *)twl, make_fft_plan(nl));
    CK(te);
  } else {
    if ((rc = mrl_expr_launch_zfwd(ctx, d.nonlin_expr, d.nonlin_var, d.nonlin_inputs_dev, p->time, c, d.g_out_real_dev, A, B,
                                   rows, nl, ncp)))
      return rc;
  }
  PASS_MARK();
  // P2: middle axis forward on both fields (3-D only)
  if (dim == 3) {
    if ((rc = ypass(2, 0))) return rc;
    PASS_MARK();
  }
  }
  if (!(phases & 2)) return MRL_OK;
  // P3: first axis forward on both + k-space update + first axis inverse
  FusedIO<T> io;
  memset(&io, 0, sizeof io);
  io.inC = A;
  io.inG = B;
  io.outU = A;
  io.n = ctx->n[0];
  io.ncols = dim == 3 ? ctx->n[1] * ncp : ncp;
  io.nouter = 1;
  io.pitch = io.ncols;
  io.outer_stride = 0;
  io.scale = T(1);
  SpectralUpdate<T> up;
  memset(&up, 0, sizeof up);
  up.kx = (const T *)ctx->kaxis_dev[0];
  up.ky = (const T *)ctx->kaxis_dev[1];
  up.kz = (const T *)ctx->kaxis_dev[2];
  up.kmode = dim == 3 ? MRL_KMODE_3D : MRL_KMODE_2D;
  up.nzc = ncp;
  up.nzv = nc;
  up.closed_M = d.M_closed_form;
  up.Mfac = (T)d.M_factor;
  up.Mbuf = (const T *)d.M_real_dev;
  up.has_L = d.has_L;
  up.closed_L = d.L_closed_form;
  up.Lfac = (T)d.L_factor;
  up.Lbuf = (const T *)d.L_real_dev;
  up.dt = (T)dt;
  up.b0 = (T)(dt * beta[0]);
  up.nold = nold;
  const int H = d.history;
  for (int i = 0; i < nold; ++i) {
    up.bold[i] = (T)(dt * beta[i + 1]);
    up.Nold[i] = (const cx<T> *)p->ring[((p->cur - 1 - i) % (H + 1) + (H + 1)) % (H + 1)];
  }
  up.Nout = H > 0 ? (cx<T> *)p->ring[p->cur] : nullptr;
  ctx->launches++;
  cudaError_t fe = launch_fused_tma<T>(ctx->lc(), io, up, (const cx<T> *)tw0, io.n);
  if (fe == cudaErrorNotSupported && !padded) fe = launch_fused<T>(ctx->lc(), io, up, (const cx<T> *)tw0, make_fft_plan(io.n));
  CK(fe);
  PASS_MARK();
  // P4: middle axis inverse
  if (dim == 3) {
    if ((rc = ypass(1, 1))) return rc;
    PASS_MARK();
  }
  // P5: last-axis c2r with the 1/N normalisation
  double N = 1;
  for (int a = 0; a < dim; ++a) N *= ctx->n[a];
  ctx->launches++;
  cudaError_t ze = launch_zinv_pairs_tma<T>(ctx->lc(), A, ncp, c, rows, nl, (T)(1.0 / N), (const cx<T> *)twl);
  if (ze == cudaErrorNotSupported && !padded)
    ze = launch_zinv_pairs<T>(ctx->lc(), A, c, rows, nl, (T)(1.0 / N), (const cx<T> *)twl, make_fft_plan(nl));
  CK(ze);
  PASS_MARK();
  return MRL_OK;
}

extern "C" int mrl_split_substep(mrl_split_plan *p, void *c, double dt, const double *beta, int nold) {
  if (!p || !c || !beta || nold < 0) return mrl_fail(MRL_ERR_INVALID, "mrl_split_substep: bad arguments");
  if (nold > p->stored)
    return mrl_fail(MRL_ERR_INVALID, "mrl_split_substep: %d old states requested, %d stored", nold, p->stored);
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? split_substep_impl<double>(p, (double *)c, dt, beta, nold)
                                      : split_substep_impl<float>(p, (float *)c, dt, beta, nold);
}

// `count` substeps, each followed by mrl_split_advance_state, with the same dt / beta / nold (the steady
// state of AdamsBashforthMoulton inside one MOOSE step).  On small grids a substep is a handful of
// microsecond-sized kernels and the launch gaps dominate: one period of the sequence (the ring of old
// nonlinear terms repeats after history+1 substeps) is captured as a CUDA graph and replayed.
extern "C" int mrl_split_substeps(mrl_split_plan *p, void *c, double dt, const double *beta, int nold, int count) {
  if (!p || !c || !beta || nold < 0 || count < 0) return mrl_fail(MRL_ERR_INVALID, "mrl_split_substeps: bad arguments");
  mrl_context *ctx = p->ctx;
  CK(cudaSetDevice(ctx->device));
  const int period = p->desc.history + 1;
  auto one = [&]() -> int {
    int rc = mrl_split_substep(p, c, dt, beta, nold);
    if (rc) return rc;
    return mrl_split_advance_state(p, nullptr);
  };
  static const bool use_graph = !getenv("MRL_NO_GRAPH");
  int done = 0;
  // graphs need a capturable (non-legacy) stream, a full ring, and enough substeps to amortise the capture
  if (use_graph && ctx->stream != nullptr && ctx->stream != cudaStreamLegacy && p->stored >= p->desc.history && nold <= p->stored &&
      count >= 4 * period) {
    mrl_split_plan::GraphKey key;
    key.c = c;
    key.dt = dt;
    key.nold = nold;
    key.cur = p->cur;
    for (int i = 0; i < 5 && i <= nold; ++i) key.beta[i] = beta[i];
    if (!p->graph || !(key == p->graph_key)) {
      // lazily initialised state (twiddles, kernel attributes, NVRTC modules) must exist before capturing
      for (int i = 0; i < period; ++i, ++done) {
        int rc = one();
        if (rc) return rc;
      }
      if (p->graph) {
        cudaGraphExecDestroy(p->graph);
        p->graph = nullptr;
      }
      const int64_t l0 = ctx->launches;
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      int rc = MRL_OK;
      for (int i = 0; i < period && !rc; ++i) rc = one();
      cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
      if (rc || ce != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        return rc ? rc : mrl_fail(MRL_ERR_CUDA, "mrl_split_substeps: stream capture failed: %s", cudaGetErrorString(ce));
      }
      p->graph_launches = ctx->launches - l0;
      ctx->launches = l0;  // nothing ran during the capture
      ce = cudaGraphInstantiate(&p->graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) return mrl_fail(MRL_ERR_CUDA, "mrl_split_substeps: cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
      p->graph_key = key;  // the ring position is back where the capture started (one full period)
    }
    for (; count - done >= period; done += period) {
      CK(cudaGraphLaunch(p->graph, ctx->stream));
      ctx->launches += p->graph_launches;
    }
  }
  for (; done < count; ++done) {
    int rc = one();
    if (rc) return rc;
  }
  return MRL_OK;
}

extern "C" int mrl_split_forward(mrl_split_plan *p, const void *c) {
  if (!p || !c) return mrl_fail(MRL_ERR_INVALID, "mrl_split_forward: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  const double b0[5] = {1, 0, 0, 0, 0};
  return p->ctx->precision == MRL_F64 ? split_substep_impl<double>(p, (double *)c, 0.0, b0, 0, nullptr, 1)
                                      : split_substep_impl<float>(p, (float *)c, 0.0, b0, 0, nullptr, 1);
}

extern "C" int mrl_split_finish(mrl_split_plan *p, void *c, double dt, const double *beta, int nold) {
  if (!p || !c || !beta || nold < 0) return mrl_fail(MRL_ERR_INVALID, "mrl_split_finish: bad arguments");
  if (nold > p->stored) return mrl_fail(MRL_ERR_INVALID, "mrl_split_finish: %d old states requested, %d stored", nold, p->stored);
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? split_substep_impl<double>(p, (double *)c, dt, beta, nold, nullptr, 2)
                                      : split_substep_impl<float>(p, (float *)c, dt, beta, nold, nullptr, 2);
}

extern "C" int mrl_split_set_time(mrl_split_plan *p, double t) {
  if (!p) return mrl_fail(MRL_ERR_INVALID, "null plan");
  p->time = t;
  return MRL_OK;
}

extern "C" int mrl_split_substep_timed(mrl_split_plan *p, void *c, double dt, const double *beta, int nold, float *pass_ms) {
  if (!p || !c || !beta || nold < 0 || !pass_ms) return mrl_fail(MRL_ERR_INVALID, "mrl_split_substep_timed: bad arguments");
  if (nold > p->stored) return mrl_fail(MRL_ERR_INVALID, "mrl_split_substep_timed: %d old states requested, %d stored", nold, p->stored);
  CK(cudaSetDevice(p->ctx->device));
  const int np = mrl_split_launches_per_substep(p);
  cudaEvent_t ev[8];
  for (int i = 0; i <= np; ++i) CK(cudaEventCreate(&ev[i]));
  int rc = p->ctx->precision == MRL_F64 ? split_substep_impl<double>(p, (double *)c, dt, beta, nold, ev)
                                        : split_substep_impl<float>(p, (float *)c, dt, beta, nold, ev);
  if (!rc) {
    CK(cudaStreamSynchronize(p->ctx->stream));
    for (int i = 0; i < np; ++i) CK(cudaEventElapsedTime(&pass_ms[i], ev[i], ev[i + 1]));
  }
  for (int i = 0; i <= np; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

// ------------------------------------------------------------------------------ slab decomposition
extern "C" int mrl_partition(int64_t total, int nranks, const double *weights, int64_t *count) {
  // partitionHepler, include/actions/DomainAction.h:249-280 (integer weights in the reference;
  // the arithmetic below is the same for integral values)
  if (total < 1 || nranks < 1 || !count) return mrl_fail(MRL_ERR_INVALID, "mrl_partition: bad arguments");
  long long remaining_w = 0;
  std::vector<long long> w(nranks, 1);
  for (int r = 0; r < nranks; ++r) {
    if (weights) w[r] = (long long)weights[r];
    remaining_w += w[r];
  }
  long long left = total;
  for (int r = 0; r < nranks; ++r) {
    if (remaining_w == 0) return mrl_fail(MRL_ERR_INVALID, "Internal partitioning error. remaining_total_weight 0 == 0");
    long long n = (left * w[r]) / remaining_w;
    if (n < 1) n = 1;  // assign at least one layer
    count[r] = n;
    remaining_w -= w[r];
    if (left < n) return mrl_fail(MRL_ERR_INVALID, "Internal partitioning error.");
    left -= n;
  }
  count[nranks - 1] += left;  // remainder to the last slice
  return MRL_OK;
}

extern "C" int mrl_domain_set_slab(mrl_context *ctx, int dim, const int64_t *n, const double *mn, const double *mx, int rank,
                                   int nranks) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return mrl_fail(MRL_ERR_INVALID, "mrl_domain_set_slab: bad arguments");
  if (dim != 3) return mrl_fail(MRL_ERR_UNSUPPORTED, "slab decomposition is implemented for dim = 3");
  int rc = mrl_domain_set(ctx, dim, n, mn, mx);
  if (rc) return rc;
  if (n[0] % nranks || n[1] % nranks)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "slab decomposition needs nx and ny divisible by the number of ranks (%lld, %lld, %d)",
                    (long long)n[0], (long long)n[1], nranks);
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->nyl = (int)(n[1] / nranks);
  ctx->nxl = (int)(n[0] / nranks);
  ctx->y0 = rank * ctx->nyl;
  ctx->x0 = rank * ctx->nxl;
  return MRL_OK;
}

extern "C" int mrl_domain_local(const mrl_context *ctx, int64_t *rs, int64_t *rb, int64_t *ks, int64_t *kb) {
  if (!ctx || !ctx->dim) return mrl_fail(MRL_ERR_INVALID, "domain not set");
  if (ctx->dist) {
    int64_t re[3], ke[3], b0[3], b1[3];
    int rc = mrl_dist_bounds(ctx, ctx->rank, rb ? rb : b0, re, kb ? kb : b1, ke);
    if (rc) return rc;
    for (int d = 0; d < 3; ++d) {
      if (rs) rs[d] = ctx->n[d];
      if (ks) ks[d] = ctx->nr[d];
    }
    return MRL_OK;
  }
  const bool slab = ctx->nranks > 1 || ctx->nyl;
  for (int d = 0; d < 3; ++d) {
    if (rs) rs[d] = (slab && d == 1) ? ctx->nyl : ctx->n[d];
    if (rb) rb[d] = (slab && d == 1) ? ctx->y0 : 0;
    if (ks) ks[d] = (slab && d == 0) ? ctx->nxl : ctx->nr[d];
    if (kb) kb[d] = (slab && d == 0) ? ctx->x0 : 0;
  }
  return MRL_OK;
}

static int slab_pitch(const mrl_context *ctx) {
  const int per128 = ctx->precision == MRL_F64 ? 8 : 16;
  return (ctx->nr[2] + per128 - 1) / per128 * per128;
}

extern "C" int mrl_slab_sizes(const mrl_context *ctx, int64_t *field, int64_t *chunk, int *pitch) {
  if (!ctx || ctx->dim != 3 || !ctx->nyl) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_sizes: slab domain not set");
  const int ncp = slab_pitch(ctx);
  if (field) *field = (int64_t)ctx->n[0] * ctx->nyl * ncp;
  if (chunk) *chunk = (int64_t)ctx->nxl * ctx->nyl * ncp;
  if (pitch) *pitch = ncp;
  return MRL_OK;
}

extern "C" int mrl_slab_plan_create(mrl_context *ctx, const mrl_split_desc *d, void *send_fwd, void *recv_fwd, void *send_bwd,
                                    mrl_slab_plan **out) {
  if (!ctx || !d || !send_fwd || !recv_fwd || !send_bwd || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_plan_create: bad arguments");
  if (ctx->dim != 3 || !ctx->nyl) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_plan_create: slab domain not set");
  if (ctx->pencil) return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan: the domain is pencil-decomposed");
  if (d->nonlin_kind != MRL_NONLIN_DOUBLE_WELL && d->nonlin_kind != MRL_NONLIN_EXPR) return mrl_fail(MRL_ERR_INVALID, "unknown nonlin_kind");
  if (d->nonlin_kind == MRL_NONLIN_EXPR && !d->nonlin_expr) return mrl_fail(MRL_ERR_INVALID, "nonlin_expr is NULL");
  if (!d->M_closed_form || (d->has_L && !d->L_closed_form))
    return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan: closed-form k-space factors only");
  // ctx->gn is the global grid (ctx->n[1] is the local extent on a context set up with mrl_domain_set_dist)
  if (!tma_enabled() || !tma_size(ctx->gn[0]) || !tma_size(ctx->gn[1]) || !tma_size(ctx->gn[2]) || ctx->nyl > 256)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan: every axis must be 128, 256, 512 or 1024 points (ny/P <= 256)");
  if (ctx->gn[0] != (long long)ctx->nxl * ctx->nranks || ctx->gn[1] != (long long)ctx->nyl * ctx->nranks)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan: the fused exchange needs equal slabs (nx = %d, ny = %d over %d ranks)", ctx->gn[0], ctx->gn[1],
                    ctx->nranks);
  for (int r = 0; ctx->dist && r < ctx->nranks; ++r)
    if (ctx->xcount[r] != ctx->nxl || ctx->ycount[r] != ctx->nyl) return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan: the fused exchange needs equal slabs");
  if (d->history < 0 || d->history > 4) return mrl_fail(MRL_ERR_INVALID, "history must be in [0,4]");
  CK(cudaSetDevice(ctx->device));
  mrl_slab_plan *p = new mrl_slab_plan();
  p->ctx = ctx;
  p->desc = *d;
  p->send_fwd = send_fwd;
  p->recv_fwd = recv_fwd;
  p->send_bwd = send_bwd;
  p->ncp = slab_pitch(ctx);
  p->field = (long long)ctx->n[0] * ctx->nyl * p->ncp;
  p->chunk = (long long)ctx->nxl * ctx->nyl * p->ncp;
  const size_t esz = ctx->precision == MRL_F64 ? 16 : 8;
  cudaError_t e = cudaMemsetAsync(send_fwd, 0, 2 * p->field * esz, ctx->stream);  // padding columns stay zero
  if (e == cudaSuccess) e = cudaMemsetAsync(recv_fwd, 0, 2 * p->field * esz, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(send_bwd, 0, p->field * esz, ctx->stream);
  for (int i = 0; e == cudaSuccess && d->history > 0 && i < d->history + 1; ++i) {
    void *q = nullptr;
    e = cudaMalloc(&q, p->field * esz);
    if (e == cudaSuccess) e = cudaMemsetAsync(q, 0, p->field * esz, ctx->stream);
    if (e == cudaSuccess) p->ring.push_back(q);
  }
  if (e != cudaSuccess) {
    mrl_slab_plan_destroy(p);
    return mrl_fail(MRL_ERR_CUDA, "slab plan allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return MRL_OK;
}

extern "C" int mrl_slab_plan_destroy(mrl_slab_plan *p) {
  if (!p) return MRL_OK;
  mrl_quiesce(p->ctx);
  if (p->s_aux) {
    cudaStreamSynchronize(p->s_aux);
    cudaStreamDestroy(p->s_aux);
    cudaEventDestroy(p->ev_begin);
    cudaEventDestroy(p->ev_done);
    for (auto &e : p->ev_chunk) cudaEventDestroy(e);
  }
  for (void *q : p->ring) cudaFree(q);
  for (void *q : p->opened) cudaIpcCloseMemHandle(q);
  cudaFree(p->peer_recv_tab);
  cudaFree(p->peer_send_tab);
  cudaFree(p->flag1_tab);
  cudaFree(p->flag2_tab);
  for (auto st : p->s_copy) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  for (auto e : p->ev_copy) cudaEventDestroy(e);
  if (p->owns) {
    cudaFree(p->send_fwd);
    cudaFree(p->recv_fwd);
    cudaFree(p->ret_stage);
    if (p->copy) cudaFree(p->send_bwd);
  }
  delete p;
  return MRL_OK;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int mrl_slab_plan_create_peer(mrl_context *ctx, const mrl_split_desc *d, mrl_slab_plan **out) {
  if (!ctx || !d || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_plan_create_peer: bad arguments");
  if (ctx->dim != 3 || !ctx->nyl) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_plan_create_peer: slab domain not set");
  CK(cudaSetDevice(ctx->device));
  const bool f64 = ctx->precision == MRL_F64;
  const size_t esz = f64 ? 16 : 8;
  const int ncp = slab_pitch(ctx);
  // the blocked staging layouts are cut into column blocks of the fused pass's tile width, the same for both axes
  const int wx = f64 ? fused_tma_tk<double>(ctx->n[0]) : fused_tma_tk<float>(ctx->n[0]);
  const int wy = f64 ? fused_tma_tk<double>(ctx->gn[1]) : fused_tma_tk<float>(ctx->gn[1]);
  if (!wx || wx != wy || ncp % wx)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan (peer mode): nx = %d and ny = %d need pipelined configurations of the same tile width", ctx->n[0],
                    ctx->gn[1]);
  if (ctx->nxl > 256) return mrl_fail(MRL_ERR_UNSUPPORTED, "slab plan (peer mode): nx/P <= 256");
  const int kb = ncp / wx;
  const size_t fbytes = (size_t)ctx->n[0] * ctx->nyl * ncp * esz;
  void *sf = nullptr, *rf = nullptr, *rs = nullptr;
  const char *xm = getenv("MRL_SLAB_EXCHANGE");
  if (xm && strcmp(xm, "copy") == 0) {
    // MRL_SLAB_EXCHANGE=copy: exchanges by the copy engines on the plain staged layouts instead of the bulk stores fused
    // into the passes (the default).  Measured on 2 B200 at 512^3 (profiles/r2t_*): 2.79 vs 2.23 ms per substep - the
    // strided (2-D) peer copies of a y-chunk do not overlap the passes, so the phase costs passes + copies
    const size_t flag_off = align256(2 * fbytes);
    cudaError_t e = cudaMalloc(&sf, 2 * fbytes);
    if (e == cudaSuccess) e = cudaMalloc(&rf, flag_off + 256 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&rs, fbytes);
    if (e == cudaSuccess) e = cudaMemset((char *)rf + flag_off, 0, 256 * sizeof(unsigned long long));
    int rc = e == cudaSuccess ? mrl_slab_plan_create(ctx, d, sf, rf, rs, out) : mrl_fail(MRL_ERR_CUDA, "slab plan allocation failed: %s", cudaGetErrorString(e));
    if (rc) {
      cudaFree(sf);
      cudaFree(rf);
      cudaFree(rs);
      return rc;
    }
    mrl_slab_plan *p = *out;
    p->owns = true;
    p->copy = true;
    p->flag_off = (long long)flag_off;
    if (const char *v = getenv("MRL_SLAB_CHUNKS")) p->chunks = atoi(v) > 0 ? atoi(v) : 1;
    if (const char *v = getenv("MRL_SLAB_YCHUNKS")) p->y_chunks = atoi(v) > 0 ? atoi(v) : 1;
    return MRL_OK;
  }
  // behind the spectra recv_fwd carries the barrier flags (one 8-byte slot per rank) and the arrival counters
  // counters1 / counters2 [nranks][kb] of the forward / return exchange
  const size_t flag_off = align256(2 * fbytes), c1_off = flag_off + 256 * sizeof(unsigned long long);
  const size_t cbytes = align256((size_t)ctx->nranks * kb * sizeof(unsigned long long)), c2_off = c1_off + cbytes;
  cudaError_t e = cudaMalloc(&sf, 2 * fbytes);
  if (e == cudaSuccess) e = cudaMalloc(&rf, c2_off + cbytes);
  if (e == cudaSuccess) e = cudaMalloc(&rs, fbytes);
  if (e == cudaSuccess) e = cudaMemset((char *)rf + flag_off, 0, c2_off + cbytes - flag_off);
  if (e == cudaSuccess) e = cudaMemset(rs, 0, fbytes);
  if (e != cudaSuccess) {
    cudaFree(sf);
    cudaFree(rf);
    cudaFree(rs);
    return mrl_fail(MRL_ERR_CUDA, "slab plan allocation failed: %s", cudaGetErrorString(e));
  }
  // send_bwd is not used in peer mode: pass recv_fwd as a placeholder for the argument check
  int rc = mrl_slab_plan_create(ctx, d, sf, rf, rf, out);
  if (rc) {
    cudaFree(sf);
    cudaFree(rf);
    cudaFree(rs);
    return rc;
  }
  mrl_slab_plan *p = *out;
  p->owns = true;
  p->send_bwd = nullptr;
  p->ret_stage = rs;
  p->tk = wx;
  p->kb = kb;
  p->flag_off = (long long)flag_off;
  p->c1_off = (long long)c1_off;
  p->c2_off = (long long)c2_off;
  if (const char *v = getenv("MRL_SLAB_CHUNKS")) p->chunks = atoi(v) > 0 ? atoi(v) : 1;
  if (const char *v = getenv("MRL_SLAB_XCTAS")) p->x_ctas = atoi(v) > 0 ? atoi(v) : 96;
  // MRL_SLAB_SYNC = flags: arrival counters per column block instead of the two barriers of a substep; the tiles then
  // run column-block major so that a phase can trail the one that feeds it
  const char *sy = getenv("MRL_SLAB_SYNC");
  p->sync_flags = sy && !strcmp(sy, "flags");
  p->kzb_major = p->sync_flags ? 1 : 0;
  if (const char *v = getenv("MRL_SLAB_KZB_MAJOR")) p->kzb_major = atoi(v) != 0;
  if (const char *v = getenv("MRL_SLAB_INV_CTAS")) p->inv_ctas = atoi(v) > 0 ? atoi(v) : 0;
  if (!p->sync_flags) p->inv_ctas = 0;
  return MRL_OK;
}

extern "C" int mrl_slab_barrier(mrl_slab_plan *p) {
  if (!p || !(p->peer || (p->copy && p->peer_recv_tab))) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_barrier: needs a peer-mode plan with imported handles");
  if (p->sync_flags) return MRL_OK;  // the phases synchronise through the arrival counters
  if (p->ctx->nranks > 32) return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_slab_barrier: at most 32 ranks");
  CK(cudaSetDevice(p->ctx->device));
  p->ctx->launches++;
  CK(launch_slab_barrier(p->ctx->lc(), p->peer_recv_tab, p->flag_off, p->ctx->rank, p->ctx->nranks, ++p->epoch));
  return MRL_OK;
}

extern "C" int mrl_slab_ipc_export(mrl_slab_plan *p, void *handles) {
  if (!p || !handles || !p->owns) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_ipc_export: needs a peer-mode plan");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CK(cudaSetDevice(p->ctx->device));
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], p->copy ? p->send_fwd : p->ret_stage));  // where the return exchange lands
  CK(cudaIpcGetMemHandle(&h[1], p->recv_fwd));
  memcpy(handles, h, sizeof h);
  return MRL_OK;
}

extern "C" int mrl_slab_ipc_import(mrl_slab_plan *p, const void *all) {
  if (!p || !all || !p->owns) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_ipc_import: needs a peer-mode plan");
  mrl_context *ctx = p->ctx;
  CK(cudaSetDevice(ctx->device));
  const int P = ctx->nranks;
  std::vector<unsigned long long> sendp(P), recvp(P), f1(P), f2(P);
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all;
  for (int s = 0; s < P; ++s) {
    if (s == ctx->rank) {
      sendp[s] = (unsigned long long)(p->copy ? p->send_fwd : p->ret_stage);
      recvp[s] = (unsigned long long)p->recv_fwd;
    } else {
      void *a = nullptr, *b = nullptr;
      CK(cudaIpcOpenMemHandle(&a, h[2 * s], cudaIpcMemLazyEnablePeerAccess));
      p->opened.push_back(a);
      CK(cudaIpcOpenMemHandle(&b, h[2 * s + 1], cudaIpcMemLazyEnablePeerAccess));
      p->opened.push_back(b);
      sendp[s] = (unsigned long long)a;
      recvp[s] = (unsigned long long)b;
    }
    f1[s] = recvp[s] + (unsigned long long)p->c1_off;
    f2[s] = recvp[s] + (unsigned long long)p->c2_off;
  }
  const size_t tb = P * sizeof(unsigned long long);
  CK(cudaMalloc(&p->peer_send_tab, tb));
  CK(cudaMalloc(&p->peer_recv_tab, tb));
  CK(cudaMalloc(&p->flag1_tab, tb));
  CK(cudaMalloc(&p->flag2_tab, tb));
  CK(cudaMemcpy(p->peer_send_tab, sendp.data(), tb, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->peer_recv_tab, recvp.data(), tb, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->flag1_tab, f1.data(), tb, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->flag2_tab, f2.data(), tb, cudaMemcpyHostToDevice));
  if (p->copy) {
    for (int s = 0; s < P; ++s) {
      p->h_peer_ret.push_back((char *)sendp[s]);
      p->h_peer_recv.push_back((char *)recvp[s]);
      cudaStream_t st;
      cudaEvent_t ev;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      p->s_copy.push_back(st);
      p->ev_copy.push_back(ev);
    }
    return MRL_OK;  // the passes stay the staged-layout ones: p->peer remains false
  }
  p->peer = true;
  return MRL_OK;
}

extern "C" int mrl_slab_advance_state(mrl_slab_plan *p, int *stored) {
  if (!p) return mrl_fail(MRL_ERR_INVALID, "null plan");
  const int H = p->desc.history;
  if (H > 0) {
    if (p->stored < H) p->stored++;
    p->cur = (p->cur + 1) % (H + 1);
  }
  if (stored) *stored = p->stored;
  return MRL_OK;
}

// x pass (axis 0) on the local [nx][nyl][ncp] slabs of send_fwd (staged mode: the exchange is the caller's all-to-all)
template <class T> static int slab_xpass(mrl_slab_plan *p, int nfields, int inverse) {
  mrl_context *ctx = p->ctx;
  cx<T> *A = (cx<T> *)p->send_fwd;
  StridedIO<T> sio;
  memset(&sio, 0, sizeof sio);
  for (int f = 0; f < nfields; ++f) sio.in[f] = sio.out[f] = A + f * p->field;
  sio.nfields = nfields;
  sio.n = ctx->n[0];
  sio.ncols = ctx->nyl * p->ncp;
  sio.nouter = 1;
  sio.pitch = (long long)ctx->nyl * p->ncp;
  sio.outer_stride = (long long)sio.n * sio.pitch;
  sio.scale = T(1);
  sio.inverse = inverse;
  const void *tw;
  int rc = ctx->twiddles(sio.n, &tw);
  if (rc) return rc;
  ctx->launches++;
  CK(launch_strided_tma<T>(ctx->lc(), sio, (const cx<T> *)tw, sio.n));
  return MRL_OK;
}

// peer mode: forward x pass of both spectra of the y-chunk [y0, y0 + ych), rows pushed into the peers' staging R
template <class T> static int slab_xfwd_peer(mrl_slab_plan *p, int y0, int ych, const LaunchCtx &lc) {
  mrl_context *ctx = p->ctx;
  SlabXIO<T> io;
  memset(&io, 0, sizeof io);
  io.n = ctx->n[0];
  io.nyl = ctx->nyl;
  io.kb = p->kb;
  io.nf = 2;
  io.nranks = ctx->nranks;
  io.rank = ctx->rank;
  io.nxl = ctx->nxl;
  io.y0 = y0;
  io.ych = ych;
  io.kzb_major = p->kzb_major;
  io.scale = T(1);
  io.peer_tab = (const unsigned long long *)p->peer_recv_tab;
  io.field = p->field;
  io.flag_tab = p->sync_flags ? (const unsigned long long *)p->flag1_tab : nullptr;
  const void *tw;
  int rc = ctx->twiddles(io.n, &tw);
  if (rc) return rc;
  ctx->launches++;
  CK(launch_slab_xfwd<T>(lc, (const cx<T> *)p->send_fwd, io, (const cx<T> *)tw, io.n));
  return MRL_OK;
}

// peer mode: inverse x pass from the return staging S into the natural slab (field 0 of send_fwd)
template <class T> static int slab_xinv_peer(mrl_slab_plan *p, const LaunchCtx &lc) {
  mrl_context *ctx = p->ctx;
  SlabXIO<T> io;
  memset(&io, 0, sizeof io);
  io.n = ctx->n[0];
  io.nyl = ctx->nyl;
  io.kb = p->kb;
  io.nf = 1;
  io.nranks = ctx->nranks;
  io.rank = ctx->rank;
  io.nxl = ctx->nxl;
  io.kzb_major = p->kzb_major;
  io.scale = T(1);
  io.out = (cx<T> *)p->send_fwd;
  io.out_pitch = (long long)ctx->nyl * p->ncp;
  if (p->sync_flags) {
    io.flag_wait = (const unsigned long long *)((const char *)p->recv_fwd + p->c2_off);
    io.flag_expect = p->n_update * (unsigned long long)ctx->nxl;  // one arrival per (source, x of the source) and column block
  }
  const void *tw;
  int rc = ctx->twiddles(io.n, &tw);
  if (rc) return rc;
  ctx->launches++;
  CK(launch_slab_xinv<T>(lc, (const cx<T> *)p->ret_stage, io, (const cx<T> *)tw, io.n));
  return MRL_OK;
}

static int slab_aux_init(mrl_slab_plan *p, int nchunks) {
  if (!p->s_aux) {
    CK(cudaStreamCreateWithFlags(&p->s_aux, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&p->ev_begin, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
  }
  while ((int)p->ev_chunk.size() < nchunks) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->ev_chunk.push_back(e);
  }
  return MRL_OK;
}

template <class T> static int slab_forward_impl(mrl_slab_plan *p, const T *c) {
  mrl_context *ctx = p->ctx;
  const mrl_split_desc &d = p->desc;
  const int nl = ctx->n[2];
  cx<T> *A = (cx<T> *)p->send_fwd;
  const void *twl;
  int rc = ctx->twiddles(nl, &twl);
  if (rc) return rc;
  NonlinDesc nlz{0, {d.nonlin_params[0], d.nonlin_params[1], d.nonlin_params[2], 0}};
  const int C = p->chunks, nyl = ctx->nyl;
  if (p->copy) {
    if (p->h_peer_recv.empty()) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_forward: peers' buffers not imported (mrl_slab_ipc_import)");
    // y-chunk by y-chunk: z pass, x pass in place (both local), then the copy engines carry x-block s of the chunk to
    // rank s while the next chunk is being transformed
    const int P = ctx->nranks, me = ctx->rank, nxl = ctx->nxl;
    const int nch = (C > 1 && nyl % C == 0 && (nyl / C) % 8 == 0) ? C : 1, ych = nyl / nch;
    if ((rc = slab_aux_init(p, nch))) return rc;
    const size_t esz = sizeof(cx<T>), rowb = (size_t)nyl * p->ncp * esz;
    const void *twx;
    if ((rc = ctx->twiddles(ctx->n[0], &twx))) return rc;
    // MRL_SLAB_COPY_SMS: SMs the passes leave free (in case the driver runs the strided peer copies as kernels)
    static const int spare = getenv("MRL_SLAB_COPY_SMS") ? atoi(getenv("MRL_SLAB_COPY_SMS")) : 0;
    const LaunchCtx lcp{ctx->stream, spare > 0 && spare < ctx->sm_count ? ctx->sm_count - spare : ctx->sm_count};
    for (int i = 0; i < nch; ++i) {
      if (d.nonlin_kind == MRL_NONLIN_EXPR) {
        const int rowmap[3] = {nch > 1 ? ych : 0, nyl, i * ych};
        if ((rc = mrl_expr_launch_zfwd_rows(ctx, d.nonlin_expr, d.nonlin_var, d.nonlin_inputs_dev, 0.0, c, d.g_out_real_dev, A, A + p->field,
                                            (long long)ctx->n[0] * ych, nl, p->ncp, rowmap, ctx->stream, lcp.sm_count)))
          return rc;
      } else {
        ctx->launches++;
        CK(launch_zfwd_nonlin_tma<T>(lcp, c, (T *)d.g_out_real_dev, A, A + p->field, (long long)ctx->n[0] * ych, nl, p->ncp, nlz,
                                     (const cx<T> *)twl, nch > 1 ? RowMap{ych, nyl, i * ych} : RowMap{0, 0, 0}));
      }
      StridedIO<T> sio;
      memset(&sio, 0, sizeof sio);
      for (int f = 0; f < 2; ++f) sio.in[f] = sio.out[f] = A + f * p->field + (long long)i * ych * p->ncp;
      sio.nfields = 2;
      sio.n = ctx->n[0];
      sio.ncols = ych * p->ncp;
      sio.nouter = 1;
      sio.pitch = (long long)nyl * p->ncp;
      sio.outer_stride = (long long)sio.n * sio.pitch;
      sio.scale = T(1);
      ctx->launches++;
      CK(launch_strided_tma<T>(lcp, sio, (const cx<T> *)twx, sio.n));
      CK(cudaEventRecord(p->ev_chunk[i], ctx->stream));
      for (int k = 0; k < P; ++k) {
        const int s = (me + 1 + k) % P;  // the own block last
        CK(cudaStreamWaitEvent(p->s_copy[s], p->ev_chunk[i], 0));
        for (int f = 0; f < 2; ++f) {
          const char *src = (const char *)(A + f * p->field + ((long long)s * nxl * nyl + (long long)i * ych) * p->ncp);
          char *dst = p->h_peer_recv[s] + ((size_t)f * p->field + ((size_t)me * nxl * nyl + (size_t)i * ych) * p->ncp) * esz;
          CK(cudaMemcpy2DAsync(dst, rowb, src, rowb, (size_t)ych * p->ncp * esz, (size_t)nxl, cudaMemcpyDefault, p->s_copy[s]));
        }
      }
    }
    for (int s = 0; s < P; ++s) {
      CK(cudaEventRecord(p->ev_copy[s], p->s_copy[s]));
      CK(cudaStreamWaitEvent(ctx->stream, p->ev_copy[s], 0));
    }
    return MRL_OK;
  }
  if (p->peer) p->n_forward++;
  if (C > 1 && p->peer && nyl % C == 0 && (nyl / C) % 8 == 0) {
    // z pass of chunk i+1 on the main stream while the x pass of chunk i pushes its rows over NVLink
    // from a few SMs on the aux stream (that pass is bound by the link, not by the SMs)
    const int ych = nyl / C;
    if ((rc = slab_aux_init(p, C))) return rc;
    const int xc = p->x_ctas < ctx->sm_count ? p->x_ctas : ctx->sm_count / 2;
    CK(cudaEventRecord(p->ev_begin, ctx->stream));
    CK(cudaStreamWaitEvent(p->s_aux, p->ev_begin, 0));
    for (int i = 0; i < C; ++i) {
      const LaunchCtx lz{ctx->stream, i == 0 ? ctx->sm_count : ctx->sm_count - xc};
      const LaunchCtx lx{p->s_aux, i == C - 1 ? ctx->sm_count : xc};
      ctx->launches++;
      if (d.nonlin_kind == MRL_NONLIN_EXPR) {
        const int rowmap[3] = {ych, nyl, i * ych};
        ctx->launches--;  // counted by the launcher
        if ((rc = mrl_expr_launch_zfwd_rows(ctx, d.nonlin_expr, d.nonlin_var, d.nonlin_inputs_dev, 0.0, c, d.g_out_real_dev, A, A + p->field,
                                            (long long)ctx->n[0] * ych, nl, p->ncp, rowmap, lz.stream, lz.sm_count)))
          return rc;
      } else {
        CK(launch_zfwd_nonlin_tma<T>(lz, c, (T *)d.g_out_real_dev, A, A + p->field, (long long)ctx->n[0] * ych, nl, p->ncp, nlz,
                                     (const cx<T> *)twl, RowMap{ych, nyl, i * ych}));
      }
      CK(cudaEventRecord(p->ev_chunk[i], ctx->stream));
      CK(cudaStreamWaitEvent(p->s_aux, p->ev_chunk[i], 0));
      if ((rc = slab_xfwd_peer<T>(p, i * ych, ych, lx))) return rc;
    }
    CK(cudaEventRecord(p->ev_done, p->s_aux));
    CK(cudaStreamWaitEvent(ctx->stream, p->ev_done, 0));
    return MRL_OK;
  }
  if (d.nonlin_kind == MRL_NONLIN_EXPR) {
    if ((rc = mrl_expr_launch_zfwd(ctx, d.nonlin_expr, d.nonlin_var, d.nonlin_inputs_dev, 0.0, c, d.g_out_real_dev, A, A + p->field,
                                   (long long)ctx->n[0] * ctx->nyl, nl, p->ncp)))
      return rc;
  } else {
    ctx->launches++;
    CK(launch_zfwd_nonlin_tma<T>(ctx->lc(), c, (T *)d.g_out_real_dev, A, A + p->field, (long long)ctx->n[0] * ctx->nyl, nl, p->ncp, nlz,
                                 (const cx<T> *)twl));
  }
  if (p->peer) return slab_xfwd_peer<T>(p, 0, nyl, ctx->lc());
  return slab_xpass<T>(p, 2, 0);
}

template <class T> static int slab_update_impl(mrl_slab_plan *p, double dt, const double *beta, int nold) {
  mrl_context *ctx = p->ctx;
  const mrl_split_desc &d = p->desc;
  cx<T> *S = (cx<T> *)p->recv_fwd;
  FusedIO<T> io;
  memset(&io, 0, sizeof io);
  io.inC = S;
  io.inG = S + p->field;
  io.outU = (cx<T> *)p->send_bwd;
  io.n = ctx->gn[1];
  io.ncols = p->ncp;
  io.nouter = ctx->nxl;
  io.pitch = p->ncp;
  io.scale = T(1);
  io.slab = 1;
  io.nyl = ctx->nyl;
  io.nranks = ctx->nranks;
  if (p->peer) {  // rows of the updated field go straight back into the owners' return staging
    p->n_update++;
    io.slab = 2;
    io.peer_tab = (const unsigned long long *)p->peer_send_tab;
    io.peer_x0 = ctx->x0;
    io.kzb_major = p->kzb_major;
    io.nx = ctx->n[0];
    io.rank = ctx->rank;
    io.ring_old = p->ring.empty() ? p->recv_fwd : p->ring[0];
    if (p->sync_flags) {
      io.flag_wait = (const unsigned long long *)((const char *)p->recv_fwd + p->c1_off);
      io.flag_expect = p->n_forward * 2ull * (unsigned long long)ctx->nyl;  // two spectra x nyl tiles per source and column block
      io.flag_tab = (const unsigned long long *)p->flag2_tab;
    }
  }
  SpectralUpdate<T> up;
  memset(&up, 0, sizeof up);
  up.kx = (const T *)ctx->kaxis_dev[0];
  up.ky = (const T *)ctx->kaxis_dev[1];
  up.kz = (const T *)ctx->kaxis_dev[2];
  up.kmode = MRL_KMODE_3D_SLAB;
  up.nzc = p->ncp;
  up.nzv = ctx->nr[2];
  up.x0 = ctx->dist ? 0 : ctx->x0;  // a decomposed context holds the local slice of the kx axis
  up.closed_M = 1;
  up.Mfac = (T)d.M_factor;
  up.has_L = d.has_L;
  up.closed_L = 1;
  up.Lfac = (T)d.L_factor;
  up.dt = (T)dt;
  up.b0 = (T)(dt * beta[0]);
  up.nold = nold;
  const int H = d.history;
  for (int i = 0; i < nold; ++i) {
    up.bold[i] = (T)(dt * beta[i + 1]);
    up.Nold[i] = (const cx<T> *)p->ring[((p->cur - 1 - i) % (H + 1) + (H + 1)) % (H + 1)];
  }
  up.Nout = H > 0 ? (cx<T> *)p->ring[p->cur] : nullptr;
  const void *tw;
  int rc = ctx->twiddles(io.n, &tw);
  if (rc) return rc;
  p->inverse_issued = false;
  if (p->peer && p->sync_flags && p->inv_ctas > 0 && p->inv_ctas < ctx->sm_count) {
    // the x inverse pass trails the fused y pass column block by column block on its own SMs: both kernels are resident
    // together (one CTA per SM each), the inverse one waiting on the arrival counters of the return exchange
    if ((rc = slab_aux_init(p, 0))) return rc;
    CK(cudaEventRecord(p->ev_begin, ctx->stream));
    ctx->launches++;
    CK(launch_fused_tma<T>(LaunchCtx{ctx->stream, ctx->sm_count - p->inv_ctas}, io, up, (const cx<T> *)tw, io.n));
    CK(cudaStreamWaitEvent(p->s_aux, p->ev_begin, 0));
    if ((rc = slab_xinv_peer<T>(p, LaunchCtx{p->s_aux, p->inv_ctas}))) return rc;
    CK(cudaEventRecord(p->ev_done, p->s_aux));
    p->inverse_issued = true;
    return MRL_OK;
  }
  if (p->copy) {
    // the fused y pass in x-chunks; behind every chunk the copy engines carry the y-block of rank s back to rank s (it
    // lands at x = x0 .. of the array its inverse x pass transforms), while the next chunk is being computed
    if (p->h_peer_ret.empty()) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_update: peers' buffers not imported (mrl_slab_ipc_import)");
    const int P = ctx->nranks, me = ctx->rank, nxl = ctx->nxl, nyl = ctx->nyl;
    const int nch = (p->y_chunks > 1 && nxl % p->y_chunks == 0) ? p->y_chunks : 1, xch = nxl / nch;
    if ((rc = slab_aux_init(p, nch))) return rc;
    const size_t esz = sizeof(cx<T>);
    const long long plane = (long long)nyl * p->ncp;  // elements of one x layer of one rank block
    const FusedIO<T> io_all = io;
    const SpectralUpdate<T> up_all = up;
    for (int i = 0; i < nch; ++i) {
      const long long sh = (long long)i * xch * plane;
      FusedIO<T> ioc = io_all;
      SpectralUpdate<T> upc = up_all;
      ioc.inC = io_all.inC + sh;
      ioc.inG = io_all.inG + sh;
      ioc.outU = io_all.outU + sh;
      ioc.nouter = xch;
      ioc.nouter_full = nxl;
      upc.x0 = up_all.x0 + i * xch;
      if (upc.Nout) upc.Nout = up_all.Nout + sh;
      for (int k = 0; k < 4; ++k)
        if (upc.Nold[k]) upc.Nold[k] = up_all.Nold[k] + sh;
      ctx->launches++;
      CK(launch_fused_tma<T>(ctx->lc(), ioc, upc, (const cx<T> *)tw, ioc.n));
      CK(cudaEventRecord(p->ev_chunk[i], ctx->stream));
      for (int k = 0; k < P; ++k) {
        const int s = (me + 1 + k) % P;
        CK(cudaStreamWaitEvent(p->s_copy[s], p->ev_chunk[i], 0));
        const char *src = (const char *)((const cx<T> *)p->send_bwd + ((long long)s * nxl + (long long)i * xch) * plane);
        char *dst = p->h_peer_ret[s] + (size_t)(((long long)me * nxl + (long long)i * xch) * plane) * esz;
        CK(cudaMemcpyAsync(dst, src, (size_t)xch * plane * esz, cudaMemcpyDefault, p->s_copy[s]));
      }
    }
    for (int s = 0; s < P; ++s) {
      CK(cudaEventRecord(p->ev_copy[s], p->s_copy[s]));
      CK(cudaStreamWaitEvent(ctx->stream, p->ev_copy[s], 0));
    }
    return MRL_OK;
  }
  ctx->launches++;
  CK(launch_fused_tma<T>(ctx->lc(), io, up, (const cx<T> *)tw, io.n));
  return MRL_OK;
}

template <class T> static int slab_inverse_impl(mrl_slab_plan *p, T *c) {
  mrl_context *ctx = p->ctx;
  int rc;
  if (p->peer) {
    if (p->inverse_issued) {
      CK(cudaStreamWaitEvent(ctx->stream, p->ev_done, 0));
      p->inverse_issued = false;
    } else if ((rc = slab_xinv_peer<T>(p, ctx->lc()))) {
      return rc;
    }
  } else if ((rc = slab_xpass<T>(p, 1, 1))) {
    return rc;
  }
  const int nl = ctx->n[2];
  const void *twl;
  if ((rc = ctx->twiddles(nl, &twl))) return rc;
  const double N = (double)ctx->gn[0] * ctx->gn[1] * ctx->gn[2];
  ctx->launches++;
  CK(launch_zinv_pairs_tma<T>(ctx->lc(), (const cx<T> *)p->send_fwd, p->ncp, c, (long long)ctx->n[0] * ctx->nyl, nl, (T)(1.0 / N),
                              (const cx<T> *)twl));
  return MRL_OK;
}

extern "C" int mrl_slab_forward(mrl_slab_plan *p, const void *c) {
  if (!p || !c) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_forward: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? slab_forward_impl<double>(p, (const double *)c) : slab_forward_impl<float>(p, (const float *)c);
}
extern "C" int mrl_slab_update(mrl_slab_plan *p, double dt, const double *beta, int nold) {
  if (!p || !beta || nold < 0) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_update: bad arguments");
  if (nold > p->stored) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_update: %d old states requested, %d stored", nold, p->stored);
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? slab_update_impl<double>(p, dt, beta, nold) : slab_update_impl<float>(p, dt, beta, nold);
}
extern "C" int mrl_slab_inverse(mrl_slab_plan *p, void *c) {
  if (!p || !c) return mrl_fail(MRL_ERR_INVALID, "mrl_slab_inverse: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? slab_inverse_impl<double>(p, (double *)c) : slab_inverse_impl<float>(p, (float *)c);
}
