// marlin_b200 - C ABI of the de Geus FFT mechanics solve (host logic; kernels in k_mech.cu and
// the batched FFT passes).  Reference: FFTMechanics::computeBuffer
// (src/tensor_computes/FFTMechanics.C:96-163), conjugateGradientSolve
// (include/utils/MarlinUtils.h:57-131), HyperElasticIsotropic::computeBuffer
// (src/tensor_computes/HyperElasticIsotropic.C:42-52).
#include <cmath>
#include <cstring>
#include <vector>

#include "mrl_internal.h"

using namespace mrl;

namespace mrl {
enum { SC_RZ = 0, SC_PAP = 1, SC_ALPHA = 2, SC_RES2 = 3, SC_BETA = 4, SC_TMP = 5, SC_COUNT = 8 };
enum { VOP_DOT = 0, VOP_CG_XR = 1, VOP_XPBY = 2, VOP_AXPY = 3, VOP_SUB = 4, VOP_COPY = 5 };
enum { FIN_STORE = 0, FIN_ALPHA = 1, FIN_RES = 2 };
template <class T>
cudaError_t launch_mech_pointwise(const LaunchCtx &lc, int dim, int mode, const T *F, const T *K, const T *mu, const T *x, const double *xc,
                                  T *out, long long n, double scale, const T *r = nullptr, T *xw = nullptr, const double *scal = nullptr);
template <class T> cudaError_t launch_vec_final(const LaunchCtx &lc, int fin, int slot, const double *partials, int nblk, double *scal);
template <class T>
cudaError_t launch_mech_project(const LaunchCtx &lc, int dim, cx<T> *A, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp);
template <class T>
cudaError_t launch_vec(const LaunchCtx &lc, int op, const T *a, const T *b, T *y, T *z, double *scal, double s, long long n, int fin, int slot,
                       double *partials, int nblk);
template <class T> cudaError_t launch_add_const9(const LaunchCtx &lc, int dim, T *y, const double *s, long long n);
template <class T> cudaError_t launch_components(const LaunchCtx &lc, const T *in, T *out, long long n, int ncomp, int to_soa);
template <class T> cudaError_t launch_von_mises(const LaunchCtx &lc, int dim, const T *s, T *out, long long n);
template <class T>
cudaError_t launch_disp_contract(const LaunchCtx &lc, int dim, cx<T> *H, const T *kx, const T *ky, const T *kz, int n0, int n1, int nzc, int ncp);
template <class T>
cudaError_t launch_disp_nodal(const LaunchCtx &lc, int dim, const T *uper, T *out, const T *const *ax, const int *n, const double *A);
}  // namespace mrl

#define CK(call)                                                                                                      \
  do {                                                                                                                \
    cudaError_t e_ = (call);                                                                                          \
    if (e_ != cudaSuccess)                                                                                            \
      return mrl_fail(MRL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);      \
  } while (0)

struct mrl_mech_plan {
  mrl_context *ctx = nullptr;
  mrl_mech_desc desc;
  const void *K = nullptr, *mu = nullptr;
  long long n = 0;  // voxels
  int dim = 3, nc = 9;  // tensors are dim x dim (FFTMechanics.C:50-58), nc = dim*dim components
  int ncp = 0;
  void *spec = nullptr;                                                    // [nc][n0][n1][ncp] complex
  void *tmp = nullptr, *rhs = nullptr, *x = nullptr, *r = nullptr, *p = nullptr, *Ap = nullptr, *Fk = nullptr;  // [nc][n] real
  double *scal = nullptr, *partials = nullptr, *host = nullptr;
  int nblk = 0;
  bool fused_x = true;  // x pass fused with the Green projection (sizes with a TMA configuration)
  // decomposed domain (mrl_domain_set_dist / _pencil): the transforms go through `dist`, the inner products of the CG and
  // Newton recurrences are summed over the ranks by the caller's `allreduce` (host side, like the reference's MPI reductions)
  mrl_dist *dist = nullptr;
  mrl_allreduce_fn allreduce = nullptr;
  void *allreduce_user = nullptr;
  bool fused_tangent = true;  // tangent fused into the first FFT pass (3-D, last axis 256 or 512, x pass fused)
  int dot_count = 0;    // partial sums the last inverse pass left in `partials` (0: the inner product was not fused)
};

template <class T> static int project_G_passes(mrl_mech_plan *p, const T *A, T *out, double sign, const T *dot_with);

extern "C" int mrl_mech_plan_destroy(mrl_mech_plan *p) {
  if (!p) return MRL_OK;
  mrl_quiesce(p->ctx);
  for (void *q : {p->spec, p->tmp, p->rhs, p->x, p->r, p->p, p->Ap, p->Fk, (void *)p->scal, (void *)p->partials}) cudaFree(q);
  if (p->host) cudaFreeHost(p->host);
  delete p;
  return MRL_OK;
}

extern "C" int mrl_mech_plan_create(mrl_context *ctx, const mrl_mech_desc *d, const void *K, const void *mu, mrl_mech_plan **out) {
  if (!ctx || !d || !K || !mu || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_plan_create: bad arguments");
  if (ctx->dim != 3 && ctx->dim != 2)
    return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_mech_plan_create: the CUDA mechanics path is 2-D or 3-D (dim = %d)", ctx->dim);
  CK(cudaSetDevice(ctx->device));
  mrl_mech_plan *p = new mrl_mech_plan();
  p->ctx = ctx;
  p->desc = *d;
  p->K = K;
  p->mu = mu;
  p->n = ctx->total();
  p->dim = ctx->dim;
  p->nc = ctx->dim * ctx->dim;
  p->fused_x = ctx->dim == 3 && !ctx->dist;
  p->fused_tangent = p->fused_x;
  if (p->desc.l_max_its <= 0) p->desc.l_max_its = (long long)ctx->gn[0] * ctx->gn[1] * ctx->gn[2];  // FFTMechanics.C:63-64: default = number of cells
  p->ncp = ctx->dist ? ctx->nr[ctx->dim - 1] : mrl_fftb_pitch(ctx);
  const size_t esz = ctx->precision == MRL_F64 ? 8 : 4;
  const size_t vbytes = p->nc * (size_t)p->n * esz;
  const size_t sbytes = ctx->dist ? p->nc * (size_t)ctx->nr[0] * ctx->nr[1] * ctx->nr[2] * 2 * esz
                                  : p->nc * (size_t)ctx->n[0] * (ctx->dim == 3 ? ctx->n[1] : 1) * p->ncp * 2 * esz;
  p->nblk = ctx->sm_count * 4;
  cudaError_t e = cudaMalloc(&p->spec, sbytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(p->spec, 0, sbytes, ctx->stream);
  for (void **q : {&p->tmp, &p->rhs, &p->x, &p->r, &p->p, &p->Ap, &p->Fk})
    if (e == cudaSuccess) e = cudaMalloc(q, vbytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&p->scal, SC_COUNT * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&p->partials, p->nblk * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost((void **)&p->host, SC_COUNT * sizeof(double));
  if (e != cudaSuccess) {
    mrl_mech_plan_destroy(p);
    return mrl_fail(MRL_ERR_CUDA, "mechanics plan allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return MRL_OK;
}

extern "C" int mrl_mech_plan_set_dist(mrl_mech_plan *p, mrl_dist *dist, mrl_allreduce_fn allreduce_sum, void *user) {
  if (!p || !dist || !allreduce_sum) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_plan_set_dist: bad arguments");
  if (!p->ctx->dist) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_plan_set_dist: the plan's domain is not decomposed");
  p->dist = dist;
  p->allreduce = allreduce_sum;
  p->allreduce_user = user;
  return MRL_OK;
}

// ---------------------------------------------------------------------------- operators
template <class T> static int vec(mrl_mech_plan *p, int op, const T *a, const T *b, T *y, T *z, double s, int fin = FIN_STORE, int slot = SC_TMP);

// dot_with != nullptr: alpha = rz / (dot_with . out) follows (FIN_ALPHA), with the inner product riding in the store of
// the last inverse pass where that pass can carry it
template <class T> static int finish_dot(mrl_mech_plan *p, const T *dot_with, const T *out);
template <class T> static int project_G(mrl_mech_plan *p, const T *A, T *out, double sign, const T *dot_with = nullptr) {
  int rc = project_G_passes<T>(p, A, out, sign, dot_with);
  return rc ? rc : finish_dot<T>(p, dot_with, out);
}

// alpha = rz / (dot_with . out) after an inverse transform that may have carried the inner product in its store
template <class T> static int finish_dot(mrl_mech_plan *p, const T *dot_with, const T *out) {
  if (!dot_with) return MRL_OK;
  if (p->dot_count > 0) {
    p->ctx->launches++;
    CK(launch_vec_final<T>(p->ctx->lc(), FIN_ALPHA, SC_TMP, p->partials, p->dot_count, p->scal));
    return MRL_OK;
  }
  return vec<T>(p, VOP_DOT, dot_with, out, nullptr, nullptr, 0, FIN_ALPHA, SC_TMP);
}

// the part of project_G after the z and y forward passes (3-D): p->spec holds the partial spectra
template <class T> static int project_G_from_spec(mrl_mech_plan *p, T *out, double sign, const T *dot_with) {
  mrl_context *ctx = p->ctx;
  p->dot_count = 0;
  const T *kx = (const T *)ctx->kaxis_dev[0], *ky = (const T *)ctx->kaxis_dev[1], *kz = (const T *)ctx->kaxis_dev[2];
  int rc;
  if (p->fused_x) {
    const void *tw;
    if ((rc = ctx->twiddles(ctx->n[0], &tw))) return rc;
    cudaError_t e = launch_mech_fused_tma<T>(ctx->lc(), (cx<T> *)p->spec, kx, ky, kz, ctx->n[0], ctx->n[1], ctx->nr[2], p->ncp, (const cx<T> *)tw);
    if (e == cudaSuccess) {
      ctx->launches++;
      if ((rc = mrl_fftb_inverse(ctx, p->spec, out, p->nc, p->ncp, sign / (double)p->n, 1, dot_with, p->partials, p->nblk, &p->dot_count))) return rc;
      return finish_dot<T>(p, dot_with, out);
    }
    if (e != cudaErrorNotSupported) CK(e);
    p->fused_x = false;  // no pipelined configuration for the x axis of this grid: separate passes from here on
  }
  if ((rc = mrl_fftb_strided(ctx, p->spec, p->nc, p->ncp, 0, 0))) return rc;
  ctx->launches++;
  CK(launch_mech_project<T>(ctx->lc(), p->dim, (cx<T> *)p->spec, kx, ky, kz, ctx->nr[0], ctx->nr[1], ctx->nr[2], p->ncp));
  if ((rc = mrl_fftb_inverse(ctx, p->spec, out, p->nc, p->ncp, sign / (double)p->n, 0, dot_with, p->partials, p->nblk, &p->dot_count))) return rc;
  return finish_dot<T>(p, dot_with, out);
}

template <class T> static int project_G_passes(mrl_mech_plan *p, const T *A, T *out, double sign, const T *dot_with) {
  // out = sign * irfftn( Ghat4 : rfftn(A) ), FFTMechanics.C:104-105
  mrl_context *ctx = p->ctx;
  p->dot_count = 0;
  const int nc = p->nc;
  const T *kx = (const T *)ctx->kaxis_dev[0], *ky = (const T *)ctx->kaxis_dev[1], *kz = (const T *)ctx->kaxis_dev[2];
  // 2-D: the half-spectrum axis is y; the projection kernel sees [nc][n0][1][ncp] with q = (kx, ky)
  const T *klast = p->dim == 3 ? kz : ky;
  const int n1 = p->dim == 3 ? ctx->nr[1] : 1, nlast = ctx->nr[p->dim - 1];
  int rc;
  if (ctx->dist) {
    // decomposed domain: transforms with their exchanges, the projection on this rank's wavevectors (local k-axes)
    if (!p->dist) return mrl_fail(MRL_ERR_INVALID, "mechanics on a decomposed domain needs mrl_mech_plan_set_dist");
    if ((rc = mrl_dist_rfftn(p->dist, A, p->spec, nc))) return rc;
    ctx->launches++;
    CK(launch_mech_project<T>(ctx->lc(), p->dim, (cx<T> *)p->spec, kx, ky, klast, ctx->nr[0], n1, nlast, p->ncp));
    return mrl_dist_irfftn_scaled(p->dist, p->spec, out, nc, sign);
  }
  if (p->fused_x) {
    // z, y forward; x forward + projection + x inverse in ONE pass over the spectra; y, z inverse
    if ((rc = mrl_fftb_forward(ctx, A, p->spec, nc, p->ncp, 1))) return rc;
    const void *tw;
    if ((rc = ctx->twiddles(ctx->n[0], &tw))) return rc;
    cudaError_t e = launch_mech_fused_tma<T>(ctx->lc(), (cx<T> *)p->spec, kx, ky, kz, ctx->n[0], ctx->n[1], ctx->nr[2], p->ncp,
                                             (const cx<T> *)tw);
    if (e == cudaSuccess) {
      ctx->launches++;
      return mrl_fftb_inverse(ctx, p->spec, out, nc, p->ncp, sign / (double)p->n, 1, dot_with, p->partials, p->nblk, &p->dot_count);
    }
    if (e != cudaErrorNotSupported) CK(e);
    p->fused_x = false;  // no pipelined configuration for this size: finish with the separate passes
    if ((rc = mrl_fftb_strided(ctx, p->spec, nc, p->ncp, 0, 0))) return rc;
  } else if ((rc = mrl_fftb_forward(ctx, A, p->spec, nc, p->ncp))) {
    return rc;
  }
  ctx->launches++;
  CK(launch_mech_project<T>(ctx->lc(), p->dim, (cx<T> *)p->spec, kx, ky, klast, ctx->nr[0], n1, nlast, p->ncp));
  return mrl_fftb_inverse(ctx, p->spec, out, nc, p->ncp, sign / (double)p->n, 0, dot_with, p->partials, p->nblk, &p->dot_count);
}

// r_update != nullptr: x is the CG direction, first replaced by r_update + beta x (beta on the device).
// cg_dot: alpha = rz / (x . out) follows.
template <class T>
static int apply_GK(mrl_mech_plan *p, const T *F, const T *x, const double *xconst, T *out, double sign, const T *r_update = nullptr,
                    bool cg_dot = false) {
  // out = sign * G( K4(F) : x ), FFTMechanics.C:107-112
  mrl_context *ctx = p->ctx;
  if (p->fused_tangent && !xconst && x) {
    // tangent (and direction update) in the load of the first FFT pass: the product never goes through HBM
    MechTangentIO<T> io;
    io.F = F;
    io.K = (const T *)p->K;
    io.mu = (const T *)p->mu;
    io.p = const_cast<T *>(x);
    io.r = r_update;
    io.scal = p->scal;
    io.n = p->n;
    io.nrows = (long long)ctx->n[0] * ctx->n[1];
    io.out = (cx<T> *)p->spec;
    io.ncp = p->ncp;
    const void *twz;
    int rc = ctx->twiddles(ctx->n[2], &twz);
    if (rc) return rc;
    cudaError_t e = launch_mech_tangent_zfwd<T>(ctx->lc(), io, (const cx<T> *)twz, ctx->n[2]);
    if (e == cudaSuccess) {
      ctx->launches++;
      if ((rc = mrl_fftb_strided(ctx, p->spec, p->nc, p->ncp, 1, 0))) return rc;  // y forward
      return project_G_from_spec<T>(p, out, sign, cg_dot ? x : nullptr);
    }
    if (e != cudaErrorNotSupported) CK(e);
    p->fused_tangent = false;
  }
  ctx->launches++;
  CK(launch_mech_pointwise<T>(ctx->lc(), p->dim, r_update ? 3 : xconst ? 2 : 1, F, (const T *)p->K, (const T *)p->mu, x, xconst, (T *)p->tmp, p->n,
                              1.0, r_update, const_cast<T *>(x), p->scal));
  return project_G<T>(p, (const T *)p->tmp, out, sign, cg_dot ? x : nullptr);
}

template <class T> static int vec(mrl_mech_plan *p, int op, const T *a, const T *b, T *y, T *z, double s, int fin, int slot) {
  p->ctx->launches++;
  const bool reduces = op == VOP_DOT || op == VOP_CG_XR;
  if (!(p->dist && reduces)) {
    CK(launch_vec<T>(p->ctx->lc(), op, a, b, y, z, p->scal, s, p->nc * p->n, fin, slot, p->partials, p->nblk));
    return MRL_OK;
  }
  // decomposed domain: this rank's sum -> host -> sum over the ranks -> device, then the dependent scalars
  CK(launch_vec<T>(p->ctx->lc(), op, a, b, y, z, p->scal, s, p->nc * p->n, FIN_STORE, SC_TMP, p->partials, p->nblk));
  CK(cudaMemcpyAsync(p->host, p->scal + SC_TMP, sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream));
  CK(cudaStreamSynchronize(p->ctx->stream));
  if (p->allreduce(p->allreduce_user, p->host, 1)) return mrl_fail(MRL_ERR_INVALID, "mechanics: the caller's allreduce failed");
  CK(cudaMemcpyAsync(p->partials, p->host, sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
  CK(launch_vec_final<T>(p->ctx->lc(), fin, slot, p->partials, 1, p->scal));
  CK(cudaStreamSynchronize(p->ctx->stream));  // p->host is reused by the next reduction
  return MRL_OK;
}
static int read_scalar(mrl_mech_plan *p, int slot, double *v) {
  CK(cudaMemcpyAsync(p->host, p->scal, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream));
  CK(cudaStreamSynchronize(p->ctx->stream));
  *v = p->host[slot];
  return MRL_OK;
}

// conjugateGradientSolve, include/utils/MarlinUtils.h:57-131 (identity preconditioner): solves
// G(K4(Fk) : x) = rhs, x warm-started from p->x.  x_zero: p->x is known to be all zeros.
template <class T> static int cg_solve(mrl_mech_plan *p, bool x_zero, int *iterations) {
  const T *Fk = (const T *)p->Fk;
  T *x = (T *)p->x, *r = (T *)p->r, *pp = (T *)p->p, *Ap = (T *)p->Ap;
  const T *b = (const T *)p->rhs;
  int rc;
  double bb;
  if ((rc = vec<T>(p, VOP_DOT, b, b, nullptr, nullptr, 0, FIN_STORE, SC_TMP))) return rc;
  if ((rc = read_scalar(p, SC_TMP, &bb))) return rc;
  const double b_norm = std::sqrt(bb);
  *iterations = 0;
  if (b_norm == 0.0) return MRL_OK;  // :63-66
  if (x_zero) {
    if ((rc = vec<T>(p, VOP_COPY, b, nullptr, r, nullptr, 0))) return rc;  // r = b - A(0)
  } else {
    if ((rc = apply_GK<T>(p, Fk, x, nullptr, Ap, 1.0))) return rc;
    if ((rc = vec<T>(p, VOP_SUB, b, Ap, nullptr, r, 0))) return rc;
  }
  if ((rc = vec<T>(p, VOP_COPY, r, nullptr, pp, nullptr, 0))) return rc;
  if ((rc = vec<T>(p, VOP_DOT, r, r, nullptr, nullptr, 0, FIN_STORE, SC_RZ))) return rc;
  const long long maxit = p->desc.l_max_its;
  for (long long k = 0; k < maxit; ++k) {
    // p = r + beta p of the previous iteration rides in the load of the tangent kernel; p.Ap -> alpha rides in the store of
    // the last inverse pass (no separate pass over p and Ap)
    if ((rc = apply_GK<T>(p, Fk, pp, nullptr, Ap, 1.0, k > 0 ? r : nullptr, true))) return rc;
    if ((rc = vec<T>(p, VOP_CG_XR, pp, Ap, x, r, 0, FIN_RES))) return rc;  // x, r, |r|^2, beta
    double res2;
    if ((rc = read_scalar(p, SC_RES2, &res2))) return rc;
    *iterations = (int)(k + 1);
    if (std::sqrt(res2) <= p->desc.l_tol * b_norm) return MRL_OK;
  }
  return MRL_OK;
}

template <class T> static int solve_impl(mrl_mech_plan *p, T *F, const double *applied, T *P, mrl_mech_stats *st) {
  mrl_context *ctx = p->ctx;
  const mrl_mech_desc &d = p->desc;
  const size_t vbytes = p->nc * (size_t)p->n * sizeof(T);
  int rc;
  memset(st, 0, sizeof *st);
  // _u = F; constitutive model evaluated at F (:114-116) - the tangent stays at this state until
  // the first Newton update even though the applied strain is added to _u right away (:118-123)
  CK(cudaMemcpyAsync(p->Fk, F, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
  const double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if ((rc = apply_GK<T>(p, (const T *)p->Fk, nullptr, applied ? applied : zero9, (T *)p->rhs, -1.0))) return rc;
  if (applied) {
    ctx->launches++;
    CK(launch_add_const9<T>(ctx->lc(), p->dim, F, applied, p->n));
  }
  double fn2;
  if ((rc = vec<T>(p, VOP_DOT, F, F, nullptr, nullptr, 0, FIN_STORE, SC_TMP))) return rc;
  if ((rc = read_scalar(p, SC_TMP, &fn2))) return rc;
  const double Fn = std::sqrt(fn2);
  CK(cudaMemsetAsync(p->x, 0, vbytes, ctx->stream));
  bool x_zero = true;
  int iiter = 0;
  while (true) {
    int its = 0;
    if ((rc = cg_solve<T>(p, x_zero, &its))) return rc;
    if (st->cg_solves < 64) st->cg_iterations[st->cg_solves] = its;
    st->cg_solves++;
    st->cg_iterations_total += its;
    if (its > 0) x_zero = false;
    // _u += dFm; constitutive; b = -G(P)   (:137-143)
    if ((rc = vec<T>(p, VOP_AXPY, (const T *)p->x, nullptr, F, nullptr, 1.0))) return rc;
    CK(cudaMemcpyAsync(p->Fk, F, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches++;
    CK(launch_mech_pointwise<T>(ctx->lc(), p->dim, 0, (const T *)p->Fk, (const T *)p->K, (const T *)p->mu, nullptr, nullptr, P, p->n, 1.0));
    if ((rc = project_G<T>(p, P, (T *)p->rhs, -1.0))) return rc;
    double x2;
    if ((rc = vec<T>(p, VOP_DOT, (const T *)p->x, (const T *)p->x, nullptr, nullptr, 0, FIN_STORE, SC_TMP))) return rc;
    if ((rc = read_scalar(p, SC_TMP, &x2))) return rc;
    const double anorm = std::sqrt(x2), rnorm = anorm / Fn;
    st->final_anorm = anorm;
    st->final_rnorm = rnorm;
    if ((rnorm < d.nl_rel_tol || anorm < d.nl_abs_tol) && iiter > 0) break;  // :145-155
    iiter++;
    if (iiter > d.nl_max_its)
      return mrl_fail(MRL_ERR_INVALID, "Exceeded the maximum number of nonlinear iterations without converging.");
  }
  st->newton_iterations = iiter + 1;
  return MRL_OK;
}

#define DISPATCH(ctx, fn, ...) ((ctx)->precision == MRL_F64 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

extern "C" int mrl_mech_constitutive(mrl_mech_plan *p, const void *F, void *P) {
  if (!p || !F || !P) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_constitutive: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  p->ctx->launches++;
  if (p->ctx->precision == MRL_F64)
    CK(launch_mech_pointwise<double>(p->ctx->lc(), p->dim, 0, (const double *)F, (const double *)p->K, (const double *)p->mu, nullptr, nullptr,
                                     (double *)P, p->n, 1.0));
  else
    CK(launch_mech_pointwise<float>(p->ctx->lc(), p->dim, 0, (const float *)F, (const float *)p->K, (const float *)p->mu, nullptr, nullptr,
                                    (float *)P, p->n, 1.0));
  return MRL_OK;
}
extern "C" int mrl_mech_apply_G(mrl_mech_plan *p, const void *A, void *out) {
  if (!p || !A || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_apply_G: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? project_G<double>(p, (const double *)A, (double *)out, 1.0)
                                      : project_G<float>(p, (const float *)A, (float *)out, 1.0);
}
extern "C" int mrl_mech_apply_GK(mrl_mech_plan *p, const void *F, const void *x, void *out) {
  if (!p || !F || !x || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_apply_GK: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? apply_GK<double>(p, (const double *)F, (const double *)x, nullptr, (double *)out, 1.0)
                                      : apply_GK<float>(p, (const float *)F, (const float *)x, nullptr, (float *)out, 1.0);
}
extern "C" int mrl_mech_solve(mrl_mech_plan *p, void *F, const double *applied, void *P, mrl_mech_stats *stats) {
  if (!p || !F || !P || !stats) return mrl_fail(MRL_ERR_INVALID, "mrl_mech_solve: bad arguments");
  CK(cudaSetDevice(p->ctx->device));
  return p->ctx->precision == MRL_F64 ? solve_impl<double>(p, (double *)F, applied, (double *)P, stats)
                                      : solve_impl<float>(p, (float *)F, applied, (float *)P, stats);
}
extern "C" int mrl_components(mrl_context *ctx, const void *in, void *out, int64_t n, int ncomp, int to_soa) {
  if (!ctx || !in || !out || n < 1 || ncomp < 1 || in == out) return mrl_fail(MRL_ERR_INVALID, "mrl_components: bad arguments");
  CK(cudaSetDevice(ctx->device));
  ctx->launches++;
  if (ctx->precision == MRL_F64) CK(launch_components<double>(ctx->lc(), (const double *)in, (double *)out, n, ncomp, to_soa));
  else CK(launch_components<float>(ctx->lc(), (const float *)in, (float *)out, n, ncomp, to_soa));
  return MRL_OK;
}

extern "C" int mrl_von_mises(mrl_context *ctx, const void *stress, void *out) {
  if (!ctx || !stress || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_von_mises: bad arguments");
  if (ctx->dim != 2 && ctx->dim != 3) return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_von_mises: Unsupported problem dimension %d", ctx->dim);
  CK(cudaSetDevice(ctx->device));
  ctx->launches++;
  if (ctx->precision == MRL_F64) CK(launch_von_mises<double>(ctx->lc(), ctx->dim, (const double *)stress, (double *)out, ctx->total()));
  else CK(launch_von_mises<float>(ctx->lc(), ctx->dim, (const float *)stress, (float *)out, ctx->total()));
  return MRL_OK;
}

// ComputeDisplacements::computeBuffer (src/tensor_computes/ComputeDisplacements.C:53-107)
template <class T> static int displacements_impl(mrl_context *ctx, const T *F, T *out) {
  const int D = ctx->dim, nc = D * D;
  const long long n = ctx->total();
  const int ncp = mrl_fftb_pitch(ctx);
  const size_t vbytes = (size_t)nc * n * sizeof(T);
  const size_t sbytes = (size_t)nc * ctx->n[0] * (D == 3 ? ctx->n[1] : 1) * ncp * 2 * sizeof(T);
  void *tmp = nullptr, *spec = nullptr;
  CK(cudaMalloc(&tmp, vbytes));
  cudaError_t e = cudaMalloc(&spec, sbytes);
  if (e != cudaSuccess) {
    cudaFree(tmp);
    return mrl_fail(MRL_ERR_CUDA, "mrl_displacements: allocation failed: %s", cudaGetErrorString(e));
  }
  auto done = [&](int rc) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    cudaFree(spec);
    return rc;
  };
  // Fbox = <F> (DomainAction::average), H = F - Fbox
  double Fbox[9], A[9], neg[9];
  for (int c = 0; c < nc; ++c) {
    int rc = mrl_reduce(ctx, MRL_SUM, F + (long long)c * n, n, &Fbox[c]);
    if (rc) return done(rc);
    Fbox[c] /= (double)n;
    neg[c] = -Fbox[c];
    A[c] = Fbox[c] - (c / D == c % D ? 1.0 : 0.0);
  }
  if (cudaMemcpyAsync(tmp, F, vbytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) return done(mrl_fail(MRL_ERR_CUDA, "copy failed"));
  ctx->launches++;
  if (launch_add_const9<T>(ctx->lc(), D, (T *)tmp, neg, n) != cudaSuccess) return done(mrl_fail(MRL_ERR_CUDA, "launch failed"));
  if (cudaMemsetAsync(spec, 0, sbytes, ctx->stream) != cudaSuccess) return done(mrl_fail(MRL_ERR_CUDA, "memset failed"));
  int rc = mrl_fftb_forward(ctx, tmp, spec, nc, ncp);
  if (rc) return done(rc);
  const T *kx = (const T *)ctx->kaxis_dev[0], *ky = (const T *)ctx->kaxis_dev[1], *kz = (const T *)ctx->kaxis_dev[2];
  ctx->launches++;
  if (launch_disp_contract<T>(ctx->lc(), D, (cx<T> *)spec, kx, ky, D == 3 ? kz : ky, ctx->n[0], D == 3 ? ctx->n[1] : 1, ctx->nr[D - 1], ncp) != cudaSuccess)
    return done(mrl_fail(MRL_ERR_CUDA, "launch failed"));
  // u_periodic = irfftn of the first D spectra (into tmp)
  if ((rc = mrl_fftb_inverse(ctx, spec, tmp, D, ncp, 1.0 / (double)n))) return done(rc);
  const T *ax[3] = {(const T *)ctx->axis_dev[0], (const T *)ctx->axis_dev[1], (const T *)ctx->axis_dev[2]};
  ctx->launches++;
  if (launch_disp_nodal<T>(ctx->lc(), D, (const T *)tmp, out, ax, ctx->n, A) != cudaSuccess) return done(mrl_fail(MRL_ERR_CUDA, "launch failed"));
  return done(MRL_OK);
}

extern "C" int mrl_displacements(mrl_context *ctx, const void *F, void *out) {
  if (!ctx || !F || !out) return mrl_fail(MRL_ERR_INVALID, "mrl_displacements: bad arguments");
  if (ctx->dim != 2 && ctx->dim != 3) return mrl_fail(MRL_ERR_UNSUPPORTED, "mrl_displacements: Unsupported problem dimension %d", ctx->dim);
  CK(cudaSetDevice(ctx->device));
  return ctx->precision == MRL_F64 ? displacements_impl<double>(ctx, (const double *)F, (double *)out)
                                   : displacements_impl<float>(ctx, (const float *)F, (float *)out);
}
