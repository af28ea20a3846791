/* marlin_b200 - C ABI of the B200-native spectral time-step path.
 *
 * This is the drop-in boundary: the entry points below are what Marlin's host objects call
 * instead of dispatching libTorch ops.  Each one cites the reference interface it replaces
 * (paths relative to the idaholab/marlin repository root).  Plain pointers and sizes only;
 * all `dev` pointers are CUDA device pointers on the context's device; all calls are
 * asynchronous on the context's stream unless stated otherwise.  Every function returns 0 on
 * success or a negative mrl_status; mrl_last_error() gives the message (thread-local).
 * There is no CPU fallback: without a CUDA device mrl_create() fails.
 *
 * Layout contract (src/actions/DomainAction.C:227-338, :854-867, :1054-1066):
 *   real fields     C-order [nx][ny][nz]           (z fastest), `dim` leading dims used
 *   reciprocal      C-order [nx][ny][nz/2+1] complex (re,im interleaved), half spectrum on the
 *                   LAST spatial axis, forward unnormalised, inverse scaled by 1/N
 *   batched fields  [batch][...] (batch slowest)
 *   axes            cell centred linspace(min+dx/2, max-dx/2, n); reciprocal axes
 *                   2*pi*fftfreq(n,dx) (2*pi*rfftfreq on the last axis), Nyquist kept
 */
#ifndef MARLIN_B200_H
#define MARLIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrl_context mrl_context;
typedef struct mrl_split_plan mrl_split_plan;

enum mrl_status {
  MRL_OK = 0,
  MRL_ERR_INVALID = -1,   /* bad argument */
  MRL_ERR_CUDA = -2,      /* CUDA runtime error */
  MRL_ERR_NO_DEVICE = -3, /* no usable CUDA device (there is no CPU path) */
  MRL_ERR_UNSUPPORTED = -4,
  MRL_ERR_PARSE = -5      /* expression syntax / semantic error */
};
enum mrl_precision { MRL_F64 = 0, MRL_F32 = 1 };

/* ---- context ---------------------------------------------------------------------------
 * Replaces the global device / precision selection of src/utils/MarlinUtils.C:17-59 and
 * src/base/MarlinApp.C:27-54 ([Domain] device_names, floating_precision).               */
int mrl_create(int device, int precision, mrl_context **out);
int mrl_destroy(mrl_context *ctx);
const char *mrl_last_error(void);
const char *mrl_version(void);
int mrl_device_count(int *count); /* CUDA devices visible to this process (device_names / local-rank assignment,
                                     src/actions/DomainAction.C:196-197)                                         */
int mrl_set_stream(mrl_context *ctx, void *cuda_stream); /* run on a caller-owned stream */
int mrl_own_stream(mrl_context *ctx);                    /* create a stream owned by the context and run on it
                                                            (capturable: enables the CUDA-graph paths)        */
int mrl_synchronize(mrl_context *ctx);                   /* block until the stream drains */
int mrl_precision_of(const mrl_context *ctx);
int mrl_launch_count(const mrl_context *ctx, int64_t *count); /* kernels launched so far */

/* ---- domain ----------------------------------------------------------------------------
 * DomainAction ctor + gridChanged(), src/actions/DomainAction.C:94-224, :227-338.
 * n/min/max have 3 entries; entries >= dim are ignored (treated as n=1).                  */
int mrl_domain_set(mrl_context *ctx, int dim, const int64_t *n, const double *min, const double *max);
/* getShape()/getReciprocalShape(), include/actions/DomainAction.h:31-73 (3 entries, 1-padded) */
int mrl_domain_shape(const mrl_context *ctx, int64_t *real_shape, int64_t *reciprocal_shape);
/* getAxis(d)/getReciprocalAxis(d): copies the n[d] (or n[d]/2+1) float64 axis values to host.
 * Pure host arithmetic, usable without a device through mrl_axis_values().                 */
int mrl_domain_axis(const mrl_context *ctx, int d, int reciprocal, double *host_out);
int mrl_axis_values(int64_t n, double min, double max, int reciprocal, int half, double *host_out);

/* ---- device memory (PlainTensorBuffer::init / makeCPUCopy,
 *      src/tensor_buffers/PlainTensorBuffer.C:30-52) ----------------------------------- */
int mrl_malloc(mrl_context *ctx, size_t bytes, void **dev);
int mrl_free(mrl_context *ctx, void *dev);
int mrl_memset(mrl_context *ctx, void *dev, int value, size_t bytes);
int mrl_upload(mrl_context *ctx, void *dev, const void *host, size_t bytes);   /* async H2D */
int mrl_download(mrl_context *ctx, void *host, const void *dev, size_t bytes); /* async D2H */
int mrl_copy(mrl_context *ctx, void *dst_dev, const void *src_dev, size_t bytes);
/* Staged transfers on the context's own copy streams, ordered against the compute stream by
 * events so that they overlap the kernels of neighbouring steps (the reference overlaps output
 * with compute through makeCPUCopy + one std::thread per output object,
 * src/problems/TensorProblem.C:225-240, src/tensor_outputs/TensorOutput.C:44-81).
 *  mrl_download_staged: D2H of `dev` as it is once the work queued so far has finished; returns
 *    at once; the compute stream may run ahead (but must not overwrite `dev` - use a second buffer).
 *  mrl_upload_staged: H2D into `dev` without waiting for the compute stream (it waits only for a
 *    staged download of the same `dev`); work queued on the compute stream afterwards sees the data.
 *  mrl_staged_wait: blocks the host until every staged transfer has completed.
 * Host memory should be pinned.                                                            */
int mrl_upload_staged(mrl_context *ctx, void *dev, const void *host, size_t bytes);
int mrl_download_staged(mrl_context *ctx, void *host, const void *dev, size_t bytes);
int mrl_staged_wait(mrl_context *ctx);

/* ---- FFT: DomainAction::fft / ifft (include/actions/DomainAction.h:75-76;
 *      fftSerial src/actions/DomainAction.C:854-867, ifft :1054-1066) and the
 *      ForwardFFT / InverseFFT operators (src/tensor_computes/PerformFFT.C:34-40).
 *      in and out must not overlap.  `batch` leading (slowest) fields.                    */
int mrl_rfftn(mrl_context *ctx, const void *in_real_dev, void *out_cplx_dev, int batch);
int mrl_irfftn(mrl_context *ctx, const void *in_cplx_dev, void *out_real_dev, int batch);

/* ---- k-space constants: ReciprocalLaplacianFactor::computeBuffer
 *      (src/tensor_computes/ReciprocalLaplacianFactor.C:30, kind 0: -k2*factor) and
 *      ReciprocalLaplacianSquareFactor::computeBuffer
 *      (src/tensor_computes/ReciprocalLaplacianSquareFactor.C:31, kind 1: k2*k2*factor).
 *      out: real, reciprocal shape.                                                        */
enum mrl_kfactor_kind { MRL_KFACTOR_LAPLACIAN = 0, MRL_KFACTOR_LAPLACIAN_SQUARE = 1 };
int mrl_kfactor(mrl_context *ctx, int kind, double factor, void *out_real_dev);

/* ---- un-fused pointwise pieces of the solver (generic path) ----------------------------
 * out = a*b, a real / b complex, reciprocal shape (ParsedCompute 'Mbar*mubar').           */
int mrl_mul_real_complex(mrl_context *ctx, const void *a_real_dev, const void *b_cplx_dev, void *out_cplx_dev);
/* ubar = (cbar + dt*beta0*N + sum_i dt*beta_{i+1}*Nold[i]) / (1 - dt*L); L may be NULL.
 * AdamsBashforthMoulton::substep, src/tensor_solver/AdamsBashforthMoulton.C:94-99.        */
int mrl_ab_update(mrl_context *ctx, void *ubar_dev, const void *cbar_dev, const void *N_dev, const void *L_real_dev,
                  double dt, const double *beta, int nold, const void *const *Nold_dev);

/* Coupled semi-implicit update: solves, for every wavevector, (I - dt*L) ubar = rhs with a dense
 * real nvar x nvar operator (LU with partial pivoting) - the batched at::linalg_solve of
 * AdamsBashforthMoultonCoupled::substep, src/tensor_solver/AdamsBashforthMoultonCoupled.C:131-171.
 * L_real_dev[r*nvar + c] is the reciprocal-space buffer that multiplies unknown c in equation r
 * (NULL = zero).  drop_imag != 0 reproduces the reference's cast of the right-hand side to the real
 * dtype of the first linear operator (:141,167): only Re(rhs) enters the solve.  out may alias rhs.
 * nvar <= 6.                                                                                */
int mrl_coupled_solve(mrl_context *ctx, int nvar, const void *const *L_real_dev, const void *const *rhs_cplx_dev,
                      void *const *out_cplx_dev, double dt, int drop_imag);

/* Broyden inverse-Jacobian iteration per wavevector (BroydenSolver::substep,
 * src/tensor_solver/BroydenSolver.C:118-168).  M: nvar*nvar complex reciprocal-space fields,
 * component major ([r*nvar + c][points]), owned by the caller and kept between calls.
 *   step  : sk = -M R;  unew = u + 0.5*sk                         (the reference's fixed 0.5)
 *   update: yk = Rnew - R;  d = sk^T yk;  M += (sk - M yk) sk^T / d  where |d| > 1e-12
 * nvar <= 6.                                                                                  */
int mrl_broyden_step(mrl_context *ctx, int nvar, const void *M_cplx_dev, const void *const *R_cplx_dev, const void *const *u_cplx_dev,
                     void *const *sk_out_dev, void *const *unew_out_dev);
int mrl_broyden_update(mrl_context *ctx, int nvar, void *M_cplx_dev, const void *const *sk_cplx_dev, const void *const *R_cplx_dev,
                       const void *const *Rnew_cplx_dev);

/* ---- reductions behind the postprocessors (src/postprocessors/Tensor*Postprocessor.C) --
 * Synchronous: returns the value on the host.                                             */
enum mrl_reduce_op { MRL_SUM = 0, MRL_MIN = 1, MRL_MAX = 2, MRL_SUMSQ = 3 };
int mrl_reduce(mrl_context *ctx, int op, const void *real_dev, int64_t count, double *host_out);

/* ---- runtime expressions: ParsedCompute (src/tensor_computes/ParsedCompute.C:24-47 parameters,
 *      :50-181 constructor, :184-265 computeBuffer) and ParsedJITTensor
 *      (src/utils/ParsedJITTensor.C:22-156: parse, differentiate, compile, eval).
 * The expression is parsed, differentiated w.r.t. `derivatives` in order, simplified, lowered to
 * ONE CUDA kernel and compiled for sm_100a with NVRTC.  Grammar, simplification and derivative rules
 * are the reference's (include/utils/MarlinExpressionParser.h:383-427,
 * src/utils/MarlinExpressionParser.C:51-1104).                                              */
typedef struct mrl_expr mrl_expr;
enum mrl_var_layout {
  MRL_VAR_REAL = 0,          /* real field, real shape [nx][ny][nz]            */
  MRL_VAR_RECIP_REAL = 1,    /* real field, reciprocal shape [nx][ny][nz/2+1]   */
  MRL_VAR_RECIP_COMPLEX = 2, /* complex field, reciprocal shape                 */
  MRL_VAR_SCALAR = 3,        /* one real value in device memory (0-d tensor)    */
  MRL_VAR_REAL_COMPLEX = 4   /* complex field, real shape                       */
};
enum mrl_expand { MRL_EXPAND_NONE = 0, MRL_EXPAND_REAL = 1, MRL_EXPAND_RECIPROCAL = 2 };
typedef struct mrl_expr_desc {
  const char *expression;
  int nvars;                         /* `inputs`                                             */
  const char *const *var_names;
  const int *var_layouts;            /* mrl_var_layout per input (NULL = all MRL_VAR_REAL)    */
  int nderivatives;                  /* `derivatives`, applied in order                      */
  const char *const *derivatives;
  int nconstants;                    /* `constant_names` with already evaluated values       */
  const char *const *constant_names;
  const double *constant_values;
  int extra_symbols;                 /* x y z kx ky kz k2 t pi e i                           */
  int expand;                        /* mrl_expand                                           */
} mrl_expr_desc;
int mrl_expr_compile(mrl_context *ctx, const mrl_expr_desc *desc, mrl_expr **out);
int mrl_expr_destroy(mrl_expr *e);
/* space: 0 = a single value (all-constant expression), 1 = real shape, 2 = reciprocal shape */
int mrl_expr_result(const mrl_expr *e, int *space, int *is_complex);
/* toString() of the AST after derivatives + simplification (reference formatting) */
int mrl_expr_string(const mrl_expr *e, char *buf, size_t cap);
/* out[p] = f(inputs[0][p], ..., t); inputs in `inputs` order (unused ones may be NULL) */
int mrl_expr_eval(mrl_expr *e, const void *const *inputs_dev, double t, void *out_dev);
/* Host-only pieces (no device needed): the simplified string; a `constant_expressions` value
 * (libMesh FParser stand-in, ParsedCompute.C:104-123; may use pi, e and the given constants);
 * and a dry run that generates the CUDA source and compiles it with NVRTC for sm_100a.     */
int mrl_expr_simplified(const mrl_expr_desc *desc, char *buf, size_t cap);
int mrl_expr_constant(const char *expression, int nconst, const char *const *names, const double *values, double *out);
int mrl_expr_check(const mrl_expr_desc *desc, int precision, char *cuda_source_buf, size_t cap);
/* same dry run for the first pass of the fused split plan specialised for the expression (last-axis
 * length n; staged_var as mrl_split_desc.nonlin_var) */
int mrl_expr_check_fused(const mrl_expr_desc *desc, int precision, int n, int staged_var);

/* ---- fused semi-implicit substep (one solver variable) ---------------------------------
 * Replaces, for the canonical split-operator pattern
 *     g   = F(c)                      ParsedCompute   (src/tensor_computes/ParsedCompute.C:184)
 *     g^  = fft(g), c^ = fft(c)       ForwardFFT      (src/tensor_computes/PerformFFT.C:34)
 *     N   = Mbar * g^                 ParsedCompute
 *     u^  = (c^ + sum dt*beta*N_i)/(1 - dt*L);  c = ifft(u^)
 *                                     AdamsBashforthMoulton::substep
 *                                     (src/tensor_solver/AdamsBashforthMoulton.C:60-101)
 * the ~15 libTorch launches of one substep by five HBM passes (three in 2-D).            */
enum mrl_nonlin_kind {
  MRL_NONLIN_DOUBLE_WELL = 0, /* F(c) = d/dc[A (c-a)^2 (b-c)^2]; params = {A,a,b} */
  MRL_NONLIN_EXPR = 1         /* F given by a compiled expression handle (mrl_expr) */
};
typedef struct mrl_split_desc {
  int nonlin_kind;
  double nonlin_params[4];
  void *nonlin_expr;        /* mrl_expr*, when nonlin_kind == MRL_NONLIN_EXPR */
  int M_closed_form;        /* 1: Mbar = -k2*M_factor computed on the fly; 0: read M_real_dev;
                               2: Mbar = 1 (the nonlinear term is fft(g) itself)              */
  double M_factor;
  const void *M_real_dev;   /* real, reciprocal shape */
  int has_L;                /* linear_reciprocal given ("0" in the input = none) */
  int L_closed_form;        /* 1: L = k2*k2*L_factor on the fly; 0: read L_real_dev */
  double L_factor;
  const void *L_real_dev;
  int history;              /* old nonlinear terms kept = max(predictor_order, corrector_order) - 1 */
  void *g_out_real_dev;     /* optional: receives g = F(c) each substep (NULL = not materialised) */
  /* MRL_NONLIN_EXPR only: the expression's inputs, in its `inputs` order.  nonlin_var is the index
   * of the solver variable among them (its value arrives through the pass's staged tile; -1: the
   * variable is not an input); the other inputs are real-space real fields read in place.      */
  int nonlin_var;
  const void *nonlin_inputs_dev[16];
} mrl_split_desc;

int mrl_split_plan_create(mrl_context *ctx, const mrl_split_desc *desc, mrl_split_plan **out);
int mrl_split_plan_destroy(mrl_split_plan *plan);
/* One predictor substep, c updated in place.  beta[0] multiplies the new nonlinear term,
 * beta[1..nold] the stored old ones (newest first); nold <= states currently stored.      */
int mrl_split_substep(mrl_split_plan *plan, void *c_real_dev, double dt, const double *beta, int nold);
/* The same substep in two halves, for solvers with several coupled variables
 * (SplitOperatorBase::getVariables, src/tensor_solver/SplitOperatorBase.C:39-64): the reference
 * evaluates the whole root compute (every variable's nonlinearity, from the OLD fields) before it
 * updates any variable, so call mrl_split_forward (passes P1-P2) on every plan first, then
 * mrl_split_finish (P3-P5, writes c) on every plan.                                          */
/* `count` substeps, each followed by mrl_split_advance_state, all with the same dt / beta / nold
 * (the steady state inside one MOOSE step of TensorSolver::computeBuffer, TensorSolver.C:93-110).
 * On a capturable stream the periodic part runs as a replayed CUDA graph (small grids are launch bound).
 * Results are identical to calling mrl_split_substep + mrl_split_advance_state `count` times.        */
int mrl_split_substeps(mrl_split_plan *plan, void *c_real_dev, double dt, const double *beta, int nold, int count);
int mrl_split_forward(mrl_split_plan *plan, const void *c_real_dev);
int mrl_split_finish(mrl_split_plan *plan, void *c_real_dev, double dt, const double *beta, int nold);
/* sub-time `t` seen by an MRL_NONLIN_EXPR expression (TensorSolver.C:95, _sub_time) */
int mrl_split_set_time(mrl_split_plan *plan, double t);
/* TensorBuffer<T>::advanceState (include/tensor_buffers/TensorBuffer.h:64-79): the newest
 * nonlinear term becomes old state 0.  Returns the number of stored states through *stored. */
int mrl_split_advance_state(mrl_split_plan *plan, int *stored);
int mrl_split_clear_states(mrl_split_plan *plan);
/* Same as mrl_split_substep but brackets every pass with CUDA events on the context's stream
 * and returns the per-pass device times in milliseconds (pass_ms has
 * mrl_split_launches_per_substep() entries).  Synchronous; for measurement only.          */
int mrl_split_substep_timed(mrl_split_plan *plan, void *c_real_dev, double dt, const double *beta, int nold,
                            float *pass_ms);
/* HBM passes / kernels one substep launches (for bench bookkeeping) */
int mrl_split_launches_per_substep(const mrl_split_plan *plan);

/* ---- de Geus finite-strain FFT mechanics ------------------------------------------------
 * FFTMechanics (src/tensor_computes/FFTMechanics.C:48-163) with the HyperElasticIsotropic
 * constitutive model (src/tensor_computes/HyperElasticIsotropic.C:42-52) and the matrix-free CG
 * of include/utils/MarlinUtils.h:57-131.  The domain is 3-D or 2-D; tensors are D x D with
 * D = dim (test/tests/mechanics/mech3d.i, mech.i).  Tensor fields are COMPONENT-MAJOR here:
 * [D*D][nx][ny][nz] with component c = D*i + j; the reference layout [nx][ny][nz][D][D] is
 * converted with mrl_components().  Ghat4 (D^4 complex values per wavevector) and C4 / K4 (D^4
 * values per voxel) are never materialised: the contractions are evaluated in closed form from
 * q and from (F, K, mu).  K and mu are real [nx][ny][nz] fields owned by the caller.        */
typedef struct mrl_mech_plan mrl_mech_plan;
typedef struct mrl_mech_desc {
  double l_tol;       /* CG relative tolerance on |b|                    (FFTMechanics.C:33) */
  int64_t l_max_its;  /* <= 0: number of cells                            (:63-64)            */
  double nl_rel_tol, nl_abs_tol;
  int nl_max_its;
} mrl_mech_desc;
typedef struct mrl_mech_stats {
  int newton_iterations, cg_solves, cg_iterations_total;
  int cg_iterations[64];
  double final_rnorm, final_anorm;
} mrl_mech_stats;
int mrl_mech_plan_create(mrl_context *ctx, const mrl_mech_desc *desc, const void *K_real_dev, const void *mu_real_dev,
                         mrl_mech_plan **out);
int mrl_mech_plan_destroy(mrl_mech_plan *plan);
/* Mechanics on a decomposed domain (a context set up with mrl_domain_set_dist / mrl_domain_set_pencil; FFTMechanics over
 * DomainAction::fft / ifft in FFT_SLAB / FFT_PENCIL mode, src/tensor_computes/FFTMechanics.C:96-163): the transforms of
 * the plan go through `dist`, the Green projection acts on the rank's wavevectors, and every inner product of the CG /
 * Newton recurrences (include/utils/MarlinUtils.h:57-131) is summed over the ranks through `allreduce_sum` (in place on
 * `count` host doubles, 0 = success; the host objects pass their communicator).  Collective: all ranks solve together. */
typedef int (*mrl_allreduce_fn)(void *user, double *values, int count);
struct mrl_dist;
int mrl_mech_plan_set_dist(mrl_mech_plan *plan, struct mrl_dist *dist, mrl_allreduce_fn allreduce_sum, void *user);
/* P = F . S(F)                      HyperElasticIsotropic::computeBuffer                    */
int mrl_mech_constitutive(mrl_mech_plan *plan, const void *F_dev, void *P_dev);
/* out = irfftn( Ghat4 : rfftn(A) )  FFTMechanics.C:104-105                                   */
int mrl_mech_apply_G(mrl_mech_plan *plan, const void *A_dev, void *out_dev);
/* out = G( K4(F) : x )              FFTMechanics.C:107-112 (the CG operator)                 */
int mrl_mech_apply_GK(mrl_mech_plan *plan, const void *F_dev, const void *x_dev, void *out_dev);
/* One FFTMechanics::computeBuffer: F is updated in place (F + applied strain + Newton
 * increments), P receives the stress of the final state.  applied9: row-major D x D applied
 * macroscopic strain (D*D values) or NULL.  Synchronous (iteration counts depend on
 * device-side norms).                                                                        */
int mrl_mech_solve(mrl_mech_plan *plan, void *F_dev, const double *applied9, void *P_dev, mrl_mech_stats *stats);
/* von Mises stress of a component-major D x D stress field (ComputeVonMisesStress::computeBuffer,
 * src/tensor_computes/ComputeVonMisesStress.C:30-67; 3-D and 2-D forms as coded there).        */
int mrl_von_mises(mrl_context *ctx, const void *stress_dev, void *out_real_dev);
/* Displacement field of a periodic deformation gradient on the (n+1)^D NODAL grid
 * (ComputeDisplacements::computeBuffer, src/tensor_computes/ComputeDisplacements.C:53-107):
 * u = (<F> - I) X + irfftn( rfftn(F - <F>) q (-i) / |q|^2 ), then (bi/tri)linear resampling with
 * align_corners = true.  F: component-major [D*D][cells]; out: component-major [D][nodes].
 * Synchronous (uses device reductions for <F>).                                               */
int mrl_displacements(mrl_context *ctx, const void *F_dev, void *out_nodal_dev);
/* [n][ncomp] (components fastest, the reference layout) <-> [ncomp][n]; in != out.           */
int mrl_components(mrl_context *ctx, const void *in_dev, void *out_dev, int64_t n, int ncomp, int to_component_major);

/* ---- multi-GPU slab decomposition -------------------------------------------------------
 * DomainAction::partitionSlabs / partitionHepler (src/actions/DomainAction.C:511-566,
 * include/actions/DomainAction.h:249-280) and fftSlab / ifftSlab (:870-1019): real space is
 * split along y, reciprocal space along x, z is never split.  One context per rank (= per GPU).
 * Differences from the reference, stated here because they are visible at the boundary:
 *   - the half spectrum (r2c on z) travels, not the reference's full c2c spectrum: same
 *     results, half the bytes, and the local reciprocal block is an x-slice of the serial layout;
 *   - the exchange itself is done by the host between the three phases below (NCCL all-to-all
 *     through torch.distributed in this repository, MPI_Alltoall in Marlin): chunks are
 *     contiguous in both directions, so no pack / unpack pass exists;
 *   - the fused path needs nx % nranks == 0 and ny % nranks == 0 (equal slabs).            */
/* partitionHepler: weighted split of `total` layers over nranks (weights NULL = equal). Host only. */
int mrl_partition(int64_t total, int nranks, const double *weights, int64_t *count);
int mrl_domain_set_slab(mrl_context *ctx, int dim, const int64_t *n, const double *min, const double *max, int rank,
                        int nranks);
/* local shapes and first global index: real [nx][ny/P][nz], reciprocal [nx/P][ny][nz/2+1] */
int mrl_domain_local(const mrl_context *ctx, int64_t *real_shape, int64_t *real_begin, int64_t *recip_shape,
                     int64_t *recip_begin);

/* ---- generic decomposed transforms (any grid size, unequal parts) -------------------------------------
 * Behind DomainAction::fft / ifft for [Domain] parallel_mode = FFT_SLAB and FFT_PENCIL: one process per GPU.
 *
 * FFT_SLAB  (DomainAction::partitionSlabs / fftSlab / ifftSlab, src/actions/DomainAction.C:511-566, :870-938,
 *   :941-1019; 2-D and 3-D): real space split along y, reciprocal space along x with the slab sizes of partitionHepler
 *   (include/actions/DomainAction.h:249-280; `weights` = [Domain] device_weights, NULL = equal).
 *   mrl_domain_set_dist makes the context's domain the LOCAL one: real shape [nx][ny_local](,[nz]), reciprocal
 *   shape [nx_local][ny](,[nz/2+1]) - an x-slice of the serial-mode layout (half spectrum on the last axis; the
 *   reference exchanges the full c2c spectrum, twice the bytes, same fields).
 * FFT_PENCIL (partitionPencils / fftPencil / ifftPencil, :569-742, :1022-1047, :1106-1404; 3-D, nranks = Py * Pz with
 *   the reference's choice of factors): rank = iz * Py + iy holds real [nx][ny / Py][nz / Pz] and reciprocal
 *   [(nx/2+1) / Py][ny / Pz][nz], half spectrum on x, exactly the reference's layout and k-axes.
 *
 * In both modes the device axes are the local slices, so that every pointwise entry point (mrl_kfactor,
 * mrl_expr_eval, mrl_ab_update, mrl_reduce ...) works on a rank's part unchanged.  mrl_rfftn / mrl_irfftn are
 * refused on such a context; the transforms are mrl_dist_rfftn / mrl_dist_irfftn, whose exchanges are strided
 * peer-to-peer copies over NVLink into staging buffers shared through CUDA IPC, bracketed by device-side barriers. */
typedef struct mrl_dist mrl_dist;
int mrl_domain_set_dist(mrl_context *ctx, int dim, const int64_t *n, const double *min, const double *max, int rank,
                        int nranks, const double *weights);
int mrl_domain_set_pencil(mrl_context *ctx, int dim, const int64_t *n, const double *min, const double *max, int rank,
                          int nranks);
/* the factorisation nranks = Py * Pz partitionPencils chooses for an nx x ny x nz grid (DomainAction.C:574-613: both
 * factors > 1 and fitting the domain, as close to a square as possible, the first such pair in its search order).
 * Host only; MRL_ERR_INVALID with the reference's message when there is none.                                */
int mrl_pencil_factors(int nranks, const int64_t *n, int *py, int *pz);
/* slab mode: global grid size and the slab table: counts / begins of the real-space y slabs and the reciprocal x
 * slabs of every rank (each array nranks entries; any pointer may be NULL)                                  */
int mrl_dist_partition(const mrl_context *ctx, int64_t *n_global, int64_t *y_count, int64_t *y_begin, int64_t *x_count,
                       int64_t *x_begin);
/* any mode: [begin, end) of `rank`'s part in real and in reciprocal space, 3 entries each (DomainAction::getLocalBounds,
 * :1543-1556; any pointer may be NULL)                                                                      */
int mrl_dist_bounds(const mrl_context *ctx, int rank, int64_t *real_begin, int64_t *real_end, int64_t *recip_begin,
                    int64_t *recip_end);
int mrl_dist_create(mrl_context *ctx, mrl_dist **out);
int mrl_dist_destroy(mrl_dist *d);
/* This rank's record for the peers (CUDA IPC handle of its shared staging allocation + the offsets of the staging
 * areas: MRL_DIST_IPC_BYTES bytes) / import of all ranks' records in rank order (nranks x MRL_DIST_IPC_BYTES); the
 * host objects exchange them over their rendezvous (host/shim/comm.h)                                        */
#define MRL_DIST_IPC_BYTES 128
int mrl_dist_ipc_export(mrl_dist *d, void *record);
int mrl_dist_ipc_import(mrl_dist *d, const void *all_records);
/* DomainAction::fft / ifft on `batch` fields (batch slowest): local real <-> local reciprocal part; forward
 * unnormalised, inverse scaled by 1/N.  Collective: every rank must call with the same batch.               */
int mrl_dist_rfftn(mrl_dist *d, const void *in_real_dev, void *out_cplx_dev, int batch);
int mrl_dist_irfftn(mrl_dist *d, const void *in_cplx_dev, void *out_real_dev, int batch);

typedef struct mrl_slab_plan mrl_slab_plan;
/* Element counts (complex elements) of the exchange buffers the caller must allocate:
 *   send_fwd : 2*field   two fields back to back, each [nx][ny/P][pitch]; rank s's chunk of field f
 *                        starts at f*field + s*chunk
 *   recv_fwd : 2*field   each [P][nx/P][ny/P][pitch] (chunk s = what rank s sent)
 *   send_bwd : field     same staged layout; chunk s goes back to rank s
 *   (the return exchange lands in field 0 of send_fwd)                                    */
int mrl_slab_sizes(const mrl_context *ctx, int64_t *field_elems, int64_t *chunk_elems, int *pitch);
int mrl_slab_plan_create(mrl_context *ctx, const mrl_split_desc *desc, void *send_fwd_dev, void *recv_fwd_dev,
                         void *send_bwd_dev, mrl_slab_plan **out);
/* Peer mode - the all-to-all fused into the passes: buffers are owned by the library and shared
 * through CUDA IPC; phase 1's x pass stages every result tile in shared memory and pushes it with bulk
 * asynchronous copies into a blocked staging layout in the HBM of the rank that owns its x block,
 * phase 2's fused pass pushes its rows into the owners' y slabs the same way (over NVLink, overlapped
 * with the pass's own HBM traffic), so no exchange call, no staging copy and no pack/unpack exist.
 * The nonlinearity may be the built-in double well or a compiled expression (MRL_NONLIN_EXPR: it is
 * compiled into the first pass as on one GPU); the k-space factors must be closed forms; every axis
 * 128 / 256 / 512 / 1024 points, equal slabs.  The context may be set up with mrl_domain_set_slab or
 * with mrl_domain_set_dist (the host objects' FFT_SLAB mode).  Protocol per substep:
 *   mrl_slab_forward -> cross-rank barrier -> mrl_slab_update -> cross-rank barrier -> mrl_slab_inverse
 * handles: 2 x 64 bytes per rank (cudaIpcMemHandle_t of the return staging, recv_fwd).
 * MRL_SLAB_EXCHANGE=copy (environment, read at plan creation) selects an alternative kept for
 * comparison: plain staged layouts and peer copies by the copy engines (slower on the measured boxes). */
int mrl_slab_plan_create_peer(mrl_context *ctx, const mrl_split_desc *desc, mrl_slab_plan **out);
int mrl_slab_ipc_export(mrl_slab_plan *plan, void *handles_128_bytes);
int mrl_slab_ipc_import(mrl_slab_plan *plan, const void *all_handles /* nranks x 128 bytes, rank order */);
/* Peer mode only: stream-ordered barrier across the ranks of the plan (flags in the IPC-mapped
 * buffers, system-scope release/acquire; no host involvement, no NCCL call).  Every rank must call
 * it the same number of times.  Separates a pass that stores into the peers' memory from the pass
 * that reads it.                                                                              */
int mrl_slab_barrier(mrl_slab_plan *plan);
int mrl_slab_plan_destroy(mrl_slab_plan *plan);
/* phase 1: z r2c of (c + i F(c)) and x forward on the local slab -> send_fwd                */
int mrl_slab_forward(mrl_slab_plan *plan, const void *c_real_dev);
/* phase 2 (after the forward exchange): y forward, semi-implicit update, y inverse -> send_bwd */
int mrl_slab_update(mrl_slab_plan *plan, double dt, const double *beta, int nold);
/* phase 3 (after the return exchange into field 0 of send_fwd): x inverse, z c2r -> c        */
int mrl_slab_inverse(mrl_slab_plan *plan, void *c_real_dev);
int mrl_slab_advance_state(mrl_slab_plan *plan, int *stored);

#ifdef __cplusplus
}
#endif
#endif /* MARLIN_B200_H */
