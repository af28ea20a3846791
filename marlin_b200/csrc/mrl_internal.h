// Internal host-side state behind the opaque handles of include/marlin_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/marlin_b200.h"
#include "mrl_launch.h"

int mrl_fail(int code, const char *fmt, ...);
struct mrl_context;
// waits for the context's stream (or the whole device if the context is already destroyed)
void mrl_quiesce(const mrl_context *ctx);

struct mrl_context {
  int device = 0;
  int precision = MRL_F64;
  cudaStream_t stream = 0;
  cudaStream_t owned_stream = nullptr;  // mrl_own_stream
  int sm_count = 148;
  int64_t launches = 0;
  // domain
  int dim = 0;
  int n[3] = {1, 1, 1};   // real shape
  int nr[3] = {1, 1, 1};  // reciprocal shape
  double min[3] = {0, 0, 0}, max[3] = {1, 1, 1};
  std::vector<double> axis_h[3], kaxis_h[3];
  void *axis_dev[3] = {nullptr, nullptr, nullptr};
  void *kaxis_dev[3] = {nullptr, nullptr, nullptr};
  // slab decomposition (nranks == 1: serial)
  int rank = 0, nranks = 1;
  int nyl = 0, nxl = 0, y0 = 0, x0 = 0;  // local extents / first global index (y real, x reciprocal)
  // generic slab decomposition (mrl_domain_set_dist): n / nr above are the LOCAL shapes, gn the global grid
  bool dist = false;
  int gn[3] = {1, 1, 1};
  std::vector<int64_t> ycount, ybegin, xcount, xbegin;  // slab: per rank; pencil: per y-part index (rank % py)
  // pencil decomposition (mrl_domain_set_pencil): rank = iz * py + iy
  bool pencil = false;
  int py = 1, pz = 1;
  std::vector<int64_t> zcount, zbegin, y2count, y2begin;  // per z-part index (rank / py): real-space z, reciprocal ky
  // caches
  std::map<int, void *> tw;
  void *scratch_ptr = nullptr;
  size_t scratch_bytes = 0;
  void *reduce_dev = nullptr;
  void *reduce_host = nullptr;
  // staged transfers (mrl_upload_staged / mrl_download_staged)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in = nullptr, ev_compute = nullptr;
  std::map<const void *, cudaEvent_t> dl_done;  // device buffer -> completion of its last staged download

  mrl::LaunchCtx lc() const { return mrl::LaunchCtx{stream, sm_count}; }
  long long total() const { return (long long)n[0] * n[1] * n[2]; }
  long long rtotal() const { return (long long)nr[0] * nr[1] * nr[2]; }
  int twiddles(int n, const void **out);
  int scratch(size_t bytes, void **out);
};

struct mrl_split_plan {
  mrl_context *ctx = nullptr;
  mrl_split_desc desc;
  void *A = nullptr, *B = nullptr;  // partial spectra (one allocation, B = A + rtotal)
  std::vector<void *> ring;         // history+1 nonlinear-term slots
  int cur = 0, stored = 0;
  int ncp = 0;                      // row pitch of the work spectra (>= n_last/2+1)
  double time = 0.0;                // sub-time seen by an expression nonlinearity
  // mrl_split_substeps: one period (history+1 substeps) of the steady-state sequence captured as a CUDA graph
  cudaGraphExec_t graph = nullptr;
  struct GraphKey {
    const void *c = nullptr;
    double dt = 0, beta[5] = {0, 0, 0, 0, 0};  // (not the time: fused plans reject expressions that read it)
    int nold = -1, cur = -1;
    bool operator==(const GraphKey &o) const {
      for (int i = 0; i < 5; ++i)
        if (beta[i] != o.beta[i]) return false;
      return c == o.c && dt == o.dt && nold == o.nold && cur == o.cur;
    }
  } graph_key;
  int64_t graph_launches = 0;       // kernel launches inside the captured period
};

struct mrl_slab_plan {
  mrl_context *ctx = nullptr;
  mrl_split_desc desc;
  void *send_fwd = nullptr, *recv_fwd = nullptr, *send_bwd = nullptr;  // caller-owned
  std::vector<void *> ring;  // history+1 nonlinear-term slots, staged layout
  int cur = 0, stored = 0;
  int ncp = 0;
  long long field = 0, chunk = 0;
  // peer mode: the exchanges are bulk stores into the peers' blocked staging arrays (mrl_passes_slab.cuh)
  bool owns = false, peer = false;
  void *ret_stage = nullptr;           // S: blocked return staging [kb][nx][nyl][W], written by every rank's fused y pass
  int tk = 0, kb = 0;                  // column-block width W of the blocked layouts, blocks per z row
  std::vector<void *> opened;          // pointers from cudaIpcOpenMemHandle
  void *peer_recv_tab = nullptr;       // device arrays of nranks base pointers: R (recv_fwd) ...
  void *peer_send_tab = nullptr;       // ... and S (ret_stage)
  // arrival counters instead of barriers between the phases (MRL_SLAB_SYNC=flags)
  bool sync_flags = false;
  int kzb_major = 0;
  long long c1_off = 0, c2_off = 0;    // byte offsets of counters1 / counters2 [nranks][kb] behind recv_fwd
  void *flag1_tab = nullptr, *flag2_tab = nullptr;  // device arrays: the peers' counters1 / counters2 bases
  unsigned long long n_forward = 0, n_update = 0;   // phases issued so far (expected counter values)
  // x inverse pass of column block k overlapped with the fused y pass (flags mode): CTAs given to the inverse pass
  int inv_ctas = 0;
  bool inverse_issued = false;
  // forward phase split into y-chunks: the z pass of chunk i+1 (main stream) overlaps the x pass +
  // peer stores of chunk i (aux stream); 1 = one pass each
  int chunks = 4, x_ctas = 96;  // measured on 2 B200: forward phase 1.25 -> 1.11 ms at 512^3 (profiles/r1w_*)
  cudaStream_t s_aux = nullptr;
  std::vector<cudaEvent_t> ev_chunk;
  cudaEvent_t ev_begin = nullptr, ev_done = nullptr;
  long long flag_off = 0;              // byte offset of the barrier flags behind recv_fwd (peer mode)
  unsigned long long epoch = 0;        // barriers issued so far
  // copy mode (MRL_SLAB_EXCHANGE=copy; an experiment kept behind the switch): the passes work on the plain staged layouts
  // and the exchanges are peer-to-peer copies by the copy engines (tools/p2p_probe.py: 760 GB/s per direction with both
  // directions busy, against ~500 GB/s for stores issued by the SMs), issued y-chunk by y-chunk behind the passes
  bool copy = false;
  std::vector<char *> h_peer_recv, h_peer_ret;   // every rank's recv_fwd / landing array of the return exchange (its send_fwd)
  std::vector<cudaStream_t> s_copy;               // one stream per destination rank
  std::vector<cudaEvent_t> ev_copy;
  int y_chunks = 2;                               // fused y pass in x-chunks, so that the return copies trail it
};

// single passes on explicit shapes (generic kernels for any length, pipelined ones where a configuration exists)
//   complex pass along the middle axis of [nouter][n][ncols] (in == out allowed); inverse: conj . FFT . conj, unnormalised
int mrl_pass_strided(mrl_context *ctx, const void *in, void *out, int n, long long ncols, long long nouter, int inverse);
//   last-axis r2c / c2r of `rows` rows of length n (half spectra of n/2+1)
int mrl_pass_zfwd(mrl_context *ctx, const void *in_real, void *out_cplx, long long rows, int n);
int mrl_pass_zinv(mrl_context *ctx, const void *in_cplx, void *out_real, long long rows, int n, double scale);

// Internal batched real transforms on [batch][n0][n1][n2] fields with a last-axis spectrum pitch
// ncp >= n2/2+1 (mrl_fftb_pitch: padded to 128 bytes when every axis runs on the TMA kernels).
// forward: unnormalised; inverse: `work` is transformed in place (destroyed), result * scale.
int mrl_fftb_pitch(const mrl_context *ctx);
// first_axis > 0: the strided axes below it are skipped (left to a fused pass of the caller).
int mrl_fftb_forward(mrl_context *ctx, const void *in_real, void *out_cplx, int batch, int ncp, int first_axis = 0);
int mrl_fftb_strided(mrl_context *ctx, void *spec_cplx, int batch, int ncp, int axis, int inverse);  // one complex pass, in place
// dot_with != nullptr: sum(out * dot_with) rides in the store of the last pass where that pass can carry it; *dot_count
// partial sums land in dot_partials (capacity entries), *dot_count == 0 means the caller has to compute it itself
int mrl_fftb_inverse(mrl_context *ctx, void *work_cplx, void *out_real, int batch, int ncp, double scale, int first_axis = 0,
                     const void *dot_with = nullptr, double *dot_partials = nullptr, int dot_capacity = 0, int *dot_count = nullptr);

// mrl_dist_irfftn with an extra factor on the result (mechanics: sign of the projected field)
int mrl_dist_irfftn_scaled(mrl_dist *d, const void *in_cplx, void *out_real, int batch, double scale);

namespace mrl {
FFTPlanDev make_fft_plan(int n);
template <class T>
cudaError_t launch_reduce(const LaunchCtx &lc, int op, const T *in, long long count, double *partials, int nblk);
// timeout_cycles: a peer that never arrives traps the kernel instead of hanging the device (8e9 ~ 4 s)
cudaError_t launch_slab_barrier(const LaunchCtx &lc, const void *recv_tab, long long flag_off, int rank, int nranks, unsigned long long epoch,
                                long long timeout_cycles = 8000000000ll);
// small layout kernels of the generic distributed transforms (k_dist.cu)
template <class T> cudaError_t launch_real_to_complex(const LaunchCtx &lc, const T *in, cx<T> *out, long long n);
template <class T> cudaError_t launch_complex_real_scale(const LaunchCtx &lc, const cx<T> *in, T *out, long long n, T scale);
template <class T> cudaError_t launch_expand_half(const LaunchCtx &lc, const cx<T> *in, cx<T> *out, long long rows, int n);
//   dst[i0 * d0 + i1 * d1 + k] = src[i0 * s0 + i1 * s1 + k], i0 < n0, i1 < n1, k < w (strides in elements; dst may be peer memory)
template <class T>
cudaError_t launch_copy3d(const LaunchCtx &lc, cx<T> *dst, long long d0, long long d1, const cx<T> *src, long long s0, long long s1, long long n0,
                          long long n1, long long w);
//   a[k][c] = conj a[n - k][c] for n/2 < k < n, c < ncols (Hermitian completion of a half spectrum along the slow axis)
template <class T> cudaError_t launch_hermitian_rows(const LaunchCtx &lc, cx<T> *a, int n, long long ncols);
}  // namespace mrl

// expression-specialised first pass (mrl_expr_zfwd.cu); returns MRL status
int mrl_expr_launch_zfwd(mrl_context *ctx, void *expr, int staged_var, const void *const *inputs, double t, const void *c,
                         void *g_out, void *outC, void *outG, long long rows, int n, int ncp);
// the same on a y-chunk of a slab (rowmap = RowMap{ych, nyl, y0} of mrl_passes.cuh, NULL = all rows), on `stream` (NULL = the
// context's) with a grid of at most sm_count CTAs (0 = all SMs)
int mrl_expr_launch_zfwd_rows(mrl_context *ctx, void *expr, int staged_var, const void *const *inputs, double t, const void *c, void *g_out,
                              void *outC, void *outG, long long rows, int n, int ncp, const int *rowmap, void *stream, int sm_count);
