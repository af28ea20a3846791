// Field reductions behind the postprocessors (sum / min / max / sum of squares).
#include "k_common.cuh"
#include "mrl_internal.h"

namespace mrl {

template <class T, int OP> __global__ void __launch_bounds__(256) k_reduce(const T *in, long long count, double *partials) {
  double acc = OP == MRL_MIN ? 1.0 / 0.0 : (OP == MRL_MAX ? -1.0 / 0.0 : 0.0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)in[i];
    if (OP == MRL_SUM) acc += v;
    else if (OP == MRL_SUMSQ) acc += v * v;
    else if (OP == MRL_MIN) acc = v < acc ? v : acc;
    else acc = v > acc ? v : acc;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, acc, o);
    if (OP == MRL_MIN) acc = w < acc ? w : acc;
    else if (OP == MRL_MAX) acc = w > acc ? w : acc;
    else acc += w;
  }
  __shared__ double sm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = sm[lane & 7];
    for (int o = 4; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, acc, o);
      if (OP == MRL_MIN) acc = w < acc ? w : acc;
      else if (OP == MRL_MAX) acc = w > acc ? w : acc;
      else acc += w;
    }
    if (lane == 0) partials[blockIdx.x] = acc;
  }
}

template <class T>
cudaError_t launch_reduce(const LaunchCtx &lc, int op, const T *in, long long count, double *partials, int nblk) {
  switch (op) {
    case MRL_SUM: k_reduce<T, MRL_SUM><<<nblk, 256, 0, lc.stream>>>(in, count, partials); break;
    case MRL_MIN: k_reduce<T, MRL_MIN><<<nblk, 256, 0, lc.stream>>>(in, count, partials); break;
    case MRL_MAX: k_reduce<T, MRL_MAX><<<nblk, 256, 0, lc.stream>>>(in, count, partials); break;
    default: k_reduce<T, MRL_SUMSQ><<<nblk, 256, 0, lc.stream>>>(in, count, partials); break;
  }
  return cudaGetLastError();
}
template cudaError_t launch_reduce<double>(const LaunchCtx &, int, const double *, long long, double *, int);
template cudaError_t launch_reduce<float>(const LaunchCtx &, int, const float *, long long, double *, int);

// Cross-GPU barrier of the peer-mode slab plan.  Thread s signals rank s (a system-scope release store
// of the epoch into that rank's flag slot for this rank) and then waits for rank s's signal in the local
// slot.  Launched after a pass whose peer stores must be visible before the next pass reads them: the
// stream order + the release/acquire pair give that.  A peer that never arrives trips the timeout
// (~4 s) and traps instead of hanging the device.
__global__ void k_slab_barrier(const unsigned long long *recv_tab, long long flag_off, int rank, int nranks, unsigned long long epoch,
                               long long timeout_cycles) {
  const int s = threadIdx.x;
  if (s >= nranks) return;
  unsigned long long *theirs = reinterpret_cast<unsigned long long *>(recv_tab[s] + flag_off) + rank;
  unsigned long long *mine = reinterpret_cast<unsigned long long *>(recv_tab[rank] + flag_off) + s;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v < epoch && clock64() - t0 > timeout_cycles) __trap();
  } while (v < epoch);
}
cudaError_t launch_slab_barrier(const LaunchCtx &lc, const void *recv_tab, long long flag_off, int rank, int nranks, unsigned long long epoch,
                                long long timeout_cycles) {
  k_slab_barrier<<<1, 32, 0, lc.stream>>>((const unsigned long long *)recv_tab, flag_off, rank, nranks, epoch, timeout_cycles);
  return cudaGetLastError();
}

}  // namespace mrl
