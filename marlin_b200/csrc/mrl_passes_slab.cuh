// marlin_b200 - x passes of the multi-GPU slab decomposition with the all-to-all fused in as BULK peer stores.
//
// DomainAction::fftSlab / ifftSlab (src/actions/DomainAction.C:870-938, :941-1019) transform the local axes,
// exchange x-blocks against y-blocks between all ranks (MPI_Isend / MPI_Recv per peer) and transform the remaining
// axis.  Here the exchange is part of the passes on either side of it: a pass stages its result tile in shared memory
// and one thread pushes it into the HBM of the ranks that own the rows with cp.async.bulk (shared -> global over
// NVLink, peer buffers mapped with CUDA IPC).  What a tile holds for one destination rank is laid out CONTIGUOUSLY on
// the receiving side, so every transfer is one nxl*128-byte (8 KB at 512^3 on 8 GPUs) bulk write instead of 128-byte
// row segments written by individual threads (measured in round 1: 128-byte segments reach ~600 GB/s of the 770 GB/s
// a peer copy gets on NVLink 5).
//
// Blocked layouts (W = TK complex columns = one 128-byte segment of a z row, kb = ncp / W column blocks per row):
//   R  receive staging of the forward exchange, per field:  [source rank][kzb][yl][xl][W]
//        written by k_slab_xfwd of every rank, read by k_fused_tma (slab = 2) through a 5-D tensor map
//   S  receive staging of the return exchange:               [kzb][x][yl][W]      (x global)
//        written by k_fused_tma (slab = 2) of every rank, read by k_slab_xinv through a 4-D tensor map
// Arrival counters (optional, `flag_*`): instead of a barrier between the phases every producer tile bumps a
// counter [source rank][kzb] in the destination's memory (red.release.sys) once its bulk writes have completed, and a
// consumer tile waits (ld.acquire.sys) until all contributions to its column block have landed; with the kzb-major
// tile order on both sides the consumer phase trails the producer phase block by block.
#pragma once
#include "mrl_passes_tma.cuh"

namespace mrl {

template <class T> struct SlabXIO {
  int n;                 // nx (transform length)
  int nyl, kb;           // local y extent; column blocks per z row
  int nf;                // fields in the pass (forward: 2, inverse: 1)
  int nranks, rank, nxl;
  int y0, ych;           // forward: only the y-chunk [y0, y0 + ych) of the slab
  int kzb_major;         // tile order: column block slowest (pipelined phases) or y slowest
  T scale;
  // forward: peers' R arrays; inverse: local natural output [x][yl][ncp]
  const unsigned long long *peer_tab;
  long long field;       // elements per field of R
  cx<T> *out;
  long long out_pitch;   // nyl * ncp
  // arrival counters.  forward: flag_tab[s] = base of rank s's counters1 [source][kb], bumped per finished tile;
  // inverse: flag_wait = this rank's counters2 [source][kb], a tile waits for flag_expect from every source
  const unsigned long long *flag_tab;
  const unsigned long long *flag_wait;
  unsigned long long flag_expect;
};

// ======================================================================== forward x pass + bulk scatter
// Input: natural slab [nf][nx][nyl][ncp] through the 3-D tensor map of k_strided_tma (real view [nf][nx][2*nyl*ncp]).
// Shared memory: NS input slots (armed NS tiles ahead) + one staging tile per group.
template <class T, class C, int TK, int NG, int NS>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_slab_xfwd(const MRL_GRID_CONSTANT TensorMap tm, SlabXIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  cx<T> *stage = slots + (size_t)NS * TILE;                                        // [NG][TILE]
  uint64_t *full = reinterpret_cast<uint64_t *>(stage + (size_t)NG * TILE);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
  const int ntiles = io.nf * io.ych * io.kb;
  const int nloc = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ncp = io.kb * TK;

  auto decode = [&](int tile, int &f, int &yl, int &kzb) {
    if (io.kzb_major) {
      kzb = tile / (io.nf * io.ych);
      const int r = tile - kzb * (io.nf * io.ych);
      f = r / io.ych;
      yl = io.y0 + (r - f * io.ych);
    } else {
      f = tile / (io.ych * io.kb);
      const int r = tile - f * (io.ych * io.kb);
      yl = io.y0 + r / io.kb;
      kzb = r % io.kb;
    }
  };
  auto issue = [&](int j) {
    int f, yl, kzb;
    decode(blockIdx.x + j * gridDim.x, f, yl, kzb);
    const int s = j % NS;
    mbar_expect_tx(&full[s], (uint32_t)(TILE * sizeof(cx<T>)));
    MRL_UNROLL
    for (int b = 0; b < NBOX; ++b)
      tma_load_3d(slots + (size_t)s * TILE + b * BOXR * TK, &tm, &full[s], (yl * ncp + kzb * TK) * 2, b * BOXR, f);
  };
  auto signal = [&](int kzb) {  // this thread's bulk writes for a tile of column block kzb have completed
    fence_proxy_async_global();
    for (int i = 0; i < io.nranks; ++i) {
      const int s = (io.rank + 1 + i) % io.nranks;
      red_release_sys_add(reinterpret_cast<unsigned long long *>(io.flag_tab[s]) + (size_t)io.rank * io.kb + kzb, 1ull);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (tid == 0)
    for (int j = 0; j < NS && j < nloc; ++j) issue(j);

  const GroupBarrier bar{1 + g, GT};
  cx<T> *stg = stage + (size_t)g * TILE;
  const SmTile<T, TK> so{stg, col};
  int prev_kzb = -1;
  for (int j = g; j < nloc; j += NG) {
    const int s = j % NS;
    mbar_wait(&full[s], (uint32_t)((j / NS) & 1));
    const SmTile<T, TK> sm{slots + (size_t)s * TILE, col};
    int f, yl, kzb;
    decode(blockIdx.x + j * gridDim.x, f, yl, kzb);
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = sm.ld(t + TP * e);
    bar.sync();
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, [&] {
      if (gt == 0 && j + NS < nloc) issue(j + NS);
    });
    if (gt == 0) bulk_wait_read_all();  // the previous tile's bulk copies have read the staging tile
    bar.sync();
    MRL_UNROLL
    for (int e = 0; e < E; ++e) so.st(t + TP * e, mk<T>(v[e].x * io.scale, v[e].y * io.scale));
    bar.sync_release();                 // staging complete and visible to the async proxy
    if (gt == 0) {
      const size_t chunk = (size_t)io.nxl * TK;
      for (int i = 0; i < io.nranks; ++i) {
        const int d = (io.rank + 1 + i) % io.nranks;  // start with the neighbour: spreads the ingress over the ranks
        cx<T> *dst = reinterpret_cast<cx<T> *>(io.peer_tab[d]) + (long long)f * io.field +
                     (((long long)io.rank * io.kb + kzb) * io.nyl + yl) * (long long)chunk;
        bulk_store_1d(dst, stg + (size_t)d * chunk, (uint32_t)(chunk * sizeof(cx<T>)));
      }
      bulk_commit();
      if (io.flag_tab) {
        if (prev_kzb >= 0) {
          bulk_wait_done<1>();          // everything but the group just committed has landed
          signal(prev_kzb);
        }
        prev_kzb = kzb;
      }
    }
  }
  if (gt == 0) {
    bulk_wait_done<0>();                // the kernel's completion must imply the peers hold the data
    if (io.flag_tab && prev_kzb >= 0) signal(prev_kzb);
    fence_proxy_async_global();
  }
}

// ======================================================================== inverse x pass from the blocked staging
// Input S = [kb][nx][nyl][W] through a 4-D tensor map (box = W columns x 1 y x BOXR rows of x x 1 block);
// output: natural slab [nx][nyl][ncp] by direct 128-byte row-segment stores (local HBM).
template <class T, class C, int TK, int NG, int NS>
__global__ void __launch_bounds__(NG *TK *C::TP, 1)
    k_slab_xinv(const MRL_GRID_CONSTANT TensorMap tm, SlabXIO<T> io, const cx<T> *tw_g) {
  constexpr int N = C::N, TP = C::TP, E = C::E, GT = TK * TP;
  constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
  constexpr int TILE = N * TK;
  MRL_DYN_SMEM(smem_raw);
  const bool nofft = g_debug_nofft != 0;
  cx<T> *slots = reinterpret_cast<cx<T> *>(align128(smem_raw));
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)NS * TILE);
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int col = gt % TK, t = gt / TK;
  TwRegs<T, C> twr;
  twr.init(tw_g, t);
  const int ntiles = io.nyl * io.kb;
  const int nloc = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ncp = io.kb * TK;

  auto decode = [&](int tile, int &yl, int &kzb) {
    if (io.kzb_major) {
      kzb = tile / io.nyl;
      yl = tile - kzb * io.nyl;
    } else {
      yl = tile / io.kb;
      kzb = tile - yl * io.kb;
    }
  };
  auto issue = [&](int j) {
    int yl, kzb;
    decode(blockIdx.x + j * gridDim.x, yl, kzb);
    if (io.flag_wait) {
      for (int s = 0; s < io.nranks; ++s) wait_counter(io.flag_wait + (size_t)s * io.kb + kzb, io.flag_expect);
      fence_proxy_async();  // the acquired peer writes precede the bulk read issued next
    }
    const int s = j % NS;
    mbar_expect_tx(&full[s], (uint32_t)(TILE * sizeof(cx<T>)));
    MRL_UNROLL
    for (int b = 0; b < NBOX; ++b) tma_load_4d(slots + (size_t)s * TILE + b * BOXR * TK, &tm, &full[s], 0, yl, b * BOXR, kzb);
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
    int yl, kzb;
    decode(blockIdx.x + j * gridDim.x, yl, kzb);
    cx<T> v[E];
    MRL_UNROLL
    for (int e = 0; e < E; ++e) v[e] = conj(sm.ld(t + TP * e));
    bar.sync();
    fft_or_skip<T, C>(nofft, v, t, sm, twr, bar, [&] {
      if (gt == 0 && j + NS < nloc) issue(j + NS);
    });
    cx<T> *dst = io.out + (long long)yl * ncp + kzb * TK + col;
    MRL_UNROLL
    for (int e = 0; e < E; ++e) dst[(long long)(t + TP * e) * io.out_pitch] = mk<T>(v[e].x * io.scale, -v[e].y * io.scale);
  }
}

}  // namespace mrl
