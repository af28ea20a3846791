// Minimal host emulation of the CUDA execution model - TEST TOOL ONLY.
//
// Lets tests/emu/*.cpp compile the product's kernel headers (marlin_b200/csrc/*.cuh) with
// g++ and run them block by block on the CPU so index math, barriers and shared-memory
// hazards can be checked without a GPU.  Each CUDA thread is a ucontext fiber;
// __syncthreads() yields to the next fiber of the block.  Never linked into the product
// library (libmarlin_b200.so has no CPU path).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };

namespace emu {
inline uint3_emu g_threadIdx, g_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char *g_dyn_smem = nullptr;
inline std::vector<ucontext_t> g_ctx;
inline ucontext_t g_main;
inline std::vector<int> g_state;  // 0 = runnable, 1 = at barrier, 2 = done
inline int g_cur = 0;
inline std::function<void()> g_body;
inline long g_barrier_count = 0;

inline void set_tid(int t) {
  g_threadIdx.x = t % g_blockDim.x;
  g_threadIdx.y = (t / g_blockDim.x) % g_blockDim.y;
  g_threadIdx.z = t / (g_blockDim.x * g_blockDim.y);
}
inline void fiber_entry() {
  g_body();
  g_state[g_cur] = 2;
  swapcontext(&g_ctx[g_cur], &g_main);
}
inline std::vector<int> g_bar_id, g_bar_n;  // barrier a waiting fiber sits at / its thread count
inline void block_at(int id, int n) {
  g_state[g_cur] = 1;
  g_bar_id[g_cur] = id;
  g_bar_n[g_cur] = n;
  int me = g_cur;
  swapcontext(&g_ctx[me], &g_main);
  set_tid(me);
}
inline void syncthreads() { block_at(0, -1); }
// bar.sync id, n: releases when n fibers wait at barrier `id`
inline void named_barrier(int id, int n) { block_at(id, n); }
// cooperative spin (e.g. mbarrier try_wait loops): stay runnable, let the others run
inline void yield() {
  int me = g_cur;
  swapcontext(&g_ctx[me], &g_main);
  set_tid(me);
}
// Run one block: round-robin the fibers; barrier 0 releases when every live fiber reached it,
// a named barrier when its thread count is reached.
inline void run_block(int nthreads, size_t stack_bytes) {
  g_ctx.assign(nthreads, ucontext_t());
  g_state.assign(nthreads, 0);
  g_bar_id.assign(nthreads, 0);
  g_bar_n.assign(nthreads, 0);
  std::vector<std::vector<unsigned char>> stacks(nthreads, std::vector<unsigned char>(stack_bytes));
  for (int t = 0; t < nthreads; ++t) {
    getcontext(&g_ctx[t]);
    g_ctx[t].uc_stack.ss_sp = stacks[t].data();
    g_ctx[t].uc_stack.ss_size = stack_bytes;
    g_ctx[t].uc_link = &g_main;
    makecontext(&g_ctx[t], (void (*)())fiber_entry, 0);
  }
  long idle_rounds = 0;
  while (true) {
    int done = 0;
    for (int t = 0; t < nthreads; ++t) {
      if (g_state[t] == 2) { ++done; continue; }
      if (g_state[t] == 1) continue;
      g_cur = t;
      set_tid(t);
      swapcontext(&g_main, &g_ctx[t]);
      if (g_state[t] == 2) ++done;
    }
    if (done == nthreads) break;
    // release barriers
    int count[128] = {0}, need[128] = {0}, live = nthreads - done, released = 0;
    for (int t = 0; t < nthreads; ++t)
      if (g_state[t] == 1) {
        count[g_bar_id[t]]++;
        need[g_bar_id[t]] = g_bar_n[t] < 0 ? live : g_bar_n[t];
      }
    for (int id = 0; id < 128; ++id)
      if (count[id] && count[id] >= need[id]) {
        ++g_barrier_count;
        ++released;
        for (int t = 0; t < nthreads; ++t)
          if (g_state[t] == 1 && g_bar_id[t] == id) g_state[t] = 0;
      }
    idle_rounds = released ? 0 : idle_rounds + 1;
    if (idle_rounds > 100000) {
      fprintf(stderr, "emu: deadlock (no barrier released for 100000 scheduler rounds)\n");
      abort();
    }
  }
}
// warp shuffle: the 32 fibers of a warp meet at a per-warp barrier, publish, meet again, read
inline double g_shfl_slot[64][32];
template <class V> inline V shfl_sync(V val, int src_lane) {
  const int lin = g_cur, warp = lin / 32, lane = lin % 32;
  static_assert(sizeof(V) <= 8, "shfl_sync: value too large");
  const int nthr = (int)(g_blockDim.x * g_blockDim.y * g_blockDim.z);
  const int wn = nthr - warp * 32 < 32 ? nthr - warp * 32 : 32;  // threads that exist in this warp
  memcpy(&g_shfl_slot[warp][lane], &val, sizeof(V));
  named_barrier(32 + warp, wn);
  V out;
  memcpy(&out, &g_shfl_slot[warp][src_lane & 31], sizeof(V));
  named_barrier(32 + warp, wn);
  return out;
}
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f, size_t stack_bytes = 96 * 1024) {
  std::vector<unsigned char> smem(smem_bytes + 64);
  g_dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 63) & ~uintptr_t(63));
  g_gridDim = grid;
  g_blockDim = block;
  g_body = f;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = {bx, by, bz};
        run_block(block.x * block.y * block.z, stack_bytes);
      }
}
}  // namespace emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define __syncthreads() emu::syncthreads()
#define __shfl_sync(mask, val, lane) emu::shfl_sync(val, lane)
template <class T> inline T __ldg(const T *p) { return *p; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#define __shared__ static
