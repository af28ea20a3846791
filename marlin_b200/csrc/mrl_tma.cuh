// marlin_b200 - asynchronous-copy plumbing for sm_100a: mbarrier, TMA (cp.async.bulk and
// cp.async.bulk.tensor), named barriers.  Thin inline-PTX wrappers; the host emulation in
// tests/emu provides synchronous stand-ins so the index logic of the pipelined kernels can
// be checked on the CPU.
#pragma once
#include "mrl_fft.cuh"

#if defined(__CUDACC_RTC__)
// NVRTC (expression-specialised passes): no system headers; the tensor map is an opaque 128-byte blob
typedef unsigned long long uint64_t;
typedef unsigned int uint32_t;
typedef unsigned long size_t;
struct alignas(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
#elif !defined(MRL_EMU)
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched at run time, no -lcuda)
#include <stdint.h>
#endif

namespace mrl {

#if !defined(MRL_EMU)

typedef CUtensorMap TensorMap;
#define MRL_GRID_CONSTANT __grid_constant__

MRL_DI uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

MRL_DI void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (TMA unit)
MRL_DI void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MRL_DI void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MRL_DI void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MRL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MRL_DONE;\n"
      "bra MRL_WAIT;\n"
      "MRL_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// order generic-proxy shared-memory accesses before subsequent async-proxy (TMA) accesses
MRL_DI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 3-D tiled TMA load: box of the tensor map at element coordinates (c0 fastest, c1, c2)
MRL_DI void tma_load_3d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
MRL_DI void tma_load_4d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
MRL_DI void tma_load_5d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// 1-D bulk copy shared -> global (also to peer memory mapped over NVLink): asynchronous, tracked by the issuing
// thread's bulk async-groups.  bytes: multiple of 16, both addresses 16-byte aligned.
MRL_DI void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
MRL_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed group has been read (the buffer may be overwritten)
MRL_DI void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the newest `N` committed groups have completed their global writes
template <int N> MRL_DI void bulk_wait_done() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// order async-proxy global writes (completed bulk stores) before subsequent generic-proxy accesses of this thread
MRL_DI void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
MRL_DI unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
MRL_DI void red_release_sys_add(unsigned long long *p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 1-D bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned)
MRL_DI void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"((uint64_t)gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// barrier among the `nthreads` threads that use id `id` (1..15; 0 is __syncthreads)
MRL_DI void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

#else  // --------------------------------------------------------------- host emulation

struct TensorMap {  // what the emulation needs of a tiled map (rank <= 5; unused trailing dims: dim 1, box 1)
  const unsigned char *base;
  int esize;
  long long dim[5], stride[5];  // stride in bytes (stride[0] = esize)
  int box[5];
};
#define MRL_GRID_CONSTANT

MRL_DI uint32_t smem_u32(const void *p) { return (uint32_t)(uintptr_t)p; }
MRL_DI void mbar_init(uint64_t *bar, int) { *bar = 0; }
MRL_DI void mbar_init_fence() {}
// copies are performed synchronously at issue, so arming a phase also completes it: the
// word counts armed phases, and a waiter spins (cooperatively) until its phase exists
MRL_DI void mbar_expect_tx(uint64_t *bar, uint32_t) { *bar += 1; }
MRL_DI void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (*(volatile uint64_t *)bar == 0 || ((*(volatile uint64_t *)bar - 1) & 1) != parity) emu::yield();
}
MRL_DI void fence_proxy_async() {}
MRL_DI void tma_load_5d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  unsigned char *d = (unsigned char *)smem_dst;
  for (int q = 0; q < tm->box[4]; ++q)
    for (int m = 0; m < tm->box[3]; ++m)
      for (int k = 0; k < tm->box[2]; ++k)
        for (int j = 0; j < tm->box[1]; ++j)
          for (int i = 0; i < tm->box[0]; ++i) {
            const long long a = c0 + i, b = c1 + j, c = c2 + k, e = c3 + m, f = c4 + q;
            const bool in = a < tm->dim[0] && b < tm->dim[1] && c < tm->dim[2] && e < tm->dim[3] && f < tm->dim[4];
            if (in)
              memcpy(d, tm->base + a * tm->stride[0] + b * tm->stride[1] + c * tm->stride[2] + e * tm->stride[3] + f * tm->stride[4], tm->esize);
            else
              memset(d, 0, tm->esize);
            d += tm->esize;
          }
  (void)bar;
}
MRL_DI void tma_load_4d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3) {
  tma_load_5d(smem_dst, tm, bar, c0, c1, c2, c3, 0);
}
MRL_DI void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes) { memcpy(gdst, smem_src, bytes); }
MRL_DI void bulk_commit() {}
MRL_DI void bulk_wait_read_all() {}
template <int N> MRL_DI void bulk_wait_done() {}
MRL_DI void fence_proxy_async_global() {}
MRL_DI unsigned long long ld_acquire_sys(const unsigned long long *p) { return *(volatile const unsigned long long *)p; }
MRL_DI void red_release_sys_add(unsigned long long *p, unsigned long long v) { *p += v; }
MRL_DI void tma_load_3d(void *smem_dst, const TensorMap *tm, uint64_t *bar, int c0, int c1, int c2) {
  tma_load_4d(smem_dst, tm, bar, c0, c1, c2, 0);
}
MRL_DI void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *) { memcpy(smem_dst, gsrc, bytes); }
MRL_DI void named_bar_sync(int id, int nthreads) { emu::named_barrier(id, nthreads); }

#endif

// spin until *p >= expect (system scope); traps after ~4 s so that a lost peer cannot hang the device
MRL_DI void wait_counter(const unsigned long long *p, unsigned long long expect) {
#if defined(MRL_EMU)
  (void)p; (void)expect;
#else
  if (ld_acquire_sys(p) >= expect) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < expect) {
    __nanosleep(64);
    if (clock64() - t0 > 8000000000ll) __trap();
  }
#endif
}

// Barrier policy for RegFFT: one group of threads of a CTA (CtaSync is the whole CTA).
struct GroupBarrier {
  int id, nthreads;
  MRL_DI void sync() const { named_bar_sync(id, nthreads); }
  // every thread orders its generic-proxy accesses before the TMA write one of them issues next
  MRL_DI void sync_release() const {
    fence_proxy_async();
    named_bar_sync(id, nthreads);
  }
};

}  // namespace mrl
