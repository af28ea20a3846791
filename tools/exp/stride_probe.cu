// Development probe: memory-side ceiling of "tile = R rows x 128 B, rows S bytes apart" traffic
// (the access pattern of the x-axis passes) as a function of the row stride S.
// Each CTA copies tiles in -> out with 128-bit accesses: 8 threads cover one 128-byte row segment.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) k_probe(const double2 *in, double2 *out, int rows, long long stride16, int ntiles, int ncb, long long outer16, int blk, long long inner16) {
  const int lane8 = threadIdx.x & 7, r0 = threadIdx.x >> 3, rstep = blockDim.x >> 3;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int o = tile / ncb, cb = tile - o * ncb;
    const long long base = blk ? (long long)(cb / 33) * blk * inner16 + (long long)(cb % 33) * 8 : o * outer16 + (long long)cb * 8;
    const double2 *src = in + base + lane8;
    double2 *dst = out + base + lane8;
    double2 v[8];
    for (int r = r0; r < rows; r += rstep * 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) { int rr = r + u * rstep; if (rr < rows) v[u] = src[blk ? (long long)(rr / blk) * stride16 + (rr % blk) * inner16 : (long long)rr * stride16]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { int rr = r + u * rstep; if (rr < rows) dst[blk ? (long long)(rr / blk) * stride16 + (rr % blk) * inner16 : (long long)rr * stride16] = v[u]; }
    }
  }
}
int main() {
  // {rows, ncols, pad, nouter}: tiles of `rows` x 128 B with a row stride of ncols*16 bytes; the same 1.1 GB
  // is cut into nouter slices so that the stride (and with it the number of 2 MB pages a tile touches)
  // varies while the traffic stays the same
  const int cases[][4] = {{512, 135168, 0, 1}, {512, 264, 0, 512}, {512, 135168, 8, 1}, {512, 135168, 16, 1}, {512, 135168, 32, 1}, {512, 135168, 64, 1}};
  // third entry reused as the block size of a BLOCKED layout [x/blk][y][x%blk][kz]: rows of one block are
  // ncp = 264 elements apart, blocks are ny*blk*ncp apart (same 1.1 GB array, ny = 512)
  for (auto &c : cases) {
    const int rows = c[0];
    const int blk = c[2];
    const long long ncols = c[1], stride16 = blk ? 512LL * blk * 264 : ncols;
    const long long inner16 = 264;
    const int nouter = c[3], ncb = blk ? 512 * 33 : (int)(ncols / 8);
    const int ntiles = ncb * nouter;
    const long long outer16 = (long long)rows * stride16;
    const size_t bytes = blk ? (size_t)512 * 512 * 264 * 16 : (size_t)nouter * rows * stride16 * 16;
    double2 *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 2;
    if (ncols < 1000) { /* y-like: many outer slices, emulate by one big 'rows' dimension */ }
    for (int rep = 0; rep < 2; ++rep) k_probe<<<grid, 512>>>(a, b, rows, stride16, ntiles, ncb, outer16, blk, inner16);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int rep = 0; rep < reps; ++rep) k_probe<<<grid, 512>>>(a, b, rows, stride16, ntiles, ncb, outer16, blk, inner16);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double moved = blk ? 2.0 * 512 * 512 * 264 * 16 : 2.0 * nouter * rows * ncols * 16;
    printf("nouter=%d rows=%d ncols=%lld pad=%d stride=%lld B : %.3f ms  %.0f GB/s  (%s)\n", nouter, rows, ncols, c[2], stride16 * 16, ms, moved / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(a); cudaFree(b);
  }
  return 0;
}
