// Probe: SM-driven device -> pinned-host copy of the Symmetric(:U) staircase (row-major lower triangle), rows written with 128-bit stores,
// `nblocks` CTAs only (the rest of the GPU stays free for the Gram kernel). Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared
// -Xcompiler -fPIC -o libzcprobe.so zero_copy_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void stair_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n, int64_t ld2 /*double2 per row*/, int64_t r_begin,
                                  int64_t r_end) {
  // one warp per row, rows dealt round-robin to the warps of the grid; row i has i + 1 elements -> ceil((i + 1) / 2) double2
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = r_begin + warp0; r < r_end; r += nwarps) {
    const int64_t w2 = (r + 2) / 2;
    const double2* s = src + r * ld2;
    double2* d = dst + r * ld2;
    int64_t k = lane;
    for (; k + 96 < w2; k += 128) {
      const double2 a = s[k], b = s[k + 32], c = s[k + 64], e = s[k + 96];
      d[k] = a; d[k + 32] = b; d[k + 64] = c; d[k + 96] = e;
    }
    for (; k < w2; k += 32) d[k] = s[k];
  }
}

extern "C" int zc_stair_copy(const double* src_dev, double* dst_host_mapped, int64_t n, int64_t ld, int64_t r_begin, int64_t r_end, int nblocks, int threads,
                             void* stream) {
  stair_copy_kernel<<<nblocks, threads, 0, (cudaStream_t)stream>>>((const double2*)src_dev, (double2*)dst_host_mapped, n, ld / 2, r_begin, r_end);
  return (int)cudaGetLastError();
}
