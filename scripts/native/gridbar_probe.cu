// Micro-benchmarks behind the grid barrier and the scalar sections of eigh_sytrd_reg_kernel (csrc/eigh.cuh): cycles per grid barrier for
// several designs on G CTAs x 256 threads (cooperative launch), the round trip of a strong load of a line another SM has just written,
// and dependent-chain latencies of FP64 add / fma / div / sqrt / rsqrt and of a 64-bit warp reduction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gridbar_probe gridbar_probe.cu ; run: ./gridbar_probe [G]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// mode 0: atomicAdd counter (threadfence + atomicAdd), thread 0 polls with ld.acquire
// mode 1: red.release counter, thread 0 polls relaxed, fence.acq_rel after
// mode 2: packed flags all-to-all, warp 0 polls (independent relaxed loads), fence.acq_rel after
// mode 3: as 2 with flags 128 bytes apart
// mode 4: master-collect: slots 32 B apart polled by warp 0 of CTA 0, go word polled by lane 0 of the others
// mode 5: as 2 with __nanosleep(20) between rounds
// mode 6: as 1 but the counter is split in 8 sub-counters (CTA b adds to counter b % 8), thread 0..7 of warp 0 poll one each
template <int MODE>
__global__ void __launch_bounds__(256, 1) bar_kernel(unsigned long long* bar, double* data, int iters, long long* out) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    const unsigned long long step = it;
    // a little traffic before the barrier, like the real kernel: every warp publishes one double
    if (lane == 0) __stcg(data + (it & 1) * 4096 + b * 8 + (tid >> 5), (double)it);
    __syncthreads();
    if (MODE == 0) {
      if (tid == 0) {
        __threadfence();
        atomicAdd(bar, 1ull);
        while (ld_acquire(bar) < step * G) {}
      }
    } else if (MODE == 1) {
      if (tid == 0) {
        red_release_add(bar, 1ull);
        while (ld_relaxed(bar) < step * G) {}
        fence_acq_rel();
      }
    } else if (MODE == 2 || MODE == 3 || MODE == 5) {
      constexpr int SP = MODE == 3 ? 16 : 1;
      if (tid < 32) {
        if (lane == 0) st_release(bar + (size_t)b * SP, step);
        for (;;) {
          unsigned long long f[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) f[q] = lane + 32 * q < G ? ld_relaxed(bar + (size_t)(lane + 32 * q) * SP) : ~0ull;
          bool all = true;
#pragma unroll
          for (int q = 0; q < 5; ++q) all = all && f[q] >= step;
          if (__all_sync(0xffffffffu, all)) break;
          if (MODE == 5) __nanosleep(20);
        }
        fence_acq_rel();
      }
    } else if (MODE == 4) {
      if (tid < 32) {
        if (b == 0) {
          for (;;) {
            unsigned long long f[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) { const int c = lane + 32 * q; f[q] = (c < G && c > 0) ? ld_relaxed(bar + 4 * c) : ~0ull; }
            bool all = true;
#pragma unroll
            for (int q = 0; q < 5; ++q) all = all && f[q] >= step;
            if (__all_sync(0xffffffffu, all)) break;
          }
          fence_acq_rel();
          if (lane == 0) st_release(bar + 2048, step);
        } else if (lane == 0) {
          st_release(bar + 4 * b, step);
          while (ld_relaxed(bar + 2048) < step) {}
          fence_acq_rel();
        }
      }
    } else if (MODE == 6) {
      if (tid < 32) {
        if (lane == 0) red_release_add(bar + 16 * (b & 7), 1ull);
        const unsigned long long want = step * (unsigned long long)((G - (lane & 7) + 7) / 8);
        for (;;) {
          const unsigned long long f = ld_relaxed(bar + 16 * (lane & 7));
          if (__all_sync(0xffffffffu, f >= want)) break;
        }
        fence_acq_rel();
      }
    }
    __syncthreads();
    // consume what the others published (one strong load per thread)
    const double x = __ldcg(data + (it & 1) * 4096 + (tid % G) * 8 + (tid & 7));
    if (x != (double)it && out) out[1] = -1;  // a stale value: the barrier is broken
  }
  if (tid == 0 && b == 0) out[0] = (clock64() - t0) / iters;
}

__global__ void chain_kernel(double* io, long long* out) {
  double x = io[0], y = io[1];
  const int N = 512;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = x + y;
  long long t1 = clock64();
  out[0] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = fma(x, y, y);
  t1 = clock64();
  out[1] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = y / x;
  t1 = clock64();
  out[2] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = sqrt(x + y);
  t1 = clock64();
  out[3] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = rsqrt(x + y);
  t1 = clock64();
  out[4] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  }
  t1 = clock64();
  out[5] = (t1 - t0) / N;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) x = __drcp_rn(x) + y;
  t1 = clock64();
  out[6] = (t1 - t0) / N;
  io[2] = x;
}

template <int MODE>
long long run(int G, unsigned long long* bar, double* data, long long* dout, int iters) {
  cudaMemset(bar, 0, 4096 * 8);
  cudaMemset(dout, 0, 64);
  void* args[] = {&bar, &data, &iters, &dout};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)bar_kernel<MODE>, dim3(G), dim3(256), args, 0, 0);
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return -1; }
  long long h[2];
  cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
  return h[1] < 0 ? -h[0] : h[0];
}

int main(int argc, char** argv) {
  unsigned long long* bar;
  double* data;
  long long* dout;
  cudaMalloc(&bar, 4096 * 8);
  cudaMalloc(&data, 8192 * 8);
  cudaMalloc(&dout, 64);
  const int iters = 2000;
  printf("{\"unit\": \"SM cycles per iteration (publish + barrier + consume; negative = stale read seen)\", \"cases\": [\n");
  const int Gs[] = {16, 63, 125, 148};
  for (int gi = 0; gi < 4; ++gi) {
    const int G = argc > 1 ? atoi(argv[1]) : Gs[gi];
    printf(" {\"ctas\": %d, \"atomic_counter\": %lld, \"red_release_counter\": %lld, \"flags_all_to_all\": %lld, \"flags_128B_apart\": %lld, "
           "\"master_collect\": %lld, \"flags_nanosleep\": %lld, \"eight_counters\": %lld}%s\n",
           G, run<0>(G, bar, data, dout, iters), run<1>(G, bar, data, dout, iters), run<2>(G, bar, data, dout, iters), run<3>(G, bar, data, dout, iters),
           run<4>(G, bar, data, dout, iters), run<5>(G, bar, data, dout, iters), run<6>(G, bar, data, dout, iters), (gi == 3 || argc > 1) ? "" : ",");
    if (argc > 1) break;
  }
  double hio[3] = {1.25, 1.0000001, 0};
  double* dio;
  cudaMalloc(&dio, 24);
  cudaMemcpy(dio, hio, 24, cudaMemcpyHostToDevice);
  chain_kernel<<<1, 32>>>(dio, dout);
  cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, dout, 56, cudaMemcpyDeviceToHost);
  printf("],\n \"dependent_chain_cycles\": {\"dadd\": %lld, \"dfma\": %lld, \"ddiv\": %lld, \"dsqrt_plus_add\": %lld, \"drsqrt_plus_add\": %lld, "
         "\"warp_sum_f64\": %lld, \"drcp_plus_add\": %lld}}\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  return 0;
}
