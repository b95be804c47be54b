// Microbenchmark: throughput of scattered global reductions (no return value) on B200, by operand type and
// footprint.  Every lane adds to a pseudo-random element of an array; one 32-byte sector per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate red_rate.cu && ./red_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <typename T, int MODE>
__global__ void red_kernel(T *a, uint32_t mask, int iters, int stride_elems) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    s = hash32(s + 0x9e3779b9u * (uint32_t)i);
    const size_t k = (size_t)(s & mask) * stride_elems;
    if (MODE == 0) atomicAdd(a + k, (T)1);            // RED
    else if (MODE == 1) { T v = a[k]; if (v == (T)-12345) a[k] = 0; }   // plain load
    else a[k] = (T)i;                                  // plain store
  }
}

template <typename T, int MODE>
float run(T *buf, size_t n_slots, int stride_elems, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256;
  red_kernel<T, MODE><<<blocks, threads>>>(buf, (uint32_t)(n_slots - 1), 8, stride_elems);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  red_kernel<T, MODE><<<blocks, threads>>>(buf, (uint32_t)(n_slots - 1), iters, stride_elems);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return (float)((double)blocks * threads * iters / (ms * 1e-3) / 1e9);
}

int main() {
  void *buf; cudaMalloc(&buf, (size_t)1 << 30); cudaMemset(buf, 0, (size_t)1 << 30);
  const int iters = 2000;
  printf("%-28s %12s %12s %12s %12s\n", "footprint (one slot / 32 B)", "f64 RED", "f32 RED", "u64 RED", "u32 RED");
  for (int lg = 17; lg <= 24; ++lg) {   // 4 MB .. 512 MB of 32-byte slots
    size_t slots = (size_t)1 << lg;
    printf("%8.0f MB                  %9.1f G/s %9.1f G/s %9.1f G/s %9.1f G/s\n", slots * 32 / 1048576.0,
           run<double, 0>((double *)buf, slots, 4, iters), run<float, 0>((float *)buf, slots, 8, iters),
           run<unsigned long long, 0>((unsigned long long *)buf, slots, 4, iters),
           run<unsigned int, 0>((unsigned int *)buf, slots, 8, iters));
  }
  printf("\n%-28s %12s %12s\n", "footprint", "f64 load", "f64 store");
  for (int lg = 17; lg <= 24; lg += 1) {
    size_t slots = (size_t)1 << lg;
    printf("%8.0f MB                  %9.1f G/s %9.1f G/s\n", slots * 32 / 1048576.0,
           run<double, 1>((double *)buf, slots, 4, iters), run<double, 2>((double *)buf, slots, 4, iters));
  }
  return 0;
}
