// Micro-benchmark: latency of one mbarrier hand-off between two warps of a CTA (arrive -> waiter resumes),
// for the waiting styles used in the kernels.  Also: tcgen05.commit -> waiter resumes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ucf_vit_b200/csrc -I include -o scripts/micro/mbar_hop \
//      scripts/micro/mbar_hop.cu ucf_vit_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace ucf;

template <int MODE>   // 0: try_wait loop (mbar_wait), 1: test_wait spin, 2: try_wait with 32 lanes polling
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
  if (MODE == 1) { while (!mbar_test_wait(bar, parity)) { } }
  else mbar_wait(bar, parity);
}

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, long long* out, int busy_warps) {
  __shared__ uint64_t bars[2];
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == 0) {                       // ping
    if (MODE == 2 || lane == 0) {
      for (int i = 0; i < iters; ++i) {
        if (lane == 0) mbar_arrive(&bars[0]);
        wait_mode<MODE>(&bars[1], i & 1);
      }
    }
  } else if (warp == 1) {                // pong
    if (MODE == 2 || lane == 0) {
      for (int i = 0; i < iters; ++i) {
        wait_mode<MODE>(&bars[0], i & 1);
        if (lane == 0) mbar_arrive(&bars[1]);
      }
    }
  } else if (warp - 2 < busy_warps) {    // optional ALU-busy warps on the other schedulers
    float x = threadIdx.x;
    for (int i = 0; i < iters * 40; ++i) x = fmaf(x, 1.0001f, 0.5f);
    if (x == 123.f) out[1] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 20000;
  for (int busy = 0; busy <= 2; busy += 2) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, 128>>>(iters, d, busy);
        if (mode == 1) k<1><<<1, 128>>>(iters, d, busy);
        if (mode == 2) k<2><<<1, 128>>>(iters, d, busy);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("busy warps %d, %-34s: %.0f cycles per hop (%s)\n", busy,
             mode == 0 ? "try_wait loop, one lane" : mode == 1 ? "test_wait spin, one lane" : "try_wait loop, 32 lanes", double(c) / (2.0 * iters),
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
