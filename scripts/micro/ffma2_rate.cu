// Micro-benchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a, and of a 32x64B-row TMA-free baseline.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_rate scripts/micro/ffma2_rate.cu && /tmp/ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) {          // scalar: 16 FFMA
          x[i].x = fmaf(x[i].x, a, b);
          x[i].y = fmaf(x[i].y, a, b);
        } else if (MODE == 1) {   // packed, register operands
          x[i] = __ffma2_rn(x[i], a2, b2);
        } else {                  // packed, Horner-style: acc = acc * u + const (u varies per chain)
          x[i] = __ffma2_rn(x[i], x[(i + 1) & 7], make_float2(0.125f, 0.125f));
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(t1 - t0);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 4 * 256 * sizeof(float));
  const int iters = 2000;
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    const int threads = 128 * warps_per_smsp;
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
        if (mode == 1) k<1><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
        if (mode == 2) k<2><<<148, threads>>>(d, iters, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
      }
      float cyc;
      cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
      const double fma_per_thread = double(iters) * 4 * 16;   // scalar-equivalent FMAs
      printf("warps/SMSP %d mode %d (%s): %.0f cycles, %.2f scalar-FMA/clk/SMSP-lane-group (32 lanes) => %.1f FMA/clk/SM\n",
             warps_per_smsp, mode, mode == 0 ? "FFMA" : mode == 1 ? "FFMA2 reg" : "FFMA2 dep-chain mix", cyc,
             fma_per_thread * warps_per_smsp / cyc, fma_per_thread * warps_per_smsp * 4 * 32 / cyc);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
