// Micro-benchmark: tcgen05.ld (tensor memory -> registers) throughput per SM for the thread-per-row access the
// attention softmax uses (32x32b shapes), as a function of the vector length, the number of loads in flight per
// warp and the number of warps per scheduler; plus MUFU.EX2 throughput for reference.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ucf_vit_b200/csrc -I include -o scripts/micro/tmem_ld_rate \
//      scripts/micro/tmem_ld_rate.cu ucf_vit_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace ucf;

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, "
      "%27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, "
      "%52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
        "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
        "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
        "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}

// MODE 0: x16 loads, 1: x32, 2: x64 ; INFL = loads issued before each wait; columns walked modulo 128
template <int MODE, int INFL>
__global__ void __launch_bounds__(256, 1) ld_kernel(int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (MODE == 0) {
        uint32_t v[INFL][16];
#pragma unroll
        for (int k = 0; k < INFL; ++k) tmem_ld16(tm + ((i * INFL + k) * 16) % 256, v[k]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < INFL; ++k) acc += __uint_as_float(v[k][0] ^ v[k][15]);
      } else if (MODE == 1) {
        uint32_t v[INFL][32];
#pragma unroll
        for (int k = 0; k < INFL; ++k) tmem_ld32(tm + ((i * INFL + k) * 32) % 256, v[k]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < INFL; ++k) acc += __uint_as_float(v[k][0] ^ v[k][31]);
      } else {
        uint32_t v[INFL][64];
#pragma unroll
        for (int k = 0; k < INFL; ++k) ld64(tm + ((i * INFL + k) * 64) % 256, v[k]);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < INFL; ++k) acc += __uint_as_float(v[k][0] ^ v[k][63]);
      }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tslot, 512);
}

__global__ void __launch_bounds__(256, 1) mufu_kernel(int iters, int nwarps, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = -0.001f * (threadIdx.x + k);
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fast_ex2(a[k]) - 1.0f;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 123.456f) sink[0] = s;
}

template <int MODE, int INFL>
static void run(int nwarps, int grid, long long* d, float* sink) {
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { ld_kernel<MODE, INFL><<<grid, 256>>>(iters, nwarps, d, sink); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const int cols = MODE == 0 ? 16 : MODE == 1 ? 32 : 64;
  const double bytes = double(iters) * INFL * cols * 32 * 4 * nwarps;
  printf("tcgen05.ld 32x32b.x%-2d  %d in flight, %d warps, grid %3d: %7.1f cycles per load, %6.1f B/clk/SM  %s\n", cols, INFL, nwarps, grid,
         double(c) / (double(iters) * INFL), bytes / double(c), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 16);
  for (int grid : {1, 148}) {
    for (int nw : {1, 4, 8}) {
      run<0, 1>(nw, grid, d, sink); run<0, 4>(nw, grid, d, sink);
      run<1, 1>(nw, grid, d, sink); run<1, 2>(nw, grid, d, sink); run<1, 4>(nw, grid, d, sink);
      run<2, 1>(nw, grid, d, sink); run<2, 2>(nw, grid, d, sink);
    }
  }
  for (int nw : {1, 4, 8}) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) { mufu_kernel<<<148, 256>>>(iters, nw, d, sink); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.EX2 (+1 FADD) %d warps: %.2f ex2 per clk per SM\n", nw, double(iters) * 8 * 32 * nw / double(c));
  }
  return 0;
}
