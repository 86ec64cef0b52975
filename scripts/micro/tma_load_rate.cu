// Micro-benchmark: TMA load throughput per SM (L2 -> shared memory), few SMs vs all SMs, for the GEMM's
// operand boxes (128 rows x 64 bf16 = 16 KB, SWIZZLE_128B).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ucf_vit_b200/csrc -I include -o scripts/micro/tma_load_rate \
//      scripts/micro/tma_load_rate.cu ucf_vit_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace ucf;

constexpr int STAGES = 6, BOX_BYTES = 16384;

template <int NPROD>   // 1: one thread issues both boxes of a stage; 2: two warps issue one box each
__global__ void __launch_bounds__(96) k(const __grid_constant__ CUtensorMap tm, int iters, int rows, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 2 * BOX_BYTES);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  const int row_blocks = rows / 128;
  if (threadIdx.x == 0 || (NPROD == 2 && threadIdx.x == 64)) {   // producer(s): two 16 KB boxes per stage (like A + B of one k-block)
    const int me = threadIdx.x == 0 ? 0 : 1;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      const int kb = it % 12;
      const int rb = (blockIdx.x * 7 + it / 12) % row_blocks;
      if (NPROD == 1) {
        mbar_expect_tx(&full[s], 2 * BOX_BYTES);
        tma_load_2d(smem + (s * 2 + 0) * BOX_BYTES, &tm, &full[s], kb * 64, rb * 128);
        tma_load_2d(smem + (s * 2 + 1) * BOX_BYTES, &tm, &full[s], kb * 64, ((rb + 3) % row_blocks) * 128);
      } else {
        mbar_expect_tx(&full[s], BOX_BYTES);
        tma_load_2d(smem + (s * 2 + me) * BOX_BYTES, &tm, &full[s], kb * 64, ((rb + 3 * me) % row_blocks) * 128);
      }
    }
  } else if (threadIdx.x == 32) {     // consumer: releases the stage as soon as it has landed
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  const int rows = 50432, cols = 768;
  void* d; cudaMalloc(&d, size_t(rows) * cols * 2);
  cudaMemset(d, 0, size_t(rows) * cols * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int smem = STAGES * 2 * BOX_BYTES + 1024 + 256;
  CUtensorMap tm;
  uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)}, strides[1] = {uint64_t(cols) * 2};
  uint32_t box[2] = {64, 128};
  if (make_tmap(&tm, d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap fail\n"); return 1; }
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int nprod : {1, 2})
  for (int grid : {148, 16}) {
    const int iters = 2048;
    for (int rep = 0; rep < 3; ++rep) {
      if (nprod == 1) k<1><<<grid, 96, smem>>>(tm, iters, rows, cyc); else k<2><<<grid, 96, smem>>>(tm, iters, rows, cyc);
      cudaDeviceSynchronize();
    }
    printf("%d producer thread(s): ", nprod);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = double(iters) * 2 * BOX_BYTES;
    printf("%3d CTAs: %8lld cycles for %d stages of 32 KB -> %.1f B/clk/SM, chip %.2f TB/s at 1.9 GHz  (%s)\n", grid, mx, iters,
           bytes / mx, bytes / mx * grid * 1.9e9 / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
