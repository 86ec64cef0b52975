// Micro-benchmark: TMA store throughput per SM for small boxes (32 rows x 64 B vs 32 rows x 128 B vs 64 x 128 B).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ucf_vit_b200/csrc -I include -o scripts/micro/tma_store_rate \
//      scripts/micro/tma_store_rate.cu ucf_vit_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include "common.cuh"
using namespace ucf;

// grid = 148 CTAs x 256 threads (8 warps); every warp stores `iters` boxes from its own staging buffer to
// distinct places of a [rows, 3072] bf16 matrix (like the fc1 epilogue).
__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, int iters, int box_cols, int box_rows,
                                         int n_outstanding, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* my = smem + warp * 8192 * 2;
  for (int i = lane; i < 4096; i += 32) reinterpret_cast<uint32_t*>(my)[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  if (lane == 0) {
    const int cols_per_row = 3072 / box_cols;
    for (int it = 0; it < iters; ++it) {
      const int idx = (blockIdx.x * 8 + warp) * iters + it;
      const int c = (idx % cols_per_row) * box_cols, r = (idx / cols_per_row) * box_rows;
      tma_store_2d(&tm, my + (it & 1) * 8192, c, r);
      tma_store_commit();
      if (n_outstanding == 1) tma_store_wait_read<1>(); else tma_store_wait_read<3>();
    }
    tma_store_wait_all<0>();
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  const int rows = 65536, cols = 3072;
  void* d; cudaMalloc(&d, size_t(rows) * cols * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
  struct { int bc, br; CUtensorMapSwizzle sw; const char* name; } cfgs[] = {
      {32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "32 rows x 64 B (2 KB)"}, {64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "32 rows x 128 B (4 KB)"},
      {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "64 rows x 128 B (8 KB)"}, {16, 32, CU_TENSOR_MAP_SWIZZLE_32B, "32 rows x 32 B (1 KB)"}};
  for (auto& c : cfgs) {
    CUtensorMap tm;
    uint64_t dims[2] = {uint64_t(cols), uint64_t(rows)}, strides[1] = {uint64_t(cols) * 2};
    uint32_t box[2] = {uint32_t(c.bc), uint32_t(c.br)};
    if (make_tmap(&tm, d, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, c.sw)) { printf("tmap fail\n"); return 1; }
    for (int grid : {148, 16}) {          // all SMs (HBM-limited?) vs a few SMs (per-SM TMA limit)
      const int nout = 3, iters = 256;
      for (int rep = 0; rep < 2; ++rep) { k<<<grid, 256, 8 * 16384 + 1024>>>(tm, iters, c.bc, c.br, nout, cyc); cudaDeviceSynchronize(); }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = double(iters) * 8 * c.bc * 2 * c.br;
      printf("%-26s %3d CTAs: %8lld cycles for %d boxes/warp x 8 warps -> %.1f cycles/box/SM, %.1f B/clk/SM  (%s)\n", c.name, grid, mx,
             iters, double(mx) / (iters * 8), bytes / mx, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
