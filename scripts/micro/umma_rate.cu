// Micro-benchmark: issue rate of SMALL tcgen05.mma shapes (M = 128, K = 16, N = 64 / 128 / 256) as the attention
// kernels use them -- dependent chains into one accumulator vs. interleaved accumulators, A operand from shared
// memory (SS) vs. from tensor memory (TS) -- plus a functional check of the TS operand layout (bf16 A in TMEM:
// lane = row m, 32-bit column c holds elements k = 2c (low half) and k = 2c + 1 (high half)).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ucf_vit_b200/csrc -I include -o scripts/micro/umma_rate \
//      scripts/micro/umma_rate.cu ucf_vit_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace ucf;

// MODE 0: SS, A K-major (P tile, 128B swizzle), B MN-major (V tile)     -> the PV product
// MODE 1: SS, A K-major, B K-major                                       -> the S = Q K^T product
// MODE 2: TS, A from tensor memory, B MN-major                           -> PV with P in TMEM
// MODE 3: SS, A MN-major, B MN-major                                     -> dV = P^T dO / dK = dS^T Q
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 512);
  // operands: any finite bit pattern will do
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (warp == 1) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MODE == 3, MODE != 1);
    // accumulators: NACC x N columns from column 0 (N * NACC <= 384); TS A operand at column 448
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; ++i) {
        const int kk = i & 7;
        const uint32_t d = tm + (i % NACC) * N;
        const uint32_t acc = i >= NACC ? 1u : 0u;
        if (MODE == 0)
          umma_bf16(d, umma_smem_desc(a_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                    umma_smem_desc(b_addr + kk * 2048, 16384, 1024), idesc, acc);
        else if (MODE == 1)
          umma_bf16(d, umma_smem_desc(a_addr + (kk & 3) * 32, 16, 1024), umma_smem_desc(b_addr + (kk & 3) * 32, 16, 1024), idesc, acc);
        else if (MODE == 2)
          umma_bf16_ts(d, tm + 448 + kk * 8, umma_smem_desc(b_addr + kk * 2048, 16384, 1024), idesc, acc);
        else
          umma_bf16(d, umma_smem_desc(a_addr + kk * 2048, 16384, 1024), umma_smem_desc(b_addr + kk * 2048, 16384, 1024), idesc, acc);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// Functional check of the TS form: D[m][n] = sum_k A[m][k] B[n][k] with A in tensor memory.
// A[m][k] = (m % 13) + k / 32.0 (exact in bf16 for k < 16), B[n][k] = (n == k): D[m][n] = A[m][n] for n < 16.
__global__ void __launch_bounds__(128, 1) ts_layout_kernel(float* out /* [128][16] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 512);
  // B: K-major, 16 rows (n) of 128 B (64 k), SWIZZLE_128B
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int n = i / 64, k = i % 64;
    const uint32_t off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16(n == k ? 1.0f : 0.0f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  // A into tensor memory columns 256..263: thread (lane of TMEM) m = threadIdx.x
  {
    const int m = threadIdx.x;
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float lo = (m % 13) + (2 * c) / 32.0f, hi = (m % 13) + (2 * c + 1) / 32.0f;
      v[c] = c < 8 ? pack_bf16x2(lo, hi) : 0u;
    }
    tmem_st32(tm + 256 + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    if (elect_one()) {
      umma_bf16_ts(tm, tm + 256, umma_smem_desc(smem_u32(smem), 16, 1024), umma_idesc_bf16(128, 16, false, false), 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t d[16];
  tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16), d);
  tmem_wait_ld();
  for (int n = 0; n < 16; ++n) out[threadIdx.x * 16 + n] = __uint_as_float(d[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
  (void)lane;
}

template <int MODE, int N, int NACC>
static void run(const char* what, int grid, long long* d) {
  const int n_mma = 4096;
  cudaFuncSetAttribute(rate_kernel<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<MODE, N, NACC><<<grid, 128, 100 * 1024>>>(n_mma, d);
    cudaDeviceSynchronize();
  }
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double per = double(c) / n_mma;
  printf("%-44s N=%3d acc=%d grid=%3d: %6.1f cycles / MMA  (ideal %5.1f, %4.0f%% of the tensor peak)  %s\n", what, N, NACC, grid, per,
         N / 2.0, 100.0 * (N / 2.0) / per, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* o; cudaMalloc(&o, 128 * 16 * 4);
  cudaFuncSetAttribute(ts_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024);
  ts_layout_kernel<<<1, 128, 8 * 1024>>>(o);
  cudaDeviceSynchronize();
  std::vector<float> h(128 * 16);
  cudaMemcpy(h.data(), o, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 16; ++n)
      if (h[m * 16 + n] != (m % 13) + n / 32.0f) ++bad;
  printf("TS operand layout (lane = m, column c = elements 2c | 2c+1): %s (%d mismatches)  %s\n", bad ? "WRONG" : "ok", bad,
         cudaGetErrorString(cudaGetLastError()));
  if (bad) {
    for (int m = 0; m < 3; ++m) { printf("  row %d:", m); for (int n = 0; n < 16; ++n) printf(" %.3f", h[m * 16 + n]); printf("\n"); }
    for (int m = 32; m < 34; ++m) { printf("  row %d:", m); for (int n = 0; n < 16; ++n) printf(" %.3f", h[m * 16 + n]); printf("\n"); }
  }
  for (int grid : {1, 148}) {
    run<0, 64, 1>("SS  A K-major smem, B MN-major (PV)", grid, d);
    run<0, 64, 2>("SS  A K-major smem, B MN-major (PV)", grid, d);
    run<0, 64, 4>("SS  A K-major smem, B MN-major (PV)", grid, d);
    run<0, 128, 1>("SS  A K-major smem, B MN-major", grid, d);
    run<1, 128, 1>("SS  A K-major, B K-major (S = QK^T)", grid, d);
    run<1, 128, 2>("SS  A K-major, B K-major (S = QK^T)", grid, d);
    run<1, 256, 1>("SS  A K-major, B K-major", grid, d);
    run<3, 64, 1>("SS  A MN-major, B MN-major (dV, dK)", grid, d);
    run<3, 64, 2>("SS  A MN-major, B MN-major (dV, dK)", grid, d);
    run<2, 64, 1>("TS  A tensor memory, B MN-major (PV)", grid, d);
    run<2, 64, 2>("TS  A tensor memory, B MN-major (PV)", grid, d);
    run<2, 128, 1>("TS  A tensor memory, B MN-major", grid, d);
  }
  return 0;
}
