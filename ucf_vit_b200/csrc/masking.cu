// MAE token masking (reference simple/arch.py:663-702): the shuffle / restore
// permutations and the mask from one kernel, the kept-token gather, and the decoder-side
// "append mask tokens, un-shuffle, add position embedding" as one gather -- plus their gradients.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// One CTA per sample.  rank[j] = #{k : noise[k] < noise[j] or (noise[k] == noise[j] and k < j)} is the
// position of token j in the ascending (stable) order, i.e. argsort(argsort(noise)) = ids_restore, and
// ids_shuffle is its inverse permutation.  L^2 comparisons per sample out of shared memory: 38 k at L = 196.
__global__ void __launch_bounds__(256)
mask_plan_kernel(const float* __restrict__ noise, int L, int len_keep, long long* __restrict__ ids_shuffle,
                 long long* __restrict__ ids_restore, float* __restrict__ mask) {
  extern __shared__ float row[];
  const long long base = static_cast<long long>(blockIdx.x) * L;
  for (int j = threadIdx.x; j < L; j += 256) row[j] = __ldg(noise + base + j);
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += 256) {
    const float v = row[j];
    int rank = 0;
    for (int k = 0; k < j; ++k) rank += (row[k] <= v) ? 1 : 0;       // earlier equal values sort first
    for (int k = j + 1; k < L; ++k) rank += (row[k] < v) ? 1 : 0;
    ids_restore[base + j] = rank;
    ids_shuffle[base + rank] = j;
    mask[base + j] = rank >= len_keep ? 1.f : 0.f;
  }
}

__device__ __forceinline__ void unpack8(const uint4 q, float* v) {
  const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <bool PRM_BF16>
__device__ __forceinline__ void load8_param(const void* p, long long off, float* v) {
  if (PRM_BF16) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + off)), v);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

// out[b, i, :] = (0 <= idx[b, i] < Ls ? src[b, idx[b, i], :] : fill[:]) + pos[b * pos_bstride + i * D + :]
// One thread moves 8 consecutive elements (16 bytes of bf16).  Rows copied without fill / pos stay raw bits.
template <bool PRM_BF16>
__global__ void __launch_bounds__(256)
gather_tokens_kernel(const __nv_bfloat16* __restrict__ src, const long long* __restrict__ idx,
                     const void* __restrict__ fill, const void* __restrict__ pos, __nv_bfloat16* __restrict__ out,
                     int Ls, int Lo, int D, long long pos_bstride, long long nvec) {
  const int dv = D >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < nvec; t += stride) {
    const long long r = t / dv;
    const int d = static_cast<int>(t - r * dv) << 3;
    const long long b = r / Lo;
    const int i = static_cast<int>(r - b * Lo);
    const long long j = __ldg(idx + r);
    const bool from_src = j >= 0 && j < Ls;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (from_src) q = __ldg(reinterpret_cast<const uint4*>(src + (b * Ls + j) * D + d));
    if (pos != nullptr || (!from_src && fill != nullptr)) {
      float v[8];
      if (from_src) unpack8(q, v);
      else if (fill != nullptr) load8_param<PRM_BF16>(fill, d, v);
      else { for (int e = 0; e < 8; ++e) v[e] = 0.f; }
      if (pos != nullptr) {
        float pv[8];
        load8_param<PRM_BF16>(pos, b * pos_bstride + static_cast<long long>(i) * D + d, pv);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += pv[e];
      }
      q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    *reinterpret_cast<uint4*>(out + r * D + d) = q;
  }
}

// dsrc[b, idx[b, i], :] = dout[b, i, :] for the rows that came from src
__global__ void __launch_bounds__(256)
scatter_tokens_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ idx,
                      __nv_bfloat16* __restrict__ dsrc, int Ls, int Lo, int D, long long nvec) {
  const int dv = D >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < nvec; t += stride) {
    const long long r = t / dv;
    const int d = static_cast<int>(t - r * dv) << 3;
    const long long j = __ldg(idx + r);
    if (j < 0 || j >= Ls) continue;
    const long long b = r / Lo;
    *reinterpret_cast<uint4*>(dsrc + (b * Ls + j) * D + d) = __ldg(reinterpret_cast<const uint4*>(dout + r * D + d));
  }
}

// dfill[:] += sum of the dout rows that were filled (idx outside [0, Ls)).  Block = 32 column pairs x 8 row
// lanes over a 64-column strip and a slab of rows, like colsum_bf16_kernel.
__global__ void __launch_bounds__(256)
fill_grad_kernel(const __nv_bfloat16* __restrict__ dout, const long long* __restrict__ idx, float* __restrict__ dfill,
                 long long rows, int Ls, int D, int rows_per_block) {
  __shared__ float2 part[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + cx * 2;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, rows);
  float2 acc = make_float2(0.f, 0.f);
  if (col < D) {
    for (long long r = r0 + ry; r < r1; r += 8) {
      const long long j = __ldg(idx + r);
      if (j >= 0 && j < Ls) continue;
      const float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dout + r * D + col)));
      acc.x += a.x; acc.y += a.y;
    }
  }
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < D) {
    float2 s = part[0][cx];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += part[i][cx].x; s.y += part[i][cx].y; }
    atomicAdd(&dfill[col], s.x);
    atomicAdd(&dfill[col + 1], s.y);
  }
}

static int tok_grid(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_mask_plan(const float* noise, int B, int L, int len_keep, long long* ids_shuffle,
                             long long* ids_restore, float* mask, void* stream) {
  if (B < 0 || L < 0 || len_keep < 0 || len_keep > L) {
    set_last_error("mask_plan: need B, L >= 0 and 0 <= len_keep <= L (got %d, %d, %d)", B, L, len_keep); return UCF_ERR_BAD_ARG;
  }
  if (B == 0 || L == 0) return UCF_OK;
  if (!noise || !ids_shuffle || !ids_restore || !mask) { set_last_error("mask_plan: null pointer"); return UCF_ERR_BAD_ARG; }
  if (L > 12288) { set_last_error("mask_plan: L=%d exceeds the 12288 tokens one CTA ranks in shared memory", L); return UCF_ERR_UNSUPPORTED; }
  mask_plan_kernel<<<B, 256, sizeof(float) * L, reinterpret_cast<cudaStream_t>(stream)>>>(noise, L, len_keep, ids_shuffle,
                                                                                         ids_restore, mask);
  return check_launch("mask_plan_kernel");
}

extern "C" int ucf_gather_tokens(const void* src, const long long* idx, const void* fill, const void* pos, void* out,
                                 int B, int Ls, int Lo, int D, long long pos_bstride, int param_dtype, void* stream) {
  if (B < 0 || Ls < 0 || Lo < 0 || D <= 0) { set_last_error("gather_tokens: bad shape"); return UCF_ERR_BAD_ARG; }
  if (D % 8) { set_last_error("gather_tokens: D=%d must be a multiple of 8", D); return UCF_ERR_BAD_ARG; }
  const long long nvec = static_cast<long long>(B) * Lo * (D / 8);
  if (nvec == 0) return UCF_OK;
  if (!idx || !out || (Ls > 0 && !src)) { set_last_error("gather_tokens: null pointer"); return UCF_ERR_BAD_ARG; }
  if ((fill || pos) && param_dtype != UCF_DTYPE_F32 && param_dtype != UCF_DTYPE_BF16) {
    set_last_error("gather_tokens: fill / pos must be f32 or bf16"); return UCF_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(fill) |
       reinterpret_cast<uintptr_t>(pos)) & 15) {
    set_last_error("gather_tokens: pointers must be 16-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = tok_grid(nvec);
  if (param_dtype == UCF_DTYPE_BF16)
    gather_tokens_kernel<true><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), idx, fill, pos,
                                                     static_cast<__nv_bfloat16*>(out), Ls, Lo, D, pos_bstride, nvec);
  else
    gather_tokens_kernel<false><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), idx, fill, pos,
                                                      static_cast<__nv_bfloat16*>(out), Ls, Lo, D, pos_bstride, nvec);
  return check_launch("gather_tokens_kernel");
}

extern "C" int ucf_scatter_tokens(const void* dout, const long long* idx, void* dsrc, float* dfill, int B, int Ls,
                                  int Lo, int D, int zero_first, void* stream) {
  if (B < 0 || Ls < 0 || Lo < 0 || D <= 0) { set_last_error("scatter_tokens: bad shape"); return UCF_ERR_BAD_ARG; }
  if (D % 8) { set_last_error("scatter_tokens: D=%d must be a multiple of 8", D); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dsrc)) & 15) {
    set_last_error("scatter_tokens: pointers must be 16-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  if (dsrc && zero_first && 1LL * B * Ls > 0) {
    const cudaError_t e = cudaMemsetAsync(dsrc, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(B) * Ls * D, st);
    if (e != cudaSuccess) { set_last_error("scatter_tokens: memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  }
  const long long rows = static_cast<long long>(B) * Lo;
  const long long nvec = rows * (D / 8);
  if (nvec == 0) return UCF_OK;
  if (!dout || !idx) { set_last_error("scatter_tokens: null pointer"); return UCF_ERR_BAD_ARG; }
  if (dsrc && 1LL * B * Ls > 0) {
    scatter_tokens_kernel<<<tok_grid(nvec), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), idx,
                                                          static_cast<__nv_bfloat16*>(dsrc), Ls, Lo, D, nvec);
    const int rc = check_launch("scatter_tokens_kernel");
    if (rc != UCF_OK) return rc;
  }
  if (dfill) {
    const int strips = (D + 63) / 64;
    long long want_y = (static_cast<long long>(num_sms()) * 8 + strips - 1) / strips;
    long long rows_per_block = (rows + want_y - 1) / want_y;
    if (rows_per_block < 64) rows_per_block = 64;
    const long long gy = (rows + rows_per_block - 1) / rows_per_block;
    fill_grad_kernel<<<dim3(strips, static_cast<unsigned>(gy)), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dout), idx, dfill, rows, Ls, D, static_cast<int>(rows_per_block));
    return check_launch("fill_grad_kernel");
  }
  return UCF_OK;
}
