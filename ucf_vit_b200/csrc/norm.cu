// LayerNorm forward / backward: HBM-bandwidth-bound, one warp per row, 16-byte vector access,
// warp-shuffle reductions, row kept in registers between the statistics and the normalise pass
// (algorithmic traffic: fwd 1 read + 1 write, bwd 3 reads (dy, x, dres) + 1 write).
// Replaces nn.LayerNorm(D, eps=1e-6) of Block.norm1/norm2 and VIT.norm
// (/root/reference/src/UCF_VIT/simple/arch.py:170,266; building_blocks.py:212,226).
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

template <bool BF16>
__device__ __forceinline__ void load8(const void* base, long long idx, float (&f)[8]) {
  if (BF16) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
__device__ __forceinline__ void store8_bf16(void* base, long long idx, const float (&f)[8]) {
  uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                       pack_bf16x2(f[6], f[7]));
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = o;
}

// NCH = number of 256-element chunks a warp covers per row (D <= 256*NCH)
template <int NCH, bool X_BF16, bool P_BF16>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ gamma, const void* __restrict__ beta,
                     void* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);

  for (long long row = warp_global; row < rows; row += nwarps) {
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        load8<X_BF16>(x, row * D + col, v[c]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[c][e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[c][e] = 0.f;
      }
    }
    const float mu = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[c][e] - mu; ss += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float g[8], b[8], o[8];
        if (gamma) load8<P_BF16>(gamma, col, g); else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = 1.f;
        }
        if (beta) load8<P_BF16>(beta, col, b); else {
#pragma unroll
          for (int e = 0; e < 8; ++e) b[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[c][e] - mu) * rs * g[e] + b[e];
        store8_bf16(y, row * D + col, o);
      }
    }
  }
}

// Backward.  Each warp walks rows with a grid stride, keeps per-lane partial dgamma/dbeta in
// registers, then the block reduces them through shared memory and issues one atomicAdd per column.
// The row (dy, x) is held as the RAW 16-byte bf16 vectors between the reduction pass and the
// write pass (not as 2 x 8 floats): <= 128 registers, so two 256-thread CTAs stay resident per SM.
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <int NCH, bool P_BF16>
__global__ void __launch_bounds__(256, (NCH <= 3 ? 2 : 1))
layernorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const void* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const void* __restrict__ dres, void* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, long long rows, int D) {
  extern __shared__ float red[];   // [2][D]
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(dres);

  // gamma is re-read per row (L1-resident, D*4 bytes) instead of pinning 8*NCH registers
  float dg[NCH][8], db[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { dg[c][e] = 0.f; db[c][e] = 0.f; }
  }
  auto load_gamma = [&](int col, float (&gv)[8]) {
    if (gamma) load8<P_BF16>(gamma, col, gv);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) gv[e] = 1.f;
    }
  };
  auto load_row = [&](long long row, uint4 (&a)[NCH], uint4 (&b)[NCH]) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D && row < rows) {
        a[c] = __ldg(reinterpret_cast<const uint4*>(dyp + row * D + col));
        b[c] = __ldg(reinterpret_cast<const uint4*>(xp + row * D + col));
      } else {
        a[c] = make_uint4(0, 0, 0, 0);
        b[c] = make_uint4(0, 0, 0, 0);
      }
    }
  };

  // software pipeline: the next row's dy / x are in flight while the current row is reduced
  uint4 dyn[NCH], xn[NCH];
  load_row(warp_global, dyn, xn);
  for (long long row = warp_global; row < rows; row += nwarps) {
    uint4 dyr[NCH], xr[NCH], rr[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { dyr[c] = dyn[c]; xr[c] = xn[c]; }
    if (dres) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
        rr[c] = (col < D) ? __ldg(reinterpret_cast<const uint4*>(rp + row * D + col)) : make_uint4(0, 0, 0, 0);
      }
    }
    const float mu = mean[row], rs = rstd[row];
    load_row(row + nwarps, dyn, xn);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float dyv[8], xv[8], gv[8];
        unpack8(dyr[c], dyv);
        unpack8(xr[c], xv);
        load_gamma(col, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mu) * rs;
          const float gy = dyv[e] * gv[e];
          s1 += gy;
          s2 = fmaf(gy, xh, s2);
          dg[c][e] = fmaf(dyv[e], xh, dg[c][e]);
          db[c][e] += dyv[e];
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float dyv[8], xv[8], gv[8], o[8];
        unpack8(dyr[c], dyv);
        unpack8(xr[c], xv);
        load_gamma(col, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mu) * rs;
          o[e] = rs * (dyv[e] * gv[e] - s1 - xh * s2);
        }
        if (dres) {
          float r[8];
          unpack8(rr[c], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += r[e];
        }
        store8_bf16(dx, row * D + col, o);
      }
    }
  }

  if (dgamma == nullptr && dbeta == nullptr) return;
  float* red_g = red;
  float* red_b = red + D;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&red_g[col + e], dg[c][e]);
        atomicAdd(&red_b[col + e], db[c][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], red_g[i]);
    if (dbeta) atomicAdd(&dbeta[i], red_b[i]);
  }
}

static int ln_grid(long long rows) {
  const long long warps_per_block = 8;
  long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;   // 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace ucf

using namespace ucf;

#define UCF_LN_DISPATCH_NCH(D, MACRO)                       \
  if (D <= 256) { MACRO(1) } else if (D <= 512) { MACRO(2) } \
  else if (D <= 768) { MACRO(3) } else if (D <= 1024) { MACRO(4) } \
  else if (D <= 1536) { MACRO(6) } else if (D <= 2048) { MACRO(8) } \
  else if (D <= 4096) { MACRO(16) }

extern "C" int ucf_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                                 float* rstd, long long rows, int D, float eps, int x_dtype, int param_dtype,
                                 void* stream) {
  if (rows <= 0) return UCF_OK;
  if (D <= 0 || D % 8 != 0 || D > 4096) {
    set_last_error("layernorm_fwd: D=%d must be a multiple of 8 and <= 4096", D);
    return UCF_ERR_BAD_ARG;
  }
  if (!x || !y) { set_last_error("layernorm_fwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_grid(rows);
  const bool xb = x_dtype == UCF_DTYPE_BF16, pb = param_dtype == UCF_DTYPE_BF16;
#define LAUNCH(NCH)                                                                                         \
  if (xb && pb) layernorm_fwd_kernel<NCH, true, true><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);  \
  else if (xb) layernorm_fwd_kernel<NCH, true, false><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps); \
  else if (pb) layernorm_fwd_kernel<NCH, false, true><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps); \
  else layernorm_fwd_kernel<NCH, false, false><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
  UCF_LN_DISPATCH_NCH(D, LAUNCH)
#undef LAUNCH
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int ucf_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 long long rows, int D, int param_dtype, void* stream) {
  if (rows <= 0) return UCF_OK;
  if (D <= 0 || D % 8 != 0 || D > 2048) {
    set_last_error("layernorm_bwd: D=%d must be a multiple of 8 and <= 2048", D);
    return UCF_ERR_BAD_ARG;
  }
  if (!dy || !x || !mean || !rstd || !dx) { set_last_error("layernorm_bwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // fewer, fatter blocks than forward: every block ends with 2*D global atomics
  long long blocks = (rows + 7) / 8;
  const long long cap = static_cast<long long>(ucf::num_sms()) * (D <= 768 ? 2 : 1);   // resident CTAs per SM, persistent
  if (blocks > cap) blocks = cap;
  const int grid = static_cast<int>(blocks < 1 ? 1 : blocks);
  const size_t smem = 2 * static_cast<size_t>(D) * sizeof(float);
  const bool pb = param_dtype == UCF_DTYPE_BF16;
#define LAUNCH(NCH)                                                                                      \
  if (NCH <= 8) {                                                                                        \
    if (pb) layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), true><<<grid, 256, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D); \
    else layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), false><<<grid, 256, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D);   \
  }
  UCF_LN_DISPATCH_NCH(D, LAUNCH)
#undef LAUNCH
  return check_launch("layernorm_bwd_kernel");
}
