// 1x1x1 convolutions of the UNETR decoder on channels-last bf16 tensors -- the residual projections `conv3` of MONAI's
// UnetResBlock (in_chans -> 16 at 128^3, 32 -> 16, ...) and the output head UnetOutBlock (16 -> num_classes), built at
// /root/reference/src/UCF_VIT/simple/arch.py:808-940,960-993 -- forward, data gradient and weight / bias gradient.
//
// In channels-last memory a 1x1x1 convolution is y[v, :] = W x[v, :] + b per voxel with a 4..32-wide W: HBM-bound
// (2 (Ci + Co) bytes per voxel against 2 Ci Co FLOP), far too narrow for a tensor-core tile; the library runs them as
// sm_80 implicit GEMMs at 1.8-2.2 ms per layer and direction at 16 x 128^3 voxels (profiles/r02_unetr_kernel_profile_*.log),
// 4-8x the time of one read + one write.  Here: one thread per voxel, the voxel's channels in registers, W broadcast from
// shared memory as float4 (forward / data gradient, which is the same kernel on W^T); the weight gradient is a
// register-tiled outer-product accumulation over voxel tiles staged in shared memory with a fixed-order two-stage reduction.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

template <int C>
__device__ __forceinline__ void pw_load(const __nv_bfloat16* p, float (&f)[C]) {      // C in {4, 8, 16, 32}
  if (C == 4) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  } else {
#pragma unroll
    for (int i = 0; i < C / 8; ++i) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(p) + i);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(w[j]);
        f[i * 8 + 2 * j] = a.x; f[i * 8 + 2 * j + 1] = a.y;
      }
    }
  }
}
template <int C>
__device__ __forceinline__ void pw_store(__nv_bfloat16* p, const float (&f)[C]) {
  if (C == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
  } else {
#pragma unroll
    for (int i = 0; i < C / 8; ++i)
      reinterpret_cast<uint4*>(p)[i] = make_uint4(pack_bf16x2(f[i * 8], f[i * 8 + 1]), pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]),
                                                  pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]), pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]));
  }
}

// y[v, co] = sum_ci x[v, ci] w[co, ci] (+ bias[co])
template <int CI, int CO>
__global__ void __launch_bounds__(256)
pointwise_conv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      __nv_bfloat16* __restrict__ y, long long V) {
  __shared__ __align__(16) float ws[CO * CI];
  __shared__ float bs[CO];
  for (int i = threadIdx.x; i < CO * CI; i += 256) ws[i] = w[i];
  for (int i = threadIdx.x; i < CO; i += 256) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  for (long long v = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; v < V; v += stride) {
    float xr[CI], o[CO];
    pw_load<CI>(x + v * CI, xr);
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      float acc = bs[co];
#pragma unroll
      for (int c4 = 0; c4 < CI / 4; ++c4) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + co * CI + c4 * 4);      // same address for every lane: broadcast
        acc = fmaf(xr[c4 * 4], w4.x, acc);
        acc = fmaf(xr[c4 * 4 + 1], w4.y, acc);
        acc = fmaf(xr[c4 * 4 + 2], w4.z, acc);
        acc = fmaf(xr[c4 * 4 + 3], w4.w, acc);
      }
      o[co] = acc;
    }
    pw_store<CO>(y + v * CO, o);
  }
}

// partial[cta][co * CI + ci] = sum over the CTA's voxels of dy[v, co] x[v, ci];  partial[cta][CO * CI + co] = sum dy[v, co]
constexpr int PW_TV = 128;   // voxels per staged tile
template <int CI, int CO>
__global__ void __launch_bounds__(256)
pointwise_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                       long long V) {
  constexpr int TI = CI / 4, TO = CO / 4, OUT_T = TI * TO;       // 4 x 4 output micro-tiles
  constexpr int GROUPS = 256 / OUT_T;                            // voxel groups working side by side
  static_assert(OUT_T <= 256 && 256 % OUT_T == 0, "micro-tiles must divide the block");
  constexpr int STAGE = PW_TV * (CI + CO), FOLD = 256 * 16 + 256 * 4;    // the fold buffer reuses the staging memory
  __shared__ __align__(16) float buf[STAGE > FOLD ? STAGE : FOLD];
  float* xs = buf;
  float* ys = buf + PW_TV * CI;
  float* red = buf;
  const int t = threadIdx.x, ot = t % OUT_T, g = t / OUT_T;
  const int cot = ot / TI, cit = ot % TI;
  float acc[4][4], db[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    db[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const long long tiles = (V + PW_TV - 1) / PW_TV;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long v0 = tile * PW_TV;
    __syncthreads();                                             // previous tile fully consumed
    for (int i = t; i < PW_TV * CI / 4; i += 256) {
      const long long e = v0 * CI + static_cast<long long>(i) * 4;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < V * CI) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(x + e));
        const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
        f = make_float4(a.x, a.y, b.x, b.y);
      }
      reinterpret_cast<float4*>(xs)[i] = f;
    }
    for (int i = t; i < PW_TV * CO / 4; i += 256) {
      const long long e = v0 * CO + static_cast<long long>(i) * 4;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < V * CO) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(dy + e));
        const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
        f = make_float4(a.x, a.y, b.x, b.y);
      }
      reinterpret_cast<float4*>(ys)[i] = f;
    }
    __syncthreads();
#pragma unroll 4
    for (int v = g; v < PW_TV; v += GROUPS) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + v * CI + cit * 4);
      const float4 yv = *reinterpret_cast<const float4*>(ys + v * CO + cot * 4);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ya[i], xa[j], acc[i][j]);
        if (cit == 0) db[i] += ya[i];
      }
    }
  }
  // fold the voxel groups (fixed order), then one row of CO * CI + CO floats per CTA
  __syncthreads();
  float* mine = red + (g * OUT_T + ot) * 16;                     // GROUPS * OUT_T * 16 = 4096 floats
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[i * 4 + j] = acc[i][j];
  float* rdb = red + 4096 + t * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) rdb[i] = db[i];
  __syncthreads();
  float* out = partial + static_cast<long long>(blockIdx.x) * (CO * CI + CO);
  for (int k = t; k < CO * CI; k += 256) {
    const int co = k / CI, ci = k - co * CI;
    const int o2 = (co / 4) * TI + ci / 4, e = (co % 4) * 4 + ci % 4;
    float s = 0.f;
    for (int gg = 0; gg < GROUPS; ++gg) s += red[(gg * OUT_T + o2) * 16 + e];
    out[k] = s;
  }
  for (int co = t; co < CO; co += 256) {
    const int o2 = (co / 4) * TI;                                // the micro-tile column with cit == 0
    float s = 0.f;
    for (int gg = 0; gg < GROUPS; ++gg) s += red[4096 + (gg * OUT_T + o2) * 4 + co % 4];
    out[CO * CI + co] = s;
  }
}

// out[i] = sum over rows of partial[row][i] (fixed order)
__global__ void __launch_bounds__(256)
pointwise_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ dbias, int rows,
                              int n_w, int n_b) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_w + n_b) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += partial[static_cast<long long>(r) * (n_w + n_b) + i];
  if (i < n_w) dw[i] = s;
  else if (dbias) dbias[i - n_w] = s;
}

static bool pw_ok(int c) { return c == 4 || c == 8 || c == 16 || c == 32; }

template <int CI>
static int pw_fwd_co(int Co, const __nv_bfloat16* x, const float* w, const float* b, __nv_bfloat16* y, long long V, int grid,
                     cudaStream_t st) {
  switch (Co) {
    case 4: pointwise_conv_kernel<CI, 4><<<grid, 256, 0, st>>>(x, w, b, y, V); break;
    case 8: pointwise_conv_kernel<CI, 8><<<grid, 256, 0, st>>>(x, w, b, y, V); break;
    case 16: pointwise_conv_kernel<CI, 16><<<grid, 256, 0, st>>>(x, w, b, y, V); break;
    default: pointwise_conv_kernel<CI, 32><<<grid, 256, 0, st>>>(x, w, b, y, V); break;
  }
  return check_launch("pointwise_conv_kernel");
}
template <int CI>
static int pw_wgrad_co(int Co, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* ws, long long V, int grid, cudaStream_t st) {
  switch (Co) {
    case 4: pointwise_wgrad_kernel<CI, 4><<<grid, 256, 0, st>>>(x, dy, ws, V); break;
    case 8: pointwise_wgrad_kernel<CI, 8><<<grid, 256, 0, st>>>(x, dy, ws, V); break;
    case 16: pointwise_wgrad_kernel<CI, 16><<<grid, 256, 0, st>>>(x, dy, ws, V); break;
    default: pointwise_wgrad_kernel<CI, 32><<<grid, 256, 0, st>>>(x, dy, ws, V); break;
  }
  return check_launch("pointwise_wgrad_kernel");
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_pointwise_conv_supported(int Ci, int Co) { return pw_ok(Ci) && pw_ok(Co); }

extern "C" int ucf_pointwise_conv_ctas(long long V) {
  if (V <= 0) return 0;
  const long long tiles = (V + PW_TV - 1) / PW_TV, want = static_cast<long long>(num_sms()) * 4;
  return static_cast<int>(tiles < want ? tiles : want);
}

extern "C" int ucf_pointwise_conv(const void* x, const float* w, const float* bias, void* y, long long V, int Ci, int Co,
                                  void* stream) {
  if (V <= 0) return UCF_OK;
  if (!x || !w || !y) { set_last_error("pointwise_conv: null pointer"); return UCF_ERR_BAD_ARG; }
  if (!ucf_pointwise_conv_supported(Ci, Co)) {
    set_last_error("pointwise_conv: Ci=%d, Co=%d not served (each of 4, 8, 16, 32)", Ci, Co);
    return UCF_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) {
    set_last_error("pointwise_conv: tensors must be 16-byte aligned");
    return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (V + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  switch (Ci) {
    case 4: return pw_fwd_co<4>(Co, xp, w, bias, yp, V, grid, st);
    case 8: return pw_fwd_co<8>(Co, xp, w, bias, yp, V, grid, st);
    case 16: return pw_fwd_co<16>(Co, xp, w, bias, yp, V, grid, st);
    default: return pw_fwd_co<32>(Co, xp, w, bias, yp, V, grid, st);
  }
}

extern "C" int ucf_pointwise_conv_wgrad(const void* x, const void* dy, float* dw, float* dbias, long long V, int Ci, int Co,
                                        float* workspace, void* stream) {
  if (!x || !dy || !dw || !workspace) { set_last_error("pointwise_conv_wgrad: null pointer"); return UCF_ERR_BAD_ARG; }
  if (V <= 0 || !ucf_pointwise_conv_supported(Ci, Co)) {
    set_last_error("pointwise_conv_wgrad: V=%lld, Ci=%d, Co=%d not served (channels each of 4, 8, 16, 32)", V, Ci, Co);
    return UCF_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 7) {
    set_last_error("pointwise_conv_wgrad: tensors must be 8-byte aligned");
    return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ucf_pointwise_conv_ctas(V);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* dyp = static_cast<const __nv_bfloat16*>(dy);
  int rc;
  switch (Ci) {
    case 4: rc = pw_wgrad_co<4>(Co, xp, dyp, workspace, V, grid, st); break;
    case 8: rc = pw_wgrad_co<8>(Co, xp, dyp, workspace, V, grid, st); break;
    case 16: rc = pw_wgrad_co<16>(Co, xp, dyp, workspace, V, grid, st); break;
    default: rc = pw_wgrad_co<32>(Co, xp, dyp, workspace, V, grid, st); break;
  }
  if (rc) return rc;
  const int n_w = Co * Ci;
  pointwise_wgrad_reduce_kernel<<<(n_w + Co + 255) / 256, 256, 0, st>>>(workspace, dw, dbias, grid, n_w, Co);
  return check_launch("pointwise_wgrad_reduce_kernel");
}
