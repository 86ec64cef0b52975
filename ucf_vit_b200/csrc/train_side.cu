// The two bandwidth-bound steps either side of the transformer path in a training step
// (SURVEY.md §8f ranks 2 and 3):
//   * the reconstruction loss against the on-the-fly patchified target (no patchified copy of the
//     image, no (pred - y)^2 temporaries): one read of pred + image forward, one more plus the dpred
//     write backward;
//   * the AdamW update over many parameter tensors per launch (7 fp32 streams per element).
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// ------------------------------------------------------------------------------------------------
// multi-tensor AdamW
// ------------------------------------------------------------------------------------------------
constexpr int kAdamTensors = 24;   // 24 * 40 B of kernel arguments

struct AdamWArgs {
  float* p[kAdamTensors];
  const float* g[kAdamTensors];
  float* m[kAdamTensors];
  float* v[kAdamTensors];
  long long n[kAdamTensors];
};

struct AdamWScalars {
  float decay;        // 1 - lr * weight_decay
  float beta1, beta2;
  float one_m_beta1, one_m_beta2;
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2_sqrt; // 1 / sqrt(1 - beta2^t)
  float eps;
  float gsign;        // -1 when maximizing
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWScalars& s) {
  g *= s.gsign;
  p *= s.decay;
  m = fmaf(s.one_m_beta1, g - m, m);                 // lerp(m, g, 1 - beta1)
  v = fmaf(s.beta2, v, s.one_m_beta2 * g * g);
  const float denom = fmaf(sqrtf(v), s.inv_bc2_sqrt, s.eps);
  p -= s.step_size * (m / denom);
}

template <bool VEC>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamWArgs a, const AdamWScalars s) {
  float* __restrict__ p = a.p[blockIdx.y];
  const float* __restrict__ g = a.g[blockIdx.y];
  float* __restrict__ m = a.m[blockIdx.y];
  float* __restrict__ v = a.v[blockIdx.y];
  const long long n = a.n[blockIdx.y];
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  if (VEC) {
    const long long nv = n >> 2;
    for (long long i = tid; i < nv; i += nthreads) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adamw_one(pp.x, gg.x, mm.x, vv.x, s);
      adamw_one(pp.y, gg.y, mm.y, vv.y, s);
      adamw_one(pp.z, gg.z, mm.z, vv.z, s);
      adamw_one(pp.w, gg.w, mm.w, vv.w, s);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (nv << 2) + tid; i < n; i += nthreads) {   // < 4 trailing elements
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, s);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  } else {
    for (long long i = tid; i < n; i += nthreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i], mm, vv, s);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reconstruction loss against the patchified image
//
// pred [B, L, P*C] (P = p0*p1*p2 pixels of a patch, channel fastest) is compared with
// img [B, C, G0*p0, G1*p1, G2*p2], token l = (g0*G1 + g1)*G2 + g2, pixel q = (q0*p1 + q1)*p2 + q2.
// One CTA walks whole tokens; a thread owns one pixel (all C channels) at a time, so image reads
// are contiguous runs of p2 (or p1 when p2 == 1) elements and pred reads cover the token row densely.
// ------------------------------------------------------------------------------------------------
struct PatchGeom {
  int B, C, L;
  int G0, G1, G2;
  int p0, p1, p2;
  int sx, sy;            // image strides (elements) of axis 0 / axis 1 inside one channel plane; axis 2 is 1
  long long plane;       // X*Y*Z
};

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// Pixel q = (q0*p1 + q1)*p2 + q2 of a patch as mixed-radix digits.  A thread visits q = tid, tid + 256, ...:
// the digits advance by the digits of 256 with carries, so the loops below contain no division.
struct PixelWalk {
  int q, q0, q1, q2;
  __device__ __forceinline__ PixelWalk(const PatchGeom& gm) {
    q = threadIdx.x;
    q2 = q % gm.p2;
    const int r = q / gm.p2;
    q1 = r % gm.p1; q0 = r / gm.p1;
  }
  __device__ __forceinline__ int offset(const PatchGeom& gm) const { return q0 * gm.sx + q1 * gm.sy + q2; }
  __device__ __forceinline__ void next(const PatchGeom& gm, int d0, int d1, int d2) {
    q += 256;
    q2 += d2; if (q2 >= gm.p2) { q2 -= gm.p2; ++q1; }
    q1 += d1; if (q1 >= gm.p1) { q1 -= gm.p1; ++q0; }
    q0 += d0;
  }
};

// A CTA owns a contiguous run of tokens; the (b, g0, g1, g2) odometer advances without divisions.
struct TokenWalk {
  int bl, end, b, g0, g1, g2;
  __device__ __forceinline__ TokenWalk(const PatchGeom& gm) {
    const int BL = gm.B * gm.L;
    const int per = (BL + gridDim.x - 1) / gridDim.x;
    bl = min(BL, static_cast<int>(blockIdx.x) * per);
    end = min(BL, bl + per);
    b = bl / gm.L;
    const int l = bl - b * gm.L;
    g2 = l % gm.G2; g1 = (l / gm.G2) % gm.G1; g0 = l / (gm.G2 * gm.G1);
  }
  __device__ __forceinline__ long long image_base(const PatchGeom& gm) const {
    return static_cast<long long>(b) * gm.C * gm.plane + static_cast<long long>(g0) * gm.p0 * gm.sx +
           static_cast<long long>(g1) * gm.p1 * gm.sy + static_cast<long long>(g2) * gm.p2;
  }
  __device__ __forceinline__ void next(const PatchGeom& gm) {
    ++bl;
    if (++g2 == gm.G2) { g2 = 0; if (++g1 == gm.G1) { g1 = 0; if (++g0 == gm.G0) { g0 = 0; ++b; } } }
  }
};

template <typename TP, typename TI>
__global__ void __launch_bounds__(256, 4)
patch_mse_fwd_kernel(const TP* __restrict__ pred, const TI* __restrict__ img, const float* __restrict__ mask,
                     const PatchGeom gm, double* __restrict__ partials) {
  const int P = gm.p0 * gm.p1 * gm.p2;
  const int d2 = 256 % gm.p2, d1 = (256 / gm.p2) % gm.p1, d0 = 256 / (gm.p2 * gm.p1);
  const PixelWalk first(gm);
  float acc = 0.f;
  for (TokenWalk t(gm); t.bl < t.end; t.next(gm)) {
    const float w = mask ? __ldg(mask + t.bl) : 1.f;
    if (w == 0.f) continue;
    const TI* ib = img + t.image_base(gm);
    const TP* pb = pred + static_cast<long long>(t.bl) * P * gm.C;
    float tok = 0.f;
    for (PixelWalk x = first; x.q < P; x.next(gm, d0, d1, d2)) {
      const TI* ip = ib + x.offset(gm);
      const TP* pp = pb + static_cast<long long>(x.q) * gm.C;
      for (int c = 0; c < gm.C; ++c) {
        const float d = ldf(pp + c) - ldf(ip + c * gm.plane);
        tok = fmaf(d, d, tok);
      }
    }
    acc = fmaf(w, tok, acc);
  }
  // block reduction in double: the order is fixed, so the loss is bit-reproducible run to run
  __shared__ double red[8];
  double d = static_cast<double>(acc);
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}

// out[0] = loss, out[1] = 1 / denominator (kept for the backward pass)
__global__ void __launch_bounds__(256)
patch_mse_finish_kernel(const double* __restrict__ partials, int n_partials, const float* __restrict__ mask,
                        int BL, double elems_per_token, float* __restrict__ out) {
  __shared__ double red[8];
  __shared__ double red_m[8];
  double s = 0.0, ms = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += blockDim.x) s += partials[i];
  if (mask) for (int i = threadIdx.x; i < BL; i += blockDim.x) ms += static_cast<double>(mask[i]);
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ms += __shfl_xor_sync(0xffffffffu, ms, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red_m[threadIdx.x >> 5] = ms; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, tm = 0.0;
    for (int i = 0; i < 8; ++i) { t += red[i]; tm += red_m[i]; }
    const double denom = (mask ? tm : static_cast<double>(BL)) * elems_per_token;
    out[0] = static_cast<float>(t / denom);     // 0/0 -> NaN like the reference when the mask is all zero
    out[1] = static_cast<float>(1.0 / denom);
  }
}

template <typename TP, typename TI>
__global__ void __launch_bounds__(256, 4)
patch_mse_bwd_kernel(const TP* __restrict__ pred, const TI* __restrict__ img, const float* __restrict__ mask,
                     const float* __restrict__ fwd_out, const float* __restrict__ grad_out, const PatchGeom gm,
                     TP* __restrict__ dpred) {
  const int P = gm.p0 * gm.p1 * gm.p2;
  const int d2 = 256 % gm.p2, d1 = (256 / gm.p2) % gm.p1, d0 = 256 / (gm.p2 * gm.p1);
  const PixelWalk first(gm);
  const float coef0 = 2.f * __ldg(fwd_out + 1) * __ldg(grad_out);
  for (TokenWalk t(gm); t.bl < t.end; t.next(gm)) {
    const float w = mask ? __ldg(mask + t.bl) : 1.f;
    TP* db = dpred + static_cast<long long>(t.bl) * P * gm.C;
    if (w == 0.f) {
      for (int e = threadIdx.x; e < P * gm.C; e += 256) stf(db + e, 0.f);
      continue;
    }
    const float coef = coef0 * w;
    const TI* ib = img + t.image_base(gm);
    const TP* pb = pred + static_cast<long long>(t.bl) * P * gm.C;
    for (PixelWalk x = first; x.q < P; x.next(gm, d0, d1, d2)) {
      const TI* ip = ib + x.offset(gm);
      const long long e = static_cast<long long>(x.q) * gm.C;
      for (int c = 0; c < gm.C; ++c) stf(db + e + c, coef * (ldf(pb + e + c) - ldf(ip + c * gm.plane)));
    }
  }
}

static int patch_geom(const char* who, int B, int C, int G0, int G1, int G2, int p0, int p1, int p2, PatchGeom* gm) {
  if (B <= 0 || C <= 0 || G0 <= 0 || G1 <= 0 || G2 <= 0 || p0 <= 0 || p1 <= 0 || p2 <= 0) {
    set_last_error("%s: every dimension must be positive", who); return UCF_ERR_BAD_ARG;
  }
  const long long L = 1LL * G0 * G1 * G2;
  if (1LL * B * L > 0x7fffffffLL || 1LL * p0 * p1 * p2 * C > 0x7fffffffLL) {
    set_last_error("%s: B*L and patch elements must fit in 31 bits", who); return UCF_ERR_BAD_ARG;
  }
  const long long sy = 1LL * G2 * p2, sx = sy * G1 * p1;
  if (sx * p0 > 0x7fffffffLL) {
    set_last_error("%s: one slab of p0 image rows must stay below 2^31 elements", who); return UCF_ERR_BAD_ARG;
  }
  gm->B = B; gm->C = C; gm->L = static_cast<int>(L); gm->G0 = G0; gm->G1 = G1; gm->G2 = G2;
  gm->p0 = p0; gm->p1 = p1; gm->p2 = p2;
  gm->sy = static_cast<int>(sy);
  gm->sx = static_cast<int>(sx);
  gm->plane = sx * G0 * p0;
  return UCF_OK;
}

static int patch_mse_grid(int BL) {
  int g = num_sms() * 16;
  if (g > UCF_PATCH_MSE_MAX_BLOCKS) g = UCF_PATCH_MSE_MAX_BLOCKS;
  return BL < g ? BL : g;
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_adamw_multi(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const long long* counts, double lr, double beta1,
                               double beta2, double eps, double weight_decay, long long step, int maximize,
                               void* stream) {
  if (n <= 0) return UCF_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !counts) {
    set_last_error("adamw_multi: null pointer table"); return UCF_ERR_BAD_ARG;
  }
  if (step < 1) { set_last_error("adamw_multi: step must be >= 1 (got %lld)", step); return UCF_ERR_BAD_ARG; }
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0)) {
    set_last_error("adamw_multi: betas must lie in [0, 1)"); return UCF_ERR_BAD_ARG;
  }
  AdamWScalars s;
  s.decay = static_cast<float>(1.0 - lr * weight_decay);
  s.beta1 = static_cast<float>(beta1); s.beta2 = static_cast<float>(beta2);
  s.one_m_beta1 = static_cast<float>(1.0 - beta1); s.one_m_beta2 = static_cast<float>(1.0 - beta2);
  s.step_size = static_cast<float>(lr / (1.0 - pow(beta1, static_cast<double>(step))));
  s.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, static_cast<double>(step))));
  s.eps = static_cast<float>(eps);
  s.gsign = maximize ? -1.f : 1.f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int base = 0; base < n; base += kAdamTensors) {
    const int cnt = n - base < kAdamTensors ? n - base : kAdamTensors;
    AdamWArgs a;
    long long nmax = 0;
    bool vec = true;
    for (int i = 0; i < kAdamTensors; ++i) {
      const int j = base + (i < cnt ? i : 0);
      a.p[i] = params[j]; a.g[i] = grads[j]; a.m[i] = exp_avg[j]; a.v[i] = exp_avg_sq[j];
      a.n[i] = i < cnt ? counts[j] : 0;
      if (i < cnt) {
        if (counts[j] < 0 || (counts[j] > 0 && (!params[j] || !grads[j] || !exp_avg[j] || !exp_avg_sq[j]))) {
          set_last_error("adamw_multi: tensor %d has a null pointer or a negative count", j); return UCF_ERR_BAD_ARG;
        }
        const uintptr_t bits = reinterpret_cast<uintptr_t>(params[j]) | reinterpret_cast<uintptr_t>(grads[j]) |
                               reinterpret_cast<uintptr_t>(exp_avg[j]) | reinterpret_cast<uintptr_t>(exp_avg_sq[j]);
        if (bits & 3) { set_last_error("adamw_multi: tensor %d is not 4-byte aligned", j); return UCF_ERR_BAD_ARG; }
        if (bits & 15) vec = false;
        if (counts[j] > nmax) nmax = counts[j];
      }
    }
    if (nmax == 0) continue;
    long long gx = (nmax / (vec ? 4 : 1) + 255) / 256;
    const long long cap = (static_cast<long long>(num_sms()) * 16 + cnt - 1) / cnt;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(cnt));
    if (vec) adamw_multi_kernel<true><<<grid, 256, 0, st>>>(a, s);
    else adamw_multi_kernel<false><<<grid, 256, 0, st>>>(a, s);
    const int rc = check_launch("adamw_multi_kernel");
    if (rc != UCF_OK) return rc;
  }
  return UCF_OK;
}

#define UCF_PATCH_MSE_DISPATCH(KERNEL, ...)                                                                  \
  do {                                                                                                       \
    if (pred_dtype == UCF_DTYPE_F32 && img_dtype == UCF_DTYPE_F32) KERNEL(float, float, __VA_ARGS__);        \
    else if (pred_dtype == UCF_DTYPE_F32 && img_dtype == UCF_DTYPE_BF16) KERNEL(float, __nv_bfloat16, __VA_ARGS__); \
    else if (pred_dtype == UCF_DTYPE_BF16 && img_dtype == UCF_DTYPE_F32) KERNEL(__nv_bfloat16, float, __VA_ARGS__); \
    else KERNEL(__nv_bfloat16, __nv_bfloat16, __VA_ARGS__);                                                  \
  } while (0)

static int patch_mse_dtypes(const char* who, int pred_dtype, int img_dtype) {
  const bool okp = pred_dtype == UCF_DTYPE_F32 || pred_dtype == UCF_DTYPE_BF16;
  const bool oki = img_dtype == UCF_DTYPE_F32 || img_dtype == UCF_DTYPE_BF16;
  if (!okp || !oki) { set_last_error("%s: pred and image must be f32 or bf16", who); return UCF_ERR_BAD_ARG; }
  return UCF_OK;
}

extern "C" int ucf_patch_mse_fwd(const void* pred, int pred_dtype, const void* img, int img_dtype, const float* mask,
                                 int B, int C, int G0, int G1, int G2, int p0, int p1, int p2, double* workspace,
                                 float* out, void* stream) {
  PatchGeom gm;
  int rc = patch_geom("patch_mse_fwd", B, C, G0, G1, G2, p0, p1, p2, &gm);
  if (rc == UCF_OK) rc = patch_mse_dtypes("patch_mse_fwd", pred_dtype, img_dtype);
  if (rc != UCF_OK) return rc;
  if (!pred || !img || !workspace || !out) { set_last_error("patch_mse_fwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int BL = gm.B * gm.L;
  const int grid = patch_mse_grid(BL);
#define UCF_FWD(TP, TI, dummy)                                                                         \
  patch_mse_fwd_kernel<TP, TI><<<grid, 256, 0, st>>>(static_cast<const TP*>(pred), static_cast<const TI*>(img), \
                                                     mask, gm, workspace)
  UCF_PATCH_MSE_DISPATCH(UCF_FWD, 0);
#undef UCF_FWD
  rc = check_launch("patch_mse_fwd_kernel");
  if (rc != UCF_OK) return rc;
  patch_mse_finish_kernel<<<1, 256, 0, st>>>(workspace, grid, mask, BL, static_cast<double>(p0) * p1 * p2 * C, out);
  return check_launch("patch_mse_finish_kernel");
}

extern "C" int ucf_patch_mse_bwd(const void* pred, int pred_dtype, const void* img, int img_dtype, const float* mask,
                                 const float* fwd_out, const float* grad_out, int B, int C, int G0, int G1, int G2,
                                 int p0, int p1, int p2, void* dpred, void* stream) {
  PatchGeom gm;
  int rc = patch_geom("patch_mse_bwd", B, C, G0, G1, G2, p0, p1, p2, &gm);
  if (rc == UCF_OK) rc = patch_mse_dtypes("patch_mse_bwd", pred_dtype, img_dtype);
  if (rc != UCF_OK) return rc;
  if (!pred || !img || !fwd_out || !grad_out || !dpred) {
    set_last_error("patch_mse_bwd: null pointer"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = patch_mse_grid(gm.B * gm.L);
#define UCF_BWD(TP, TI, dummy)                                                                         \
  patch_mse_bwd_kernel<TP, TI><<<grid, 256, 0, st>>>(static_cast<const TP*>(pred), static_cast<const TI*>(img), \
                                                     mask, fwd_out, grad_out, gm, static_cast<TP*>(dpred))
  UCF_PATCH_MSE_DISPATCH(UCF_BWD, 0);
#undef UCF_BWD
  return check_launch("patch_mse_bwd_kernel");
}
