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
constexpr int kAdamTensors = 24;     // per launch: 24 * 40 B of pointers and counts + 25 chunk offsets
constexpr int kAdamChunk = 4096;     // elements one CTA updates: 256 threads x 4 float4

struct AdamWArgs {
  float* p[kAdamTensors];
  const float* g[kAdamTensors];
  float* m[kAdamTensors];
  float* v[kAdamTensors];
  long long n[kAdamTensors];
  int first_chunk[kAdamTensors + 1];   // CTA b works on tensor t with first_chunk[t] <= b < first_chunk[t + 1]
};

struct AdamWScalars {
  float decay;        // 1 - lr * weight_decay
  float beta1, beta2;
  float one_m_beta1, one_m_beta2;
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2_sqrt; // 1 / sqrt(1 - beta2^t)
  float eps;
  float gsign;        // -1 when maximizing
  // CUDA-graph capturable form (ucf_adamw_multi_dev): learning rate and step count live in device memory and the three
  // step-dependent scalars above are derived from them inside the kernel, so a captured launch stays valid as both change
  const float* lr_dev;
  const float* step_dev;
  float weight_decay;
};

__device__ __forceinline__ AdamWScalars adamw_resolve(AdamWScalars s) {
  if (s.lr_dev != nullptr) {
    const float lr = *s.lr_dev, t = *s.step_dev;
    s.decay = 1.0f - lr * s.weight_decay;
    s.step_size = lr / (1.0f - powf(s.beta1, t));
    s.inv_bc2_sqrt = rsqrtf(1.0f - powf(s.beta2, t));
  }
  return s;
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWScalars& s) {
  g *= s.gsign;
  p *= s.decay;
  m = fmaf(s.one_m_beta1, g - m, m);                 // lerp(m, g, 1 - beta1)
  v = fmaf(s.beta2, v, s.one_m_beta2 * g * g);
  const float denom = fmaf(sqrtf(v), s.inv_bc2_sqrt, s.eps);
  p -= s.step_size * (m / denom);
}

// One CTA per 4096-element chunk of one tensor, so tensors of any mix of sizes load the machine evenly.
// VEC: all four streams of every tensor are 16-byte aligned -> float4, four independent chunks of loads in
// flight per thread; otherwise a scalar walk over the same chunk.
template <bool VEC>
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamWArgs a, const AdamWScalars s_in) {
  const AdamWScalars s = adamw_resolve(s_in);
  int t = 0;
#pragma unroll
  for (int i = 1; i < kAdamTensors; ++i) t += (static_cast<int>(blockIdx.x) >= a.first_chunk[i]) ? 1 : 0;
  const long long base = static_cast<long long>(static_cast<int>(blockIdx.x) - a.first_chunk[t]) * kAdamChunk;
  const long long left = a.n[t] - base;
  const int cnt = left < kAdamChunk ? static_cast<int>(left) : kAdamChunk;
  float* __restrict__ p = a.p[t] + base;
  const float* __restrict__ g = a.g[t] + base;
  float* __restrict__ m = a.m[t] + base;
  float* __restrict__ v = a.v[t] + base;
  if (VEC) {
    float4 pp[4], gg[4], mm[4], vv[4];
    const int nv = cnt >> 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + 256 * j;
      if (i < nv) {
        pp[j] = reinterpret_cast<const float4*>(p)[i];
        gg[j] = __ldg(reinterpret_cast<const float4*>(g) + i);
        mm[j] = reinterpret_cast<const float4*>(m)[i];
        vv[j] = reinterpret_cast<const float4*>(v)[i];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + 256 * j;
      if (i < nv) {
        adamw_one(pp[j].x, gg[j].x, mm[j].x, vv[j].x, s);
        adamw_one(pp[j].y, gg[j].y, mm[j].y, vv[j].y, s);
        adamw_one(pp[j].z, gg[j].z, mm[j].z, vv[j].z, s);
        adamw_one(pp[j].w, gg[j].w, mm[j].w, vv[j].w, s);
        reinterpret_cast<float4*>(p)[i] = pp[j];
        reinterpret_cast<float4*>(m)[i] = mm[j];
        reinterpret_cast<float4*>(v)[i] = vv[j];
      }
    }
    const int i = (nv << 2) + threadIdx.x;             // < 4 trailing elements of the tensor's last chunk
    if (i < cnt) {
      float q = p[i], mq = m[i], vq = v[i];
      adamw_one(q, g[i], mq, vq, s);
      p[i] = q; m[i] = mq; v[i] = vq;
    }
  } else {
    for (int i = threadIdx.x; i < cnt; i += 256) {
      float q = p[i], mq = m[i], vq = v[i];
      adamw_one(q, g[i], mq, vq, s);
      p[i] = q; m[i] = mq; v[i] = vq;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reconstruction loss against the patchified image
//
// pred [B, L, P*C] (P = p0*p1*p2 pixels of a patch, channel fastest) is compared with
// img [B, C, G0*p0, G1*p1, G2*p2], token l = (g0*G1 + g1)*G2 + g2, pixel q = (q0*p1 + q1)*p2 + q2.
// Fast path ("quad"): p2 % 4 == 0 and C <= 4.  A work item is four consecutive pixels along the image's
// fastest axis: C 16-byte image loads (one per channel plane) against the 4*C contiguous pred values they
// pair with, all issued before the first subtraction.  Items are dealt to CTAs as contiguous ranges of the
// flattened (token, quad) index, so small patches (64 quads at p = 16) fill every lane.
// Generic path: one pixel per thread, scalar loads, any geometry.
// ------------------------------------------------------------------------------------------------
struct PatchGeom {
  int B, C, L;
  int G0, G1, G2;
  int p0, p1, p2;
  int sx, sy;            // image strides (elements) of axis 0 / axis 1 inside one channel plane; axis 2 is 1
  long long plane;       // X*Y*Z
  long long items;       // quad path: B*L*(P/4)
  int per_cta;           // quad path: items per CTA
};

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
// four consecutive elements as floats (16-byte / 8-byte aligned)
__device__ __forceinline__ void ld4(const float* p, float* o) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float* o) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// sum of two per-thread doubles over the CTA, in a fixed order (bit-reproducible); valid on thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
  }
}

// every CTA adds its strided share of the mask to the second half of the workspace
__device__ __forceinline__ double mask_share(const float* __restrict__ mask, int BL) {
  double ms = 0.0;
  if (mask)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < BL; i += gridDim.x * 256) ms += static_cast<double>(__ldg(mask + i));
  return ms;
}

// (token, quad) of a CTA-local item index; offsets of the quad in pred and in channel plane 0 of the image
struct QuadItem {
  int token;
  long long pred_off, img_off;
  __device__ __forceinline__ QuadItem(const PatchGeom& gm, unsigned token0, unsigned r, unsigned Qn, int C) {
    const unsigned tk = r / Qn, quad = r - tk * Qn;
    token = static_cast<int>(token0 + tk);
    const unsigned b = static_cast<unsigned>(token) / static_cast<unsigned>(gm.L);
    const unsigned l = static_cast<unsigned>(token) - b * gm.L;
    const unsigned g2 = l % gm.G2, lr = l / gm.G2, g1 = lr % gm.G1, g0 = lr / gm.G1;
    const unsigned q = quad * 4u;
    const unsigned q2 = q % gm.p2, qr = q / gm.p2, q1 = qr % gm.p1, q0 = qr / gm.p1;
    pred_off = (static_cast<long long>(token) * Qn + quad) * (4 * C);
    img_off = static_cast<long long>(b) * C * gm.plane + static_cast<long long>(g0 * gm.p0 + q0) * gm.sx +
              static_cast<long long>(g1 * gm.p1 + q1) * gm.sy + (g2 * gm.p2 + q2);
  }
};

template <typename TP, typename TI, int C>
__global__ void __launch_bounds__(256)
patch_mse_fwd_quad_kernel(const TP* __restrict__ pred, const TI* __restrict__ img, const float* __restrict__ mask,
                          const PatchGeom gm, double* __restrict__ partials) {
  const unsigned Qn = static_cast<unsigned>(gm.p0 * gm.p1 * gm.p2) >> 2;
  const long long start = static_cast<long long>(blockIdx.x) * gm.per_cta;
  const long long stop = start + gm.per_cta < gm.items ? start + gm.per_cta : gm.items;
  const unsigned token0 = static_cast<unsigned>(start / Qn), rem0 = static_cast<unsigned>(start - 1LL * token0 * Qn);
  const int n_here = stop > start ? static_cast<int>(stop - start) : 0;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_here; i += 256) {
    const QuadItem it(gm, token0, rem0 + i, Qn, C);
    const float w = mask ? __ldg(mask + it.token) : 1.f;
    if (w == 0.f) continue;
    float iv[C][4], pv[4 * C];
#pragma unroll
    for (int c = 0; c < C; ++c) ld4(img + it.img_off + c * gm.plane, iv[c]);
#pragma unroll
    for (int j = 0; j < C; ++j) ld4(pred + it.pred_off + 4 * j, pv + 4 * j);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 4 * C; ++e) {
      const float d = pv[e] - iv[e % C][e / C];
      s = fmaf(d, d, s);
    }
    acc = fmaf(w, s, acc);
  }
  double d = static_cast<double>(acc), ms = mask_share(mask, gm.B * gm.L);
  block_sum2(d, ms);
  if (threadIdx.x == 0) { partials[blockIdx.x] = d; partials[UCF_PATCH_MSE_MAX_BLOCKS + blockIdx.x] = ms; }
}

template <typename TP, typename TI, int C>
__global__ void __launch_bounds__(256)
patch_mse_bwd_quad_kernel(const TP* __restrict__ pred, const TI* __restrict__ img, const float* __restrict__ mask,
                          const float* __restrict__ fwd_out, const float* __restrict__ grad_out, const PatchGeom gm,
                          TP* __restrict__ dpred) {
  const unsigned Qn = static_cast<unsigned>(gm.p0 * gm.p1 * gm.p2) >> 2;
  const long long start = static_cast<long long>(blockIdx.x) * gm.per_cta;
  const long long stop = start + gm.per_cta < gm.items ? start + gm.per_cta : gm.items;
  const unsigned token0 = static_cast<unsigned>(start / Qn), rem0 = static_cast<unsigned>(start - 1LL * token0 * Qn);
  const int n_here = stop > start ? static_cast<int>(stop - start) : 0;
  const float coef0 = 2.f * __ldg(fwd_out + 1) * __ldg(grad_out);
  for (int i = threadIdx.x; i < n_here; i += 256) {
    const QuadItem it(gm, token0, rem0 + i, Qn, C);
    const float w = mask ? __ldg(mask + it.token) : 1.f;
    float pv[4 * C];
    if (w == 0.f) {
#pragma unroll
      for (int e = 0; e < 4 * C; ++e) pv[e] = 0.f;
    } else {
      float iv[C][4];
#pragma unroll
      for (int c = 0; c < C; ++c) ld4(img + it.img_off + c * gm.plane, iv[c]);
#pragma unroll
      for (int j = 0; j < C; ++j) ld4(pred + it.pred_off + 4 * j, pv + 4 * j);
      const float coef = coef0 * w;
#pragma unroll
      for (int e = 0; e < 4 * C; ++e) pv[e] = coef * (pv[e] - iv[e % C][e / C]);
    }
#pragma unroll
    for (int j = 0; j < C; ++j) st4(dpred + it.pred_off + 4 * j, pv + 4 * j);
  }
}

// ---- generic path -----------------------------------------------------------------------------
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
  double d = static_cast<double>(acc), ms = mask_share(mask, gm.B * gm.L);
  block_sum2(d, ms);
  if (threadIdx.x == 0) { partials[blockIdx.x] = d; partials[UCF_PATCH_MSE_MAX_BLOCKS + blockIdx.x] = ms; }
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

// workspace = [MAX per-CTA sums of w * (pred - target)^2 | MAX per-CTA shares of sum(mask)]
// out[0] = loss, out[1] = 1 / denominator (kept for the backward pass)
__global__ void __launch_bounds__(256)
patch_mse_finish_kernel(const double* __restrict__ partials, int n_partials, int has_mask, int BL,
                        double elems_per_token, float* __restrict__ out) {
  double s = 0.0, ms = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += 256) {
    s += partials[i];
    ms += partials[UCF_PATCH_MSE_MAX_BLOCKS + i];
  }
  block_sum2(s, ms);
  if (threadIdx.x == 0) {
    const double denom = (has_mask ? ms : static_cast<double>(BL)) * elems_per_token;
    out[0] = static_cast<float>(s / denom);     // 0/0 -> NaN like the reference when the mask is all zero
    out[1] = static_cast<float>(1.0 / denom);
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
  gm->items = 0; gm->per_cta = 0;
  return UCF_OK;
}

static int patch_mse_max_grid() {
  const int g = num_sms() * 16;
  return g < UCF_PATCH_MSE_MAX_BLOCKS ? g : UCF_PATCH_MSE_MAX_BLOCKS;
}

// Quad path when the fastest patch axis is a multiple of four pixels, C <= 4 and both tensors are aligned for
// 4-element vector access; fills gm->items / per_cta and returns the grid, else returns 0 (generic path).
static int patch_mse_quad_grid(PatchGeom* gm, const void* pred, int pred_dtype, const void* img, int img_dtype,
                               const void* dpred) {
  if (gm->C > 4 || gm->p2 % 4) return 0;
  const uintptr_t pa = pred_dtype == UCF_DTYPE_F32 ? 15 : 7, ia = img_dtype == UCF_DTYPE_F32 ? 15 : 7;
  if ((reinterpret_cast<uintptr_t>(pred) & pa) || (reinterpret_cast<uintptr_t>(dpred) & pa) ||
      (reinterpret_cast<uintptr_t>(img) & ia)) return 0;
  const long long Qn = 1LL * gm->p0 * gm->p1 * gm->p2 / 4;
  gm->items = 1LL * gm->B * gm->L * Qn;
  long long grid = (gm->items + 255) / 256;
  if (grid > patch_mse_max_grid()) grid = patch_mse_max_grid();
  long long per = (gm->items + grid - 1) / grid;
  per = (per + 255) / 256 * 256;
  if (per + Qn > 0x7fffffffLL) return 0;
  gm->per_cta = static_cast<int>(per);
  return static_cast<int>((gm->items + per - 1) / per);
}

// ------------------------------------------------------------------------------------------------
// Dice + binary-cross-entropy loss of the SAP driver (utils/metrics.py:95-121)
//
// pred = sigmoid(logits)[:, 1:], true = targets[:, 1:], flattened:
//   loss = w * mean(BCE(pred, true)) + (1 - w) * (1 - (2 sum(pred*true) + sm) / (sum(pred) + sum(true) + sm))
// One pass reads both tensors once and carries the four sums; channel 0 is never touched.
// ------------------------------------------------------------------------------------------------
struct DiceGeom {
  long long HW;        // elements of one channel plane
  long long slab;      // (C - 1) * HW: the elements of one sample that count
  long long pitch;     // C * HW
  long long n;         // B * slab
};

__device__ __forceinline__ float dice_prob(float x, bool act) { return act ? 1.f / (1.f + expf(-x)) : x; }
// torch.nn.functional.binary_cross_entropy clamps both logarithms at -100
__device__ __forceinline__ float bce_term(float s, float t) {
  return -(t * fmaxf(logf(s), -100.f) + (1.f - t) * fmaxf(logf(1.f - s), -100.f));
}

template <typename TL, typename TT, bool VEC>
__global__ void __launch_bounds__(256)
dice_bce_fwd_kernel(const TL* __restrict__ logits, const TT* __restrict__ targets, const DiceGeom gm, int act,
                    double* __restrict__ partials) {
  float inter = 0.f, sp = 0.f, st = 0.f, bce = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const long long tid = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (VEC) {
    for (long long i = tid * 4; i < gm.n; i += stride * 4) {
      const long long b = i / gm.slab, off = b * gm.pitch + gm.HW + (i - b * gm.slab);
      float x[4], t[4];
      ld4(logits + off, x);
      ld4(targets + off, t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float s = dice_prob(x[e], act);
        inter = fmaf(s, t[e], inter); sp += s; st += t[e]; bce += bce_term(s, t[e]);
      }
    }
  } else {
    for (long long i = tid; i < gm.n; i += stride) {
      const long long b = i / gm.slab, off = b * gm.pitch + gm.HW + (i - b * gm.slab);
      const float s = dice_prob(ldf(logits + off), act), t = ldf(targets + off);
      inter = fmaf(s, t, inter); sp += s; st += t; bce += bce_term(s, t);
    }
  }
  double a = inter, c = bce, d = sp, e = st;
  block_sum2(a, c);
  __syncthreads();
  block_sum2(d, e);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = a;
    partials[UCF_PATCH_MSE_MAX_BLOCKS + blockIdx.x] = c;
    partials[2 * UCF_PATCH_MSE_MAX_BLOCKS + blockIdx.x] = d;
    partials[3 * UCF_PATCH_MSE_MAX_BLOCKS + blockIdx.x] = e;
  }
}

// out[0] = loss, out[1] = 2*I + smooth, out[2] = sum(pred) + sum(true) + smooth, out[3] = 1 / n
__global__ void __launch_bounds__(256)
dice_bce_finish_kernel(const double* __restrict__ partials, int n_partials, double n, float weight, float smooth,
                       float* __restrict__ out) {
  double a = 0.0, c = 0.0, d = 0.0, e = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += 256) {
    a += partials[i];
    c += partials[UCF_PATCH_MSE_MAX_BLOCKS + i];
    d += partials[2 * UCF_PATCH_MSE_MAX_BLOCKS + i];
    e += partials[3 * UCF_PATCH_MSE_MAX_BLOCKS + i];
  }
  block_sum2(a, c);
  __syncthreads();
  block_sum2(d, e);
  if (threadIdx.x == 0) {
    const double num = 2.0 * a + smooth, den = d + e + smooth;
    out[0] = static_cast<float>(weight * (c / n) + (1.0 - weight) * (1.0 - num / den));
    out[1] = static_cast<float>(num);
    out[2] = static_cast<float>(den);
    out[3] = static_cast<float>(1.0 / n);
  }
}

// d loss / d logits; one thread per element of the FULL tensor so channel 0 receives its zeros in the same pass
template <typename TL, typename TT, bool VEC>
__global__ void __launch_bounds__(256)
dice_bce_bwd_kernel(const TL* __restrict__ logits, const TT* __restrict__ targets, const float* __restrict__ fwd_out,
                    const float* __restrict__ grad_out, const DiceGeom gm, long long total, float weight, int act,
                    TL* __restrict__ dlogits) {
  const float g = __ldg(grad_out), num = __ldg(fwd_out + 1), den = __ldg(fwd_out + 2), inv_n = __ldg(fwd_out + 3);
  const float k_bce = g * weight * inv_n;
  const float k_t = -g * (1.f - weight) * 2.f / den;           // d dice / d pred = -2 t / den + num / den^2
  const float k_0 = g * (1.f - weight) * num / (den * den);
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  const long long tid = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  constexpr int W = VEC ? 4 : 1;
  for (long long i = tid * W; i < total; i += stride * W) {
    const bool counted = (i % gm.pitch) >= gm.HW;                // HW % 4 == 0 on the vector path: uniform per quad
    float x[W], t[W], r[W];
    if (counted) {
      if (VEC) { ld4(logits + i, x); ld4(targets + i, t); }
      else { x[0] = ldf(logits + i); t[0] = ldf(targets + i); }
    }
#pragma unroll
    for (int e = 0; e < W; ++e) {
      r[e] = 0.f;
      if (counted) {
        const float s = dice_prob(x[e], act);
        const float ds = k_bce * (s - t[e]) / fmaxf(s * (1.f - s), 1e-12f) + k_t * t[e] + k_0;
        r[e] = act ? ds * s * (1.f - s) : ds;
      }
    }
    if (VEC) st4(dlogits + i, r);
    else stf(dlogits + i, r[0]);
  }
}

static int dice_geom(const char* who, int B, int C, long long HW, DiceGeom* gm) {
  if (B <= 0 || C < 2 || HW <= 0) { set_last_error("%s: need B >= 1, C >= 2 and a non-empty plane", who); return UCF_ERR_BAD_ARG; }
  gm->HW = HW; gm->slab = (C - 1) * HW; gm->pitch = C * HW; gm->n = B * gm->slab;
  return UCF_OK;
}


// ------------------------------------------------------------------------------------------------
// Dice + cross-entropy loss of the UNETR driver: monai.losses.DiceCELoss(to_onehot_y=True,
// softmax=True, squared_pred=..., smooth_nr, smooth_dr) as configured at training_scripts/train_unetr_simple.py:38.
//   p = softmax(logits, 1); t = one_hot(target); per (b, c) over the S spatial positions
//   I = sum p t,  Q = sum p^2 (squared_pred) or sum p,  T = sum t
//   loss = ld * mean_{b,c}(1 - (2 I + snr) / (Q + T + sdr)) + lce * mean_{b,s}(-log p[target])
// logits [B, C, S] (C planes of S contiguous values), target [B, S] class indices.  One read of both tensors in each
// direction; a CTA works on voxels of ONE sample, so it carries 3 C + 1 partial sums (fp32 per thread, double across
// threads and CTAs, fixed order).  The finish kernel leaves, per (b, c), the two coefficients of d dice / d p = a t + b p
// (b p^(squared ? 1 : 0)) and the cross-entropy scale on the device for the backward pass.
// ------------------------------------------------------------------------------------------------
constexpr int kDiceCeMaxC = 8;

__device__ __forceinline__ int ld_class(const uint8_t* p) { return *p; }
__device__ __forceinline__ int ld_class(const long long* p) { return static_cast<int>(*p); }
__device__ __forceinline__ int ld_class(const float* p) { return static_cast<int>(*p); }

template <typename TL, typename TT>
__global__ void __launch_bounds__(256)
dice_ce_fwd_kernel(const TL* __restrict__ logits, const TT* __restrict__ target, int C, long long S, int nb, int squared,
                   double* __restrict__ partials, long long cs, long long ss) {     // logit (b, c, s) at b C S + c cs + s ss
  const int b = blockIdx.x / nb, j = blockIdx.x - b * nb;
  const long long per = (S + nb - 1) / nb;
  const long long s0 = j * per, s1 = s0 + per < S ? s0 + per : S;
  float si[kDiceCeMaxC], sq[kDiceCeMaxC], st[kDiceCeMaxC], ce = 0.f;
#pragma unroll
  for (int c = 0; c < kDiceCeMaxC; ++c) si[c] = sq[c] = st[c] = 0.f;
  const TL* lg = logits + static_cast<long long>(b) * C * S;
  const TT* tg = target + static_cast<long long>(b) * S;
  for (long long s = s0 + threadIdx.x; s < s1; s += 256) {
    float z[kDiceCeMaxC], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) { z[c] = ldf(lg + c * cs + s * ss); m = fmaxf(m, z[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) { z[c] = __expf(z[c] - m); sum += z[c]; }
    const float inv = 1.f / sum;
    const int t = ld_class(tg + s);
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) {
        const float pc = z[c] * inv;
        sq[c] += squared ? pc * pc : pc;
        if (c == t) { si[c] += pc; st[c] += 1.f; ce -= logf(fmaxf(pc, 1e-38f)); }
      }
  }
  double* out = partials + static_cast<long long>(blockIdx.x) * (3 * kDiceCeMaxC + 1);
#pragma unroll
  for (int c = 0; c < kDiceCeMaxC; ++c) {
    double a = si[c], q = sq[c];
    block_sum2(a, q);
    __syncthreads();
    double t2 = st[c], e = (c == 0) ? ce : 0.0;
    block_sum2(t2, e);
    __syncthreads();
    if (threadIdx.x == 0) {
      out[c] = a; out[kDiceCeMaxC + c] = q; out[2 * kDiceCeMaxC + c] = t2;
      if (c == 0) out[3 * kDiceCeMaxC] = e;
    }
  }
}

// out[0] = loss; out[1 + (b C + c)] = a_bc; out[1 + B C + (b C + c)] = b_bc; out[1 + 2 B C] = lambda_ce / (B S)
__global__ void __launch_bounds__(256)
dice_ce_finish_kernel(const double* __restrict__ partials, int B, int C, int nb, double S, int squared, float snr, float sdr,
                      float ld, float lce, float* __restrict__ out) {
  __shared__ double dice_s[256], ce_s[256];
  double dice = 0.0, ce = 0.0;
  for (int bc = threadIdx.x; bc < B * C; bc += 256) {
    const int b = bc / C, c = bc - b * C;
    double I = 0.0, Q = 0.0, T = 0.0;
    for (int j = 0; j < nb; ++j) {
      const double* pp = partials + static_cast<long long>(b * nb + j) * (3 * kDiceCeMaxC + 1);
      I += pp[c]; Q += pp[kDiceCeMaxC + c]; T += pp[2 * kDiceCeMaxC + c];
      if (c == 0) ce += pp[3 * kDiceCeMaxC];
    }
    const double num = 2.0 * I + snr, den = Q + T + sdr, w = static_cast<double>(ld) / (B * C);
    dice += 1.0 - num / den;
    out[1 + bc] = static_cast<float>(-2.0 / den * w);                                  // d dice / d p: coefficient of t
    out[1 + B * C + bc] = static_cast<float>((squared ? 2.0 : 1.0) * num / (den * den) * w);   // coefficient of p (squared) or 1
  }
  dice_s[threadIdx.x] = dice; ce_s[threadIdx.x] = ce;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0, e = 0.0;
    for (int i = 0; i < 256; ++i) { d += dice_s[i]; e += ce_s[i]; }
    out[0] = static_cast<float>(ld * d / (B * C) + lce * e / (B * S));
    out[1 + 2 * B * C] = static_cast<float>(lce / (B * S));
  }
}

template <typename TL, typename TT>
__global__ void __launch_bounds__(256)
dice_ce_bwd_kernel(const TL* __restrict__ logits, const TT* __restrict__ target, const float* __restrict__ fwd_out,
                   const float* __restrict__ grad_out, int B, int C, long long S, int squared, TL* __restrict__ dlogits,
                   long long cs, long long ss) {
  const float g = *grad_out, ces = fwd_out[1 + 2 * B * C];
  const long long total = static_cast<long long>(B) * S;
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += stride) {
    const int b = static_cast<int>(i / S);
    const long long s = i - b * S;
    const TL* lg = logits + static_cast<long long>(b) * C * S + s * ss;
    float z[kDiceCeMaxC], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) { z[c] = ldf(lg + c * cs); m = fmaxf(m, z[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) { z[c] = __expf(z[c] - m); sum += z[c]; }
    const float inv = 1.f / sum;
    const int t = ld_class(target + i);
    float G[kDiceCeMaxC], dot = 0.f;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) {
        z[c] *= inv;                                                     // p_c
        const float a = fwd_out[1 + b * C + c], bb = fwd_out[1 + B * C + b * C + c];
        G[c] = (c == t ? a : 0.f) + (squared ? bb * z[c] : bb);         // d(dice term) / d p_c
        dot = fmaf(z[c], G[c], dot);
      }
    TL* dl = dlogits + static_cast<long long>(b) * C * S + s * ss;
#pragma unroll
    for (int c = 0; c < kDiceCeMaxC; ++c)
      if (c < C) stf(dl + c * cs, g * (z[c] * (G[c] - dot) + ces * (z[c] - (c == t ? 1.f : 0.f))));
  }
}

}  // namespace ucf

using namespace ucf;

static int adamw_launch(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, const long long* counts, const AdamWScalars& s, void* stream);

extern "C" int ucf_adamw_multi(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const long long* counts, double lr, double beta1,
                               double beta2, double eps, double weight_decay, long long step, int maximize,
                               void* stream) {
  if (n <= 0) return UCF_OK;
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
  s.lr_dev = nullptr; s.step_dev = nullptr; s.weight_decay = static_cast<float>(weight_decay);
  return adamw_launch(n, params, grads, exp_avg, exp_avg_sq, counts, s, stream);
}

extern "C" int ucf_adamw_multi_dev(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                                   float* const* exp_avg_sq, const long long* counts, const float* lr_dev, double beta1,
                                   double beta2, double eps, double weight_decay, const float* step_dev, int maximize,
                                   void* stream) {
  if (n <= 0) return UCF_OK;
  if (!lr_dev || !step_dev) { set_last_error("adamw_multi_dev: null lr / step pointer"); return UCF_ERR_BAD_ARG; }
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0)) {
    set_last_error("adamw_multi_dev: betas must lie in [0, 1)"); return UCF_ERR_BAD_ARG;
  }
  AdamWScalars s;
  s.decay = 1.f; s.step_size = 0.f; s.inv_bc2_sqrt = 1.f;      // derived from *lr_dev / *step_dev inside the kernel
  s.beta1 = static_cast<float>(beta1); s.beta2 = static_cast<float>(beta2);
  s.one_m_beta1 = static_cast<float>(1.0 - beta1); s.one_m_beta2 = static_cast<float>(1.0 - beta2);
  s.eps = static_cast<float>(eps);
  s.gsign = maximize ? -1.f : 1.f;
  s.lr_dev = lr_dev; s.step_dev = step_dev; s.weight_decay = static_cast<float>(weight_decay);
  return adamw_launch(n, params, grads, exp_avg, exp_avg_sq, counts, s, stream);
}

static int adamw_launch(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, const long long* counts, const AdamWScalars& s, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !counts) {
    set_last_error("adamw_multi: null pointer table"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int base = 0; base < n; base += kAdamTensors) {
    const int cnt = n - base < kAdamTensors ? n - base : kAdamTensors;
    AdamWArgs a;
    long long chunks = 0;
    bool vec = true;
    for (int i = 0; i < kAdamTensors; ++i) {
      const int j = base + (i < cnt ? i : 0);
      a.p[i] = params[j]; a.g[i] = grads[j]; a.m[i] = exp_avg[j]; a.v[i] = exp_avg_sq[j];
      a.n[i] = i < cnt ? counts[j] : 0;
      a.first_chunk[i] = static_cast<int>(chunks);
      if (i < cnt) {
        if (counts[j] < 0 || (counts[j] > 0 && (!params[j] || !grads[j] || !exp_avg[j] || !exp_avg_sq[j]))) {
          set_last_error("adamw_multi: tensor %d has a null pointer or a negative count", j); return UCF_ERR_BAD_ARG;
        }
        const uintptr_t bits = reinterpret_cast<uintptr_t>(params[j]) | reinterpret_cast<uintptr_t>(grads[j]) |
                               reinterpret_cast<uintptr_t>(exp_avg[j]) | reinterpret_cast<uintptr_t>(exp_avg_sq[j]);
        if (bits & 3) { set_last_error("adamw_multi: tensor %d is not 4-byte aligned", j); return UCF_ERR_BAD_ARG; }
        if (counts[j] > 0 && (bits & 15)) vec = false;
        chunks += (counts[j] + kAdamChunk - 1) / kAdamChunk;
        if (chunks > 0x7fffffffLL) { set_last_error("adamw_multi: more than 2^31 chunks in one launch"); return UCF_ERR_BAD_ARG; }
      }
    }
    a.first_chunk[kAdamTensors] = static_cast<int>(chunks);
    if (chunks == 0) continue;
    if (vec) adamw_multi_kernel<true><<<static_cast<unsigned>(chunks), 256, 0, st>>>(a, s);
    else adamw_multi_kernel<false><<<static_cast<unsigned>(chunks), 256, 0, st>>>(a, s);
    const int rc = check_launch("adamw_multi_kernel");
    if (rc != UCF_OK) return rc;
  }
  return UCF_OK;
}

static int patch_mse_dtypes(const char* who, int pred_dtype, int img_dtype) {
  const bool okp = pred_dtype == UCF_DTYPE_F32 || pred_dtype == UCF_DTYPE_BF16;
  const bool oki = img_dtype == UCF_DTYPE_F32 || img_dtype == UCF_DTYPE_BF16;
  if (!okp || !oki) { set_last_error("%s: pred and image must be f32 or bf16", who); return UCF_ERR_BAD_ARG; }
  return UCF_OK;
}

namespace {

struct FwdLaunch {
  const void* pred; const void* img; const float* mask; PatchGeom gm; double* ws; int grid; cudaStream_t st;
  template <typename TP, typename TI, int C> void run() const {
    if (C == 0)
      patch_mse_fwd_kernel<TP, TI><<<grid, 256, 0, st>>>(static_cast<const TP*>(pred), static_cast<const TI*>(img), mask, gm, ws);
    else
      patch_mse_fwd_quad_kernel<TP, TI, (C == 0 ? 1 : C)><<<grid, 256, 0, st>>>(static_cast<const TP*>(pred),
                                                                              static_cast<const TI*>(img), mask, gm, ws);
  }
};

struct BwdLaunch {
  const void* pred; const void* img; const float* mask; const float* fwd_out; const float* grad_out; PatchGeom gm;
  void* dpred; int grid; cudaStream_t st;
  template <typename TP, typename TI, int C> void run() const {
    if (C == 0)
      patch_mse_bwd_kernel<TP, TI><<<grid, 256, 0, st>>>(static_cast<const TP*>(pred), static_cast<const TI*>(img), mask,
                                                         fwd_out, grad_out, gm, static_cast<TP*>(dpred));
    else
      patch_mse_bwd_quad_kernel<TP, TI, (C == 0 ? 1 : C)><<<grid, 256, 0, st>>>(
          static_cast<const TP*>(pred), static_cast<const TI*>(img), mask, fwd_out, grad_out, gm, static_cast<TP*>(dpred));
  }
};

// C = 0 selects the generic kernels
template <typename Launch, typename TP, typename TI> void by_channels(const Launch& l, int c) {
  switch (c) {
    case 1: l.template run<TP, TI, 1>(); break;
    case 2: l.template run<TP, TI, 2>(); break;
    case 3: l.template run<TP, TI, 3>(); break;
    case 4: l.template run<TP, TI, 4>(); break;
    default: l.template run<TP, TI, 0>(); break;
  }
}

template <typename Launch> void by_dtypes(const Launch& l, int pred_dtype, int img_dtype, int c) {
  const bool pf = pred_dtype == UCF_DTYPE_F32, imf = img_dtype == UCF_DTYPE_F32;
  if (pf && imf) by_channels<Launch, float, float>(l, c);
  else if (pf) by_channels<Launch, float, __nv_bfloat16>(l, c);
  else if (imf) by_channels<Launch, __nv_bfloat16, float>(l, c);
  else by_channels<Launch, __nv_bfloat16, __nv_bfloat16>(l, c);
}

}  // namespace

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
  int grid = patch_mse_quad_grid(&gm, pred, pred_dtype, img, img_dtype, nullptr);
  const bool quad = grid > 0;
  if (!quad) grid = BL < patch_mse_max_grid() ? BL : patch_mse_max_grid();
  by_dtypes(FwdLaunch{pred, img, mask, gm, workspace, grid, st}, pred_dtype, img_dtype, quad ? C : 0);
  rc = check_launch(quad ? "patch_mse_fwd_quad_kernel" : "patch_mse_fwd_kernel");
  if (rc != UCF_OK) return rc;
  patch_mse_finish_kernel<<<1, 256, 0, st>>>(workspace, grid, mask != nullptr, BL, static_cast<double>(p0) * p1 * p2 * C, out);
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
  const int BL = gm.B * gm.L;
  int grid = patch_mse_quad_grid(&gm, pred, pred_dtype, img, img_dtype, dpred);
  const bool quad = grid > 0;
  if (!quad) grid = BL < patch_mse_max_grid() ? BL : patch_mse_max_grid();
  by_dtypes(BwdLaunch{pred, img, mask, fwd_out, grad_out, gm, dpred, grid, st}, pred_dtype, img_dtype, quad ? C : 0);
  return check_launch(quad ? "patch_mse_bwd_quad_kernel" : "patch_mse_bwd_kernel");
}

namespace {

template <typename TL, typename TT>
void dice_fwd_launch(const void* logits, const void* targets, const DiceGeom& gm, int act, double* ws, int grid, bool vec,
                     cudaStream_t st) {
  if (vec) dice_bce_fwd_kernel<TL, TT, true><<<grid, 256, 0, st>>>(static_cast<const TL*>(logits), static_cast<const TT*>(targets), gm, act, ws);
  else dice_bce_fwd_kernel<TL, TT, false><<<grid, 256, 0, st>>>(static_cast<const TL*>(logits), static_cast<const TT*>(targets), gm, act, ws);
}

template <typename TL, typename TT>
void dice_bwd_launch(const void* logits, const void* targets, const float* fwd_out, const float* grad_out, const DiceGeom& gm,
                     long long total, float weight, int act, void* dlogits, int grid, bool vec, cudaStream_t st) {
  if (vec)
    dice_bce_bwd_kernel<TL, TT, true><<<grid, 256, 0, st>>>(static_cast<const TL*>(logits), static_cast<const TT*>(targets),
                                                            fwd_out, grad_out, gm, total, weight, act, static_cast<TL*>(dlogits));
  else
    dice_bce_bwd_kernel<TL, TT, false><<<grid, 256, 0, st>>>(static_cast<const TL*>(logits), static_cast<const TT*>(targets),
                                                             fwd_out, grad_out, gm, total, weight, act, static_cast<TL*>(dlogits));
}

bool dice_vec_ok(const DiceGeom& gm, const void* a, int a_dtype, const void* b, int b_dtype, const void* c) {
  if (gm.HW % 4) return false;
  const uintptr_t ma = a_dtype == UCF_DTYPE_F32 ? 15 : 7, mb = b_dtype == UCF_DTYPE_F32 ? 15 : 7;
  return !((reinterpret_cast<uintptr_t>(a) & ma) || (reinterpret_cast<uintptr_t>(c) & ma) || (reinterpret_cast<uintptr_t>(b) & mb));
}

int dice_grid(long long items) {
  long long g = (items + 255) / 256;
  if (g > patch_mse_max_grid()) g = patch_mse_max_grid();
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int ucf_dice_bce_fwd(const void* logits, int logits_dtype, const void* targets, int targets_dtype, int B, int C,
                                long long HW, float weight, float smooth, int act, double* workspace, float* out,
                                void* stream) {
  DiceGeom gm;
  int rc = dice_geom("dice_bce_fwd", B, C, HW, &gm);
  if (rc == UCF_OK) rc = patch_mse_dtypes("dice_bce_fwd", logits_dtype, targets_dtype);
  if (rc != UCF_OK) return rc;
  if (!logits || !targets || !workspace || !out) { set_last_error("dice_bce_fwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = dice_vec_ok(gm, logits, logits_dtype, targets, targets_dtype, nullptr);
  const int grid = dice_grid(vec ? gm.n / 4 : gm.n);
  const bool lf = logits_dtype == UCF_DTYPE_F32, tf = targets_dtype == UCF_DTYPE_F32;
  if (lf && tf) dice_fwd_launch<float, float>(logits, targets, gm, act, workspace, grid, vec, st);
  else if (lf) dice_fwd_launch<float, __nv_bfloat16>(logits, targets, gm, act, workspace, grid, vec, st);
  else if (tf) dice_fwd_launch<__nv_bfloat16, float>(logits, targets, gm, act, workspace, grid, vec, st);
  else dice_fwd_launch<__nv_bfloat16, __nv_bfloat16>(logits, targets, gm, act, workspace, grid, vec, st);
  rc = check_launch("dice_bce_fwd_kernel");
  if (rc != UCF_OK) return rc;
  dice_bce_finish_kernel<<<1, 256, 0, st>>>(workspace, grid, static_cast<double>(gm.n), weight, smooth, out);
  return check_launch("dice_bce_finish_kernel");
}

extern "C" int ucf_dice_bce_bwd(const void* logits, int logits_dtype, const void* targets, int targets_dtype,
                                const float* fwd_out, const float* grad_out, int B, int C, long long HW, float weight,
                                int act, void* dlogits, void* stream) {
  DiceGeom gm;
  int rc = dice_geom("dice_bce_bwd", B, C, HW, &gm);
  if (rc == UCF_OK) rc = patch_mse_dtypes("dice_bce_bwd", logits_dtype, targets_dtype);
  if (rc != UCF_OK) return rc;
  if (!logits || !targets || !fwd_out || !grad_out || !dlogits) { set_last_error("dice_bce_bwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = dice_vec_ok(gm, logits, logits_dtype, targets, targets_dtype, dlogits);
  const long long total = static_cast<long long>(B) * gm.pitch;
  const int grid = dice_grid(vec ? total / 4 : total);
  const bool lf = logits_dtype == UCF_DTYPE_F32, tf = targets_dtype == UCF_DTYPE_F32;
  if (lf && tf) dice_bwd_launch<float, float>(logits, targets, fwd_out, grad_out, gm, total, weight, act, dlogits, grid, vec, st);
  else if (lf) dice_bwd_launch<float, __nv_bfloat16>(logits, targets, fwd_out, grad_out, gm, total, weight, act, dlogits, grid, vec, st);
  else if (tf) dice_bwd_launch<__nv_bfloat16, float>(logits, targets, fwd_out, grad_out, gm, total, weight, act, dlogits, grid, vec, st);
  else dice_bwd_launch<__nv_bfloat16, __nv_bfloat16>(logits, targets, fwd_out, grad_out, gm, total, weight, act, dlogits, grid, vec, st);
  return check_launch("dice_bce_bwd_kernel");
}

namespace {
template <typename TL>
int dice_ce_dispatch_fwd(const void* logits, const void* target, int tdt, int C, long long S, int nb, int squared, double* ws,
                         int grid, cudaStream_t st, long long cs, long long ss) {
  const TL* lg = static_cast<const TL*>(logits);
  if (tdt == UCF_DTYPE_U8) dice_ce_fwd_kernel<TL, uint8_t><<<grid, 256, 0, st>>>(lg, static_cast<const uint8_t*>(target), C, S, nb, squared, ws, cs, ss);
  else if (tdt == UCF_DTYPE_I64) dice_ce_fwd_kernel<TL, long long><<<grid, 256, 0, st>>>(lg, static_cast<const long long*>(target), C, S, nb, squared, ws, cs, ss);
  else dice_ce_fwd_kernel<TL, float><<<grid, 256, 0, st>>>(lg, static_cast<const float*>(target), C, S, nb, squared, ws, cs, ss);
  return check_launch("dice_ce_fwd_kernel");
}
template <typename TL>
int dice_ce_dispatch_bwd(const void* logits, const void* target, int tdt, const float* fwd_out, const float* grad_out, int B, int C,
                         long long S, int squared, void* dlogits, int grid, cudaStream_t st, long long cs, long long ss) {
  const TL* lg = static_cast<const TL*>(logits);
  TL* dl = static_cast<TL*>(dlogits);
  if (tdt == UCF_DTYPE_U8) dice_ce_bwd_kernel<TL, uint8_t><<<grid, 256, 0, st>>>(lg, static_cast<const uint8_t*>(target), fwd_out, grad_out, B, C, S, squared, dl, cs, ss);
  else if (tdt == UCF_DTYPE_I64) dice_ce_bwd_kernel<TL, long long><<<grid, 256, 0, st>>>(lg, static_cast<const long long*>(target), fwd_out, grad_out, B, C, S, squared, dl, cs, ss);
  else dice_ce_bwd_kernel<TL, float><<<grid, 256, 0, st>>>(lg, static_cast<const float*>(target), fwd_out, grad_out, B, C, S, squared, dl, cs, ss);
  return check_launch("dice_ce_bwd_kernel");
}
int dice_ce_check(const char* who, int ldt, int tdt, int B, int C, long long S) {
  if (B <= 0 || C < 2 || C > kDiceCeMaxC || S <= 0) { set_last_error("%s: need B > 0, 2 <= C <= %d, S > 0 (got B=%d C=%d S=%lld)", who, kDiceCeMaxC, B, C, S); return UCF_ERR_BAD_ARG; }
  if (ldt != UCF_DTYPE_F32 && ldt != UCF_DTYPE_BF16) { set_last_error("%s: logits must be f32 or bf16", who); return UCF_ERR_BAD_ARG; }
  if (tdt != UCF_DTYPE_U8 && tdt != UCF_DTYPE_I64 && tdt != UCF_DTYPE_F32) { set_last_error("%s: target must hold class indices as u8, i64 or f32", who); return UCF_ERR_BAD_ARG; }
  return UCF_OK;
}
}  // namespace

extern "C" int ucf_dice_ce_blocks_per_sample(int B, long long S) {
  long long nb = (S + 8191) / 8192;
  const long long cap = B > 0 ? (2048 / B > 1 ? 2048 / B : 1) : 1;
  if (nb > cap) nb = cap;
  return static_cast<int>(nb < 1 ? 1 : nb);
}

extern "C" int ucf_dice_ce_fwd(const void* logits, int logits_dtype, const void* target, int target_dtype, int B, int C,
                               long long S, int squared_pred, float smooth_nr, float smooth_dr, float lambda_dice,
                               float lambda_ce, double* workspace, float* out, int channels_last, void* stream) {
  int rc = dice_ce_check("dice_ce_fwd", logits_dtype, target_dtype, B, C, S);
  const long long cs = channels_last ? 1 : S, ss = channels_last ? C : 1;
  if (rc != UCF_OK) return rc;
  if (!logits || !target || !workspace || !out) { set_last_error("dice_ce_fwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = ucf_dice_ce_blocks_per_sample(B, S);
  rc = logits_dtype == UCF_DTYPE_F32 ? dice_ce_dispatch_fwd<float>(logits, target, target_dtype, C, S, nb, squared_pred, workspace, B * nb, st, cs, ss)
                                     : dice_ce_dispatch_fwd<__nv_bfloat16>(logits, target, target_dtype, C, S, nb, squared_pred, workspace, B * nb, st, cs, ss);
  if (rc != UCF_OK) return rc;
  dice_ce_finish_kernel<<<1, 256, 0, st>>>(workspace, B, C, nb, static_cast<double>(S), squared_pred, smooth_nr, smooth_dr, lambda_dice,
                                           lambda_ce, out);
  return check_launch("dice_ce_finish_kernel");
}

extern "C" int ucf_dice_ce_bwd(const void* logits, int logits_dtype, const void* target, int target_dtype, const float* fwd_out,
                               const float* grad_out, int B, int C, long long S, int squared_pred, void* dlogits, int channels_last,
                               void* stream) {
  int rc = dice_ce_check("dice_ce_bwd", logits_dtype, target_dtype, B, C, S);
  const long long cs = channels_last ? 1 : S, ss = channels_last ? C : 1;
  if (rc != UCF_OK) return rc;
  if (!logits || !target || !fwd_out || !grad_out || !dlogits) { set_last_error("dice_ce_bwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = dice_grid(static_cast<long long>(B) * S);
  return logits_dtype == UCF_DTYPE_F32
             ? dice_ce_dispatch_bwd<float>(logits, target, target_dtype, fwd_out, grad_out, B, C, S, squared_pred, dlogits, grid, st, cs, ss)
             : dice_ce_dispatch_bwd<__nv_bfloat16>(logits, target, target_dtype, fwd_out, grad_out, B, C, S, squared_pred, dlogits, grid, st, cs, ss);
}
