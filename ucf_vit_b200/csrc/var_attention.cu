// Channel (variable) aggregation cross-attention core: N_a query tokens (normally 1) attend over
// the V per-variable tokens of one spatial location.  V is tiny, so this is HBM-bandwidth-bound:
// one warp per (row, head) streams the kv projection once with an online softmax; no tensor cores.
// Replaces the SDPA call inside VariableMapping_Attention.forward
// (/root/reference/src/UCF_VIT/simple/building_blocks.py:339-367) as used by
// VIT.aggregate_variables (/root/reference/src/UCF_VIT/simple/arch.py:414-432).
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// q  : [Bq, Na, H, HD] bf16, Bq == rows or 1 (query shared by every row)
// kv : [rows, V, 2, H, HD] bf16
// o  : [rows, Na, H, HD] bf16      lse : [rows, Na, H] fp32
template <int EPL>   // elements per lane = HD / 32
__global__ void __launch_bounds__(256)
var_attn_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                    __nv_bfloat16* __restrict__ o, float* __restrict__ lse, long long rows, int Na, int V, int H,
                    int q_shared, float scale) {
  constexpr int HD = EPL * 32;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long total = rows * Na * H;
  for (long long item = warp0; item < total; item += nwarps) {
    const int h = static_cast<int>(item % H);
    const long long r2 = item / H;
    const int a = static_cast<int>(r2 % Na);
    const long long row = r2 / Na;
    const __nv_bfloat16* qp = q + (((q_shared ? 0 : row) * Na + a) * H + h) * HD + lane * EPL;
    float qv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) qv[e] = __bfloat162float(qp[e]) * scale;
    float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
    const __nv_bfloat16* kvp = kv + row * V * 2 * H * HD + h * HD + lane * EPL;
    for (int v = 0; v < V; ++v) {
      const __nv_bfloat16* kp = kvp + static_cast<long long>(v) * 2 * H * HD;
      const __nv_bfloat16* vp = kp + H * HD;
      float s = 0.f, vv[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) { s += qv[e] * __bfloat162float(kp[e]); vv[e] = __bfloat162float(vp[e]); }
      s = warp_sum(s);
      const float mn = fmaxf(m, s);
      const float alpha = __expf(m - mn), p = __expf(s - mn);
      l = l * alpha + p;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * alpha + p * vv[e];
      m = mn;
    }
    const float inv = 1.f / l;
    __nv_bfloat16* op = o + ((row * Na + a) * H + h) * HD + lane * EPL;
#pragma unroll
    for (int e = 0; e < EPL; ++e) op[e] = __float2bfloat16(acc[e] * inv);
    if (lane == 0) lse[(row * Na + a) * H + h] = m + __logf(l);
  }
}

// dkv : [rows, V, 2, H, HD] bf16 (fully written); dq_acc : fp32 [Bq, Na, H, HD] (accumulated with
// atomics when the query is shared, plain store otherwise).  One warp per (row, head), looping a.
template <int EPL>
__global__ void __launch_bounds__(256)
var_attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                    const float* __restrict__ lse, __nv_bfloat16* __restrict__ dkv, float* __restrict__ dq_acc,
                    long long rows, int Na, int V, int H, int q_shared, float scale) {
  constexpr int HD = EPL * 32;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long total = rows * H;
  for (long long item = warp0; item < total; item += nwarps) {
    const int h = static_cast<int>(item % H);
    const long long row = item / H;
    const __nv_bfloat16* kvp = kv + row * V * 2 * H * HD + h * HD + lane * EPL;
    __nv_bfloat16* dkvp = dkv + row * V * 2 * H * HD + h * HD + lane * EPL;
    for (int a = 0; a < Na; ++a) {
      const long long oi = ((row * Na + a) * H + h);
      const long long qi = (((q_shared ? 0 : row) * Na + a) * H + h);
      float qv[EPL], dov[EPL], dq[EPL];
      float delta = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        qv[e] = __bfloat162float(q[qi * HD + lane * EPL + e]);
        dov[e] = __bfloat162float(d_o[oi * HD + lane * EPL + e]);
        delta += dov[e] * __bfloat162float(o[oi * HD + lane * EPL + e]);
        dq[e] = 0.f;
      }
      delta = warp_sum(delta);
      const float L = lse[oi];
      for (int v = 0; v < V; ++v) {
        const long long off = static_cast<long long>(v) * 2 * H * HD;
        float kk[EPL], vv[EPL], s = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          kk[e] = __bfloat162float(kvp[off + e]);
          vv[e] = __bfloat162float(kvp[off + H * HD + e]);
          s += qv[e] * kk[e];
          dp += dov[e] * vv[e];
        }
        s = warp_sum(s) * scale;
        dp = warp_sum(dp);
        const float p = __expf(s - L);
        const float ds = p * (dp - delta) * scale;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          dq[e] += ds * kk[e];
          float gk = ds * qv[e], gv = p * dov[e];
          if (a > 0) {
            gk += __bfloat162float(dkvp[off + e]);
            gv += __bfloat162float(dkvp[off + H * HD + e]);
          }
          dkvp[off + e] = __float2bfloat16(gk);
          dkvp[off + H * HD + e] = __float2bfloat16(gv);
        }
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        if (q_shared) atomicAdd(&dq_acc[qi * HD + lane * EPL + e], dq[e]);
        else dq_acc[qi * HD + lane * EPL + e] = dq[e];
      }
    }
  }
}

}  // namespace ucf

using namespace ucf;

static int va_grid(long long warps) {
  long long blocks = (warps + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

extern "C" int ucf_var_attention_fwd(const void* q, const void* kv, void* o, float* lse, long long rows, int Na,
                                     int V, int H, int hd, int q_shared, float scale, void* stream) {
  if (rows <= 0 || Na <= 0 || V <= 0 || H <= 0) { set_last_error("var_attention_fwd: empty problem"); return UCF_ERR_BAD_ARG; }
  if (hd != 32 && hd != 64 && hd != 128) { set_last_error("var_attention_fwd: head_dim %d not in {32,64,128}", hd); return UCF_ERR_UNSUPPORTED; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = va_grid(rows * Na * H);
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q); auto KV = reinterpret_cast<const __nv_bfloat16*>(kv);
  auto O = reinterpret_cast<__nv_bfloat16*>(o);
  if (hd == 32) var_attn_fwd_kernel<1><<<grid, 256, 0, st>>>(Q, KV, O, lse, rows, Na, V, H, q_shared, scale);
  else if (hd == 64) var_attn_fwd_kernel<2><<<grid, 256, 0, st>>>(Q, KV, O, lse, rows, Na, V, H, q_shared, scale);
  else var_attn_fwd_kernel<4><<<grid, 256, 0, st>>>(Q, KV, O, lse, rows, Na, V, H, q_shared, scale);
  return check_launch("var_attn_fwd_kernel");
}

extern "C" int ucf_var_attention_bwd(const void* q, const void* kv, const void* o, const void* d_o, const float* lse,
                                     void* dkv, float* dq_acc, long long rows, int Na, int V, int H, int hd,
                                     int q_shared, float scale, void* stream) {
  if (rows <= 0 || Na <= 0 || V <= 0 || H <= 0) { set_last_error("var_attention_bwd: empty problem"); return UCF_ERR_BAD_ARG; }
  if (hd != 32 && hd != 64 && hd != 128) { set_last_error("var_attention_bwd: head_dim %d not in {32,64,128}", hd); return UCF_ERR_UNSUPPORTED; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (q_shared) {
    cudaError_t e = cudaMemsetAsync(dq_acc, 0, sizeof(float) * static_cast<size_t>(Na) * H * hd, st);
    if (e != cudaSuccess) { set_last_error("var_attention_bwd: memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  }
  const int grid = va_grid(rows * H);
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q); auto KV = reinterpret_cast<const __nv_bfloat16*>(kv);
  auto O = reinterpret_cast<const __nv_bfloat16*>(o); auto DO = reinterpret_cast<const __nv_bfloat16*>(d_o);
  auto DKV = reinterpret_cast<__nv_bfloat16*>(dkv);
  if (hd == 32) var_attn_bwd_kernel<1><<<grid, 256, 0, st>>>(Q, KV, O, DO, lse, DKV, dq_acc, rows, Na, V, H, q_shared, scale);
  else if (hd == 64) var_attn_bwd_kernel<2><<<grid, 256, 0, st>>>(Q, KV, O, DO, lse, DKV, dq_acc, rows, Na, V, H, q_shared, scale);
  else var_attn_bwd_kernel<4><<<grid, 256, 0, st>>>(Q, KV, O, DO, lse, DKV, dq_acc, rows, Na, V, H, q_shared, scale);
  return check_launch("var_attn_bwd_kernel");
}
