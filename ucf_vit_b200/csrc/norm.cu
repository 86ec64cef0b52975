// LayerNorm forward / backward: HBM-bandwidth-bound, one warp per row, 16-byte vector access,
// warp-shuffle reductions, row kept in registers between the statistics and the normalise pass
// (algorithmic traffic: fwd 1 read + 1 write, bwd 3 reads (dy, x, dres) + 1 write).
// Replaces nn.LayerNorm(D, eps=1e-6) of Block.norm1/norm2 and VIT.norm
// (/root/reference/src/UCF_VIT/simple/arch.py:170,266; building_blocks.py:212,226).
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// ---- packed fp32x2 helpers (Blackwell FFMA2 / FADD2 / FMUL2: two fp32 lanes per instruction) -------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {   // bf16x2 -> fp32x2: one shift + one mask
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ void unpack4(const uint4& r, float2 (&f)[4]) {
  f[0] = bf2_to_f2(r.x); f[1] = bf2_to_f2(r.y); f[2] = bf2_to_f2(r.z); f[3] = bf2_to_f2(r.w);
}
template <bool BF16>
__device__ __forceinline__ void load4x2(const void* base, long long idx, float2 (&f)[4]) {   // 8 elements
  if (BF16) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    unpack4(r, f);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4));
    f[0] = f2(a.x, a.y); f[1] = f2(a.z, a.w); f[2] = f2(b.x, b.y); f[3] = f2(b.z, b.w);
  }
}
__device__ __forceinline__ uint4 pack4(const float2 (&f)[4]) {
  return make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y),
                    pack_bf16x2(f[3].x, f[3].y));
}

// NCH = number of 256-element chunks a warp covers per row (D <= 256*NCH)
template <int NCH, bool X_BF16, bool P_BF16>
__global__ void __launch_bounds__(256, (NCH <= 3 ? 2 : 1))
layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ gamma, const void* __restrict__ beta,
                     void* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);

  // gamma / beta of this lane's columns live in registers for the whole (persistent) kernel when the row
  // is short enough: re-loading them per row put a long-scoreboard stall in front of every write pass
  constexpr bool HOIST = NCH <= 3;
  float2 gr[HOIST ? NCH : 1][4], br[HOIST ? NCH : 1][4];
  if (HOIST) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
#pragma unroll
      for (int e = 0; e < 4; ++e) { gr[c][e] = f2(1.f, 1.f); br[c][e] = f2(0.f, 0.f); }
      if (col < D) {
        if (gamma) load4x2<P_BF16>(gamma, col, gr[c]);
        if (beta) load4x2<P_BF16>(beta, col, br[c]);
      }
    }
  }
  // bf16 rows are software-prefetched one row ahead (raw 16-byte vectors): a warp otherwise has no load in
  // flight while it reduces and writes, and 24 resident warps x 1.5 KB is short of the bytes-in-flight that
  // HBM3e latency x bandwidth asks for
  uint4 nxt[NCH];
  if (X_BF16 && warp_global < rows) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (c * 256 + lane * 8 < D)
        nxt[c] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + warp_global * D + lane * 8 + c * 256));
  }
  for (long long row = warp_global; row < rows; row += nwarps) {
    const long long base = row * D + lane * 8;
    float2 v[NCH][4];
    float2 s = f2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c * 256 + lane * 8 < D) {
        if (X_BF16) unpack4(nxt[c], v[c]);
        else load4x2<X_BF16>(x, base + c * 256, v[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) s = __fadd2_rn(s, v[c][e]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[c][e] = f2(0.f, 0.f);
      }
    }
    if (X_BF16 && row + nwarps < rows) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c * 256 + lane * 8 < D)
          nxt[c] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + (row + nwarps) * D + lane * 8 + c * 256));
    }
    const float mu = warp_sum(s.x + s.y) * inv_d;
    const float2 nmu = f2(-mu, -mu);
    float2 ss = f2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c * 256 + lane * 8 < D) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[c][e] = __fadd2_rn(v[c][e], nmu);          // centred values are reused by the write pass
          ss = __ffma2_rn(v[c][e], v[c][e], ss);
        }
      }
    }
    const float rs = rsqrtf(warp_sum(ss.x + ss.y) * inv_d + eps);
    const float2 rs2 = f2(rs, rs);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
    if (y == nullptr) continue;          // statistics-only pass (ucf_layernorm_stats: the LayerNorm -> GEMM fusion)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float2 g[4], b[4], o[4];
        if (!HOIST) {
          if (gamma) load4x2<P_BF16>(gamma, col, g);
          if (beta) load4x2<P_BF16>(beta, col, b);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 t = __fmul2_rn(v[c][e], rs2);
          if (HOIST) t = __ffma2_rn(t, gr[c][e], br[c][e]);           // (gamma = 1 / beta = 0 when absent)
          else if (gamma) t = beta ? __ffma2_rn(t, g[e], b[e]) : __fmul2_rn(t, g[e]);
          else if (beta) t = __fadd2_rn(t, b[e]);
          o[e] = t;
        }
        *reinterpret_cast<uint4*>(yp + base + c * 256) = pack4(o);
      }
    }
  }
}

// Backward.  Each warp walks rows with a grid stride, keeps per-lane partial dgamma/dbeta in
// registers, then the block reduces them through shared memory and issues one atomicAdd per column.
// Issue-bound before memory-bound (ncu: IPC 2.1, 52 % issue slots at 40 % DRAM), so the math is
// arranged for the fewest instructions, on packed fp32x2 lanes:
//   t = dy*g,  s1 = sum t,  s2' = sum t*x            (s2 = rs*(s2' - mu*s1))
//   dgamma += rs*(dy*x) - (rs*mu)*dy,  dbeta += dy
//   dx = rs*t + B*x + C   with  B = -rs^2*s2/D,  C = -B*mu - rs*s1/D      (+ dres)
// The row (dy, x) is held as the RAW 16-byte bf16 vectors between the two passes.
template <int NCH, bool P_BF16>
__global__ void __launch_bounds__(256, (NCH <= 3 ? 2 : 1))
layernorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, const void* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const void* __restrict__ dres, void* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, long long rows, int D) {
  extern __shared__ float red[];   // [2][D] dgamma / dbeta partials, then [D] gamma as fp32, then the row prefetch buffers
  float* gam_s = red + 2 * D;
  // Row prefetch (NCH <= 4): every warp owns two buffers of {dy, x, dres} rows in shared memory and fills the
  // next row's with cp.async while it works on the current one.  Without it a warp had no load in flight while
  // it computed, and ncu charged 29 % of all samples to the first use of a freshly loaded row; prefetching into
  // registers is not an option at 126 registers and two CTAs per SM.
  constexpr bool PREFETCH = NCH <= 4;
  constexpr int ROWB = NCH * 512;                                    // bytes of one bf16 row slot
  uint8_t* pre = reinterpret_cast<uint8_t*>(red + 3 * D) + (threadIdx.x >> 5) * (2 * 3 * ROWB);
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    gam_s[i] = gamma ? (P_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gamma)[i])
                               : reinterpret_cast<const float*>(gamma)[i])
                     : 1.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(dres);
  __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx);

  float2 dg[NCH][4], db[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[c][e] = f2(0.f, 0.f); db[c][e] = f2(0.f, 0.f); }
  }
  // gamma is re-read per row from shared memory (two LDS.128) instead of pinning 8*NCH registers; the
  // global (L1-hit) loads used before stalled both passes on the long scoreboard
  auto load_gamma = [&](int col, float2 (&gv)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(gam_s + col);
    const float4 b = *reinterpret_cast<const float4*>(gam_s + col + 4);
    gv[0] = f2(a.x, a.y); gv[1] = f2(a.z, a.w); gv[2] = f2(b.x, b.y); gv[3] = f2(b.z, b.w);
  };

  auto prefetch_row = [&](long long row, int buf) {
    uint8_t* b = pre + buf * 3 * ROWB + lane * 16;
    const long long base = row * D + lane * 8;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c * 256 + lane * 8 < D) {
        cp_async16(b + c * 512, dyp + base + c * 256);
        cp_async16(b + ROWB + c * 512, xp + base + c * 256);
        if (dres) cp_async16(b + 2 * ROWB + c * 512, rp + base + c * 256);
      }
    }
  };
  float mu_n = 0.f, rs_n = 0.f;
  if (PREFETCH && warp_global < rows) {
    prefetch_row(warp_global, 0);
    mu_n = __ldg(mean + warp_global);
    rs_n = __ldg(rstd + warp_global);
  }
  if (PREFETCH) cp_async_commit();
  int buf = 0;
  for (long long row = warp_global; row < rows; row += nwarps, buf ^= 1) {
    const long long base = row * D + lane * 8;
    uint4 dyr[NCH], xr[NCH], rr[NCH];
    float mu, rs;
    if (PREFETCH) {
      mu = mu_n; rs = rs_n;
      __syncwarp();                                  // every lane is done reading the buffer about to be refilled
      if (row + nwarps < rows) {
        prefetch_row(row + nwarps, buf ^ 1);
        mu_n = __ldg(mean + row + nwarps);
        rs_n = __ldg(rstd + row + nwarps);
      }
      cp_async_commit();
      cp_async_wait<1>();                            // this row's group has landed (each lane reads only its own copies)
      const uint8_t* b = pre + buf * 3 * ROWB + lane * 16;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * 256 + lane * 8 < D) {
          dyr[c] = *reinterpret_cast<const uint4*>(b + c * 512);
          xr[c] = *reinterpret_cast<const uint4*>(b + ROWB + c * 512);
          if (dres) rr[c] = *reinterpret_cast<const uint4*>(b + 2 * ROWB + c * 512);
        } else {
          dyr[c] = make_uint4(0, 0, 0, 0);
          xr[c] = make_uint4(0, 0, 0, 0);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * 256 + lane * 8 < D) {
          dyr[c] = __ldg(reinterpret_cast<const uint4*>(dyp + base + c * 256));
          xr[c] = __ldg(reinterpret_cast<const uint4*>(xp + base + c * 256));
          if (dres) rr[c] = __ldg(reinterpret_cast<const uint4*>(rp + base + c * 256));
        } else {
          dyr[c] = make_uint4(0, 0, 0, 0);
          xr[c] = make_uint4(0, 0, 0, 0);
        }
      }
      mu = mean[row]; rs = rstd[row];
    }
    const float2 rs2 = f2(rs, rs), nrm2 = f2(-rs * mu, -rs * mu);
    float2 s1 = f2(0.f, 0.f), s2 = f2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float2 dyv[4], xv[4], gv[4];
        unpack4(dyr[c], dyv);
        unpack4(xr[c], xv);
        load_gamma(col, gv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __fmul2_rn(dyv[e], gv[e]);
          s1 = __fadd2_rn(s1, t);
          s2 = __ffma2_rn(t, xv[e], s2);
          const float2 u = __fmul2_rn(dyv[e], xv[e]);
          dg[c][e] = __ffma2_rn(u, rs2, dg[c][e]);
          dg[c][e] = __ffma2_rn(dyv[e], nrm2, dg[c][e]);
          db[c][e] = __fadd2_rn(db[c][e], dyv[e]);
        }
      }
    }
    const float S1 = warp_sum(s1.x + s1.y);
    const float S2p = warp_sum(s2.x + s2.y);
    const float S2 = rs * (S2p - mu * S1);           // sum t * xhat
    const float Bc = -rs * rs * S2 * inv_d;
    const float Cc = -Bc * mu - rs * S1 * inv_d;
    const float2 B2 = f2(Bc, Bc), C2 = f2(Cc, Cc);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < D) {
        float2 dyv[4], xv[4], gv[4], o[4];
        unpack4(dyr[c], dyv);
        unpack4(xr[c], xv);
        load_gamma(col, gv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __fmul2_rn(dyv[e], gv[e]);
          o[e] = __ffma2_rn(t, rs2, __ffma2_rn(xv[e], B2, C2));
        }
        if (dres) {
          float2 r[4];
          unpack4(rr[c], r);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = __fadd2_rn(o[e], r[e]);
        }
        *reinterpret_cast<uint4*>(dxp + base + c * 256) = pack4(o);
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
      for (int e = 0; e < 4; ++e) {
        atomicAdd(&red_g[col + 2 * e], dg[c][e].x);
        atomicAdd(&red_g[col + 2 * e + 1], dg[c][e].y);
        atomicAdd(&red_b[col + 2 * e], db[c][e].x);
        atomicAdd(&red_b[col + 2 * e + 1], db[c][e].y);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], red_g[i]);
    if (dbeta) atomicAdd(&dbeta[i], red_b[i]);
  }
}

// Wide rows (2048 < D <= 8192: the adaptive token_embeds' LayerNorm over patch_dim = C*p^d, 4096 for 3-D p = 16,
// arch.py:282-289).  Same algebra as above, but the row no longer fits in registers: a warp walks its row in 256-column
// steps twice (the second pass hits L1/L2) and the per-column dgamma / dbeta partials live in shared memory
// (shared atomics: different warps of the CTA hit the same column).  One LayerNorm per step on this path.
template <bool P_BF16>
__global__ void __launch_bounds__(256)
layernorm_bwd_wide_kernel(const void* __restrict__ dy, const void* __restrict__ x, const void* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ rstd,
                          const void* __restrict__ dres, void* __restrict__ dx, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, long long rows, int D) {
  extern __shared__ float red[];   // [2][D] dgamma / dbeta partials, then [D] gamma as fp32
  float* red_g = red;
  float* red_b = red + D;
  float* gam_s = red + 2 * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    red_g[i] = 0.f; red_b[i] = 0.f;
    gam_s[i] = gamma ? (P_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gamma)[i])
                               : reinterpret_cast<const float*>(gamma)[i])
                     : 1.0f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(dres);
  __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
  const bool want_pg = dgamma != nullptr || dbeta != nullptr;
  for (long long row = warp_global; row < rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    const long long base = row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int col = lane * 8; col < D; col += 256) {
      float2 dyv[4], xv[4];
      unpack4(__ldg(reinterpret_cast<const uint4*>(dyp + base + col)), dyv);
      unpack4(__ldg(reinterpret_cast<const uint4*>(xp + base + col)), xv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g0 = gam_s[col + 2 * e], g1 = gam_s[col + 2 * e + 1];
        const float t0 = dyv[e].x * g0, t1 = dyv[e].y * g1;
        s1 += t0 + t1;
        s2 = fmaf(t0, xv[e].x, fmaf(t1, xv[e].y, s2));
        if (want_pg) {
          atomicAdd(&red_g[col + 2 * e], rs * (dyv[e].x * xv[e].x) - rs * mu * dyv[e].x);
          atomicAdd(&red_g[col + 2 * e + 1], rs * (dyv[e].y * xv[e].y) - rs * mu * dyv[e].y);
          atomicAdd(&red_b[col + 2 * e], dyv[e].x);
          atomicAdd(&red_b[col + 2 * e + 1], dyv[e].y);
        }
      }
    }
    const float S1 = warp_sum(s1);
    const float S2 = rs * (warp_sum(s2) - mu * S1);
    const float Bc = -rs * rs * S2 * inv_d;
    const float Cc = -Bc * mu - rs * S1 * inv_d;
    for (int col = lane * 8; col < D; col += 256) {
      float2 dyv[4], xv[4], o[4];
      unpack4(__ldg(reinterpret_cast<const uint4*>(dyp + base + col)), dyv);
      unpack4(__ldg(reinterpret_cast<const uint4*>(xp + base + col)), xv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e].x = fmaf(dyv[e].x * gam_s[col + 2 * e], rs, fmaf(xv[e].x, Bc, Cc));
        o[e].y = fmaf(dyv[e].y * gam_s[col + 2 * e + 1], rs, fmaf(xv[e].y, Bc, Cc));
      }
      if (dres) {
        float2 r[4];
        unpack4(__ldg(reinterpret_cast<const uint4*>(rp + base + col)), r);
#pragma unroll
        for (int e = 0; e < 4; ++e) { o[e].x += r[e].x; o[e].y += r[e].y; }
      }
      *reinterpret_cast<uint4*>(dxp + base + col) = pack4(o);
    }
  }
  if (!want_pg) return;
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], red_g[i]);
    if (dbeta) atomicAdd(&dbeta[i], red_b[i]);
  }
}

// Statistics only (bf16 rows): one pass, sum and sum of squares together, two rows in flight per warp.  8 bytes out
// per row; the whole kernel is one read of x.  var = E[x^2] - mean^2 in fp32 over bf16 data (|x| of activations is O(1..100)).
__global__ void __launch_bounds__(256)
layernorm_stats_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                            long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.0f / static_cast<float>(D);
  const int nvec = D >> 3;                                   // 16-byte vectors per row
  for (long long row = warp_global * 2; row < rows; row += nwarps * 2) {
    const bool two = row + 1 < rows;
    const uint4* r0 = reinterpret_cast<const uint4*>(x + row * D);
    const uint4* r1 = reinterpret_cast<const uint4*>(x + (row + (two ? 1 : 0)) * D);
    float2 s0 = f2(0.f, 0.f), q0 = f2(0.f, 0.f), s1 = f2(0.f, 0.f), q1 = f2(0.f, 0.f);
    for (int v = lane; v < nvec; v += 32) {
      const uint4 a = __ldg(r0 + v), b = __ldg(r1 + v);
      float2 fa[4], fb[4];
      unpack4(a, fa);
      unpack4(b, fb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s0 = __fadd2_rn(s0, fa[e]); q0 = __ffma2_rn(fa[e], fa[e], q0);
        s1 = __fadd2_rn(s1, fb[e]); q1 = __ffma2_rn(fb[e], fb[e], q1);
      }
    }
    float a0 = s0.x + s0.y, b0 = q0.x + q0.y, a1 = s1.x + s1.y, b1 = q1.x + q1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); b0 += __shfl_xor_sync(0xffffffffu, b0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    }
    if (lane == 0) {
      const float m0 = a0 * inv_d;
      mean_out[row] = m0;
      rstd_out[row] = rsqrtf(fmaxf(b0 * inv_d - m0 * m0, 0.f) + eps);
      if (two) {
        const float m1 = a1 * inv_d;
        mean_out[row + 1] = m1;
        rstd_out[row + 1] = rsqrtf(fmaxf(b1 * inv_d - m1 * m1, 0.f) + eps);
      }
    }
  }
}

static int ln_grid(long long rows) {
  const long long warps_per_block = 8;
  long long blocks = (rows + warps_per_block - 1) / warps_per_block;
  const long long cap = static_cast<long long>(num_sms()) * 2;   // persistent: the 2 resident CTAs per SM, no tail wave
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

extern "C" int ucf_layernorm_stats(const void* x, float* mean, float* rstd, long long rows, int D, float eps, int x_dtype,
                                   void* stream) {
  if (rows <= 0) return UCF_OK;
  if (D <= 0 || D % 8 != 0 || D > 4096) {
    set_last_error("layernorm_stats: D=%d must be a multiple of 8 and <= 4096", D);
    return UCF_ERR_BAD_ARG;
  }
  if (!x || !mean || !rstd) { set_last_error("layernorm_stats: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool xb = x_dtype == UCF_DTYPE_BF16;
  if (xb && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    long long blocks = (rows + 15) / 16;                        // 8 warps x 2 rows per block
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    layernorm_stats_bf16_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), mean, rstd, rows, D, eps);
    return check_launch("layernorm_stats_bf16_kernel");
  }
  const int grid = ln_grid(rows);
  const void* none = nullptr;
  void* no_y = nullptr;
#define LAUNCH(NCH)                                                                                              \
  if (xb) layernorm_fwd_kernel<NCH, true, false><<<grid, 256, 0, st>>>(x, none, none, no_y, mean, rstd, rows, D, eps);  \
  else layernorm_fwd_kernel<NCH, false, false><<<grid, 256, 0, st>>>(x, none, none, no_y, mean, rstd, rows, D, eps);
  UCF_LN_DISPATCH_NCH(D, LAUNCH)
#undef LAUNCH
  return check_launch("layernorm_fwd_kernel(stats)");
}

extern "C" int ucf_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 long long rows, int D, int param_dtype, void* stream) {
  if (rows <= 0) return UCF_OK;
  if (D <= 0 || D % 8 != 0 || D > 8192) {
    set_last_error("layernorm_bwd: D=%d must be a multiple of 8 and <= 8192", D);
    return UCF_ERR_BAD_ARG;
  }
  if (!dy || !x || !mean || !rstd || !dx) { set_last_error("layernorm_bwd: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (D > 2048) {
    long long wblocks = (rows + 7) / 8;
    const long long wcap = static_cast<long long>(ucf::num_sms()) * 2;
    if (wblocks > wcap) wblocks = wcap;
    const size_t wsmem = 3 * static_cast<size_t>(D) * sizeof(float);
    if (param_dtype == UCF_DTYPE_BF16) {
      cudaFuncSetAttribute(layernorm_bwd_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wsmem));
      layernorm_bwd_wide_kernel<true><<<static_cast<int>(wblocks), 256, wsmem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D);
    } else {
      cudaFuncSetAttribute(layernorm_bwd_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wsmem));
      layernorm_bwd_wide_kernel<false><<<static_cast<int>(wblocks), 256, wsmem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D);
    }
    return check_launch("layernorm_bwd_wide_kernel");
  }
  // fewer, fatter blocks than forward: every block ends with 2*D global atomics
  long long blocks = (rows + 7) / 8;
  const long long cap = static_cast<long long>(ucf::num_sms()) * (D <= 768 ? 2 : 1);   // resident CTAs per SM, persistent
  if (blocks > cap) blocks = cap;
  const int grid = static_cast<int>(blocks < 1 ? 1 : blocks);
  const int nch = D <= 256 ? 1 : D <= 512 ? 2 : D <= 768 ? 3 : D <= 1024 ? 4 : 0;      // row prefetch buffers: 8 warps x 2 x 3 rows
  const size_t smem = 3 * static_cast<size_t>(D) * sizeof(float) + static_cast<size_t>(nch) * 512 * 3 * 2 * 8;
  const bool pb = param_dtype == UCF_DTYPE_BF16;
  // (more than 48 KB of dynamic shared memory with the prefetch buffers: opt in; the call is cheap and idempotent)
#define LAUNCH(NCH)                                                                                      \
  if (NCH <= 8) {                                                                                        \
    if (pb) {                                                                                            \
      cudaFuncSetAttribute(layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
      layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), true><<<grid, 256, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D); \
    } else {                                                                                             \
      cudaFuncSetAttribute(layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
      layernorm_bwd_kernel<(NCH <= 8 ? NCH : 8), false><<<grid, 256, smem, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D);   \
    }                                                                                                    \
  }
  UCF_LN_DISPATCH_NCH(D, LAUNCH)
#undef LAUNCH
  return check_launch("layernorm_bwd_kernel");
}
