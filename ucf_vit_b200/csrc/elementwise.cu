// Bandwidth-bound helpers around the tensor-core kernels: dtype casts, bias-gradient column sums,
// and the cast+patchify pass that turns Conv{2,3}d(k=s=p) into a plain GEMM.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

__global__ void __launch_bounds__(256)
cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
      uint4 o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
      *reinterpret_cast<uint4*>(dst + i) = o;
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
    }
  }
}

struct CastMultiArgs {
  const float* src[8];
  __nv_bfloat16* dst[8];
  long long n[8];
};
// blockIdx.y selects the tensor; same body as the single-tensor kernel
__global__ void __launch_bounds__(256) cast_f32_to_bf16_multi_kernel(const CastMultiArgs a) {
  const float* __restrict__ src = a.src[blockIdx.y];
  __nv_bfloat16* __restrict__ dst = a.dst[blockIdx.y];
  const long long n = a.n[blockIdx.y];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + i));
      const float4 y = __ldg(reinterpret_cast<const float4*>(src + i + 4));
      *reinterpret_cast<uint4*>(dst + i) =
          make_uint4(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w), pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
    }
  }
}

template <bool ACC>
__global__ void __launch_bounds__(256)
cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(src + i));
      const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
      float4 o0 = make_float4(a.x, a.y, b.x, b.y), o1 = make_float4(c.x, c.y, d.x, d.y);
      if (ACC) {
        const float4 p0 = *reinterpret_cast<const float4*>(dst + i), p1 = *reinterpret_cast<const float4*>(dst + i + 4);
        o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
        o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
      }
      *reinterpret_cast<float4*>(dst + i) = o0;
      *reinterpret_cast<float4*>(dst + i + 4) = o1;
    } else {
      for (long long j = i; j < n; ++j) dst[j] = (ACC ? dst[j] : 0.f) + __bfloat162float(src[j]);
    }
  }
}

// Column sums of a [M, N] bf16 matrix.  Block = 32 column-pairs x 8 row lanes; each block owns a
// 64-column strip and a slab of rows, so every warp reads 128 contiguous bytes per row.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long M, int N,
                   long long ld, int rows_per_block) {
  __shared__ float2 part[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + cx * 2;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = min(r0 + rows_per_block, M);
  float2 acc = make_float2(0.f, 0.f);
  if (col < N) {   // N is even (16-byte pitch contract)
    long long r = r0 + ry;
    for (; r + 24 < r1; r += 32) {
      uint32_t v0 = __ldg(reinterpret_cast<const uint32_t*>(x + r * ld + col));
      uint32_t v1 = __ldg(reinterpret_cast<const uint32_t*>(x + (r + 8) * ld + col));
      uint32_t v2 = __ldg(reinterpret_cast<const uint32_t*>(x + (r + 16) * ld + col));
      uint32_t v3 = __ldg(reinterpret_cast<const uint32_t*>(x + (r + 24) * ld + col));
      float2 a = unpack_bf16x2(v0), b = unpack_bf16x2(v1), c = unpack_bf16x2(v2), d = unpack_bf16x2(v3);
      acc.x += (a.x + b.x) + (c.x + d.x);
      acc.y += (a.y + b.y) + (c.y + d.y);
    }
    for (; r < r1; r += 8) {
      float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(x + r * ld + col)));
      acc.x += a.x; acc.y += a.y;
    }
  }
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < N) {
    float2 s = part[0][cx];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += part[i][cx].x; s.y += part[i][cx].y; }
    atomicAdd(&out[col], s.x);
    if (col + 1 < N) atomicAdd(&out[col + 1], s.y);
  }
}

// x [B, C, G0*p, G1*p(, G2*p)] -> out bf16 [B*G0*G1(*G2), C*p^dims].  One thread produces 8
// consecutive K elements (one 16-byte store); the innermost patch axis (p) is contiguous in x.
// VEC consecutive elements of one patch row per thread (VEC | p; VEC == 8: one 16-byte store, the layout every
// ViT-B/16-style model uses; smaller VEC serves patch sizes 4 / 2 / odd and cropped images whose rows are not
// 16-byte aligned).  S0/S1/S2 are the image's REAL spatial sizes (>= G*p: like the strided convolution, pixels
// past the last whole patch are ignored); ldo >= K is the output row pitch.
// XT: element type of x -- 0 fp32, 1 bf16, 2 uint8 (raw pixels: a quarter of the fp32 host->device bytes)
template <int XT, int VEC>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int G0, int G1,
                int G2, int p, int dims, int S0, int S1, int S2, long long ldo, long long total_vec) {
  const int Kp = (dims == 2) ? p * p : p * p * p;
  const int K = C * Kp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total_vec; t += stride) {
    const long long e = t * VEC;
    const long long row = e / K;
    int k = static_cast<int>(e - row * K);
    const int kcol = k;
    const int c = k / Kp; k -= c * Kp;
    long long src;
    if (dims == 2) {
      const int p0 = k / p, p1 = k - p0 * p;
      const int g1 = static_cast<int>(row % G1);
      const long long r2 = row / G1;
      const int g0 = static_cast<int>(r2 % G0);
      const long long b = r2 / G0;
      src = ((b * C + c) * S0 + (static_cast<long long>(g0) * p + p0)) * S1 + static_cast<long long>(g1) * p + p1;
    } else {
      const int p0 = k / (p * p); k -= p0 * p * p;
      const int p1 = k / p, p2 = k - p1 * p;
      const int g2 = static_cast<int>(row % G2);
      long long r2 = row / G2;
      const int g1 = static_cast<int>(r2 % G1); r2 /= G1;
      const int g0 = static_cast<int>(r2 % G0);
      const long long b = r2 / G0;
      src = (((b * C + c) * S0 + (static_cast<long long>(g0) * p + p0)) * S1 + (static_cast<long long>(g1) * p + p1)) * S2 +
            static_cast<long long>(g2) * p + p2;
    }
    __nv_bfloat16* dst = out + row * ldo + kcol;
    if (VEC == 8) {
      if (XT == 1) {
        *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + src));
      } else if (XT == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + src));
        const float4 b2 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + src + 4));
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b2.x, b2.y),
                                                    pack_bf16x2(b2.z, b2.w));
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(x) + src));
        *reinterpret_cast<uint4*>(dst) = make_uint4(
            pack_bf16x2(static_cast<float>(q.x & 255u), static_cast<float>((q.x >> 8) & 255u)),
            pack_bf16x2(static_cast<float>((q.x >> 16) & 255u), static_cast<float>(q.x >> 24)),
            pack_bf16x2(static_cast<float>(q.y & 255u), static_cast<float>((q.y >> 8) & 255u)),
            pack_bf16x2(static_cast<float>((q.y >> 16) & 255u), static_cast<float>(q.y >> 24)));
      }
    } else if (VEC == 4) {
      if (XT == 1) {
        *reinterpret_cast<uint2*>(dst) = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + src));
      } else if (XT == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + src));
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
      } else {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(x) + src));
        *reinterpret_cast<uint2*>(dst) = make_uint2(
            pack_bf16x2(static_cast<float>(q & 255u), static_cast<float>((q >> 8) & 255u)),
            pack_bf16x2(static_cast<float>((q >> 16) & 255u), static_cast<float>(q >> 24)));
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        dst[i] = XT == 1 ? reinterpret_cast<const __nv_bfloat16*>(x)[src + i]
               : XT == 0 ? __float2bfloat16_rn(reinterpret_cast<const float*>(x)[src + i])
                         : __float2bfloat16_rn(static_cast<float>(reinterpret_cast<const uint8_t*>(x)[src + i]));
    }
  }
}

// x[b, n, :] = (n < P ? prefix[n] : tok[b, n - P, :]) + pos[(pos_bstride ? b : 0), n - pos_off, :]
// (class-token concat + position-embedding add of VIT._pos_embed, one pass instead of cat + add).
// tok bf16 [B, L, D]; prefix/pos of dtype f32 or bf16; out bf16 [B, P + L, D]; 8 elements per thread.
template <bool PRM_BF16>
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok, const void* __restrict__ prefix,
                       const void* __restrict__ pos, __nv_bfloat16* __restrict__ out, int B, int L, int P, int D,
                       long long pos_bstride, int pos_off) {
  const int N = L + P;
  const int dv = D / 8;
  const long long total = static_cast<long long>(B) * N * dv;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int d = static_cast<int>(t % dv) * 8;
    const long long r = t / dv;
    const int n = static_cast<int>(r % N);
    const long long b = r / N;
    float v[8];
    if (n < P) {
      if (PRM_BF16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(prefix) + n * D + d));
        float2 a = unpack_bf16x2(q.x), c = unpack_bf16x2(q.y), e = unpack_bf16x2(q.z), f = unpack_bf16x2(q.w);
        v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y; v[4] = e.x; v[5] = e.y; v[6] = f.x; v[7] = f.y;
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(prefix) + n * D + d));
        const float4 c = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(prefix) + n * D + d + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      }
    } else {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(tok + (b * L + (n - P)) * D + d));
      float2 a = unpack_bf16x2(q.x), c = unpack_bf16x2(q.y), e = unpack_bf16x2(q.z), f = unpack_bf16x2(q.w);
      v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y; v[4] = e.x; v[5] = e.y; v[6] = f.x; v[7] = f.y;
    }
    if (pos != nullptr && n >= pos_off) {
      const long long pi = b * pos_bstride + static_cast<long long>(n - pos_off) * D + d;
      if (PRM_BF16) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(pos) + pi));
        float2 a = unpack_bf16x2(q.x), c = unpack_bf16x2(q.y), e = unpack_bf16x2(q.z), f = unpack_bf16x2(q.w);
        v[0] += a.x; v[1] += a.y; v[2] += c.x; v[3] += c.y; v[4] += e.x; v[5] += e.y; v[6] += f.x; v[7] += f.y;
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pos) + pi));
        const float4 c = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pos) + pi + 4));
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += c.x; v[5] += c.y; v[6] += c.z; v[7] += c.w;
      }
    }
    *reinterpret_cast<uint4*>(out + r * D + d) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// out[i0, i1, i2, :] = x[i0, i1, i2, :] + e[i0*s0 + i1*s1 + i2*s2 + :]   (stride 0 = broadcast along that axis)
template <bool PRM_BF16>
__global__ void __launch_bounds__(256)
add_bcast_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ e, __nv_bfloat16* __restrict__ out,
                 int n1, int n2, int D, long long s0, long long s1, long long s2, long long nvec) {
  const int dv = D >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < nvec; t += stride) {
    const long long r = t / dv;
    const int d = static_cast<int>(t - r * dv) << 3;
    const long long q = r / n2;
    const int i2 = static_cast<int>(r - q * n2);
    const long long i0 = q / n1;
    const int i1 = static_cast<int>(q - i0 * n1);
    const long long eo = i0 * s0 + i1 * s1 + i2 * s2 + d;
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + r * D + d));
    float2 a = unpack_bf16x2(xv.x), b = unpack_bf16x2(xv.y), c = unpack_bf16x2(xv.z), f = unpack_bf16x2(xv.w);
    float v[8] = {a.x, a.y, b.x, b.y, c.x, c.y, f.x, f.y};
    if (PRM_BF16) {
      const uint4 ev = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e) + eo));
      a = unpack_bf16x2(ev.x); b = unpack_bf16x2(ev.y); c = unpack_bf16x2(ev.z); f = unpack_bf16x2(ev.w);
      v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += f.x; v[7] += f.y;
    } else {
      const float4 e0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e) + eo));
      const float4 e1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e) + eo + 4));
      v[0] += e0.x; v[1] += e0.y; v[2] += e0.z; v[3] += e0.w; v[4] += e1.x; v[5] += e1.y; v[6] += e1.z; v[7] += e1.w;
    }
    *reinterpret_cast<uint4*>(out + r * D + d) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

static int ew_grid(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return UCF_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_last_error("cast_f32_to_bf16: pointers must be 16-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  cast_f32_to_bf16_kernel<<<ew_grid((n + 7) / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_f32_to_bf16_kernel");
}

extern "C" int ucf_cast_f32_to_bf16_multi(int n, const float* const* srcs, void* const* dsts, const long long* counts,
                                          void* stream) {
  if (n <= 0) return UCF_OK;
  if (n > 8 || !srcs || !dsts || !counts) { set_last_error("cast_f32_to_bf16_multi: 1..8 tensors, non-null arrays"); return UCF_ERR_BAD_ARG; }
  CastMultiArgs a;
  long long nmax = 0;
  for (int i = 0; i < 8; ++i) {
    const int j = i < n ? i : 0;
    a.src[i] = srcs[j]; a.dst[i] = static_cast<__nv_bfloat16*>(dsts[j]); a.n[i] = i < n ? counts[j] : 0;
    if (i < n) {
      if (!srcs[i] || !dsts[i] || (reinterpret_cast<uintptr_t>(srcs[i]) & 15) || (reinterpret_cast<uintptr_t>(dsts[i]) & 15)) {
        set_last_error("cast_f32_to_bf16_multi: pointer %d null or not 16-byte aligned", i); return UCF_ERR_BAD_ARG;
      }
      if (counts[i] > nmax) nmax = counts[i];
    }
  }
  if (nmax <= 0) return UCF_OK;
  int gx = ew_grid((nmax + 7) / 8, 256);
  const int cap = (num_sms() * 16 + n - 1) / n;          // ~16 CTAs per SM over all tensors
  if (gx > cap) gx = cap;
  cast_f32_to_bf16_multi_kernel<<<dim3(gx, n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("cast_f32_to_bf16_multi_kernel");
}

extern "C" int ucf_cast_bf16_to_f32(const void* src, float* dst, long long n, int accumulate, void* stream) {
  if (n <= 0) return UCF_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_last_error("cast_bf16_to_f32: pointers must be 16-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid((n + 7) / 8, 256);
  if (accumulate) cast_bf16_to_f32_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  else cast_bf16_to_f32_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  return check_launch("cast_bf16_to_f32_kernel");
}

extern "C" int ucf_colsum_bf16(const void* x, float* out, long long M, int N, long long ld, int accumulate,
                               void* stream) {
  if (M <= 0 || N <= 0) return UCF_OK;
  if (N % 2 || ld % 2 || (reinterpret_cast<uintptr_t>(x) & 3)) {
    set_last_error("colsum_bf16: N and ld must be even, x 4-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, st);
    if (e != cudaSuccess) { set_last_error("colsum_bf16: memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  }
  const int strips = (N + 63) / 64;
  // aim for ~8 blocks per SM overall
  long long want_y = (static_cast<long long>(num_sms()) * 8 + strips - 1) / strips;
  long long rows_per_block = (M + want_y - 1) / want_y;
  if (rows_per_block < 64) rows_per_block = 64;
  const long long gy = (M + rows_per_block - 1) / rows_per_block;
  dim3 grid(strips, static_cast<unsigned>(gy));
  colsum_bf16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, M, N, ld,
                                            static_cast<int>(rows_per_block));
  return check_launch("colsum_bf16_kernel");
}

extern "C" int ucf_patchify(const void* x, void* out, int B, int C, int G0, int G1, int G2, int p, int dims,
                            int S0, int S1, int S2, long long ld_out, int x_dtype, void* stream) {
  if (dims != 2 && dims != 3) { set_last_error("patchify: dims must be 2 or 3"); return UCF_ERR_BAD_ARG; }
  if (p <= 0 || B < 0 || C <= 0 || G0 < 0 || G1 < 0 || G2 < 0) { set_last_error("patchify: bad shape"); return UCF_ERR_BAD_ARG; }
  if (dims == 2) { G2 = 1; S2 = 1; }
  if (S0 < G0 * p || S1 < G1 * p || (dims == 3 && S2 < G2 * p)) {
    set_last_error("patchify: image %dx%dx%d is smaller than grid x patch", S0, S1, S2); return UCF_ERR_BAD_ARG;
  }
  const long long Kp = (dims == 2) ? 1LL * p * p : 1LL * p * p * p;
  const long long K = C * Kp;
  if (ld_out < K || ld_out % 8) { set_last_error("patchify: ld_out must be >= K and a multiple of 8"); return UCF_ERR_BAD_ARG; }
  const long long rows = 1LL * B * G0 * G1 * G2;
  const long long total = rows * K;
  if (total <= 0) return UCF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ld_out > K) {      // zero the pad columns (the GEMM's K extent is ld_out)
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(rows) * ld_out, st);
    if (e != cudaSuccess) { set_last_error("patchify: memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  }
  // widest vector that divides the patch row and keeps every source / destination address aligned
  if (x_dtype != UCF_DTYPE_F32 && x_dtype != UCF_DTYPE_BF16 && x_dtype != UCF_DTYPE_U8) {
    set_last_error("patchify: x_dtype must be f32, bf16 or u8"); return UCF_ERR_BAD_ARG;
  }
  const int esz = x_dtype == UCF_DTYPE_BF16 ? 2 : x_dtype == UCF_DTYPE_U8 ? 1 : 4;
  const long long inner = dims == 2 ? S1 : S2;
  int vec = 1;
  for (int v = 8; v >= 4; v >>= 1) {
    const uintptr_t align = v * esz > 16 ? 16 : v * esz;          // widest single load the kernel issues
    if (p % v == 0 && inner % v == 0 && reinterpret_cast<uintptr_t>(x) % align == 0) { vec = v; break; }
  }
  if ((reinterpret_cast<uintptr_t>(out) & 15)) { set_last_error("patchify: out must be 16-byte aligned"); return UCF_ERR_BAD_ARG; }
  const long long nvec = total / vec;
  const int grid = ew_grid(nvec, 256);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define UCF_PATCHIFY(T_, V_) patchify_kernel<T_, V_><<<grid, 256, 0, st>>>(x, o, B, C, G0, G1, G2, p, dims, S0, S1, S2, ld_out, nvec)
#define UCF_PATCHIFY_T(V_)                                                  \
  {                                                                         \
    if (x_dtype == UCF_DTYPE_BF16) UCF_PATCHIFY(1, V_);                     \
    else if (x_dtype == UCF_DTYPE_U8) UCF_PATCHIFY(2, V_);                  \
    else UCF_PATCHIFY(0, V_);                                               \
  }
  if (vec == 8) UCF_PATCHIFY_T(8)
  else if (vec == 4) UCF_PATCHIFY_T(4)
  else UCF_PATCHIFY_T(1)
#undef UCF_PATCHIFY_T
#undef UCF_PATCHIFY
  return check_launch("patchify_kernel");
}

extern "C" int ucf_assemble_tokens(const void* tok, const void* prefix, const void* pos, void* out, int B, int L, int P,
                                   int D, long long pos_bstride, int pos_off, int param_dtype, void* stream) {
  if (B <= 0 || L < 0 || P < 0 || D <= 0) { set_last_error("assemble_tokens: bad shape"); return UCF_ERR_BAD_ARG; }
  if (D % 8) { set_last_error("assemble_tokens: D=%d must be a multiple of 8", D); return UCF_ERR_BAD_ARG; }
  if (P > 0 && !prefix) { set_last_error("assemble_tokens: prefix tokens requested but prefix is NULL"); return UCF_ERR_BAD_ARG; }
  const long long nvec = static_cast<long long>(B) * (L + P) * (D / 8);
  if (nvec == 0) return UCF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid(nvec, 256);
  if (param_dtype == UCF_DTYPE_BF16)
    assemble_tokens_kernel<true><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(tok), prefix, pos,
                                                       reinterpret_cast<__nv_bfloat16*>(out), B, L, P, D, pos_bstride, pos_off);
  else
    assemble_tokens_kernel<false><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(tok), prefix, pos,
                                                        reinterpret_cast<__nv_bfloat16*>(out), B, L, P, D, pos_bstride, pos_off);
  return check_launch("assemble_tokens_kernel");
}

extern "C" int ucf_add_bcast(const void* x, const void* e, void* out, long long n0, int n1, int n2, int D, long long s0,
                             long long s1, long long s2, int param_dtype, void* stream) {
  if (n0 < 0 || n1 <= 0 || n2 <= 0 || D <= 0) { set_last_error("add_bcast: bad shape"); return UCF_ERR_BAD_ARG; }
  if (D % 8 || s0 % 8 || s1 % 8 || s2 % 8 || s0 < 0 || s1 < 0 || s2 < 0) {
    set_last_error("add_bcast: D and the strides of e must be non-negative multiples of 8 elements"); return UCF_ERR_BAD_ARG;
  }
  if (param_dtype != UCF_DTYPE_F32 && param_dtype != UCF_DTYPE_BF16) {
    set_last_error("add_bcast: e must be f32 or bf16"); return UCF_ERR_BAD_ARG;
  }
  const long long nvec = n0 * n1 * n2 * (D / 8);
  if (nvec == 0) return UCF_OK;
  if (!x || !e || !out || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    set_last_error("add_bcast: pointers must be non-null and 16-byte aligned"); return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid(nvec, 256);
  if (param_dtype == UCF_DTYPE_BF16)
    add_bcast_kernel<true><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), e, static_cast<__nv_bfloat16*>(out),
                                                 n1, n2, D, s0, s1, s2, nvec);
  else
    add_bcast_kernel<false><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), e, static_cast<__nv_bfloat16*>(out),
                                                  n1, n2, D, s0, s1, s2, nvec);
  return check_launch("add_bcast_kernel");
}
