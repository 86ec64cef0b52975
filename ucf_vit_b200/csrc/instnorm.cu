// InstanceNorm (+ residual) (+ LeakyReLU) of the UNETR convolutional decoder on channels-last bf16 tensors:
// HBM-bandwidth-bound, 16-byte vector access, fixed-order (reproducible) two-stage reductions.
//
// Replaces the MONAI block bodies the reference builds its decoder from
// (/root/reference/src/UCF_VIT/simple/arch.py:808-940 -> monai UnetResBlock / UnetBasicBlock):
//     out = lrelu(norm1(conv1(x)));  out = norm2(conv2(out));  res = norm3(conv3(x)) | x;  out = lrelu(out + res)
// with norm = nn.InstanceNorm{2,3}d(C) (affine=False, eps=1e-5, biased variance) and lrelu = LeakyReLU(0.01).
// PyTorch runs each InstanceNorm as three batch-norm kernels plus layout copies and the activation / add as separate
// passes (60 % of the bf16 decoder step, profiles/r02_unetr_kernel_profile_*.log); here one unit
//     y = lrelu( IN(a) [+ IN(b) | + b] )
// costs: forward  1 read of a (statistics) + 1 read of a (+ b) + 1 write of y;
//        backward 1 read of dy, y, a (+ b) (sums) + 1 read of the same + 1 write of da (+ db).
//
// Layout: x[n, s, c], c fastest ("NDHWC" / torch.channels_last{,_3d}), S = D*H*W positions, C channels.
// A CTA of 256 threads covers `vecs = C / VEC` channel vectors x `lanes = 256 / vecs` rows per sweep, so consecutive threads
// read consecutive 16-byte vectors (one fully coalesced 4 KB line per sweep) and a thread keeps the same VEC channels for
// its whole life: the per-(n, c) scalars it needs live in registers.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

constexpr int IN_THREADS = 256;

template <int VEC> struct InVec;
template <> struct InVec<8> { using T = uint4; };
template <> struct InVec<4> { using T = uint2; };
template <> struct InVec<2> { using T = uint32_t; };

template <int VEC>
__device__ __forceinline__ void in_unpack(const typename InVec<VEC>::T& r, float (&f)[VEC]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < VEC / 2; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <int VEC>
__device__ __forceinline__ typename InVec<VEC>::T in_pack(const float (&f)[VEC]) {
  typename InVec<VEC>::T r;
  uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < VEC / 2; ++i) w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return r;
}
template <int VEC>
__device__ __forceinline__ typename InVec<VEC>::T in_ld(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const typename InVec<VEC>::T*>(p));
}
template <int VEC>
__device__ __forceinline__ void in_ldf(const float* p, float (&f)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = __ldg(p + i);
}

struct InGeom {
  int vecs, lanes, rows_per_cta, chunks;
};
static InGeom in_geom(int N, long long S, int C, int VEC) {
  InGeom g;
  g.vecs = C / VEC;
  g.lanes = IN_THREADS / g.vecs;
  int iters = 32;
  auto nchunks = [&](int it) { return (S + static_cast<long long>(g.lanes) * it - 1) / (static_cast<long long>(g.lanes) * it); };
  while (iters > 4 && static_cast<long long>(N) * nchunks(iters) < 4LL * num_sms()) iters /= 2;
  g.rows_per_cta = g.lanes * iters;
  g.chunks = static_cast<int>(nchunks(iters));
  return g;
}
static int in_vec(int C, const void* const* ptrs, int nptr) {
  int v = (C % 8 == 0) ? 8 : (C % 4 == 0) ? 4 : (C % 2 == 0) ? 2 : 0;
  for (int i = 0; i < nptr; ++i) {
    if (!ptrs[i]) continue;
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptrs[i]);
    while (v > 2 && (a % (2 * v)) != 0) v /= 2;
    if (a % 4 != 0) return 0;
  }
  if (v && C / v > IN_THREADS) return 0;
  return v;
}

// CTA-wide sum over the row lanes of K x VEC per-thread partials, fixed order; result written to
// dst[k * C + channel] by the threads that own a column.
template <int VEC, int K>
__device__ __forceinline__ void in_cta_reduce(const float (&acc)[K][VEC], float* red, float* dst, int C, int vecs, int lanes,
                                              bool active) {
  const int t = threadIdx.x;
  if (active) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int e = 0; e < VEC; ++e) red[(t * K + k) * VEC + e] = acc[k][e];
  }
  __syncthreads();
  for (int j = t; j < K * C; j += IN_THREADS) {
    const int k = j / C, col = j - k * C;
    const int cv = col / VEC, e = col - cv * VEC;
    float s = 0.f;
    for (int rl = 0; rl < lanes; ++rl) s += red[((rl * vecs + cv) * K + k) * VEC + e];
    dst[j] = s;
  }
}

// partial[n][chunk][0][c] = sum x, [1][c] = sum x^2 over the chunk's rows
template <int VEC>
__global__ void __launch_bounds__(IN_THREADS)
inorm_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, long long S, int C, int vecs, int lanes,
                   int rows_per_cta) {
  __shared__ float red[IN_THREADS * 2 * VEC];
  const int t = threadIdx.x, n = blockIdx.y;
  const bool active = t < vecs * lanes;
  float acc[2][VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[0][e] = acc[1][e] = 0.f;
  if (active) {
    const int cv = t % vecs, rl = t / vecs;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < S) ? r0 + rows_per_cta : S;
    const __nv_bfloat16* base = x + static_cast<long long>(n) * S * C + cv * VEC;
    long long r = r0 + rl;
    for (; r + 3LL * lanes < r1; r += 4LL * lanes) {
      typename InVec<VEC>::T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = in_ld<VEC>(base + (r + static_cast<long long>(u) * lanes) * C);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[VEC];
        in_unpack<VEC>(v[u], f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { acc[0][e] += f[e]; acc[1][e] = fmaf(f[e], f[e], acc[1][e]); }
      }
    }
    for (; r < r1; r += lanes) {
      float f[VEC];
      in_unpack<VEC>(in_ld<VEC>(base + r * C), f);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { acc[0][e] += f[e]; acc[1][e] = fmaf(f[e], f[e], acc[1][e]); }
    }
  }
  in_cta_reduce<VEC, 2>(acc, red, partial + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * 2 * C, C, vecs, lanes, active);
}

// MODE 0: partial [N][chunks][2][C] -> out [N][2][C] = (mean, rstd);  MODE 1: partial [N][chunks][3][C] -> out [N][3][C] = sums / S
// One CTA per sample walks the sample's partials as one flat array with a stride that is a multiple of K*C, so a thread
// always meets the same (k, c) column and consecutive threads read consecutive floats; fp64 from here on, fixed order.
constexpr int FIN_THREADS = 1024;
template <int MODE>
__global__ void __launch_bounds__(FIN_THREADS)
inorm_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int chunks, int C, long long S, float eps) {
  constexpr int K = MODE == 0 ? 2 : 3;
  __shared__ double sm[FIN_THREADS];
  __shared__ double tot[FIN_THREADS];
  const int KC = K * C;
  const int n = blockIdx.x, t = threadIdx.x;
  const float* p = partial + static_cast<long long>(n) * chunks * KC;
  float* o = out + static_cast<long long>(n) * KC;
  const double inv = 1.0 / static_cast<double>(S);
  for (int col0 = 0; col0 < KC; col0 += FIN_THREADS) {        // one pass unless K * C > 1024
    const int width = (KC - col0 < FIN_THREADS) ? KC - col0 : FIN_THREADS;
    const int lanes = FIN_THREADS / width;                     // chunk lanes working side by side
    double acc = 0.0;
    if (t < lanes * width) {
      const int col = col0 + t % width;
      for (int ch = t / width; ch < chunks; ch += lanes) acc += static_cast<double>(p[static_cast<long long>(ch) * KC + col]);
    }
    sm[t] = acc;
    __syncthreads();
    if (t < width) {
      double s = 0.0;
      for (int l = 0; l < lanes; ++l) s += sm[l * width + t];
      tot[t] = s;
    }
    __syncthreads();
    if (MODE == 1) {
      if (t < width) o[col0 + t] = static_cast<float>(tot[t] * inv);
    } else if (KC <= FIN_THREADS) {
      if (t < C) {
        const double mean = tot[t] * inv;
        double var = tot[C + t] * inv - mean * mean;
        if (var < 0.0) var = 0.0;
        o[t] = static_cast<float>(mean);
        o[C + t] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      }
    } else {
      if (t < width) o[col0 + t] = static_cast<float>(tot[t]);   // raw sums; turned into (mean, rstd) below
    }
    __syncthreads();
  }
  if (MODE == 0 && KC > FIN_THREADS) {
    for (int c = t; c < C; c += FIN_THREADS) {
      const double mean = static_cast<double>(o[c]) * inv;
      double var = static_cast<double>(o[C + c]) * inv - mean * mean;
      if (var < 0.0) var = 0.0;
      o[c] = static_cast<float>(mean);
      o[C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
}

// y = lrelu( (a - mean_a) rstd_a  [+ (b - mean_b) rstd_b  |  + b] )
template <int VEC, int BMODE>   // BMODE 0: no b, 1: raw b, 2: normalised b
__global__ void __launch_bounds__(IN_THREADS)
inorm_apply_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                   const float* __restrict__ stats_a, const float* __restrict__ stats_b, long long S, int C, int vecs, int lanes,
                   int rows_per_cta, float slope) {
  const int t = threadIdx.x, n = blockIdx.y;
  if (t >= vecs * lanes) return;
  const int cv = t % vecs, rl = t / vecs;
  float ma[VEC], ra[VEC], mb[VEC], rb[VEC];
  in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + cv * VEC, ma);
  in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + C + cv * VEC, ra);
  if (BMODE == 2) {
    in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + cv * VEC, mb);
    in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + C + cv * VEC, rb);
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) ma[e] = -ma[e] * ra[e];            // z = a * rstd + (-mean * rstd)
  if (BMODE == 2) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) ma[e] = fmaf(-mb[e], rb[e], ma[e]);   // both shifts in one constant
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = (r0 + rows_per_cta < S) ? r0 + rows_per_cta : S;
  const long long off = static_cast<long long>(n) * S * C + cv * VEC;
  auto one = [&](const typename InVec<VEC>::T& va, const typename InVec<VEC>::T& vb, long long r) {
    float fa[VEC], fb[VEC], o[VEC];
    in_unpack<VEC>(va, fa);
    if (BMODE) in_unpack<VEC>(vb, fb);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float z = fmaf(fa[e], ra[e], ma[e]);
      if (BMODE == 1) z += fb[e];
      if (BMODE == 2) z = fmaf(fb[e], rb[e], z);
      o[e] = z > 0.f ? z : z * slope;
    }
    *reinterpret_cast<typename InVec<VEC>::T*>(y + off + r * C) = in_pack<VEC>(o);
  };
  long long r = r0 + rl;
  for (; r + 3LL * lanes < r1; r += 4LL * lanes) {
    typename InVec<VEC>::T va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      va[u] = in_ld<VEC>(a + off + (r + static_cast<long long>(u) * lanes) * C);
      if (BMODE) vb[u] = in_ld<VEC>(b + off + (r + static_cast<long long>(u) * lanes) * C);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(va[u], vb[u], r + static_cast<long long>(u) * lanes);
  }
  for (; r < r1; r += lanes) {
    typename InVec<VEC>::T va = in_ld<VEC>(a + off + r * C), vb = va;
    if (BMODE) vb = in_ld<VEC>(b + off + r * C);
    one(va, vb, r);
  }
}

// dz = dy * lrelu'(y);  partial[n][chunk][0] = sum dz, [1] = sum dz * a_hat, [2] = sum dz * b_hat (BMODE 2, else 0)
template <int VEC, int BMODE>
__global__ void __launch_bounds__(IN_THREADS)
inorm_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                        const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                        const float* __restrict__ stats_a, const float* __restrict__ stats_b, float* __restrict__ partial,
                        long long S, int C, int vecs, int lanes, int rows_per_cta, float slope) {
  __shared__ float red[IN_THREADS * 3 * VEC];
  const int t = threadIdx.x, n = blockIdx.y;
  const bool active = t < vecs * lanes;
  float acc[3][VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[0][e] = acc[1][e] = acc[2][e] = 0.f;
  if (active) {
    const int cv = t % vecs, rl = t / vecs;
    float ma[VEC], ra[VEC], mb[VEC], rb[VEC];
    in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + cv * VEC, ma);
    in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + C + cv * VEC, ra);
    if (BMODE == 2) {
      in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + cv * VEC, mb);
      in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + C + cv * VEC, rb);
    }
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
    const long long r1 = (r0 + rows_per_cta < S) ? r0 + rows_per_cta : S;
    const long long off = static_cast<long long>(n) * S * C + cv * VEC;
    auto one = [&](const typename InVec<VEC>::T& vd, const typename InVec<VEC>::T& vy, const typename InVec<VEC>::T& va,
                   const typename InVec<VEC>::T& vb) {
      float fd[VEC], fy[VEC], fa[VEC], fb[VEC];
      in_unpack<VEC>(vd, fd);
      in_unpack<VEC>(va, fa);
      if (y) in_unpack<VEC>(vy, fy);
      if (BMODE == 2) in_unpack<VEC>(vb, fb);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float dz = (y && !(fy[e] > 0.f)) ? fd[e] * slope : fd[e];
        acc[0][e] += dz;
        acc[1][e] = fmaf(dz, (fa[e] - ma[e]) * ra[e], acc[1][e]);
        if (BMODE == 2) acc[2][e] = fmaf(dz, (fb[e] - mb[e]) * rb[e], acc[2][e]);
      }
    };
    long long r = r0 + rl;
    for (; r + lanes < r1; r += 2LL * lanes) {
      typename InVec<VEC>::T vd[2], vy[2], va[2], vb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long o = off + (r + static_cast<long long>(u) * lanes) * C;
        vd[u] = in_ld<VEC>(dy + o);
        va[u] = in_ld<VEC>(a + o);
        vy[u] = y ? in_ld<VEC>(y + o) : vd[u];
        vb[u] = BMODE == 2 ? in_ld<VEC>(b + o) : vd[u];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) one(vd[u], vy[u], va[u], vb[u]);
    }
    for (; r < r1; r += lanes) {
      const long long o = off + r * C;
      const typename InVec<VEC>::T vd = in_ld<VEC>(dy + o), va = in_ld<VEC>(a + o);
      one(vd, y ? in_ld<VEC>(y + o) : vd, va, BMODE == 2 ? in_ld<VEC>(b + o) : vd);
    }
  }
  in_cta_reduce<VEC, 3>(acc, red, partial + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * 3 * C, C, vecs, lanes, active);
}

// da = rstd_a (dz - c1 - a_hat c2a);  db = dz (BMODE 1) | rstd_b (dz - c1 - b_hat c2b) (BMODE 2)
template <int VEC, int BMODE>
__global__ void __launch_bounds__(IN_THREADS, BMODE == 2 ? 3 : 0)
inorm_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                       const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                       const float* __restrict__ stats_a, const float* __restrict__ stats_b, const float* __restrict__ coef,
                       __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ db, long long S, int C, int vecs, int lanes,
                       int rows_per_cta, float slope) {
  const int t = threadIdx.x, n = blockIdx.y;
  if (t >= vecs * lanes) return;
  const int cv = t % vecs, rl = t / vecs;
  // da = dz * ra + a * A1 + A0  with  A1 = -ra^2 c2a,  A0 = -ra c1 - ma A1     (same form for b)
  float ra[VEC], a1[VEC], a0[VEC], rb[VEC], b1[VEC], b0[VEC];
  {
    float ma[VEC], c1[VEC], c2[VEC];
    in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + cv * VEC, ma);
    in_ldf<VEC>(stats_a + static_cast<long long>(n) * 2 * C + C + cv * VEC, ra);
    in_ldf<VEC>(coef + static_cast<long long>(n) * 3 * C + cv * VEC, c1);
    in_ldf<VEC>(coef + static_cast<long long>(n) * 3 * C + C + cv * VEC, c2);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      a1[e] = -ra[e] * ra[e] * c2[e];
      a0[e] = -ra[e] * c1[e] - ma[e] * a1[e];
    }
    if (BMODE == 2) {
      float mb[VEC], c3[VEC];
      in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + cv * VEC, mb);
      in_ldf<VEC>(stats_b + static_cast<long long>(n) * 2 * C + C + cv * VEC, rb);
      in_ldf<VEC>(coef + static_cast<long long>(n) * 3 * C + 2 * C + cv * VEC, c3);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        b1[e] = -rb[e] * rb[e] * c3[e];
        b0[e] = -rb[e] * c1[e] - mb[e] * b1[e];
      }
    }
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = (r0 + rows_per_cta < S) ? r0 + rows_per_cta : S;
  const long long off = static_cast<long long>(n) * S * C + cv * VEC;
  auto one = [&](const typename InVec<VEC>::T& vd, const typename InVec<VEC>::T& vy, const typename InVec<VEC>::T& va,
                 const typename InVec<VEC>::T& vb, long long o) {
    float fd[VEC], fy[VEC], fa[VEC], fb[VEC], oa[VEC], ob[VEC];
    in_unpack<VEC>(vd, fd);
    in_unpack<VEC>(va, fa);
    if (y) in_unpack<VEC>(vy, fy);
    if (BMODE == 2) in_unpack<VEC>(vb, fb);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float dz = (y && !(fy[e] > 0.f)) ? fd[e] * slope : fd[e];
      oa[e] = fmaf(dz, ra[e], fmaf(fa[e], a1[e], a0[e]));
      if (BMODE == 1) ob[e] = dz;
      if (BMODE == 2) ob[e] = fmaf(dz, rb[e], fmaf(fb[e], b1[e], b0[e]));
    }
    *reinterpret_cast<typename InVec<VEC>::T*>(da + o) = in_pack<VEC>(oa);
    if (BMODE) *reinterpret_cast<typename InVec<VEC>::T*>(db + o) = in_pack<VEC>(ob);
  };
  long long r = r0 + rl;
  for (; BMODE != 2 && r + lanes < r1; r += 2LL * lanes) {   // two operands: one row in flight, 80 registers, 3 CTAs per SM
    typename InVec<VEC>::T vd[2], vy[2], va[2], vb[2];
    long long o[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      o[u] = off + (r + static_cast<long long>(u) * lanes) * C;
      vd[u] = in_ld<VEC>(dy + o[u]);
      va[u] = in_ld<VEC>(a + o[u]);
      vy[u] = y ? in_ld<VEC>(y + o[u]) : vd[u];
      vb[u] = BMODE == 2 ? in_ld<VEC>(b + o[u]) : vd[u];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) one(vd[u], vy[u], va[u], vb[u], o[u]);
  }
  for (; r < r1; r += lanes) {
    const long long o = off + r * C;
    const typename InVec<VEC>::T vd = in_ld<VEC>(dy + o), va = in_ld<VEC>(a + o);
    one(vd, y ? in_ld<VEC>(y + o) : vd, va, BMODE == 2 ? in_ld<VEC>(b + o) : vd, o);
  }
}

static int in_check(const char* what, int N, long long S, int C, int vec) {
  if (N <= 0 || S <= 0 || C <= 0 || N > 65535) {
    set_last_error("%s: N=%d S=%lld C=%d out of range (1 <= N <= 65535)", what, N, S, C);
    return UCF_ERR_BAD_ARG;
  }
  if (vec == 0) {
    set_last_error("%s: C=%d must be even and <= 2048 (C / vector width <= 256), tensors 4-byte aligned", what, C);
    return UCF_ERR_UNSUPPORTED;
  }
  return UCF_OK;
}

}  // namespace ucf

using namespace ucf;

#define UCF_IN_VEC_DISPATCH(vec, M) \
  switch (vec) {                    \
    case 8: M(8) break;             \
    case 4: M(4) break;             \
    default: M(2) break;            \
  }

extern "C" int ucf_inorm_chunks(int N, long long S, int C) {
  if (N <= 0 || S <= 0 || C <= 0) return 0;
  int worst = 0;
  for (int v : {8, 4, 2}) {        // the vector width also depends on pointer alignment: size for the widest grid
    if (C % v != 0 || C / v > IN_THREADS) continue;
    const int ch = in_geom(N, S, C, v).chunks;
    if (ch > worst) worst = ch;
  }
  return worst;
}

extern "C" int ucf_inorm_stats(const void* x, int N, long long S, int C, float eps, float* workspace, float* stats,
                               void* stream) {
  const void* ptrs[] = {x};
  const int vec = in_vec(C, ptrs, 1);
  if (int e = in_check("inorm_stats", N, S, C, vec)) return e;
  if (!x || !workspace || !stats) { set_last_error("inorm_stats: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const InGeom g = in_geom(N, S, C, vec);
  const dim3 grid(g.chunks, N);
#define M(V) inorm_stats_kernel<V><<<grid, IN_THREADS, 0, st>>>(static_cast<const __nv_bfloat16*>(x), workspace, S, C, g.vecs, g.lanes, g.rows_per_cta);
  UCF_IN_VEC_DISPATCH(vec, M)
#undef M
  if (int e = check_launch("inorm_stats_kernel")) return e;
  inorm_finish_kernel<0><<<N, FIN_THREADS, 0, st>>>(workspace, stats, g.chunks, C, S, eps);
  return check_launch("inorm_finish_kernel");
}

extern "C" int ucf_inorm_apply(const void* a, const float* stats_a, const void* b, const float* stats_b, void* y, int N,
                               long long S, int C, float slope, void* stream) {
  const void* ptrs[] = {a, b, y};
  const int vec = in_vec(C, ptrs, 3);
  if (int e = in_check("inorm_apply", N, S, C, vec)) return e;
  if (!a || !stats_a || !y || (stats_b && !b)) { set_last_error("inorm_apply: null pointer"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const InGeom g = in_geom(N, S, C, vec);
  const dim3 grid(g.chunks, N);
  const int bmode = !b ? 0 : (stats_b ? 2 : 1);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(b);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
#define M(V)                                                                                                                  \
  if (bmode == 0) inorm_apply_kernel<V, 0><<<grid, IN_THREADS, 0, st>>>(ap, bp, yp, stats_a, stats_b, S, C, g.vecs, g.lanes, g.rows_per_cta, slope); \
  else if (bmode == 1) inorm_apply_kernel<V, 1><<<grid, IN_THREADS, 0, st>>>(ap, bp, yp, stats_a, stats_b, S, C, g.vecs, g.lanes, g.rows_per_cta, slope); \
  else inorm_apply_kernel<V, 2><<<grid, IN_THREADS, 0, st>>>(ap, bp, yp, stats_a, stats_b, S, C, g.vecs, g.lanes, g.rows_per_cta, slope);
  UCF_IN_VEC_DISPATCH(vec, M)
#undef M
  return check_launch("inorm_apply_kernel");
}

extern "C" int ucf_inorm_bwd(const void* dy, const void* y, const void* a, const float* stats_a, const void* b,
                             const float* stats_b, void* da, void* db, int N, long long S, int C, float slope,
                             float* workspace, float* coef, void* stream) {
  const void* ptrs[] = {dy, y, a, b, da, db};
  const int vec = in_vec(C, ptrs, 6);
  if (int e = in_check("inorm_bwd", N, S, C, vec)) return e;
  const int bmode = !db ? 0 : (stats_b ? 2 : 1);
  if (!dy || !a || !stats_a || !da || !workspace || !coef || (bmode == 2 && !b)) {
    set_last_error("inorm_bwd: null pointer");
    return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const InGeom g = in_geom(N, S, C, vec);
  const dim3 grid(g.chunks, N);
  const __nv_bfloat16* dyp = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* yp = static_cast<const __nv_bfloat16*>(y);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(b);
  __nv_bfloat16* dap = static_cast<__nv_bfloat16*>(da);
  __nv_bfloat16* dbp = static_cast<__nv_bfloat16*>(db);
#define M(V)                                                                                                                      \
  if (bmode == 2) inorm_bwd_reduce_kernel<V, 2><<<grid, IN_THREADS, 0, st>>>(dyp, yp, ap, bp, stats_a, stats_b, workspace, S, C, g.vecs, g.lanes, g.rows_per_cta, slope); \
  else inorm_bwd_reduce_kernel<V, 0><<<grid, IN_THREADS, 0, st>>>(dyp, yp, ap, bp, stats_a, stats_b, workspace, S, C, g.vecs, g.lanes, g.rows_per_cta, slope);
  UCF_IN_VEC_DISPATCH(vec, M)
#undef M
  if (int e = check_launch("inorm_bwd_reduce_kernel")) return e;
  inorm_finish_kernel<1><<<N, FIN_THREADS, 0, st>>>(workspace, coef, g.chunks, C, S, 0.f);
  if (int e = check_launch("inorm_finish_kernel")) return e;
#define M(V)                                                                                                                      \
  if (bmode == 0) inorm_bwd_apply_kernel<V, 0><<<grid, IN_THREADS, 0, st>>>(dyp, yp, ap, bp, stats_a, stats_b, coef, dap, dbp, S, C, g.vecs, g.lanes, g.rows_per_cta, slope); \
  else if (bmode == 1) inorm_bwd_apply_kernel<V, 1><<<grid, IN_THREADS, 0, st>>>(dyp, yp, ap, bp, stats_a, stats_b, coef, dap, dbp, S, C, g.vecs, g.lanes, g.rows_per_cta, slope); \
  else inorm_bwd_apply_kernel<V, 2><<<grid, IN_THREADS, 0, st>>>(dyp, yp, ap, bp, stats_a, stats_b, coef, dap, dbp, S, C, g.vecs, g.lanes, g.rows_per_cta, slope);
  UCF_IN_VEC_DISPATCH(vec, M)
#undef M
  return check_launch("inorm_bwd_apply_kernel");
}
