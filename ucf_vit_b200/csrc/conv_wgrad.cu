// Weight gradient of the UNETR decoder's 3x3x3 convolutions (stride 1, padding 1, no bias) on channels-last bf16 tensors:
//     dW[co][ci][kd][kh][kw] = sum_{n,z,y,x} dY[n, z, y, x, co] * X[n, z+kd-1, y+kh-1, x+kw-1, ci]
// (autograd of the nn.Conv3d inside MONAI's UnetResBlock / UnetBasicBlock that /root/reference/src/UCF_VIT/simple/arch.py:808-940
// builds).  At this decoder's widths (16-64 channels) the library answers with a legacy sm_80 implicit-GEMM kernel at
// ~130 TFLOP/s (27 + 4 ms of the 107 ms UNETR-128 step, profiles/r02_unetr_kernel_profile_fused.log).
//
// Shape of the work: a GEMM with M = Co (16 / 32), N = Ci (16..64) per filter tap and K = all voxels (up to 33 M).  tcgen05 does
// not fit: one UMMA would be M = 64 (padded) x N <= 64, operand-bandwidth-bound, and the 27 shifted views of an activation tile
// cannot be expressed as swizzled UMMA descriptors over ONE shared-memory halo tile -- so this kernel keeps
// the halo tile in shared memory, gathers fragments with ldmatrix (per-lane row addresses: a tap is just an address offset) and
// issues warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).
//
//   grid  = (persistent CTAs, 3 / KD filter planes);  block = 9 warps, warp w owns taps (kd', kh, kw) = (all KD planes of the
//           CTA, w / 3, w % 3).  KD = 3 (narrow layers, Ci * Co <= 512: one CTA does all 27 taps, the dY tile and the halo are
//           loaded once) or KD = 1 (wider layers: 3x the accumulator registers would not fit, the planes go to grid.y)
//   tile  = 4 x 4 x 8 output voxels (K = 128 per tile), halo: (4 + KD - 1) x 6 x 10 input voxels, zero-filled outside
//   smem  = double-buffered {halo tile [240][Ci], dY tile [128][Co]}, 16-byte chunks XOR-swizzled by the row so that the 8 rows
//           of an ldmatrix 8x8 block fall into 8 different bank groups; the next tile is prefetched into registers under the MMAs
//   acc   = (Co/16) x (Ci/8) m16n8 fragments per warp, kept in registers over ALL tiles of the CTA
//   out   = fp32 partial [CTA][tap][Co][Ci]; a second kernel sums the CTAs in a fixed order (reproducible) into
//           dW [Co][Ci][3][3][3] fp32
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

constexpr int CW_TZ = 4, CW_TY = 4, CW_TX = 8, CW_VOX = CW_TZ * CW_TY * CW_TX;       // 128 output voxels per tile
constexpr int CW_HY = CW_TY + 2, CW_HX = CW_TX + 2;                                   // halo extent in y and x
constexpr int CW_WARPS = 9, CW_THREADS = CW_WARPS * 32;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` of row `r` in a tile whose rows are Q chunks long (Q = 2, 4 or 8)
template <int Q>
__device__ __forceinline__ uint32_t cw_off(int r, int c) {
  return static_cast<uint32_t>(r * Q + (c ^ ((r / (8 / Q)) % Q))) * 16u;
}

template <int CI, int CO, int KD>
__global__ void __launch_bounds__(CW_THREADS, (CI * CO <= 1024) ? 2 : 1)
conv3d_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                    int N, int D, int H, int W, int tiles_z, int tiles_y, int tiles_x) {
  constexpr int QI = CI / 8, QO = CO / 8;                 // 16-byte chunks per voxel row
  constexpr int MT = CO / 16, NT = CI / 8;                // m16 tiles, n8 tiles
  constexpr int CW_HALO = (CW_TZ + KD - 1) * CW_HY * CW_HX;   // 240 (one plane) or 360 (all three) input voxels
  constexpr int X_BYTES = CW_HALO * CI * 2, Y_BYTES = CW_VOX * CO * 2;
  constexpr int X_CHUNKS = CW_HALO * QI, Y_CHUNKS = CW_VOX * QO, CHUNKS = X_CHUNKS + Y_CHUNKS;
  constexpr int PER_THREAD = (CHUNKS + CW_THREADS - 1) / CW_THREADS;
  extern __shared__ __align__(128) uint8_t smem[];        // [2][X_BYTES + Y_BYTES]
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int kd0 = (KD == 3) ? 0 : blockIdx.y, kh = warp / 3, kw = warp % 3;
  const long long tiles_per_n = static_cast<long long>(tiles_z) * tiles_y * tiles_x;
  const long long tiles = tiles_per_n * N;

  float acc[KD][MT][NT][4];
#pragma unroll
  for (int p = 0; p < KD; ++p)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[p][m][n][e] = 0.f;

  uint4 pre[PER_THREAD];
  auto prefetch = [&](long long tile) {
    const int n = static_cast<int>(tile / tiles_per_n);
    long long r = tile - n * tiles_per_n;
    const int z0 = static_cast<int>(r / (tiles_y * tiles_x)) * CW_TZ;
    r %= tiles_y * tiles_x;
    const int y0 = static_cast<int>(r / tiles_x) * CW_TY, x0 = static_cast<int>(r % tiles_x) * CW_TX;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int c = t + i * CW_THREADS;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c < X_CHUNKS) {
        const int h = c / QI, q = c - h * QI;
        const int hz = h / (CW_HY * CW_HX), hy = (h / CW_HX) % CW_HY, hx = h % CW_HX;
        const int z = z0 + hz + kd0 - 1, y = y0 + hy - 1, xx = x0 + hx - 1;
        if (z >= 0 && z < D && y >= 0 && y < H && xx >= 0 && xx < W)
          v = __ldg(reinterpret_cast<const uint4*>(x + (((static_cast<long long>(n) * D + z) * H + y) * W + xx) * CI + q * 8));
      } else if (c < CHUNKS) {
        const int cc = c - X_CHUNKS;
        const int vx = cc / QO, q = cc - vx * QO;
        const int z = z0 + vx / (CW_TY * CW_TX), y = y0 + (vx / CW_TX) % CW_TY, xx = x0 + vx % CW_TX;
        v = __ldg(reinterpret_cast<const uint4*>(dy + (((static_cast<long long>(n) * D + z) * H + y) * W + xx) * CO + q * 8));
      }
      pre[i] = v;
    }
  };
  auto stash = [&](int buf) {
    uint8_t* base = smem + buf * (X_BYTES + Y_BYTES);
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int c = t + i * CW_THREADS;
      if (c < X_CHUNKS) {
        const int h = c / QI, q = c - h * QI;
        *reinterpret_cast<uint4*>(base + cw_off<QI>(h, q)) = pre[i];
      } else if (c < CHUNKS) {
        const int cc = c - X_CHUNKS;
        const int vx = cc / QO, q = cc - vx * QO;
        *reinterpret_cast<uint4*>(base + X_BYTES + cw_off<QO>(vx, q)) = pre[i];
      }
    }
  };

  // per-lane ldmatrix geometry: lane l feeds row (l % 8) of 8x8 block (l / 8)
  const int blk = lane >> 3, row8 = lane & 7;
  // A (dY^T): blocks (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15)
  const int a_k = row8 + 8 * (blk >> 1), a_q = blk & 1;
  // B (X):    blocks (k 0-7, n-tile j), (k 8-15, n-tile j), (k 0-7, n-tile j+1), (k 8-15, n-tile j+1)
  const int b_k = row8 + 8 * (blk & 1), b_q = blk >> 1;

  long long tile = blockIdx.x;
  int buf = 0;
  if (tile < tiles) { prefetch(tile); stash(0); }
  __syncthreads();
  for (; tile < tiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < tiles) prefetch(next);
    const uint32_t xs = smem_base + buf * (X_BYTES + Y_BYTES), ys = xs + X_BYTES;
#pragma unroll 2
    for (int ks = 0; ks < CW_VOX / 16; ++ks) {
      // voxel of this lane's A / B row: v = 16 ks + k;  (tz, ty, tx) = (v / 32, (v / 8) % 4, v % 8)
      const int va = 16 * ks + a_k;
      const int vb = 16 * ks + b_k;
      uint32_t a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) ldmatrix_x4_trans(a[m], ys + cw_off<QO>(va, 2 * m + a_q));
#pragma unroll
      for (int p = 0; p < KD; ++p) {
        const int hb = (((vb >> 5) + p) * CW_HY + ((vb >> 3) & 3) + kh) * CW_HX + (vb & 7) + kw;
#pragma unroll
        for (int n = 0; n < NT; n += 2) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, xs + cw_off<QI>(hb, n + b_q));
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            mma_bf16_16816(acc[p][m][n], a[m], b[0], b[1]);
            mma_bf16_16816(acc[p][m][n + 1], a[m], b[2], b[3]);
          }
        }
      }
    }
    if (next < tiles) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // partial[cta][tap][co][ci]; fragment (m16n8): c0,c1 -> (row lane/4, cols 2(lane%4)+{0,1}); c2,c3 -> row + 8
#pragma unroll
  for (int p = 0; p < KD; ++p) {
    float* out = partial + ((static_cast<long long>(blockIdx.x) * 27 + (kd0 + p) * 9 + warp) * CO) * CI;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int co = m * 16 + (lane >> 2), ci = n * 8 + 2 * (lane & 3);
        *reinterpret_cast<float2*>(out + co * CI + ci) = make_float2(acc[p][m][n][0], acc[p][m][n][1]);
        *reinterpret_cast<float2*>(out + (co + 8) * CI + ci) = make_float2(acc[p][m][n][2], acc[p][m][n][3]);
      }
  }
}

// dW[co][ci][tap] = sum over CTAs of partial[cta][tap][co][ci] (fixed order)
__global__ void __launch_bounds__(256)
conv3d_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ctas, int CO, int CI) {
  const int i = blockIdx.x * 256 + threadIdx.x;           // index into [tap][co][ci]
  const int total = 27 * CO * CI;
  if (i >= total) return;
  float s = 0.f;
  for (int c = 0; c < ctas; ++c) s += partial[static_cast<long long>(c) * total + i];
  const int tap = i / (CO * CI), r = i - tap * CO * CI;
  dw[r * 27 + tap] = s;                                    // r = co * CI + ci
}

template <int CI, int CO, int KD>
static int cw_launch(const void* x, const void* dy, float* ws, float* dw, int N, int D, int H, int W, int ctas, cudaStream_t st) {
  constexpr int SMEM = 2 * ((CW_TZ + KD - 1) * CW_HY * CW_HX * CI * 2 + CW_VOX * CO * 2);
  static DeviceOnce once;
  bool& attr = once.flag();
  if (SMEM > 48 * 1024 && !attr) {
    if (cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_kernel<CI, CO, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM))
      return static_cast<int>(e);
    attr = true;
  }
  conv3d_wgrad_kernel<CI, CO, KD><<<dim3(ctas, 3 / KD), CW_THREADS, SMEM, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), ws, N, D, H, W, D / CW_TZ, H / CW_TY, W / CW_TX);
  if (int e = check_launch("conv3d_wgrad_kernel")) return e;
  conv3d_wgrad_reduce_kernel<<<(27 * CO * CI + 255) / 256, 256, 0, st>>>(ws, dw, ctas, CO, CI);
  return check_launch("conv3d_wgrad_reduce_kernel");
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_conv3d_wgrad_supported(int Ci, int Co, int D, int H, int W) {
  const bool ch = (Ci == 16 && Co == 16) || (Ci == 32 && Co == 16) || (Ci == 32 && Co == 32) || (Ci == 64 && Co == 32);
  return ch && D > 0 && H > 0 && W > 0 && D % CW_TZ == 0 && H % CW_TY == 0 && W % CW_TX == 0;
}

extern "C" int ucf_conv3d_wgrad_ctas(int N, int D, int H, int W) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  const long long tiles = static_cast<long long>(N) * (D / CW_TZ) * (H / CW_TY) * (W / CW_TX);
  const long long want = static_cast<long long>(num_sms()) * 2;                // x grid.y filter-plane groups
  return static_cast<int>(tiles < want ? (tiles > 0 ? tiles : 1) : want);
}

extern "C" int ucf_conv3d_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Ci, int Co,
                                float* workspace, void* stream) {
  if (!x || !dy || !dw || !workspace) { set_last_error("conv3d_wgrad: null pointer"); return UCF_ERR_BAD_ARG; }
  if (N <= 0 || !ucf_conv3d_wgrad_supported(Ci, Co, D, H, W)) {
    set_last_error("conv3d_wgrad: (Ci, Co) = (%d, %d) with D, H, W = %d, %d, %d is not served (channel pairs (16,16) (32,16) "
                   "(32,32) (64,32); D, H multiples of 4, W of 8)", Ci, Co, D, H, W);
    return UCF_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) {
    set_last_error("conv3d_wgrad: tensors must be 16-byte aligned");
    return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ctas = ucf_conv3d_wgrad_ctas(N, D, H, W);
  if (Ci == 16 && Co == 16) return cw_launch<16, 16, 3>(x, dy, workspace, dw, N, D, H, W, ctas, st);
  if (Ci == 32 && Co == 16) return cw_launch<32, 16, 3>(x, dy, workspace, dw, N, D, H, W, ctas, st);
  if (Ci == 32 && Co == 32) return cw_launch<32, 32, 1>(x, dy, workspace, dw, N, D, H, W, ctas, st);
  return cw_launch<64, 32, 1>(x, dy, workspace, dw, N, D, H, W, ctas, st);
}
