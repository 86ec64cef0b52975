// SAP adaptive-patching front end on the device: the two OpenCV calls the reference makes on every natural (uint8) image
// before it builds the quadtree,
//     grey_img = cv.GaussianBlur(img, (k, k), 0);  edges = cv.Canny(grey_img, c, c + 50)
// (/root/reference/src/UCF_VIT/dataloaders/transform.py:33-34), restated bit-for-bit in integer arithmetic
// (oracle/canny_np.py is the pinned CPU statement; OpenCV itself is the witness in the tests).
//
// HBM-bound byte / integer work (no tensor cores): every pass stages a tile + halo in shared memory.
//   blur       1 read + 1 write of the H x W x C bytes
//   nms        1 read of the blurred bytes, 1 write of the H x W class map (2 strong / 0 candidate / 1 none)
//   hysteresis sweeps over the class map: a tile converges in shared memory, tiles exchange through their halos, the
//              sweep is repeated until no tile changed (flag read back by the host: the launcher synchronises `stream`)
//   finalize   class map -> 0 / 255
// The host pays 39 + 226 ms for a 4096 x 4096 x 3 image (8 threads); see profiles/ for the device numbers.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i = (i < 0 ? -i : i) % period;
  return i >= n ? period - i : i;
}

// ---- Gaussian blur, k = 3 or 5, sigma 0: integer taps [1 2 1] / [1 4 6 4 1], one rounding at the end ------------------
constexpr int BL_TH = 32, BL_TWB = 512, BL_THREADS = 256;   // tile: 32 rows x 512 byte columns (x and channel interleaved)

template <int K>
__global__ void __launch_bounds__(BL_THREADS)
blur_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int C, int aligned4) {
  constexpr int R = K / 2;
  constexpr int MAXHALO = R * 4;                       // bytes of halo per side at C = 4
  constexpr int SW = BL_TWB + 2 * MAXHALO;
  __shared__ uint8_t tile[(BL_TH + 2 * R) * SW];
  const int halo = R * C;
  const int rowbytes = W * C;
  const int b0 = blockIdx.x * BL_TWB, y0 = blockIdx.y * BL_TH;
  const int sw = BL_TWB + 2 * halo;
  // a warp per tile row: interior bytes are a straight coalesced copy, only bytes outside the image pay for the
  // floor-division by C and the reflection (4 integer divisions)
  const int lane = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < BL_TH + 2 * R; r += BL_THREADS / 32) {
    const int yy = y0 + r - R;
    const uint8_t* srow = src + static_cast<long long>((yy >= 0 && yy < H) ? yy : reflect101(yy, H)) * rowbytes;
    for (int cb = lane; cb < sw; cb += 32) {
      const int gb = b0 + cb - halo;                   // byte column; pixel = floor(gb / C) (gb may be negative)
      int src_b = gb;
      if (gb < 0 || gb >= rowbytes) {
        const int x = (gb >= 0) ? gb / C : -((-gb + C - 1) / C);
        src_b = reflect101(x, W) * C + (gb - x * C);
      }
      tile[r * SW + cb] = srow[src_b];
    }
  }
  __syncthreads();
  constexpr int w3[3] = {1, 2, 1};
  constexpr int w5[5] = {1, 4, 6, 4, 1};
  constexpr int SHIFT = (K == 3) ? 4 : 8;
  const int col4 = threadIdx.x % (BL_TWB / 4), rg = threadIdx.x / (BL_TWB / 4);
  const int gb = b0 + col4 * 4;
  if (gb >= rowbytes) return;
  for (int r = rg; r < BL_TH; r += BL_THREADS / (BL_TWB / 4)) {
    const int gy = y0 + r;
    if (gy >= H) break;
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const uint8_t* row = tile + (r + i) * SW + col4 * 4;        // column offset 0 here = byte column -halo of the output
      int h[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int wj = (K == 3) ? w3[j] : w5[j];
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] += wj * row[e + j * C];
      }
      const int wi = (K == 3) ? w3[i] : w5[i];
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += wi * h[e];
    }
    uint8_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = static_cast<uint8_t>((acc[e] + (1 << (SHIFT - 1))) >> SHIFT);
    uint8_t* out = dst + static_cast<long long>(gy) * rowbytes + gb;
    if (aligned4 && gb + 4 <= rowbytes) {
      *reinterpret_cast<uint32_t*>(out) = o[0] | (o[1] << 8) | (o[2] << 16) | (static_cast<uint32_t>(o[3]) << 24);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (gb + e < rowbytes) out[e] = o[e];
    }
  }
}

// ---- Sobel 3x3 (replicated border) + channel of largest |dx| + |dy| + non-maximum suppression + double threshold ------
constexpr int NM_TH = 32, NM_TW = 64, NM_THREADS = 256;
constexpr int CANNY_TG22 = 13573;   // (int)(0.41421356237309504880 * (1 << 15) + 0.5)

__global__ void __launch_bounds__(NM_THREADS)
canny_nms_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ pmap, int H, int W, int C, int low, int high) {
  constexpr int PH = NM_TH + 4, PW = NM_TW + 4;       // pixels: tile + 2
  constexpr int MH = NM_TH + 2, MW = NM_TW + 2;       // gradients: tile + 1
  __shared__ uint8_t px[PH * PW * 4];
  __shared__ short s_mag[MH * MW], s_dx[MH * MW], s_dy[MH * MW];
  const int x0 = blockIdx.x * NM_TW, y0 = blockIdx.y * NM_TH;
  for (int i = threadIdx.x; i < PH * PW; i += NM_THREADS) {
    const int r = i / PW, c = i - r * PW;
    int gy = y0 + r - 2, gx = x0 + c - 2;
    gy = gy < 0 ? 0 : (gy >= H ? H - 1 : gy);          // BORDER_REPLICATE
    gx = gx < 0 ? 0 : (gx >= W ? W - 1 : gx);
    const uint8_t* p = img + (static_cast<long long>(gy) * W + gx) * C;
    for (int ch = 0; ch < C; ++ch) px[i * 4 + ch] = p[ch];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MH * MW; i += NM_THREADS) {
    const int r = i / MW, c = i - r * MW;
    const int gy = y0 + r - 1, gx = x0 + c - 1;
    short bm = 0, bdx = 0, bdy = 0;                    // magnitude is 0 outside the image
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      int best = -1;
      const uint8_t* q = px + (r * PW + c) * 4;        // top-left of the 3x3 window (pixel tile is offset by one more)
      for (int ch = 0; ch < C; ++ch) {
        const int a00 = q[ch], a01 = q[4 + ch], a02 = q[8 + ch];
        const int a10 = q[PW * 4 + ch], a12 = q[PW * 4 + 8 + ch];
        const int a20 = q[2 * PW * 4 + ch], a21 = q[2 * PW * 4 + 4 + ch], a22 = q[2 * PW * 4 + 8 + ch];
        const int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        const int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
        const int m = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
        if (m > best) { best = m; bdx = static_cast<short>(dx); bdy = static_cast<short>(dy); }
      }
      bm = static_cast<short>(best);
    }
    s_mag[i] = bm; s_dx[i] = bdx; s_dy[i] = bdy;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NM_TH * NM_TW; i += NM_THREADS) {
    const int r = i / NM_TW, c = i - r * NM_TW;
    const int gy = y0 + r, gx = x0 + c;
    if (gy >= H || gx >= W) continue;
    const int k = (r + 1) * MW + (c + 1);
    const int m = s_mag[k];
    uint8_t cls = 1;
    if (m > low) {
      const int xs = s_dx[k], ys = s_dy[k];
      const int x = xs < 0 ? -xs : xs;
      const int y = (ys < 0 ? -ys : ys) << 15;
      const int tg22x = x * CANNY_TG22;
      bool keep;
      if (y < tg22x) {
        keep = m > s_mag[k - 1] && m >= s_mag[k + 1];
      } else {
        const int tg67x = tg22x + (x << 16);
        if (y > tg67x) {
          keep = m > s_mag[k - MW] && m >= s_mag[k + MW];
        } else {
          const int s = ((xs ^ ys) < 0) ? -1 : 1;
          keep = m > s_mag[k - MW - s] && m > s_mag[k + MW + s];
        }
      }
      if (keep) cls = (m > high) ? 2 : 0;
    }
    pmap[static_cast<long long>(gy) * W + gx] = cls;
  }
}

// ---- hysteresis: candidates (0) 8-connected to a strong pixel (2) become strong ---------------------------------------
constexpr int HY_T = 64, HY_THREADS = 256;

__global__ void __launch_bounds__(HY_THREADS)
canny_hysteresis_kernel(uint8_t* __restrict__ pmap, int H, int W, int* __restrict__ changed_flag) {
  constexpr int SW = HY_T + 2;
  __shared__ uint8_t t_s[SW * SW];
  volatile uint8_t* t = t_s;                           // neighbours are written by other threads between the barriers
  const int x0 = blockIdx.x * HY_T, y0 = blockIdx.y * HY_T;
  int has_candidate = 0;
  for (int i = threadIdx.x; i < SW * SW; i += HY_THREADS) {
    const int r = i / SW, c = i - r * SW;
    const int gy = y0 + r - 1, gx = x0 + c - 1;
    const uint8_t v = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? pmap[static_cast<long long>(gy) * W + gx] : 1;
    t[i] = v;
    has_candidate |= (v == 0) && r >= 1 && r <= HY_T && c >= 1 && c <= HY_T;
  }
  if (!__syncthreads_or(has_candidate)) return;
  // each thread owns a 4 x 4 patch: a change inside the patch propagates within the same pass (row-major sweep, then the
  // reverse sweep), so long chains converge in few block-wide iterations
  const int pr = (threadIdx.x / 16) * 4 + 1, pc = (threadIdx.x % 16) * 4 + 1;
  int any = 0;
  for (;;) {
    int ch = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      for (int q = 0; q < 16; ++q) {
        const int qq = pass ? 15 - q : q;
        const int r = pr + qq / 4, c = pc + qq % 4;
        const int k = r * SW + c;
        if (t[k] != 0) continue;
        const bool hit = t[k - SW - 1] == 2 || t[k - SW] == 2 || t[k - SW + 1] == 2 || t[k - 1] == 2 || t[k + 1] == 2 ||
                         t[k + SW - 1] == 2 || t[k + SW] == 2 || t[k + SW + 1] == 2;
        if (hit) { t[k] = 2; ch = 1; }
      }
    }
    any |= ch;
    if (!__syncthreads_or(ch)) break;
  }
  if (any) {
    for (int q = 0; q < 16; ++q) {
      const int r = pr + q / 4, c = pc + q % 4;
      const int gy = y0 + r - 1, gx = x0 + c - 1;
      if (gy < H && gx < W && t[r * SW + c] == 2) pmap[static_cast<long long>(gy) * W + gx] = 2;
    }
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) *changed_flag = 1;
}

__global__ void __launch_bounds__(256)
canny_finalize_kernel(const uint8_t* __restrict__ pmap, uint8_t* __restrict__ edges, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(pmap) | reinterpret_cast<uintptr_t>(edges)) & 3) == 0) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(pmap + i);
    uint32_t o = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) o |= (((v >> (8 * e)) & 0xff) == 2 ? 0xffu : 0u) << (8 * e);
    *reinterpret_cast<uint32_t*>(edges + i) = o;
  } else {
    for (long long j = i; j < n && j < i + 4; ++j) edges[j] = pmap[j] == 2 ? 255 : 0;
  }
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_gaussian_blur_u8(const void* src, void* dst, int H, int W, int C, int ksize, void* stream) {
  if (H <= 0 || W <= 0) return UCF_OK;
  if (!src || !dst || src == dst) { set_last_error("gaussian_blur_u8: null or aliased pointers"); return UCF_ERR_BAD_ARG; }
  if (C < 1 || C > 4 || (ksize != 1 && ksize != 3 && ksize != 5)) {
    set_last_error("gaussian_blur_u8: C=%d must be 1..4 and ksize=%d one of 1, 3, 5 (the reference's sths)", C, ksize);
    return UCF_ERR_UNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rowbytes = static_cast<long long>(W) * C;
  if (ksize == 1) {
    if (cudaError_t e = cudaMemcpyAsync(dst, src, rowbytes * H, cudaMemcpyDeviceToDevice, st)) return static_cast<int>(e);
    return UCF_OK;
  }
  const dim3 grid(static_cast<unsigned>((rowbytes + BL_TWB - 1) / BL_TWB), (H + BL_TH - 1) / BL_TH);
  if (grid.y > 65535) { set_last_error("gaussian_blur_u8: H=%d too large", H); return UCF_ERR_UNSUPPORTED; }
  const int aligned4 = (rowbytes % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 4 == 0);
  if (ksize == 3) blur_u8_kernel<3><<<grid, BL_THREADS, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), H, W, C, aligned4);
  else blur_u8_kernel<5><<<grid, BL_THREADS, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), H, W, C, aligned4);
  return check_launch("blur_u8_kernel");
}

extern "C" int ucf_canny_u8(const void* img, int H, int W, int C, double low_thresh, double high_thresh, void* class_map,
                            void* edges, int* flag_dev, int* flag_host_pinned, int* sweeps_out, void* stream) {
  if (H <= 0 || W <= 0) return UCF_OK;
  if (!img || !class_map || !edges || !flag_dev || !flag_host_pinned) { set_last_error("canny_u8: null pointer"); return UCF_ERR_BAD_ARG; }
  if (C < 1 || C > 4) { set_last_error("canny_u8: C=%d must be 1..4", C); return UCF_ERR_UNSUPPORTED; }
  if ((H + NM_TH - 1) / NM_TH > 65535) { set_last_error("canny_u8: H=%d too large", H); return UCF_ERR_UNSUPPORTED; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (low_thresh > high_thresh) { const double t = low_thresh; low_thresh = high_thresh; high_thresh = t; }
  const int low = static_cast<int>(floor(low_thresh)), high = static_cast<int>(floor(high_thresh));
  uint8_t* pm = static_cast<uint8_t*>(class_map);
  canny_nms_kernel<<<dim3((W + NM_TW - 1) / NM_TW, (H + NM_TH - 1) / NM_TH), NM_THREADS, 0, st>>>(
      static_cast<const uint8_t*>(img), pm, H, W, C, low, high);
  if (int e = check_launch("canny_nms_kernel")) return e;
  const dim3 hgrid((W + HY_T - 1) / HY_T, (H + HY_T - 1) / HY_T);
  int sweeps = 0;
  for (;;) {
    if (cudaError_t e = cudaMemsetAsync(flag_dev, 0, sizeof(int), st)) return static_cast<int>(e);
    canny_hysteresis_kernel<<<hgrid, HY_THREADS, 0, st>>>(pm, H, W, flag_dev);
    if (int e = check_launch("canny_hysteresis_kernel")) return e;
    ++sweeps;
    if (cudaError_t e = cudaMemcpyAsync(flag_host_pinned, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st)) return static_cast<int>(e);
    if (cudaError_t e = cudaStreamSynchronize(st)) return static_cast<int>(e);
    if (*flag_host_pinned == 0) break;
    if (sweeps > H + W) { set_last_error("canny_u8: hysteresis did not converge in %d sweeps", sweeps); return UCF_ERR_BAD_ARG; }
  }
  if (sweeps_out) *sweeps_out = sweeps;
  const long long n = static_cast<long long>(H) * W;
  canny_finalize_kernel<<<static_cast<unsigned>((n + 1023) / 1024), 256, 0, st>>>(pm, static_cast<uint8_t*>(edges), n);
  return check_launch("canny_finalize_kernel");
}
