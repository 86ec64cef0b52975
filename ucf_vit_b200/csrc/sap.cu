// Adaptive patching (SAP): greedy quadtree / octree construction (host, integer, bit-exact with the
// reference's node order) and the per-leaf resampling gather / scatter kernels (device).
// Replaces FixedQuadTree._build_tree / serialize / deserialize
// (/root/reference/src/UCF_VIT/dataloaders/quadtree.py:115-137,144-174,209-221, Rect.set_area :25-36)
// and the FixedOctTree twins (/root/reference/src/UCF_VIT/dataloaders/octree.py:72-150,201-213).
//
// Tree build: the reference re-scans a Python list for the first maximum (O(L^2)) and slices the
// edge map per candidate.  Here: a priority queue ordered by (value desc, DFS path asc) -- children
// replace their parent in place, so list order == DFS order -- O(L log L) queue work.  Box sums: uint8
// edge maps (the Canny output the drivers feed) use an integer summed-area table over 8 x 8 pixel cells built
// in one SIMD pass (psadbw) plus direct sums of the thin unaligned borders of a box -- exact, and 1/64 of the
// memory of a per-pixel table, which was 134 MB and 135 ms at 4096^2.  Float maps use the same cell table in double;
// volumes keep a per-voxel table (float) or direct sums (uint8).  A batch entry point builds the trees of several images on host
// threads (the calls release the GIL).
// Gather/scatter: 2-D = OpenCV INTER_CUBIC (Keys a=-0.75, half-pixel centres, replicated border, NO
// antialiasing: each output needs 16 taps, so traffic is L*p^2*16*C reads -- latency-, not
// bandwidth-bound); uint8 images use OpenCV's 11-bit fixed-point coefficients.  3-D = align-corners
// trilinear (scipy RegularGridInterpolator on linspace(0,s,s) grids).
#if defined(__SSE2__) || defined(_M_X64)
#include <emmintrin.h>
#define UCF_HOST_SSE2 1
#endif
#include <algorithm>
#include <cstring>
#include <atomic>
#include <queue>
#include <thread>
#include <vector>

#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// ------------------------------------------------------------------------------------------------
// host: tree construction
// ------------------------------------------------------------------------------------------------
// Sum of n bytes.  SSE2: psadbw adds 16 bytes per instruction into two 64-bit lanes; elsewhere a SWAR loop
// over 8-byte words with 16-bit lanes folded every 128 words (128 * 2 * 255 < 2^16).
static inline unsigned long long sum_bytes(const uint8_t* p, int n) {
  unsigned long long s = 0;
  int i = 0;
#ifdef UCF_HOST_SSE2
  __m128i acc = _mm_setzero_si128();
  const __m128i zero = _mm_setzero_si128();
  for (; i + 16 <= n; i += 16)
    acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + i)), zero));
  unsigned long long lanes[2];
  _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), acc);
  s = lanes[0] + lanes[1];
#else
  const unsigned long long m = 0x00FF00FF00FF00FFull;
  while (i + 8 <= n) {
    unsigned long long acc = 0;
    int k = 0;
    for (; k < 128 && i + 8 <= n; ++k, i += 8) {
      unsigned long long v;
      memcpy(&v, p + i, 8);
      acc += (v & m) + ((v >> 8) & m);
    }
    s += (acc & 0xFFFF) + ((acc >> 16) & 0xFFFF) + ((acc >> 32) & 0xFFFF) + (acc >> 48);
  }
#endif
  for (; i < n; ++i) s += p[i];
  return s;
}

// Exact box sums of a uint8 map by direct summation over its contiguous rows.
struct DirectSumU8 {
  const uint8_t* d;
  int n0, n1, n2;   // 2-D: [n0 rows, n1 cols], n2 unused; 3-D: [z, y, x]
  unsigned long long sum2(int x1, int x2, int y1, int y2) const {
    unsigned long long s = 0;
    for (int y = y1; y < y2; ++y) s += sum_bytes(d + static_cast<size_t>(y) * n1 + x1, x2 - x1);
    return s;
  }
  unsigned long long sum3(int x1, int x2, int y1, int y2, int z1, int z2) const {
    unsigned long long s = 0;
    for (int z = z1; z < z2; ++z)
      for (int y = y1; y < y2; ++y) s += sum_bytes(d + (static_cast<size_t>(z) * n1 + y) * n2 + x1, x2 - x1);
    return s;
  }
};

// uint8 edge maps, 2-D: summed-area table over 8 x 8 pixel cells (one pass over the image, 1/64 of the entries
// of a per-pixel table: 2 MB at 4096^2, cache resident) + direct sums of the < 8 pixel wide border strips of a
// box whose edges are not multiples of 8.  Exact in integers, so values and node order match a per-pixel sum.
struct CellSatU8 {
  DirectSumU8 px;
  int ch, cw;                            // whole cells per column / row
  std::vector<unsigned long long> s;     // (ch + 1) x (cw + 1)
  CellSatU8(const uint8_t* d, int H, int W) : px{d, H, W, 0}, ch(H / 8), cw(W / 8),
                                              s(static_cast<size_t>(H / 8 + 1) * (W / 8 + 1), 0ull) {
    std::vector<unsigned long long> cell(static_cast<size_t>(cw) + 2);
    for (int cy = 0; cy < ch; ++cy) {
      std::fill(cell.begin(), cell.end(), 0ull);
      for (int r = 0; r < 8; ++r) {
        const uint8_t* row = d + static_cast<size_t>(cy * 8 + r) * W;
        int cx = 0;
#ifdef UCF_HOST_SSE2
        const __m128i zero = _mm_setzero_si128();
        for (; cx + 2 <= cw; cx += 2) {    // psadbw: bytes 0-7 and 8-15 -> two 64-bit lanes = two adjacent cells
          const __m128i v = _mm_sad_epu8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(row + cx * 8)), zero);
          __m128i* acc = reinterpret_cast<__m128i*>(&cell[cx]);
          _mm_storeu_si128(acc, _mm_add_epi64(_mm_loadu_si128(acc), v));
        }
#endif
        for (; cx < cw; ++cx) cell[cx] += sum_bytes(row + cx * 8, 8);
      }
      unsigned long long run = 0;
      const size_t up = static_cast<size_t>(cy) * (cw + 1), here = up + (cw + 1);
      for (int cx = 0; cx < cw; ++cx) {
        run += cell[cx];
        s[here + cx + 1] = s[up + cx + 1] + run;
      }
    }
  }
  unsigned long long cells(int cx1, int cx2, int cy1, int cy2) const {
    auto at = [&](int y, int x) { return s[static_cast<size_t>(y) * (cw + 1) + x]; };
    return at(cy2, cx2) - at(cy1, cx2) - at(cy2, cx1) + at(cy1, cx1);
  }
  unsigned long long sum2(int x1, int x2, int y1, int y2) const {
    const int ax1 = (x1 + 7) & ~7, ax2 = x2 & ~7, ay1 = (y1 + 7) & ~7, ay2 = y2 & ~7;
    if (ax1 >= ax2 || ay1 >= ay2) return px.sum2(x1, x2, y1, y2);
    return cells(ax1 >> 3, ax2 >> 3, ay1 >> 3, ay2 >> 3) +
           px.sum2(x1, x2, y1, ay1) + px.sum2(x1, x2, ay2, y2) +        // rows above / below the aligned core
           px.sum2(x1, ax1, ay1, ay2) + px.sum2(ax2, x2, ay1, ay2);     // columns left / right of it
  }
};

// Float edge maps, 2-D: the same 8 x 8 cell table in double (the reference sums each box with numpy; any
// double-precision evaluation order agrees with it to ~1e-9 of a unit at 4096^2, far below the int() step).
template <typename T>
struct CellSatF {
  const T* d;
  int H, W, ch, cw;
  std::vector<double> s;                 // (ch + 1) x (cw + 1)
  CellSatF(const T* dom, int h, int w) : d(dom), H(h), W(w), ch(h / 8), cw(w / 8),
                                         s(static_cast<size_t>(h / 8 + 1) * (w / 8 + 1), 0.0) {
    std::vector<double> cell(static_cast<size_t>(cw) + 1);
    for (int cy = 0; cy < ch; ++cy) {
      std::fill(cell.begin(), cell.end(), 0.0);
      for (int r = 0; r < 8; ++r) {
        const T* row = d + static_cast<size_t>(cy * 8 + r) * W;
        for (int cx = 0; cx < cw; ++cx) {
          const T* q = row + cx * 8;
          cell[cx] += ((static_cast<double>(q[0]) + q[1]) + (static_cast<double>(q[2]) + q[3])) +
                      ((static_cast<double>(q[4]) + q[5]) + (static_cast<double>(q[6]) + q[7]));
        }
      }
      double run = 0.0;
      const size_t up = static_cast<size_t>(cy) * (cw + 1), here = up + (cw + 1);
      for (int cx = 0; cx < cw; ++cx) {
        run += cell[cx];
        s[here + cx + 1] = s[up + cx + 1] + run;
      }
    }
  }
  double direct(int x1, int x2, int y1, int y2) const {
    double t = 0.0;
    for (int y = y1; y < y2; ++y) {
      const T* row = d + static_cast<size_t>(y) * W;
      double rs = 0.0;
      for (int x = x1; x < x2; ++x) rs += static_cast<double>(row[x]);
      t += rs;
    }
    return t;
  }
  double sum(int x1, int x2, int y1, int y2) const {
    const int ax1 = (x1 + 7) & ~7, ax2 = x2 & ~7, ay1 = (y1 + 7) & ~7, ay2 = y2 & ~7;
    if (ax1 >= ax2 || ay1 >= ay2) return direct(x1, x2, y1, y2);
    auto at = [&](int y, int x) { return s[static_cast<size_t>(y) * (cw + 1) + x]; };
    const int cx1 = ax1 >> 3, cx2 = ax2 >> 3, cy1 = ay1 >> 3, cy2 = ay2 >> 3;
    return (at(cy2, cx2) - at(cy1, cx2) - at(cy2, cx1) + at(cy1, cx1)) + direct(x1, x2, y1, ay1) + direct(x1, x2, ay2, y2) +
           direct(x1, ax1, ay1, ay2) + direct(ax2, x2, ay1, ay2);
  }
};

template <typename Acc>
struct Sat3 {   // (Z+1) x (Y+1) x (X+1)
  int Z, Y, X;
  std::vector<Acc> s;
  template <typename T>
  Sat3(const T* d, int z, int y, int x) : Z(z), Y(y), X(x), s(static_cast<size_t>(z + 1) * (y + 1) * (x + 1), Acc(0)) {
    auto idx = [&](int k, int j, int i) { return (static_cast<size_t>(k) * (Y + 1) + j) * (X + 1) + i; };
    for (int k = 0; k < z; ++k)
      for (int j = 0; j < y; ++j) {
        Acc row = 0;
        for (int i = 0; i < x; ++i) {
          row += static_cast<Acc>(d[(static_cast<size_t>(k) * y + j) * x + i]);
          s[idx(k + 1, j + 1, i + 1)] = row + s[idx(k + 1, j, i + 1)] + s[idx(k, j + 1, i + 1)] - s[idx(k, j, i + 1)];
        }
      }
  }
  Acc sum(int x1, int x2, int y1, int y2, int z1, int z2) const {
    auto at = [&](int k, int j, int i) { return s[(static_cast<size_t>(k) * (Y + 1) + j) * (X + 1) + i]; };
    return at(z2, y2, x2) - at(z1, y2, x2) - at(z2, y1, x2) - at(z2, y2, x1) + at(z1, y1, x2) + at(z1, y2, x1) +
           at(z2, y1, x1) - at(z1, y1, x1);
  }
};

struct Leaf {
  long long value;
  unsigned long long path;   // DFS path, most significant digits first
  int depth;
  int c[6];
};
struct LeafLess {   // priority_queue top = largest value, then smallest path (first in list order)
  bool operator()(const Leaf& a, const Leaf& b) const {
    if (a.value != b.value) return a.value < b.value;
    return a.path > b.path;
  }
};

template <typename ValueFn>
static int build_tree(int ndim, const int root[6], int fixed_length, ValueFn value_of, int32_t* boxes, long long* values) {
  const int fan = ndim == 2 ? 4 : 8;
  const int bits = ndim == 2 ? 2 : 3;
  std::priority_queue<Leaf, std::vector<Leaf>, LeafLess> pq;
  Leaf r{};
  for (int i = 0; i < 6; ++i) r.c[i] = root[i];
  r.value = value_of(r.c);
  r.path = 0;
  r.depth = 0;
  pq.push(r);
  int count = 1;
  while (count < fixed_length) {
    Leaf best = pq.top();
    if (best.c[1] - best.c[0] == 2) break;                 // selected leaf is 2 wide: stop (reference :124-125)
    if ((best.depth + 1) * bits > 64) break;               // cannot happen for sizes < 2^21
    pq.pop();
    const int x1 = best.c[0], x2 = best.c[1], y1 = best.c[2], y2 = best.c[3], z1 = best.c[4], z2 = best.c[5];
    const int xm = (x1 + x2) / 2, ym = (y1 + y2) / 2, zm = (z1 + z2) / 2;
    for (int k = 0; k < fan; ++k) {
      Leaf ch{};
      if (ndim == 2) {
        // order lt, rt, lb, rb with "t" = the HIGHER-y half (reference :128-135)
        const bool right = k & 1, low = k >= 2;
        ch.c[0] = right ? xm : x1; ch.c[1] = right ? x2 : xm;
        ch.c[2] = low ? y1 : ym;   ch.c[3] = low ? ym : y2;
        ch.c[4] = 0; ch.c[5] = 0;
      } else {
        // x fastest, then y, then z (octree.py:85-100)
        const bool xr = k & 1, yr = k & 2, zr = k & 4;
        ch.c[0] = xr ? xm : x1; ch.c[1] = xr ? x2 : xm;
        ch.c[2] = yr ? ym : y1; ch.c[3] = yr ? y2 : ym;
        ch.c[4] = zr ? zm : z1; ch.c[5] = zr ? z2 : zm;
      }
      ch.value = value_of(ch.c);
      ch.depth = best.depth + 1;
      ch.path = best.path | (static_cast<unsigned long long>(k) << (64 - bits * ch.depth));
      pq.push(ch);
    }
    count += fan - 1;
  }
  std::vector<Leaf> leaves;
  leaves.reserve(count);
  while (!pq.empty()) { leaves.push_back(pq.top()); pq.pop(); }
  std::sort(leaves.begin(), leaves.end(), [](const Leaf& a, const Leaf& b) { return a.path < b.path; });
  const int nc = ndim == 2 ? 4 : 6;
  for (size_t i = 0; i < leaves.size(); ++i) {
    for (int j = 0; j < nc; ++j) boxes[i * nc + j] = leaves[i].c[j];
    if (values) values[i] = leaves[i].value;
  }
  return static_cast<int>(leaves.size());
}

// ------------------------------------------------------------------------------------------------
// device: resampling
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
  c[3] = 1.f - c[0] - c[1] - c[2];
}
// destination index d of a (src -> dst) resize: first tap and fractional offset, as cv::resize
__device__ __forceinline__ void cubic_axis(int d, int src, int dst, int& tap0, float (&c)[4]) {
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  const int sx = static_cast<int>(floorf(fx));
  fx -= static_cast<float>(sx);
  tap0 = sx - 1;
  cubic_coeffs(fx, c);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// seq_out [L, p, p, C] float32; one block per sequence slot
template <typename T>
__global__ void __launch_bounds__(256)
sap_gather2d_kernel(const T* __restrict__ img, int H, int W, int C, const int32_t* __restrict__ boxes, int n_leaves,
                    int p, float* __restrict__ seq, long long* __restrict__ seq_size, double* __restrict__ seq_pos) {
  const int leaf = blockIdx.x;
  float* out = seq + static_cast<size_t>(leaf) * p * p * C;
  const int total = p * p * C;
  if (leaf >= n_leaves) {   // padding slot: zero patch, size 0, centre (-1,-1)
    for (int t = threadIdx.x; t < total; t += blockDim.x) out[t] = 0.f;
    if (threadIdx.x == 0) { seq_size[leaf] = 0; seq_pos[2 * leaf] = -1.0; seq_pos[2 * leaf + 1] = -1.0; }
    return;
  }
  const int x1 = boxes[4 * leaf], x2 = boxes[4 * leaf + 1], y1 = boxes[4 * leaf + 2], y2 = boxes[4 * leaf + 3];
  const int sw = x2 - x1, sh = y2 - y1;
  if (threadIdx.x == 0) {
    seq_size[leaf] = sw;
    seq_pos[2 * leaf] = (x2 + x1) / 2.0;
    seq_pos[2 * leaf + 1] = (y2 + y1) / 2.0;
  }
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int c = t % C;
    const int px = (t / C) % p;
    const int py = t / (C * p);
    if (sw == p && sh == p) {   // same size: cv::resize copies
      out[t] = static_cast<float>(img[(static_cast<size_t>(y1 + py) * W + x1 + px) * C + c]);
      continue;
    }
    int tx, ty;
    float cx[4], cy[4];
    cubic_axis(px, sw, p, tx, cx);
    cubic_axis(py, sh, p, ty, cy);
    if (sizeof(T) == 1) {
      // OpenCV fixed point: coefficients * 2048 rounded to short, int32 passes, (v + 2^21) >> 22
      int ix[4], iy[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ix[k] = clampi(__float2int_rn(cx[k] * 2048.f), -32768, 32767);
        iy[k] = clampi(__float2int_rn(cy[k] * 2048.f), -32768, 32767);
      }
      long long acc = 0;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = y1 + clampi(ty + ky, 0, sh - 1);
        int row = 0;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = x1 + clampi(tx + kx, 0, sw - 1);
          row += static_cast<int>(img[(static_cast<size_t>(yy) * W + xx) * C + c]) * ix[kx];
        }
        acc += static_cast<long long>(row) * iy[ky];
      }
      const long long v = (acc + (1ll << 21)) >> 22;
      out[t] = static_cast<float>(v < 0 ? 0 : (v > 255 ? 255 : v));
    } else {
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = y1 + clampi(ty + ky, 0, sh - 1);
        float row = 0.f;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = x1 + clampi(tx + kx, 0, sw - 1);
          row = __fmaf_rn(static_cast<float>(img[(static_cast<size_t>(yy) * W + xx) * C + c]), cx[kx], row);
        }
        acc = __fmaf_rn(row, cy[ky], acc);
      }
      out[t] = acc;
    }
  }
}

// mask [H, W, C] float32 <- per-leaf bicubic upsample of trunc(seq[leaf]) ; grid (n_leaves, slices)
__global__ void __launch_bounds__(256)
sap_scatter2d_kernel(const float* __restrict__ seq, const int32_t* __restrict__ boxes, int p, int C, int H, int W,
                     int truncate_to_int, float* __restrict__ mask) {
  const int leaf = blockIdx.x;
  const int x1 = boxes[4 * leaf], x2 = boxes[4 * leaf + 1], y1 = boxes[4 * leaf + 2], y2 = boxes[4 * leaf + 3];
  const int sw = x2 - x1, sh = y2 - y1;
  const float* src = seq + static_cast<size_t>(leaf) * p * p * C;
  const long long total = static_cast<long long>(sw) * sh * C;
  for (long long t = static_cast<long long>(blockIdx.y) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.y) * blockDim.x) {
    const int c = static_cast<int>(t % C);
    const int ox = static_cast<int>((t / C) % sw);
    const int oy = static_cast<int>(t / (static_cast<long long>(C) * sw));
    auto fetch = [&](int yy, int xx) {
      float v = src[(yy * p + xx) * C + c];
      return truncate_to_int ? truncf(v) : v;   // seq.astype(int) in the reference
    };
    float r;
    if (sw == p && sh == p) {
      r = fetch(oy, ox);
    } else {
      int tx, ty;
      float cx[4], cy[4];
      cubic_axis(ox, p, sw, tx, cx);
      cubic_axis(oy, p, sh, ty, cy);
      r = 0.f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = clampi(ty + ky, 0, p - 1);
        float row = 0.f;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) row = __fmaf_rn(fetch(yy, clampi(tx + kx, 0, p - 1)), cx[kx], row);
        r = __fmaf_rn(row, cy[ky], r);
      }
    }
    mask[(static_cast<size_t>(y1 + oy) * W + x1 + ox) * C + c] = r;
  }
}

// align-corners linear axis (scipy RegularGridInterpolator on linspace(0, s, s) -> linspace(0, s, n))
__device__ __forceinline__ void lin_axis(int j, int src, int dst, int& i0, double& w) {
  if (src <= 1 || dst <= 1) { i0 = 0; w = 0.0; return; }
  const double t = static_cast<double>(j) * (src - 1) / static_cast<double>(dst - 1);
  int i = static_cast<int>(floor(t));
  if (i > src - 2) i = src - 2;
  if (i < 0) i = 0;
  i0 = i;
  w = t - i;
}

// vol [Z, Y, X, C] float32 -> seq [L, p, p, p, C]
__global__ void __launch_bounds__(256)
sap_gather3d_kernel(const float* __restrict__ vol, int Z, int Y, int X, int C, const int32_t* __restrict__ boxes,
                    int n_leaves, int p, float* __restrict__ seq, long long* __restrict__ seq_size,
                    double* __restrict__ seq_pos) {
  const int leaf = blockIdx.x;
  const long long total = static_cast<long long>(p) * p * p * C;
  float* out = seq + static_cast<size_t>(leaf) * total;
  if (leaf >= n_leaves) {
    for (long long t = threadIdx.x; t < total; t += blockDim.x) out[t] = 0.f;
    if (threadIdx.x == 0) { seq_size[leaf] = 0; seq_pos[3 * leaf] = seq_pos[3 * leaf + 1] = seq_pos[3 * leaf + 2] = -1.0; }
    return;
  }
  const int* b = boxes + 6 * leaf;
  const int x1 = b[0], x2 = b[1], y1 = b[2], y2 = b[3], z1 = b[4], z2 = b[5];
  const int sx = x2 - x1, sy = y2 - y1, sz = z2 - z1;
  if (threadIdx.x == 0) {
    seq_size[leaf] = sx;
    seq_pos[3 * leaf] = (x2 + x1) / 2.0; seq_pos[3 * leaf + 1] = (y2 + y1) / 2.0; seq_pos[3 * leaf + 2] = (z2 + z1) / 2.0;
  }
  for (long long t = threadIdx.x; t < total; t += blockDim.x) {
    const int c = static_cast<int>(t % C);
    long long r = t / C;
    const int k2 = static_cast<int>(r % p); r /= p;     // crop axis 2 (x)
    const int k1 = static_cast<int>(r % p);             // crop axis 1 (y)
    const int k0 = static_cast<int>(r / p);             // crop axis 0 (z)
    int i0, i1, i2; double w0, w1, w2;
    lin_axis(k0, sz, p, i0, w0);
    lin_axis(k1, sy, p, i1, w1);
    lin_axis(k2, sx, p, i2, w2);
    auto at = [&](int dz, int dy, int dx) {
      const int zz = z1 + min(i0 + dz, sz - 1), yy = y1 + min(i1 + dy, sy - 1), xx = x1 + min(i2 + dx, sx - 1);
      return static_cast<double>(vol[((static_cast<size_t>(zz) * Y + yy) * X + xx) * C + c]);
    };
    double acc = 0.0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          acc += (dz ? w0 : 1.0 - w0) * (dy ? w1 : 1.0 - w1) * (dx ? w2 : 1.0 - w2) * at(dz, dy, dx);
    out[t] = static_cast<float>(acc);
  }
}

// mask [Z, Y, X, C] float32 <- trilinear resample of seq[leaf] (p^3) to the leaf box
__global__ void __launch_bounds__(256)
sap_scatter3d_kernel(const float* __restrict__ seq, const int32_t* __restrict__ boxes, int p, int C, int Z, int Y, int X,
                     float* __restrict__ mask) {
  const int leaf = blockIdx.x;
  const int* b = boxes + 6 * leaf;
  const int x1 = b[0], x2 = b[1], y1 = b[2], y2 = b[3], z1 = b[4], z2 = b[5];
  const int sx = x2 - x1, sy = y2 - y1, sz = z2 - z1;
  const float* src = seq + static_cast<size_t>(leaf) * p * p * p * C;
  const long long total = static_cast<long long>(sx) * sy * sz * C;
  for (long long t = static_cast<long long>(blockIdx.y) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.y) * blockDim.x) {
    const int c = static_cast<int>(t % C);
    long long r = t / C;
    const int ox = static_cast<int>(r % sx); r /= sx;
    const int oy = static_cast<int>(r % sy);
    const int oz = static_cast<int>(r / sy);
    int i0, i1, i2; double w0, w1, w2;
    lin_axis(oz, p, sz, i0, w0);
    lin_axis(oy, p, sy, i1, w1);
    lin_axis(ox, p, sx, i2, w2);
    auto at = [&](int dz, int dy, int dx) {
      const int zz = min(i0 + dz, p - 1), yy = min(i1 + dy, p - 1), xx = min(i2 + dx, p - 1);
      return static_cast<double>(src[((static_cast<size_t>(zz) * p + yy) * p + xx) * C + c]);
    };
    double acc = 0.0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          acc += (dz ? w0 : 1.0 - w0) * (dy ? w1 : 1.0 - w1) * (dx ? w2 : 1.0 - w2) * at(dz, dy, dx);
    mask[((static_cast<size_t>(z1 + oz) * Y + y1 + oy) * X + x1 + ox) * C + c] = static_cast<float>(acc);
  }
}

}  // namespace ucf

using namespace ucf;

extern "C" int ucf_sap_build_tree_host(const void* domain_host, int domain_dtype, int ndim, int n0, int n1, int n2,
                                       double norm_factor, int fixed_length, int32_t* boxes_host,
                                       long long* values_host) {
  if (!domain_host || !boxes_host || fixed_length < 1 || (ndim != 2 && ndim != 3) || n0 <= 0 || n1 <= 0 ||
      (ndim == 3 && n2 <= 0) || norm_factor == 0.0) {
    set_last_error("sap_build_tree: bad arguments");
    return UCF_ERR_BAD_ARG;
  }
  if (domain_dtype != UCF_DTYPE_U8 && domain_dtype != UCF_DTYPE_F32 && domain_dtype != UCF_DTYPE_F64) {
    set_last_error("sap_build_tree: domain dtype must be u8, f32 or f64");
    return UCF_ERR_BAD_ARG;
  }
  if (ndim == 2) {
    const int H = n0, W = n1;
    const int root[6] = {0, W, 0, H, 0, 0};
    if (domain_dtype == UCF_DTYPE_U8) {
      const CellSatU8 ds(static_cast<const uint8_t*>(domain_host), H, W);
      const long long nf = static_cast<long long>(norm_factor);
      auto val = [&](const int* c) { return static_cast<long long>(static_cast<double>(ds.sum2(c[0], c[1], c[2], c[3])) / static_cast<double>(nf)); };
      return build_tree(2, root, fixed_length, val, boxes_host, values_host);
    }
    if (domain_dtype == UCF_DTYPE_F32) {
      const CellSatF<float> sat(static_cast<const float*>(domain_host), H, W);
      auto val = [&](const int* c) { return static_cast<long long>(sat.sum(c[0], c[1], c[2], c[3]) / norm_factor); };
      return build_tree(2, root, fixed_length, val, boxes_host, values_host);
    }
    const CellSatF<double> sat(static_cast<const double*>(domain_host), H, W);
    auto val = [&](const int* c) { return static_cast<long long>(sat.sum(c[0], c[1], c[2], c[3]) / norm_factor); };
    return build_tree(2, root, fixed_length, val, boxes_host, values_host);
  }
  // 3-D: reference root is Cube(0,h,0,w,0,d) while contains() indexes [z,y,x]: cubic tiles only
  if (n0 != n1 || n1 != n2) {
    set_last_error("sap_build_tree: the reference octree is only self-consistent for cubic tiles (got %dx%dx%d)", n0, n1, n2);
    return UCF_ERR_BAD_ARG;
  }
  const int root[6] = {0, n0, 0, n1, 0, n2};
  if (domain_dtype == UCF_DTYPE_U8) {
    const DirectSumU8 ds{static_cast<const uint8_t*>(domain_host), n0, n1, n2};
    auto val = [&](const int* c) { return static_cast<long long>(static_cast<double>(ds.sum3(c[0], c[1], c[2], c[3], c[4], c[5])) / norm_factor); };
    return build_tree(3, root, fixed_length, val, boxes_host, values_host);
  }
  if (domain_dtype == UCF_DTYPE_F32) {
    Sat3<double> sat(static_cast<const float*>(domain_host), n0, n1, n2);
    auto val = [&](const int* c) { return static_cast<long long>(sat.sum(c[0], c[1], c[2], c[3], c[4], c[5]) / norm_factor); };
    return build_tree(3, root, fixed_length, val, boxes_host, values_host);
  }
  Sat3<double> sat(static_cast<const double*>(domain_host), n0, n1, n2);
  auto val = [&](const int* c) { return static_cast<long long>(sat.sum(c[0], c[1], c[2], c[3], c[4], c[5]) / norm_factor); };
  return build_tree(3, root, fixed_length, val, boxes_host, values_host);
}

extern "C" int ucf_sap_build_tree_batch_host(const void* const* domains_host, int n_images, int domain_dtype, int ndim,
                                             int n0, int n1, int n2, double norm_factor, int fixed_length,
                                             int32_t* boxes_host, long long* values_host, int* n_leaves_host,
                                             int n_threads) {
  if (n_images < 0 || !n_leaves_host || (n_images > 0 && (!domains_host || !boxes_host))) {
    set_last_error("sap_build_tree_batch: bad arguments");
    return UCF_ERR_BAD_ARG;
  }
  if (n_images == 0) return UCF_OK;
  if (fixed_length < 1 || (ndim != 2 && ndim != 3)) { set_last_error("sap_build_tree_batch: bad arguments"); return UCF_ERR_BAD_ARG; }
  const int nc = ndim == 2 ? 4 : 6;
  const size_t rows = static_cast<size_t>(UCF_SAP_TREE_ROWS(fixed_length, ndim));
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  int workers = n_threads > 0 ? n_threads : static_cast<int>(hw);
  if (workers > n_images) workers = n_images;
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n_images; i = next.fetch_add(1)) {
      n_leaves_host[i] = ucf_sap_build_tree_host(domains_host[i], domain_dtype, ndim, n0, n1, n2, norm_factor, fixed_length,
                                                 boxes_host + static_cast<size_t>(i) * rows * nc,
                                                 values_host ? values_host + static_cast<size_t>(i) * rows : nullptr);
    }
  };
  if (workers <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(workers - 1);
    for (int t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
  }
  for (int i = 0; i < n_images; ++i)
    if (n_leaves_host[i] < 0) {        // the message of a failing image lives in its worker's thread-local slot: restate it
      const int code = n_leaves_host[i];
      set_last_error("sap_build_tree_batch: image %d was rejected (code %d): bad shape, dtype or a non-cubic volume", i, code);
      return code;
    }
  return UCF_OK;
}

extern "C" int ucf_sap_gather(const void* img, int img_dtype, int ndim, int n0, int n1, int n2, int C,
                              const int32_t* boxes, int n_leaves, int fixed_length, int p, float* seq,
                              long long* seq_size, double* seq_pos, void* stream) {
  if (!img || !boxes || !seq || !seq_size || !seq_pos || n_leaves < 0 || n_leaves > fixed_length || p < 1 || C < 1) {
    set_last_error("sap_gather: bad arguments");
    return UCF_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ndim == 2) {
    if (img_dtype == UCF_DTYPE_U8)
      sap_gather2d_kernel<uint8_t><<<fixed_length, 256, 0, st>>>(static_cast<const uint8_t*>(img), n0, n1, C, boxes, n_leaves, p, seq, seq_size, seq_pos);
    else if (img_dtype == UCF_DTYPE_F32)
      sap_gather2d_kernel<float><<<fixed_length, 256, 0, st>>>(static_cast<const float*>(img), n0, n1, C, boxes, n_leaves, p, seq, seq_size, seq_pos);
    else { set_last_error("sap_gather: 2-D image dtype must be u8 or f32"); return UCF_ERR_BAD_ARG; }
    return check_launch("sap_gather2d_kernel");
  }
  if (ndim == 3) {
    if (img_dtype != UCF_DTYPE_F32) { set_last_error("sap_gather: 3-D volume dtype must be f32"); return UCF_ERR_BAD_ARG; }
    sap_gather3d_kernel<<<fixed_length, 256, 0, st>>>(static_cast<const float*>(img), n0, n1, n2, C, boxes, n_leaves, p, seq, seq_size, seq_pos);
    return check_launch("sap_gather3d_kernel");
  }
  set_last_error("sap_gather: ndim must be 2 or 3");
  return UCF_ERR_BAD_ARG;
}

extern "C" int ucf_sap_scatter(const float* seq, int ndim, int n0, int n1, int n2, int C, const int32_t* boxes,
                               int n_leaves, int p, int truncate_to_int, float* mask, void* stream) {
  if (!seq || !boxes || !mask || n_leaves < 0 || p < 1 || C < 1) { set_last_error("sap_scatter: bad arguments"); return UCF_ERR_BAD_ARG; }
  if (n_leaves == 0) return UCF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(n_leaves, 8);
  if (ndim == 2) {
    sap_scatter2d_kernel<<<grid, 256, 0, st>>>(seq, boxes, p, C, n0, n1, truncate_to_int, mask);
    return check_launch("sap_scatter2d_kernel");
  }
  if (ndim == 3) {
    sap_scatter3d_kernel<<<grid, 256, 0, st>>>(seq, boxes, p, C, n0, n1, n2, mask);
    return check_launch("sap_scatter3d_kernel");
  }
  set_last_error("sap_scatter: ndim must be 2 or 3");
  return UCF_ERR_BAD_ARG;
}
