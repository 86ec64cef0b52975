// bf16 GEMM with fused epilogues on tcgen05 / TMEM / TMA (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )        fp32 accumulation in tensor memory
//
// This one kernel family serves every dense contraction of the transformer block
// (/root/reference/src/UCF_VIT/simple/building_blocks.py:115-128,150-191):
//   forward   x*W^T          A K-major,  B K-major   (nn.Linear weight is [out,in])
//   dgrad     dY*W           A K-major,  B MN-major  (W read in place, no transpose copy)
//   wgrad     dY^T*X         A MN-major, B MN-major  (split-K, fp32 TMA reduce-add into dW)
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM
// alloc), warps 2..9 = epilogue (TMEM -> registers -> swizzled smem -> TMA store).  The
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits;
  int kb_total, kb_per_split;
  const void* bias;   // length N or null
  int bias_is_bf16;
  float* bias_grad;   // wgrad only (A MN-major, fp32 reduce-add): [M] += row sums of A, i.e. the bias gradient
  // EPI_DELTA only (the attention-output dgrad dO = dX1 * Wproj, aux = O): delta[b, h, n] = sum_d dO * O per head,
  // the row statistic the attention backward kernel needs -- taken from the accumulator tile in the epilogue
  float* delta;       // [M / tokens, heads, tokens] fp32
  int tokens, heads, head_dim;
  // EPI_LN only (LayerNorm folded into the GEMM that consumes it): C[r, n] = rstd[r] * (acc[r, n] - mean[r] * colsum[n]) + bias[n]
  // with A = the RAW (un-normalised) rows, B = W * diag(gamma), colsum[n] = sum_k B[n, k], bias = W beta + b
  const float* ln_mean;     // [M]
  const float* ln_rstd;     // [M]
  const float* ln_colsum;   // [N]
};

enum { EPI_BIAS = 0, EPI_BIAS_RESIDUAL = 1, EPI_BIAS_GELU_AUX = 2, EPI_DGELU = 3, EPI_F32_ADD = 4, EPI_DELTA = 5, EPI_LN = 6 };

// Bias-gradient warps of the wgrad kernels.  In wgrad, A = dY^T (MN-major: 64 token rows x 128
// channels per stage, two 64-channel boxes of 128-byte swizzled rows), so the bias gradient
// db[c] = sum_tokens dY[token, c] is the row sum of A -- the data is already in shared memory.
// Two extra warps re-read each stage of the n_blk == 0 tiles AFTER the MMAs that consumed it have
// retired (they wait on the same `empty` barrier as the producer) and hand the stage back through
// `bias_done`; the producer waits for both.  Replaces a separate pass over dY (colsum kernel).
template <int STAGES, int STAGE_BYTES, int NBW>
__device__ __forceinline__ void bias_grad_warp_loop(const GemmParams& p, uint8_t* stage_base, uint64_t* empty_bar,
                                                    uint64_t* bias_done, int bw, int lane, int first_tile,
                                                    int tile_stride, int total_tiles, int m_rows_per_tile,
                                                    int m_sub_off) {
  uint32_t kiter = 0;
  const int c = lane & 15;                       // 16-byte chunk (8 channels) of the 128-channel row
  const int box = c >> 3, cin = c & 7;
  const int rsel = lane >> 4;                    // this lane reads rows of one parity
  for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
    const int n_blk = tile % p.tiles_n;
    const int rest = tile / p.tiles_n;
    const int m_blk = rest % p.tiles_m;
    const int split = rest / p.tiles_m;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const bool work = n_blk == 0 && p.bias_grad != nullptr;
    float2 acc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = make_float2(0.f, 0.f);
    for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
      const int s = kiter % STAGES;
      const uint32_t ph = (kiter / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph);              // the MMAs that read this fill of stage s have retired
      if (work) {
        const uint8_t* a = stage_base + s * STAGE_BYTES + box * 8192;
#pragma unroll 4
        for (int j = 0; j < 32 / NBW; ++j) {
          const int r = bw * (64 / NBW) + 2 * j + rsel;  // token row inside the stage (rows past K are zero-filled)
          const uint4 v = *reinterpret_cast<const uint4*>(a + r * 128 + ((cin ^ (r & 7)) << 4));
          acc[0] = __fadd2_rn(acc[0], bf16x2_to_f32x2(v.x));
          acc[1] = __fadd2_rn(acc[1], bf16x2_to_f32x2(v.y));
          acc[2] = __fadd2_rn(acc[2], bf16x2_to_f32x2(v.z));
          acc[3] = __fadd2_rn(acc[3], bf16x2_to_f32x2(v.w));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bias_done[s]);
    }
    if (work) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[e].x += __shfl_xor_sync(0xffffffffu, acc[e].x, 16);
        acc[e].y += __shfl_xor_sync(0xffffffffu, acc[e].y, 16);
      }
      if (lane < 16) {
        const int ch = m_blk * m_rows_per_tile + m_sub_off + c * 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ch + 2 * e < p.M) atomicAdd(&p.bias_grad[ch + 2 * e], acc[e].x);
          if (ch + 2 * e + 1 < p.M) atomicAdd(&p.bias_grad[ch + 2 * e + 1], acc[e].y);
        }
      }
    }
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool HAS_AUX = (EPI == EPI_BIAS_RESIDUAL || EPI == EPI_BIAS_GELU_AUX || EPI == EPI_DGELU);
  static constexpr bool AUX_IN = (EPI == EPI_BIAS_RESIDUAL || EPI == EPI_DGELU);
  // 8 epilogue warps (two per TMEM lane quarter, each owning half of the tile's columns) work in
  // chunks of 32 rows x 32 columns: bf16 -> 64-byte rows (SWIZZLE_64B), fp32 -> 128-byte rows.
  static constexpr int EPI_WARPS = 8;
  static constexpr int CW = 32;
  static constexpr int OUT_BUF = (EPI == EPI_F32_ADD) ? 4096 : 2048;
  static constexpr int AUX_BUF = 2048;
  static constexpr int OUT_NBUF = (EPI == EPI_F32_ADD) ? 1 : 2;   // wgrad: light epilogue, spend smem on stages
  static constexpr int OUT_STAGE_BYTES = EPI_WARPS * OUT_NBUF * OUT_BUF;
  static constexpr int AUX_STAGE_BYTES = HAS_AUX ? EPI_WARPS * 2 * AUX_BUF : 0;
  static constexpr bool BIASW = A_MN && EPI == EPI_F32_ADD;   // wgrad: two extra bias-gradient warps
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + (BIASW ? 64 : 0);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + OUT_STAGE_BYTES + AUX_STAGE_BYTES +
                                    BN * 4 + (3 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
  static_assert(2 * BN <= 512, "two accumulators must fit in TMEM");
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GemmCfg<BN, STAGES, A_MN, B_MN, EPI>::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, STAGES, A_MN, B_MN, EPI>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK;
  constexpr int A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic)
  uint8_t* stage_base = smem;
  uint8_t* epi_out = smem + STAGES * STAGE_BYTES;
  uint8_t* epi_aux = epi_out + Cfg::OUT_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(epi_aux + Cfg::AUX_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;   // [EPI_WARPS][2]
  uint64_t* bias_done = aux_bar + 2 * Cfg::EPI_WARPS;   // [STAGES] (wgrad bias-gradient warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bias_done + STAGES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler: uniform branches / registers
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (Cfg::HAS_AUX) tma_prefetch_desc(&tmAux);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], Cfg::EPI_WARPS);
    }
    for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    for (int i = 0; i < STAGES; ++i) mbar_init(&bias_done[i], 2);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {   // converged warp; one elected lane issues the TMA loads (descriptors stay in uniform registers)
      uint32_t kiter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.tiles_n;
        const int rest = tile / p.tiles_n;
        const int m_blk = rest % p.tiles_m;
        const int split = rest / p.tiles_m;
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (Cfg::BIASW) mbar_wait(&bias_done[s], ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            uint8_t* a_dst = stage_base + s * STAGE_BYTES;
            uint8_t* b_dst = a_dst + A_BYTES;
            if (!A_MN) {
              tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
            } else {
              tma_load_2d(a_dst, &tmA, &full_bar[s], m0, kb * BK);
              tma_load_2d(a_dst + 8192, &tmA, &full_bar[s], m0 + 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], n0 + j * 64, kb * BK);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      uint32_t kiter = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int rest = tile / p.tiles_n;
        const int split = rest / p.tiles_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = A_MN ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc(a_addr + k * 32, 16, 1024);
              const uint64_t bdesc = B_MN ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc(b_addr + k * 32, 16, 1024);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);   // smem slot reusable once these MMAs retire
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&tfull_bar[buf]);   // accumulator complete
        __syncwarp();
      }
    }
  } else if (Cfg::BIASW && warp >= 2 + Cfg::EPI_WARPS) {
    // ------------------------------------------------------------------ bias-gradient warps (wgrad)
    bias_grad_warp_loop<STAGES, STAGE_BYTES, 2>(p, stage_base, empty_bar, bias_done, warp - 2 - Cfg::EPI_WARPS, lane,
                                             blockIdx.x, gridDim.x, total_tiles, BM, 0);
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int ew = warp - 2;            // 0..7
    const int half = ew >> 2;           // which half of the tile's columns this warp owns
    const int etid = threadIdx.x - 64;
    constexpr int CW = Cfg::CW;
    constexpr int NCHUNK = BN / 2 / CW;   // chunks per warp
    constexpr int OUT_BUF = Cfg::OUT_BUF, AUX_BUF = Cfg::AUX_BUF;
    uint8_t* my_out = epi_out + ew * Cfg::OUT_NBUF * OUT_BUF;
    uint8_t* my_aux = epi_aux + ew * 2 * AUX_BUF;
    uint64_t* my_aux_bar = aux_bar + ew * 2;
    // byte offset of this lane's 16-byte chunk j inside a staging buffer
    //   bf16: 64-byte rows, SWIZZLE_64B  -> chunk ^= (row >> 1) & 3
    //   fp32: 128-byte rows, SWIZZLE_128B -> chunk ^= row & 7
    const uint32_t sw64 = (lane >> 1) & 3, sw128 = lane & 7;
    uint32_t cc = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % p.tiles_n;
      const int rest = tile / p.tiles_n;
      const int m_blk = rest % p.tiles_m;
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int buf = it & 1;
      const int r0 = m0 + q * 32;             // first row of this warp's 32-row slab
      const int cbase = half * (BN / 2);      // first column (inside the tile) of this warp's half

      if (EPI != EPI_F32_ADD && EPI != EPI_DGELU) {
        named_bar_sync(1, 32 * Cfg::EPI_WARPS);
        for (int i = etid; i < BN; i += 32 * Cfg::EPI_WARPS) {
          float b = 0.f;
          if (p.bias != nullptr && n0 + i < p.N)
            b = p.bias_is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n0 + i])
                               : reinterpret_cast<const float*>(p.bias)[n0 + i];
          bias_s[i] = b;
        }
        named_bar_sync(1, 32 * Cfg::EPI_WARPS);
      }
      const bool active = r0 < p.M && n0 + cbase < p.N;
      if (Cfg::AUX_IN && lane == 0 && active) {
        mbar_expect_tx(&my_aux_bar[cc & 1], AUX_BUF);
        tma_load_2d(my_aux + (cc & 1) * AUX_BUF, &tmAux, &my_aux_bar[cc & 1], n0 + cbase, r0);
      }
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * BN + cbase + (static_cast<uint32_t>(q * 32) << 16);

      if (active) {
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = n0 + cbase + c * CW;
          if (col0 >= p.N) break;
          const uint32_t b = cc & 1;
          if (Cfg::AUX_IN && lane == 0 && c + 1 < NCHUNK && col0 + CW < p.N) {
            mbar_expect_tx(&my_aux_bar[b ^ 1], AUX_BUF);
            tma_load_2d(my_aux + (b ^ 1) * AUX_BUF, &tmAux, &my_aux_bar[b ^ 1], col0 + CW, r0);
          }
          uint32_t v[32];
          tmem_ld32(t_row + c * CW, v);
          tmem_wait_ld();
          if (lane == 0) {                           // staging buffer about to be rewritten is drained
            if (Cfg::OUT_NBUF == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
          }
          __syncwarp();
          if (Cfg::AUX_IN) mbar_wait(&my_aux_bar[b], (cc >> 1) & 1);

          const uint32_t ob = (Cfg::OUT_NBUF == 2) ? b : 0;
          if (EPI == EPI_F32_ADD) {
            uint8_t* out_row = my_out + ob * OUT_BUF + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(out_row + ((j ^ sw128) << 4)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint8_t* out_row = my_out + ob * OUT_BUF + lane * 64;
            uint8_t* aux_row = my_aux + b * AUX_BUF + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                f[e] = make_float2(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1]));
              if (EPI != EPI_DGELU) {
                const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[cbase + c * CW + 8 * j]);
                const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[cbase + c * CW + 8 * j + 4]);
                f[0] = __fadd2_rn(f[0], make_float2(b0.x, b0.y));
                f[1] = __fadd2_rn(f[1], make_float2(b0.z, b0.w));
                f[2] = __fadd2_rn(f[2], make_float2(b1.x, b1.y));
                f[3] = __fadd2_rn(f[3], make_float2(b1.z, b1.w));
              }
              const uint32_t sw = (static_cast<uint32_t>(j) ^ sw64) << 4;
              if (EPI == EPI_BIAS_RESIDUAL) {
                const uint4 r = *reinterpret_cast<const uint4*>(aux_row + sw);
                f[0] = __fadd2_rn(f[0], bf16x2_to_f32x2(r.x));
                f[1] = __fadd2_rn(f[1], bf16x2_to_f32x2(r.y));
                f[2] = __fadd2_rn(f[2], bf16x2_to_f32x2(r.z));
                f[3] = __fadd2_rn(f[3], bf16x2_to_f32x2(r.w));
              } else if (EPI == EPI_DGELU) {
                const uint4 r = *reinterpret_cast<const uint4*>(aux_row + sw);
                f[0] = __fmul2_rn(f[0], gelu_grad_poly2(bf16x2_to_f32x2(r.x)));
                f[1] = __fmul2_rn(f[1], gelu_grad_poly2(bf16x2_to_f32x2(r.y)));
                f[2] = __fmul2_rn(f[2], gelu_grad_poly2(bf16x2_to_f32x2(r.z)));
                f[3] = __fmul2_rn(f[3], gelu_grad_poly2(bf16x2_to_f32x2(r.w)));
              } else if (EPI == EPI_BIAS_GELU_AUX) {
                // pre-activation is rounded to bf16 first so that backward (which re-reads the
                // stored bf16 z) differentiates exactly the function forward evaluated.
                uint4 zq = make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y),
                                      pack_bf16x2(f[2].x, f[2].y), pack_bf16x2(f[3].x, f[3].y));
                *reinterpret_cast<uint4*>(aux_row + sw) = zq;
                f[0] = gelu_poly2(bf16x2_to_f32x2(zq.x));
                f[1] = gelu_poly2(bf16x2_to_f32x2(zq.y));
                f[2] = gelu_poly2(bf16x2_to_f32x2(zq.z));
                f[3] = gelu_poly2(bf16x2_to_f32x2(zq.w));
              }
              *reinterpret_cast<uint4*>(out_row + sw) =
                  make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y),
                             pack_bf16x2(f[3].x, f[3].y));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (EPI == EPI_F32_ADD) {
              tma_reduce_add_2d(&tmC, my_out + ob * OUT_BUF, col0, r0);
            } else {
              tma_store_2d(&tmC, my_out + ob * OUT_BUF, col0, r0);
              if (EPI == EPI_BIAS_GELU_AUX) tma_store_2d(&tmAux, my_aux + b * AUX_BUF, col0, r0);
            }
            tma_store_commit();
          }
          ++cc;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: a cluster of two CTAs (one TPC) works on a 256 x BN tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and HALF of the B tile, so the
// shared-memory traffic per MMA flop is halved and a pipeline stage is 16 KB smaller.  Only the
// leader CTA (cluster rank 0) issues MMAs; its commits are multicast to both CTAs' barriers; both
// CTAs' TMA loads credit the leader's "full" barrier; each CTA's epilogue warps drain the 128
// accumulator rows that live in their own TMEM.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
struct Gemm2Cfg {
  static constexpr int BM = 128, BK = 64;                    // per CTA; the pair covers 256 rows
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;          // this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool HAS_AUX = (EPI == EPI_BIAS_RESIDUAL || EPI == EPI_BIAS_GELU_AUX || EPI == EPI_DGELU || EPI == EPI_DELTA);
  static constexpr bool AUX_IN = (EPI == EPI_BIAS_RESIDUAL || EPI == EPI_DGELU || EPI == EPI_DELTA);
  // The GELU / GELU' epilogues carry ~20 issue slots per element pair plus a second tensor: with two
  // warps per scheduler the epilogue of tile i did not fit under the MMAs of tile i+1 (measured 234 us
  // against 178 us for the plain kernel, issue slots 40 % busy -- latency-, not throughput-bound).
  // They run 16 epilogue warps (four per TMEM lane quarter, 64 columns each) with single staging
  // buffers; the light epilogues keep 8 warps (128 columns each) with double buffers.
  static constexpr bool HEAVY = (EPI == EPI_BIAS_GELU_AUX || EPI == EPI_DGELU);
  static constexpr int EPI_WARPS = HEAVY ? 16 : 8;
  static constexpr int CW = 32;
  static constexpr int OUT_BUF = (EPI == EPI_F32_ADD) ? 4096 : 2048;
  static constexpr int AUX_BUF = 2048;
  static constexpr int OUT_NBUF = (HEAVY || EPI == EPI_F32_ADD) ? 1 : 2;   // wgrad: one epilogue per ~100 k-blocks, spend smem on stages
  static constexpr int AUX_NBUF = (HEAVY && !AUX_IN) ? 1 : 2;   // aux as an input is prefetched one chunk ahead
  static constexpr int OUT_STAGE_BYTES = EPI_WARPS * OUT_NBUF * OUT_BUF;
  static constexpr int AUX_STAGE_BYTES = HAS_AUX ? EPI_WARPS * AUX_NBUF * AUX_BUF : 0;
  static constexpr bool BIASW = A_MN && EPI == EPI_F32_ADD;   // wgrad: two extra bias-gradient warps
  // warps: 0 = TMA producer (A operand), 1 = MMA issuer, 2.. = epilogue, then the two bias-gradient warps
  // (wgrad), last = second TMA producer (B operand).  One issuing thread tops out at 55 B/clk of TMA loads,
  // two reach 70 (scripts/micro/tma_load_rate.cu) -- and a 256x256 pair tile needs 64 B/clk per CTA.
  static constexpr int BIAS_WARPS = 4;     // with two, re-reading a stage after its MMAs took ~350 cycles and delayed its refill
  static constexpr int PRODUCER2_WARP = 2 + EPI_WARPS + (BIASW ? BIAS_WARPS : 0);
  static constexpr int THREADS = 32 * (PRODUCER2_WARP + 1);
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_STAGE_BYTES + AUX_STAGE_BYTES + BN * 4 * (EPI == EPI_LN ? 2 : 1) +
                                    (3 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
  static_assert(BN == 256, "pair kernel is instantiated for BN = 256");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg<BN, STAGES, A_MN, B_MN, EPI>::THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                  const GemmParams p) {
  using Cfg = Gemm2Cfg<BN, STAGES, A_MN, B_MN, EPI>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, HN = BN / 2;
  constexpr int A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic)
  uint8_t* stage_base = smem;
  uint8_t* epi_out = smem + STAGES * STAGE_BYTES;
  uint8_t* epi_aux = epi_out + Cfg::OUT_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(epi_aux + Cfg::AUX_STAGE_BYTES);
  float* csum_s = bias_s + BN;                                  // EPI_LN: column sums of the gamma-scaled weight
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + BN * (EPI == EPI_LN ? 2 : 1));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;   // [EPI_WARPS][2]
  uint64_t* bias_done = aux_bar + 2 * Cfg::EPI_WARPS;   // [STAGES] (wgrad bias-gradient warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bias_done + STAGES);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler: uniform branches / registers
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (Cfg::HAS_AUX) tma_prefetch_desc(&tmAux);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);      // the leader's two producers arrive (+ 2 CTAs' transaction bytes)
      mbar_init(&empty_bar[s], 1);     // multicast commit from the leader's MMA thread
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);                       // multicast commit
      mbar_init(&tempty_bar[b], 2 * Cfg::EPI_WARPS);     // epilogue warps of BOTH CTAs (leader's copy is used)
    }
    for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    for (int i = 0; i < STAGES; ++i) mbar_init(&bias_done[i], Cfg::BIAS_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;       // tiles_m counts 256-row tiles here
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 || warp == Cfg::PRODUCER2_WARP) {
    // ------------------------------------------------------------------ TMA producers (both CTAs; converged warps):
    // warp 0 loads the A operand of every stage, the last warp the B operand
    const bool load_a = warp == 0;
    {
      uint32_t kiter = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int n_blk = tile % p.tiles_n;
        const int rest = tile / p.tiles_n;
        const int m_blk = rest % p.tiles_m;
        const int split = rest / p.tiles_m;
        const int m0 = m_blk * 2 * BM + rank * BM;        // this CTA's 128 rows
        const int n0 = n_blk * BN + rank * HN;            // this CTA's half of the B tile
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (Cfg::BIASW) mbar_wait(&bias_done[s], ph ^ 1);
          if (elect_one()) {
            const uint32_t full0 = mapa_u32(smem_u32(&full_bar[s]), 0);     // the leader's barrier
            uint8_t* a_dst = stage_base + s * STAGE_BYTES;
            uint8_t* b_dst = a_dst + A_BYTES;
            if (load_a) {
              if (leader) mbar_expect_tx(&full_bar[s], 2 * A_BYTES);        // both CTAs' A tiles
              if (!A_MN) {
                tma_load_2d_2sm(a_dst, &tmA, full0, kb * BK, m0);
              } else {
                tma_load_2d_2sm(a_dst, &tmA, full0, m0, kb * BK);
                tma_load_2d_2sm(a_dst + 8192, &tmA, full0, m0 + 64, kb * BK);
              }
            } else {
              if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES);   // both CTAs' halves of B
              if (!B_MN) {
                tma_load_2d_2sm(b_dst, &tmB, full0, kb * BK, n0);
              } else {
#pragma unroll
                for (int j = 0; j < HN / 64; ++j)
                  tma_load_2d_2sm(b_dst + j * 8192, &tmB, full0, n0 + j * 64, kb * BK);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp walks the loop converged; one elected lane issues (see elect_one()).
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      constexpr uint32_t A_KSTEP = A_MN ? (2048 >> 4) : (32 >> 4), B_KSTEP = B_MN ? (2048 >> 4) : (32 >> 4);
      uint32_t kiter = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const int rest = tile / p.tiles_n;
        const int split = rest / p.tiles_m;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
          const int s = kiter % STAGES;
          const uint32_t ph = (kiter / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
          const uint64_t adesc = A_MN ? umma_smem_desc(a_addr, 8192, 1024) : umma_smem_desc(a_addr, 16, 1024);
          const uint64_t bdesc = B_MN ? umma_smem_desc(b_addr, 8192, 1024) : umma_smem_desc(b_addr, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(d_tmem, adesc + k * A_KSTEP, bdesc + k * B_KSTEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[s], 3);     // both CTAs may refill this stage
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(&tfull_bar[buf], 3);     // accumulator complete in both CTAs' TMEM
        __syncwarp();
      }
    }
  } else if (Cfg::BIASW && warp >= 2 + Cfg::EPI_WARPS && warp < Cfg::PRODUCER2_WARP) {
    // ------------------------------------------------------------------ bias-gradient warps (wgrad, both CTAs)
    bias_grad_warp_loop<STAGES, STAGE_BYTES, Cfg::BIAS_WARPS>(p, stage_base, empty_bar, bias_done, warp - 2 - Cfg::EPI_WARPS, lane,
                                             cluster_id, num_clusters, total_tiles, 2 * BM, static_cast<int>(rank) * BM);
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    const int ew = warp - 2;
    constexpr int G = Cfg::EPI_WARPS / 4;        // column groups: each warp owns BN / G columns of its 32 rows
    const int grp = ew >> 2;
    const int etid = threadIdx.x - 64;
    constexpr int CW = Cfg::CW;
    constexpr int NCHUNK = BN / G / CW;
    constexpr int OUT_BUF = Cfg::OUT_BUF, AUX_BUF = Cfg::AUX_BUF;
    constexpr int ONB = Cfg::OUT_NBUF, ANB = Cfg::AUX_NBUF;
    uint8_t* my_out = epi_out + ew * ONB * OUT_BUF;
    uint8_t* my_aux = epi_aux + ew * ANB * AUX_BUF;
    uint64_t* my_aux_bar = aux_bar + ew * 2;
    const uint32_t sw64 = (lane >> 1) & 3, sw128 = lane & 7;
    uint32_t cc = 0;
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const int n_blk = tile % p.tiles_n;
      const int rest = tile / p.tiles_n;
      const int m_blk = rest % p.tiles_m;
      const int m0 = m_blk * 2 * BM + rank * BM, n0 = n_blk * BN;
      const int buf = it & 1;
      const int r0 = m0 + q * 32;
      const int cbase = grp * (BN / G);

      if (EPI != EPI_F32_ADD && EPI != EPI_DGELU && EPI != EPI_DELTA) {
        named_bar_sync(1, 32 * Cfg::EPI_WARPS);
        for (int i = etid; i < BN; i += 32 * Cfg::EPI_WARPS) {
          float b = 0.f;
          if (p.bias != nullptr && n0 + i < p.N)
            b = p.bias_is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n0 + i])
                               : reinterpret_cast<const float*>(p.bias)[n0 + i];
          bias_s[i] = b;
          if (EPI == EPI_LN) csum_s[i] = n0 + i < p.N ? p.ln_colsum[n0 + i] : 0.f;
        }
        named_bar_sync(1, 32 * Cfg::EPI_WARPS);
      }
      const bool active = r0 < p.M && n0 + cbase < p.N;
      float2 ln_rs2 = make_float2(1.f, 1.f), ln_nm2 = make_float2(0.f, 0.f);      // EPI_LN: this thread's row statistics
      if (EPI == EPI_LN && r0 + lane < p.M) {
        const float rs = p.ln_rstd[r0 + lane], mu = p.ln_mean[r0 + lane];
        ln_rs2 = make_float2(rs, rs);
        ln_nm2 = make_float2(-mu * rs, -mu * rs);
      }
      if (Cfg::AUX_IN && lane == 0 && active) {
        mbar_expect_tx(&my_aux_bar[cc & 1], AUX_BUF);
        tma_load_2d(my_aux + (cc & 1) * AUX_BUF, &tmAux, &my_aux_bar[cc & 1], n0 + cbase, r0);
      }
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * BN + cbase + (static_cast<uint32_t>(q * 32) << 16);

      float dsum = 0.f;      // EPI_DELTA: running sum_d dO * O of this thread's row over the current head
      if (active) {
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = n0 + cbase + c * CW;
          if (col0 >= p.N) break;
          const uint32_t b = cc & 1;
          const uint32_t ob = (ONB == 2) ? b : 0u, ab = (ANB == 2) ? b : 0u;
          if (Cfg::AUX_IN && lane == 0 && c + 1 < NCHUNK && col0 + CW < p.N) {
            mbar_expect_tx(&my_aux_bar[b ^ 1], AUX_BUF);
            tma_load_2d(my_aux + (b ^ 1) * AUX_BUF, &tmAux, &my_aux_bar[b ^ 1], col0 + CW, r0);
          }
          uint32_t v[32];
          tmem_ld32(t_row + c * CW, v);
          tmem_wait_ld();
          if (Cfg::AUX_IN) mbar_wait(&my_aux_bar[b], (cc >> 1) & 1);
          if (EPI == EPI_F32_ADD) {
            if (lane == 0) tma_store_wait_read<ONB - 1>();
            __syncwarp();
            uint8_t* out_row = my_out + ob * OUT_BUF + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(out_row + ((j ^ sw128) << 4)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint8_t* out_row = my_out + ob * OUT_BUF + lane * 64;
            uint8_t* aux_row = my_aux + ab * AUX_BUF + lane * 64;
            // results are formed in registers first, so that waiting for the staging buffer's previous TMA
            // store (single-buffered in the heavy epilogues) overlaps the math
            uint4 o_pk[4], z_pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                f[e] = make_float2(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1]));
              if (EPI == EPI_LN) {          // rstd * acc - (mean * rstd) * colsum, then the folded bias below
                const float4 c0 = *reinterpret_cast<const float4*>(&csum_s[cbase + c * CW + 8 * j]);
                const float4 c1 = *reinterpret_cast<const float4*>(&csum_s[cbase + c * CW + 8 * j + 4]);
                f[0] = __ffma2_rn(f[0], ln_rs2, __fmul2_rn(ln_nm2, make_float2(c0.x, c0.y)));
                f[1] = __ffma2_rn(f[1], ln_rs2, __fmul2_rn(ln_nm2, make_float2(c0.z, c0.w)));
                f[2] = __ffma2_rn(f[2], ln_rs2, __fmul2_rn(ln_nm2, make_float2(c1.x, c1.y)));
                f[3] = __ffma2_rn(f[3], ln_rs2, __fmul2_rn(ln_nm2, make_float2(c1.z, c1.w)));
              }
              if (EPI != EPI_DGELU && EPI != EPI_DELTA) {
                const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[cbase + c * CW + 8 * j]);
                const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[cbase + c * CW + 8 * j + 4]);
                f[0] = __fadd2_rn(f[0], make_float2(b0.x, b0.y));
                f[1] = __fadd2_rn(f[1], make_float2(b0.z, b0.w));
                f[2] = __fadd2_rn(f[2], make_float2(b1.x, b1.y));
                f[3] = __fadd2_rn(f[3], make_float2(b1.z, b1.w));
              }
              const uint32_t sw = (static_cast<uint32_t>(j) ^ sw64) << 4;
              if (EPI == EPI_BIAS_RESIDUAL) {
                const uint4 r = *reinterpret_cast<const uint4*>(aux_row + sw);
                f[0] = __fadd2_rn(f[0], bf16x2_to_f32x2(r.x));
                f[1] = __fadd2_rn(f[1], bf16x2_to_f32x2(r.y));
                f[2] = __fadd2_rn(f[2], bf16x2_to_f32x2(r.z));
                f[3] = __fadd2_rn(f[3], bf16x2_to_f32x2(r.w));
              } else if (EPI == EPI_DELTA) {
                const uint4 r = *reinterpret_cast<const uint4*>(aux_row + sw);
                float2 d2 = __fmul2_rn(f[0], bf16x2_to_f32x2(r.x));
                d2 = __ffma2_rn(f[1], bf16x2_to_f32x2(r.y), d2);
                d2 = __ffma2_rn(f[2], bf16x2_to_f32x2(r.z), d2);
                d2 = __ffma2_rn(f[3], bf16x2_to_f32x2(r.w), d2);
                dsum += d2.x + d2.y;
              } else if (EPI == EPI_DGELU) {
                const uint4 r = *reinterpret_cast<const uint4*>(aux_row + sw);
                // half of the elements on the MUFU form (2 MUFU + 8 FMA-pipe ops), half on the polynomial
                // (17 FMA-pipe ops): the two pipes work side by side
                f[0] = __fmul2_rn(f[0], gelu_erf_grad2(bf16x2_to_f32x2(r.x)));
                f[1] = __fmul2_rn(f[1], gelu_grad_poly2(bf16x2_to_f32x2(r.y)));
                f[2] = __fmul2_rn(f[2], gelu_erf_grad2(bf16x2_to_f32x2(r.z)));
                f[3] = __fmul2_rn(f[3], gelu_grad_poly2(bf16x2_to_f32x2(r.w)));
              } else if (EPI == EPI_BIAS_GELU_AUX) {
                const uint4 zq = make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y),
                                            pack_bf16x2(f[2].x, f[2].y), pack_bf16x2(f[3].x, f[3].y));
                z_pk[j] = zq;
                f[0] = gelu_erf2(bf16x2_to_f32x2(zq.x));      // MUFU form / polynomial alternate (see GELU' above)
                f[1] = gelu_poly2(bf16x2_to_f32x2(zq.y));
                f[2] = gelu_erf2(bf16x2_to_f32x2(zq.z));
                f[3] = gelu_poly2(bf16x2_to_f32x2(zq.w));
              }
              o_pk[j] = make_uint4(pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y), pack_bf16x2(f[2].x, f[2].y),
                                   pack_bf16x2(f[3].x, f[3].y));
            }
            if (lane == 0) tma_store_wait_read<ONB - 1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t sw = (static_cast<uint32_t>(j) ^ sw64) << 4;
              *reinterpret_cast<uint4*>(out_row + sw) = o_pk[j];
              if (EPI == EPI_BIAS_GELU_AUX) *reinterpret_cast<uint4*>(aux_row + sw) = z_pk[j];
            }
          }
          if (EPI == EPI_DELTA) {
            // a head's columns end with this chunk (head_dim 32: every chunk, 64: every second; column groups start on
            // multiples of 128, so heads never straddle warps): store the row's delta for that head
            const int cend = col0 + CW;
            if (cend % p.head_dim == 0) {
              const int r = r0 + lane;
              if (r < p.M) {
                const int bi = r / p.tokens, n = r - bi * p.tokens;
                const int hh = cend / p.head_dim - 1;
                p.delta[(static_cast<long long>(bi) * p.heads + hh) * p.tokens + n] = dsum;
              }
              dsum = 0.f;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (EPI == EPI_F32_ADD) {
              tma_reduce_add_2d(&tmC, my_out + ob * OUT_BUF, col0, r0);
            } else {
              tma_store_2d(&tmC, my_out + ob * OUT_BUF, col0, r0);
              if (EPI == EPI_BIAS_GELU_AUX) tma_store_2d(&tmAux, my_aux + ab * AUX_BUF, col0, r0);
            }
            tma_store_commit();
          }
          ++cc;
        }
      }
      tc_fence_before();
      __syncwarp();
      // the accumulator buffer is free once the epilogue warps of BOTH CTAs are done: arrive on the
      // leader's barrier (the only one the MMA thread waits on)
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[buf]);
        else mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();          // nobody tears TMEM / smem down while the peer may still touch it
  tc_fence_after();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_num_sms[64] = {};
int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = g_num_sms[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                       const CUtensorMap& tAux, const GemmParams& p, int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN, STAGES, A_MN, B_MN, EPI>;
  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN, EPI>;
  static DeviceOnce once;
  bool& attr_set = once.flag();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gemm: cudaFuncSetAttribute(%d B smem): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tA, tB, tC, tAux, p);
  return check_launch("gemm_bf16_kernel");
}

static int g_debug_max_clusters = 0;   // profiling aid (ucf_debug_set_gemm_max_clusters): run on fewer SM pairs

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_gemm2(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                        const CUtensorMap& tAux, const GemmParams& p, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN, STAGES, A_MN, B_MN, EPI>;
  auto kern = gemm2_bf16_kernel<BN, STAGES, A_MN, B_MN, EPI>;
  static DeviceOnce once;
  bool& attr_set = once.flag();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gemm2: cudaFuncSetAttribute(%d B smem): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n * p.splits;
  int clusters = num_sms() / 2;
  if (g_debug_max_clusters > 0 && clusters > g_debug_max_clusters) clusters = g_debug_max_clusters;
  if (clusters > total) clusters = total;
  kern<<<2 * clusters, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tA, tB, tC, tAux, p);   // __cluster_dims__(2,1,1)
  return check_launch("gemm2_bf16_kernel");
}

}  // namespace ucf

using namespace ucf;

/* profiling aid (not part of the public header): limit the CTA-pair kernels to n clusters (0 = all SMs) */
extern "C" void ucf_debug_set_gemm_max_clusters(int n) { ucf::g_debug_max_clusters = n; }

struct DeltaArgs { float* delta; int tokens, heads; const float* ln_mean; const float* ln_rstd; const float* ln_colsum; };
static int gemm_bf16_impl(const void* A, const void* B, void* C, const void* bias, void* aux,
                             int M, int N, int K, long long lda, long long ldb, long long ldc,
                             long long ldaux, int a_layout, int b_layout, int epilogue,
                             int bias_dtype, int splits, int tile_n, void* bias_grad, void* stream,
                             const DeltaArgs* dargs = nullptr);

// ---- profiling aid of bench.py (not part of the public header): CUDA-event timing of every GEMM launch on the
// launching stream, so the roofline figure is measured live inside the timed region wherever the launch comes from
// (Python op wrapper or ucf_block_fwd / ucf_block_bwd)
namespace {
struct GemmTimingRec { double flops; cudaEvent_t e0, e1; };
bool g_gemm_timing = false;
std::vector<GemmTimingRec> g_gemm_recs;
}  // namespace
extern "C" void ucf_debug_gemm_timing(int enable) { g_gemm_timing = enable != 0; }
/* sums the records taken so far (synchronises on their events), then clears them */
extern "C" int ucf_debug_gemm_timing_summary(double* flops, double* ms, long long* launches) {
  double fl = 0.0, t = 0.0;
  for (auto& r : g_gemm_recs) {
    cudaEventSynchronize(r.e1);
    float dt = 0.f;
    if (cudaEventElapsedTime(&dt, r.e0, r.e1) == cudaSuccess) { fl += r.flops; t += dt; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (flops) *flops = fl;
  if (ms) *ms = t;
  if (launches) *launches = static_cast<long long>(g_gemm_recs.size());
  g_gemm_recs.clear();
  return UCF_OK;
}

extern "C" int ucf_gemm_bf16(const void* A, const void* B, void* C, const void* bias, void* aux,
                             int M, int N, int K, long long lda, long long ldb, long long ldc,
                             long long ldaux, int a_layout, int b_layout, int epilogue,
                             int bias_dtype, int splits, int tile_n, void* bias_grad, void* stream) {
  if (!g_gemm_timing)
    return gemm_bf16_impl(A, B, C, bias, aux, M, N, K, lda, ldb, ldc, ldaux, a_layout, b_layout, epilogue, bias_dtype, splits,
                          tile_n, bias_grad, stream);
  GemmTimingRec r;
  r.flops = 2.0 * M * N * K;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaEventRecord(r.e0, st);
  const int rc = gemm_bf16_impl(A, B, C, bias, aux, M, N, K, lda, ldb, ldc, ldaux, a_layout, b_layout, epilogue, bias_dtype,
                                splits, tile_n, bias_grad, stream);
  cudaEventRecord(r.e1, st);
  g_gemm_recs.push_back(r);
  return rc;
}

extern "C" int ucf_ln_gemm_supported(int M, int N, int K) { return M >= 512 && N >= 512 && N <= 4096 && K > 0 && K % 8 == 0; }

extern "C" int ucf_ln_gemm(const void* x, const void* w_gamma, void* y, const float* bias_folded, const float* colsum,
                           const float* mean, const float* rstd, int M, int N, int K, long long ldx, long long ldw, long long ldy,
                           void* stream) {
  if (!bias_folded || !colsum || !mean || !rstd) { set_last_error("ln_gemm: null vector"); return UCF_ERR_BAD_ARG; }
  if (!ucf_ln_gemm_supported(M, N, K)) {
    set_last_error("ln_gemm: shape M=%d N=%d K=%d has no fused kernel (ucf_ln_gemm_supported)", M, N, K);
    return UCF_ERR_UNSUPPORTED;
  }
  DeltaArgs d{nullptr, 0, 0, mean, rstd, colsum};
  GemmTimingRec r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_gemm_timing) {
    r.flops = 2.0 * M * N * K;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
  const int rc = gemm_bf16_impl(x, w_gamma, y, bias_folded, nullptr, M, N, K, ldx, ldw, ldy, 0, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_K_MAJOR,
                                EPI_LN, UCF_DTYPE_F32, 1, 512, nullptr, stream, &d);
  if (g_gemm_timing) { cudaEventRecord(r.e1, st); g_gemm_recs.push_back(r); }
  return rc;
}

extern "C" int ucf_gemm_dgrad_delta_supported(int M, int N, int K, int heads) {
  if (heads <= 0 || N % heads) return 0;
  const int hd = N / heads;
  return (hd == 32 || hd == 64) && N % 128 == 0 && M >= 512 && N >= 512 && N <= 4096 && K > 0;
}

extern "C" int ucf_gemm_dgrad_delta(const void* dY, const void* W, void* dX, const void* O, float* delta, int M, int N, int K,
                                    long long lddy, long long ldw, long long lddx, long long ldo, int tokens, int heads,
                                    void* stream) {
  if (!delta || !O || tokens <= 0 || M % tokens) {
    set_last_error("gemm_dgrad_delta: delta / O null or M not a multiple of tokens"); return UCF_ERR_BAD_ARG;
  }
  if (!ucf_gemm_dgrad_delta_supported(M, N, K, heads)) {
    set_last_error("gemm_dgrad_delta: shape M=%d N=%d heads=%d has no fused kernel (ucf_gemm_dgrad_delta_supported)", M, N, heads);
    return UCF_ERR_UNSUPPORTED;
  }
  DeltaArgs d{delta, tokens, heads, nullptr, nullptr, nullptr};
  GemmTimingRec r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g_gemm_timing) {
    r.flops = 2.0 * M * N * K;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
  const int rc = gemm_bf16_impl(dY, W, dX, nullptr, const_cast<void*>(O), M, N, K, lddy, ldw, lddx, ldo, UCF_LAYOUT_K_MAJOR,
                                UCF_LAYOUT_MN_MAJOR, EPI_DELTA, 0, 1, 512, nullptr, stream, &d);
  if (g_gemm_timing) { cudaEventRecord(r.e1, st); g_gemm_recs.push_back(r); }
  return rc;
}

static int gemm_bf16_impl(const void* A, const void* B, void* C, const void* bias, void* aux,
                             int M, int N, int K, long long lda, long long ldb, long long ldc,
                             long long ldaux, int a_layout, int b_layout, int epilogue,
                             int bias_dtype, int splits, int tile_n, void* bias_grad, void* stream,
                             const DeltaArgs* dargs) {
  if (M <= 0 || N <= 0 || K <= 0) { set_last_error("gemm: empty problem M=%d N=%d K=%d", M, N, K); return UCF_ERR_BAD_ARG; }
  if (!A || !B || !C) { set_last_error("gemm: null operand"); return UCF_ERR_BAD_ARG; }
  if (epilogue < 0 || epilogue > 6 || ((epilogue == EPI_DELTA || epilogue == EPI_LN) && !dargs)) { set_last_error("gemm: bad epilogue %d", epilogue); return UCF_ERR_BAD_ARG; }
  const bool a_mn = a_layout == UCF_LAYOUT_MN_MAJOR, b_mn = b_layout == UCF_LAYOUT_MN_MAJOR;
  const bool has_aux = epilogue == EPI_BIAS_RESIDUAL || epilogue == EPI_BIAS_GELU_AUX || epilogue == EPI_DGELU || epilogue == EPI_DELTA;
  if (has_aux && !aux) { set_last_error("gemm: epilogue %d needs aux", epilogue); return UCF_ERR_BAD_ARG; }
  const int out_es = (epilogue == EPI_F32_ADD) ? 4 : 2;
  if ((lda * 2) % 16 || (ldb * 2) % 16 || (ldc * out_es) % 16 || (has_aux && (ldaux * 2) % 16) ||
      (reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) |
       reinterpret_cast<uintptr_t>(aux)) & 15) {
    set_last_error("gemm: pointers and row pitches must be 16-byte aligned (lda=%lld ldb=%lld ldc=%lld ldaux=%lld)",
                   lda, ldb, ldc, ldaux);
    return UCF_ERR_BAD_ARG;
  }
  if (epilogue != EPI_F32_ADD) splits = 1;
  if (splits < 1) splits = 1;

  // tile_n: 128 / 256 = single-CTA kernels; 512 = CTA-pair kernel (256 x 256 tile per 2-CTA cluster)
  // auto (tile_n == 0): tall problems with several 256-wide column tiles go to the CTA-pair kernel
  // (measured +7..11 % on the ViT-B shapes); everything else to the single-CTA kernels.
  const bool pair_supported =
      (!a_mn && !b_mn && (epilogue == EPI_BIAS || epilogue == EPI_BIAS_RESIDUAL || epilogue == EPI_BIAS_GELU_AUX || epilogue == EPI_LN)) ||
      (!a_mn && b_mn && (epilogue == EPI_BIAS || epilogue == EPI_DGELU || epilogue == EPI_DELTA)) || (a_mn && b_mn && epilogue == EPI_F32_ADD);
  if (tile_n == 0 && epilogue != EPI_F32_ADD && epilogue != EPI_DELTA && epilogue != EPI_LN) {
    // Persistent kernels work through their tiles in rounds of one tile per SM (pair kernel: per 2-SM cluster), so a
    // small problem is decided by how well its tile count fills whole rounds, not by the per-tile efficiency alone:
    // M = 8192 x N = 768 is 96 pair tiles on 74 clusters (2 rounds, 65 % busy) but 384 128x128 tiles on 148 SMs (3 rounds
    // of a quarter of the work each).  Predicted time = rounds x tile area per SM / relative tile efficiency (pair 1.0,
    // 128x256 0.92, 128x128 0.80: measured on the ViT-B shapes, profiles/r01_gemm_pair_vs_single.log).  The wgrad form keeps
    // its own round-filling split-K choice (ucf_wgrad_splits).
    const long long sms = num_sms();
    auto rounds = [](long long tiles, long long workers) { return (tiles + workers - 1) / workers; };
    const long long tm128 = (M + 127) / 128, tm256 = (M + 255) / 256, tn128 = (N + 127) / 128, tn256 = (N + 255) / 256;
    double best = 1e300;
    if (pair_supported && M >= 512 && N >= 512 && N <= 4096) {
      best = static_cast<double>(rounds(tm256 * tn256, sms / 2 > 0 ? sms / 2 : 1)) * (128.0 * 256.0);
      tile_n = 512;
    }
    const double c256 = static_cast<double>(rounds(tm128 * tn256, sms)) * (128.0 * 256.0) / 0.92;
    const double c128 = static_cast<double>(rounds(tm128 * tn128, sms)) * (128.0 * 128.0) / 0.80;
    if (c256 < best * 0.97) { best = c256; tile_n = 256; }
    if (c128 < best * 0.97) { best = c128; tile_n = 128; }
  } else if (tile_n == 0 && pair_supported && M >= 512 && N >= 512 && N <= 4096) {
    tile_n = 512;
  }
  const bool pair = tile_n == 512;
  int BN = pair ? 256 : tile_n;
  if (BN != 128 && BN != 256) BN = (N <= 128 || (N % 256 != 0 && N % 128 == 0 && N < 1024)) ? 128 : 256;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = pair ? (M + 255) / 256 : (M + 127) / 128;
  p.tiles_n = (N + BN - 1) / BN;
  p.kb_total = (K + 63) / 64;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.bias = bias;
  p.bias_is_bf16 = bias_dtype == UCF_DTYPE_BF16;
  p.bias_grad = static_cast<float*>(bias_grad);
  p.delta = nullptr; p.tokens = 1; p.heads = 1; p.head_dim = 64;
  p.ln_mean = p.ln_rstd = p.ln_colsum = nullptr;
  if (epilogue == EPI_DELTA) {
    p.delta = dargs->delta; p.tokens = dargs->tokens; p.heads = dargs->heads; p.head_dim = N / dargs->heads;
  }
  if (epilogue == EPI_LN) { p.ln_mean = dargs->ln_mean; p.ln_rstd = dargs->ln_rstd; p.ln_colsum = dargs->ln_colsum; }
  if (bias_grad && !(a_mn && epilogue == EPI_F32_ADD)) {
    set_last_error("gemm: bias_grad is only produced by the wgrad form (A MN-major, UCF_EPI_F32_ADD)");
    return UCF_ERR_BAD_ARG;
  }

  CUtensorMap tA, tB, tC, tAux;
  memset(&tAux, 0, sizeof(tAux));
  int rc;
  {
    // K-major: dims {K, M}, box {64, 128}.  MN-major (stored [K, M]): dims {M, K}, box {64, 64}.
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; box[0] = 64; box[1] = 128; }
    else       { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = 64; }
    strides[0] = static_cast<uint64_t>(lda) * 2;
    if ((rc = make_tmap(&tA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if (!b_mn) { dims[0] = K; dims[1] = N; box[0] = 64; box[1] = pair ? BN / 2 : BN; }
    else       { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = 64; }
    strides[0] = static_cast<uint64_t>(ldb) * 2;
    if ((rc = make_tmap(&tB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    // epilogue boxes: 32 rows x 32 columns (fp32: 128-byte rows, bf16: 64-byte rows)
    dims[0] = N; dims[1] = M; box[0] = 32; box[1] = 32;
    if (epilogue == EPI_F32_ADD) {
      strides[0] = static_cast<uint64_t>(ldc) * 4;
      if ((rc = make_tmap(&tC, C, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    } else {
      strides[0] = static_cast<uint64_t>(ldc) * 2;
      if ((rc = make_tmap(&tC, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    if (has_aux) {
      strides[0] = static_cast<uint64_t>(ldaux) * 2;
      if ((rc = make_tmap(&tAux, aux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

#define UCF_GEMM2_CASE(stg, amn, bmn, epi)                                            \
  if (pair && a_mn == amn && b_mn == bmn && epilogue == epi)                          \
    return launch_gemm2<256, stg, amn, bmn, epi>(tA, tB, tC, tAux, p, st);
  UCF_GEMM2_CASE(6, false, false, EPI_BIAS)
  UCF_GEMM2_CASE(5, false, false, EPI_BIAS_RESIDUAL)
  UCF_GEMM2_CASE(5, false, false, EPI_BIAS_GELU_AUX)
  UCF_GEMM2_CASE(6, false, true, EPI_BIAS)
  UCF_GEMM2_CASE(4, false, true, EPI_DGELU)
  UCF_GEMM2_CASE(5, false, true, EPI_DELTA)
  UCF_GEMM2_CASE(5, false, false, EPI_LN)
  UCF_GEMM2_CASE(6, true, true, EPI_F32_ADD)
#undef UCF_GEMM2_CASE
  if (pair) {
    set_last_error("gemm: no CTA-pair kernel for a_layout=%d b_layout=%d epilogue=%d", a_layout, b_layout, epilogue);
    return UCF_ERR_UNSUPPORTED;
  }

#define UCF_GEMM_CASE(bn, stg, amn, bmn, epi)                                        \
  if (BN == bn && a_mn == amn && b_mn == bmn && epilogue == epi)                     \
    return launch_gemm<bn, stg, amn, bmn, epi>(tA, tB, tC, tAux, p, 0, st);

  // forward
  UCF_GEMM_CASE(256, 4, false, false, EPI_BIAS)
  UCF_GEMM_CASE(128, 6, false, false, EPI_BIAS)
  UCF_GEMM_CASE(256, 3, false, false, EPI_BIAS_RESIDUAL)
  UCF_GEMM_CASE(128, 5, false, false, EPI_BIAS_RESIDUAL)
  UCF_GEMM_CASE(256, 3, false, false, EPI_BIAS_GELU_AUX)
  UCF_GEMM_CASE(128, 5, false, false, EPI_BIAS_GELU_AUX)
  UCF_GEMM_CASE(256, 4, false, false, EPI_F32_ADD)
  UCF_GEMM_CASE(128, 6, false, false, EPI_F32_ADD)
  // dgrad
  UCF_GEMM_CASE(256, 4, false, true, EPI_BIAS)
  UCF_GEMM_CASE(128, 6, false, true, EPI_BIAS)
  UCF_GEMM_CASE(256, 3, false, true, EPI_DGELU)
  UCF_GEMM_CASE(128, 5, false, true, EPI_DGELU)
  UCF_GEMM_CASE(256, 3, false, true, EPI_BIAS_RESIDUAL)
  UCF_GEMM_CASE(128, 5, false, true, EPI_BIAS_RESIDUAL)
  // wgrad
  UCF_GEMM_CASE(256, 4, true, true, EPI_F32_ADD)
  UCF_GEMM_CASE(128, 6, true, true, EPI_F32_ADD)
  UCF_GEMM_CASE(256, 4, true, true, EPI_BIAS)
  UCF_GEMM_CASE(128, 6, true, true, EPI_BIAS)
  UCF_GEMM_CASE(128, 6, true, false, EPI_BIAS)
  UCF_GEMM_CASE(128, 6, true, false, EPI_F32_ADD)
#undef UCF_GEMM_CASE
  set_last_error("gemm: no kernel for a_layout=%d b_layout=%d epilogue=%d tile_n=%d", a_layout, b_layout, epilogue, BN);
  return UCF_ERR_UNSUPPORTED;
}
