// Fused softmax attention forward on tcgen05 (sm_100a):  O = softmax(Q K^T * scale) V.
// Replaces the FLASH / CK / DEFAULT / NONE branches of Attention.forward
// (/root/reference/src/UCF_VIT/simple/building_blocks.py:163-187): non-causal, no mask, dropout 0.
//
// Persistent, one CTA per SM, work item = (batch, head, pair of 128-query tiles).  Two softmax
// warpgroups ping-pong on the two query tiles of the item while sharing its K/V tiles:
//   warp 0        TMA producer: Q pair (double-buffered across items), K_j / V_j tiles (128 keys)
//                 through a 4-slot ring -- it runs ahead into the next item, so loads, barrier
//                 set-up and TMEM allocation are paid once per CTA, not once per (b, h)
//   warp 1        tcgen05.mma issuer: S_g = Q_g K_j^T -> TMEM, PV_g = P_g V_j -> TMEM; the MMAs of
//                 one warpgroup overlap the softmax of the other
//   warps 2..5    softmax warpgroup 0 (query tile 0), thread == query row
//   warps 6..9    softmax warpgroup 1 (query tile 1)
// Online max/sum live in registers (exp2 domain); P_j is written as bf16 into 128B-swizzled
// shared memory (the A operand of the PV MMA); O accumulates in registers and is rescaled when the
// running max moves.  q/k/v are read in place from the packed qkv projection through 4-D tensor
// maps {hd, H, N, B}; rows past N are zero-filled by TMA and masked here.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

struct AttnFwdParams {
  int B, H, Nq, Nk;
  int nqp;            // query-tile pairs per (b, h)
  int items;          // B * H * nqp
  float scale_log2;   // scale * log2(e)
  float* lse;         // [B, H, Nq]
  long long* timeline;   // optional clock64 stamps of CTA 0 (profiling aid; NULL in production)
  int stagger_cycles;    // short-key kernel: start offset of warpgroup 1
};

template <int HD>
struct AttnFwdCfg {
  static constexpr int BQ = 128, BKV = 128;
  static constexpr int Q_BYTES = BQ * HD * 2;          // one query tile
  static constexpr int KV_BYTES = BKV * HD * 2;
  static constexpr int P_BYTES = BQ * BKV * 2;
  static constexpr int RING = 4;
  static constexpr int NBAR = 2 + 2 + 2 * RING + 8;
  static constexpr int SMEM_BYTES = 1024 + 4 * Q_BYTES + RING * KV_BYTES + 2 * P_BYTES + NBAR * 8 + 16;
  static constexpr int TMEM_COLS = 512;                // WG g: S at g*256, PV at g*256+128
  static constexpr int ROW_BYTES = HD * 2;             // 128 (HD=64) or 64 (HD=32)
  static constexpr int ATOM_BYTES = 8 * ROW_BYTES;     // swizzle atom: 8 rows
  static constexpr int THREADS = 64 + 256;
  static_assert(SMEM_BYTES <= 232448, "smem");
};

// shared-memory descriptor for a tile whose rows are ROW_BYTES wide (128 -> SWIZZLE_128B,
// 64 -> SWIZZLE_64B).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t attn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = umma_smem_desc(saddr, lbo, sbo);
  if (ROW_BYTES == 64) d = (d & ~(7ull << 61)) | (4ull << 61);   // SWIZZLE_64B
  return d;
}

template <int HD>
__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV, RING = Cfg::RING;
  constexpr int RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic)
  uint8_t* q_s = smem;                                  // [2 buffers][2 tiles]
  uint8_t* kv_s = q_s + 4 * Cfg::Q_BYTES;               // [RING]
  uint8_t* p_s = kv_s + RING * Cfg::KV_BYTES;           // [2 warpgroups]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = bars + 2;            // [2]
  uint64_t* kv_full = bars + 4;            // [RING]
  uint64_t* kv_empty = bars + 4 + RING;    // [RING]
  uint64_t* s_full = bars + 4 + 2 * RING;  // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_full + 4;           // [2]
  uint64_t* pv_full = s_full + 6;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  // warp index through a shuffle: warp-uniform for the compiler, so the producer / MMA warps run converged and
  // keep TMA / UMMA descriptors in uniform registers (no per-instruction ELECT / R2UR waterfall)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nkv = (p.Nk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < RING; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_empty[g], 4);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_full[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (b, h, first query row); does warpgroup 1 have any valid rows?
  auto decode = [&](int item, int& b, int& h, int& q0) {
    const int qp = item % p.nqp;
    const int bh = item / p.nqp;
    h = bh % p.H;
    b = bh / p.H;
    q0 = qp * 2 * BQ;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane issues)
    {
      uint32_t it = 0, r = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int b, h, q0;
        decode(item, b, h, q0);
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&q_full[qb], 2 * Cfg::Q_BYTES);
          tma_load_4d(q_s + (qb * 2 + 0) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0, b);
          tma_load_4d(q_s + (qb * 2 + 1) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0 + BQ, b);
        }
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
          for (int t = 0; t < 2; ++t, ++r) {
            const int slot = r % RING;
            mbar_wait(&kv_empty[slot], ((r / RING) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&kv_full[slot], Cfg::KV_BYTES);
              tma_load_4d(kv_s + slot * Cfg::KV_BYTES, t == 0 ? &tmK : &tmV, &kv_full[slot], 0, h, j * BKV, b);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, false, true);
      const uint32_t p_addr0 = smem_u32(p_s);
      uint32_t it = 0, r = 0;        // item counter, ring counter (K and V tiles alternate)
      uint32_t tc[2] = {0, 0};       // per-warpgroup processed-tile counters (barrier parities)
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int b, h, q0;
        decode(item, b, h, q0);
        const int ng = (q0 + BQ < p.Nq) ? 2 : 1;      // warpgroup 1 idles when its tile is past Nq
        const int qb = it & 1;
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        const uint32_t q_addr = smem_u32(q_s + qb * 2 * Cfg::Q_BYTES);

        auto issue_s = [&](int g, uint32_t k_addr) {
          mbar_wait(&s_empty[g], (tc[g] & 1) ^ 1);
          tc_fence_after();
          if (p.timeline && blockIdx.x == 0 && tc[g] < 4) p.timeline[(g * 4 + tc[g]) * 8 + 0] = clock64();   // S issue
          const uint32_t d = tmem_base + g * 256;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              umma_bf16(d, attn_desc<RB>(q_addr + g * Cfg::Q_BYTES + k * 32, 16, AB), attn_desc<RB>(k_addr + k * 32, 16, AB),
                        idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[g]);
          }
          __syncwarp();
        };
        auto issue_pv = [&](int g, uint32_t v_addr) {
          mbar_wait(&p_full[g], tc[g] & 1);
          tc_fence_after();
          if (p.timeline && blockIdx.x == 0 && tc[g] < 4) p.timeline[(g * 4 + tc[g]) * 8 + 1] = clock64();   // PV issue
          const uint32_t d = tmem_base + g * 256 + 128;
          const uint32_t pa = p_addr0 + g * Cfg::P_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk)
              umma_bf16(d, umma_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        attn_desc<RB>(v_addr + kk * 2 * AB, 0, AB), idesc_pv, kk > 0 ? 1u : 0u);
            umma_commit(&pv_full[g]);
          }
          __syncwarp();
          ++tc[g];
        };

        // tile 0: both S products
        {
          const int sk = r % RING;
          mbar_wait(&kv_full[sk], (r / RING) & 1);
          const uint32_t k_addr = smem_u32(kv_s + sk * Cfg::KV_BYTES);
          for (int g = 0; g < ng; ++g) issue_s(g, k_addr);
          if (elect_one()) umma_commit(&kv_empty[sk]);
          __syncwarp();
          ++r;
        }
        for (int j = 0; j < nkv; ++j) {
          const int sv = r % RING;
          mbar_wait(&kv_full[sv], (r / RING) & 1);
          const uint32_t v_addr = smem_u32(kv_s + sv * Cfg::KV_BYTES);
          ++r;
          const bool more = j + 1 < nkv;
          int sk = 0;
          uint32_t k_addr = 0;
          if (more) {
            sk = r % RING;
            mbar_wait(&kv_full[sk], (r / RING) & 1);
            k_addr = smem_u32(kv_s + sk * Cfg::KV_BYTES);
            ++r;
          }
          for (int g = 0; g < ng; ++g) {
            issue_pv(g, v_addr);                 // waits for P_g(j)
            if (more) issue_s(g, k_addr);        // S_g(j+1): overlaps the other warpgroup's softmax
          }
          if (elect_one()) {
            umma_commit(&kv_empty[sv]);
            if (more) umma_commit(&kv_empty[sk]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&q_empty[qb]);               // every S product of this item has been issued
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 256, tmem_pv = tmem_s + 128;
    const uint32_t row_sw = row & 7, lrow_sw = lane & 7;
    uint8_t* p_row = p_s + g * Cfg::P_BYTES + row * 128;
    uint8_t* o_stage = p_s + g * Cfg::P_BYTES + qd * 4096;   // O staging reuses this warpgroup's P buffer
    uint32_t tc = 0;

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int b, h, q0;
      decode(item, b, h, q0);
      const int qt0 = q0 + g * BQ;                  // first query row of this warpgroup's tile
      if (qt0 >= p.Nq) continue;                    // (only warpgroup 1 can be idle)
      float2 o_acc[HD / 2];
#pragma unroll
      for (int i = 0; i < HD / 2; ++i) o_acc[i] = make_float2(0.f, 0.f);
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
      // the previous item's O tile may still be leaving the staging area (== P buffer)
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();

      for (int j = 0; j < nkv; ++j, ++tc) {
        mbar_wait(&s_full[g], tc & 1);
        tc_fence_after();
        const bool stamp = p.timeline && blockIdx.x == 0 && tc < 4 && (threadIdx.x & 127) == 64;
        if (stamp) p.timeline[(g * 4 + tc) * 8 + 2] = clock64();   // S visible
        const int kbase = j * BKV;
        const int nvalid = min(BKV, p.Nk - kbase);          // keys of this tile that exist
        const int nchunk = (nvalid + 31) >> 5;
        // pass 1: row maximum
        float m_tile = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_s + lane_addr + c * 32, v);
          tmem_wait_ld();
          if ((c + 1) * 32 <= nvalid) {
#pragma unroll
            for (int e = 0; e < 32; e += 2)      // 3-input max (FMNMX3)
              m_tile = fmaxf(m_tile, fmaxf(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e < nvalid) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
          }
        }
        if (stamp) p.timeline[(g * 4 + tc) * 8 + 3] = clock64();   // pass 1 done
        const float m_new = fmaxf(m_run, m_tile * p.scale_log2);
        const float alpha = exp2f(m_run - m_new);            // first tile: exp2(-inf) = 0
        if (j > 0) {                                          // fold PV_{j-1}; also frees the P buffer
          mbar_wait(&pv_full[g], (tc - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_pv + lane_addr + c * 32, v);
            tmem_wait_ld();
            const float2 ap2 = mk2(alpha_prev);
#pragma unroll
            for (int e = 0; e < 16; ++e)
              o_acc[c * 16 + e] = __ffma2_rn(o_acc[c * 16 + e], ap2,
                                             make_float2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])));
          }
        }
        // pass 2: probabilities -> bf16 -> swizzled smem
        float2 l2 = make_float2(0.f, 0.f);
        const float2 sc2 = mk2(p.scale_log2), nm2 = mk2(-m_new);
#pragma unroll 1
        for (int c = 0; c < BKV / 32; ++c) {
          uint32_t pk[16];
          if (c < nchunk) {
            uint32_t v[32];
            tmem_ld32(tmem_s + lane_addr + c * 32, v);
            tmem_wait_ld();
            const bool full = (c + 1) * 32 <= nvalid;
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, nm2);
              float2 pp = make_float2(fast_ex2(t.x), fast_ex2(t.y));
              if (!full) {
                if (c * 32 + e >= nvalid) pp.x = 0.f;
                if (c * 32 + e + 1 >= nvalid) pp.y = 0.f;
              }
              l2 = __fadd2_rn(l2, pp);
              pk[e >> 1] = pack_bf16x2(pp.x, pp.y);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          }
          uint8_t* blk = p_row + (c >> 1) * 16384;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
            *reinterpret_cast<uint4*>(blk + ((chunk ^ row_sw) << 4)) =
                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive_warp(&s_empty[g]);        // S_g may be overwritten by the next QK^T
        fence_proxy_async_smem();
        mbar_arrive_warp(&p_full[g]);         // P_g(j) visible to the tensor core
        if (stamp) p.timeline[(g * 4 + tc) * 8 + 4] = clock64();   // P written
        l_run = fmaf(l_run, alpha, l2.x + l2.y);
        m_run = m_new;
        alpha_prev = alpha;
      }
      // last PV of the item
      mbar_wait(&pv_full[g], (tc - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_pv + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 r = __fmul2_rn(__ffma2_rn(o_acc[c * 16 + q4 * 4 + e], mk2(alpha_prev),
                                                   make_float2(__uint_as_float(v[q4 * 8 + 2 * e]), __uint_as_float(v[q4 * 8 + 2 * e + 1]))),
                                        mk2(inv_l));
            f[2 * e] = r.x; f[2 * e + 1] = r.y;
          }
          const uint32_t chunk = static_cast<uint32_t>(c * 4 + q4);
          uint8_t* dst = (RB == 128) ? o_stage + lane * 128 + ((chunk ^ lrow_sw) << 4)
                                     : o_stage + lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                      pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
      // NOTE: the PV accumulator has been read; the next S/PV of this warpgroup cannot be issued
      // before this warpgroup's next s_empty / p_full arrivals, so TMEM reuse is ordered.
      tc_fence_before();
      if (qt0 + row < p.Nq)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qt0 + row] = (m_run + log2f(l_run)) * 0.69314718055994531f;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && qt0 + qd * 32 < p.Nq) {
        tma_store_4d(&tmO, o_stage, 0, h, qt0 + qd * 32, b);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// profiling aid of the short-key kernel: clock64 stamps of CTA 0's first 4 items, 16 slots per (item, warpgroup)
#define UCF_FTL(k_, g_, idx_)                                                                          \
  do {                                                                                                 \
    if (p.timeline && blockIdx.x == 0 && (k_) < 4) p.timeline[((k_) * 2 + (g_)) * 16 + (idx_)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Short-key schedule (Nk <= 256: at most two key tiles -- ViT-B/16 at 224 px has 197 tokens).
// With every score of a row available at once the softmax is exact in one sweep: S0 = Q K0^T and
// S1 = Q K1^T go to tensor memory up front (2 x 128 columns per warpgroup), the row maximum is taken
// over both, P0 / P1 are exponentiated against the FINAL maximum and O = P0 V0 + P1 V1 accumulates in
// tensor memory (aliasing S0's columns) -- no running maximum, no rescaling, no O accumulator in
// registers.  Each warpgroup has its OWN MMA-issuing warp (warps 1 and 10) walking an independent
// command stream (S -> PV0 -> PV1 per item) with blocking waits, so the warpgroups are free to run out
// of phase and share the MUFU pipe instead of colliding on it.  The tail key tile is trimmed to
// a multiple of 16 keys in both the S1 and the PV1 products.
// Same shared-memory layout, barriers and producer as the general kernel (the 4-slot ring holds exactly
// one item's K0, V0, K1, V1); kv_empty / q_empty count two arrivals (one commit per stream).
template <int HD>
__global__ void __launch_bounds__(352, 1)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV, RING = Cfg::RING;
  constexpr int RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                  // [2 buffers][2 tiles]
  uint8_t* kv_s = q_s + 4 * Cfg::Q_BYTES;               // [RING]
  uint8_t* p_s = kv_s + RING * Cfg::KV_BYTES;           // [2 warpgroups]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = bars + 2;            // [2]
  uint64_t* kv_full = bars + 4;            // [RING]
  uint64_t* kv_empty = bars + 4 + RING;    // [RING]
  uint64_t* s_full = bars + 4 + 2 * RING;  // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_full + 4;           // [2]
  uint64_t* pv_full = s_full + 6;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nkv = (p.Nk + BKV - 1) / BKV;               // 1 or 2
  const int nvalid1 = p.Nk - BKV;                       // valid keys of the tail tile (nkv == 2)
  const int n1 = (nvalid1 + 15) & ~15;                  // ... rounded up to an MMA K / N step

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2); }
    for (int i = 0; i < RING; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_empty[g], 4);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_full[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& b, int& h, int& q0) {
    const int qp = item % p.nqp;
    const int bh = item / p.nqp;
    h = bh % p.H;
    b = bh / p.H;
    q0 = qp * 2 * BQ;
  };
  const int my_items = (p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp)
    uint32_t r = 0;
    for (int k = 0; k < my_items; ++k) {
      int b, h, q0;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), b, h, q0);
      const int qb = k & 1;
      mbar_wait(&q_empty[qb], ((k >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[qb], 2 * Cfg::Q_BYTES);
        tma_load_4d(q_s + (qb * 2 + 0) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0, b);
        tma_load_4d(q_s + (qb * 2 + 1) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0 + BQ, b);
      }
      __syncwarp();
      // K tiles first (both S products are issued up front), then the V tiles
      for (int t = 0; t < 2 * nkv; ++t, ++r) {
        const int slot = r % RING;
        const int j = (nkv == 2) ? (t & 1) : 0;          // order: K0, K1, V0, V1
        const bool is_v = (nkv == 2) ? (t >= 2) : (t == 1);
        mbar_wait(&kv_empty[slot], ((r / RING) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[slot], Cfg::KV_BYTES);
          tma_load_4d(kv_s + slot * Cfg::KV_BYTES, is_v ? &tmV : &tmK, &kv_full[slot], 0, h, j * BKV, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------------ MMA issuers: one warp per warpgroup stream
    // (S -> PV0 -> PV1 per item, blocking waits; the two streams never wait for each other)
    const int g = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, false, true);
    const uint32_t idesc_s1 = umma_idesc_bf16(BQ, n1 > 0 ? n1 : 16, false, false);
    const uint32_t pa = smem_u32(p_s) + g * Cfg::P_BYTES;
    const uint32_t d_s = tmem_base + g * 256;
    uint32_t nvalid_items = 0;     // s_empty parity
    uint32_t pc = 0;               // p_full parity
    for (int k = 0; k < my_items; ++k) {
      int b, h, q0;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), b, h, q0);
      const bool valid = q0 + g * BQ < p.Nq;
      const int qb = k & 1;
      const uint32_t r0 = static_cast<uint32_t>(k) * 2 * nkv;          // ring position of this item's first tile
      const uint32_t rk0 = r0, rk1 = r0 + 1, rv0 = r0 + nkv, rv1 = r0 + 3;   // producer order K0, K1, V0, V1 (or K0, V0)
      mbar_wait(&q_full[qb], (k >> 1) & 1);
      mbar_wait(&kv_full[rk0 % RING], (rk0 / RING) & 1);
      if (nkv == 2) mbar_wait(&kv_full[rk1 % RING], (rk1 / RING) & 1);
      if (valid) mbar_wait(&s_empty[g], (nvalid_items & 1) ^ 1);
      tc_fence_after();
      UCF_FTL(k, g, 8);
      if (elect_one()) {
        if (valid) {
          const uint32_t q_addr = smem_u32(q_s + (qb * 2 + g) * Cfg::Q_BYTES);
          const uint32_t k0_addr = smem_u32(kv_s + (rk0 % RING) * Cfg::KV_BYTES);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(d_s, attn_desc<RB>(q_addr + kk * 32, 16, AB), attn_desc<RB>(k0_addr + kk * 32, 16, AB), idesc_s,
                      kk > 0 ? 1u : 0u);
          if (nkv == 2) {
            const uint32_t k1_addr = smem_u32(kv_s + (rk1 % RING) * Cfg::KV_BYTES);
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk)
              umma_bf16(d_s + 128, attn_desc<RB>(q_addr + kk * 32, 16, AB), attn_desc<RB>(k1_addr + kk * 32, 16, AB),
                        idesc_s1, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[g]);
        }
        umma_commit(&kv_empty[rk0 % RING]);
        if (nkv == 2) umma_commit(&kv_empty[rk1 % RING]);
        umma_commit(&q_empty[qb]);
      }
      __syncwarp();
      if (valid) ++nvalid_items;
      for (int j = 0; j < nkv; ++j) {
        const uint32_t rv = j == 0 ? rv0 : rv1;
        // (an idle warpgroup's stream still waits for the tile before releasing it: its release must not
        // run a whole item ahead of the other stream's)
        mbar_wait(&kv_full[rv % RING], (rv / RING) & 1);
        if (valid) {
          mbar_wait(&p_full[g], pc & 1);
          ++pc;
          tc_fence_after();
          UCF_FTL(k, g, j == 0 ? 9 : 10);
        }
        if (elect_one()) {
          if (valid) {
            const uint32_t v_addr = smem_u32(kv_s + (rv % RING) * Cfg::KV_BYTES);
            const int ksteps = j == 0 ? BKV / 16 : n1 / 16;
            for (int kk = 0; kk < ksteps; ++kk)
              umma_bf16(d_s, umma_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        attn_desc<RB>(v_addr + kk * 2 * AB, 0, AB), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&pv_full[g]);
          }
          umma_commit(&kv_empty[rv % RING]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 256;          // S0 at +0, S1 at +128, O aliases S0
    const uint32_t row_sw = row & 7, lrow_sw = lane & 7;
    uint8_t* p_row = p_s + g * Cfg::P_BYTES + row * 128;
    uint8_t* o_stage = p_s + g * Cfg::P_BYTES + qd * 4096;
    uint32_t tcnt = 0;       // valid items processed (s_full parity)
    uint32_t pcnt = 0;       // P tiles published (pv_full parity)
    const int nchunk1 = nkv == 2 ? (nvalid1 + 31) >> 5 : 0;
    const float2 sc2 = mk2(p.scale_log2);

    for (int k = 0; k < my_items; ++k) {
      int b, h, q0;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), b, h, q0);
      const int qt0 = q0 + g * BQ;
      if (qt0 >= p.Nq) continue;
      const bool st = (threadIdx.x & 127) == 64;
      if (st) UCF_FTL(k, g, 7);
      mbar_wait(&s_full[g], tcnt & 1);
      tc_fence_after();
      // Warpgroup 1 holds its first tile back by half an item: the two warpgroups then alternate on the
      // MUFU pipe (one runs its exponentials while the other waits for S / takes the row maximum / writes
      // O) instead of halving each other's rate; nothing couples the two streams, so the offset persists.
      if (g == 1 && tcnt == 0 && my_items > 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.stagger_cycles) { }
      }
      if (st) UCF_FTL(k, g, 0);
      // ---- pass 1: row maximum over every valid key
      float m_row = -INFINITY;
      const int nv0 = min(BKV, p.Nk);
      auto max_chunk = [&](const uint32_t (&v)[32], int nv) {
        if (nv >= 32) {
#pragma unroll
          for (int e = 0; e < 32; e += 2)
            m_row = fmaxf(m_row, fmaxf(__uint_as_float(v[e]), __uint_as_float(v[e + 1])));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e < nv) m_row = fmaxf(m_row, __uint_as_float(v[e]));
        }
      };
      // two tensor-memory loads in flight per wait
#pragma unroll 1
      for (int c = 0; c < 4 + nchunk1; c += 2) {
        const int nva = c < 4 ? nv0 - c * 32 : nvalid1 - (c - 4) * 32;            // valid columns left in chunk c
        const int nvb = c + 1 < 4 ? nv0 - (c + 1) * 32 : (c + 1 < 4 + nchunk1 ? nvalid1 - (c + 1 - 4) * 32 : 0);
        uint32_t va[32], vb[32];
        if (nva > 0) tmem_ld32(tmem_s + lane_addr + c * 32, va);
        if (nvb > 0) tmem_ld32(tmem_s + lane_addr + (c + 1) * 32, vb);
        tmem_wait_ld();
        if (nva > 0) max_chunk(va, nva);
        if (nvb > 0) max_chunk(vb, nvb);
      }
      if (st) UCF_FTL(k, g, 1);
      const float m2 = m_row * p.scale_log2;
      const float2 nm2 = mk2(-m2);
      float2 l2 = make_float2(0.f, 0.f);
      // one 32-column chunk of probabilities -> 16 packed bf16 pairs (masked past `nv` valid columns)
      auto exp_chunk = [&](int c, int nv, uint32_t (&pk)[16]) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, nm2);
          float2 pp = make_float2(fast_ex2(t.x), fast_ex2(t.y));
          if (nv < 32) {
            if (e >= nv) pp.x = 0.f;
            if (e + 1 >= nv) pp.y = 0.f;
          }
          l2 = __fadd2_rn(l2, pp);
          pk[e >> 1] = pack_bf16x2(pp.x, pp.y);
        }
      };
      auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {      // chunk c (0..3) of the P tile
        uint8_t* blk = p_row + (c >> 1) * 16384;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
          *reinterpret_cast<uint4*>(blk + ((chunk ^ row_sw) << 4)) =
              make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      };
      // ---- pass 2a: P0   (the previous item's O tile must have left the staging area == P buffer)
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        const int nv = nv0 - c * 32;
        if (nv > 0) exp_chunk(c, nv, pk);
        else {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = 0u;
        }
        store_chunk(c, pk);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive_warp(&p_full[g]);           // P0 visible; S0's columns may now receive O = P0 V0
      if (st) UCF_FTL(k, g, 2);
      // ---- pass 2b: P1 is formed in registers while PV0 still reads the P buffer
      if (nkv == 2) {
        uint32_t pk1[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int nv = nvalid1 - c * 32;
          if (c < nchunk1) exp_chunk(4 + c, nv, pk1[c]);
        }
        if (st) UCF_FTL(k, g, 11);
        mbar_wait(&pv_full[g], pcnt & 1);       // PV0 retired: the P buffer is free
        ++pcnt;
        if (st) UCF_FTL(k, g, 3);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nchunk1) store_chunk(c, pk1[c]);
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive_warp(&p_full[g]);         // P1 visible
        if (st) UCF_FTL(k, g, 4);
      }
      // ---- epilogue: O / l
      mbar_wait(&pv_full[g], pcnt & 1);
      ++pcnt;
      tc_fence_after();
      if (st) UCF_FTL(k, g, 5);
      const float l_row = l2.x + l2.y;
      const float inv_l = 1.0f / l_row;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q4 * 8 + e]) * inv_l;
          const uint32_t chunk = static_cast<uint32_t>(c * 4 + q4);
          uint8_t* dst = (RB == 128) ? o_stage + lane * 128 + ((chunk ^ lrow_sw) << 4)
                                     : o_stage + lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                      pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&s_empty[g]);          // S0 / S1 / O of this warpgroup are free for the next item
      if (qt0 + row < p.Nq)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qt0 + row] = (m2 + log2f(l_row)) * 0.69314718055994531f;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && qt0 + qd * 32 < p.Nq) {
        tma_store_4d(&tmO, o_stage, 0, h, qt0 + qd * 32, b);
        tma_store_commit();
      }
      if (st) UCF_FTL(k, g, 6);
      ++tcnt;
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ---------------------------------------------------------------------------------------------
// Single-sweep schedule (round 2; default for every shape).
//
// What bounded the two kernels above was tensor-memory READ bandwidth (~64-85 B/clk/SM): the general kernel
// read 320 columns per row and key tile (S for the row maximum, S again for the exponentials, the PV
// partial product), the short-key kernel 480 per item.  Here S is read ONCE and O never leaves tensor
// memory until the item's epilogue:
//   * optimistic stabiliser: the exponentials of tile j are taken against `m_used`, a value that was a row
//     maximum at some earlier point (tile 0: the maximum of the first 32 keys).  fp32 / bf16 carry 8 exponent
//     bits, so P = exp2(s - m_used) stays exact in relative terms as long as s - m_used < ~100; the true row
//     maximum is tracked on the side (one FMNMX3 per pair) and only when it exceeds m_used by more than
//     kRescaleThreshold (= 60, i.e. a factor 2^60) does the warp take the slow path: O (in tensor memory) and
//     l are multiplied by exp2(m_used - m_new) and the tile is swept again.  The result is the exact softmax
//     either way (lse = m_used + log2 l); the slow path is a correctness net for pathological rows
//     (tests/test_gpu_kernels.py::test_attention_fwd_stabiliser_jumps drives it).
//   * O accumulates in tensor memory over the key tiles (PV with accumulate), no per-tile read-back, no
//     rescaling in the common case, no O registers.
//   * the MMA warp walks ONE continuous stream of key tiles across work items: S_g(next tile) -- also the
//     first tile of the NEXT item -- is issued right behind PV_g(this tile), so each warpgroup's
//     softmax -> PV -> S chain never drains at an item boundary, and warpgroup 1 is started half a
//     softmax period late once (the offset persists: nothing couples the two chains) so that one
//     warpgroup's MMAs run under the other's exponentials.
//   * a fraction of the exponentials is evaluated on the FMA pipe (Cody-Waite + degree-3 minimax
//     polynomial, rel. error 7.5e-5, far below bf16 rounding of P) to unload the 16-lane MUFU unit.
// Same shared-memory layout, barriers, producer and tensor maps as the general kernel above.
// ---------------------------------------------------------------------------------------------
constexpr float kRescaleThreshold = 60.0f;

// 2^t for t <= ~100 on the FMA / ALU pipes (no MUFU): t = n + f, n = round(t), f in [-0.5, 0.5]
__device__ __forceinline__ float2 exp2_poly2(float2 t) {
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  const float2 magic = mk2(12582912.0f);                   // 1.5 * 2^23: low mantissa bits hold round(t)
  const float2 fi = __fadd2_rn(t, magic);
  const float2 n = __fadd2_rn(fi, mk2(-12582912.0f));
  const float2 f = __fadd2_rn(t, make_float2(-n.x, -n.y));
  float2 q = __ffma2_rn(mk2(0.05517132207751274f), f, mk2(0.24261054396629333f));
  q = __ffma2_rn(q, f, mk2(0.6932609677314758f));
  q = __ffma2_rn(q, f, mk2(0.9999281167984009f));
  return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(fi.x) << 23)),
                     __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(fi.y) << 23)));
}

// tcgen05.wait::ld that also names the destination registers of the load in flight, so the compiler cannot
// schedule a read (or a copy) of them above the wait
__device__ __forceinline__ void tmem_wait_ld_pin(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
        "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
        "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
        "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :: "memory");
}

template <int HD, int POLY>     // POLY: pairs (of 16) per 32-column chunk whose exp2 runs on the FMA pipe
__global__ void __launch_bounds__(320, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV, RING = Cfg::RING;
  constexpr int RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                  // [2 buffers][2 tiles]
  uint8_t* kv_s = q_s + 4 * Cfg::Q_BYTES;               // [RING]
  uint8_t* p_s = kv_s + RING * Cfg::KV_BYTES;           // [2 warpgroups]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = bars + 2;            // [2]
  uint64_t* kv_full = bars + 4;            // [RING]
  uint64_t* kv_empty = bars + 4 + RING;    // [RING]
  uint64_t* s_full = bars + 4 + 2 * RING;  // [2]   S_g(tile) is in tensor memory
  uint64_t* p_full = s_full + 4;           // [2]   P_g(tile) is in shared memory (and S_g has been read)
  uint64_t* pv_full = s_full + 6;          // [2]   the item's last PV_g has retired: O_g is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nkv = (p.Nk + BKV - 1) / BKV;
  const int nvalid_last = p.Nk - (nkv - 1) * BKV;       // keys of the last tile that exist (1..128)
  const int ksteps_last = (nvalid_last + 15) >> 4;      // PV k-steps of the last tile (columns past it are never read)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < RING; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_full[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& b, int& h, int& q0) {
    const int qp = item % p.nqp;
    const int bh = item / p.nqp;
    h = bh % p.H;
    b = bh / p.H;
    q0 = qp * 2 * BQ;
  };
  const int my_items = (p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // first query row of the k-th item of this CTA
  auto item_q0 = [&](int k) { return ((static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x)) % p.nqp) * 2 * BQ; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane issues)
    uint32_t r = 0;
    for (int k = 0; k < my_items; ++k) {
      int b, h, q0;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), b, h, q0);
      const int qb = k & 1;
      mbar_wait(&q_empty[qb], ((k >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[qb], 2 * Cfg::Q_BYTES);
        tma_load_4d(q_s + (qb * 2 + 0) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0, b);
        tma_load_4d(q_s + (qb * 2 + 1) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, h, q0 + BQ, b);
      }
      __syncwarp();
      for (int j = 0; j < nkv; ++j) {
        for (int t = 0; t < 2; ++t, ++r) {
          const int slot = r % RING;
          mbar_wait(&kv_empty[slot], ((r / RING) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&kv_full[slot], Cfg::KV_BYTES);
            tma_load_4d(kv_s + slot * Cfg::KV_BYTES, t == 0 ? &tmK : &tmV, &kv_full[slot], 0, h, j * BKV, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one continuous stream of key tiles
    if (my_items > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, false, true);
      const uint32_t p_addr0 = smem_u32(p_s);
      uint32_t pc[2] = {0, 0};       // P tiles consumed per warpgroup (p_full parity)

      auto issue_s = [&](int g, int k, uint32_t k_addr) {      // S_g of a tile of item k
        const uint32_t q_addr = smem_u32(q_s + ((k & 1) * 2 + g) * Cfg::Q_BYTES);
        const uint32_t d = tmem_base + g * 256;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(d, attn_desc<RB>(q_addr + kk * 32, 16, AB), attn_desc<RB>(k_addr + kk * 32, 16, AB), idesc_s,
                      kk > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };

      uint32_t r = 0;                // ring position of the current tile's K (V follows at r + 1)
      // prologue: S of the first tile; warpgroup 1 starts `stagger_cycles` late, once
      {
        mbar_wait(&q_full[0], 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(kv_s);
        const int q0 = item_q0(0);
        issue_s(0, 0, k_addr);
        if (q0 + BQ < p.Nq) {
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger_cycles) { }
          issue_s(1, 0, k_addr);
        }
        if (elect_one()) umma_commit(&kv_empty[0]);
        __syncwarp();
      }
      for (int k = 0; k < my_items; ++k) {
        const int q0 = item_q0(k);
        const int ng = (q0 + BQ < p.Nq) ? 2 : 1;
        for (int j = 0; j < nkv; ++j, r += 2) {
          const bool last_tile = j + 1 == nkv;
          const bool has_next = !(last_tile && k + 1 == my_items);
          const int kn = last_tile ? k + 1 : k;                 // item of the next tile
          const int sv = (r + 1) % RING;
          mbar_wait(&kv_full[sv], ((r + 1) / RING) & 1);
          const uint32_t v_addr = smem_u32(kv_s + sv * Cfg::KV_BYTES);
          int sk = 0, ng_next = 0;
          uint32_t k_addr = 0;
          if (has_next) {
            if (last_tile) mbar_wait(&q_full[kn & 1], (kn >> 1) & 1);
            sk = (r + 2) % RING;
            mbar_wait(&kv_full[sk], ((r + 2) / RING) & 1);
            k_addr = smem_u32(kv_s + sk * Cfg::KV_BYTES);
            ng_next = last_tile ? ((item_q0(kn) + BQ < p.Nq) ? 2 : 1) : ng;
          }
          const int ksteps = last_tile ? ksteps_last : BKV / 16;
          for (int g = 0; g < 2; ++g) {
            if (g < ng) {
              mbar_wait(&p_full[g], pc[g] & 1);
              ++pc[g];
              tc_fence_after();
              const uint32_t d = tmem_base + g * 256 + 128;
              const uint32_t pa = p_addr0 + g * Cfg::P_BYTES;
              if (elect_one()) {
                for (int kk = 0; kk < ksteps; ++kk)
                  umma_bf16(d, umma_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                            attn_desc<RB>(v_addr + kk * 2 * AB, 0, AB), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                if (last_tile) umma_commit(&pv_full[g]);
              }
              __syncwarp();
            }
            if (has_next && g < ng_next) issue_s(g, kn, k_addr);
          }
          if (elect_one()) {
            umma_commit(&kv_empty[sv]);
            if (has_next) umma_commit(&kv_empty[sk]);
            if (last_tile) umma_commit(&q_empty[k & 1]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * 256, tmem_o = tmem_s + 128;
    const uint32_t row_sw = row & 7, lrow_sw = lane & 7;
    uint8_t* p_row = p_s + g * Cfg::P_BYTES + row * 128;
    uint8_t* o_stage = p_s + g * Cfg::P_BYTES + qd * 4096;   // O staging reuses this warp's own P rows
    uint32_t tc = 0;         // key tiles processed (s_full parity)
    uint32_t ic = 0;         // items processed (pv_full parity)
    const float2 sc2 = mk2(p.scale_log2);

    for (int k = 0; k < my_items; ++k) {
      int b, h, q0;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), b, h, q0);
      const int qt0 = q0 + g * BQ;                  // first query row of this warpgroup's tile
      if (qt0 >= p.Nq) continue;                    // (only warpgroup 1 can be idle)
      // a warp whose 32 rows are all past Nq keeps the barrier protocol but does no work (its rows of P / O are
      // never stored: the TMA store clips them)
      const bool dead = qt0 + qd * 32 >= p.Nq;
      float m_used = 0.f, l_run = 0.f;
      // the previous item's O tile may still be leaving the staging area (== this warp's P rows)
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();

      for (int j = 0; j < nkv; ++j, ++tc) {
        mbar_wait(&s_full[g], tc & 1);
        tc_fence_after();
        if (!dead) {
          const int nvalid = j + 1 == nkv ? nvalid_last : BKV;
          const int nchunk = (nvalid + 31) >> 5;
          float mx;
          float2 l2;
          // one sweep over the tile: exponentials against m_used -> bf16 P in shared memory; returns the row
          // maximum (raw scores) in mx and the row sum in l2.  `init`: first tile, m_used := max of the first chunk.
          auto sweep = [&](bool init) {
            mx = -INFINITY;
            l2 = make_float2(0.f, 0.f);
            uint32_t v[2][32];
            tmem_ld32(tmem_s + lane_addr, v[0]);
            tmem_wait_ld_pin(v[0]);
            if (init) {
              float m0 = -INFINITY;
              if (nvalid >= 32) {
#pragma unroll
                for (int e = 0; e < 32; e += 2) m0 = fmaxf(m0, fmaxf(__uint_as_float(v[0][e]), __uint_as_float(v[0][e + 1])));
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                  if (e < nvalid) m0 = fmaxf(m0, __uint_as_float(v[0][e]));
              }
              m_used = m0 * p.scale_log2;
            }
            const float2 nm2 = mk2(-m_used);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < nchunk) {
                if (c + 1 < 4 && c + 1 < nchunk) tmem_ld32(tmem_s + lane_addr + (c + 1) * 32, v[(c + 1) & 1]);
                const uint32_t (&vc)[32] = v[c & 1];
                uint32_t pk[16];
                const int nv = nvalid - c * 32;
                if (nv >= 32) {
#pragma unroll
                  for (int e = 0; e < 32; e += 2) {
                    const float2 s2 = make_float2(__uint_as_float(vc[e]), __uint_as_float(vc[e + 1]));
                    mx = fmaxf(mx, fmaxf(s2.x, s2.y));
                    const float2 t = __ffma2_rn(s2, sc2, nm2);
                    const float2 pp = (e >> 1) >= 16 - POLY ? exp2_poly2(t) : make_float2(fast_ex2(t.x), fast_ex2(t.y));
                    l2 = __fadd2_rn(l2, pp);
                    pk[e >> 1] = pack_bf16x2(pp.x, pp.y);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 32; e += 2) {
                    const float2 s2 = make_float2(__uint_as_float(vc[e]), __uint_as_float(vc[e + 1]));
                    const float2 t = __ffma2_rn(s2, sc2, nm2);
                    float2 pp = make_float2(fast_ex2(t.x), fast_ex2(t.y));
                    if (e < nv) mx = fmaxf(mx, s2.x); else pp.x = 0.f;
                    if (e + 1 < nv) mx = fmaxf(mx, s2.y); else pp.y = 0.f;
                    l2 = __fadd2_rn(l2, pp);
                    pk[e >> 1] = pack_bf16x2(pp.x, pp.y);
                  }
                }
                uint8_t* blk = p_row + (c >> 1) * 16384;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
                  *reinterpret_cast<uint4*>(blk + ((chunk ^ row_sw) << 4)) =
                      make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
                }
                if (c + 1 < 4 && c + 1 < nchunk) tmem_wait_ld_pin(v[(c + 1) & 1]);
              }
            }
          };
          sweep(j == 0);
          const float mt = mx * p.scale_log2;
          if (__any_sync(0xffffffffu, mt - m_used > kRescaleThreshold)) {
            // slow path: move the stabiliser up to the row maximum, rescale what has been accumulated, redo the tile
            const float m_new = fmaxf(m_used, mt);
            const float alpha = exp2f(m_used - m_new);
            if (j > 0) {        // O_g is quiescent: PV_g(j-1) retired before S_g(j) was committed
#pragma unroll
              for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tmem_o + lane_addr + c * 32, o);
                tmem_wait_ld_pin(o);
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st32(tmem_o + lane_addr + c * 32, o);
              }
              tmem_wait_st();
            }
            l_run *= alpha;
            m_used = m_new;
            sweep(false);
          }
          l_run += l2.x + l2.y;
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive_warp(&p_full[g]);         // P_g(j) visible to the tensor core; S_g may be overwritten
      }
      // ---- epilogue: O / l  (the last PV of the item has retired; it was also the last reader of P)
      mbar_wait(&pv_full[g], ic & 1);
      ++ic;
      tc_fence_after();
      if (!dead) {
        const float inv_l = 1.0f / l_run;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_o + lane_addr + c * 32, v);
          tmem_wait_ld_pin(v);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q4 * 8 + e]) * inv_l;
            const uint32_t chunk = static_cast<uint32_t>(c * 4 + q4);
            uint8_t* dst = (RB == 128) ? o_stage + lane * 128 + ((chunk ^ lrow_sw) << 4)
                                       : o_stage + lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                        pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
        }
        // (the next PV_g that overwrites O_g waits for this warp's next p_full arrival: TMEM reuse is ordered)
        tc_fence_before();
        if (qt0 + row < p.Nq)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qt0 + row] = (m_used + log2f(l_run)) * 0.69314718055994531f;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmO, o_stage, 0, h, qt0 + qd * 32, b);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

int make_bnhd_tmap(CUtensorMap* tm, const void* ptr, int B, int H, int N, int hd, long long sb, long long sn,
                   long long sh, int box_rows, CUtensorMapDataType dt, int elem_bytes, int box_cols) {
  uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(sh) * elem_bytes, static_cast<uint64_t>(sn) * elem_bytes,
                         static_cast<uint64_t>(sb) * elem_bytes};
  uint32_t box[4] = {static_cast<uint32_t>(box_cols), 1u, static_cast<uint32_t>(box_rows), 1u};
  const int inner_bytes = box_cols * elem_bytes;
  CUtensorMapSwizzle swz = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : inner_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                         : inner_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
  return make_tmap(tm, ptr, dt, 4, dims, strides, box, swz);
}

static bool strides_ok(long long sb, long long sn, long long sh) {
  return sb % 8 == 0 && sn % 8 == 0 && sh % 8 == 0;
}

static int g_fwd_stagger = -1;     // < 0: derived from the problem size
static int g_fwd_variant = 0;      // 0: single-sweep kernel (default); 1: round-1 kernels (short-key / general); profiling aid
static int g_fwd_poly = 4;         // single-sweep kernel: exp2 pairs per 16 evaluated on the FMA pipe (0, 4 or 8)

template <typename K>
static int set_smem_attr(K kernel, int bytes, bool& done) {
  if (done) return UCF_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_last_error("attention_fwd: smem attr: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  done = true;
  return UCF_OK;
}

template <int HD>
static int launch_attn_fwd(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const CUtensorMap& tO,
                           AttnFwdParams p, cudaStream_t st) {
  using Cfg = AttnFwdCfg<HD>;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  int rc;
  if (g_fwd_variant == 0) {
    static bool a0 = false, a4 = false, a8 = false;
    p.stagger_cycles = g_fwd_stagger >= 0 ? g_fwd_stagger : 600;
    if (g_fwd_poly == 0) {
      if ((rc = set_smem_attr(attn_fwd2_kernel<HD, 0>, Cfg::SMEM_BYTES, a0))) return rc;
      attn_fwd2_kernel<HD, 0><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
    } else if (g_fwd_poly == 8) {
      if ((rc = set_smem_attr(attn_fwd2_kernel<HD, 8>, Cfg::SMEM_BYTES, a8))) return rc;
      attn_fwd2_kernel<HD, 8><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
    } else {
      if ((rc = set_smem_attr(attn_fwd2_kernel<HD, 4>, Cfg::SMEM_BYTES, a4))) return rc;
      attn_fwd2_kernel<HD, 4><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
    }
    return check_launch("attn_fwd2_kernel");
  }
  if (p.Nk <= 2 * Cfg::BKV && g_fwd_variant != 2) {
    static bool attr2 = false;
    if ((rc = set_smem_attr(attn_fwd_short_kernel<HD>, Cfg::SMEM_BYTES, attr2))) return rc;
    attn_fwd_short_kernel<HD><<<grid, Cfg::THREADS + 32, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
    return check_launch("attn_fwd_short_kernel");
  }
  static bool attr = false;
  if ((rc = set_smem_attr(attn_fwd_kernel<HD>, Cfg::SMEM_BYTES, attr))) return rc;
  attn_fwd_kernel<HD><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
  return check_launch("attn_fwd_kernel");
}

}  // namespace ucf

using namespace ucf;

static long long* g_fwd_timeline = nullptr;
/* profiling aid (not part of the public header): device buffer of 64 int64 receiving clock64 stamps */
extern "C" void ucf_debug_set_attn_fwd_timeline(void* dev_ptr) { g_fwd_timeline = static_cast<long long*>(dev_ptr); }
/* profiling aids: variant 0 = single-sweep kernel (default), 1 = round-1 short-key / general kernels, 2 = round-1 general
   kernel for every Nk; poly = exp2 pairs per 16 on the FMA pipe in the single-sweep kernel (0, 4, 8) */
extern "C" void ucf_debug_set_attn_fwd_variant(int variant) { ucf::g_fwd_variant = variant; }
extern "C" void ucf_debug_set_attn_fwd_poly(int pairs) { ucf::g_fwd_poly = pairs; }
extern "C" void ucf_debug_force_general_attn_fwd(int on) { ucf::g_fwd_variant = on ? 2 : 0; }
extern "C" void ucf_debug_set_attn_fwd_stagger(int cycles) { ucf::g_fwd_stagger = cycles; }

extern "C" int ucf_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                                 int B, int H, int Nq, int Nk, int hd,
                                 long long q_sb, long long q_sn, long long q_sh,
                                 long long k_sb, long long k_sn, long long k_sh,
                                 long long v_sb, long long v_sn, long long v_sh,
                                 long long o_sb, long long o_sn, long long o_sh,
                                 float scale, void* stream) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk <= 0) { set_last_error("attention_fwd: empty problem"); return UCF_ERR_BAD_ARG; }
  if (hd != 64 && hd != 32) {
    set_last_error("attention_fwd: head_dim %d not supported (32 or 64)", hd);
    return UCF_ERR_UNSUPPORTED;
  }
  if (!q || !k || !v || !o || !lse) { set_last_error("attention_fwd: null pointer"); return UCF_ERR_BAD_ARG; }
  if (!strides_ok(q_sb, q_sn, q_sh) || !strides_ok(k_sb, k_sn, k_sh) || !strides_ok(v_sb, v_sn, v_sh) ||
      !strides_ok(o_sb, o_sn, o_sh) ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(o)) & 15)) {
    set_last_error("attention_fwd: pointers and strides must be 16-byte aligned");
    return UCF_ERR_BAD_ARG;
  }
  CUtensorMap tQ, tK, tV, tO;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = make_bnhd_tmap(&tQ, q, B, H, Nq, hd, q_sb, q_sn, q_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tK, k, B, H, Nk, hd, k_sb, k_sn, k_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tV, v, B, H, Nk, hd, v_sb, v_sn, v_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tO, o, B, H, Nq, hd, o_sb, o_sn, o_sh, 32, bf, 2, hd))) return rc;
  AttnFwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.nqp = (Nq + 255) / 256;
  const long long items = static_cast<long long>(B) * H * p.nqp;
  if (items > 0x7fffffffLL) { set_last_error("attention_fwd: too many work items"); return UCF_ERR_BAD_ARG; }
  p.items = static_cast<int>(items);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.timeline = g_fwd_timeline;
  // half of an item's period; the kernel is bound by tensor-memory reads (64 B/clk/SM: S twice + O once per
  // warpgroup and item), ~9 cycles per 32-bit column of a 128-row tile and warpgroup pair
  {
    const int c0 = ((Nk < 128 ? Nk : 128) + 31) / 32 * 32, c1 = Nk > 128 ? (Nk - 128 + 31) / 32 * 32 : 0;
    p.stagger_cycles = g_fwd_stagger >= 0 ? g_fwd_stagger : 9 * (2 * (c0 + c1) + hd);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return hd == 64 ? launch_attn_fwd<64>(tQ, tK, tV, tO, p, st) : launch_attn_fwd<32>(tQ, tK, tV, tO, p, st);
}
