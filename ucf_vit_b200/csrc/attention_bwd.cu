// Fused softmax attention backward on tcgen05 (sm_100a).
// Gradient of Attention.forward's core (/root/reference/src/UCF_VIT/simple/building_blocks.py:163-187),
// which the reference leaves to autograd over SDPA / xformers.
//
// One CTA per (128-key tile j, head, batch element); it loops over 128-query tiles i:
//     S  = Q_i K_j^T            dP = dO_i V_j^T                    (tensor core -> TMEM)
//     P  = exp(S*scale - lse)   dS = scale * P o (dP - delta)      (registers, thread == query row)
//     dV_j += P^T dO_i          dK_j += dS^T Q_i                   (accumulate in TMEM over i)
//     dQ_i  = dS K_j  -> fp32 TMA reduce-add into dq_acc           (summed over j by the TMA unit)
// P and dS are written once, as bf16, into 128B-swizzled shared memory and consumed both as a
// K-major operand (dS K) and as an MN-major operand (P^T dO, dS^T Q) -- same bytes, two
// descriptors.  delta = rowsum(dO o O) comes from a small bandwidth-bound pre-pass; dq_acc is
// converted to bf16 by a post-pass.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

int make_bnhd_tmap(CUtensorMap* tm, const void* ptr, int B, int H, int N, int hd, long long sb, long long sn,
                   long long sh, int box_rows, CUtensorMapDataType dt, int elem_bytes, int box_cols);

struct AttnBwdParams {
  int B, H, Nq, Nk;
  int nkt;              // key tiles per (b, h)
  int items;            // LONG: B * H * nkt work items (b, h, key tile); SHORT: B * H work items (b, h)
  float scale, scale_log2;
  const float* lse;     // [B,H,Nq]
  const float* delta;   // [B,H,Nq]
  long long* timeline;  // optional [256] clock64 stamps of CTA 0 (profiling aid; NULL in production)
};

template <int HD>
struct AttnBwdCfg {
  static constexpr int T = 128;
  static constexpr int TILE_BYTES = T * HD * 2;       // one Q / K / V / dO tile
  static constexpr int PS_BYTES = T * T * 2;          // P or dS
  static constexpr int NBAR = 24;
  // K,V x2 (per item) | Q x2 | dO x2 | P | dS x2.   dQ / dK / dV staging reuses dead dS rows.
  static constexpr int SMEM_BYTES = 1024 + 4 * TILE_BYTES + 4 * TILE_BYTES + 3 * PS_BYTES + NBAR * 8 + 16;
  static constexpr int ROW_BYTES = HD * 2;
  static constexpr int ATOM_BYTES = 8 * ROW_BYTES;
  static_assert(SMEM_BYTES <= 232448, "smem");
};

// profiling aid: clock64 stamps of CTA 0's first 8 tiles, 32 slots per tile (NULL in production)
#define UCF_TL(t_, k_)                                                                   \
  do {                                                                                   \
    if (p.timeline && blockIdx.x == 0 && (t_) < 8) p.timeline[(t_) * 32 + (k_)] = clock64(); \
  } while (0)

template <int ROW_BYTES>
__device__ __forceinline__ uint64_t bwd_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = umma_smem_desc(saddr, lbo, sbo);
  if (ROW_BYTES == 64) d = (d & ~(7ull << 61)) | (4ull << 61);   // SWIZZLE_64B
  return d;
}

// Persistent: one CTA per SM walks ONE continuous stream of (item, query-tile) pairs, item =
// (batch, head, 128-key tile).  The tensor core is the critical resource (all five products read
// both operands from shared memory: ~224 KB of operand traffic per tile), so everything else is
// arranged to keep it fed (cycle counts from the clock64 timeline hook at N = 197, hd = 64:
// S/dP 770, exp/dS math 1700-1800, dV+dQ+dK 2200):
//   * K/V are double-buffered per item and Q/dO stream through a 2-slot ring, so the producer
//     runs a whole item ahead -- no load bubble at item boundaries;
//   * per tile the issue order is S/dP(t+1), dV(t), dQ(t), dK(t): the next tile's S/dP goes first so
//     that its exp/dS math runs on the CUDA cores under all three products of tile t; dS and the
//     dQ accumulator are double-buffered, the single P buffer is written last (P waits in
//     registers until dV(t) has retired);
//   * the compute warps drain dQ(t-1) after the math of tile t, and an item's dK/dV after the
//     math of the NEXT item's first tile, so draining never delays math.
//
// Two schedules share the code.  LONG (any Nq): a work item is (b, h, key tile); Q/dO tiles stream through
// the ring once per key tile and every tile's dQ leaves through a fp32 TMA reduce-add into dq_acc (summed
// over key tiles by the L2), converted to bf16 by a post-pass.  SHORT (Nq <= 256, i.e. at most two query
// tiles -- ViT-B/16 at 224 px has 197 tokens): a work item is (b, h) and the CTA walks ALL its key tiles, so
// the two Q/dO tiles are loaded once and stay resident, and dQ_i accumulates over the key tiles in its own
// TMEM buffer and is written once, as bf16 -- no fp32 workspace, no zero fill, no reduce traffic, no cast.
template <int HD, bool SHORT>
__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQacc, const __grid_constant__ CUtensorMap tmdK,
                const __grid_constant__ CUtensorMap tmdV, const AttnBwdParams p) {
  using Cfg = AttnBwdCfg<HD>;
  constexpr int T = Cfg::T, TB = Cfg::TILE_BYTES, RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;
  constexpr int PSB = Cfg::PS_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic)
  uint8_t* kv_s = smem;             // [2 items][K, V]
  uint8_t* q_s = kv_s + 4 * TB;     // [2]
  uint8_t* do_s = q_s + 2 * TB;     // [2]
  uint8_t* p_s = do_s + 2 * TB;     // [1]
  uint8_t* ds_s = p_s + PSB;        // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + 2 * PSB);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* qdo_full = bars + 4;     // [2]
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;
  uint64_t* sdp_empty = bars + 9;
  uint64_t* pds_full = bars + 10;    // [2]
  uint64_t* dq_full = bars + 12;     // [2]  (also: every product of that tile has retired)
  uint64_t* dq_empty = bars + 14;    // [2]
  uint64_t* dkv_full = bars + 16;
  uint64_t* dkv_empty = bars + 17;
  uint64_t* p_free = bars + 18;      // dV(t) retired: P may be rewritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  // warp index through a shuffle: warp-uniform for the compiler, so the producer / MMA warps run converged and
  // keep TMA / UMMA descriptors in uniform registers (no per-instruction ELECT / R2UR waterfall)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nq = (p.Nq + T - 1) / T;
  const int nkt = p.nkt;
  // sub-item = (b, h, key tile): nq tiles each.  LONG: one per work item; SHORT: nkt consecutive per work item.
  const int my_items = ((p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)) *
                       (SHORT ? nkt : 1);
  const uint32_t total_tiles = static_cast<uint32_t>(my_items) * nq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQacc); tma_prefetch_desc(&tmdK); tma_prefetch_desc(&tmdV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1);
      mbar_init(&pds_full[i], 8);
      mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 8);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, 8);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, 8);
    mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 320,
                 t_dq = tmem_base + 384;   // two dQ buffers: +0, +64

  // n-th sub-item of this CTA -> (b, h, key tile index)
  auto decode = [&](uint32_t n, int& b, int& h, int& jt) {
    int bh;
    if (SHORT) {
      jt = static_cast<int>(n) % nkt;
      bh = static_cast<int>(blockIdx.x) + (static_cast<int>(n) / nkt) * static_cast<int>(gridDim.x);
    } else {
      const int item = static_cast<int>(blockIdx.x) + static_cast<int>(n) * static_cast<int>(gridDim.x);
      jt = item % nkt;
      bh = item / nkt;
    }
    h = bh % p.H;
    b = bh / p.H;
  };
  // Q/dO ring slot and dQ TMEM buffer of tile t = (sub-item n, query tile i) are picked by a counter u
  // of query-tile INSTANCES: LONG u = t (fresh Q/dO and dQ every tile); SHORT u counts (work item, i)
  // pairs, so the tiles of all key tiles of one work item share the slot.
  auto qinst = [&](uint32_t t, uint32_t n, int i) -> uint32_t {
    return SHORT ? (n / static_cast<uint32_t>(nkt)) * nq + i : t;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane issues)
    {
      uint32_t t = 0;
      for (uint32_t n = 0; n < static_cast<uint32_t>(my_items); ++n) {
        int b, h, jt;
        decode(n, b, h, jt);
        const int kb = n & 1;
        mbar_wait(&kv_empty[kb], ((n >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[kb], 2 * TB);
          tma_load_4d(kv_s + (kb * 2 + 0) * TB, &tmK, &kv_full[kb], 0, h, jt * T, b);
          tma_load_4d(kv_s + (kb * 2 + 1) * TB, &tmV, &kv_full[kb], 0, h, jt * T, b);
        }
        __syncwarp();
        for (int i = 0; i < nq; ++i, ++t) {
          if (SHORT && jt > 0) continue;            // resident since the work item's first key tile
          const uint32_t u = qinst(t, n, i);
          const int slot = u & 1;
          mbar_wait(&qdo_empty[slot], ((u >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&qdo_full[slot], 2 * TB);
            tma_load_4d(q_s + slot * TB, &tmQ, &qdo_full[slot], 0, h, i * T, b);
            tma_load_4d(do_s + slot * TB, &tmdO, &qdo_full[slot], 0, h, i * T, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    if (total_tiles > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(T, T, false, false);     // S, dP
      constexpr uint32_t idesc_tn = umma_idesc_bf16(T, HD, true, true);      // dV, dK
      constexpr uint32_t idesc_dq = umma_idesc_bf16(T, HD, false, true);     // dQ
      const uint32_t p_addr = smem_u32(p_s);

      auto issue_sdp = [&](uint32_t t) {        // S, dP of global tile t (item n = t / nq)
        const uint32_t n = t / nq;
        UCF_TL(t, 7);
        if (t % nq == 0) mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
        UCF_TL(t, 8);
        const uint32_t u = qinst(t, n, static_cast<int>(t % nq));
        const int slot = u & 1;
        const uint32_t k_addr = smem_u32(kv_s + ((n & 1) * 2 + 0) * TB), v_addr = smem_u32(kv_s + ((n & 1) * 2 + 1) * TB);
        const uint32_t q_addr = smem_u32(q_s + slot * TB), do_addr = smem_u32(do_s + slot * TB);
        mbar_wait(&qdo_full[slot], (u >> 1) & 1);     // SHORT, later key tiles: phase already complete
        UCF_TL(t, 9);
        mbar_wait(sdp_empty, (t & 1) ^ 1);
        tc_fence_after();
        UCF_TL(t, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(t_s, bwd_desc<RB>(q_addr + k * 32, 16, AB), bwd_desc<RB>(k_addr + k * 32, 16, AB), idesc_qk, k > 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(t_dp, bwd_desc<RB>(do_addr + k * 32, 16, AB), bwd_desc<RB>(v_addr + k * 32, 16, AB), idesc_qk, k > 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };

      // are the operands of tile t's S/dP in shared memory yet?  (non-blocking; lane 0 decides for the warp)
      auto sdp_inputs_ready = [&](uint32_t t) -> bool {
        const uint32_t n = t / nq;
        const uint32_t u = qinst(t, n, static_cast<int>(t % nq));
        bool ok = mbar_try_wait(&qdo_full[u & 1], (u >> 1) & 1);
        if (ok && t % nq == 0) ok = mbar_try_wait(&kv_full[n & 1], (n >> 1) & 1);
        return __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
      };

      issue_sdp(0);
      for (uint32_t t = 0; t < total_tiles; ++t) {
        const uint32_t n = t / nq;
        const int i = static_cast<int>(t - n * nq);
        const uint32_t u = qinst(t, n, i);
        const int slot = u & 1, db = t & 1;
        const int jt = SHORT ? static_cast<int>(n) % nkt : 0;
        const bool first_kt = !SHORT || jt == 0, last_kt = !SHORT || jt == nkt - 1;
        const uint32_t k_addr = smem_u32(kv_s + ((n & 1) * 2 + 0) * TB);
        const uint32_t q_addr = smem_u32(q_s + slot * TB), do_addr = smem_u32(do_s + slot * TB);
        const uint32_t ds_addr = smem_u32(ds_s + db * PSB);
        // next tile's S/dP first -- as soon as the compute warps have read S/dP(t) out of tensor memory,
        // i.e. before P(t)/dS(t) are even written: its exp/dS math then overlaps all three products of tile t
        // (at a work-item boundary the next Q/dO may still be in flight: then tile t's products go first
        // instead of stalling the in-order issue behind the load)
        const bool sdp_first = t + 1 < total_tiles && sdp_inputs_ready(t + 1);
        if (sdp_first) issue_sdp(t + 1);
        mbar_wait(&pds_full[db], (t >> 1) & 1);                      // P(t), dS(t) are in shared memory
        if (i == 0) mbar_wait(dkv_empty, (n & 1) ^ 1);               // previous item's dK/dV have left TMEM
        tc_fence_after();
        UCF_TL(t, 1);
        // dV += P^T dO_t   (reduction over the 128 query rows, 16 per MMA); releases the single P buffer
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16(t_dv, umma_smem_desc(p_addr + kk * 2048, 16384, 1024), bwd_desc<RB>(do_addr + kk * 2 * AB, 0, AB),
                      idesc_tn, (i > 0 || kk > 0) ? 1u : 0u);
          umma_commit(p_free);
        }
        __syncwarp();
        if (first_kt) {
          mbar_wait(&dq_empty[slot], ((u >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        UCF_TL(t, 11);
        // dQ_t (+)= dS K_j   (reduction over the 128 keys; SHORT: accumulated over the key tiles)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16(t_dq + slot * 64, umma_smem_desc(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      bwd_desc<RB>(k_addr + kk * 2 * AB, 0, AB), idesc_dq, (!first_kt || kk > 0) ? 1u : 0u);
          // dK += dS^T Q_t
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16(t_dk, umma_smem_desc(ds_addr + kk * 2048, 16384, 1024), bwd_desc<RB>(q_addr + kk * 2 * AB, 0, AB),
                      idesc_tn, (i > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&dq_full[db]);          // every product of tile t has retired
          if (last_kt) umma_commit(&qdo_empty[slot]);
          if (i + 1 == nq) {
            umma_commit(dkv_full);
            umma_commit(&kv_empty[n & 1]);
          }
        }
        __syncwarp();
        if (t + 1 < total_tiles && !sdp_first) issue_sdp(t + 1);
        UCF_TL(t, 12);
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps
    // 8 warps: two per TMEM lane quarter; warp (qd, half) owns rows qd*32.. and the key columns
    // [half*64, half*64+64) of S / dP, and the head-dim columns [half*HD/2, ...) of dQ / dK / dV.
    const int cw = warp - 2;
    const int qd = warp & 3;
    const int half = cw >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t row_sw = row & 7;
    const uint32_t lrow_sw = lane & 7;
    constexpr int HH = HD / 2;                                // head-dim columns per warp (32 or 16)
    const uint32_t blk_off = half * 16384 + row * 128;        // this thread's row in its 64-key block of P / dS
    const uint32_t stage_off = half * 16384 + qd * 4096;      // 4 KB of this warp's OWN dS rows, reused for staging
    const float LOG2E = 1.4426950408889634f;

    // dQ of global tile t (all its products have retired): TMEM -> fp32 staging in the dead rows of
    // dS[t&1] -> TMA reduce-add into dq_acc
    // `drain` is false for the SHORT schedule's tiles before the last key tile: the wait still happens
    // (it is what frees dS[t&1] for tile t+2), dQ stays in TMEM.
    auto flush_dq = [&](uint32_t t, uint32_t u, bool drain, int b, int h, int qrow0) {
      const uint32_t db = t & 1, slot = u & 1;
      mbar_wait(&dq_full[db], (t >> 1) & 1);
      if (!drain) return;
      tc_fence_after();
      if (threadIdx.x == 64) UCF_TL(t, 4);
      uint8_t* my_dq = ds_s + db * PSB + stage_off;
      uint32_t v[HH];
      if (HH == 32) tmem_ld32(t_dq + slot * 64 + lane_addr + half * HH, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      else tmem_ld16(t_dq + slot * 64 + lane_addr + half * HH, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      tmem_wait_ld();
      if (SHORT) {       // final values: bf16 rows of HH elements (64 B SWIZZLE_64B / 32 B SWIZZLE_32B)
#pragma unroll
        for (int g = 0; g < HH / 8; ++g) {
          uint8_t* dst = (HH == 32) ? my_dq + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4)
                                    : my_dq + lane * 32 + ((static_cast<uint32_t>(g) ^ ((lane >> 2) & 1)) << 4);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
        }
      } else {           // partial sums: fp32 rows of HH elements (128 B SWIZZLE_128B / 64 B SWIZZLE_64B)
#pragma unroll
        for (int g = 0; g < HH / 4; ++g) {
          uint8_t* dst = (HH == 32) ? my_dq + lane * 128 + ((static_cast<uint32_t>(g) ^ lrow_sw) << 4)
                                    : my_dq + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&dq_empty[slot]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && qrow0 + qd * 32 < p.Nq) {
        if (SHORT) tma_store_4d(&tmdQacc, my_dq, half * HH, h, qrow0 + qd * 32, b);
        else tma_reduce_add_4d(&tmdQacc, my_dq, half * HH, h, qrow0 + qd * 32, b);
        tma_store_commit();
      }
      if (threadIdx.x == 64) UCF_TL(t, 5);
    };
    // dK / dV of this CTA's n-th item (its last tile is `t_last`, whose dQ staging has been issued)
    auto flush_dkv = [&](uint32_t n, uint32_t t_last, int b, int h, int k0) {
      mbar_wait(dkv_full, n & 1);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();        // dQ(t_last) staging shares these rows
      __syncwarp();
      uint8_t* st_dv = ds_s + (t_last & 1) * PSB + stage_off;
      uint8_t* st_dk = st_dv + 2048;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint8_t* st = which == 0 ? st_dv : st_dk;
        const uint32_t t_src = (which == 0 ? t_dv : t_dk) + half * HH;
        uint32_t v[HH];
        if (HH == 32) tmem_ld32(t_src + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        else tmem_ld16(t_src + lane_addr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_wait_ld();
        // rows of HH bf16 = 64 B (SWIZZLE_64B) or 32 B (SWIZZLE_32B)
#pragma unroll
        for (int g = 0; g < HH / 8; ++g) {
          uint8_t* dst = (HH == 32) ? st + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4)
                                    : st + lane * 32 + ((static_cast<uint32_t>(g) ^ ((lane >> 2) & 1)) << 4);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
        }
      }
      tc_fence_before();
      mbar_arrive_warp(dkv_empty);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && k0 + qd * 32 < p.Nk) {
        tma_store_4d(&tmdV, st_dv, half * HH, h, k0 + qd * 32, b);
        tma_store_4d(&tmdK, st_dk, half * HH, h, k0 + qd * 32, b);
        tma_store_commit();
      }
    };

    int pb_ = 0, ph_ = 0, pk0_ = 0;      // coordinates of the previous sub-item (drained one tile late)
    bool plast_ = true;                  // ... and whether it was its work item's last key tile
    uint32_t t = 0;
    for (uint32_t n = 0; n < static_cast<uint32_t>(my_items); ++n) {
      int b, h, jt;
      decode(n, b, h, jt);
      const int k0 = jt * T;
      const bool last_kt = !SHORT || jt == nkt - 1;
      const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Nq;
      for (int i = 0; i < nq; ++i, ++t) {
        const uint32_t db = t & 1;
        const int qrow = i * T + row;
        float lse_v = INFINITY, delta = 0.f;   // rows past Nq: P = exp2(-inf) = 0
        if (qrow < p.Nq) {
          lse_v = p.lse[stat_base + qrow];
          delta = p.delta[stat_base + qrow];
        }
        // every earlier TMA store of this warp has finished reading its staging rows (dS rows)
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        mbar_wait(sdp_full, t & 1);
        tc_fence_after();
        if (threadIdx.x == 64) UCF_TL(t, 2);
        if (lane == 0) UCF_TL(t, 24 + cw);
        const float nlse = -lse_v * LOG2E;
        const float2 sl2 = mk2(p.scale_log2), nlse2 = mk2(nlse), sc2 = mk2(p.scale), nds2 = mk2(-delta * p.scale);
        uint8_t* p_row = p_s + blk_off;
        uint8_t* ds_row = ds_s + db * PSB + blk_off;
        uint32_t pk[2][16];       // P stays in registers until dV(t-1) has released the single P buffer
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld32(t_s + lane_addr + half * 64 + c * 32, sv);
          tmem_ld32(t_dp + lane_addr + half * 64 + c * 32, dv);
          tmem_wait_ld();
          uint32_t dk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 tt = __ffma2_rn(make_float2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2, nlse2);
            const float2 pp = make_float2(fast_ex2(tt.x), fast_ex2(tt.y));
            // dS = scale * P * (dP - delta)
            const float2 dd = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])), sc2, nds2));
            pk[c][e >> 1] = pack_bf16x2(pp.x, pp.y);
            dk[e >> 1] = pack_bf16x2(dd.x, dd.y);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t off = ((static_cast<uint32_t>(c * 4 + g)) ^ row_sw) << 4;
            *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive_warp(sdp_empty);                          // S / dP may be overwritten
        if (lane == 0 && cw == 0) UCF_TL(t, 13);
        if (t > 0) mbar_wait(p_free, (t - 1) & 1);       // dV(t-1) has finished reading P
        if (lane == 0 && cw == 0) UCF_TL(t, 14);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t off = ((static_cast<uint32_t>(c * 4 + g)) ^ row_sw) << 4;
            *reinterpret_cast<uint4*>(p_row + off) = make_uint4(pk[c][4 * g], pk[c][4 * g + 1], pk[c][4 * g + 2], pk[c][4 * g + 3]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(&pds_full[db]);
        if (threadIdx.x == 64) UCF_TL(t, 3);
        if (lane == 0) UCF_TL(t, 16 + cw);
        // drain what the tensor core finished while this tile's math ran
        if (t > 0) {
          if (i > 0) {
            flush_dq(t - 1, qinst(t - 1, n, i - 1), last_kt, b, h, (i - 1) * T);
          } else {                                   // previous tile closed the previous sub-item
            flush_dq(t - 1, qinst(t - 1, n - 1, nq - 1), plast_, pb_, ph_, (nq - 1) * T);
            flush_dkv(n - 1, t - 1, pb_, ph_, pk0_);
          }
        }
      }
      pb_ = b; ph_ = h; pk0_ = k0; plast_ = last_kt;
    }
    if (total_tiles > 0) {
      const uint32_t nl = static_cast<uint32_t>(my_items) - 1;
      flush_dq(total_tiles - 1, qinst(total_tiles - 1, nl, nq - 1), true, pb_, ph_, (nq - 1) * T);
      flush_dkv(nl, total_tiles - 1, pb_, ph_, pk0_);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Transposed schedule (round 2; default).
//
// What bounded the kernel above was SHARED-MEMORY bandwidth, not the tensor core: all five products read both
// operands from shared memory and P / dS are written there as well -- ~304 KB per 128x128 tile against
// 128 B/clk = 2400 cycles, with an M=128 N=64 K=16 MMA streaming 6 KB = 48 cycles for 32 cycles of tensor work
// (scripts/micro/umma_rate.cu: SS 48.1 cycles per MMA, A operand from tensor memory 32.1).  Here the scores are
// produced TRANSPOSED, S^T = K_j Q_i^T and dP^T = V_j dO_i^T (TMEM lane = key, column = query), so that
//     P^T  = exp2(S^T * scale*log2e - lse_i*log2e)      dS^T = scale * P^T o (dP^T - delta_i)
// are exactly the A operands of the two accumulating products
//     dV_j += P^T dO_i          dK_j += dS^T Q_i        (A from TENSOR MEMORY: bf16 pairs written back with
//                                                        tcgen05.st over columns of S^T / dP^T this thread has read)
// and only dQ_i = dS K_j still takes dS from shared memory (the same bytes, read as an MN-major A operand).
// Shared-memory traffic per tile drops to ~208 KB, the P buffer and its hand-shake disappear, and the
// S^T/dP^T(t+1) products are issued in order right behind dV(t) / dK(t), which is also what makes the aliasing
// safe.  lse_i / delta_i now vary along a thread's columns: the compute warps stage the tile's 2 x 128 statistics
// (pre-multiplied) in shared memory one tile ahead and read them back as broadcast vectors.
// Everything else -- producer, K/V double buffer, Q/dO ring, SHORT / LONG schedules, dQ reduce-add, flushes --
// is the kernel above.
template <int HD>
struct AttnBwd2Cfg {
  static constexpr int T = 128;
  static constexpr int TILE_BYTES = T * HD * 2;
  static constexpr int PS_BYTES = T * T * 2;
  static constexpr int NBAR = 24;
  static constexpr int STAT_BYTES = 2 * 2 * T * 4;     // [2 buffers][nlse | ndelta*scale][128] fp32
  // K,V x2 (per item) | Q x2 | dO x2 | dS^T x2 | stats.   dQ / dK / dV staging reuses dead dS^T rows.
  static constexpr int SMEM_BYTES = 1024 + 8 * TILE_BYTES + 2 * PS_BYTES + STAT_BYTES + NBAR * 8 + 16;
  static constexpr int ROW_BYTES = HD * 2;
  static constexpr int ATOM_BYTES = 8 * ROW_BYTES;
  static_assert(SMEM_BYTES <= 232448, "smem");
};

template <int HD, bool SHORT>
__global__ void __launch_bounds__(320, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmdQacc, const __grid_constant__ CUtensorMap tmdK,
                 const __grid_constant__ CUtensorMap tmdV, const AttnBwdParams p) {
  using Cfg = AttnBwd2Cfg<HD>;
  constexpr int T = Cfg::T, TB = Cfg::TILE_BYTES, RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;
  constexpr int PSB = Cfg::PS_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* kv_s = smem;             // [2 items][K, V]
  uint8_t* q_s = kv_s + 4 * TB;     // [2]
  uint8_t* do_s = q_s + 2 * TB;     // [2]
  uint8_t* ds_s = do_s + 2 * TB;    // [2]  dS^T: row = key, 128 queries contiguous (two 64-wide swizzled blocks)
  float* stat_s = reinterpret_cast<float*>(ds_s + 2 * PSB);    // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stat_s) + Cfg::STAT_BYTES);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* qdo_full = bars + 4;     // [2]
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;
  uint64_t* pds_full = bars + 10;    // [2]  P^T / dS^T of the tile are in tensor memory, dS^T in shared memory
  uint64_t* dq_full = bars + 12;     // [2]  (also: every product of that tile has retired)
  uint64_t* dq_empty = bars + 14;    // [2]
  uint64_t* dkv_full = bars + 16;
  uint64_t* dkv_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int nq = (p.Nq + T - 1) / T;
  const int nkt = p.nkt;
  const int my_items = ((p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)) *
                       (SHORT ? nkt : 1);
  const uint32_t total_tiles = static_cast<uint32_t>(my_items) * nq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQacc); tma_prefetch_desc(&tmdK); tma_prefetch_desc(&tmdV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1);
      mbar_init(&pds_full[i], 8);
      mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 8);
    }
    mbar_init(sdp_full, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S^T (fp32, then P^T as bf16 pairs over its own columns) | dP^T (then dS^T) | dV | dK | dQ x2
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 320,
                 t_dq = tmem_base + 384;

  auto decode = [&](uint32_t n, int& b, int& h, int& jt) {
    int bh;
    if (SHORT) {
      jt = static_cast<int>(n) % nkt;
      bh = static_cast<int>(blockIdx.x) + (static_cast<int>(n) / nkt) * static_cast<int>(gridDim.x);
    } else {
      const int item = static_cast<int>(blockIdx.x) + static_cast<int>(n) * static_cast<int>(gridDim.x);
      jt = item % nkt;
      bh = item / nkt;
    }
    h = bh % p.H;
    b = bh / p.H;
  };
  auto qinst = [&](uint32_t t, uint32_t n, int i) -> uint32_t {
    return SHORT ? (n / static_cast<uint32_t>(nkt)) * nq + i : t;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (unchanged)
    uint32_t t = 0;
    for (uint32_t n = 0; n < static_cast<uint32_t>(my_items); ++n) {
      int b, h, jt;
      decode(n, b, h, jt);
      const int kb = n & 1;
      mbar_wait(&kv_empty[kb], ((n >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[kb], 2 * TB);
        tma_load_4d(kv_s + (kb * 2 + 0) * TB, &tmK, &kv_full[kb], 0, h, jt * T, b);
        tma_load_4d(kv_s + (kb * 2 + 1) * TB, &tmV, &kv_full[kb], 0, h, jt * T, b);
      }
      __syncwarp();
      for (int i = 0; i < nq; ++i, ++t) {
        if (SHORT && jt > 0) continue;
        const uint32_t u = qinst(t, n, i);
        const int slot = u & 1;
        mbar_wait(&qdo_empty[slot], ((u >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&qdo_full[slot], 2 * TB);
          tma_load_4d(q_s + slot * TB, &tmQ, &qdo_full[slot], 0, h, i * T, b);
          tma_load_4d(do_s + slot * TB, &tmdO, &qdo_full[slot], 0, h, i * T, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (total_tiles > 0) {
      constexpr uint32_t idesc_st = umma_idesc_bf16(T, T, false, false);     // S^T = K Q^T, dP^T = V dO^T
      constexpr uint32_t idesc_ts = umma_idesc_bf16(T, HD, false, true);     // dV, dK: A from tensor memory, B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16(T, HD, true, true);      // dQ: A = dS^T read MN-major, B = K MN-major

      auto issue_sdp = [&](uint32_t t) {
        const uint32_t n = t / nq;
        if (t % nq == 0) mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
        const uint32_t u = qinst(t, n, static_cast<int>(t % nq));
        const int slot = u & 1;
        const uint32_t k_addr = smem_u32(kv_s + ((n & 1) * 2 + 0) * TB), v_addr = smem_u32(kv_s + ((n & 1) * 2 + 1) * TB);
        const uint32_t q_addr = smem_u32(q_s + slot * TB), do_addr = smem_u32(do_s + slot * TB);
        mbar_wait(&qdo_full[slot], (u >> 1) & 1);
        tc_fence_after();
        UCF_TL(t, 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(t_s, bwd_desc<RB>(k_addr + k * 32, 16, AB), bwd_desc<RB>(q_addr + k * 32, 16, AB), idesc_st, k > 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(t_dp, bwd_desc<RB>(v_addr + k * 32, 16, AB), bwd_desc<RB>(do_addr + k * 32, 16, AB), idesc_st, k > 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };

      issue_sdp(0);
      for (uint32_t t = 0; t < total_tiles; ++t) {
        const uint32_t n = t / nq;
        const int i = static_cast<int>(t - n * nq);
        const uint32_t u = qinst(t, n, i);
        const int slot = u & 1, db = t & 1;
        const int jt = SHORT ? static_cast<int>(n) % nkt : 0;
        const bool first_kt = !SHORT || jt == 0, last_kt = !SHORT || jt == nkt - 1;
        const uint32_t k_addr = smem_u32(kv_s + ((n & 1) * 2 + 0) * TB);
        const uint32_t q_addr = smem_u32(q_s + slot * TB), do_addr = smem_u32(do_s + slot * TB);
        const uint32_t ds_addr = smem_u32(ds_s + db * PSB);
        mbar_wait(&pds_full[db], (t >> 1) & 1);                      // P^T(t), dS^T(t) are in place
        if (i == 0) mbar_wait(dkv_empty, (n & 1) ^ 1);               // previous item's dK/dV have left tensor memory
        tc_fence_after();
        UCF_TL(t, 1);
        // dV += P^T dO_t, dK += dS^T Q_t: A = bf16 pairs in tensor memory, 8 columns per 16 queries; the two
        // 64-query halves sit at columns 0.. and 64.. (each written over columns its own warp had read)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16_ts(t_dv, t_s + (kk >> 2) * 64 + (kk & 3) * 8, bwd_desc<RB>(do_addr + kk * 2 * AB, 0, AB), idesc_ts,
                         (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16_ts(t_dk, t_dp + (kk >> 2) * 64 + (kk & 3) * 8, bwd_desc<RB>(q_addr + kk * 2 * AB, 0, AB), idesc_ts,
                         (i > 0 || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
        // next tile's S^T / dP^T: in order behind dV / dK, which are the last readers of the aliased columns
        if (t + 1 < total_tiles) issue_sdp(t + 1);
        if (first_kt) {
          mbar_wait(&dq_empty[slot], ((u >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        UCF_TL(t, 11);
        // dQ_t (+)= dS K_j: A = dS^T from shared memory read MN-major (rows = keys), B = K_j MN-major
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < T / 16; ++kk)
            umma_bf16(t_dq + slot * 64, umma_smem_desc(ds_addr + kk * 2048, 16384, 1024), bwd_desc<RB>(k_addr + kk * 2 * AB, 0, AB),
                      idesc_dq, (!first_kt || kk > 0) ? 1u : 0u);
          umma_commit(&dq_full[db]);          // every product of tile t has retired
          if (last_kt) umma_commit(&qdo_empty[slot]);
          if (i + 1 == nq) {
            umma_commit(dkv_full);
            umma_commit(&kv_empty[n & 1]);
          }
        }
        __syncwarp();
        UCF_TL(t, 12);
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps
    // 8 warps: two per TMEM lane quarter; warp (qd, half) owns KEY rows qd*32.. and the QUERY columns
    // [half*64, half*64+64) of S^T / dP^T, and the head-dim columns [half*HD/2, ...) of dQ / dK / dV.
    const int cw = warp - 2;
    const int qd = warp & 3;
    const int half = cw >> 2;
    const int row = qd * 32 + lane;
    const int ctid = threadIdx.x - 64;                          // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t row_sw = row & 7;
    const uint32_t lrow_sw = lane & 7;
    constexpr int HH = HD / 2;
    const uint32_t blk_off = half * 16384 + row * 128;          // this thread's key row in its 64-query block of dS^T
    const uint32_t stage_off = half * 16384 + qd * 4096;
    const float LOG2E = 1.4426950408889634f;

    auto flush_dq = [&](uint32_t t, uint32_t u, bool drain, int b, int h, int qrow0) {
      const uint32_t db = t & 1, slot = u & 1;
      mbar_wait(&dq_full[db], (t >> 1) & 1);
      if (!drain) return;
      tc_fence_after();
      if (threadIdx.x == 64) UCF_TL(t, 4);
      uint8_t* my_dq = ds_s + db * PSB + stage_off;
      uint32_t v[HH];
      if (HH == 32) tmem_ld32(t_dq + slot * 64 + lane_addr + half * HH, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      else tmem_ld16(t_dq + slot * 64 + lane_addr + half * HH, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      tmem_wait_ld();
      if (SHORT) {
#pragma unroll
        for (int g = 0; g < HH / 8; ++g) {
          uint8_t* dst = (HH == 32) ? my_dq + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4)
                                    : my_dq + lane * 32 + ((static_cast<uint32_t>(g) ^ ((lane >> 2) & 1)) << 4);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
        }
      } else {
#pragma unroll
        for (int g = 0; g < HH / 4; ++g) {
          uint8_t* dst = (HH == 32) ? my_dq + lane * 128 + ((static_cast<uint32_t>(g) ^ lrow_sw) << 4)
                                    : my_dq + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&dq_empty[slot]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && qrow0 + qd * 32 < p.Nq) {
        if (SHORT) tma_store_4d(&tmdQacc, my_dq, half * HH, h, qrow0 + qd * 32, b);
        else tma_reduce_add_4d(&tmdQacc, my_dq, half * HH, h, qrow0 + qd * 32, b);
        tma_store_commit();
      }
      if (threadIdx.x == 64) UCF_TL(t, 5);
    };
    auto flush_dkv = [&](uint32_t n, uint32_t t_last, int b, int h, int k0) {
      mbar_wait(dkv_full, n & 1);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      uint8_t* st_dv = ds_s + (t_last & 1) * PSB + stage_off;
      uint8_t* st_dk = st_dv + 2048;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint8_t* st = which == 0 ? st_dv : st_dk;
        const uint32_t t_src = (which == 0 ? t_dv : t_dk) + half * HH;
        uint32_t v[HH];
        if (HH == 32) tmem_ld32(t_src + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        else tmem_ld16(t_src + lane_addr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < HH / 8; ++g) {
          uint8_t* dst = (HH == 32) ? st + lane * 64 + ((static_cast<uint32_t>(g) ^ ((lane >> 1) & 3)) << 4)
                                    : st + lane * 32 + ((static_cast<uint32_t>(g) ^ ((lane >> 2) & 1)) << 4);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
        }
      }
      tc_fence_before();
      mbar_arrive_warp(dkv_empty);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && k0 + qd * 32 < p.Nk) {
        tma_store_4d(&tmdV, st_dv, half * HH, h, k0 + qd * 32, b);
        tma_store_4d(&tmdK, st_dk, half * HH, h, k0 + qd * 32, b);
        tma_store_commit();
      }
    };

    // statistics of query tile i of (b, h) -> stat_s[buf]: [0][q] = -lse*log2e (rows past Nq: -inf => P = 0),
    // [1][q] = -delta*scale; thread ctid covers one value
    auto stage_stats = [&](int buf, int b, int h, int i) {
      const int q = ctid & 127, which = ctid >> 7;
      const int qrow = i * T + q;
      float val = which == 0 ? -INFINITY : 0.f;
      if (qrow < p.Nq) {
        const long long idx = (static_cast<long long>(b) * p.H + h) * p.Nq + qrow;
        val = which == 0 ? -p.lse[idx] * LOG2E : -p.delta[idx] * p.scale;
      }
      stat_s[(buf * 2 + which) * T + q] = val;
    };

    if (total_tiles > 0) {
      int b0, h0, jt0;
      decode(0, b0, h0, jt0);
      stage_stats(0, b0, h0, 0);
    }
    named_bar_sync(1, 256);

    int pb_ = 0, ph_ = 0, pk0_ = 0;
    bool plast_ = true;
    uint32_t t = 0;
    for (uint32_t n = 0; n < static_cast<uint32_t>(my_items); ++n) {
      int b, h, jt;
      decode(n, b, h, jt);
      const int k0 = jt * T;
      const bool last_kt = !SHORT || jt == nkt - 1;
      for (int i = 0; i < nq; ++i, ++t) {
        const uint32_t db = t & 1;
        // statistics of the NEXT tile into the other buffer (published by the named barrier that ends this tile)
        if (t + 1 < total_tiles) {
          if (i + 1 < nq) stage_stats(db ^ 1, b, h, i + 1);
          else {
            int b2, h2, jt2;
            decode(n + 1, b2, h2, jt2);
            stage_stats(db ^ 1, b2, h2, 0);
          }
        }
        if (lane == 0) tma_store_wait_read<0>();       // staging rows (== dS^T rows) of earlier stores are free
        __syncwarp();
        mbar_wait(sdp_full, t & 1);
        tc_fence_after();
        if (threadIdx.x == 64) UCF_TL(t, 2);
        const float2 sl2 = mk2(p.scale_log2), sc2 = mk2(p.scale);
        const float* nl_s = stat_s + (db * 2 + 0) * T + half * 64;
        const float* nd_s = stat_s + (db * 2 + 1) * T + half * 64;
        uint8_t* ds_row = ds_s + db * PSB + blk_off;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld32(t_s + lane_addr + half * 64 + c * 32, sv);
          tmem_ld32(t_dp + lane_addr + half * 64 + c * 32, dv);
          tmem_wait_ld();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 nl4 = *reinterpret_cast<const float4*>(nl_s + c * 32 + e);     // broadcast reads
            const float4 nd4 = *reinterpret_cast<const float4*>(nd_s + c * 32 + e);
            const float2 ta = __ffma2_rn(make_float2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2, make_float2(nl4.x, nl4.y));
            const float2 tb = __ffma2_rn(make_float2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), sl2, make_float2(nl4.z, nl4.w));
            const float2 pa = make_float2(fast_ex2(ta.x), fast_ex2(ta.y));
            const float2 pb = make_float2(fast_ex2(tb.x), fast_ex2(tb.y));
            // dS = scale * P * (dP - delta)
            const float2 da = __fmul2_rn(pa, __ffma2_rn(make_float2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])), sc2, make_float2(nd4.x, nd4.y)));
            const float2 dbv = __fmul2_rn(pb, __ffma2_rn(make_float2(__uint_as_float(dv[e + 2]), __uint_as_float(dv[e + 3])), sc2, make_float2(nd4.z, nd4.w)));
            pk[e >> 1] = pack_bf16x2(pa.x, pa.y);
            pk[(e >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
            dk[e >> 1] = pack_bf16x2(da.x, da.y);
            dk[(e >> 1) + 1] = pack_bf16x2(dbv.x, dbv.y);
          }
          // bf16 pairs back over columns this thread has already read: queries [half*64 + c*32, +32) -> 16 columns
          tmem_st16(t_s + lane_addr + half * 64 + c * 16, pk);
          tmem_st16(t_dp + lane_addr + half * 64 + c * 16, dk);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t off = ((static_cast<uint32_t>(c * 4 + g)) ^ row_sw) << 4;
            *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive_warp(&pds_full[db]);
        if (threadIdx.x == 64) UCF_TL(t, 3);
        // drain what the tensor core finished while this tile's math ran
        if (t > 0) {
          if (i > 0) {
            flush_dq(t - 1, qinst(t - 1, n, i - 1), last_kt, b, h, (i - 1) * T);
          } else {
            flush_dq(t - 1, qinst(t - 1, n - 1, nq - 1), plast_, pb_, ph_, (nq - 1) * T);
            flush_dkv(n - 1, t - 1, pb_, ph_, pk0_);
          }
        }
        named_bar_sync(1, 256);      // next tile's statistics are visible; this tile's are no longer read
      }
      pb_ = b; ph_ = h; pk0_ = k0; plast_ = last_kt;
    }
    if (total_tiles > 0) {
      const uint32_t nl = static_cast<uint32_t>(my_items) - 1;
      flush_dq(total_tiles - 1, qinst(total_tiles - 1, nl, nq - 1), true, pb_, ph_, (nq - 1) * T);
      flush_dkv(nl, total_tiles - 1, pb_, ph_, pk0_);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]; (HD/8) lanes cooperate on one (b,n,h).
template <int HD>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                  int B, int H, int N, long long o_sb, long long o_sn, long long o_sh, long long do_sb,
                  long long do_sn, long long do_sh) {
  constexpr int G = HD / 8;
  const long long total = static_cast<long long>(B) * N * H;
  const long long gstride = (static_cast<long long>(gridDim.x) * blockDim.x) / G;
  const long long first = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  const int sub = threadIdx.x % G;
  const long long iters = (total + gstride - 1) / gstride;   // uniform trip count: shuffles stay converged
  for (long long it = 0; it < iters; ++it) {
    const long long item = first + it * gstride;
    const bool valid = item < total;
    float s = 0.f;
    long long bi = 0; int n = 0, hh = 0;
    if (valid) {
      hh = static_cast<int>(item % H);
      const long long r = item / H;
      n = static_cast<int>(r % N);
      bi = r / N;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + bi * o_sb + n * o_sn + hh * o_sh + sub * 8));
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(d_o + bi * do_sb + n * do_sn + hh * do_sh + sub * 8));
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
      s = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
    }
#pragma unroll
    for (int o2 = G / 2; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    if (valid && sub == 0) delta[(bi * H + hh) * N + n] = s;
  }
}

// dq[b,n,h,:] (bf16, strided) = (bf16) dq_acc[b,n,h,:] (fp32, contiguous [B,N,H,HD])
__global__ void __launch_bounds__(256)
attn_dq_cast_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int B, int H, int N, int hd,
                    long long sb, long long sn, long long sh) {
  const int per_head = hd / 8;
  const long long total = static_cast<long long>(B) * N * H * per_head;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int v = static_cast<int>(t % per_head);
    long long r = t / per_head;
    const int hh = static_cast<int>(r % H); r /= H;
    const int n = static_cast<int>(r % N);
    const long long bi = r / N;
    const float4 a = __ldg(reinterpret_cast<const float4*>(acc + t * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(acc + t * 8 + 4));
    *reinterpret_cast<uint4*>(dq + bi * sb + n * sn + hh * sh + v * 8) =
        make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(c.x, c.y), pack_bf16x2(c.z, c.w));
  }
}

}  // namespace ucf

using namespace ucf;

static long long* g_bwd_timeline = nullptr;
static int g_bwd_variant = 0;     // 0: transposed schedule (default); 1: round-1 kernel (profiling aid)
extern "C" void ucf_debug_set_attn_bwd_variant(int v) { g_bwd_variant = v; }
/* profiling aid (not part of the public header): device buffer of 64 int64 receiving clock64 stamps */
extern "C" void ucf_debug_set_attn_bwd_timeline(void* dev_ptr) { g_bwd_timeline = static_cast<long long*>(dev_ptr); }

static int attention_bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, void* dq, void* dk, void* dv, float* dq_acc, float* delta,
                                 int B, int H, int Nq, int Nk, int hd,
                                 long long q_sb, long long q_sn, long long q_sh,
                                 long long k_sb, long long k_sn, long long k_sh,
                                 long long v_sb, long long v_sn, long long v_sh,
                                 long long o_sb, long long o_sn, long long o_sh,
                                 long long dq_sb, long long dq_sn, long long dq_sh,
                                 long long dk_sb, long long dk_sn, long long dk_sh,
                                 long long dv_sb, long long dv_sn, long long dv_sh,
                                 float scale, void* stream, bool delta_ready);

extern "C" int ucf_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, void* dq, void* dk, void* dv, float* dq_acc, float* delta,
                                 int B, int H, int Nq, int Nk, int hd,
                                 long long q_sb, long long q_sn, long long q_sh,
                                 long long k_sb, long long k_sn, long long k_sh,
                                 long long v_sb, long long v_sn, long long v_sh,
                                 long long o_sb, long long o_sn, long long o_sh,
                                 long long dq_sb, long long dq_sn, long long dq_sh,
                                 long long dk_sb, long long dk_sn, long long dk_sh,
                                 long long dv_sb, long long dv_sn, long long dv_sh,
                                 float scale, void* stream) {
  return attention_bwd_impl(q, k, v, o, d_o, lse, dq, dk, dv, dq_acc, delta, B, H, Nq, Nk, hd, q_sb, q_sn, q_sh, k_sb, k_sn, k_sh,
                            v_sb, v_sn, v_sh, o_sb, o_sn, o_sh, dq_sb, dq_sn, dq_sh, dk_sb, dk_sn, dk_sh, dv_sb, dv_sn, dv_sh,
                            scale, stream, false);
}

extern "C" int ucf_attention_bwd_with_delta(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, void* dq, void* dk, void* dv, float* dq_acc, const float* delta,
                                 int B, int H, int Nq, int Nk, int hd,
                                 long long q_sb, long long q_sn, long long q_sh,
                                 long long k_sb, long long k_sn, long long k_sh,
                                 long long v_sb, long long v_sn, long long v_sh,
                                 long long o_sb, long long o_sn, long long o_sh,
                                 long long dq_sb, long long dq_sn, long long dq_sh,
                                 long long dk_sb, long long dk_sn, long long dk_sh,
                                 long long dv_sb, long long dv_sn, long long dv_sh,
                                 float scale, void* stream) {
  return attention_bwd_impl(q, k, v, o, d_o, lse, dq, dk, dv, dq_acc, const_cast<float*>(delta), B, H, Nq, Nk, hd, q_sb, q_sn, q_sh,
                            k_sb, k_sn, k_sh, v_sb, v_sn, v_sh, o_sb, o_sn, o_sh, dq_sb, dq_sn, dq_sh, dk_sb, dk_sn, dk_sh,
                            dv_sb, dv_sn, dv_sh, scale, stream, true);
}

static int attention_bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, void* dq, void* dk, void* dv, float* dq_acc, float* delta,
                                 int B, int H, int Nq, int Nk, int hd,
                                 long long q_sb, long long q_sn, long long q_sh,
                                 long long k_sb, long long k_sn, long long k_sh,
                                 long long v_sb, long long v_sn, long long v_sh,
                                 long long o_sb, long long o_sn, long long o_sh,
                                 long long dq_sb, long long dq_sn, long long dq_sh,
                                 long long dk_sb, long long dk_sn, long long dk_sh,
                                 long long dv_sb, long long dv_sn, long long dv_sh,
                                 float scale, void* stream, bool delta_ready) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk <= 0) { set_last_error("attention_bwd: empty problem"); return UCF_ERR_BAD_ARG; }
  if (hd != 64 && hd != 32) { set_last_error("attention_bwd: head_dim %d not supported (32 or 64)", hd); return UCF_ERR_UNSUPPORTED; }
  const bool short_q = Nq <= 256;      // dQ accumulates in tensor memory: dq_acc is not touched (may be NULL)
  if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv || (!dq_acc && !short_q) || !delta) {
    set_last_error("attention_bwd: null pointer"); return UCF_ERR_BAD_ARG;
  }
  const long long all_strides[] = {q_sb, q_sn, q_sh, k_sb, k_sn, k_sh, v_sb, v_sn, v_sh, o_sb, o_sn, o_sh,
                                   dq_sb, dq_sn, dq_sh, dk_sb, dk_sn, dk_sh, dv_sb, dv_sn, dv_sh};
  for (long long s : all_strides)
    if (s % 8) { set_last_error("attention_bwd: strides must be multiples of 8 elements"); return UCF_ERR_BAD_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  if (!short_q) {
    e = cudaMemsetAsync(dq_acc, 0, sizeof(float) * static_cast<size_t>(B) * Nq * H * hd, st);
    if (e != cudaSuccess) { set_last_error("attention_bwd: memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  }

  // d_o is addressed with o's strides (both are (B,N,H*hd) activations produced by this library)
  const long long items = static_cast<long long>(B) * Nq * H;
  if (!delta_ready) {
    const int G = hd / 8;
    long long threads = items * G;
    long long blocks = (threads + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (hd == 64)
      attn_delta_kernel<64><<<static_cast<int>(blocks), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta, B, H, Nq,
          o_sb, o_sn, o_sh, o_sb, o_sn, o_sh);
    else
      attn_delta_kernel<32><<<static_cast<int>(blocks), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(d_o), delta, B, H, Nq,
          o_sb, o_sn, o_sh, o_sb, o_sn, o_sh);
    int rc = check_launch("attn_delta_kernel");
    if (rc) return rc;
  }

  CUtensorMap tQ, tK, tV, tdO, tdQ, tdK, tdV;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = make_bnhd_tmap(&tQ, q, B, H, Nq, hd, q_sb, q_sn, q_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tK, k, B, H, Nk, hd, k_sb, k_sn, k_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tV, v, B, H, Nk, hd, v_sb, v_sn, v_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tdO, d_o, B, H, Nq, hd, o_sb, o_sn, o_sh, 128, bf, 2, hd))) return rc;
  // output boxes are 32 rows x hd/2 columns (each compute warp owns half of the head dimension)
  if (short_q) {
    if ((rc = make_bnhd_tmap(&tdQ, dq, B, H, Nq, hd, dq_sb, dq_sn, dq_sh, 32, bf, 2, hd / 2))) return rc;
  } else {
    if ((rc = make_bnhd_tmap(&tdQ, dq_acc, B, H, Nq, hd, static_cast<long long>(Nq) * H * hd, static_cast<long long>(H) * hd, hd,
                             32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, hd / 2))) return rc;
  }
  if ((rc = make_bnhd_tmap(&tdK, dk, B, H, Nk, hd, dk_sb, dk_sn, dk_sh, 32, bf, 2, hd / 2))) return rc;
  if ((rc = make_bnhd_tmap(&tdV, dv, B, H, Nk, hd, dv_sb, dv_sn, dv_sh, 32, bf, 2, hd / 2))) return rc;

  AttnBwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta;
  p.timeline = g_bwd_timeline;
  p.nkt = (Nk + 127) / 128;
  const long long n_items = static_cast<long long>(B) * H * (short_q ? 1 : p.nkt);
  if (n_items > 0x7fffffffLL) { set_last_error("attention_bwd: too many work items"); return UCF_ERR_BAD_ARG; }
  p.items = static_cast<int>(n_items);
  const int grid = p.items < num_sms() ? p.items : num_sms();
#define UCF_ATTN_BWD_LAUNCH(HD_, SHORT_)                                                                        \
  {                                                                                                             \
    static DeviceOnce once; bool& attr = once.flag();                                                           \
    if (!attr) {                                                                                                \
      e = cudaFuncSetAttribute(attn_bwd_kernel<HD_, SHORT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                               AttnBwdCfg<HD_>::SMEM_BYTES);                                                    \
      if (e != cudaSuccess) { set_last_error("attention_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; } \
      attr = true;                                                                                              \
    }                                                                                                           \
    attn_bwd_kernel<HD_, SHORT_><<<grid, 320, AttnBwdCfg<HD_>::SMEM_BYTES, st>>>(tQ, tK, tV, tdO, tdQ, tdK, tdV, p); \
  }
#define UCF_ATTN_BWD2_LAUNCH(HD_, SHORT_)                                                                       \
  {                                                                                                             \
    static DeviceOnce once; bool& attr = once.flag();                                                           \
    if (!attr) {                                                                                                \
      e = cudaFuncSetAttribute(attn_bwd2_kernel<HD_, SHORT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                               AttnBwd2Cfg<HD_>::SMEM_BYTES);                                                   \
      if (e != cudaSuccess) { set_last_error("attention_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; } \
      attr = true;                                                                                              \
    }                                                                                                           \
    attn_bwd2_kernel<HD_, SHORT_><<<grid, 320, AttnBwd2Cfg<HD_>::SMEM_BYTES, st>>>(tQ, tK, tV, tdO, tdQ, tdK, tdV, p); \
  }
  if (g_bwd_variant == 0) {
    if (hd == 64 && short_q) UCF_ATTN_BWD2_LAUNCH(64, true)
    else if (hd == 64) UCF_ATTN_BWD2_LAUNCH(64, false)
    else if (short_q) UCF_ATTN_BWD2_LAUNCH(32, true)
    else UCF_ATTN_BWD2_LAUNCH(32, false)
  } else if (hd == 64 && short_q) UCF_ATTN_BWD_LAUNCH(64, true)
  else if (hd == 64) UCF_ATTN_BWD_LAUNCH(64, false)
  else if (short_q) UCF_ATTN_BWD_LAUNCH(32, true)
  else UCF_ATTN_BWD_LAUNCH(32, false)
#undef UCF_ATTN_BWD_LAUNCH
#undef UCF_ATTN_BWD2_LAUNCH
  if ((rc = check_launch("attn_bwd_kernel"))) return rc;

  if (!short_q) {
    const long long total = static_cast<long long>(B) * Nq * H * (hd / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    attn_dq_cast_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(dq), B, H, Nq,
                                                                  hd, dq_sb, dq_sn, dq_sh);
    if ((rc = check_launch("attn_dq_cast_kernel"))) return rc;
  }
  return UCF_OK;
}
