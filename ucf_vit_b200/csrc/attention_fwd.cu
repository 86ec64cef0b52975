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
  int stagger_cycles;    // start offset of warpgroup 1
  int pair_bh;           // Nq <= 128: the two warpgroups take the single query tile of two consecutive (b, h)
};

// shared-memory descriptor for a tile whose rows are ROW_BYTES wide (128 -> SWIZZLE_128B,
// 64 -> SWIZZLE_64B).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t attn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = umma_smem_desc(saddr, lbo, sbo);
  if (ROW_BYTES == 64) d = (d & ~(7ull << 61)) | (4ull << 61);   // SWIZZLE_64B
  return d;
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
//   * P never touches shared memory: the bf16 probabilities go back into tensor memory (tcgen05.st, 64 columns
//     beside S and O) and PV is issued with its A operand FROM TENSOR MEMORY.  An SS-form M=128 N=64 K=16 MMA
//     streams 6 KB of operands out of shared memory = 48 cycles at 128 B/clk against 32 cycles of tensor work
//     (scripts/micro/umma_rate.cu: SS 48.1 cycles, TS 32.1), and the P stores competed with the tensor core for
//     that same shared-memory bandwidth (~256 KB per tile pair in all; 128 KB now).
//   * the MMA warp walks ONE continuous stream of key tiles across work items: S_g(next tile) -- also the
//     first tile of the NEXT item -- is issued right behind PV_g(this tile), so each warpgroup's
//     softmax -> PV -> S chain never drains at an item boundary, and warpgroup 1 is started half a
//     softmax period late once (the offset persists: nothing couples the two chains) so that one
//     warpgroup's MMAs run under the other's exponentials.
//   * a fraction of the exponentials is evaluated on the FMA pipe (Cody-Waite + degree-3 minimax
//     polynomial, rel. error 7.5e-5, far below bf16 rounding of P) to unload the 16-lane MUFU unit.
// Same barriers, producer and tensor maps as the general kernel above; shared memory: Q x4, a 6-slot K/V ring, O staging.
// ---------------------------------------------------------------------------------------------
constexpr float kRescaleThreshold = 60.0f;

template <int HD>
struct AttnFwdCfg {
  static constexpr int BQ = 128, BKV = 128;
  static constexpr int Q_BYTES = BQ * HD * 2;          // one query tile == one O tile
  static constexpr int KV_BYTES = BKV * HD * 2;
  static constexpr int RING = 6;                       // K/V tiles in flight
  static constexpr int NBAR = 2 + 2 + 2 * RING + 8;
  static constexpr int SMEM_BYTES = 1024 + 4 * Q_BYTES + RING * KV_BYTES + 2 * Q_BYTES + NBAR * 8 + 16;
  static constexpr int TMEM_COLS = 512;                // WG g: S at g*256 (128), O at +128 (HD), P (bf16 pairs) at +192 (64)
  static constexpr int ROW_BYTES = HD * 2;
  static constexpr int ATOM_BYTES = 8 * ROW_BYTES;
  static constexpr int THREADS = 64 + 256;             // producer, MMA issuer, 8 softmax warps
  static_assert(SMEM_BYTES <= 232448, "smem");
};

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

// profiling aid of the single-sweep kernel: clock64 stamps of CTA 0's first 24 key tiles per warpgroup, 8 slots per (tile, warpgroup)
// (compiled in with -DUCF_ATTN_TIMELINE only: the unconditional clock reads cost ~7 % of the softmax warps' issue slots)
#ifdef UCF_ATTN_TIMELINE
#define UCF_F2TL(t_, g_, idx_)                                                                          \
  do {                                                                                                  \
    if (p.timeline && blockIdx.x == 0 && (t_) < 24) p.timeline[((t_) * 2 + (g_)) * 8 + (idx_)] = clock64(); \
  } while (0)
#else
#define UCF_F2TL(t_, g_, idx_) do { } while (0)
#endif

// bf16 pair from two fp32 probabilities.  TRUNC = false: cvt.rn (F2FP.BF16.PACK_AB).  TRUNC = true: one PRMT that keeps
// the upper halves (round toward zero); the caller pre-multiplies p by (1 + 2^-9) through the exponent (kTruncBias),
// which turns truncation into round-half-up to within a quarter ulp and leaves no first-order bias in P / l.
constexpr float kTruncBias = 0.0028150156f;     // log2(1 + 2^-9)
template <bool TRUNC>
__device__ __forceinline__ uint32_t pack_p(float2 pp) {
  if (TRUNC) return __byte_perm(__float_as_uint(pp.x), __float_as_uint(pp.y), 0x7632);
  return pack_bf16x2(pp.x, pp.y);
}

template <int HD, int POLY, bool TRUNC>     // POLY: pairs (of 16) per 32-column chunk whose exp2 runs on the FMA pipe; TRUNC: see pack_p
__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV, RING = Cfg::RING;
  constexpr int RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                  // [2 buffers][2 tiles]
  uint8_t* kv_s = q_s + 4 * Cfg::Q_BYTES;               // [RING]
  uint8_t* o_s = kv_s + RING * Cfg::KV_BYTES;           // [2 warpgroups] O staging for the TMA store
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_s + 2 * Cfg::Q_BYTES);
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

  // Work item -> (batch, head, first query row) of warpgroup g; returns whether that warpgroup has rows to compute.
  // Normal mode: the two warpgroups take the two 128-query tiles of one (b, h, tile pair) and share its K/V tiles.
  // pair_bh mode (Nq <= 128, i.e. one query tile per (b, h) -- the MAE encoder's 49 kept tokens): they take the tiles of two
  // CONSECUTIVE (b, h), each with its own K/V tiles in the ring (kvsets = 2), so that warpgroup 1 is not idle and the two
  // items' load -> S -> softmax -> PV -> store chains overlap.
  const int kvsets = p.pair_bh ? 2 : 1;
  auto wg_coords = [&](int item, int g, int& b, int& h, int& q0) -> bool {
    if (p.pair_bh) {
      int bh = item * 2 + g;
      const bool ok = bh < p.B * p.H;
      if (!ok) bh = item * 2;        // an idle warpgroup's loads repeat its neighbour's tiles; nothing is computed or stored
      h = bh % p.H;
      b = bh / p.H;
      q0 = 0;
      return ok;
    }
    const int qp = item % p.nqp;
    const int bh = item / p.nqp;
    h = bh % p.H;
    b = bh / p.H;
    q0 = qp * 2 * BQ + g * BQ;
    return q0 < p.Nq;
  };
  const int my_items = (p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto item_of = [&](int k) { return static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x); };
  auto wg1_valid = [&](int k) { int b, h, q0; return wg_coords(item_of(k), 1, b, h, q0); };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane issues)
    uint32_t r = 0;
    for (int k = 0; k < my_items; ++k) {
      int bb[2], hh[2], qq[2];
      wg_coords(item_of(k), 0, bb[0], hh[0], qq[0]);
      wg_coords(item_of(k), 1, bb[1], hh[1], qq[1]);
      const int qb = k & 1;
      mbar_wait(&q_empty[qb], ((k >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[qb], 2 * Cfg::Q_BYTES);
        tma_load_4d(q_s + (qb * 2 + 0) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, hh[0], qq[0], bb[0]);
        tma_load_4d(q_s + (qb * 2 + 1) * Cfg::Q_BYTES, &tmQ, &q_full[qb], 0, hh[1], qq[1], bb[1]);
      }
      __syncwarp();
      for (int j = 0; j < nkv; ++j) {
        for (int set = 0; set < kvsets; ++set) {
          for (int t = 0; t < 2; ++t, ++r) {
            const int slot = r % RING;
            mbar_wait(&kv_empty[slot], ((r / RING) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&kv_full[slot], Cfg::KV_BYTES);
              tma_load_4d(kv_s + slot * Cfg::KV_BYTES, t == 0 ? &tmK : &tmV, &kv_full[slot], 0, hh[set], j * BKV, bb[set]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one continuous stream of key tiles
    if (my_items > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, false, true);
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

      // ring position of the current tile's K of K/V set 0 (its V follows at r + 1; set 1, pair_bh mode, at r + 2 / r + 3)
      uint32_t r = 0;
      const uint32_t tile_slots = 2 * kvsets;
      // prologue: S of the first tile; warpgroup 1 starts `stagger_cycles` late, once
      {
        mbar_wait(&q_full[0], 0);
        mbar_wait(&kv_full[0], 0);
        if (kvsets == 2) mbar_wait(&kv_full[2], 0);
        tc_fence_after();
        issue_s(0, 0, smem_u32(kv_s));
        if (wg1_valid(0)) {
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger_cycles) { }
          issue_s(1, 0, smem_u32(kv_s + (kvsets == 2 ? 2 : 0) * Cfg::KV_BYTES));
        }
        if (elect_one()) {
          umma_commit(&kv_empty[0]);
          if (kvsets == 2) umma_commit(&kv_empty[2]);
        }
        __syncwarp();
      }
      for (int k = 0; k < my_items; ++k) {
        const int ng = wg1_valid(k) ? 2 : 1;
        for (int j = 0; j < nkv; ++j, r += tile_slots) {
          const bool last_tile = j + 1 == nkv;
          const bool has_next = !(last_tile && k + 1 == my_items);
          const int kn = last_tile ? k + 1 : k;                 // item of the next tile
          int svs[2], sks[2] = {0, 0}, ng_next = 0;
          uint32_t v_addrs[2], k_addrs[2] = {0, 0};
          for (int set = 0; set < kvsets; ++set) {
            const uint32_t rv = r + 2 * set + 1;
            svs[set] = rv % RING;
            mbar_wait(&kv_full[svs[set]], (rv / RING) & 1);
            v_addrs[set] = smem_u32(kv_s + svs[set] * Cfg::KV_BYTES);
          }
          if (has_next) {
            if (last_tile) mbar_wait(&q_full[kn & 1], (kn >> 1) & 1);
            for (int set = 0; set < kvsets; ++set) {
              const uint32_t rk = r + tile_slots + 2 * set;
              sks[set] = rk % RING;
              mbar_wait(&kv_full[sks[set]], (rk / RING) & 1);
              k_addrs[set] = smem_u32(kv_s + sks[set] * Cfg::KV_BYTES);
            }
            ng_next = last_tile ? (wg1_valid(kn) ? 2 : 1) : ng;
          }
          if (kvsets == 1) { v_addrs[1] = v_addrs[0]; k_addrs[1] = k_addrs[0]; }
          const int ksteps = last_tile ? ksteps_last : BKV / 16;
          for (int g = 0; g < 2; ++g) {
            const uint32_t v_addr = v_addrs[g], k_addr = k_addrs[g];
            if (g < ng) {
              mbar_wait(&p_full[g], pc[g] & 1);       // P_g(j) is in tensor memory and S_g(j) has been read
              UCF_F2TL(pc[g], g, 3);
              ++pc[g];
              tc_fence_after();
            }
            // S_g of the NEXT tile first: it does not depend on P, and it is what the warpgroup is waiting for
            if (has_next && g < ng_next) issue_s(g, kn, k_addr);
            if (g < ng) {
              const uint32_t d = tmem_base + g * 256 + 128;
              const uint32_t pt = tmem_base + g * 256 + 192;          // P_g: bf16 pairs, 8 columns per 16 keys
              if (elect_one()) {
                for (int kk = 0; kk < ksteps; ++kk)
                  umma_bf16_ts(d, pt + kk * 8, attn_desc<RB>(v_addr + kk * 2 * AB, 0, AB), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&pv_full[g]);             // PV_g(j) retired: P_g may be rewritten (and, last tile: O_g is complete)
              }
              __syncwarp();
            }
            UCF_F2TL(pc[g] - 1, g, 4);
          }
          if (elect_one()) {
            for (int set = 0; set < kvsets; ++set) {
              umma_commit(&kv_empty[svs[set]]);
              if (has_next) umma_commit(&kv_empty[sks[set]]);
            }
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
    const uint32_t tmem_s = tmem_base + g * 256, tmem_o = tmem_s + 128, tmem_p = tmem_s + 192;
    const uint32_t lrow_sw = lane & 7;
    uint8_t* o_stage = o_s + g * Cfg::Q_BYTES + qd * (32 * RB);   // this warp's 32 rows of the O tile
    uint32_t tc = 0;         // key tiles processed (s_full / pv_full parity)
    const float2 sc2 = mk2(p.scale_log2);
    const bool stamp_thread = (threadIdx.x & 127) == 64;

    for (int k = 0; k < my_items; ++k) {
      int b, h, qt0;                                // qt0: first query row of this warpgroup's tile
      if (!wg_coords(item_of(k), g, b, h, qt0)) continue;      // (only warpgroup 1 can be idle)
      // a warp whose 32 rows are all past Nq keeps the barrier protocol but does no work (its rows of P / O are
      // never stored: the TMA store clips them)
      const bool dead = qt0 + qd * 32 >= p.Nq;
      float m_used = 0.f, l_run = 0.f;

      for (int j = 0; j < nkv; ++j, ++tc) {
        mbar_wait(&s_full[g], tc & 1);
        tc_fence_after();
        if (stamp_thread) UCF_F2TL(tc, g, 0);
        if (!dead) {
          const int nvalid = j + 1 == nkv ? nvalid_last : BKV;
          const int nchunk = (nvalid + 31) >> 5;
          float mx;
          float2 l2;
          // one sweep over the tile: exponentials against m_used -> bf16 P in tensor memory; returns the row
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
            // TRUNC: P = trunc_bf16(p * (1 + 2^-9)); the factor rides on the exponent for free and is taken out of lse
            const float2 nm2 = mk2(TRUNC ? kTruncBias - m_used : -m_used);
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
                    pk[e >> 1] = pack_p<TRUNC>(pp);
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
                    pk[e >> 1] = pack_p<TRUNC>(pp);
                  }
                }
                if (c == 0 && j > 0) mbar_wait(&pv_full[g], (tc - 1) & 1);   // PV_g(j-1) no longer reads the P columns
                tmem_st16(tmem_p + lane_addr + c * 16, pk);
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
        if (stamp_thread) UCF_F2TL(tc, g, 1);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(&p_full[g]);         // P_g(j) is in tensor memory; S_g may be overwritten
        if (stamp_thread) UCF_F2TL(tc, g, 2);
      }
      // ---- epilogue: O / l  (the last PV of the item has retired; it was also the last reader of P)
      mbar_wait(&pv_full[g], (tc - 1) & 1);
      tc_fence_after();
      if (stamp_thread) UCF_F2TL(tc - 1, g, 5);
      if (!dead) {
        const float inv_l = 1.0f / l_run;
        if (lane == 0) tma_store_wait_read<0>();      // the previous item's O rows have left the staging area
        __syncwarp();
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
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qt0 + row] =
              (m_used + log2f(l_run) - (TRUNC ? kTruncBias : 0.f)) * 0.69314718055994531f;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmO, o_stage, 0, h, qt0 + qd * 32, b);
          tma_store_commit();
        }
      }
      if (stamp_thread) UCF_F2TL(tc - 1, g, 6);
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

static int g_fwd_stagger = -1;     // < 0: default
static int g_fwd_poly = 4;         // exp2 pairs per 16 evaluated on the FMA pipe (0, 4 or 8); profiling aid

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
  p.stagger_cycles = g_fwd_stagger >= 0 ? g_fwd_stagger : 600;
#define UCF_FWD(POLY_)                                                                                            \
  {                                                                                                               \
    static DeviceOnce once; bool& done = once.flag();                                                                                     \
    if ((rc = set_smem_attr(attn_fwd_kernel<HD, POLY_, true>, Cfg::SMEM_BYTES, done))) return rc;                 \
    attn_fwd_kernel<HD, POLY_, true><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);             \
  }
  if (g_fwd_poly == 0) UCF_FWD(0)
  else if (g_fwd_poly == 8) UCF_FWD(8)
  else UCF_FWD(4)
#undef UCF_FWD
  return check_launch("attn_fwd_kernel");
}

}  // namespace ucf

using namespace ucf;

static long long* g_fwd_timeline = nullptr;
/* profiling aid (not part of the public header): device buffer of 64 int64 receiving clock64 stamps */
extern "C" void ucf_debug_set_attn_fwd_timeline(void* dev_ptr) { g_fwd_timeline = static_cast<long long*>(dev_ptr); }
/* profiling aid: exp2 pairs per 16 evaluated by the polynomial on the FMA pipe (0, 4, 8) */
extern "C" void ucf_debug_set_attn_fwd_poly(int pairs) { ucf::g_fwd_poly = pairs; }
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
  p.pair_bh = (Nq <= 128 && static_cast<long long>(B) * H >= 2) ? 1 : 0;
  const long long items = p.pair_bh ? (static_cast<long long>(B) * H + 1) / 2 : static_cast<long long>(B) * H * p.nqp;
  if (items > 0x7fffffffLL) { set_last_error("attention_fwd: too many work items"); return UCF_ERR_BAD_ARG; }
  p.items = static_cast<int>(items);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.timeline = g_fwd_timeline;
  p.stagger_cycles = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return hd == 64 ? launch_attn_fwd<64>(tQ, tK, tV, tO, p, st) : launch_attn_fwd<32>(tQ, tK, tV, tO, p, st);
}
