// Fused softmax attention forward on tcgen05 (sm_100a):  O = softmax(Q K^T * scale) V.
// Replaces the FLASH / CK / DEFAULT / NONE branches of Attention.forward
// (/root/reference/src/UCF_VIT/simple/building_blocks.py:163-187): non-causal, no mask, dropout 0.
//
// One CTA per (128-query tile, head, batch element); two CTAs co-reside per SM.
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (128 keys) through a 3-slot ring
//   warp 1      tcgen05.mma issuer: S = Q K_j^T -> TMEM[0,128), PV = P_j V_j -> TMEM[128,128+HD)
//   warps 2..5  softmax: thread == query row; online max/sum in registers (exp2 domain), P_j is
//               written as bf16 into 128B-swizzled shared memory (the A operand of the PV MMA),
//               O accumulates in registers and is rescaled when the running max moves.
// q/k/v are read in place from the packed qkv projection through 4-D tensor maps
// {hd, H, N, B}; rows past N are zero-filled by TMA and masked to -inf here.
#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

struct AttnFwdParams {
  int B, H, Nq, Nk;
  float scale_log2;   // scale * log2(e)
  float* lse;         // [B, H, Nq]
};

template <int HD>
struct AttnFwdCfg {
  static constexpr int BQ = 128, BKV = 128;
  static constexpr int Q_BYTES = BQ * HD * 2;
  static constexpr int KV_BYTES = BKV * HD * 2;
  static constexpr int P_BYTES = BQ * BKV * 2;
  static constexpr int RING = 3;
  static constexpr int SMEM_BYTES = 1024 + Q_BYTES + RING * KV_BYTES + P_BYTES + 16 * 8 + 16;
  static constexpr int TMEM_COLS = 256;   // S: [0,128)  PV: [128, 128+HD)
  static constexpr int ROW_BYTES = HD * 2;                 // 128 (HD=64) or 64 (HD=32)
  static constexpr int ATOM_BYTES = 8 * ROW_BYTES;         // swizzle atom: 8 rows
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
__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int BQ = Cfg::BQ, BKV = Cfg::BKV, RING = Cfg::RING;
  constexpr int RB = Cfg::ROW_BYTES, AB = Cfg::ATOM_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* kv_s = q_s + Cfg::Q_BYTES;
  uint8_t* p_s = kv_s + RING * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + Cfg::P_BYTES);
  uint64_t* q_full = bars;           // 1
  uint64_t* kv_full = bars + 1;      // RING
  uint64_t* kv_empty = bars + 4;     // RING
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* pv_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Nk + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < RING; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_4d(q_s, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < nkv; ++j) {
        for (int t = 0; t < 2; ++t) {
          const int r = 2 * j + t, slot = r % RING;
          mbar_wait(&kv_empty[slot], ((r / RING) & 1) ^ 1);
          mbar_expect_tx(&kv_full[slot], Cfg::KV_BYTES);
          tma_load_4d(kv_s + slot * Cfg::KV_BYTES, t == 0 ? &tmK : &tmV, &kv_full[slot], 0, h, j * BKV, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, false, false);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, false, true);
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(q_s), p_addr = smem_u32(p_s);
      for (int j = 0; j < nkv; ++j) {
        const int rk = 2 * j, rv = 2 * j + 1;
        const int sk = rk % RING, sv = rv % RING;
        // ---- S = Q K_j^T
        mbar_wait(&kv_full[sk], (rk / RING) & 1);
        mbar_wait(s_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(kv_s + sk * Cfg::KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_s, attn_desc<RB>(q_addr + k * 32, 16, AB), attn_desc<RB>(k_addr + k * 32, 16, AB),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&kv_empty[sk]);
        // ---- PV = P_j V_j   (A = P: K-major, 2 column blocks of 64 keys; B = V: MN-major)
        mbar_wait(&kv_full[sv], (rv / RING) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(kv_s + sv * Cfg::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t adesc = umma_smem_desc(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
          // V tile: rows = keys (RB bytes each), 8-key atoms AB apart; 16 keys per MMA = 2 atoms
          const uint64_t bdesc = attn_desc<RB>(v_addr + kk * 2 * AB, 0, AB);
          umma_bf16(tmem_pv, adesc, bdesc, idesc_pv, kk > 0 ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(&kv_empty[sv]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t row_sw = row & 7;
    uint8_t* p_row = p_s + row * 128;
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * BKV;
      // pass 1: row maximum
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (kbase + c * 32 + e < p.Nk) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
      }
      const float m_new = fmaxf(m_run, m_tile * p.scale_log2);   // Nk >= 1 and tile non-empty => finite
      const float alpha = exp2f(m_run - m_new);                   // first tile: exp2(-inf) = 0
      // previous tile's PV: fold into the register accumulator (also frees the P buffer)
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_pv + lane_addr + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) o_acc[c * 32 + e] = o_acc[c * 32 + e] * alpha_prev + __uint_as_float(v[e]);
        }
      }
      // pass 2: probabilities -> bf16 -> swizzled smem
      float l_tile = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = exp2f(__uint_as_float(v[e]) * p.scale_log2 - m_new);
          float p1 = exp2f(__uint_as_float(v[e + 1]) * p.scale_log2 - m_new);
          if (kbase + c * 32 + e >= p.Nk) p0 = 0.f;
          if (kbase + c * 32 + e + 1 >= p.Nk) p1 = 0.f;
          // sum what the MMA will actually see (bf16-rounded) so rows normalise exactly
          const uint32_t packed = pack_bf16x2(p0, p1);
          const float2 back = unpack_bf16x2(packed);
          l_tile += back.x + back.y;
          pk[e >> 1] = packed;
        }
        uint8_t* blk = p_row + (c >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + g);
          *reinterpret_cast<uint4*>(blk + ((chunk ^ row_sw) << 4)) =
              make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);        // S may be overwritten by the next QK^T
      fence_proxy_async_smem();
      mbar_arrive(p_full);         // P_j visible to the tensor core
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      alpha_prev = alpha;
    }
    // last PV
    mbar_wait(pv_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    // O staging reuses the P buffer (all PV MMAs have completed): 32 rows x 128 B per warp
    uint8_t* o_stage = p_s + (warp - 2) * 4096;
    const uint32_t lrow_sw = lane & 7;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_pv + lane_addr + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          f[e] = (o_acc[c * 32 + g * 8 + e] * alpha_prev + __uint_as_float(v[g * 8 + e])) * inv_l;
        const uint32_t chunk = static_cast<uint32_t>(c * 4 + g);
        uint8_t* dst = (RB == 128) ? o_stage + lane * 128 + ((chunk ^ lrow_sw) << 4)
                                   : o_stage + lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                    pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
    }
    if (q0 + row < p.Nq)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + q0 + row] = (m_run + log2f(l_run)) * 0.69314718055994531f;
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && q0 + qd * 32 < p.Nq) {
      tma_store_4d(&tmO, o_stage, 0, h, q0 + qd * 32, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
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
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
  return make_tmap(tm, ptr, dt, 4, dims, strides, box, swz);
}

static bool strides_ok(long long sb, long long sn, long long sh) {
  return sb % 8 == 0 && sn % 8 == 0 && sh % 8 == 0;
}

}  // namespace ucf

using namespace ucf;

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
  if (H > 65535 || B > 65535) { set_last_error("attention_fwd: H and B must be <= 65535"); return UCF_ERR_BAD_ARG; }
  CUtensorMap tQ, tK, tV, tO;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = make_bnhd_tmap(&tQ, q, B, H, Nq, hd, q_sb, q_sn, q_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tK, k, B, H, Nk, hd, k_sb, k_sn, k_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tV, v, B, H, Nk, hd, v_sb, v_sn, v_sh, 128, bf, 2, hd))) return rc;
  if ((rc = make_bnhd_tmap(&tO, o, B, H, Nq, hd, o_sb, o_sn, o_sh, 32, bf, 2, hd))) return rc;
  AttnFwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  dim3 grid((Nq + 127) / 128, H, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (hd == 64) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           AttnFwdCfg<64>::SMEM_BYTES);
      if (e != cudaSuccess) { set_last_error("attention_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      attr = true;
    }
    attn_fwd_kernel<64><<<grid, 192, AttnFwdCfg<64>::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
  } else {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           AttnFwdCfg<32>::SMEM_BYTES);
      if (e != cudaSuccess) { set_last_error("attention_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      attr = true;
    }
    attn_fwd_kernel<32><<<grid, 192, AttnFwdCfg<32>::SMEM_BYTES, st>>>(tQ, tK, tV, tO, p);
  }
  return check_launch("attn_fwd_kernel");
}
