// Whole transformer block per C call: ucf_block_fwd / ucf_block_bwd enqueue every kernel of
// Block.forward (/root/reference/src/UCF_VIT/simple/building_blocks.py:236-239) and of its hand-scheduled
// backward from C++, so the host pays ONE foreign call per block and direction instead of one per kernel
// (47 us of Python + ctypes each; 259 launches per ViT-B step were 12.3 ms of enqueue time).
// No new device code here: the kernels are the ones behind ucf_layernorm_*, ucf_gemm_bf16 and ucf_attention_*.
#include <algorithm>

#include "common.cuh"
#include "ucf_vit_b200.h"

namespace ucf {

// Split-K factor of the weight-gradient GEMM dW[n_out, k_in] = dY^T X over M tokens.  The persistent kernels hand
// tile x split work units to SMs (CTA-pair kernel: 256x256 tiles over SMs/2 clusters) in rounds, so the factor is
// picked to fill whole rounds: 36 tiles x 5 splits on 74 clusters is 2.43 rounds = 81 % busy, 36 x 4 is 1.95 = 97 %.
int wgrad_splits(int n_out, int k_in, long long M) {
  const long long kb = (M + 63) / 64;
  if (kb < 16) return 1;
  const int sms = num_sms();
  const bool pair = n_out >= 512 && k_in >= 512 && k_in <= 4096;      // mirrors the auto tile selection of ucf_gemm_bf16
  long long tiles;
  int workers;
  if (pair) { tiles = static_cast<long long>((n_out + 255) / 256) * ((k_in + 255) / 256); workers = std::max(1, sms / 2); }
  else      { tiles = static_cast<long long>((n_out + 127) / 128) * ((k_in + 255) / 256); workers = sms; }
  struct Cand { int s; long long units; double eff; };
  Cand cands[16];
  int nc = 0;
  double best_eff = 0.0;
  for (int s = 1; s <= 16; ++s) {
    if (kb / s < 8) break;
    const long long per = (kb + s - 1) / s;
    const long long units = tiles * ((kb + per - 1) / per);             // the launcher drops empty splits the same way
    const double eff = static_cast<double>(units) / static_cast<double>(((units + workers - 1) / workers) * workers);
    cands[nc++] = {s, units, eff};
    best_eff = std::max(best_eff, eff);
  }
  for (int i = 0; i < nc; ++i)      // fewest splits within 0.5 % of the best fill, two rounds if possible
    if (cands[i].eff >= best_eff - 0.005 && cands[i].units >= 2LL * workers) return cands[i].s;
  for (int i = 0; i < nc; ++i)
    if (cands[i].eff >= best_eff - 0.005) return cands[i].s;
  return 1;
}

static int check_block(const ucf_block_params* p, const char* who) {
  if (!p) { set_last_error("%s: null parameter block", who); return UCF_ERR_BAD_ARG; }
  if (p->B <= 0 || p->N <= 0 || p->D <= 0 || p->H <= 0 || p->hidden <= 0 || p->D % p->H != 0 || p->D % 8 != 0 ||
      p->hidden % 8 != 0) {
    set_last_error("%s: bad dimensions B=%d N=%d D=%d H=%d hidden=%d", who, p->B, p->N, p->D, p->H, p->hidden);
    return UCF_ERR_BAD_ARG;
  }
  const int hd = p->D / p->H;
  if (hd != 32 && hd != 64) {
    set_last_error("%s: head_dim %d not supported by the fused block (32 or 64)", who, hd);
    return UCF_ERR_UNSUPPORTED;
  }
  if (!p->n1_w || !p->n2_w || !p->qkv_w || !p->proj_w || !p->fc1_w || !p->fc2_w) {
    set_last_error("%s: null weight pointer", who);
    return UCF_ERR_BAD_ARG;
  }
  return UCF_OK;
}

// wgrad: dW[n_out, k_in] += dY^T X, db[n_out] += colsum(dY)
static int wgrad(const void* dy, const void* x, int n_out, int k_in, long long M, float* dw, float* db, void* st) {
  return ucf_gemm_bf16(dy, x, dw, nullptr, nullptr, n_out, k_in, static_cast<int>(M), n_out, k_in, k_in, 0,
                       UCF_LAYOUT_MN_MAJOR, UCF_LAYOUT_MN_MAJOR, UCF_EPI_F32_ADD, 0, wgrad_splits(n_out, k_in, M), 0, db, st);
}

}  // namespace ucf

using namespace ucf;

#define UCF_TRY(expr)            \
  do {                           \
    const int rc_ = (expr);      \
    if (rc_ != UCF_OK) return rc_; \
  } while (0)

extern "C" int ucf_wgrad_splits(int n_out, int k_in, long long M) { return ucf::wgrad_splits(n_out, k_in, M); }

extern "C" int ucf_block_fwd(const ucf_block_params* p, const ucf_block_acts* a, void* stream) {
  UCF_TRY(check_block(p, "block_fwd"));
  if (!a || !a->x || !a->h1 || !a->qkv || !a->o || !a->x1 || !a->h2 || !a->z || !a->u || !a->y || !a->mean1 || !a->rstd1 ||
      !a->mean2 || !a->rstd2 || !a->lse) {
    set_last_error("block_fwd: null activation pointer");
    return UCF_ERR_BAD_ARG;
  }
  const int B = p->B, N = p->N, D = p->D, H = p->H, Hd = p->hidden, hd = D / H;
  const long long M = static_cast<long long>(B) * N;
  if (M > 0x7fffffffLL) { set_last_error("block_fwd: too many tokens"); return UCF_ERR_BAD_ARG; }
  const int Mi = static_cast<int>(M);
  // fp32 masters -> this call's bf16 compute copies, one launch
  {
    const float* srcs[4];
    void* dsts[4];
    long long cnts[4];
    int n = 0;
    const void* masters[4] = {p->qkv_w_master, p->proj_w_master, p->fc1_w_master, p->fc2_w_master};
    const void* w16[4] = {p->qkv_w, p->proj_w, p->fc1_w, p->fc2_w};
    const long long sizes[4] = {3LL * D * D, 1LL * D * D, 1LL * Hd * D, 1LL * D * Hd};
    for (int i = 0; i < 4; ++i)
      if (masters[i]) { srcs[n] = static_cast<const float*>(masters[i]); dsts[n] = const_cast<void*>(w16[i]); cnts[n] = sizes[i]; ++n; }
    if (n) UCF_TRY(ucf_cast_f32_to_bf16_multi(n, srcs, dsts, cnts, stream));
  }
  UCF_TRY(ucf_layernorm_fwd(a->x, p->n1_w, p->n1_b, a->h1, a->mean1, a->rstd1, M, D, p->eps1, UCF_DTYPE_BF16, p->ln_dtype, stream));
  UCF_TRY(ucf_gemm_bf16(a->h1, p->qkv_w, a->qkv, p->qkv_b, nullptr, Mi, 3 * D, D, D, D, 3 * D, 0, UCF_LAYOUT_K_MAJOR,
                        UCF_LAYOUT_K_MAJOR, UCF_EPI_BIAS, p->bias_dtype, 1, 0, nullptr, stream));
  {
    const uint16_t* q = static_cast<const uint16_t*>(a->qkv);
    const long long sb = 3LL * N * D, sn = 3LL * D, sh = hd;
    UCF_TRY(ucf_attention_fwd(q, q + D, q + 2 * D, a->o, a->lse, B, H, N, N, hd, sb, sn, sh, sb, sn, sh, sb, sn, sh,
                              1LL * N * D, D, hd, 1.0f / sqrtf(static_cast<float>(hd)), stream));
  }
  UCF_TRY(ucf_gemm_bf16(a->o, p->proj_w, a->x1, p->proj_b, const_cast<void*>(a->x), Mi, D, D, D, D, D, D, UCF_LAYOUT_K_MAJOR,
                        UCF_LAYOUT_K_MAJOR, UCF_EPI_BIAS_RESIDUAL, p->bias_dtype, 1, 0, nullptr, stream));
  UCF_TRY(ucf_layernorm_fwd(a->x1, p->n2_w, p->n2_b, a->h2, a->mean2, a->rstd2, M, D, p->eps2, UCF_DTYPE_BF16, p->ln_dtype, stream));
  UCF_TRY(ucf_gemm_bf16(a->h2, p->fc1_w, a->u, p->fc1_b, a->z, Mi, Hd, D, D, D, Hd, Hd, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_K_MAJOR,
                        UCF_EPI_BIAS_GELU_AUX, p->bias_dtype, 1, 0, nullptr, stream));
  UCF_TRY(ucf_gemm_bf16(a->u, p->fc2_w, a->y, p->fc2_b, a->x1, Mi, D, Hd, Hd, Hd, D, D, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_K_MAJOR,
                        UCF_EPI_BIAS_RESIDUAL, p->bias_dtype, 1, 0, nullptr, stream));
  return UCF_OK;
}

extern "C" int ucf_block_bwd(const ucf_block_params* p, const ucf_block_acts* a, const ucf_block_grads* g, void* stream) {
  UCF_TRY(check_block(p, "block_bwd"));
  if (!a || !g || !a->x || !a->h1 || !a->qkv || !a->o || !a->x1 || !a->h2 || !a->z || !a->u || !a->mean1 || !a->rstd1 ||
      !a->mean2 || !a->rstd2 || !a->lse || !g->dy || !g->dx || !g->ws_a || !g->ws_b || !g->ws_c || !g->delta ||
      !g->g_qkv_w || !g->g_proj_w || !g->g_fc1_w || !g->g_fc2_w || !g->g_n1_w || !g->g_n2_w) {
    set_last_error("block_bwd: null pointer");
    return UCF_ERR_BAD_ARG;
  }
  const int B = p->B, N = p->N, D = p->D, H = p->H, Hd = p->hidden, hd = D / H;
  const long long M = static_cast<long long>(B) * N;
  const int Mi = static_cast<int>(M);
  if (N > 256 && !g->dq_acc) { set_last_error("block_bwd: N > 256 needs the dq_acc workspace"); return UCF_ERR_BAD_ARG; }
  void* dz = g->ws_a;     // [M, hidden]; later dqkv [M, 3D]
  void* dqkv = g->ws_a;
  void* dh = g->ws_b;     // dh2, then d_o, then dh1
  void* dx1 = g->ws_c;
  // ---- MLP
  UCF_TRY(wgrad(g->dy, a->u, D, Hd, M, g->g_fc2_w, g->g_fc2_b, stream));
  UCF_TRY(ucf_gemm_bf16(g->dy, p->fc2_w, dz, nullptr, a->z, Mi, Hd, D, D, Hd, Hd, Hd, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_MN_MAJOR,
                        UCF_EPI_DGELU, 0, 1, 0, nullptr, stream));
  UCF_TRY(wgrad(dz, a->h2, Hd, D, M, g->g_fc1_w, g->g_fc1_b, stream));
  UCF_TRY(ucf_gemm_bf16(dz, p->fc1_w, dh, nullptr, nullptr, Mi, D, Hd, Hd, D, D, 0, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_MN_MAJOR,
                        UCF_EPI_BIAS, 0, 1, 0, nullptr, stream));
  UCF_TRY(ucf_layernorm_bwd(dh, a->x1, p->n2_w, a->mean2, a->rstd2, g->dy, dx1, g->g_n2_w, g->g_n2_b, M, D, p->ln_dtype, stream));
  // ---- attention
  UCF_TRY(wgrad(dx1, a->o, D, D, M, g->g_proj_w, g->g_proj_b, stream));
  // d_o = dx1 * Wproj; where the fused kernel applies, its epilogue also leaves delta = rowsum(d_o o O) per head for the
  // attention backward (otherwise ucf_attention_bwd runs its own pass over d_o and O)
  const bool fused_delta = ucf_gemm_dgrad_delta_supported(Mi, D, D, H) != 0;
  if (fused_delta) {
    UCF_TRY(ucf_gemm_dgrad_delta(dx1, p->proj_w, dh, a->o, g->delta, Mi, D, D, D, D, D, D, N, H, stream));
  } else {
    UCF_TRY(ucf_gemm_bf16(dx1, p->proj_w, dh, nullptr, nullptr, Mi, D, D, D, D, D, 0, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_MN_MAJOR,
                          UCF_EPI_BIAS, 0, 1, 0, nullptr, stream));
  }
  {
    const uint16_t* q = static_cast<const uint16_t*>(a->qkv);
    uint16_t* dq = static_cast<uint16_t*>(dqkv);
    const long long sb = 3LL * N * D, sn = 3LL * D, sh = hd;
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    if (fused_delta) {
      UCF_TRY(ucf_attention_bwd_with_delta(q, q + D, q + 2 * D, a->o, dh, a->lse, dq, dq + D, dq + 2 * D, g->dq_acc, g->delta, B, H,
                                           N, N, hd, sb, sn, sh, sb, sn, sh, sb, sn, sh, 1LL * N * D, D, hd, sb, sn, sh, sb, sn, sh,
                                           sb, sn, sh, scale, stream));
    } else {
      UCF_TRY(ucf_attention_bwd(q, q + D, q + 2 * D, a->o, dh, a->lse, dq, dq + D, dq + 2 * D, g->dq_acc, g->delta, B, H, N, N, hd,
                                sb, sn, sh, sb, sn, sh, sb, sn, sh, 1LL * N * D, D, hd, sb, sn, sh, sb, sn, sh, sb, sn, sh,
                                scale, stream));
    }
  }
  UCF_TRY(wgrad(dqkv, a->h1, 3 * D, D, M, g->g_qkv_w, g->g_qkv_b, stream));
  UCF_TRY(ucf_gemm_bf16(dqkv, p->qkv_w, dh, nullptr, nullptr, Mi, D, 3 * D, 3 * D, D, D, 0, UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_MN_MAJOR,
                        UCF_EPI_BIAS, 0, 1, 0, nullptr, stream));
  UCF_TRY(ucf_layernorm_bwd(dh, a->x, p->n1_w, a->mean1, a->rstd1, dx1, g->dx, g->g_n1_w, g->g_n1_b, M, D, p->ln_dtype, stream));
  return UCF_OK;
}
