/* ucf_vit_b200.h -- C ABI of the B200-native (sm_100a) ViT transformer-block training hot path.
 *
 * Drop-in boundary.  The reference (irlyngaas/UCF-VIT) is 100 % Python: its "operator interface"
 * for this path is the set of torch library calls made by
 *   src/UCF_VIT/simple/building_blocks.py  (PatchEmbed :30-92, Mlp :94-129, Attention :131-192,
 *                                           Block :194-239, VariableMapping_Attention :301-373)
 *   src/UCF_VIT/simple/arch.py             (VIT.forward_features :434-476, _pos_embed :367-393,
 *                                           MAE.random_masking :663-681)
 *   src/UCF_VIT/dataloaders/quadtree.py    (FixedQuadTree.serialize/deserialize :144-221)
 * There is no FFI in the reference; each entry point below names the reference call it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - return 0 on success; <0 = argument/driver error (UCF_ERR_*), >0 = cudaError_t of the launch
 *   - ucf_last_error() returns a thread-local message for the last non-zero return
 *   - launchers are re-entrant, hold no mutable global state besides lazily queried device
 *     attributes, allocate nothing: every workspace is caller-provided
 *   - bf16 storage is the raw 16-bit pattern (torch.bfloat16 / __nv_bfloat16)
 */
#ifndef UCF_VIT_B200_H_
#define UCF_VIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCF_ABI_VERSION 4

enum { UCF_DTYPE_F32 = 0, UCF_DTYPE_BF16 = 1, UCF_DTYPE_U8 = 2, UCF_DTYPE_F64 = 3, UCF_DTYPE_I64 = 4 };
enum { UCF_LAYOUT_K_MAJOR = 0, UCF_LAYOUT_MN_MAJOR = 1 };
/* GEMM epilogues */
enum {
  UCF_EPI_BIAS = 0,          /* C = acc (+ bias)                       bf16 out                    */
  UCF_EPI_BIAS_RESIDUAL = 1, /* C = acc (+ bias) + aux                 bf16 out, aux bf16 in       */
  UCF_EPI_BIAS_GELU_AUX = 2, /* aux = acc (+ bias); C = gelu_erf(aux)  bf16 out, aux bf16 out      */
  UCF_EPI_DGELU = 3,         /* C = acc * gelu_erf'(aux)               bf16 out, aux bf16 in       */
  UCF_EPI_F32_ADD = 4        /* C += acc   (fp32, split-K capable)     fp32 in/out                 */
};

int ucf_abi_version(void);
const char* ucf_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
unsigned long long ucf_launch_count(void);

/* ---- dense contractions: nn.Linear forward / dgrad / wgrad ---------------------------------
 * replaces torch.nn.Linear.forward + autograd at building_blocks.py:115-128 (Mlp.fc1/act/fc2),
 * :150,159 (Attention.qkv), :154,189 (Attention.proj), :321-373 (var-agg q/kv/proj) and the
 * Conv{2,3}d patch projection :58-60,89 once patches are laid out as rows (ucf_patchify_*).
 *
 *   C[M,N] = epilogue( A[M,K] * B[N,K]^T ), fp32 accumulation on tcgen05 tensor cores.
 * a_layout/b_layout: K_MAJOR  = operand stored [rows, K] with K contiguous (pitch ld, elements)
 *                    MN_MAJOR = operand stored [K, rows] with rows contiguous (pitch ld)
 * bias: length N, dtype bias_dtype, may be NULL.  splits: split-K factor (UCF_EPI_F32_ADD only).
 * bias_grad: wgrad form only (A MN-major + UCF_EPI_F32_ADD, i.e. C = dW += dY^T X): if non-NULL, fp32
 * [M] += row sums of A = column sums of dY -- the bias gradient, taken from the dY tiles already staged
 * in shared memory (replaces a separate reduction pass over dY).  NULL otherwise.
 * tile_n: 0 = auto, 128 / 256 = single-CTA kernels, 512 = CTA-pair kernel (tcgen05 cta_group::2,
 * 256 x 256 tile per 2-CTA cluster).  Pointers and pitches*elemsize must be 16-byte aligned. */
int ucf_gemm_bf16(const void* A, const void* B, void* C, const void* bias, void* aux,
                  int M, int N, int K, long long lda, long long ldb, long long ldc, long long ldaux,
                  int a_layout, int b_layout, int epilogue, int bias_dtype, int splits, int tile_n,
                  void* bias_grad, void* stream);

/* The attention-output dgrad with the attention backward's row statistic fused into its epilogue:
 *   dX[M, N] = dY[M, K] * W[K, N]   (W = Attention.proj.weight [out = K, in = N], read MN-major in place)
 *   delta[b, h, n] = sum_{d < N/heads} dX[b*tokens + n, h*hd + d] * O[b*tokens + n, h*hd + d]
 * i.e. `d_o = d_x1 @ proj.weight` of Attention.proj's backward (building_blocks.py:154,189) plus the
 * rowsum(dO o O) pass a flash-attention backward starts with (replaces a separate 2 x M x N read).  delta is taken from
 * the fp32 accumulator tile.  O: bf16 [M, N] (pitch ldo), delta: fp32 [M / tokens, heads, tokens].  Shapes the fused
 * kernel serves: ucf_gemm_dgrad_delta_supported() != 0 (head_dim 32 / 64, N a multiple of 128 in 512..4096, M >= 512). */
int ucf_gemm_dgrad_delta_supported(int M, int N, int K, int heads);
int ucf_gemm_dgrad_delta(const void* dY, const void* W, void* dX, const void* O, float* delta, int M, int N, int K,
                         long long lddy, long long ldw, long long lddx, long long ldo, int tokens, int heads, void* stream);

/* LayerNorm folded into the projection that consumes it (north-star "fused LayerNorm+QKV projection"; replaces the pair
 * `self.attn(self.norm1(x))` -> `self.qkv(x)`, building_blocks.py:236-237 -> :150,159, when the normalised activations are
 * not needed afterwards, i.e. in forward-only / no-grad execution -- training keeps them for the weight gradient):
 *   y[r, n] = rstd[r] * (sum_k x[r, k] * Wg[n, k] - mean[r] * colsum[n]) + bias_folded[n]   ==   LN(x) W^T + b
 * with Wg = W * diag(gamma) (bf16 [N, K]), colsum[n] = sum_k Wg[n, k], bias_folded = W beta + b (both fp32 [N]), and
 * mean / rstd from ucf_layernorm_stats (one read of x, 8 bytes out per row -- instead of a read and a write of x).
 * x: bf16 [M, K] raw rows; y: bf16 [M, N].  Shapes served: ucf_ln_gemm_supported() != 0. */
int ucf_layernorm_stats(const void* x, float* mean, float* rstd, long long rows, int D, float eps, int x_dtype, void* stream);
int ucf_ln_gemm_supported(int M, int N, int K);
int ucf_ln_gemm(const void* x, const void* w_gamma, void* y, const float* bias_folded, const float* colsum, const float* mean,
                const float* rstd, int M, int N, int K, long long ldx, long long ldw, long long ldy, void* stream);

/* ---- LayerNorm (replaces nn.LayerNorm at arch.py:170,266; building_blocks.py:212,226) ------
 * x: [rows, D] (x_dtype), gamma/beta: [D] (param_dtype, may be NULL = 1/0), y: [rows, D] bf16,
 * mean/rstd: [rows] fp32 (saved for backward).  D % 8 == 0, D <= 4096. */
int ucf_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                      float* rstd, long long rows, int D, float eps, int x_dtype, int param_dtype,
                      void* stream);
/* dx = LN'(dy) (+ dres if non-NULL: fuses the residual-branch gradient add); dgamma/dbeta fp32
 * [D] are ACCUMULATED into (caller zero-fills or carries .grad). x/dy/dres/dx are bf16. */
int ucf_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                      const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                      long long rows, int D, int param_dtype, void* stream);

/* ---- fused scaled-dot-product attention (replaces building_blocks.py:163-187) --------------
 * q,k,v,o: bf16, element (b, n, h, d) at  ptr[b*sb + n*sn + h*sh + d]  (d contiguous), so the
 * packed qkv buffer of Attention.qkv is consumed in place and o is written as (B,N,H*hd).
 * lse: [B,H,N] fp32 log-sum-exp of the scaled scores (saved for backward).  hd in {32,64}.
 * non-causal, no mask, no dropout (attn_drop = 0 in every reference config). */
int ucf_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                      int B, int H, int Nq, int Nk, int hd,
                      long long q_sb, long long q_sn, long long q_sh,
                      long long k_sb, long long k_sn, long long k_sh,
                      long long v_sb, long long v_sn, long long v_sh,
                      long long o_sb, long long o_sn, long long o_sh,
                      float scale, void* stream);
/* dq_acc: fp32 workspace [B,Nq,H,hd] (zero-filled by this call; unused and may be NULL when Nq <= 256,
 * where dQ accumulates in tensor memory), delta: fp32 [B,H,Nq] workspace.
 * dq/dk/dv: bf16, addressed like q/k/v (may alias a packed dqkv buffer). */
int ucf_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* lse, void* dq, void* dk, void* dv, float* dq_acc, float* delta,
                      int B, int H, int Nq, int Nk, int hd,
                      long long q_sb, long long q_sn, long long q_sh,
                      long long k_sb, long long k_sn, long long k_sh,
                      long long v_sb, long long v_sn, long long v_sh,
                      long long o_sb, long long o_sn, long long o_sh,
                      long long dq_sb, long long dq_sn, long long dq_sh,
                      long long dk_sb, long long dk_sn, long long dk_sh,
                      long long dv_sb, long long dv_sn, long long dv_sh,
                      float scale, void* stream);
/* The same with `delta` already holding rowsum(dO o O) (written by ucf_gemm_dgrad_delta): skips the delta pass. */
int ucf_attention_bwd_with_delta(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* lse, void* dq, void* dk, void* dv, float* dq_acc, const float* delta,
                      int B, int H, int Nq, int Nk, int hd,
                      long long q_sb, long long q_sn, long long q_sh,
                      long long k_sb, long long k_sn, long long k_sh,
                      long long v_sb, long long v_sn, long long v_sh,
                      long long o_sb, long long o_sn, long long o_sh,
                      long long dq_sb, long long dq_sn, long long dq_sh,
                      long long dk_sb, long long dk_sn, long long dk_sh,
                      long long dv_sb, long long dv_sn, long long dv_sh,
                      float scale, void* stream);

/* ---- one C call per transformer block and direction ------------------------------------------
 * Replaces Block.forward (simple/building_blocks.py:236-239: x + attn(norm1(x)), then + mlp(norm2(.))) and the
 * autograd graph torch builds for it, in the configuration every reference driver uses (no qk_norm, no LayerScale,
 * drop rates 0, GELU(erf), head_dim 32 or 64).  The kernels are the ones behind ucf_layernorm_*, ucf_gemm_bf16 and
 * ucf_attention_*; what these entry points remove is the per-kernel host cost of a Python caller (one foreign call
 * instead of 8 forward / 15 backward).  M = B*N tokens.  Every buffer is caller-provided. */
typedef struct {
  int B, N, D, H, hidden;        /* batch, tokens per sample, width, heads, MLP width; head_dim = D/H */
  float eps1, eps2;              /* LayerNorm epsilons */
  int ln_dtype, bias_dtype;      /* UCF_DTYPE_F32 | UCF_DTYPE_BF16 of the LayerNorm parameters / of the Linear biases */
  const void *n1_w, *n1_b, *n2_w, *n2_b;           /* [D] (biases may be NULL) */
  const void *qkv_b, *proj_b, *fc1_b, *fc2_b;      /* [3D], [D], [hidden], [D] or NULL */
  const void *qkv_w, *proj_w, *fc1_w, *fc2_w;      /* bf16 compute copies [3D,D], [D,D], [hidden,D], [D,hidden] */
  /* fp32 master weights or NULL.  Non-NULL: ucf_block_fwd first refreshes the bf16 copy above from it (the per-step
   * cast of mixed-precision training, one launch for all four); ucf_block_bwd ignores these. */
  const void *qkv_w_master, *proj_w_master, *fc1_w_master, *fc2_w_master;
} ucf_block_params;
typedef struct {                 /* activations: written by ucf_block_fwd, read by ucf_block_bwd (all but y) */
  const void* x;                 /* bf16 [M, D] block input */
  void *h1, *qkv, *o, *x1, *h2;  /* bf16 [M,D] LN1(x) | [M,3D] packed q,k,v | [M,D] attention out | [M,D] x + proj | [M,D] LN2(x1) */
  void *z, *u;                   /* bf16 [M,hidden] fc1 pre-activation | GELU(z) */
  void* y;                       /* bf16 [M, D] block output */
  float *mean1, *rstd1, *mean2, *rstd2;   /* fp32 [M] */
  float* lse;                    /* fp32 [B, H, N] */
} ucf_block_acts;
typedef struct {
  const void* dy;                /* bf16 [M, D] gradient of the block output */
  void* dx;                      /* bf16 [M, D] gradient of the block input (written) */
  /* fp32 gradients, ACCUMULATED into (caller zero-fills or carries .grad); the bias entries may be NULL */
  float *g_n1_w, *g_n1_b, *g_qkv_w, *g_qkv_b, *g_proj_w, *g_proj_b, *g_n2_w, *g_n2_b, *g_fc1_w, *g_fc1_b, *g_fc2_w, *g_fc2_b;
  void* ws_a;                    /* bf16 workspace, M * max(hidden, 3D) elements */
  void *ws_b, *ws_c;             /* bf16 workspaces, M * D elements each */
  float* dq_acc;                 /* fp32 [B,N,H,hd] workspace; may be NULL when N <= 256 */
  float* delta;                  /* fp32 [B,H,N] workspace */
} ucf_block_grads;
int ucf_block_fwd(const ucf_block_params* params, const ucf_block_acts* acts, void* stream);
int ucf_block_bwd(const ucf_block_params* params, const ucf_block_acts* acts, const ucf_block_grads* grads, void* stream);
/* split-K factor ucf_block_bwd uses for the weight-gradient GEMM dW[n_out, k_in] over M tokens (host arithmetic only) */
int ucf_wgrad_splits(int n_out, int k_in, long long M);

/* ---- channel variable-aggregation cross-attention core (replaces the SDPA call inside
 * VariableMapping_Attention.forward, building_blocks.py:339-367; Nq = Na (1), Nk = V) ---------
 * q: bf16 [Bq, Na, H, hd] with Bq == rows, or Bq == 1 when q_shared != 0 (one learned query for
 * every row); kv: bf16 [rows, V, 2, H, hd] (raw output of the kv projection); o: bf16
 * [rows, Na, H, hd]; lse: fp32 [rows, Na, H].  hd in {32, 64, 128}. */
int ucf_var_attention_fwd(const void* q, const void* kv, void* o, float* lse, long long rows, int Na,
                          int V, int H, int hd, int q_shared, float scale, void* stream);
/* dkv: bf16 like kv (fully written).  dq_acc: fp32 [Bq, Na, H, hd] (zero-filled here and summed
 * over rows when q_shared, else plainly written). */
int ucf_var_attention_bwd(const void* q, const void* kv, const void* o, const void* d_o,
                          const float* lse, void* dkv, float* dq_acc, long long rows, int Na, int V,
                          int H, int hd, int q_shared, float scale, void* stream);

/* ---- adaptive patching (SAP) --------------------------------------------------------------
 * Tree construction (HOST pointers; integer work, bit-exact with the reference's node order):
 * replaces FixedQuadTree._build_tree (dataloaders/quadtree.py:115-137) and FixedOctTree._build_tree
 * (dataloaders/octree.py:72-102).  domain: [n0, n1] (2-D: rows, cols) or cubic [n0, n1, n2] edge map
 * of dtype u8 / f32 / f64; leaf value = int(sum(domain[box]) / norm_factor).  Greedy: split the FIRST
 * leaf holding the maximum value until fixed_length leaves exist or the chosen leaf is 2 wide.
 * Like the reference, the last split may overshoot: the tree ends with fixed_length .. fixed_length + 2 (2-D) /
 * + 6 (3-D) leaves unless fixed_length = 1 (mod 3 / mod 7), so the buffers hold UCF_SAP_TREE_ROWS(fixed_length,
 * ndim) rows: boxes_host int32 [rows, 4] (x1,x2,y1,y2) or [rows, 6] (+z1,z2) in list order; values_host int64
 * [rows] or NULL.  Returns the number of leaves (>= 1) or a negative code. */
#define UCF_SAP_TREE_ROWS(fixed_length, ndim) ((fixed_length) + ((ndim) == 2 ? 3 : 7) - 1)
int ucf_sap_build_tree_host(const void* domain_host, int domain_dtype, int ndim, int n0, int n1, int n2,
                            double norm_factor, int fixed_length, int32_t* boxes_host,
                            long long* values_host);
/* The same for n_images edge maps of one shape on host threads (n_threads <= 0: one per core, at most one per
 * image): domains_host = HOST array of n_images HOST pointers; boxes_host int32 [n_images, rows, 4|6],
 * values_host int64 [n_images, rows] or NULL with rows = UCF_SAP_TREE_ROWS(fixed_length, ndim); n_leaves_host
 * int [n_images] (leaf count per image).
 * Returns 0, or the (negative) code of the first image that was rejected. */
int ucf_sap_build_tree_batch_host(const void* const* domains_host, int n_images, int domain_dtype, int ndim,
                                  int n0, int n1, int n2, double norm_factor, int fixed_length,
                                  int32_t* boxes_host, long long* values_host, int* n_leaves_host, int n_threads);
/* Gather = FixedQuadTree.serialize (quadtree.py:144-174; cv.resize INTER_CUBIC per leaf) /
 * FixedOctTree.serialize (octree.py:104-150; align-corners trilinear).  DEVICE pointers.
 * img: [n0, n1, C] u8|f32 (2-D, HWC) or [n0, n1, n2, C] f32 (3-D, ZYXC); boxes: int32 on device;
 * seq: f32 [fixed_length, p, p(, p), C]; seq_size: int64 [fixed_length]; seq_pos: f64
 * [fixed_length, 2|3] box centres.  Slots >= n_leaves are padded: zero patch, size 0, centre -1. */
int ucf_sap_gather(const void* img, int img_dtype, int ndim, int n0, int n1, int n2, int C,
                   const int32_t* boxes, int n_leaves, int fixed_length, int p, float* seq,
                   long long* seq_size, double* seq_pos, void* stream);
/* Scatter = FixedQuadTree.deserialize + Rect.set_area (quadtree.py:209-221, :25-36) /
 * FixedOctTree.deserialize + Cube.set_area (octree.py:201-213, :28-55): each patch is resampled
 * to its leaf box and pasted into mask (f32 [n0, n1(, n2), C], caller zero-fills).
 * truncate_to_int != 0 reproduces the 2-D reference's `seq.astype(int)` before resizing. */
int ucf_sap_scatter(const float* seq, int ndim, int n0, int n1, int n2, int C, const int32_t* boxes,
                    int n_leaves, int p, int truncate_to_int, float* mask, void* stream);

/* ---- bandwidth-bound helpers ------------------------------------------------------------- */
/* dst_bf16[i] = (bf16) src_f32[i] */
int ucf_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* Up to 8 such casts in ONE launch (the per-step bf16 copies of a Block's four weight matrices): srcs / dsts /
 * counts are HOST arrays of n device pointers / element counts; every pointer 16-byte aligned. */
int ucf_cast_f32_to_bf16_multi(int n, const float* const* srcs, void* const* dsts, const long long* counts,
                               void* stream);
/* dst_f32[i] (+)= (float) src_bf16[i] */
int ucf_cast_bf16_to_f32(const void* src, float* dst, long long n, int accumulate, void* stream);
/* out[n] (+)= sum_m x[m, n]   x bf16 [M, N] pitch ld; out fp32 [N]  (bias gradients) */
int ucf_colsum_bf16(const void* x, float* out, long long M, int N, long long ld, int accumulate,
                    void* stream);
/* Conv{2,3}d(k = s = p) input -> GEMM rows (replaces the im2col inside cuDNN, building_blocks.py:58-60,89)
 * x: [B, C, S0, S1 (, S2)] of x_dtype (fp32 / bf16 / uint8 raw pixels), contiguous, S >= G*p: like the strided convolution,
 *    pixels past the last whole patch are ignored.  Any patch size (16-byte vectors when p % 8 == 0,
 *    8-byte when p % 4 == 0, scalar otherwise).
 * out: bf16 [B*G0*G1(*G2), ld_out], ld_out >= K = C*p^dims and a multiple of 8 (the GEMM's K extent; pad
 *      columns are zero-filled), K ordered (c, p0, p1(, p2)) == conv weight.view(D, -1) */
int ucf_patchify(const void* x, void* out, int B, int C, int G0, int G1, int G2, int p, int dims,
                 int S0, int S1, int S2, long long ld_out, int x_dtype, void* stream);

/* class-token concat + position-embedding add (replaces torch.cat + add in VIT._pos_embed,
 * arch.py:367-393) in one pass: out[b, n] = (n < P ? prefix[n] : tok[b, n - P]) + pos[b * pos_bstride
 * + (n - pos_off)] for n >= pos_off.  tok/out bf16; prefix [P, D] and pos of param_dtype (f32|bf16);
 * pos_bstride = 0 for a table shared by the batch, L*D (elements) for per-sample embeddings;
 * pos_off = 0 when the table has rows for the prefix tokens, P when it has not.  pos may be NULL. */
int ucf_assemble_tokens(const void* tok, const void* prefix, const void* pos, void* out, int B, int L,
                        int P, int D, long long pos_bstride, int pos_off, int param_dtype, void* stream);

/* Broadcast add of an embedding to bf16 tokens in one pass: out[i0, i1, i2, :] = x[i0, i1, i2, :] +
 * e[i0*s0 + i1*s1 + i2*s2 + :] with stride 0 along broadcast axes.  Replaces the variable-embedding add
 * `x + var_embed.unsqueeze(2)` on [B, V, L, D] tokens (simple/arch.py:456-462: s = (0, D, 0)) and DiffusionVIT's
 * time-embedding add `x + temb[:, None, :]` (arch.py:1263-1265: n = (B, 1, N), s = (D, 0, 0)).  x / out bf16
 * [n0, n1, n2, D] contiguous (out may alias x); e of param_dtype (f32 | bf16); D and strides multiples of 8. */
int ucf_add_bcast(const void* x, const void* e, void* out, long long n0, int n1, int n2, int D, long long s0,
                  long long s1, long long s2, int param_dtype, void* stream);

/* ---- MAE token masking (replaces MAE.random_masking / MAE.mask_head, simple/arch.py:663-702) ---------
 * Shuffle / restore permutations and the mask from the per-token noise in ONE launch (replaces two
 * torch.argsort calls, ones + slice-assign + gather): ids_restore[b, j] = rank of noise[b, j] in ascending order
 * (ties: lower index first), ids_shuffle = its inverse permutation, mask[b, j] = 1 if rank >= len_keep (token
 * removed) else 0.  noise / mask f32 [B, L], ids int64 [B, L]; L <= 12288. */
int ucf_mask_plan(const float* noise, int B, int L, int len_keep, long long* ids_shuffle,
                  long long* ids_restore, float* mask, void* stream);
/* out[b, i, :] = (0 <= idx[b, i] < Ls ? src[b, idx[b, i], :] : fill[:]) + pos[b * pos_bstride + i * D + :]
 * - kept-token gather of random_masking (arch.py:674-675): idx = ids_shuffle[:, :len_keep] (pass the row
 *   pitch-free copy), fill = pos = NULL;
 * - mask_head's cat(x, mask_token.repeat) + gather(ids_restore) + decoder_pos_embed add (arch.py:687-698) as one
 *   pass: src = decoder_embed(x) [B, Ls, D], idx = ids_restore [B, Lo], fill = mask_token [D], pos = embedding.
 * src / out bf16; idx int64 [B, Lo]; fill, pos of param_dtype (f32 | bf16), either may be NULL (fill NULL = zeros);
 * pos_bstride = 0 for a table shared by the batch, Lo*D for per-sample embeddings.  D % 8 == 0. */
int ucf_gather_tokens(const void* src, const long long* idx, const void* fill, const void* pos, void* out,
                      int B, int Ls, int Lo, int D, long long pos_bstride, int param_dtype, void* stream);
/* Gradient of ucf_gather_tokens: dsrc[b, idx[b, i], :] = dout[b, i, :] for rows taken from src (idx entries of a
 * sample must be distinct); zero_first != 0 zero-fills dsrc [B, Ls, D] first (rows no index names get no
 * gradient); dfill f32 [D] += sum of the dout rows that were filled.  dsrc or dfill may be NULL. */
int ucf_scatter_tokens(const void* dout, const long long* idx, void* dsrc, float* dfill, int B, int Ls, int Lo,
                       int D, int zero_first, void* stream);

/* ---- either side of the path in a training step ---------------- */
/* Reconstruction loss against the patchified image WITHOUT materialising the patchified target:
 * replaces `target = patchify(data, p, twoD); loss = masked_mse(output, target, mask)` or
 * `nn.MSELoss()(output, target)` (training_scripts/train_masked_fsdp.py:48-62, utils/misc.py:14-33,
 * utils/metrics.py:11-17), and with (G0, G1, G2, p0, p1, p2) = (1, L, 1, 1, 1, P) the adaptive variant
 * `target = rearrange(seq, 'b c s p -> b s (p c)')` (train_masked_fsdp.py:40-43).
 *   pred  [B, L, p0*p1*p2*C]  f32|bf16, L = G0*G1*G2, channel fastest inside a patch
 *   img   [B, C, G0*p0, G1*p1, G2*p2]  f32|bf16 contiguous (2-D images: G2 = p2 = 1)
 *   mask  f32 [B*L] weights (1 = token counts) or NULL for the plain mean over every element
 *   workspace  2 * UCF_PATCH_MSE_MAX_BLOCKS doubles of scratch (per-CTA partial sums of the loss and of the mask)
 *   out   f32 [2]: out[0] = loss, out[1] = 1 / denominator (input of the backward call)
 * Put the image's contiguous axis last in the geometry (2-D images as G = (1, Gy, Gx), p = (1, p, p); the
 * adaptive layout as G = (1, L, 1), p = (1, 1, P)): with p2 % 4 == 0, C <= 4 and 16-byte (f32) / 8-byte (bf16)
 * aligned tensors the kernels move four pixels per load; any other geometry takes a scalar path.
 * The per-CTA partial sums are combined in a fixed order in double: the loss is reproducible. */
#define UCF_PATCH_MSE_MAX_BLOCKS 4096
int ucf_patch_mse_fwd(const void* pred, int pred_dtype, const void* img, int img_dtype, const float* mask,
                      int B, int C, int G0, int G1, int G2, int p0, int p1, int p2, double* workspace,
                      float* out, void* stream);
/* dpred = grad_out * d loss / d pred  (same dtype and shape as pred); fwd_out is `out` of the forward
 * call, grad_out a DEVICE f32 scalar (autograd's incoming gradient).  The image gets no gradient. */
int ucf_patch_mse_bwd(const void* pred, int pred_dtype, const void* img, int img_dtype, const float* mask,
                      const float* fwd_out, const float* grad_out, int B, int C, int G0, int G1, int G2,
                      int p0, int p1, int p2, void* dpred, void* stream);

/* Dice + binary-cross-entropy loss of the SAP driver: replaces DiceBLoss.forward (utils/metrics.py:95-121, called at
 * training_scripts/train_sap_simple.py:44-45).  pred = sigmoid(logits)[:, 1:] (act != 0; act == 0 takes logits as
 * probabilities), true = targets[:, 1:]:
 *   loss = weight * mean(BCE(pred, true)) + (1 - weight) * (1 - (2 sum(pred true) + smooth) / (sum pred + sum true + smooth))
 * with torch's BCE conventions (logarithms clamped at -100; gradient denominator max(p(1-p), 1e-12)).
 * logits / targets [B, C, HW] f32|bf16 contiguous, C >= 2; workspace 4 * UCF_PATCH_MSE_MAX_BLOCKS doubles;
 * out f32 [4] = loss, 2I + smooth, denominator, 1/n (inputs of the backward call).  One read of both tensors. */
int ucf_dice_bce_fwd(const void* logits, int logits_dtype, const void* targets, int targets_dtype, int B, int C,
                     long long HW, float weight, float smooth, int act, double* workspace, float* out, void* stream);
/* dlogits [B, C, HW] (dtype of logits; channel 0 = 0) = grad_out * d loss / d logits; grad_out a DEVICE f32 scalar. */
int ucf_dice_bce_bwd(const void* logits, int logits_dtype, const void* targets, int targets_dtype,
                     const float* fwd_out, const float* grad_out, int B, int C, long long HW, float weight, int act,
                     void* dlogits, void* stream);

/* Dice + cross-entropy loss of the UNETR driver: replaces `monai.losses.DiceCELoss(to_onehot_y=True, softmax=True,
 * squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)(output, label)` (training_scripts/train_unetr_simple.py:38-39,51-52;
 * MONAI 1.4 is a dependency that is not vendored in the reference: restated from its published definition, parity unpinned).
 *   p = softmax(logits, 1), t = one_hot(target); per (b, c) over the S positions: I = sum p t, Q = sum p^2 (squared_pred
 *   != 0) or sum p, T = sum t;  loss = lambda_dice * mean_{b,c}(1 - (2I + smooth_nr) / (Q + T + smooth_dr))
 *                                      + lambda_ce * mean_{b,s}(-log p[b, target[b,s], s])
 * logits f32|bf16, 2 <= C <= 8: [B, C, S] contiguous (channels_last = 0) or [B, S, C] (channels_last != 0: the memory of a
 * torch.channels_last{,_3d} tensor, what the channels-last decoder emits; dlogits in the same layout); target [B, S] class indices as u8, i64 or f32; workspace
 * B * ucf_dice_ce_blocks_per_sample(B, S) * 25 doubles; out f32 [2 + 2 B C]: out[0] = loss, the rest are the per-(b, c)
 * coefficients and the cross-entropy scale the backward call reads.  One read of both tensors; reproducible (fixed order). */
int ucf_dice_ce_blocks_per_sample(int B, long long S);
int ucf_dice_ce_fwd(const void* logits, int logits_dtype, const void* target, int target_dtype, int B, int C, long long S,
                    int squared_pred, float smooth_nr, float smooth_dr, float lambda_dice, float lambda_ce,
                    double* workspace, float* out, int channels_last, void* stream);
/* dlogits [B, C, S] (dtype of logits) = grad_out * d loss / d logits; grad_out a DEVICE f32 scalar. */
int ucf_dice_ce_bwd(const void* logits, int logits_dtype, const void* target, int target_dtype, const float* fwd_out,
                    const float* grad_out, int B, int C, long long S, int squared_pred, void* dlogits, int channels_last,
                    void* stream);

/* Weight gradient of a 3x3x3 convolution (stride 1, padding 1) on channels-last bf16: replaces the wgrad half of autograd's
 * `convolution_backward` for the nn.Conv3d layers inside the decoder blocks above at this decoder's widths.
 *   dW[co][ci][kd][kh][kw] = sum_{n,z,y,x} dY[n,z,y,x,co] * X[n,z+kd-1,y+kh-1,x+kw-1,ci]      (zero outside the volume)
 * x: bf16 [N, D, H, W, Ci], dy: bf16 [N, D, H, W, Co], dw: fp32 [Co, Ci, 3, 3, 3] (overwritten);
 * workspace: fp32 [ucf_conv3d_wgrad_ctas(N, D, H, W) * 27 * Co * Ci].  Served: ucf_conv3d_wgrad_supported() != 0
 * ((Ci, Co) in {(16,16), (32,16), (32,32), (64,32)}, D and H multiples of 4, W of 8); reproducible (fixed-order sums). */
int ucf_conv3d_wgrad_supported(int Ci, int Co, int D, int H, int W);
int ucf_conv3d_wgrad_ctas(int N, int D, int H, int W);
int ucf_conv3d_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Ci, int Co, float* workspace,
                     void* stream);

/* 1x1x1 convolutions of the same decoder (UnetResBlock.conv3 residual projections, UnetOutBlock head; arch.py:808-940,960-993)
 * on channels-last bf16: replace `nn.Conv3d(kernel_size=1)` forward, the data gradient (the same call on the transposed
 * weight) and the weight / bias gradient of autograd's `convolution_backward`.
 *   y[v, co] = sum_ci x[v, ci] * w[co, ci] (+ bias[co]);   dw[co, ci] = sum_v dy[v, co] * x[v, ci];   dbias[co] = sum_v dy[v, co]
 * x: bf16 [V, Ci], y / dy: bf16 [V, Co] (V = all voxels of the batch), w / dw: fp32 [Co, Ci], bias / dbias: fp32 [Co] or NULL;
 * workspace: fp32 [ucf_pointwise_conv_ctas(V) * (Co * Ci + Co)].  Ci, Co each one of 4, 8, 16, 32; fixed-order sums. */
int ucf_pointwise_conv_supported(int Ci, int Co);
int ucf_pointwise_conv_ctas(long long V);
int ucf_pointwise_conv(const void* x, const float* w, const float* bias, void* y, long long V, int Ci, int Co, void* stream);
int ucf_pointwise_conv_wgrad(const void* x, const void* dy, float* dw, float* dbias, long long V, int Ci, int Co,
                             float* workspace, void* stream);

/* ---- SAP front end on the device: edge map of a natural (uint8) image -------------------------------
 * Replace `grey_img = cv.GaussianBlur(img, (k, k), 0)` and `edges = cv.Canny(grey_img, c, c + 50)`
 * (dataloaders/transform.py:33-34; opencv-python is a dependency of the reference that is not vendored in it: restated
 * from its published algorithm in oracle/canny_np.py, pinned bit-for-bit against cv2 4.13 by the tests).  Integer arithmetic,
 * results identical to OpenCV's byte for byte.
 * img / dst: u8 [H, W, C] (C interleaved, 1..4), ksize in {1, 3, 5} (sigma 0: taps [1], [1 2 1]/4, [1 4 6 4 1]/16,
 * BORDER_REFLECT_101, one round-half-up). */
int ucf_gaussian_blur_u8(const void* src, void* dst, int H, int W, int C, int ksize, void* stream);
/* Canny, aperture 3, L1 gradient norm: per pixel the channel with the largest |dx| + |dy|, non-maximum suppression,
 * thresholds floor(low) / floor(high), hysteresis over 8-neighbours.  class_map: u8 [H, W] workspace (2 strong, 0 candidate,
 * 1 none); edges: u8 [H, W] out (0 / 255); flag_dev: one device int; flag_host_pinned: one int of page-locked host memory.
 * The hysteresis repeats whole-image sweeps until none changes the map and reads the flag back after each, so this
 * launcher SYNCHRONISES `stream` (run it on a side stream / host thread); *sweeps_out (may be NULL) = sweeps taken. */
int ucf_canny_u8(const void* img, int H, int W, int C, double low_thresh, double high_thresh, void* class_map, void* edges,
                 int* flag_dev, int* flag_host_pinned, int* sweeps_out, void* stream);

/* ---- UNETR convolutional decoder: InstanceNorm (+ residual) + LeakyReLU on channels-last bf16 -------------------------
 * One call per MONAI block body the reference builds its decoder from (simple/arch.py:808-940 ->
 * monai.networks.blocks.dynunet_block.UnetResBlock / UnetBasicBlock; MONAI 1.4 is not vendored: restated from its published
 * definition, the arithmetic is pinned to torch.nn.InstanceNorm3d + LeakyReLU by tests):
 *     y = lrelu_slope( IN(a)  [ + IN(b)  |  + b ] ),   IN(x)[n, s, c] = (x - mean[n, c]) * rstd[n, c]
 * (nn.InstanceNorm{2,3}d: affine=False, biased variance, eps inside the square root; slope = 1 disables the activation).
 * Tensors are bf16 [N, S, C] with C fastest (torch.channels_last / channels_last_3d of [N, C, ...]); C even, C <= 2048.
 * stats: fp32 [N, 2, C] = (mean, rstd); workspace: fp32 [N * ucf_inorm_chunks(N, S, C) * 3 * C]; coef: fp32 [N, 3, C].
 * b == NULL: no second operand; b != NULL and stats_b == NULL: raw residual; both: normalised residual.
 * Reductions are two-stage and fixed-order (reproducible run to run). */
int ucf_inorm_chunks(int N, long long S, int C);
int ucf_inorm_stats(const void* x, int N, long long S, int C, float eps, float* workspace, float* stats, void* stream);
int ucf_inorm_apply(const void* a, const float* stats_a, const void* b, const float* stats_b, void* y, int N, long long S,
                    int C, float slope, void* stream);
/* da (and db when b took part: pass db != NULL exactly then) from dy; y is the forward output (NULL when slope == 1). */
int ucf_inorm_bwd(const void* dy, const void* y, const void* a, const float* stats_a, const void* b, const float* stats_b,
                  void* da, void* db, int N, long long S, int C, float slope, float* workspace, float* coef, void* stream);

/* One AdamW step (decoupled weight decay, no amsgrad) over n fp32 tensors that share the same
 * hyper-parameters and step count: the arithmetic of torch.optim.AdamW as configured by
 * utils/misc.py:58-84 (`configure_optimizer`: two groups, weight_decay 0 for var/pos embeddings) and
 * stepped at training_scripts/train_class_simple.py:355.
 *   p *= 1 - lr*wd;  m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;
 *   p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps)
 * params / grads / exp_avg / exp_avg_sq / counts are HOST arrays of n device pointers / element
 * counts; `step` is the 1-based count AFTER this update.  16-byte aligned tensors take the float4 path,
 * others a scalar one.  Hyper-parameters are doubles and rounded to fp32 once on the host. */
int ucf_adamw_multi(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                    float* const* exp_avg_sq, const long long* counts, double lr, double beta1, double beta2,
                    double eps, double weight_decay, long long step, int maximize, void* stream);
/* The same update with the learning rate and the step count read from DEVICE memory (fp32 scalars; *step_dev is the
 * 1-based count of this update), so the launch can be captured in a CUDA graph and replayed while a scheduler rewrites
 * *lr_dev and the caller increments *step_dev on the stream between replays.  Bias corrections are evaluated in fp32
 * inside the kernel (the host form above evaluates them in double): updates agree to ~1e-6 relative. */
int ucf_adamw_multi_dev(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, const long long* counts, const float* lr_dev, double beta1,
                        double beta2, double eps, double weight_decay, const float* step_dev, int maximize,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UCF_VIT_B200_H_ */
