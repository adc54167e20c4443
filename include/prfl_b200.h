/* prfl_b200 — C ABI of the B200 (sm_100a) kernels behind the Wan-DiT / PAVRM hot path of
 * HY-Video-PRFL.  This is the drop-in boundary: plain pointers, sizes and a cudaStream_t; no
 * torch types.  The reference has no FFI of its own (it is pure PyTorch; its hot ops are
 * third-party kernels called through torch) so each entry point cites the reference call site
 * whose third-party kernel(s) it replaces.  Paths are relative to the reference checkout.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; the caller owns all memory (the library
 *     allocates nothing persistent), and all work is enqueued on `stream` (a cudaStream_t).
 *   - return value: PRFL_OK or a negative PRFL_E_* code; prfl_last_error_string() describes the
 *     last failure on the calling thread.  Nothing throws across the ABI.
 *   - bf16 tensors are `void*` to keep this header free of CUDA headers; "f32" = float.
 *   - no CPU fallback: on a device whose compute capability is not 10.x every compute entry
 *     point returns PRFL_E_ARCH.
 */
#ifndef PRFL_B200_H_
#define PRFL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRFL_OK 0
#define PRFL_E_SHAPE (-1) /* unsupported / inconsistent sizes                    */
#define PRFL_E_ALIGN (-2) /* pointer or leading dimension not 16-byte aligned    */
#define PRFL_E_ARCH (-3)  /* device is not sm_100 / no CUDA device               */
#define PRFL_E_CUDA (-4)  /* CUDA runtime / driver error (see last error string) */

typedef void* prfl_stream_t; /* cudaStream_t */

/* ---- library -------------------------------------------------------------------------------- */
int prfl_abi_version(void);
const char* prfl_last_error_string(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's
 * `gpu_launches` claim is read from here, not estimated) */
int64_t prfl_launch_count(void);
void prfl_launch_count_reset(void);

/* ---- LayerNorm (+AdaLN modulate), fp32 residual in -> bf16 out -------------------------------
 * Replaces ATen layer_norm + pointwise chain at model.py:345,353 (norm1/norm2 then
 * `.float() * (1 + e[1]) + e[0]`), model.py:352 (norm3, affine) and WanLayerNorm model.py:125-135.
 *   y = (x - mean) * rstd                  (fp32, eps inside the sqrt)
 *   if round_bf16   : y = bf16(y)           (block 0: x is bf16 there and norm1 returns type_as(x))
 *   if gamma        : y = y * gamma + beta  (norm3)
 *   if scale        : y = y * (1 + scale) + shift
 *   out = bf16(y)
 * x: [rows, C] f32 row-major; shift/scale/gamma/beta: [C] f32 or NULL; mean/rstd: [rows] f32 or
 * NULL (saved for the backward).  C % 256 == 0, C <= 8192. */
int prfl_ln_mod_fwd(const float* x, const float* shift, const float* scale, const float* gamma, const float* beta,
                    void* out_bf16, float* mean, float* rstd, int64_t rows, int C, float eps, int round_bf16,
                    prfl_stream_t stream);
/* ---- RMSNorm over the full channel dim (+ 3-D RoPE), bf16 in -> bf16 out ----------------------
 * Replaces WanRMSNorm (model.py:106-122) applied to q/k (model.py:175-176, 216-217, 256-260) and
 * rope_apply (model.py:60-103) for self-attention.
 *   t = bf16(x * rsqrt(mean(x^2) + eps)) * w            (w f32; the rounding is model.py:119)
 *   rows < n_rot: pairs (2j, 2j+1) of every 128-wide head rotate by angle[pos0 + row][j]
 *   out = bf16(t)
 * x/out: [rows, C] bf16 with row strides ldx/ldo (elements; lets q,k live inside a fused
 * [rows, 3C] QKV buffer; in-place allowed).  cos/sin: [n_pos, 64] f32 or NULL (no RoPE: cross-attn).
 * rstd: [rows] f32 or NULL.  pos0 = first row's position (Ulysses rank offset, model.py:89-96). */
int prfl_rmsnorm_rope_fwd(const void* x_bf16, int64_t ldx, const float* w, const float* cos_tab, const float* sin_tab,
                          void* out_bf16, int64_t ldo, float* rstd, int64_t rows, int C, int64_t n_rot, int64_t pos0,
                          float eps, prfl_stream_t stream);
/* Head (model.py:379-389) runs its LayerNorm + modulate + Linear(dim -> 64) in fp32 in the reference.  Here the
 * modulated row h is emitted as a bf16 pair h = hi + lo (lo = bf16(h - hi), ~16 mantissa bits) and the Linear is three
 * tensor-core GEMMs hi.Whi + hi.Wlo + lo.Whi accumulated in fp32 (prfl_gemm_bf16, PRFL_EPI_F32 with beta): fp32-grade
 * result (relative error ~1e-5) from the bf16 tensor pipe.  out_hi/out_lo: [rows, C] bf16. */
int prfl_ln_mod_split_fwd(const float* x, const float* shift, const float* scale, void* out_hi_bf16, void* out_lo_bf16,
                          int64_t rows, int C, float eps, prfl_stream_t stream);

/* ---- backward of the two norm kernels + the token-dimension reductions ---------------------------
 * Replace what autograd derives for model.py:345,352,353 (LayerNorm + modulate / affine), model.py:106-122 + 60-103
 * (RMSNorm + RoPE) and the bias / modulation / gate / norm-weight gradient reductions over tokens.
 * Column reductions write PARTIAL sums [nparts, N] f32 (nparts = prfl_colsum_parts(rows), 256 rows each); the caller
 * adds the nparts rows (N floats each — negligible). */
int prfl_colsum_parts(int64_t rows);
/* dx_accum[rows, C] (f32) += dL/dx of prfl_ln_mod_fwd given dy (bf16).  If dshift_part/dscale_part are given they
 * receive partial sums of dy and dy * xhat: (dshift, dscale) of the modulated form, (dbeta, dgamma) of the affine form. */
int prfl_ln_mod_bwd(const float* x, const void* dy_bf16, const float* scale, const float* gamma, const float* mean,
                    const float* rstd, float* dx_accum, float* dshift_part, float* dscale_part, int64_t rows, int C,
                    prfl_stream_t stream);
/* dx (bf16, may alias dy) = dL/dx of prfl_rmsnorm_rope_fwd; gw_bf16 [rows, C] (or NULL) = dt * bf16(x * rstd), whose
 * column sum (prfl_colsum_bf16) is dL/dw. */
int prfl_rmsnorm_rope_bwd(const void* x_bf16, int64_t ldx, const float* w, const float* cos_tab, const float* sin_tab,
                          const void* dy_bf16, int64_t lddy, const float* rstd, void* dx_bf16, int64_t lddx, void* gw_bf16,
                          int64_t ldgw, int64_t rows, int C, int64_t n_rot, int64_t pos0, prfl_stream_t stream);
/* part[chunk, n] = sum over the chunk's rows of a[row, n] (bias and norm-weight gradients). */
int prfl_colsum_bf16(const void* a_bf16, int64_t lda, float* part, int64_t rows, int N, prfl_stream_t stream);
/* Backward of x_out = x_in + gate * y (model.py:348,355): dy_bf16 = bf16(dx * gate) (gate NULL => 1, i.e. a cast);
 * dgate_part (or NULL) = partial sums of dx * y. */
int prfl_gate_bwd(const float* dx, const void* y_bf16, const float* gate, void* dy_bf16, float* dgate_part, int64_t rows, int N,
                  prfl_stream_t stream);

/* ---- bf16 tcgen05 GEMM with fused epilogues ---------------------------------------------------
 * Replaces cuBLAS(Lt) behind nn.Linear on the path: q,k,v,o (model.py:156-159,175-177,200),
 * cross-attn q,k,v,o (model.py:216-225, 256-270), FFN (model.py:313-315), patch embedding as a
 * GEMM (model.py:497-498,578), text embedding (model.py:499-501) and — for the backward — the
 * dgrad / wgrad GEMMs autograd derives from them.
 *   acc[m,n] = sum_k A(m,k) * B(n,k)          bf16 operands, fp32 accumulate in TMEM
 *   a_trans = 0: A stored [M, K] (K contiguous), lda;  a_trans = 1: A stored [K, M] (M contiguous)
 *   b_trans = 0: B stored [N, K] (K contiguous), ldb;  b_trans = 1: B stored [K, N] (N contiguous)
 * Epilogue (`epi`):
 *   PRFL_EPI_BF16       out_bf16[m,n]  = bf16(acc + bias[n])
 *   PRFL_EPI_BF16_GELU  out_bf16[m,n]  = bf16(gelu_tanh(bf16(acc + bias[n])))        (model.py:314)
 *   PRFL_EPI_F32        out_f32[m,n]   = acc + bias[n]        (beta=1: out_f32 += ...; wgrad accumulation)
 *   PRFL_EPI_RESIDUAL   out_f32[m,n]   = resid[m,n] + gate[n] * bf16(acc + bias[n])   gate NULL => 1; resid NULL => out
 *                       (fuses the gated residual adds model.py:348,352,355 into o / ffn.2; resid [M, N] f32 with the
 *                       leading dimension of out: a separate input stream lets the checkpointed backward keep the
 *                       block's intermediate residual states without cloning them first)
 *   PRFL_EPI_BF16_DGELU out_bf16[m,n]  = bf16(acc * gelu_tanh'(aux_bf16[m,n]))   (FFN backward; aux is an INPUT)
 * With PRFL_EPI_BF16_GELU or PRFL_EPI_RESIDUAL a non-NULL aux_bf16 [M, N] (ldaux) is an extra OUTPUT that receives
 * bf16(acc + bias[n]) — the pre-activation / the un-gated branch output the backward needs.
 * bias: [N] f32 or NULL.  M,N,K > 0; N % 8 == 0; K % 8 == 0 unless both operands are transposed (wgrad over a ragged token count);
 * lda/ldb/ldc % 8 == 0. */
#define PRFL_EPI_BF16 0
#define PRFL_EPI_BF16_GELU 1
#define PRFL_EPI_F32 2
#define PRFL_EPI_RESIDUAL 3
#define PRFL_EPI_BF16_DGELU 4
int prfl_gemm_bf16(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, void* out,
                   int64_t ldc, const float* bias, const float* gate, const float* resid, void* aux_bf16, int64_t ldaux,
                   int M, int N, int K, int epi, int beta, prfl_stream_t stream);

/* ---- flash attention, bf16, head_dim 128, non-causal ------------------------------------------
 * Replaces flash_attn_varlen_func behind flash_attention() (attention.py:24-130) as called from
 * model.py:188-193 (self-attention, k_lens = seq_lens), model.py:221 and model.py:262-264
 * (cross-attention).  One sequence per call (the path runs B = 1 per sample).
 *   O[i,h,:] = softmax_j(scale * Q[i,h,:] . K[j,h,:]) V[j,h,:],  j < Lk
 * q/k/v/o: bf16, element (token t, head h, dim d) at base + t*ld_tok + h*ld_head + d (strides in
 * elements, multiples of 8: lets them alias a fused QKV buffer or an Ulysses staging layout).
 * lse: [H, Lq] f32 (natural-log sum-exp of the scaled scores) or NULL; needed by the backward.
 * ws: NULL, or a 16-byte aligned device workspace of prfl_attn_fwd_ws_bytes(Lq, Lk, H) bytes.  With a workspace the
 * units past the last full wave of CTAs are split along the key axis and merged by a second kernel (wave quantisation:
 * 640 equal units on 148 SMs would otherwise run 5 waves, the last 32 % full); results differ from the unsplit launch by
 * fp32 summation order only. */
int64_t prfl_attn_fwd_ws_bytes(int Lq, int Lk, int H);
/* The same policy as a pure host function (no device): out4 = {units, units of the plain launch, tail units, pieces per
 * tail unit} for a GPU with n_sms SMs. */
void prfl_attn_fwd_split_plan(int Lq, int Lk, int H, int n_sms, int* out4);
int prfl_attn_fwd(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                  int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, void* o, int64_t o_ld_tok,
                  int64_t o_ld_head, float* lse, int Lq, int Lk, int H, float scale, void* ws, prfl_stream_t stream);
/* Same kernel with the Ulysses output exchange (model.py:195-196, communication.py:91-123) fused into its epilogue:
 * this rank computed H = H_total/P heads over all Lq tokens; row i is stored directly into rank (i / L_loc)'s
 * [L_loc, H_total, 128] buffer o_peers[i / L_loc] (peer-mapped NVLink pointer; HOST array of n_peer device pointers) at
 * token i % L_loc, head head_off + h.  No NCCL call, no staging copy.  The caller barriers across ranks before reading. */
int prfl_attn_fwd_p2p(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                      int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, void* const* o_peers,
                      int n_peer, int L_loc, int head_off, int64_t o_ld_tok, int64_t o_ld_head, float* lse, int Lq, int Lk,
                      int H, float scale, void* ws, prfl_stream_t stream);
/* Backward (replaces flash_attn's bwd kernels reached through autograd from the same call sites): dq, dk, dv (bf16,
 * same addressing as q/k/v) from q, k, v, o, dout and the forward's lse.  ws: f32 workspace of
 * prfl_attn_bwd_ws_floats(Lq, H) elements (16-byte aligned), filled here with -lse*log2(e) and -rowsum(dout * o) padded
 * to whole 64-query blocks.  Two tcgen05 kernels (dK/dV with keys resident, dQ with queries resident, the resident
 * operands held in tensor memory); no atomics, deterministic. */
int64_t prfl_attn_bwd_ws_floats(int Lq, int H);
int prfl_attn_bwd(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                  int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, const void* o,
                  int64_t o_ld_tok, int64_t o_ld_head, const void* dout, int64_t do_ld_tok, int64_t do_ld_head,
                  const float* lse, float* ws, void* dq, int64_t dq_ld_tok, int64_t dq_ld_head, void* dk,
                  int64_t dk_ld_tok, int64_t dk_ld_head, void* dv, int64_t dv_ld_tok, int64_t dv_ld_head, int Lq, int Lk,
                  int H, float scale, prfl_stream_t stream);

/* ---- patchify / unpatchify -------------------------------------------------------------------
 * Patch embedding Conv3d(kernel = stride = (1,2,2)) (model.py:497-498,578-581) = gather + GEMM:
 * this gathers latent [Cin, F, H, W] f32 into patches [F*(H/2)*(W/2), Cin*4] bf16, column index
 * (c, ph, pw) matching `patch_embedding.weight.reshape(dim, -1)`.  Two sources (x then y) are
 * concatenated on the channel axis as model.py:574-575 does; y may be NULL (Cy = 0). */
int prfl_patchify(const float* x, int Cx, const float* y, int Cy, void* patches_bf16, int F, int H, int W,
                  prfl_stream_t stream);
/* Scatter-add of the above's transpose: dpatches [L, (Cx+Cy)*4] f32 -> dlatent [Cx, F, H, W] f32 (x part only). */
int prfl_patchify_bwd(const float* dpatches, int Cx, int Ctot, float* dx, int F, int H, int W, prfl_stream_t stream);
/* WanModel.unpatchify (model.py:683-705): tokens [F*h*w, 4*c] f32 -> video [c, F, 2h, 2w] f32
 * ('fhwpqrc->cfphqwr' with patch (1,2,2)); `inverse` != 0 runs the transpose map (backward). */
int prfl_unpatchify(const float* tokens, float* video, int c, int F, int h, int w, int inverse, prfl_stream_t stream);

/* ---- PAVRM single-query attention pooling ------------------------------------------------------
 * QueryAttention with one learnable query (network.py:44-110): because the query is a single
 * vector, q.(x Wk^T + bk) = x.(Wk^T q) + const and sum_l p_l (x_l Wv^T + bv) = (sum_l p_l x_l) Wv^T + bv,
 * so the [L, C] x [C, 2C] in-proj GEMM collapses to two streaming passes over the features:
 *   scores[l, h] = x[l, :] . wk_eff[h, :]                                  (pass 1)
 *   pooled[h, :] = sum_l softmax_l(scores[:, h])[l] * x[l, :]               (pass 2)
 * x: [L, C] f32; wk_eff: [NH, C] f32 (already scaled by 1/sqrt(hd)); scores: [L, NH] f32 workspace;
 * stats: [2*NH] f32 workspace (row max, sum); pooled: [NH, C] f32 (zeroed by the call). NH <= 8. */
int prfl_sq_pool_fwd(const float* x, const float* wk_eff, float* scores, float* stats, float* pooled, int64_t L,
                     int C, int NH, prfl_stream_t stream);
/* Backward: dx[l,:] (+)= sum_h p[l,h] * dpooled[h,:] + ds[l,h] * wk_eff[h,:],
 * ds[l,h] = p[l,h] * (x[l,:].dpooled[h,:] - pooled[h,:].dpooled[h,:]).  ds: [L, NH] f32 out (or NULL); the
 * caller gets dwk_eff = ds^T x from it.  accumulate != 0 adds into dx. */
int prfl_sq_pool_bwd(const float* x, const float* wk_eff, const float* scores, const float* stats,
                     const float* pooled, const float* dpooled, float* dx, float* ds, int64_t L, int C, int NH,
                     int accumulate, prfl_stream_t stream);

/* ---- ring / context-parallel attention merge ----------------------------------------------------
 * Replaces the out/lse update of xfuser's ring attention behind xFuserLongContextAttention
 * (diffusers_lite/wan/distributed/xdit_context_parallel.py:214-219): folds the attention over one more key block
 * (o_new bf16 [L, H, 128] strided, lse_new [H, L] natural log, both from prfl_attn_fwd) into the running fp32 result
 * (o_acc [L, H, 128] contiguous, lse_acc [H, L]).  first != 0 initialises the running result.  out_bf16 (may be NULL):
 * strided [L, H, 128] destination that also receives bf16(o_acc) — pass it on the last block. */
int prfl_attn_merge(float* o_acc, float* lse_acc, const void* o_new_bf16, int64_t n_ld_tok, int64_t n_ld_head,
                    const float* lse_new, int first, void* out_bf16, int64_t o_ld_tok, int64_t o_ld_head, int L, int H,
                    prfl_stream_t stream);

/* ---- misc elementwise --------------------------------------------------------------------------*/
/* dst_bf16[i] = bf16(src_f32[i]) — fp32 master weights -> bf16 operands (what autocast does per call,
 * done once here). n % 8 == 0 not required. */
int prfl_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, prfl_stream_t stream);
/* Ulysses staging (communication.py:60-123 restated): the reference does reshape/transpose/contiguous before AND
 * after each all_to_all_single; with the [P][L_loc][H/P][128] wire layout only one copy per exchange is left:
 *   mode 0 (pack, before the q/k/v exchange)        packed[p][t][hl][:] = strided[t][p*(H/P) + hl][:]
 *   mode 1 (unpack, after the attention-out exchange) strided[t][p*(H/P) + hl][:] = packed[p][t][hl][:]
 * The received q/k/v buffer already is [L, H/P, 128] in global token order, and the attention output [L, H/P, 128]
 * already is the send buffer.  strided: element (t, h, d) at t*ld_tok + h*ld_head + d, t < L_loc, h < H. */
int prfl_a2a_pack(void* strided, int64_t ld_tok, int64_t ld_head, void* packed, int L_loc, int H, int P, int mode,
                  prfl_stream_t stream);

/* The q/k/v exchange (model.py:183-186) as direct peer stores: peer_recv[p][(rank*L_loc + t), hl, :] = strided[t][p*(H/P)+hl][:]
 * where peer_recv is a HOST array of P peer-mapped device pointers to each rank's [P*L_loc, H/P, 128] receive buffer. */
int prfl_a2a_scatter_p2p(const void* strided, int64_t ld_tok, int64_t ld_head, void* const* peer_recv, int L_loc, int H, int P,
                         int rank, prfl_stream_t stream);
/* The reverse exchange (the attention-output all-to-all of model.py:196 and, in backward, the dq / dk / dv exchanges that
 * autograd derives from model.py:183-186) as direct peer stores: this rank holds all P*L_loc tokens of its Hl = H/P heads
 * (src, element (t, hl, d) at t*src_ld_tok + hl*src_ld_head + d); token chunk p goes to rank p at this rank's head offset:
 *   peer_dst[p][t*dst_ld_tok + (rank*Hl + hl)*dst_ld_head + d] = src[(p*L_loc + t), hl, d]
 * peer_dst: HOST array of P peer-mapped device pointers (e.g. each rank's [L_loc, H, 128] buffer, or a column block of its
 * fused [L_loc, 3*H*128] dQKV buffer: dst_ld_tok = 3*H*128). */
int prfl_a2a_gather_p2p(const void* src, int64_t src_ld_tok, int64_t src_ld_head, void* const* peer_dst, int64_t dst_ld_tok,
                        int64_t dst_ld_head, int L_loc, int Hl, int P, int rank, prfl_stream_t stream);

/* ---- PRFL chain glue: scheduler step ------------------------------------------------------------
 * One FlowUniPCMultistepScheduler.step (diffusers_lite/wan/utils/fm_solvers_unipc.py:655-739: convert_model_output
 * :318-321, UniC corrector :486-626, UniP predictor :350-484; called from train_prfl.py:690,734 and
 * text2video.py:298-303) as one kernel over the latent (all fp32, n elements):
 *   v         = model_output, or model_output_uncond + guide_scale * (model_output - model_output_uncond) when
 *               model_output_uncond != NULL (classifier-free guidance, text2video.py:295-296)
 *   x0        = sample - sigma * v
 *   corrected = c[0] last_sample + c[1] x0 + c[2] hist0 + c[3] hist1 + c[4] hist2      (only if corr_coef != NULL)
 *   prev      = p[0] (corrected | sample) + p[1] x0 + p[2] hist0 + p[3] hist1 + p[4] hist2
 * hist0..2 = the scheduler's previous x0 predictions, newest first (NULL when absent; their coefficients must be 0).
 * corr_coef / pred_coef are HOST arrays of 5 floats folded from the step's sigmas by the caller (scheduler.py). */
int prfl_unipc_step(const float* sample, const float* model_output, const float* model_output_uncond, float guide_scale,
                    const float* last_sample, const float* hist0, const float* hist1, const float* hist2, float sigma,
                    const float* corr_coef, const float* pred_coef, float* x0_out, float* corrected_out, float* prev_out,
                    int64_t n, prfl_stream_t stream);
/* Backward of the step (it is linear): ya = a * g, yb = b * g (yb may be NULL). */
int prfl_scale2_f32(const float* g, float a, float* ya, float b, float* yb, int64_t n, prfl_stream_t stream);

/* ---- sharded optimizer ---------------------------------------------------------------------------
 * AdamW (decoupled weight decay, torch.optim.AdamW semantics; train_prfl.py:482-491, 825-830) on one rank's fp32 shard of
 * an FSDP unit: grad (already reduce-scattered), master weights and both moments, n elements each, updated in place by
 * one kernel.  clip_coef_dev: DEVICE pointer to the clip_grad_norm_ coefficient (NULL = 1); step = 1-based update count.
 * master_bf16_out (may be NULL): n bf16 elements receiving bf16(master) in the same pass — this rank's slice of the resident
 * bf16 operand buffer, all-gathered in bf16 afterwards (what torch.autocast re-derives from FSDP's fp32 all-gather,
 * fsdp_utils.py:86-109, on every use). */
int prfl_adamw_step(const float* grad, float* master, float* exp_avg, float* exp_avg_sq, const float* clip_coef_dev,
                    void* master_bf16_out, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    prfl_stream_t stream);
/* acc[0] (DEVICE double) += sum_i x[i]^2 — the local part of FSDP.clip_grad_norm_ (train_prfl.py:825). */
int prfl_sumsq_f32(const float* x, int64_t n, double* acc, prfl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PRFL_B200_H_ */
