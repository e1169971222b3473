/* cinema_b200 -- C-ABI of the B200 (sm_100a) kernels behind the CineMA MAE-ViT hot path.
 *
 * The reference (mathpluscode/CineMA) has no FFI / plugin layer: its hot path is Python
 * nn.Modules calling ATen / cuBLAS / cuDNN / SDPA.  The drop-in boundary is therefore the
 * nn.Module API (cinema_b200/*.py mirrors it), and this header is the native layer that
 * replaces the library calls *below* that API.  Each entry point names the reference call
 * site(s) (file:line under /root/reference) whose GPU work it takes over.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (torch's caching allocator); nothing is allocated, freed or retained by the library
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync
 *   - return 0 on success, non-zero on failure (cb_last_error() gives the message);
 *     nothing throws across the ABI
 *   - bf16 buffers are `void*` of 2-byte elements; leading dimensions / strides are in
 *     ELEMENTS unless the name says bytes
 */
#ifndef CINEMA_B200_H_
#define CINEMA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CINEMA_B200_ABI_VERSION 1

enum { CB_DT_BF16 = 0, CB_DT_F32 = 1, CB_DT_U8 = 2, CB_DT_I16 = 3, CB_DT_U16 = 4 };
enum { CB_EPI_NONE = 0, CB_EPI_GELU = 1, CB_EPI_GELU_BWD = 2 };

/* ---- library ------------------------------------------------------------------------ */
int cb_version(void);
const char* cb_last_error(void);
int cb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- GEMM (tcgen05 / TMEM / TMA) ----------------------------------------------------- *
 * C[M,N] = alpha * sum_k A(m,k) B(n,k), fp32 accumulation in tensor memory.
 *   a_mn_major = 0: A stored [M,K] (row pitch lda)   1: A stored [K,M]
 *   b_mn_major = 0: B stored [N,K] (row pitch ldb)   1: B stored [K,N]
 * epilogue (applied in this order): + bias[n]; CB_EPI_GELU (out <- bf16 pre-activation
 * [may be NULL], out2 <- bf16 GELU); CB_EPI_GELU_BWD (acc *= GELU'(aux[m,n])); + residual[m,n]
 * (fp32); store to out as bf16 / fp32, or red.add into fp32 out when accumulate=1; out2 (if
 * given, non-GELU) receives a bf16 copy of the stored value.
 * split_k: 0 = auto (only ever >1 when accumulate=1).  block_n: 0 = auto, or 64/128/256.
 * colsum (fp32 [N], may be NULL; bf16-output epilogues without GELU only): colsum[n] += sum_m of the bf16 values stored to
 * out -- the bias gradient of the Linear layer that consumes `out` as its output gradient (no separate column-sum pass).
 * row_scale (fp32 [ceil(M / rows_per_group)], may be NULL; plain epilogue, non-accumulating): the value alpha*acc+bias of
 * row m is multiplied by row_scale[m / rows_per_group] before the residual is added -- timm DropPath (stochastic depth,
 * per-sample keep / keep_prob factors) of cinema/vit.py:562,577,606,608 fused into the branch's last Linear.
 * Replaces nn.Linear forward/backward: cinema/vit.py:472-477,498-499,520 (q, kv, proj),
 * timm Mlp fc1/GELU/fc2 (cinema/vit.py:570-575), cinema/vit.py:294-298,342 (PatchEmbed.proj),
 * cinema/convvit.py:121,205 (linear), cinema/convvit.py:252,284 (k==s down convs as GEMM),
 * cinema/mae/mae.py:395,567 (dec_linear), cinema/mae/mae.py:435-440,594 (pred heads). */
int cb_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, int M,
                 int N, int K, void* out, long long ldo, int out_dtype, int accumulate, void* out2, long long ldo2,
                 const float* bias, const float* residual, long long ldr, const void* aux, long long ldaux,
                 int epilogue, float alpha, int split_k, int block_n, float* colsum, const float* row_scale,
                 int rows_per_group, void* stream);

/* EXPERIMENTAL -- exported for the next round's segmentation decoder, not used by any product path yet and not validated
 * on a GPU in round 1 (arithmetic pinned on the CPU: tools/conv_rowspace_prototype.py, tests/test_conv_rowspace_design.py).
 * 3^n "same" convolution over a zero-haloed channel-last ROW SPACE as one K-concatenated tcgen05 GEMM:
 *   out[r, :] = sum_t X[r + row_off[t], :] . W[:, t*c_in : (t+1)*c_in]^T (+ bias) (+ residual),  rows outside [0, rows) read 0.
 * X [rows, c_in] bf16 (c_in % 64 == 0), W [c_out, n_taps*c_in] bf16 tap-major, row_off: HOST array of n_taps <= 27 ints.
 * Replaces nn.Conv3d / nn.Conv2d with kernel 3, padding "same" of ConvResBlock (cinema/conv.py:314-315) in the UNETR decoder
 * (cinema/segmentation/convunetr.py:66-81,340-388); the same entry point computes the input gradient (negated offsets,
 * transposed weights). */
int cb_conv_gemm_bf16(const void* X, long long ldx, long long rows, int c_in, const void* W, long long ldw, int c_out,
                      int n_taps, const int* row_off, void* out, long long ldo, int out_dtype, const float* bias,
                      const float* residual, long long ldr, int block_n, void* stream);

/* column sums of a bf16 [M,N] matrix accumulated (atomically) into fp32 out[N]: bias gradients.
 * Replaces the reduce kernels autograd runs for nn.Linear bias (same call sites as above). */
int cb_colsum_bf16(const void* X, long long ldx, int M, int N, float* out, void* stream);

/* ---- fused attention (tcgen05 flash attention) ---------------------------------------- *
 * O = softmax(Q K^T * scale) V per (batch, head); head_dim in {32, 64}.  Q/K/V/O are bf16 views
 * addressed as ptr + b*stride_b + token*stride_n + head*stride_h (+ dim), strides in elements,
 * so the fused [B,N,3,H,d] projection output is consumed in place (no permute copies).
 * lse[b,h,q] = log-sum-exp of the scaled scores (fp32), kept for the backward.
 * Replaces F.scaled_dot_product_attention at cinema/vit.py:505-511 and the permutes at :498-500,519. */
int cb_attention_fwd(const void* q, long long q_sb, long long q_sn, long long q_sh, const void* k, long long k_sb,
                     long long k_sn, long long k_sh, const void* v, long long v_sb, long long v_sn, long long v_sh,
                     void* o, long long o_sb, long long o_sn, long long o_sh, float* lse, int B, int H, int Nq, int Nk,
                     int head_dim, float scale, void* stream);
/* dQ, dK, dV (bf16, same addressing scheme).  Caller-provided workspaces: delta (fp32 scratch, 2 * B*H*NqP floats with
 * NqP = Nq rounded up to a multiple of 128: the per-query vectors -lse*log2(e) and scale*rowsum(O o dO), padded so
 * that the kernel fetches them with 512-byte bulk copies) and dq_acc (fp32 scratch, B*H*Nq*head_dim). */
int cb_attention_bwd(const void* q, long long q_sb, long long q_sn, long long q_sh, const void* k, long long k_sb,
                     long long k_sn, long long k_sh, const void* v, long long v_sb, long long v_sn, long long v_sh,
                     const void* o, long long o_sb, long long o_sn, long long o_sh, const void* d_o, long long do_sb,
                     long long do_sn, long long do_sh, const float* lse, void* dq, long long dq_sb, long long dq_sn,
                     long long dq_sh, void* dk, long long dk_sb, long long dk_sn, long long dk_sh, void* dv,
                     long long dv_sb, long long dv_sn, long long dv_sh, float* delta, float* dq_acc, int B, int H,
                     int Nq, int Nk, int head_dim, float scale, float* dq_colsum, float* dk_colsum, float* dv_colsum,
                     void* stream);
/* dq_colsum / dk_colsum / dv_colsum (fp32 [H * head_dim] each, any may be NULL): += column sums over (batch, token) of the
 * bf16 values stored to dQ / dK / dV -- the bias gradients of the q / k / v projections (qkv_bias=True,
 * cinema/vit.py:472-473), fused into the dQ conversion pass and the dK / dV epilogue instead of three more passes. */

/* ---- LayerNorm (row-wise over the last dim, fp32 statistics) -------------------------- *
 * y = (x - mean) * rstd * gamma + beta.  x fp32 [M,D] (row pitch ldx).  y16 (bf16) and/or y32
 * (fp32) may be NULL.  mean / rstd (fp32 [M]) may be NULL for inference.  act=1 applies exact-erf GELU to the
 * normalised output (ConvNormActBlock, cinema/conv.py:257-273).
 * Replaces nn.LayerNorm at cinema/vit.py:549,564,650,738, cinema/convvit.py:254,290 and the
 * permute+LayerNorm+permute of ConvLayerNorm (cinema/conv.py:169-187) on channel-last rows. */
int cb_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, int M, int D, float eps,
                     void* y16, long long ldy16, float* y32, long long ldy32, float* mean, float* rstd, int act,
                     void* stream);
/* beta_act != NULL: the forward was run with act=1; dy is first multiplied by GELU'(LN output) (needs beta).
 * dx = LN'(dy) (+ dres if given) -> dx32 (fp32) and optional bf16 copy dx16; dgamma / dbeta are
 * accumulated with fp32 atomics (may be NULL).  dy is bf16 (dy_dtype=CB_DT_BF16) or fp32.
 * dxsum (may be NULL): dxsum[c] += sum over rows of bf16(dx[row, c]) -- the bias gradient of the Linear layer whose
 * output gradient is dx16 (fuses the column-sum pass; needs D in {16,32,64,128} or 256 <= D <= 1024 with 16-byte aligned rows). */
int cb_layernorm_bwd(const void* dy, long long lddy, int dy_dtype, const float* x, long long ldx, const float* mean,
                     const float* rstd, const float* gamma, const float* dres, long long lddres, int M, int D,
                     float* dx32, long long lddx32, void* dx16, long long lddx16, float* dgamma, float* dbeta,
                     float* dxsum, const float* beta_act, void* stream);

/* ---- data movement (bit-exact) --------------------------------------------------------- */
/* fp32 -> bf16 (round-to-nearest-even) over a flat buffer: the per-step bf16 shadow of the weights. */
int cb_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);

/* (B,n) boolean mask (1 = removed) -> ascending index lists keep_idx (B,n_keep), drop_idx (B,n-n_keep)
 * and, optionally, slot[b,t] = position of token t inside its list.  Each row must hold exactly
 * n_keep zeros.  Replaces the nonzero()+index host-sync pairs behind x[~mask] / x[mask] at
 * cinema/mae/mae.py:98-99,140,550 and cinema/convvit.py:288. */
int cb_mask_to_index(const unsigned char* mask, int B, int n, int n_keep, int* keep_idx, int* drop_idx, int* slot,
                     void* stream);

/* out[b, out_off + i, :] = src[b * src_bstride + idx[b,i], :]  for i < k; rows are row_bytes wide
 * (multiple of 16).  src_bstride (rows) = 0 broadcasts one table over the batch (pos-embed);
 * out_bstride is the row count of one batch item of `out`.  Pure byte copy. */
int cb_gather_rows(const void* src, long long src_bstride, const int* idx, int B, int k, void* out,
                   long long out_bstride, long long out_off, long long row_bytes, void* stream);
/* inverse: dst[b * dst_bstride + idx[b,i], :] = src[b, src_off + i, :] (rows not listed are untouched). */
int cb_scatter_rows(const void* src, long long src_bstride, long long src_off, const int* idx, int B, int k, void* dst,
                    long long dst_bstride, long long row_bytes, void* stream);
/* out[b, out_off+i, :] = (a ? a[b, a_off+i, :] : 0) + (row ? row[:] : 0) + (table ? table[idx[b,i], :] : 0)
 * computed in fp32 (width D) and stored to out (fp32) and / or out16 (bf16), same row addressing.
 * The decoder embedding: visible tokens + pos[keep], mask_token + pos[drop]
 * (cinema/mae/mae.py:68-104,179-204) and the encoder's pos-embed add (cinema/convvit.py:205). */
int cb_embed_rows_f32(const float* a, long long a_bstride, long long a_off, const float* row, const float* table,
                      const int* idx, int B, int k, int D, float* out, void* out16, long long out_bstride,
                      long long out_off, void* stream);
/* out[d] += sum_{b<B, i<k} X[(b*bstride_rows + off + i) * D + d]  (fp32): gradients of cls / mask tokens, i.e.
 * the backward of the broadcast in cinema/vit.py:672 and cinema/mae/mae.py:98-99. */
int cb_colsum_seg_f32(const float* X, long long bstride_rows, long long off, int B, int k, int D, float* out,
                      void* stream);
/* Input pipeline, device side: MONAI ScaleIntensity(minv=0, maxv=1) of a raw-dtype batch in ONE pass (the `ScaleIntensityd`
 * of cinema/mae/pretrain.py:184 applied after the upload instead of on 16 CPU workers):
 *   out[b, i] = (float(raw[b, i]) - lo[b]) * inv[b],  inv[b] = hi[b] > lo[b] ? 1 / (hi[b] - lo[b]) : 0   (fp32)
 * raw: B samples of n_per_sample elements of raw_dtype (CB_DT_U8 / CB_DT_I16 / CB_DT_U16 / CB_DT_F32), contiguous;
 * lo / hi: fp32 [B] per-sample minimum / maximum (the shard index carries them); a constant sample maps to 0. */
int cb_scale_intensity(const void* raw, int raw_dtype, const float* lo, const float* hi, float* out, int B,
                       long long n_per_sample, void* stream);
/* Input pipeline, device side, augmented: RandZoom(keep_size) -> ScaleIntensity -> SpatialPad("end") in two passes
 * (`RandZoomd` trilinear for SAX / bicubic for LAX, `ScaleIntensityd`, `SpatialPadd`: cinema/mae/pretrain.py:163-199; MONAI
 * `Zoom` = torch `interpolate(scale_factor, align_corners=False)` of the unpadded frame, centre-padded with zeros / centre-
 * cropped back to its own size).  raw: B frames of raw_dtype stored in the model's input size `size[nd]` (host array; the
 * frame occupies the corner [0, extent[b]) of it); extent: int32 [B][3] device (per-sample frame size, unused axes 1);
 * zoom: fp32 [B] device (1 = identity, bit-exact); out: fp32 [B][prod size]; keys_ws: 2 * B uint32 device scratch (per-
 * sample min / max of the zoomed frame).  nd = 3 && cubic = 0 (trilinear) or nd = 2 && cubic = 1 (bicubic, A = -0.75). */
int cb_zoom_intensity(const void* raw, int raw_dtype, const int* extent, const float* zoom, float* out, void* keys_ws,
                      int B, int nd, const int* size, int cubic, void* stream);
/* dst[i] = bf16(src[i] * scale * f), f = 1 (scale_dev NULL), *scale_dev (group == 0) or scale_dev[i / group] (group > 0,
 * a multiple of 4 dividing n: per-sample factors, the backward of DropPath, cinema/vit.py:606,608) */
int cb_scale_cast_bf16(const float* src, void* dst, long long n, const float* scale_dev, float scale, long long group,
                       void* stream);

/* patchify / unpatchify of a contiguous (B, C, S1..Sn) tensor, n in 1..4, element size 2 or 4 bytes:
 * tokens (B, prod(grid), prod(patch)*C), channel fastest ("nchpwqdr->nhwdpqrc", cinema/vit.py:67-256).
 * inverse=0: image -> tokens; inverse=1: tokens -> image. */
int cb_patchify(const void* src, void* dst, int B, int C, int ndim, const int* spatial, const int* patch,
                int elem_bytes, int inverse, void* stream);

/* tokens of a SUBSET of patches gathered from an arbitrarily strided source, written as bf16 rows:
 * out[(b,i), e] = src[b*sb + c*sc + sum_a (g_a*p_a + o_a)*s_a], token = idx[b,i] (idx NULL = all tokens),
 * e ordered (o_1..o_n, c) when chan_last=1 (PatchEmbed / patchify order) or (c, o_1..o_n) when
 * chan_last=0 (the flattened conv-weight order of a kernel==stride convolution).
 * src_dtype: CB_DT_F32 or CB_DT_BF16.  This is patchify + x[~mask] fused (cinema/vit.py:338-342 +
 * cinema/mae/mae.py:550; cinema/convvit.py:284-288). */
int cb_gather_patches(const void* src, int src_dtype, long long sb, long long sc, const long long* sstride, int B,
                      int C, int ndim, const int* grid, const int* patch, const int* idx, int k, int chan_last,
                      void* out, void* stream);
/* backward of cb_gather_patches: scatters bf16/fp32 rows back into a strided gradient buffer; accumulate=0
 * overwrites the touched elements (buffer pre-zeroed by the caller), accumulate=1 adds to them. */
int cb_scatter_patches(const void* rows, int rows_dtype, void* dst, int dst_dtype, long long sb, long long sc,
                       const long long* sstride, int B, int C, int ndim, const int* grid, const int* patch,
                       const int* idx, int k, int chan_last, int accumulate, void* stream);

/* Rotary position embedding, standalone contract of cinema/rotary.py:25-50,108-128: x, y (B, n_tokens, H, d)
 * contiguous (fp32 or bf16), cos / sin (n_tokens, rotary_dim/2) fp32 tables, NeoX half rotation over the first
 * rotary_dim channels, the rest copied.  transpose=1 applies the inverse rotation (the backward). */
int cb_rope_apply(const void* x, void* y, int dtype, const float* cos_t, const float* sin_t, int B, int n_tokens,
                  int H, int d, int rotary_dim, int transpose, void* stream);

/* ---- masked-pixel MSE (cinema/mae/mae.py:107-152) fused with the target patchify (:597) --- *
 * image fp32 contiguous (B,C,S1..Sn); pred fp32 (B,n_drop,E) rows for the tokens drop_idx[b,j];
 * slot[b,t] from cb_mask_to_index, mask (B,n_tok).  acc (fp32[8], pre-zeroed) receives
 *   (acc[3], acc[4] must be pre-set to -inf when norm_target)
 *   [0] sum (pred-target)^2 over masked patches   [1] sum of per-patch means   [2] sum of per-patch
 *   unbiased stds   [3] max normalised target   [4] max pred   (3,4 only when norm_target)
 * diff (fp32, same shape as pred, may be NULL) receives pred - target for the backward. */
int cb_masked_mse_fwd(const float* image, int B, int C, int ndim, const int* spatial, const int* patch,
                      const unsigned char* mask, const int* slot, const float* pred, int n_drop, int norm_target,
                      float eps, float* acc, float* diff, void* stream);

/* View-averaged loss and metrics from the accumulators of cb_masked_mse_fwd (acc: n_views x 8 floats).
 * sq_count[v] = B*n_drop*E, patch_count[v] = B*n_tok (HOST arrays).  out[0] = mean of the finite per-view
 * losses (NaN if none, cinema/mae/mae.py:604-608 without the host sync); out[1+5v..] = {mse, target_mean,
 * target_std, normed_target_max, pred_max} of view v; scales[v] = d loss / d (pred - target) factor of view v. */
int cb_mae_loss_finalize(const float* acc, int n_views, const float* sq_count, const float* patch_count, float* out,
                         float* scales, void* stream);

/* ---- ConvMAE stem on the visible patches only (token-major channel-last rows x[(b, i), p, c]) ------------ *
 * keep (B, nk) ascending visible token ids, mask (B, n_tok) 1 = removed, slot from cb_mask_to_index.
 * grid_tok: ViT token grid; f: positions per token and axis at this stem level (e.g. {4,4,1} then {2,2,1}).
 * out[b, i*P + p] = id of position p of token keep[b,i] in the level grid (grid_tok * f): the index list that lets
 * cb_gather_patches read the k==s patch-conv inputs of the first stem level straight from the image. */
int cb_expand_token_index(const int* keep, int B, int nk, int ndim, const int* grid_tok, const int* f, int* out,
                          void* stream);
/* Depth-wise 5^ndim convolution, "same" zero padding, bf16 in / out, fp32 accumulation (cinema/conv.py:385,410-411:
 * dw_conv(mask * x)): neighbours are looked up through the token map, masked or out-of-image neighbours are zero.
 * w: bf16 (C, 1, 5, 5[, 5]); bias fp32 (C) or NULL.  transpose=1 correlates with the flipped kernel = gradient
 * with respect to the input. */
int cb_dwconv_tokens(const void* in, void* out, const void* w, const float* bias, const unsigned char* mask,
                     const int* slot, const int* keep, int B, int nk, int C, int ndim, const int* grid_tok, const int* f,
                     int transpose, void* stream);
/* dw (fp32, (C, 1, 5, 5[, 5])) += sum dy * shifted in;  db (fp32 (C), may be NULL) += sum dy. */
int cb_dwconv_tokens_wgrad(const void* in, const void* dy, float* dw, float* db, const unsigned char* mask,
                           const int* slot, const int* keep, int B, int nk, int C, int ndim, const int* grid_tok,
                           const int* f, void* stream);

/* ---- optimiser step over the flat arena ------------------------------------------------- *
 * out[0] += sum x^2 (global gradient norm, torch.nn.utils.clip_grad_norm_ at cinema/optim.py:206). */
int cb_sumsq_f32(const float* x, long long n, float* out, void* stream);
/* AdamW (torch.optim.AdamW semantics, decoupled weight decay) on a flat segment, fused with gradient
 * clipping and the bf16 shadow refresh.  hyper (DEVICE) = {lr, 1 - beta1^t, 1 - beta2^t}.  gnorm_sq
 * (DEVICE, may be NULL) = sum of squared gradients of the whole model BEFORE grad_scale; the applied
 * gradient is g * grad_scale * min(1, max_norm / (sqrt(gnorm_sq) * grad_scale + 1e-6)); a non-finite
 * norm skips the step (GradScaler semantics, cinema/optim.py:204-212).  p16 may be NULL. */
int cb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, long long n, const float* hyper, float beta1,
                  float beta2, float eps, float weight_decay, const float* gnorm_sq, float max_norm, float grad_scale,
                  void* stream);

/* Programmatic dependent launch (griddepcontrol) for all kernels of the library: on by default (env CB_PDL=0 turns it
 * off); returns the previous setting.  Per-kernel event timing should switch it off: with it, a kernel's prologue runs
 * in the tail of its predecessor. */
int cb_set_pdl(int enabled);

/* Segmentation loss of the fine-tuning scripts (cinema/segmentation/train.py:77-103): F.cross_entropy(ignore_index = -1) +
 * MONAI DiceLoss(include_background=False, softmax=True) over channel-first logits (B, C, S = prod(spatial)), 2 <= C <= 8,
 * logits fp32 or bf16, labels (B, S) of label_dtype 0 int64 / 1 int32 / 2 int16 / 3 uint8 (-1 = unlabelled: no cross-entropy
 * term, background for the Dice term).  Forward: one pass + a one-block finalize.  acc: (B * C * 3 + 2) fp32 scratch (zeroed
 * here); out[0..2] = loss, cross-entropy, mean Dice loss; coef: (B * C * 2 + 1) fp32 kept for the backward.  Backward: one pass,
 * dlogits (same dtype / layout as logits) = grad_out[0] * d loss / d logits. */
int cb_seg_loss_fwd(const void* logits, int logits_dtype, const void* labels, int label_dtype, int B, int C, long long S,
                    float* acc, float* out, float* coef, void* stream);
int cb_seg_loss_bwd(const void* logits, int logits_dtype, const void* labels, int label_dtype, int B, int C, long long S,
                    const float* coef, const float* grad_out, void* dlogits, void* stream);
/* Diagnostics: register a DEVICE buffer of at least 1024 long long (or NULL to switch off, the default).  While set, one
 * CTA in the middle of the grid of every second-generation attention launch stamps clock64() at the phase boundaries of
 * its softmax / compute warps and of its MMA-issuing warp (slot layout and reader: tools/attn_trace.py).  This is how
 * profiles/r0x_attention_clock_trace.md is produced; it is not part of any product path. */
int cb_attention_trace(long long* device_buf);
/* Diagnostics only, as cb_attention_trace: the leader CTA of one pair in the middle of every cb_gemm_bf16 launch stamps
 * clock64() at the phase boundaries of its producer, MMA and first epilogue warp (1024 x int64, tools/gemm_trace.py). */
int cb_gemm_trace(long long* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* CINEMA_B200_H_ */
