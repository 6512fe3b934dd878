/* libldit_b200 -- C ABI of the B200-native DiT backbone kernels.
 *
 * The reference (matteociccozzi/LayoutDiT) has no FFI of its own: its hot path is the Python
 * call `self.dit(x).hidden_states` (src/layoutdit/modeling/dit_backbone.py:47) into
 * HuggingFace `BeitModel`, which dispatches one ATen/cuBLAS/cuDNN library kernel per op.
 * Each entry point below replaces the library call(s) named beside it
 * (HF = transformers/models/beit/modeling_beit.py, transformers 5.5.0;
 *  R  = src/layoutdit/modeling/dit_backbone.py of the reference).
 *
 * Conventions
 *  - All pointers are DEVICE pointers into memory owned by the caller (PyTorch in the shipped
 *    host code).  The library never allocates, frees or retains device memory.
 *  - `stream` is a cudaStream_t passed as void*.  Every call only enqueues work on it: no
 *    synchronisation, no use of the default stream, safe to capture into a CUDA graph.
 *  - Return value: 0 = OK; negative = argument / alignment error (LDIT_E_*); positive = a
 *    cudaError_t (or 10000 + CUresult from the tensor-map encoder).  No C++ exception crosses
 *    the ABI.  ldit_error_string() describes any code.
 *  - Alignment: base pointers 16-byte aligned; row pitches multiples of 16 bytes.
 *  - bf16 = __nv_bfloat16 (2 bytes); f32 = float.
 */
#ifndef LDIT_H_
#define LDIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDIT_OK 0
#define LDIT_E_NULL (-1)       /* required pointer is NULL */
#define LDIT_E_SHAPE (-2)      /* unsupported dimension */
#define LDIT_E_ALIGN (-3)      /* pointer / pitch alignment */
#define LDIT_E_DTYPE (-4)      /* unknown dtype code */
#define LDIT_E_NO_DRIVER (-5)  /* cuTensorMapEncodeTiled not resolvable (no CUDA driver) */
#define LDIT_E_UNSUPPORTED (-6) /* experimental entry point / variant not compiled into this build */

#define LDIT_DTYPE_F32 0
#define LDIT_DTYPE_F16 1
#define LDIT_DTYPE_BF16 2

int ldit_version(void);
const char* ldit_error_string(int code);

/* nn.LayerNorm(D, eps) -- HF:458,460 (called at HF:478,495); ATen native_layer_norm.
 * x f32 [rows, D] -> y bf16 [rows, D].  D must be a multiple of 128, D <= 2048. */
int ldit_layernorm(const void* x, const void* gamma, const void* beta, void* y, int rows, int D, float eps, void* stream);

/* nn.Linear with fused bias -- query/key/value at HF:324-338 run as ONE GEMM over the
 * concatenated [3D, D] weight (key has no bias: pass zeros in that third of `bias`).
 * out bf16 [M, N] = A bf16 [M, K] x W bf16 [N, K]^T + bias f32 [N] (bias may be NULL). */
int ldit_gemm_bias(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, void* stream);

/* BeitIntermediate, HF:428-432: out bf16 [M, N] = gelu_erf(A x W^T + bias). */
int ldit_gemm_bias_gelu(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, void* stream);

/* BeitSelfOutput + layer-scale + residual (HF:383, 488-492) and BeitOutput + layer-scale +
 * residual (HF:442, 500-504):  x f32 [M, N] <- x + scale (.) (A x W^T + bias), in place.
 * scale may be NULL (layer_scale_init_value <= 0, HF:462-467). */
int ldit_gemm_bias_scale_residual(const void* A, const void* W, const void* bias, const void* scale, void* x, int M, int N,
                                  int K, void* stream);
/* (The branch scale (.) (A x W^T + bias) is rounded to bf16 before it is added: x then receives bit for bit what the two-step
 * form below adds, so the forward's numbers do not depend on which form a geometry uses.)
 * Plain accumulation without bias, scale or rounding, acc f32 [M, N] += A x W^T: the wgrad GEMMs of the backward. */
int ldit_gemm_accumulate(const void* A, const void* W, void* acc, int M, int N, int K, void* stream);

/* The same residual block in two steps, for a residual stream that is about to be normalised anyway:
 *   ldit_gemm_bias_scale:  out bf16 [M, N] = scale (.) (A x W^T + bias)          (the layer-scaled branch; scale may be NULL)
 *   ldit_add_layernorm:    x f32 [rows, D] += branch bf16 [rows, D];  y bf16 = LayerNorm(x)   (HF:488-495 / 500-504 + HF:478)
 * The GEMM then ends in a plain bf16 tile store and the residual add rides on the LayerNorm that streams x anyway.
 * y may alias branch.  D multiple of 128, D <= 2048. */
int ldit_gemm_bias_scale(const void* A, const void* W, const void* bias, const void* scale, void* out, int M, int N, int K,
                         void* stream);
int ldit_add_layernorm(void* x, const void* branch, const void* gamma, const void* beta, void* y, int rows, int D, float eps,
                       void* stream);

/* BeitIntermediate + BeitOutput + layer scale + residual (HF:428-432, 441-445, 500-504) as ONE persistent kernel:
 *   h bf16 [M, I] = gelu_erf(a bf16 [M, D] x W1 bf16 [I, D]^T + b1);   x f32 [M, D] += lam2 (.) (h x W2 bf16 [D, I]^T + b2)
 * Same arithmetic as ldit_gemm_bias_gelu followed by ldit_gemm_bias_scale_residual, bit for bit; the two GEMMs
 * share one balanced tile schedule and fc2 tiles start as soon as the rows of h they read are complete.
 *  sched  DEVICE int32 [ldit_mlp_clusters(), stride]: the tile lists from ldit_mlp_schedule(M, D, I, ...) (host
 *         function: call with host_sched == NULL to get `stride`, then with a buffer of clusters * stride ints)
 *  ready  DEVICE int32 [2 * ceil(M / 256)], zero-initialised once by the caller; left zeroed by every call
 * D and I must both be multiples of 192 or both multiples of 256 (else LDIT_E_SHAPE: use the two-call form).
 * EXPERIMENTAL (measured 14-21 % slower than the two launches): only in -DLDIT_EXPERIMENTAL builds, otherwise
 * ldit_mlp_schedule / ldit_mlp_fused return LDIT_E_UNSUPPORTED. */
int ldit_mlp_clusters(void);
int ldit_mlp_schedule(int M, int D, int I, int* host_sched, int capacity);
int ldit_mlp_fused(const void* a, const void* W1, const void* b1, void* h, const void* W2, const void* b2, const void* lam2,
                   void* x, int M, int D, int I, const int* sched, int sched_stride, int* ready, void* stream);

/* BeitEmbeddings.forward, HF:161-184 (Conv2d k16 s16 HF:218 + flatten/transpose HF:220 +
 * cat(cls) HF:176-177 + position add HF:179-180), as an im2col GEMM.
 *  pixels   [B, 3, H, W] of dtype `pixel_dtype` (LDIT_DTYPE_*), contiguous NCHW
 *  w        bf16 [D, 768]  flattened conv weight
 *  pos_bias f32 [P, D]     position rows 1..P (already resized to Gh x Gw) + conv bias
 *  cls_pos  f32 [D]        cls_token + position row 0
 *  scratch  bf16 [B*P, 768] workspace for the im2col operand
 *  x        f32 [B, P+1, D] residual stream (output)
 * H, W multiples of 16; P = (H/16)(W/16). */
int ldit_patch_embed(const void* pixels, int pixel_dtype, const void* w, const void* pos_bias, const void* cls_pos,
                     void* scratch, void* x, int B, int H, int W, int D, void* stream);

/* BeitEmbeddings.forward for 16-bit pixels as ONE TMA-fed im2col GEMM (no scratch, no separate gather / CLS kernels): the A
 * operand tiles are gathered by 5-D TMA boxes (px, py, patch column, patch row, image x channel) straight out of the NCHW
 * batch, patches outside the grid are zero-filled by the TMA unit, token row 0 of every image is written by the same kernel.
 *  pixels  [B, 3, H, W] fp16 or bf16 (pixel_dtype LDIT_DTYPE_F16 / LDIT_DTYPE_BF16; the reference feeds fp16 under autocast,
 *          R:src/layoutdit/training/trainer.py:155,168)
 *  w       [D, 768] flattened conv weight IN THE SAME 16-BIT TYPE AS THE PIXELS (tcgen05 kind::f16 traps on mixed f16 / bf16)
 * everything else as ldit_patch_embed, which takes this path by itself for bf16 pixels. */
int ldit_patch_embed_tma(const void* pixels, int pixel_dtype, const void* w, const void* pos_bias, const void* cls_pos, void* x,
                         int B, int H, int W, int D, void* stream);
/* 1 if the one-launch TMA path is expected to beat gather pass + plain GEMM for this geometry (it tiles every image on its
 * own, so narrow grids and wide D can lose), else 0; ldit_patch_embed applies the same rule to bf16 pixels. */
int ldit_patch_embed_tma_preferred(int B, int H, int W, int D);

/* ldit_patch_embed with the detector's input transform fused into the patch gather (SURVEY.md section 8 row
 * f3): GeneralizedRCNNTransform as configured at R:src/layoutdit/modeling/model.py:44-56 -- per page
 * normalize (image - mean) / std, then F.interpolate(size=(H, W), mode="bilinear", align_corners=False)
 * (TV:models/detection/transform.py normalize() / _resize_image_and_masks() with fixed_size), then batching --
 * followed by BeitEmbeddings.forward as above.  The resized batch [B, 3, H, W] is never written.
 *  pages    DEVICE array of B device pointers, page b = [3, page_hw[2b], page_hw[2b+1]] contiguous CHW of
 *           dtype `pixel_dtype` (any size >= 1x1 per page, values as the reference feeds them: floats in [0, 1])
 *  page_hw  DEVICE int32 [B, 2] = (height, width) of each page
 *  max_page_w  host-known upper bound of the page widths: sizes the shared-memory row staging of the fast
 *           gather (6 rows of max_page_w pixels must fit 200 KB); 0 = unknown, use the direct gather
 *  mean*, std*  per-channel normalisation constants (the reference: 0.5 each)
 * everything else as ldit_patch_embed; (H, W) = the fixed size the detector resizes to (224, 224). */
int ldit_patch_embed_pages(const void* const* pages, const int* page_hw, int max_page_w, int pixel_dtype, float mean0, float mean1,
                           float mean2, float std0, float std1, float std2, const void* w, const void* pos_bias, const void* cls_pos,
                           void* scratch, void* x, int B, int H, int W, int D, void* stream);

/* BeitSelfAttention core, HF:275-298 / F.scaled_dot_product_attention at HF:356-364, plus the
 * head merge HF:365-367.  qkv bf16 [B*N, 3D] (Q | K | V, heads contiguous, head_dim 64) ->
 * ctx bf16 [B*N, D].  bias_table: NULL, or f32 [heads, T] with T = (2Gh-1)(2Gw-1)+3 -- the
 * relative_position_bias_table resized to this window (HF:556-571), transposed; the kernel
 * gathers it in-tile with the index rule of HF:522-544 instead of materialising [heads,N,N]. */
int ldit_attention(const void* qkv, void* ctx, const void* bias_table, int B, int N, int heads, int Gh, int Gw, void* stream);

/* Tap extraction + F.interpolate(bilinear, align_corners=False), R:50-61.
 *  x f32 [B, 1+Gh*Gw, D] (hidden state incl. CLS) -> out bf16, channels-last memory
 *  [B, oh, ow, D] with oh = floor(Gh*scale), ow = floor(Gw*scale); scale in {4, 2, 1, 0.5}
 *  (any positive scale is accepted). */
int ldit_resample_taps(const void* x, void* out, int B, int Gh, int Gw, int D, float scale, void* stream);

/* ---- FPN on the taps: TV = torchvision/ops/feature_pyramid_network.py (0.26),
 * as instantiated at R:dit_backbone.py:80-90.  All tensors bf16 channels-last.
 *
 * Lateral 1x1 convolutions (TV:111-116, 187) are ldit_gemm_bias calls on the token grid BEFORE resampling
 * (1x1 conv and bilinear resampling commute); ldit_fpn_merge then does R:57-59's bilinear resample of that
 * lateral plus the nearest-neighbour top-down add (TV:188-190):
 *   out [B, oh, ow, C] = bilinear(lat [B, Gh, Gw, C], scale) + nearest(top [B, top_h, top_w, C] -> oh x ow)
 * with oh = floor(Gh*scale), ow = floor(Gw*scale); top may be NULL (coarsest level). */
int ldit_fpn_merge(const void* lat, const void* top, void* out, int B, int Gh, int Gw, int C, float scale, int top_h,
                   int top_w, void* stream);

/* Conv2d(Cin, Cout, 3, padding=1) + bias (the FPN layer_blocks, TV:118-124, 193) as an implicit tcgen05 GEMM.
 *  in  bf16 [B, H, W, Cin]   w bf16 [Cout, 9*Cin] = conv.weight.permute(0, 2, 3, 1) flattened ((ky, kx, cin))
 *  bias f32 [Cout] or NULL   out bf16 [B, H, W, Cout].  Cin multiple of 64, Cout multiple of 128. */
int ldit_conv3x3_bias(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int Cin, int Cout,
                      void* stream);

/* Same convolution with an fp32 output map (out f32 [B, H, W, Cout]): what the detection heads behind the FPN expect when
 * the detector runs in fp32 (torchvision's RPN / RoI heads hold fp32 weights, R:src/layoutdit/modeling/model.py:44-56);
 * the cast is the epilogue's store format, not an extra pass. */
int ldit_conv3x3_bias_f32(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int Cin, int Cout,
                          void* stream);

/* LastLevelMaxPool (TV:231-249): max_pool2d(kernel 1, stride 2) = out[b, y, x, :] = in[b, 2y, 2x, :];
 * out is [B, ceil(H/2), ceil(W/2), C].  bf16 maps; ldit_subsample2_f32 for fp32 maps. */
int ldit_subsample2(const void* in, void* out, int B, int H, int W, int C, void* stream);
int ldit_subsample2_f32(const void* in, void* out, int B, int H, int W, int C, void* stream);

/* Weight preparation, once per (weights, H, W): resize a table of rows, src f32 [h*w, C] -> dst f32 [oh*ow, C]
 * (+ add f32 [C] if not NULL), with ATen's rules for F.interpolate(size=(oh, ow), align_corners=False):
 *  bicubic != 0: the position table at a non-native patch grid (HF interpolate_pos_encoding, HF:138-159);
 *  bicubic == 0: bilinear, the relative-position bias table at a non-native window (HF:556-571).
 * Equal sizes copy exactly.  Not part of the per-forward launch sequence. */
int ldit_resize_rows(const void* src, void* dst, const void* add, int h, int w, int oh, int ow, int C, int bicubic, void* stream);

/* ---- Backward of the backbone (host side: layoutdit_b200/train.py).
 * The reference trains through torch.autograd over HF BeitModel (R:src/layoutdit/training/trainer.py:164-183, HF:469-508).
 * The eight GEMMs of a layer's backward are ldit_gemm_dgrad (dA = dY W) and ldit_gemm_wgrad (dW += dY^T A) below: the
 * forward's tcgen05 GEMM with MN-major operands, no transposed copies.  bf16 activations / gradients, fp32
 * residual-stream gradients and parameter gradients (ACCUMULATED into, like .grad). */
/* out bf16 [C, ld_out] (ld_out >= R) = in bf16 [R, C]^T  (utility: the first wgrad path used it; pad ld_out to a multiple
 * of 8 and zero the padding: a TMA operand's row pitch must be a multiple of 16 bytes) */
int ldit_transpose_bf16(const void* in, void* out, int R, int C, int ld_out, void* stream);
/* out f32 [C] += column sums of in bf16 [R, ld] over columns [0, C)  (bias gradients; C, ld even) */
int ldit_colsum_bf16(const void* in, void* out, int R, int C, int ld, void* stream);
/* erf-GELU (HF:430) on a stored pre-activation, and its backward dpre = dh * gelu'(pre); n elements, n % 8 == 0 */
int ldit_gelu(const void* pre, void* h, size_t n, void* stream);
int ldit_gelu_bwd(const void* dh, const void* pre, void* dpre, size_t n, void* stream);
/* y f32 [rows, D] = x f32 + lam (.) branch bf16   (HF:488-492 / 500-504; lam may be NULL); y may alias x */
int ldit_scale_residual(const void* x, const void* branch, const void* lam, void* y, int rows, int D, void* stream);
/* dbranch bf16 = lam (.) dy f32;  dlam f32 [D] += column sums of dy (.) branch  (dlam may be NULL) */
int ldit_scale_residual_bwd(const void* dy, const void* branch, const void* lam, void* dbranch, void* dlam, int rows, int D,
                            void* stream);
/* The same two with drop-path (HF:61-73, applied to the layer-scaled branch at HF:488-492 / 500-504): the branch of image b
 * (rows b*rows_per_image ..) is additionally scaled by row_scale f32 [rows / rows_per_image] (0 or 1 / keep_prob; NULL = 1). */
int ldit_scale_residual_rows(const void* x, const void* branch, const void* lam, const void* row_scale, int rows_per_image, void* y,
                             int rows, int D, void* stream);
int ldit_scale_residual_rows_bwd(const void* dy, const void* branch, const void* lam, const void* row_scale, int rows_per_image,
                                 void* dbranch, void* dlam, int rows, int D, void* stream);
/* nn.LayerNorm backward: dx_out f32 = (dx_in or 0) + d/dx of LN(x) gamma + beta under dy bf16; dgamma, dbeta f32 [D] +=.
 * dx_in may be NULL and may alias dx_out.  Row statistics are recomputed from x. */
int ldit_layernorm_bwd(const void* x, const void* gamma, const void* dy, const void* dx_in, void* dx_out, void* dgamma, void* dbeta,
                       int rows, int D, float eps, void* stream);
/* Input gradient of nn.Linear: dA bf16 [M, Kin] = dY bf16 [M, Nout] x W bf16 [Nout, Kin], W exactly as the forward holds it
 * (MN-major B operand: no transposed weight copy).  Nout, Kin multiples of 8. */
int ldit_gemm_dgrad(const void* dY, const void* W, void* dA, int M, int Nout, int Kin, void* stream);
/* Weight gradient of nn.Linear: dW f32 [Nw, Kw] += dY^T A with dY bf16 [T, Nw], A bf16 [T, Kw] row-major (T = tokens).
 * tcgen05 GEMM with MN-major operands (no transposed copies) and a split contraction (fp32 reduce-add; the summation
 * order of the pieces is not fixed).  Nw, Kw multiples of 8. */
int ldit_gemm_wgrad(const void* dY, const void* A, void* dW, int T, int Nw, int Kw, void* stream);
/* Backward of ldit_attention without relative-position bias: dqkv bf16 [B*N, 3D] from qkv and dctx bf16 [B*N, D].
 * tcgen05 kernel, one CTA per (image, head), P recomputed from qkv (nothing else is kept from the forward);
 * N <= 256 (two 128-row query tiles x two 128-key halves fill TMEM), LDIT_E_UNSUPPORTED beyond. */
int ldit_attention_bwd(const void* qkv, const void* dctx, void* dqkv, int B, int N, int heads, void* stream);
/* ldit_attention that also writes lse f32 [B, heads, N]: the log2-sum-exp of every row's scaled (and biased) logits --
 * what ldit_attention_bwd_flash needs to recompute P. */
int ldit_attention_lse(const void* qkv, void* ctx, const void* bias_table, void* lse, int B, int N, int heads, int Gh, int Gw,
                       void* stream);
/* Backward of ldit_attention for ANY sequence length, optionally with the relative-position table the forward used
 * (bias_table f32 [heads, T] as for ldit_attention; dbias f32 [heads, T] += its gradient, HF:522-544 index rule; both NULL
 * without a table): flash-style, one CTA per (image, head,
 * 128-key tile) with dV / dK in TMEM, walking the query tiles; P recomputed from `lse` (ldit_attention_lse), delta from
 * ctx (the forward's output) and dctx; dQ contributions are fp32 vector reductions into the workspace dq_acc
 * f32 [B*N, D] (zeroed here; summation order not fixed), cast into dqkv at the end.  delta: workspace f32 [B*heads*N]. */
int ldit_attention_bwd_flash(const void* qkv, const void* ctx, const void* lse, const void* dctx, void* dqkv, void* dq_acc, void* delta,
                             const void* bias_table, void* dbias, int B, int N, int heads, int Gh, int Gw, void* stream);
/* Adjoint of ldit_resample_taps (R:dit_backbone.py:50-61 under autograd): dout bf16 [B, floor(Gh*scale), floor(Gw*scale), D]
 * channels-last -> dx f32 [B, Gh*Gw + 1, D], rows 1..P written (the CLS row is left as it is: zero it first). */
int ldit_resample_taps_bwd(const void* dout, void* dx, int B, int Gh, int Gw, int D, float scale, void* stream);
/* out f32 [R] += sum over b of x f32 [B, R]  (gradients of the position / CLS embeddings, HF:161-184; R % 4 == 0) */
int ldit_batch_sum(const void* x, void* out, int B, int R, void* stream);

/* Bytes of the im2col scratch ldit_patch_embed needs. */
size_t ldit_patch_embed_scratch_bytes(int B, int H, int W);

/* Tuning knob: force the GEMM tile width (128, 192 or 256); 0 restores the automatic choice
 * (fewest persistent-schedule rounds x tile width).  Also settable with LDIT_GEMM_BN. */
void ldit_set_gemm_tile_n(int bn);
/* Tuning knob: 2 (default) = CTA pairs with tcgen05.mma.cta_group::2 (256 x BN tile per pair);
 * 1 = independent CTAs (128 x BN tile).  Also settable with LDIT_GEMM_CTAS. */
void ldit_set_gemm_cta_pair(int ctas);

/* 1 (default): kernels are launched with programmatic stream serialization (PDL) -- each kernel's
 * prologue overlaps the tail of its predecessor in the stream; 0: plain stream order.  Also LDIT_PDL. */
void ldit_set_pdl(int on);

/* Opt-in: every kernel enqueued ON `stream` after this call carries an L2 access-policy window (persisting) over
 * [ptr, ptr + bytes) -- meant for the fp32 residual stream x, which LayerNorms, reduce-add epilogues and taps revisit
 * all forward long while larger activations stream through L2 in between.  The window is host-side state keyed by
 * the stream (two engines, streams or devices in one process do not see each other's); (stream, NULL, 0, 0) removes
 * it.  `set_aside_cap` caps the persisting-L2 set-aside this call may request (0 = no cap): a window larger than the
 * set-aside persists only the fraction of its lines that fits (hitRatio = set-aside / window).  Side effect: grows
 * the device-wide cudaLimitPersistingL2CacheSize of the current device to min(bytes, cap, device maximum); it is never
 * shrunk again.  Returns a cudaError_t if the device refuses (MPS, MIG): the window is then off, the caller may go on. */
int ldit_set_l2_window(void* stream, void* ptr, size_t bytes, size_t set_aside_cap);

/* Bytes of workspace one forward needs for B pages of H x W at hidden size D / MLP width I, as the host module lays it
 * out in ONE allocation: residual stream x f32 [M, D], then a bf16 [M, D] (x and a contiguous: they form the L2 window),
 * then the wide buffer bf16 [M, max(3D, I, 768)] (QKV, MLP hidden and im2col scratch have disjoint lifetimes);
 * M = B ((H/16)(W/16) + 1); each part rounded up to 1 KB.  Outputs (taps / FPN maps) are the caller's. */
size_t ldit_workspace_bytes(int B, int H, int W, int D, int I);

/* Attention variant: 0 (default) = attention_v3 (two CTAs per SM, double-buffered scores).  1 = warp-level mma.sync,
 * 2 = one-tile-per-CTA tcgen05, 4 = the round-1 ping-pong tcgen05 kernel: superseded, only in -DLDIT_EXPERIMENTAL
 * builds (ldit_attention returns LDIT_E_UNSUPPORTED otherwise).  Also settable with LDIT_ATTN_IMPL. */
void ldit_set_attention_impl(int impl);
/* 1 if the library was built with -DLDIT_EXPERIMENTAL (superseded attention variants, ldit_mlp_fused), else 0. */
int ldit_has_experimental(void);

/* Diagnosis builds only (-DLDIT_A3_TIMELINE / -DLDIT_DEBUG_HOOKS): device buffers the attention / GEMM kernels record
 * clock64() stamps into.  In the product build these calls do nothing: no debug path is compiled into the kernels. */
void ldit_debug_attention_timeline(void* device_buffer);
void ldit_debug_gemm_timeline(void* device_buffer);

/* Number of kernels the library has enqueued since load / last reset (for gpu_launches). */
unsigned long long ldit_launch_count(void);
void ldit_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LDIT_H_ */
