/*
 * vnext_b200 — C ABI of the B200-native (sm_100a) hot path of VNext (SeqFormer / IDOL).
 *
 * Plain pointers and sizes only; no torch types.  Every entry point returns 0 on success and a
 * non-zero VNEXT_E* code on failure; vnext_last_error() returns the message for the calling thread
 * (the reference raises RuntimeError through AT_ASSERTM / AT_ERROR at the same places; the Python
 * binding turns a non-zero return into RuntimeError).  All `const void*` tensor arguments are
 * DEVICE pointers on the current CUDA device unless the function name ends in `_host`.  `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  Kernels are enqueued
 * asynchronously; nothing here synchronises the device except the `_host` variants.
 *
 * Reference interfaces replaced (paths under /root/reference/projects/IDOL/idol/models/ops,
 * byte-identical copy under projects/SeqFormer/seqformer/models/ops):
 *   vnext_msda_forward   <- ms_deform_attn_forward   src/ms_deform_attn.h:20-39,
 *                           ms_deform_attn_cuda_forward  src/cuda/ms_deform_attn_cuda.cu:20-80,
 *                           ms_deformable_im2col_cuda    src/cuda/ms_deform_im2col_cuda.cuh:923-954
 *   vnext_msda_backward  <- ms_deform_attn_backward  src/ms_deform_attn.h:41-61,
 *                           ms_deform_attn_cuda_backward src/cuda/ms_deform_attn_cuda.cu:83-152,
 *                           ms_deformable_col2im_cuda    src/cuda/ms_deform_im2col_cuda.cuh:956-1326
 * pybind names kept by the Python shim `MultiScaleDeformableAttention` (src/vision.cpp:13-16).
 */
#ifndef VNEXT_B200_H_
#define VNEXT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type of value / output / grad_output / grad_value */
enum {
    VNEXT_F32 = 0,  /* everything fp32 (reference: AT_DISPATCH_FLOATING_TYPES float) */
    VNEXT_F64 = 1,  /* everything fp64 (reference: double; used by gradcheck) */
    VNEXT_BF16 = 2, /* value/output(/grad_output/grad_value) bf16; sampling_loc, attn_weight and
                       their grads stay fp32; accumulation fp32.  New: the reference has no bf16. */
    VNEXT_F16 = 3   /* IEEE fp16 storage — vnext_linear_tc only (a + weight and/or out): 11 significant bits, i.e. at least what
                       a TF32 MMA keeps of an fp32 operand, at half the bytes; outputs saturate at +-65504 */
};

enum {
    VNEXT_OK = 0,
    VNEXT_EINVAL = 1,   /* bad argument (null pointer, non-positive size, unsupported dtype) */
    VNEXT_ECUDA = 2,    /* CUDA runtime / launch error (message carries cudaGetErrorString) */
    VNEXT_ENOMEM = 3
};

const char *vnext_last_error(void);
/* "vnext_b200 <version> sm_100a" */
const char *vnext_version(void);
/* number of kernels this library has launched since load (all threads); for bench `gpu_launches` */
unsigned long long vnext_launch_count(void);

/*
 * Multi-scale deformable attention, forward.
 *   value              [batch, spatial_size, num_heads, channels]             (dtype)
 *   spatial_shapes     int64 [num_levels, 2] = (H_l, W_l), DEVICE memory (as the reference passes it)
 *   level_start_index  int64 [num_levels], DEVICE memory
 *   sampling_loc       [batch, num_query, num_heads, num_levels, num_point, 2] = (x, y) in [0,1]
 *   attn_weight        [batch, num_query, num_heads, num_levels, num_point]
 *   output             [batch, num_query, num_heads*channels]  — fully overwritten (no pre-zeroing needed)
 * Semantics (bit-exact indexing): h_im = y*H - 0.5, w_im = x*W - 0.5; a sample contributes only if
 * h_im > -1 && w_im > -1 && h_im < H && w_im < W; corners outside [0,H-1]x[0,W-1] contribute 0
 * (ms_deform_im2col_cuda.cuh:55-78, 285-288).  The reference's `im2col_step` chunking is a host loop
 * with no numerical effect; the whole batch is one launch here.
 */
int vnext_msda_forward(int dtype, const void *value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const void *sampling_loc,
                       const void *attn_weight, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point, void *output,
                       void *stream);

/*
 * Forward with the module's pre-ops fused in (softmax over L*K, sampling-location arithmetic), replacing the
 * ATen passes of MSDeformAttn.forward around the op (ops/modules/ms_deform_attn.py:98-108; SeqFormer variant
 * S/models/ops/modules/ms_deform_attn.py:98-121, 140-176):
 *   offsets_logits  fp32 [batch*num_query, row_pitch]: per row M*L*K*2 raw sampling offsets (head, level, point, xy)
 *                   followed by M*L*K raw attention logits — i.e. the concatenated sampling_offsets /
 *                   attention_weights projections of the query
 *   reference_points fp32 [ref_shared ? 1 : batch, num_query, num_levels, ref_dim], ref_dim 2 (x, y) or 4 (x, y, w, h)
 *   loc = ref_xy + off / (W_l, H_l)            (ref_dim 2)      loc = ref_xy + off / num_point * ref_wh * 0.5  (ref_dim 4)
 *   attn = softmax over the L*K logits of each (query, head);   output as vnext_msda_forward.
 * Tile kernels only: channels == 32, num_levels*num_point in {4, 8, 16, 32}, value fp32 or bf16.
 */
int vnext_msda_forward_fused(int dtype, const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *offsets_logits, int row_pitch,
                             const float *reference_points, int ref_dim, int ref_shared, int batch, int spatial_size,
                             int num_heads, int channels, int num_levels, int num_query, int num_point, void *output,
                             void *stream);

/*
 * Forward for SELF-attention over the pixel grid (the encoder shape: num_query == spatial_size, query i IS pixel i of
 * the flattened pyramid — DeformableTransformerEncoderLayer, projects/IDOL/idol/models/deformable_transformer.py:
 * 230-261).  Same result as vnext_msda_forward / vnext_msda_forward_fused for ANY sampling locations; it is faster
 * when a query's samples fall near its own pixel, because each (query tile, head) stages a window of every level in
 * shared memory with TMA box loads (zero fill outside the level = the reference's border rule) and gathers from
 * there; samples that leave the window by more than `margin` pixels are fetched from global memory in the same loop.
 *   spatial_shapes_host / level_start_index_host: HOST copies of the int64 tensors (the TMA descriptors are built from
 *   them per call); levels must tile [0, spatial_size) in order.
 *   Front-end: either (sampling_loc, attn_weight) as vnext_msda_forward, or — when offsets_logits != NULL — the fused
 *   one of vnext_msda_forward_fused (offsets_logits, row_pitch, reference_points, ref_dim == 2, ref_shared).
 *   margin >= 0: the shared-memory window kernel (csrc/msda_window.cu) with that margin; margin < 0: the cache-blocked kernel
 *   (csrc/msda_cached.cu: 2-D query tiles x head, the L1 does the staging).  With margin < 0, offsets_logits != NULL and
 *   reference_points == NULL mean the rows already hold sampling LOCATIONS | normalised attention WEIGHTS in the same layout
 *   (what vnext_linear_msda_front writes): no front-end arithmetic in the sampler.
 * Needs channels == 32, num_levels <= 4, num_levels*num_point in {8, 16, 32}, value fp32 or bf16 (16-byte aligned).
 */
int vnext_msda_forward_selfattn(int dtype, const void *value, const int64_t *spatial_shapes_host,
                                const int64_t *level_start_index_host, const void *sampling_loc,
                                const void *attn_weight, const float *offsets_logits, int row_pitch,
                                const float *reference_points, int ref_dim, int ref_shared, int batch,
                                int spatial_size, int num_heads, int channels, int num_levels, int num_point,
                                int margin, void *output, void *stream);

/*
 * Backward.  grad_output [batch, num_query, num_heads*channels] (dtype).
 * Outputs are fully (over)written: grad_value like value (zero-filled then accumulated with
 * atomics, as the reference), grad_sampling_loc like sampling_loc, grad_attn_weight like attn_weight.
 * `workspace`: only for VNEXT_BF16 — fp32 scratch of vnext_msda_backward_workspace_bytes() bytes
 * (grad_value is accumulated in fp32 and rounded once); ignored (may be NULL) otherwise.
 */
int vnext_msda_backward(int dtype, const void *value, const int64_t *spatial_shapes,
                        const int64_t *level_start_index, const void *sampling_loc,
                        const void *attn_weight, const void *grad_output, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels,
                        int num_query, int num_point, void *grad_value, void *grad_sampling_loc,
                        void *grad_attn_weight, void *workspace, void *stream);

size_t vnext_msda_backward_workspace_bytes(int dtype, int batch, int spatial_size, int num_heads,
                                           int channels);

/*
 * Host-buffer variants (what a non-torch caller binds): all pointers are HOST memory (pinned or
 * pageable); device staging buffers are owned by the library and reused across calls; the call
 * copies inputs H2D, runs the kernel, copies outputs D2H and synchronises before returning.
 */
int vnext_msda_forward_host(int dtype, const void *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const void *sampling_loc,
                            const void *attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            void *output);

int vnext_msda_backward_host(int dtype, const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const void *sampling_loc,
                             const void *attn_weight, const void *grad_output, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, void *grad_value,
                             void *grad_sampling_loc, void *grad_attn_weight);

/*
 * Linear layer / 1x1 convolution on the tcgen05 tensor cores with a fused epilogue:
 *     out[M, N] = epilogue( a[M, K] . weight[N, K]^T + bias[N] )
 * fp32 storage, TF32 tensor-core math, fp32 accumulation.  `weight` is the nn.Linear / 1x1-conv weight as
 * the reference checkpoint stores it.  Replaces F.linear + separate bias/ReLU/residual/LayerNorm kernels of
 * projects/SeqFormer/seqformer/models/deformable_transformer.py:155-169 (forward_ffn, norm1/norm2) and
 * models/ops/modules/ms_deform_attn.py:94-100,114 (value/offset/weight/output projections).
 *   flags: VNEXT_EPI_RELU | VNEXT_EPI_RESIDUAL (+ residual[M, ldr]) | VNEXT_EPI_LAYERNORM (over the N == 256
 *          outputs of a row, after bias/residual; gamma, beta [N]); ReLU, when set, is applied last.
 *   bias2d [bias_period, ldb2] (optional): geometry-dependent bias added as bias2d[row % bias_period][col] — used to
 *          fold `(x + pos) W` into `x W + (pos W + b)` with the clip's position embedding projected once per layer.
 *   row_mask [bias_period] (optional): rows with row_mask[row % bias_period] != 0 get columns [0, mask_cols) zeroed
 *          (the reference's value.masked_fill(padding_mask, 0), ms_deform_attn.py:96-97).
 * lda / ldo / ldr / ldb2 are row pitches in elements.  a, weight: 16-byte aligned, lda % 4 == 0, K % 4 == 0 (TMA).
 */
enum { VNEXT_EPI_RELU = 1, VNEXT_EPI_RESIDUAL = 2, VNEXT_EPI_LAYERNORM = 4, VNEXT_EPI_GELU = 8 /* exact erf GELU, applied last */ };
int vnext_linear_tf32(const float *a, int lda, const float *weight, const float *bias, const float *bias2d,
                      int bias_period, int ldb2, const unsigned char *row_mask, int mask_cols,
                      const float *residual, int ldr, const float *ln_gamma, const float *ln_beta, float ln_eps,
                      float *out, int ldo, int M, int N, int K, int flags, void *stream);

/*
 * Same kernel, operand / output formats chosen by the caller (bf16 activations end to end; BASELINE config 4):
 *   in_dtype  (VNEXT_F32 | VNEXT_BF16 | VNEXT_F16): element type of `a` AND `weight` (fp32 -> kind::tf32, bf16 / fp16 ->
 *             kind::f16 MMAs); out_dtype may be VNEXT_F16 too (no residual then)
 *   out_dtype (VNEXT_F32 | VNEXT_BF16): element type of `out` AND `residual`
 * bias, bias2d, LayerNorm gamma/beta and all accumulation / statistics stay fp32.  bf16 rows (a, weight, out, residual)
 * must be 16-byte aligned with pitches that are multiples of 8 elements.  The reference has no bf16 path; its AMP
 * route is torch.autocast (detectron2/engine/train_loop.py:357-408).
 */
int vnext_linear_tc(int in_dtype, int out_dtype, const void *a, int lda, const void *weight, const float *bias,
                    const float *bias2d, int bias_period, int ldb2, const unsigned char *row_mask, int mask_cols,
                    const void *residual, int ldr, const float *ln_gamma, const float *ln_beta, float ln_eps, void *out,
                    int ldo, int M, int N, int K, int flags, void *stream);

/*
 * COCO run-length encoding of boolean masks on the device (csrc/rle_encode.cu) — what the reference's evaluator computes on
 * the host for every predicted mask right after the hot path: `mask_util.encode(np.array(mask[:, :, None], order="F"))[0]`
 * (projects/IDOL/idol/data/ytvis_eval.py:174-214, projects/SeqFormer/seqformer/data/ytvis_eval.py:171-211; pycocotools
 * maskApi.c rleEncode + rleToString).  masks [n_masks, H, W] bytes (row-major, 0 / non-zero); the strings of all masks are
 * written back to back into chars (device), char_offsets [n_masks + 1] int64 (device): mask i's "counts" string is
 * chars[char_offsets[i] : char_offsets[i + 1]].  char_offsets[n_masks] is the total even when it exceeds chars_capacity (then
 * only the first chars_capacity bytes were written: call again with a larger buffer).  workspace: device memory, 256-byte
 * aligned, at least vnext_rle_workspace_bytes(n_masks, H, W) bytes.  Six launches of fixed grid size (graph capturable).
 */
long long vnext_rle_workspace_bytes(int n_masks, int H, int W);
int vnext_rle_encode(const unsigned char *masks, int n_masks, int H, int W, void *workspace, long long workspace_bytes,
                     long long *char_offsets, unsigned char *chars, long long chars_capacity, void *stream);

/*
 * The encoder's offsets | logits projection with the sampler's front-end fused into the epilogue (N = 384 = 8 heads x (4 levels
 * x 4 points x 2 | 4 x 4)): out[row, 0:256] = reference_points[row % period, level, xy] + (a W^T + bias2d)[.] / (W_level | H_level),
 * out[row, 256:384] = softmax over each head's 16 logits — MSDeformAttn.forward between its two Linears and the sampling op
 * (projects/IDOL/idol/models/ops/modules/ms_deform_attn.py:98-108).  reference_points [period, 4, 2] fp32, level_wh = HOST
 * pointer to (W_0, H_0, ..., W_3, H_3) as floats; bias2d [period, 384] (row-periodic, carries the projection bias); out fp32.
 */
int vnext_linear_msda_front(int in_dtype, const void *a, int lda, const void *weight, const float *bias2d, int bias_period,
                            int ldb2, const float *reference_points, const float *level_wh, float *out, int ldo, int M, int K,
                            void *stream);

/*
 * 3x3 convolution, stride 1, zero padding 1, NHWC fp32, as an implicit GEMM on the tcgen05 tensor cores (TF32) with
 * the folded-FrozenBN bias and ReLU fused: out[T,H,W,Cout] = act(conv3x3(x[T,H,W,Cin]) + bias).  Replaces the cuDNN
 * conv + BN-affine + ReLU passes of BottleneckBlock.conv2 (detectron2/modeling/backbone/resnet.py:150-161, 194-199)
 * and of MaskHeadSmallConv (projects/SeqFormer/seqformer/models/segmentation_condInst.py:539-575).
 * weight_khwc: the conv weight re-laid at load time as [Cout, 3, 3, Cin] (K index = (ky*3 + kx)*Cin + c).
 * Needs Cin % 32 == 0, Cout % 4 == 0.  No im2col buffer: every filter tap is a shifted 4-D TMA box of the input.
 */
int vnext_conv3x3_tf32(const float *x, const float *weight_khwc, const float *bias, float *out, int T, int H, int W,
                       int Cin, int Cout, int relu, void *stream);
/* dtype-generic form (see vnext_linear_tc): bf16 input needs Cin % 64 == 0, bf16 output Cout % 8 == 0. */
int vnext_conv3x3_tc(int in_dtype, int out_dtype, const void *x, const void *weight_khwc, const float *bias, void *out,
                     int T, int H, int W, int Cin, int Cout, int relu, void *stream);
/* ... with a stride: 1 or 2 (the stride-in-3x3 convs that open res3 / res4 / res5, detectron2/modeling/backbone/
 * resnet.py:150-161 with STRIDE_IN_1X1 False, and the extra-level input projection, projects/SeqFormer/seqformer/models/
 * deformable_detr.py:70-76).  H, W are the INPUT size; out is [T, (H-1)/stride+1, (W-1)/stride+1, Cout].  The strided
 * taps are still single TMA boxes: the tensor map walks the input with element strides (stride, stride). */
int vnext_conv3x3_strided_tc(int in_dtype, int out_dtype, const void *x, const void *weight_khwc, const float *bias,
                             void *out, int T, int H, int W, int Cin, int Cout, int stride, int relu, void *stream);

/*
 * ResNet stem on the tensor cores: normalise + zero-pad the frames (as vnext_preprocess) and convolve 7x7 / stride 2 / padding 3
 * (3 -> Cout channels; BasicStem.conv1, detectron2/modeling/backbone/resnet.py:355-359) as an implicit GEMM — no im2col buffer:
 * the frames are written once as a zero-bordered 4-channel image into `scratch` ([T, Hp+6, Wp+6, 4] fp32) and each filter row is
 * one TMA box of an overlapping tensor-map view of it.  weight_packed [Cout, 7, 8, 4] fp32 = the conv weight as (ky, kx, c) with
 * kx padded 7 -> 8 and c padded 3 -> 4 by zeros.  out [T, Hp/2, Wp/2, Cout] (fp32 or bf16) = conv (+ bias, ReLU if asked).
 */
int vnext_stem_conv7x7(int in_is_u8, int out_dtype, const void *frames, const float *weight_packed, const float *bias, void *out,
                       float *scratch, int T, int H, int W, int Hp, int Wp, int Cout, const float *mean3, const float *std3, int relu,
                       void *stream);

/*
 * Multi-head self-attention over a few hundred queries (the decoder's object queries), head_dim 32:
 *   out[b, i, h*32 + c] = sum_j softmax_j(q_i . k_j / sqrt(32)) v_j[c]   per (batch b, head h)
 * with q | k | v read in place from the merged in-projection qkv [batch, n, >= 3*heads*32] (row pitch ld floats: q at
 * column 0, k at heads*32, v at 2*heads*32).  Replaces what nn.MultiheadAttention runs after its in-projection
 * (projects/SeqFormer/seqformer/models/deformable_transformer.py:280-298; scale, head split, two bmm, softmax, merge).
 * n <= 384.  fp32.
 */
int vnext_mha_small(const float *qkv, int ld, float *out, int ldo, int batch, int n, int heads, void *stream);

/*
 * Small fused pieces of the SeqFormer decoder layer (fp32, 300 queries x T frames):
 *   vnext_time_pool_ln  out[q,:] = LayerNorm(out_in[q,:] + sum_t softmax_t(logit[t,q]) * t2[t,q,:]) — temporal pooling of the
 *                       per-frame samples into the clip-level query (projects/SeqFormer/seqformer/models/
 *                       deformable_transformer.py:314-317).  t2 [T,nq,C], logit [T,nq], out_in / out [nq,C]; T <= 64, C <= 256.
 *   vnext_box_refine    out[r,0:4] = sigmoid(delta[r,:] + pad(inverse_sigmoid(ref[r,0:ref_dim]))) — iterative box refinement
 *                       (deformable_transformer.py:366-377; inverse_sigmoid eps 1e-5); ref_dim 2 or 4; rows = T * nq.
 */
int vnext_time_pool_ln(const float *out_in, const float *t2, const float *logit, const float *gamma, const float *beta,
                       float *out, int T, int nq, int C, float eps, void *stream);
int vnext_box_refine(const float *delta, const float *ref, float *out, int rows, int ref_dim, void *stream);
/* out[r,0:2] = sigmoid(inverse_sigmoid(ref[r,0:2])) * (scale_w, scale_h): reference points in pixels for the dynamic mask head
 * (projects/SeqFormer/seqformer/models/segmentation_condInst.py:311-331). */
int vnext_ref_points_px(const float *ref, float *out, int rows, int ref_dim, float scale_w, float scale_h, void *stream);

/*
 * Single-pass NHWC epilogues of the backbone / input projection (fp32, C % 4 == 0, 16-byte aligned):
 *   vnext_bias_act      out = act(x + bias[c] (+ residual))      — folded FrozenBN bias + ReLU (+ shortcut) after a
 *                       library conv (detectron2/modeling/backbone/resnet.py:194-210; layers/batch_norm.py:44-65)
 *   vnext_stem_pool     out = maxpool3x3/s2/p1(relu(x + bias))    — BasicStem tail (resnet.py:355-359); x [N,H,W,C]
 *   vnext_groupnorm_nhwc GroupNorm(groups, C == 8*groups) over x [N, HW, C], written to `out` with batch stride
 *                       `out_batch_stride` elements (lets the caller place a level directly inside the [T, S, C]
 *                       token tensor); stats_ws: vnext_groupnorm_workspace_floats(N, HW, groups) floats of scratch
 *                       (deformable_detr.py:64-78).  Deterministic: per-CTA partial sums folded in a fixed order.
 *   vnext_preprocess    frames [T,3,H,W] (uint8 if in_is_u8 else fp32, 0..255) -> (x - mean)/std, zero padded to
 *                       [T,Hp,Wp,3] channels-last (seqformer.py:413-429, util/misc.py:288-316); mean3/std3: HOST.
 */
int vnext_bias_act(const float *x, const float *bias, const float *residual, float *out, long long rows, int C,
                   int relu, void *stream);
int vnext_stem_pool(const float *x, const float *bias, float *out, int N, int H, int W, int C, void *stream);
size_t vnext_groupnorm_workspace_floats(int N, int HW, int groups);
int vnext_groupnorm_nhwc(const float *x, const float *gamma, const float *beta, float *out,
                         long long out_batch_stride, float *stats_ws, int N, int HW, int C, int groups, float eps,
                         void *stream);
int vnext_preprocess(int in_is_u8, const void *frames, float *out, int T, int H, int W, int Hp, int Wp,
                     const float *mean3, const float *std3, void *stream);

/* dtype-generic forms of the four kernels above (VNEXT_F32 | VNEXT_BF16 storage; bias / gamma / beta / statistics stay
 * fp32): the bf16 activation path of the backbone and input projection. */
int vnext_bias_act_dt(int dtype, const void *x, const float *bias, const void *residual, void *out, long long rows, int C,
                      int relu, void *stream);
int vnext_stem_pool_dt(int in_dtype, int out_dtype, const void *x, const float *bias, void *out, int N, int H, int W,
                       int C, void *stream);
int vnext_groupnorm_nhwc_dt(int in_dtype, int out_dtype, const void *x, const float *gamma, const float *beta, void *out,
                            long long out_batch_stride, float *stats_ws, int N, int HW, int C, int groups, float eps,
                            void *stream);
int vnext_preprocess_dt(int in_is_u8, int out_dtype, const void *frames, void *out, int T, int H, int W, int Hp, int Wp,
                        const float *mean3, const float *std3, void *stream);

/*
 * Swin window attention, one kernel per block (projects/IDOL/idol/backbone/swin.py:129-169 WindowAttention.forward +
 * the pad / roll / window_partition / window_reverse / roll / crop around it, :252-285):
 *   qkv   fp32 [B, H, W, 3C]   the block's qkv projection of the UN-padded, UN-rolled norm1(x) (q | k | v, heads x 32 each)
 *   qkv_bias [3C]              q/k/v of the zero-padded positions (the reference pads after norm1, so they equal the bias)
 *   rel_pos_bias_table [(2*window-1)^2, heads]
 *   out   fp32 [B, H, W, C]    softmax(q*scale @ k^T + bias + shift mask) @ v, heads concatenated — the input of `proj`
 * shift = 0 (W-MSA) or window/2 (SW-MSA; the -100 region mask of BasicLayer.forward :411-438 is generated in-kernel).
 * Requires C == heads*32 and window <= 12.
 *   vnext_layernorm: row LayerNorm over C (norm1/norm2, patch-embed / patch-merging / output norms), C % 4 == 0.
 */
int vnext_swin_window_attention(const float *qkv, const float *qkv_bias, const float *rel_pos_bias_table, float *out,
                                int B, int H, int W, int C, int heads, int window, int shift, void *stream);
/* Same contract on the tcgen05 tensor cores (TF32 MMAs, S / P / O in TMEM, per-thread row softmax); window in [4, 12]. */
int vnext_swin_window_attention_tc(const float *qkv, const float *qkv_bias, const float *rel_pos_bias_table, float *out,
                                   int B, int H, int W, int C, int heads, int window, int shift, void *stream);
int vnext_layernorm(const float *x, const float *gamma, const float *beta, float *out, long long rows, int C,
                    float eps, void *stream);
/* dtype-generic forms (VNEXT_F32 / VNEXT_BF16 storage; TF32 MMAs, fp32 statistics): the bf16 mode of the Swin backbone keeps
 * its residual stream in fp32 and hands bf16 tensors to / takes them from the bf16 GEMMs. */
int vnext_swin_window_attention_tc_dt(int io_dtype, const void *qkv, const float *qkv_bias, const float *rel_pos_bias_table,
                                      void *out, int B, int H, int W, int C, int heads, int window, int shift, void *stream);
/* bf16 storage, register-resident (flash-attention style) kernel: a warp owns 16 query rows, S / P / O never leave the register
 * file, both products are mma.sync m16n8k16 bf16 MMAs (csrc/swin_attention_mma.cu says why this beats the TMEM round trip at
 * 144 x 144 x 32).  qkv / out bf16; window in [2, 12].
 * NB: takes the relative-position-bias table TRANSPOSED, [heads, (2*window-1)^2] (one coalesced row per head). */
int vnext_swin_window_attention_bf16(const void *qkv, const float *qkv_bias, const float *rel_pos_bias_table_t, void *out, int B,
                                     int H, int W, int C, int heads, int window, int shift, void *stream);
int vnext_layernorm_dt(int in_dtype, int out_dtype, const void *x, const float *gamma, const float *beta, void *out,
                       long long rows, int C, float eps, void *stream);

/*
 * Fused AdamW over a flat fp32 segment (training step; torch.optim.AdamW semantics as built by
 * projects/IDOL/train_net.py:94-140, gradient clipping :110-125 folded into *grad_scale):
 *   g' = g * (*grad_scale);  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   w = w (1 - lr wd) - lr / bias_correction1 * m / (sqrt(v) / sqrt(bias_correction2) + eps)
 * n % 4 == 0, 16-byte aligned device buffers; grad_scale is a DEVICE pointer to one float (NULL = 1).
 */
int vnext_adamw(float *w, const float *g, float *m, float *v, long long n, float lr, float beta1, float beta2,
                float eps, float weight_decay, float bias_correction1, float bias_correction2, const float *grad_scale,
                void *stream);

/*
 * CondInst dynamic mask head + video-level mask post-processing.
 *   vnext_dynamic_mask_head: logits8[n, T, H, W] = MLP_i([ref_i - loc(x, y) | feats[t, y, x, 0:8]]) with the 10->8->8->1
 *     per-instance MLP generated by the controller (169 parameters: w0[8x10] w1[8x8] w2[8] b0[8] b1[8] b2, i.e.
 *     parse_dynamic_params' order, segmentation_condInst.py:614-637).  mask_feats [T, H, W, 8] NHWC; params [n, 169]
 *     shared by all frames (SeqFormer) or [T, n, 169] when params_per_frame != 0 (IDOL); ref_px [T, n, 2] = (x, y) in
 *     pixels; loc(x, y) = (x, y) * feat_stride + feat_stride / 2 (compute_locations :665-678).  Replaces
 *     dynamic_mask_with_coords / mask_heads_forward (:404-493) up to, not including, aligned_bilinear.
 *   vnext_mask_finalize: boolean masks out[n_images, OH, OW] (uint8 0/1) = nearest_resize(crop_{H,W}(sigmoid(
 *     bilinear_x{mask_stride}(aligned_bilinear_x2(logits8))))) > 0.5 — seqformer.py:381-396 + segmentation_condInst.py:640-662.
 */
int vnext_dynamic_mask_head(const float *mask_feats, const float *params, int params_per_frame, const float *ref_px,
                            float *out, int T, int n_inst, int H, int W, int feat_stride, void *stream);
int vnext_mask_finalize(const float *logits8, unsigned char *out, int n_images, int H8, int W8, int H, int W, int OH,
                        int OW, int mask_stride, void *stream);

/* Which kernel variant the last forward / backward on this thread dispatched to
 * ("fwd_f32_d32_tma", "fwd_generic", ...).  For tests and profiles. */
const char *vnext_msda_last_variant(void);

#ifdef __cplusplus
}
#endif
#endif /* VNEXT_B200_H_ */
