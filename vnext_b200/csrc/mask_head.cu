// CondInst dynamic mask head and the video-level mask post-processing, fused (sm_100a, fp32).
//
// dynamic_mask_head_kernel replaces `dynamic_mask_with_coords` + `mask_heads_forward`
// (projects/SeqFormer/seqformer/models/segmentation_condInst.py:404-493): the reference materialises a
// [1, n_inst*10, H/8, W/8] tensor (relative coordinates concatenated with 8 mask-feature channels, repeated per
// instance) and runs three grouped 1x1 convolutions (groups = n_inst).  Here one thread evaluates the instance's
// 10->8->8->1 MLP for one stride-8 pixel straight from the shared 8-channel feature map; the instance's 169
// generated parameters sit in shared memory.
//
// mask_finalize_kernel replaces, for the instances whole_video_inference keeps (seqformer.py:376-398):
// aligned_bilinear x2 (segmentation_condInst.py:640-662) -> F.interpolate x4 bilinear(align_corners=False) ->
// sigmoid -> crop to the un-padded frame -> nearest resize to the output size -> (> 0.5), producing the boolean
// masks in one pass over the stride-8 logits (which stay L2-resident) with no intermediate tensors.
#include "common.cuh"

namespace vnext {

// sigmoid(v) > 0.5 (seqformer.py:388-396: sigmoid, then `> 0.5`) without the exponential: the two differ only for
// 0 < v < 6e-8, where fp32 sigmoid rounds to exactly 0.5
__device__ __forceinline__ bool mask_on(float v) { return v > 0.f; }


constexpr int kMhThreads = 256;

// feats [T, H, W, 8] (NHWC); params [n, 169] (pstride_t = 0) or [T, n, 169]; ref_px [T, n, 2] (x, y) in pixels
// out [n, T, H, W] stride-8 logits
__global__ void __launch_bounds__(kMhThreads) dynamic_mask_head_kernel(const float *__restrict__ feats,
                                                                      const float *__restrict__ params,
                                                                      const float *__restrict__ ref_px,
                                                                      float *__restrict__ out, int T, int n, int H, int W,
                                                                      long long pstride_t, int stride, int chunks) {
    __shared__ __align__(16) float w[172];
    int bid = blockIdx.x;
    const int chunk = bid % chunks; bid /= chunks;
    const int inst = bid % n;
    const int t = bid / n;
    const float *prm = params + t * pstride_t + (size_t)inst * 169;
    for (int i = threadIdx.x; i < 169; i += kMhThreads) w[i] = prm[i];
    __syncthreads();
    const float rx = ref_px[((size_t)t * n + inst) * 2], ry = ref_px[((size_t)t * n + inst) * 2 + 1];
    const float *w0 = w, *w1 = w + 80, *w2 = w + 144, *b0 = w + 152, *b1 = w + 160;
    const float b2 = w[168];
    const int HW = H * W;
    const int per = (HW + chunks - 1) / chunks;
    const int p_end = min(HW, (chunk + 1) * per);
    for (int pix = chunk * per + threadIdx.x; pix < p_end; pix += kMhThreads) {
        const int y = pix / W, x = pix % W;
        const float4 fa = *reinterpret_cast<const float4 *>(feats + ((size_t)t * HW + pix) * 8);
        const float4 fb = *reinterpret_cast<const float4 *>(feats + ((size_t)t * HW + pix) * 8 + 4);
        const float in[10] = {rx - (float)(x * stride + stride / 2), ry - (float)(y * stride + stride / 2),
                              fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
        float h1[8], h2[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a = b0[o];
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(w0[o * 10 + k], in[k], a);
            h1[o] = fmaxf(a, 0.f);
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float a = b1[o];
#pragma unroll
            for (int k = 0; k < 8; ++k) a = fmaf(w1[o * 8 + k], h1[k], a);
            h2[o] = fmaxf(a, 0.f);
        }
        float r = b2;
#pragma unroll
        for (int k = 0; k < 8; ++k) r = fmaf(w2[k], h2[k], r);
        out[((size_t)inst * T + t) * HW + pix] = r;
    }
}

// stride-4 logit (aligned_bilinear x2 of the stride-8 map p8 [H8, W8]) at integer position (i, j), 0 <= i < 2*H8
__device__ __forceinline__ float logit4(const float *__restrict__ p8, int H8, int W8, int i, int j) {
    const int u = max(i - 1, 0), v = max(j - 1, 0);              // replicate pad (1, 0, 1, 0) of the (2h+1, 2w+1) upsample
    const int y0 = u >> 1, x0 = v >> 1;
    const int y1 = min(y0 + (u & 1), H8 - 1), x1 = min(x0 + (v & 1), W8 - 1);   // replicate pad (0, 1, 0, 1) of p8
    const float a = p8[y0 * W8 + x0], b = p8[y0 * W8 + x1], c = p8[y1 * W8 + x0], d = p8[y1 * W8 + x1];
    const float wy = (u & 1) ? 0.5f : 0.f, wx = (v & 1) ? 0.5f : 0.f;           // align_corners=True: source = u / 2
    const float top = a + (b - a) * wx, bot = c + (d - c) * wx;
    return top + (bot - top) * wy;
}

// p8 [n, T, H8, W8] -> out [n, T, OH, OW] uint8 (0/1).  (H, W): un-padded frame size; (OH, OW): output size.
// One thread produces PIX consecutive output pixels of a row and stores them as one word.
template <int PIX>
__global__ void __launch_bounds__(256) mask_finalize_kernel(const float *__restrict__ p8, unsigned char *__restrict__ out,
                                                            int NT, int H8, int W8, int H, int W, int OH, int OW,
                                                            int mask_stride) {
    const int H4 = 2 * H8, W4 = 2 * W8;
    const int OWg = OW / PIX;                                     // host guarantees OW % PIX == 0
    const long long total = (long long)NT * OH * OWg;
    const float inv = 1.f / (float)mask_stride;
    const float ry = (float)H / (float)OH, rx = (float)W / (float)OW;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int Xg = (int)(e % OWg);
        const long long t2 = e / OWg;
        const int Y = (int)(t2 % OH);
        const int img = (int)(t2 / OH);
        const float *m = p8 + (size_t)img * H8 * W8;
        // nearest resize (F.interpolate mode='nearest'): src = floor(dst * in / out); then
        // bilinear x mask_stride, align_corners=False: src = (dst + 0.5) / s - 0.5, clamped at 0
        const int y = min((int)floorf((float)Y * ry), H - 1);
        const float sy = fmaxf(((float)y + 0.5f) * inv - 0.5f, 0.f);
        const int y0 = min((int)sy, H4 - 1), y1 = min(y0 + 1, H4 - 1);
        const float ly = sy - (float)y0;
        unsigned int word[PIX / 4];
#pragma unroll
        for (int k = 0; k < PIX / 4; ++k) word[k] = 0u;
#pragma unroll
        for (int k = 0; k < PIX; ++k) {
            const int X = Xg * PIX + k;
            const int x = min((int)floorf((float)X * rx), W - 1);
            const float sx = fmaxf(((float)x + 0.5f) * inv - 0.5f, 0.f);
            const int x0 = min((int)sx, W4 - 1), x1 = min(x0 + 1, W4 - 1);
            const float lx = sx - (float)x0;
            const float v00 = logit4(m, H8, W8, y0, x0), v01 = logit4(m, H8, W8, y0, x1);
            const float v10 = logit4(m, H8, W8, y1, x0), v11 = logit4(m, H8, W8, y1, x1);
            const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
            if (mask_on(v)) word[k / 4] |= 1u << (8 * (k % 4));
        }
        unsigned int *o = reinterpret_cast<unsigned int *>(out + ((size_t)(img * (long long)OH + Y) * OW + (size_t)Xg * PIX));
#pragma unroll
        for (int k = 0; k < PIX / 4; ++k) o[k] = word[k];
    }
}

// Fast path for the common case (no resize, mask_stride 4): 8 consecutive output pixels of a row only ever touch a
// 2 x 4 patch of stride-4 logits, so that patch is evaluated once (32 loads instead of 128) and the eight bilinear
// weights are compile-time constants: src_x = 2g + (k - 1.5) / 4.
__global__ void __launch_bounds__(256) mask_finalize_x4_kernel(const float *__restrict__ p8, unsigned char *__restrict__ out,
                                                               int NT, int H8, int W8, int H, int W) {
    const int H4 = 2 * H8, W4 = 2 * W8;
    const int Wg = W / 8;
    const long long total = (long long)NT * H * Wg;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(e % Wg);
        const long long t2 = e / Wg;
        const int y = (int)(t2 % H);
        const int img = (int)(t2 / H);
        const float *m = p8 + (size_t)img * H8 * W8;
        const float sy = fmaxf(((float)y + 0.5f) * 0.25f - 0.5f, 0.f);
        const int y0 = min((int)sy, H4 - 1), y1 = min(y0 + 1, H4 - 1);
        const float ly = sy - (float)y0;
        float col[4];                                            // vertically interpolated logits of columns 2g-1 .. 2g+2
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(max(2 * g - 1 + j, 0), W4 - 1);
            const float a = logit4(m, H8, W8, y0, c), b = logit4(m, H8, W8, y1, c);
            col[j] = a + (b - a) * ly;
        }
        unsigned int w0 = 0u, w1 = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr float kLx[8] = {0.625f, 0.875f, 0.125f, 0.375f, 0.625f, 0.875f, 0.125f, 0.375f};
            const int j0 = k < 2 ? 0 : (k < 6 ? 1 : 2);
            float lx = kLx[k];
            if (g == 0 && k < 2) lx = 0.f;                        // src_x clamped at 0 on the left edge
            const float v = col[j0] + (col[j0 + 1] - col[j0]) * lx;
            const unsigned int bit = mask_on(v) ? 1u : 0u;
            if (k < 4) w0 |= bit << (8 * k); else w1 |= bit << (8 * (k - 4));
        }
        *reinterpret_cast<uint2 *>(out + ((size_t)(img * (long long)H + y) * W + (size_t)g * 8)) = make_uint2(w0, w1);
    }
}

__global__ void __launch_bounds__(256) mask_finalize_scalar_kernel(const float *__restrict__ p8, unsigned char *__restrict__ out,
                                                                   int NT, int H8, int W8, int H, int W, int OH, int OW,
                                                                   int mask_stride) {
    const int H4 = 2 * H8, W4 = 2 * W8;
    const long long total = (long long)NT * OH * OW;
    const float inv = 1.f / (float)mask_stride;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(e % OW);
        const long long t2 = e / OW;
        const int Y = (int)(t2 % OH);
        const int img = (int)(t2 / OH);
        const int y = min((int)floorf((float)Y * ((float)H / (float)OH)), H - 1);
        const int x = min((int)floorf((float)X * ((float)W / (float)OW)), W - 1);
        const float sy = fmaxf(((float)y + 0.5f) * inv - 0.5f, 0.f), sx = fmaxf(((float)x + 0.5f) * inv - 0.5f, 0.f);
        const int y0 = min((int)sy, H4 - 1), x0 = min((int)sx, W4 - 1);
        const int y1 = min(y0 + 1, H4 - 1), x1 = min(x0 + 1, W4 - 1);
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float *m = p8 + (size_t)img * H8 * W8;
        const float v = (1.f - ly) * ((1.f - lx) * logit4(m, H8, W8, y0, x0) + lx * logit4(m, H8, W8, y0, x1)) +
                        ly * ((1.f - lx) * logit4(m, H8, W8, y1, x0) + lx * logit4(m, H8, W8, y1, x1));
        out[e] = mask_on(v) ? 1 : 0;
    }
}

}  // namespace vnext

using namespace vnext;

extern "C" {

int vnext_dynamic_mask_head(const float *mask_feats, const float *params, int params_per_frame, const float *ref_px,
                            float *out, int T, int n_inst, int H, int W, int feat_stride, void *stream) {
    VNEXT_CHECK_ARG(T >= 0 && n_inst >= 0 && H > 0 && W > 0, "vnext_dynamic_mask_head: bad shape");
    if (T == 0 || n_inst == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(mask_feats && params && ref_px && out && (reinterpret_cast<uintptr_t>(mask_feats) & 15) == 0,
                    "vnext_dynamic_mask_head: null or unaligned pointer");
    const int HW = H * W;
    int chunks = (HW + 4095) / 4096;
    while ((long long)T * n_inst * chunks < 4ll * num_sms() && chunks < (HW + 255) / 256) chunks *= 2;
    dynamic_mask_head_kernel<<<T * n_inst * chunks, kMhThreads, 0, (cudaStream_t)stream>>>(
        mask_feats, params, ref_px, out, T, n_inst, H, W, params_per_frame ? (long long)n_inst * 169 : 0, feat_stride, chunks);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

int vnext_mask_finalize(const float *logits8, unsigned char *out, int n_images, int H8, int W8, int H, int W, int OH, int OW,
                        int mask_stride, void *stream) {
    VNEXT_CHECK_ARG(n_images >= 0 && H8 > 0 && W8 > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && mask_stride > 0,
                    "vnext_mask_finalize: bad shape");
    VNEXT_CHECK_ARG(H <= 2 * H8 * mask_stride && W <= 2 * W8 * mask_stride, "vnext_mask_finalize: frame larger than the mask grid");
    if (n_images == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(logits8 && out, "vnext_mask_finalize: null pointer");
    const bool wide = (OW % 8 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    const long long total = (long long)n_images * OH * (wide ? OW / 8 : OW);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (wide && OH == H && OW == W && mask_stride == 4 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
        mask_finalize_x4_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits8, out, n_images, H8, W8, H, W);
    else if (wide)
        mask_finalize_kernel<8><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits8, out, n_images, H8, W8, H, W, OH, OW, mask_stride);
    else
        mask_finalize_scalar_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits8, out, n_images, H8, W8, H, W, OH, OW, mask_stride);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

}  // extern "C"
