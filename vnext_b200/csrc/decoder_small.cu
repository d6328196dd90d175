// Two small fused kernels of the SeqFormer decoder layer (300 queries x T frames), replacing ~14 library launches each
// layer that are pure launch latency at this size:
//   * time_pool_ln:  out' = LayerNorm(out + sum_t softmax_t(logit[t, q]) * t2[t, q, :])    — the temporal pooling of the
//     per-frame box-stream samples into the clip-level mask stream
//     (projects/SeqFormer/seqformer/models/deformable_transformer.py:314-317: time_attention_weights, softmax over frames,
//     weighted sum, dropout1 (identity in eval), norm1);
//   * box_refine:    ref' = sigmoid(delta + inverse_sigmoid(ref))  (4-d references) or
//                    ref' = sigmoid(cat(delta[:2] + inverse_sigmoid(ref), delta[2:]))  (2-d references, first layer)
//     (deformable_transformer.py:366-377, util/misc.py inverse_sigmoid with eps 1e-5).
#include "common.cuh"

namespace vnext {

constexpr int kTpMaxT = 64;

__global__ void __launch_bounds__(256) time_pool_ln_kernel(const float *__restrict__ out_in, const float *__restrict__ t2,
                                                           const float *__restrict__ logit, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, float *__restrict__ out, int T, int nq,
                                                           int C, float eps) {
    const int q = blockIdx.x, c = threadIdx.x;
    __shared__ float red[2][8];
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, logit[t * nq + q]);
    float den = 0.f, acc = 0.f;
    for (int t = 0; t < T; ++t) {
        const float e = expf(logit[t * nq + q] - mx);
        den += e;
        if (c < C) acc = fmaf(e, t2[((size_t)t * nq + q) * C + c], acc);
    }
    const float v = c < C ? out_in[(size_t)q * C + c] + acc / den : 0.f;
    // LayerNorm over the C channels of this query (two-pass: mean, then centred second moment, as at::layer_norm)
    float s = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mean += red[0][w];
    mean /= (float)C;
    const float d = c < C ? v - mean : 0.f;
    float s2 = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s2;
    __syncthreads();
    float var = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) var += red[1][w];
    var /= (float)C;
    if (c < C) out[(size_t)q * C + c] = d * rsqrtf(var + eps) * gamma[c] + beta[c];
}

__device__ __forceinline__ float inverse_sigmoid_f(float x) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
    return logf(x1 / x2);
}

__global__ void box_refine_kernel(const float *__restrict__ delta, const float *__restrict__ ref, float *__restrict__ out,
                                  int rows, int ref_dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 4) return;
    const int r = i >> 2, k = i & 3;
    float x = delta[i];
    if (k < ref_dim) x += inverse_sigmoid_f(ref[r * ref_dim + k]);
    out[i] = 1.f / (1.f + expf(-x));
}

// ref_px[r, 0:2] = sigmoid(inverse_sigmoid(ref[r, 0:2])) * (scale_w, scale_h): the reference points of the last decoder level
// in pixels, as the mask head wants them (segmentation_condInst.py:311-331: inverse_sigmoid -> sigmoid -> scale)
__global__ void ref_points_px_kernel(const float *__restrict__ ref, float *__restrict__ out, int rows, int ref_dim, float sw,
                                     float sh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 2) return;
    const int r = i >> 1, k = i & 1;
    const float x = inverse_sigmoid_f(ref[r * ref_dim + k]);
    out[i] = (1.f / (1.f + expf(-x))) * (k ? sh : sw);
}

}  // namespace vnext

using namespace vnext;

extern "C" int vnext_time_pool_ln(const float *out_in, const float *t2, const float *logit, const float *gamma,
                                  const float *beta, float *out, int T, int nq, int C, float eps, void *stream) {
    VNEXT_CHECK_ARG(T > 0 && T <= kTpMaxT && nq >= 0 && C > 0 && C <= 256, "vnext_time_pool_ln: needs 0 < T <= %d, C <= 256 (T %d, C %d)",
                    kTpMaxT, T, C);
    if (nq == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(out_in && t2 && logit && gamma && beta && out, "vnext_time_pool_ln: null pointer");
    time_pool_ln_kernel<<<nq, 256, 0, static_cast<cudaStream_t>(stream)>>>(out_in, t2, logit, gamma, beta, out, T, nq, C, eps);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

extern "C" int vnext_box_refine(const float *delta, const float *ref, float *out, int rows, int ref_dim, void *stream) {
    VNEXT_CHECK_ARG(rows >= 0 && (ref_dim == 2 || ref_dim == 4), "vnext_box_refine: ref_dim must be 2 or 4");
    if (rows == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(delta && ref && out, "vnext_box_refine: null pointer");
    box_refine_kernel<<<(rows * 4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(delta, ref, out, rows, ref_dim);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

extern "C" int vnext_ref_points_px(const float *ref, float *out, int rows, int ref_dim, float scale_w, float scale_h, void *stream) {
    VNEXT_CHECK_ARG(rows >= 0 && (ref_dim == 2 || ref_dim == 4), "vnext_ref_points_px: ref_dim must be 2 or 4");
    if (rows == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(ref && out, "vnext_ref_points_px: null pointer");
    ref_points_px_kernel<<<(rows * 2 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ref, out, rows, ref_dim, scale_w, scale_h);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
