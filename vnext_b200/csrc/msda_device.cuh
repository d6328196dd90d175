// Device-side building blocks shared by the MSDA forward and backward kernels (sm_100a).
//
// Semantics restated from the reference kernel (projects/IDOL/idol/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:33-84, 272-296): pixel coordinates h_im = y*H - 0.5, w_im = x*W - 0.5
// (two separately rounded fp operations — no FMA contraction, so floor() lands on the same integer
// as the reference), sample kept iff h_im > -1 && w_im > -1 && h_im < H && w_im < W, each of the
// four bilinear corners kept iff it lies inside [0,H-1] x [0,W-1].
#pragma once
#include "common.cuh"

namespace vnext {

constexpr int kMaxLevels = 16;

struct LevelMeta {
    int H, W, start, pad;
};

// One bilinear sample, prepared by ONE lane: four corner offsets (in units of `unit` elements from
// the start of `value`, -1 = corner does not contribute) and the four bilinear weights.
struct SamplePrep {
    int off[4];
    float w[4];  // hh*hw, hh*lw, lh*hw, lh*lw  (0 for skipped corners / samples)
    float hh, hw, lh, lw;
    bool inside;
};

// pix_stride = elements (in `unit`s) between horizontally adjacent pixels = M * D / unit_elems
// base       = offset (in units) of (batch n, level start, head m, channel 0)
__device__ __forceinline__ SamplePrep prep_sample(float x, float y, int H, int W, int base, int pix_stride) {
    SamplePrep s;
    const float h_im = __fsub_rn(__fmul_rn(y, (float)H), 0.5f);
    const float w_im = __fsub_rn(__fmul_rn(x, (float)W), 0.5f);
    s.inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = (int)hf, w_low = (int)wf;
    s.lh = h_im - hf;
    s.lw = w_im - wf;
    s.hh = 1.f - s.lh;
    s.hw = 1.f - s.lw;
    const bool h0 = s.inside && h_low >= 0, h1 = s.inside && h_low + 1 <= H - 1;
    const bool w0 = w_low >= 0, w1 = w_low + 1 <= W - 1;
    const int o00 = base + (h_low * W + w_low) * pix_stride;
    s.off[0] = (h0 && w0) ? o00 : -1;
    s.off[1] = (h0 && w1) ? o00 + pix_stride : -1;
    s.off[2] = (h1 && w0) ? o00 + W * pix_stride : -1;
    s.off[3] = (h1 && w1) ? o00 + (W + 1) * pix_stride : -1;
    // skipped samples get zero weights so that a non-finite location cannot poison the sum
    s.w[0] = s.inside ? s.hh * s.hw : 0.f;
    s.w[1] = s.inside ? s.hh * s.lw : 0.f;
    s.w[2] = s.inside ? s.lh * s.hw : 0.f;
    s.w[3] = s.inside ? s.lh * s.lw : 0.f;
    return s;
}

__device__ __forceinline__ void load_level_meta(LevelMeta *meta, const int64_t *shapes, const int64_t *lsi, int L) {
    if (threadIdx.x < L) {
        // int64 device tensors truncated to int, as the reference does (cuh:274-277)
        meta[threadIdx.x].H = (int)shapes[2 * threadIdx.x];
        meta[threadIdx.x].W = (int)shapes[2 * threadIdx.x + 1];
        meta[threadIdx.x].start = (int)lsi[threadIdx.x];
        meta[threadIdx.x].pad = 0;
    }
}

}  // namespace vnext
