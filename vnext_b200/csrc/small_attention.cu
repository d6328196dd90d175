// Multi-head self-attention over a few hundred queries (the decoder's 300 object queries), head_dim 32 — one kernel
// for softmax(q k^T / sqrt(d)) v, reading q | k | v straight out of the merged in-projection [B, n, 3C].
//
// Replaces, per decoder layer and stream, the chain nn.MultiheadAttention runs after its in-projection
// (projects/SeqFormer/seqformer/models/deformable_transformer.py:280-298, torch/nn/functional.py multi_head_attention_forward):
// scale, three reshapes/transposes, two batched matmuls, a softmax and the transpose back — eight library launches on
// [heads, 300, 300] tensors that are pure launch latency at this size.  CUDA cores on purpose: 0.09 GFLOP per call.
//   * grid (B * heads, ceil(n / 64)): a CTA stages K^T and V of one (batch, head) in shared memory (K transposed so that
//     lanes walking consecutive keys hit consecutive banks) and serves 64 queries, 8 per warp;
//   * scores: lane j owns keys j, j+32, ...; max / sum by warp shuffles; probabilities parked in shared memory;
//   * P V: lane = output channel; the warp's 8 queries share every V row read.
#include <cmath>

#include "common.cuh"

namespace vnext {

constexpr int kSaWarps = 8, kSaQPerWarp = 8, kSaQPerCta = kSaWarps * kSaQPerWarp, kSaD = 32, kSaMaxKeysPerLane = 16;

__global__ void __launch_bounds__(kSaWarps * 32) small_mha_kernel(const float *__restrict__ qkv, int ld, float *__restrict__ out,
                                                                  int ldo, int n, int heads, int n_pad, float scale) {
    extern __shared__ __align__(16) float sa_smem[];
    float *Kt = sa_smem;                                 // [32][n_pad]
    float *V = Kt + kSaD * n_pad;                        // [n][32]
    float *P = V + (size_t)n * kSaD;                     // [warps][8][n_pad]
    float *Q = P + (size_t)kSaWarps * kSaQPerWarp * n_pad;   // [warps][8][32]
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int C = heads * kSaD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *base = qkv + (size_t)b * n * ld + h * kSaD;
    for (int i = tid; i < n * kSaD; i += blockDim.x) {   // coalesced over d
        const int j = i >> 5, d = i & 31;
        Kt[d * n_pad + j] = base[(size_t)j * ld + C + d];
        V[i] = base[(size_t)j * ld + 2 * C + d];
    }
    const int q0 = blockIdx.y * kSaQPerCta + warp * kSaQPerWarp;
#pragma unroll
    for (int t = 0; t < kSaQPerWarp; ++t) {
        const int q = q0 + t;
        Q[(warp * kSaQPerWarp + t) * kSaD + lane] = q < n ? base[(size_t)q * ld + lane] * scale : 0.f;
    }
    __syncthreads();
    if (q0 >= n) return;
    float *Pw = P + (size_t)warp * kSaQPerWarp * n_pad;
    const float *Qw = Q + warp * kSaQPerWarp * kSaD;
    const int kpl = (n + 31) >> 5;                       // keys per lane (<= kSaMaxKeysPerLane)
    float inv[kSaQPerWarp];
#pragma unroll 1
    for (int t = 0; t < kSaQPerWarp; ++t) {
        float s[kSaMaxKeysPerLane];
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r) s[r] = 0.f;
        for (int d = 0; d < kSaD; ++d) {
            const float qd = Qw[t * kSaD + d];
            const float *kr = Kt + d * n_pad + lane;
#pragma unroll
            for (int r = 0; r < kSaMaxKeysPerLane; ++r)
                if (r < kpl) s[r] = fmaf(qd, kr[r * 32], s[r]);           // (columns >= n of Kt are zero padding)
        }
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl && r * 32 + lane < n) mx = fmaxf(mx, s[r]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl) {
                const float e = (r * 32 + lane < n) ? expf(s[r] - mx) : 0.f;
                Pw[t * n_pad + r * 32 + lane] = e;
                sum += e;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        inv[t] = 1.f / sum;
    }
    __syncwarp();
    float acc[kSaQPerWarp];
#pragma unroll
    for (int t = 0; t < kSaQPerWarp; ++t) acc[t] = 0.f;
    for (int j = 0; j < n; ++j) {
        const float v = V[j * kSaD + lane];
#pragma unroll
        for (int t = 0; t < kSaQPerWarp; ++t) acc[t] = fmaf(Pw[t * n_pad + j], v, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < kSaQPerWarp; ++t) {
        const int q = q0 + t;
        if (q < n) out[((size_t)b * n + q) * ldo + h * kSaD + lane] = acc[t] * inv[t];
    }
}

}  // namespace vnext

using namespace vnext;

extern "C" int vnext_mha_small(const float *qkv, int ld, float *out, int ldo, int batch, int n, int heads, void *stream) {
    VNEXT_CHECK_ARG(batch >= 0 && n >= 0 && heads > 0 && ld >= 3 * heads * kSaD && ldo >= heads * kSaD,
                    "vnext_mha_small: bad sizes (batch %d, n %d, heads %d, ld %d, ldo %d); head_dim is 32", batch, n, heads, ld, ldo);
    if (batch == 0 || n == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(qkv && out, "vnext_mha_small: null pointer");
    VNEXT_CHECK_ARG(n <= 32 * kSaMaxKeysPerLane, "vnext_mha_small: at most %d queries (got %d)", 32 * kSaMaxKeysPerLane, n);
    const int n_pad = ((n + 31) / 32) * 32 + 1;          // odd pitch: the transposed-K stores of consecutive d do not collide
    const size_t smem = sizeof(float) * ((size_t)kSaD * n_pad + (size_t)n * kSaD + (size_t)kSaWarps * kSaQPerWarp * n_pad +
                                         (size_t)kSaWarps * kSaQPerWarp * kSaD);
    static size_t configured = 0;
    if (smem > configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(small_mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(batch * heads, (n + kSaQPerCta - 1) / kSaQPerCta);
    small_mha_kernel<<<grid, kSaWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(qkv, ld, out, ldo, n, heads, n_pad,
                                                                                      (float)(1.0 / sqrt((double)kSaD)));
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
