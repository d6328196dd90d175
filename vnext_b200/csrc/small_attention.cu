// Multi-head self-attention over a few hundred queries (the decoder's 300 object queries), head_dim 32 — one kernel
// for softmax(q k^T / sqrt(d)) v, reading q | k | v straight out of the merged in-projection [B, n, 3C].
//
// Replaces, per decoder layer and stream, the chain nn.MultiheadAttention runs after its in-projection
// (projects/SeqFormer/seqformer/models/deformable_transformer.py:280-298, torch/nn/functional.py multi_head_attention_forward):
// scale, three reshapes/transposes, two batched matmuls, a softmax and the transpose back — eight library launches on
// [heads, 300, 300] tensors that are pure launch latency at this size.  CUDA cores on purpose: 0.09 GFLOP per call.
//   * grid (B * heads, ceil(n / 32)): a CTA stages K^T and V of one (batch, head) in shared memory (K transposed so that
//     lanes walking consecutive keys hit consecutive banks) and serves 32 queries, 4 per warp;
//   * scores: lane j owns keys j, j+32, ..., 4 queries register-blocked; max / sum by warp shuffles; the normalised
//     probabilities are parked in the shared memory K^T occupied;
//   * P V: thread = (query, 4 channels).
#include <cmath>

#include "common.cuh"

namespace vnext {

// CTA = (batch, head, 32 queries), 8 warps.  Shared memory: K^T [32][n_pad] (re-used for the probabilities once the scores are
// in registers), V [n][32], the CTA's 32 scaled query rows: 83 KB at n = 300 -> two CTAs per SM.
constexpr int kSaWarps = 8, kSaQPerWarp = 4, kSaQPerCta = kSaWarps * kSaQPerWarp, kSaD = 32, kSaMaxKeysPerLane = 12;

__device__ __forceinline__ void sa_ffma2(float2 &acc, float w, float2 v) {   // acc += w * v, one packed FFMA2 (sm_100)
    unsigned long long d;
    const unsigned long long a = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
    const unsigned long long b = ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
    const unsigned long long c = ((unsigned long long)__float_as_uint(acc.y) << 32) | __float_as_uint(acc.x);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    acc.x = __uint_as_float((uint32_t)d);
    acc.y = __uint_as_float((uint32_t)(d >> 32));
}

__global__ void __launch_bounds__(kSaWarps * 32, 2) small_mha_kernel(const float *__restrict__ qkv, int ld, float *__restrict__ out,
                                                                     int ldo, int n, int heads, int n_pad, float scale, int vec) {
    extern __shared__ __align__(16) float sa_smem[];
    float *Kt = sa_smem;                                 // [32][n_pad]; after the score phase: P [32 queries][n_pad]
    float *V = Kt + kSaD * n_pad;                        // [n][32]
    float *Q = V + (size_t)n * kSaD;                     // [32 queries][32]
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int C = heads * kSaD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *base = qkv + (size_t)b * n * ld + h * kSaD;
    if (vec) {                                           // 16-byte loads; the odd pitch keeps the transposed stores conflict-free
#pragma unroll 4
        for (int i = tid; i < n * (kSaD / 4); i += blockDim.x) {
            const int j = i >> 3, d4 = (i & 7) * 4;
            const float4 k4 = *reinterpret_cast<const float4 *>(base + (size_t)j * ld + C + d4);
            const float4 v4 = *reinterpret_cast<const float4 *>(base + (size_t)j * ld + 2 * C + d4);
            Kt[(d4 + 0) * n_pad + j] = k4.x; Kt[(d4 + 1) * n_pad + j] = k4.y;
            Kt[(d4 + 2) * n_pad + j] = k4.z; Kt[(d4 + 3) * n_pad + j] = k4.w;
            *reinterpret_cast<float4 *>(V + j * kSaD + d4) = v4;
        }
    } else {
        for (int i = tid; i < n * kSaD; i += blockDim.x) {
            const int j = i >> 5, d = i & 31;
            Kt[d * n_pad + j] = base[(size_t)j * ld + C + d];
            V[i] = base[(size_t)j * ld + 2 * C + d];
        }
    }
    for (int i = tid; i < kSaD * (n_pad - n); i += blockDim.x)      // zero the padding columns of K^T (read below)
        Kt[(i / (n_pad - n)) * n_pad + n + i % (n_pad - n)] = 0.f;
    const int qc0 = blockIdx.y * kSaQPerCta;
    for (int i = tid; i < kSaQPerCta * kSaD; i += blockDim.x) {
        const int q = qc0 + (i >> 5);
        Q[i] = q < n ? base[(size_t)q * ld + (i & 31)] * scale : 0.f;
    }
    __syncthreads();
    // ---- scores: this warp's 4 queries x all keys; lane j owns keys j, j+32, ... (register-blocked: every K value read from
    //      shared memory feeds 4 FMAs)
    const int kpl = (n + 31) >> 5;
    float s[kSaQPerWarp][kSaMaxKeysPerLane];
    {
        float2 s2[kSaQPerWarp / 2][kSaMaxKeysPerLane];   // (query 2p, query 2p+1) pairs: one packed FFMA2 per key and pair
#pragma unroll
        for (int t = 0; t < kSaQPerWarp / 2; ++t)
#pragma unroll
            for (int r = 0; r < kSaMaxKeysPerLane; ++r) s2[t][r] = make_float2(0.f, 0.f);
        const float *Qw = Q + warp * kSaQPerWarp * kSaD;
#pragma unroll 4
        for (int d = 0; d < kSaD; ++d) {
            float2 qd[kSaQPerWarp / 2];
#pragma unroll
            for (int t = 0; t < kSaQPerWarp / 2; ++t) qd[t] = make_float2(Qw[(2 * t) * kSaD + d], Qw[(2 * t + 1) * kSaD + d]);
            const float *kr = Kt + d * n_pad + lane;
#pragma unroll
            for (int r = 0; r < kSaMaxKeysPerLane; ++r)
                if (r < kpl) {
                    const float kv = kr[r * 32];
#pragma unroll
                    for (int t = 0; t < kSaQPerWarp / 2; ++t) sa_ffma2(s2[t][r], kv, qd[t]);
                }
        }
#pragma unroll
        for (int t = 0; t < kSaQPerWarp / 2; ++t)
#pragma unroll
            for (int r = 0; r < kSaMaxKeysPerLane; ++r) {
                s[2 * t][r] = s2[t][r].x;
                s[2 * t + 1][r] = s2[t][r].y;
            }
    }
#pragma unroll
    for (int t = 0; t < kSaQPerWarp; ++t) {              // softmax over the keys, normalised in registers
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl && r * 32 + lane < n) mx = fmaxf(mx, s[t][r]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl) {
                s[t][r] = (r * 32 + lane < n) ? expf(s[t][r] - mx) : 0.f;
                sum += s[t][r];
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl) s[t][r] *= inv;
    }
    __syncthreads();                                     // every warp is done with K^T: its storage becomes P
    float *P = Kt;
#pragma unroll
    for (int t = 0; t < kSaQPerWarp; ++t)
#pragma unroll
        for (int r = 0; r < kSaMaxKeysPerLane; ++r)
            if (r < kpl) P[(warp * kSaQPerWarp + t) * n_pad + r * 32 + lane] = s[t][r];
    __syncthreads();
    // ---- P V: thread -> (query tid / 8, channels 4 * (tid % 8) ..): one broadcast P read + one 16-byte V read per 4 FMAs
    const int ql = tid >> 3, c4 = (tid & 7) * 4;
    const float *Pq = P + ql * n_pad;
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        const float pj = Pq[j];
        const float4 v = *reinterpret_cast<const float4 *>(V + j * kSaD + c4);
        sa_ffma2(a01, pj, make_float2(v.x, v.y));
        sa_ffma2(a23, pj, make_float2(v.z, v.w));
    }
    const int q = qc0 + ql;
    if (q < n) *reinterpret_cast<float4 *>(out + ((size_t)b * n + q) * ldo + h * kSaD + c4) = make_float4(a01.x, a01.y, a23.x, a23.y);
}

}  // namespace vnext

using namespace vnext;

extern "C" int vnext_mha_small(const float *qkv, int ld, float *out, int ldo, int batch, int n, int heads, void *stream) {
    VNEXT_CHECK_ARG(batch >= 0 && n >= 0 && heads > 0 && ld >= 3 * heads * kSaD && ldo >= heads * kSaD && ldo % 4 == 0,
                    "vnext_mha_small: bad sizes (batch %d, n %d, heads %d, ld %d, ldo %d); head_dim is 32", batch, n, heads, ld, ldo);
    if (batch == 0 || n == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(qkv && out && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "vnext_mha_small: null or unaligned pointer");
    VNEXT_CHECK_ARG(n <= 32 * kSaMaxKeysPerLane, "vnext_mha_small: at most %d queries (got %d)", 32 * kSaMaxKeysPerLane, n);
    const int n_pad = ((n + 31) / 32) * 32 + 1;          // odd pitch, >= keys-per-lane * 32
    const size_t smem = sizeof(float) * ((size_t)kSaD * n_pad + (size_t)n * kSaD + (size_t)kSaQPerCta * kSaD);
    static size_t configured = 0;
    if (smem > configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(small_mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(batch * heads, (n + kSaQPerCta - 1) / kSaQPerCta);
    small_mha_kernel<<<grid, kSaWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(qkv, ld, out, ldo, n, heads, n_pad,
                                                                                      (float)(1.0 / sqrt((double)kSaD)),
                                                                                      (ld % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15u) == 0) ? 1 : 0);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
