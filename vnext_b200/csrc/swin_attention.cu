// Swin window attention (W-MSA / SW-MSA) as ONE kernel per block, plus a row LayerNorm (sm_100a, fp32).
//
// Replaces, per Swin block of the reference (projects/IDOL/idol/backbone/swin.py:233-293 and :129-169):
// F.pad -> torch.roll -> window_partition (permute + contiguous copy) -> q*scale -> q@k^T (materialised
// [nW*B, heads, N, N] in HBM) -> relative-position-bias gather + add -> shift-mask add -> softmax -> attn@v ->
// transpose/reshape -> window_reverse -> torch.roll -> crop.  Here one CTA owns one (frame, window, head): it gathers
// the window's q/k/v rows straight out of the un-padded, un-rolled qkv projection [B, H, W, 3C] (the cyclic shift and
// the window partition are index arithmetic; zero-padded tokens — the reference pads AFTER norm1, so their q/k/v are
// the projection's bias — are synthesised from qkv_bias), keeps q, k^T, v, the head's relative-position-bias column
// and the window's shift-region ids in shared memory, computes S = qk^T + bias + mask, the row softmax and P.V with
// 4-row register tiles per warp, and scatters the result rows back to [B, H, W, C] at their un-rolled positions.
// The N x N attention matrix never exists in HBM.  head_dim = 32 (every Swin variant), window <= 12 (N <= 144).
#include "common.cuh"

namespace vnext {

constexpr int kWinThreads = 256;
constexpr int kWinWarps = 8;
constexpr int kMaxN = 144;     // 12 x 12 window
constexpr int kRows = 4;       // query rows per warp pass

struct WinAttnParams {
    const float *qkv;        // [B, H, W, 3C]  (q | k | v, each heads x 32)
    const float *qkv_bias;   // [3C]
    const float *table;      // [(2ws-1)^2, heads]
    float *out;              // [B, H, W, C]
    int B, H, W, Hp, Wp, C, heads, ws, shift;
    float scale;
};

__global__ void __launch_bounds__(kWinThreads, 2) swin_window_attention_kernel(const WinAttnParams p) {
    extern __shared__ __align__(16) unsigned char win_smem[];      // > 48 KB: dynamic
    float (*sq)[32] = reinterpret_cast<float (*)[32]>(win_smem);
    float (*skT)[kMaxN] = reinterpret_cast<float (*)[kMaxN]>(win_smem + sizeof(float) * kMaxN * 32);
    float (*sv)[32] = reinterpret_cast<float (*)[32]>(win_smem + sizeof(float) * kMaxN * 64);
    float (*sp)[kRows][kMaxN] = reinterpret_cast<float (*)[kRows][kMaxN]>(win_smem + sizeof(float) * kMaxN * 96);
    float *sbias = reinterpret_cast<float *>(win_smem + sizeof(float) * (kMaxN * 96 + kWinWarps * kRows * kMaxN));
    int *stok = reinterpret_cast<int *>(sbias + 23 * 23 + 3);   // global token row, or -1 for a zero-padded position
    int *sreg = stok + kMaxN;                                   // shift-mask region id

    const int ws = p.ws, N = ws * ws;
    const int nWx = p.Wp / ws, nWy = p.Hp / ws;
    int bid = blockIdx.x;
    const int h = bid % p.heads; bid /= p.heads;
    const int wx = bid % nWx; bid /= nWx;
    const int wy = bid % nWy;
    const int b = bid / nWy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int t = tid; t < N; t += kWinThreads) {
        const int ys = wy * ws + t / ws, xs = wx * ws + t % ws;           // coordinates in the rolled, padded frame
        int y = ys + p.shift, x = xs + p.shift;                          // torch.roll(x, -shift): rolled[i] = x[i + shift]
        if (y >= p.Hp) y -= p.Hp;
        if (x >= p.Wp) x -= p.Wp;
        stok[t] = (y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) : -1;
        const int hr = ys < p.Hp - ws ? 0 : (ys < p.Hp - p.shift ? 1 : 2);
        const int wr = xs < p.Wp - ws ? 0 : (xs < p.Wp - p.shift ? 1 : 2);
        sreg[t] = p.shift > 0 ? hr * 3 + wr : 0;
    }
    const int nb = (2 * ws - 1) * (2 * ws - 1);
    for (int i = tid; i < nb; i += kWinThreads) sbias[i] = p.table[(size_t)i * p.heads + h];
    __syncthreads();
    // q (pre-scaled), k (transposed), v: one float4 per thread per token part
    for (int i = tid; i < N * 8; i += kWinThreads) {
        const int t = i >> 3, c4 = (i & 7) * 4;
        const int tok = stok[t];
        float4 q, k, v;
        if (tok >= 0) {
            const float *r = p.qkv + (size_t)tok * 3 * p.C + h * 32 + c4;
            q = *reinterpret_cast<const float4 *>(r);
            k = *reinterpret_cast<const float4 *>(r + p.C);
            v = *reinterpret_cast<const float4 *>(r + 2 * p.C);
        } else {
            const float *r = p.qkv_bias + h * 32 + c4;
            q = *reinterpret_cast<const float4 *>(r);
            k = *reinterpret_cast<const float4 *>(r + p.C);
            v = *reinterpret_cast<const float4 *>(r + 2 * p.C);
        }
        *reinterpret_cast<float4 *>(&sq[t][c4]) = make_float4(q.x * p.scale, q.y * p.scale, q.z * p.scale, q.w * p.scale);
        skT[c4][t] = k.x; skT[c4 + 1][t] = k.y; skT[c4 + 2][t] = k.z; skT[c4 + 3][t] = k.w;
        *reinterpret_cast<float4 *>(&sv[t][c4]) = v;
    }
    __syncthreads();

    const int N4 = (N + 3) & ~3;
    for (int i0 = warp * kRows; i0 < N; i0 += kWinWarps * kRows) {
        // ---- S = q k^T for rows i0..i0+3, columns j = lane + 32*it
        float s[kRows][5];
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int it = 0; it < 5; ++it) s[r][it] = 0.f;
#pragma unroll 2
        for (int c = 0; c < 32; c += 4) {
            float4 qr[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) qr[r] = *reinterpret_cast<const float4 *>(&sq[min(i0 + r, N - 1)][c]);
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int j = min(lane + 32 * it, N - 1);
                const float k0 = skT[c][j], k1 = skT[c + 1][j], k2 = skT[c + 2][j], k3 = skT[c + 3][j];
#pragma unroll
                for (int r = 0; r < kRows; ++r)
                    s[r][it] = fmaf(qr[r].x, k0, fmaf(qr[r].y, k1, fmaf(qr[r].z, k2, fmaf(qr[r].w, k3, s[r][it]))));
            }
        }
        // ---- + relative position bias + shift mask, softmax over the row
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int i = min(i0 + r, N - 1);
            const int iy = i / ws, ix = i % ws, ireg = sreg[i];
            float mx = -INFINITY;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int j = lane + 32 * it;
                if (j < N) {
                    const int jy = j / ws, jx = j % ws;
                    float val = s[r][it] + sbias[(iy - jy + ws - 1) * (2 * ws - 1) + (ix - jx + ws - 1)];
                    if (sreg[j] != ireg) val += -100.0f;
                    s[r][it] = val;
                    mx = fmaxf(mx, val);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int j = lane + 32 * it;
                const float e = j < N ? expf(s[r][it] - mx) : 0.f;
                s[r][it] = e;
                sum += e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int j = lane + 32 * it;
                if (j < N4) sp[warp][r][j] = j < N ? s[r][it] * inv : 0.f;
            }
        }
        __syncwarp();
        // ---- O = P V: lane = channel
        float o[kRows] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < N4; j += 4) {
            float v0 = sv[min(j, N - 1)][lane], v1 = sv[min(j + 1, N - 1)][lane];
            float v2 = sv[min(j + 2, N - 1)][lane], v3 = sv[min(j + 3, N - 1)][lane];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float4 pr = *reinterpret_cast<const float4 *>(&sp[warp][r][j]);
                o[r] = fmaf(pr.x, v0, fmaf(pr.y, v1, fmaf(pr.z, v2, fmaf(pr.w, v3, o[r]))));
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int i = i0 + r;
            if (i < N) {
                const int tok = stok[i];
                if (tok >= 0) p.out[(size_t)tok * p.C + h * 32 + lane] = o[r];
            }
        }
        __syncwarp();
    }
}

// Row LayerNorm: one warp per row, C % 4 == 0, two passes over the row (C <= 32 * 4 * kLnVec).  IT / OT: fp32 or bf16
// storage of x / out (statistics and affine parameters fp32).
constexpr int kLnVec = 48;      // float4 per lane: C up to 6144
template <typename IT, typename OT>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const IT *__restrict__ x, const float4 *__restrict__ gamma,
                                                             const float4 *__restrict__ beta, OT *__restrict__ out,
                                                             long long rows, int c4, float eps) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = warp0; row < rows; row += nwarps) {
        const IT *xr = x + row * c4 * 4;
        float sum = 0.f, sumsq = 0.f;
        for (int i = lane; i < c4; i += 32) {
            const float4 v = load4(xr + i * 4);
            sum += (v.x + v.y) + (v.z + v.w);
            sumsq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        }
        const float inv_n = 1.f / (4.f * c4);
        const float mean = sum * inv_n;
        const float rstd = rsqrtf(fmaxf(sumsq * inv_n - mean * mean, 0.f) + eps);
        OT *orow = out + row * c4 * 4;
        for (int i = lane; i < c4; i += 32) {
            const float4 v = load4(xr + i * 4), g = __ldg(gamma + i), b = __ldg(beta + i);
            store4(orow + i * 4, make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                                             (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w));
        }
    }
}

}  // namespace vnext

using namespace vnext;

extern "C" {

int vnext_swin_window_attention(const float *qkv, const float *qkv_bias, const float *rel_pos_bias_table, float *out, int B,
                                int H, int W, int C, int heads, int window, int shift, void *stream) {
    VNEXT_CHECK_ARG(B >= 0 && H > 0 && W > 0 && heads > 0 && C == heads * 32,
                    "vnext_swin_window_attention: needs C == heads * 32 (got C=%d heads=%d)", C, heads);
    VNEXT_CHECK_ARG(window >= 2 && window <= 12 && shift >= 0 && shift < window,
                    "vnext_swin_window_attention: window must be in [2, 12] and 0 <= shift < window (got %d, %d)", window, shift);
    if (B == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(qkv && qkv_bias && rel_pos_bias_table && out, "vnext_swin_window_attention: null pointer");
    WinAttnParams p;
    p.qkv = qkv; p.qkv_bias = qkv_bias; p.table = rel_pos_bias_table; p.out = out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.heads = heads; p.ws = window; p.shift = shift;
    p.Hp = (H + window - 1) / window * window;
    p.Wp = (W + window - 1) / window * window;
    p.scale = 1.0f / sqrtf(32.0f);
    const long long grid = (long long)B * (p.Hp / window) * (p.Wp / window) * heads;
    VNEXT_CHECK_ARG(grid < (1ll << 31), "vnext_swin_window_attention: grid too large");
    constexpr size_t smem = sizeof(float) * (kMaxN * 96 + kWinWarps * kRows * kMaxN + 23 * 23 + 3) + sizeof(int) * 2 * kMaxN;
    static bool configured = false;
    if (!configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(swin_window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    swin_window_attention_kernel<<<(int)grid, kWinThreads, smem, (cudaStream_t)stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

int vnext_layernorm_dt(int in_dtype, int out_dtype, const void *x, const float *gamma, const float *beta, void *out, long long rows,
                       int C, float eps, void *stream) {
    VNEXT_CHECK_ARG((in_dtype == VNEXT_F32 || in_dtype == VNEXT_BF16) && (out_dtype == VNEXT_F32 || out_dtype == VNEXT_BF16),
                    "vnext_layernorm: dtypes must be fp32 (0) or bf16 (2)");
    VNEXT_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0, "vnext_layernorm: C must be a positive multiple of 4 (got %d)", C);
    if (rows == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(x && gamma && beta && out, "vnext_layernorm: null pointer");
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const float4 *g = (const float4 *)gamma, *b = (const float4 *)beta;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == VNEXT_F32 && out_dtype == VNEXT_F32)
        layernorm_rows_kernel<float, float><<<(int)blocks, 256, 0, st>>>((const float *)x, g, b, (float *)out, rows, C / 4, eps);
    else if (in_dtype == VNEXT_F32)
        layernorm_rows_kernel<float, __nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const float *)x, g, b, (__nv_bfloat16 *)out, rows, C / 4, eps);
    else if (out_dtype == VNEXT_F32)
        layernorm_rows_kernel<__nv_bfloat16, float><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16 *)x, g, b, (float *)out, rows, C / 4, eps);
    else
        layernorm_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16 *)x, g, b, (__nv_bfloat16 *)out, rows, C / 4, eps);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

int vnext_layernorm(const float *x, const float *gamma, const float *beta, float *out, long long rows, int C, float eps,
                    void *stream) {
    return vnext_layernorm_dt(VNEXT_F32, VNEXT_F32, x, gamma, beta, out, rows, C, eps, stream);
}

}  // extern "C"
