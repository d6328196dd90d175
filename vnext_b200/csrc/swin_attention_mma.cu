// Swin window attention with bf16 storage — register-resident (flash-attention style) kernel for sm_100a.
//
// Same contract as swin_window_attention_tc_kernel (swin_attention_tc.cu): one (frame, window, head) per work item; cyclic
// shift, window partition, zero-padded tokens (= the projection bias), relative-position bias and the SW-MSA region mask by
// index arithmetic (projects/IDOL/idol/backbone/swin.py:129-169, 233-293, 411-438); qkv [B, H, W, 3C] and out [B, H, W, C]
// in bf16.
//
// Why not tcgen05 here: the problem per item is 144 x 144 x 32.  With UMMA the 144 query rows are two M = 128 blocks (the
// second 87 % padding), S / P make a round trip through TMEM (tcgen05.ld / st of 160 columns per thread, twice), and each
// of the four MMAs is followed by a commit + mbarrier wait: ~25 us of serial latency per item at 4 warps per CTA, tensor
// pipe 6 % (profiles/r01_swin_attn_ncu_full.txt; 8.7 ms of the 27 ms Swin-L clip).  Here a warp owns 16 query rows and the
// whole attention row lives in its registers:
//   * CTA = ceil(N / 16) warps (9 for a 12 x 12 window), persistent over the work items; q, k, v of an item staged row-major
//     with an 80-B pitch (16-B aligned for cp.async, conflict-free for the fragment loads and ldmatrix) in one of TWO
//     stages: the cp.async gather of item i+1 flies during the products of item i;
//   * S = q k^T by mma.sync.m16n8k16 (bf16 x bf16 -> fp32): 18 key tiles x 2 k-steps, accumulators in registers;
//     scale, relative-position bias and region mask are applied to the fragments in place (per-token codes in shared
//     memory: no integer division in the loop); row max / sum by two quad shuffles each;
//   * the probabilities never leave the register file: the S fragment layout IS the A-fragment layout of the second
//     product (two adjacent 8-key tiles = one 16-key k-step), so O = P v is 9 k-steps x 4 dim tiles of the same MMA, its
//     B fragments (k = key, n = dim) fetched from the row-major V by ldmatrix.trans;
//   * 1 / sum folded into the epilogue that scatters O to the un-rolled token positions.
#include <mma.h>

#include "common.cuh"

namespace vnext {

constexpr int kMmMaxN = 144;                  // 12 x 12 window
constexpr int kMmMaxWarps = kMmMaxN / 16;     // 9
constexpr int kMmPitch = 40;                  // bf16 elements per staged q / k / v row (32 + 8: 80 B, 16-B aligned, conflict-free)
constexpr int kMmMaxTiles = kMmMaxN / 8;      // 18 key tiles of 8
constexpr int kMmTileElems = kMmMaxN * kMmPitch;

struct WinAttnMmaParams {
    const __nv_bfloat16 *qkv;
    const float *qkv_bias, *table_t;          // table_t: [heads, (2ws-1)^2] (the reference table transposed)
    __nv_bfloat16 *out;
    int B, H, W, Hp, Wp, C, heads, ws, shift, items;
    float scale;
};
struct WinAttnMmaStage {                      // everything one work item needs, double-buffered
    __nv_bfloat16 q[kMmTileElems], k[kMmTileElems], v[kMmTileElems];
    float bias[23 * 23 + 3];
    int tok[kMmMaxN];
    int code[kMmMaxN];                        // (bias-table offset of the token as a key) << 4 | shift region
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// four transposed 8 x 8 bf16 matrices: B fragments (k = key, n = dim) of two adjacent dim tiles from row-major V
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_row)));
}

__global__ void __launch_bounds__(kMmMaxWarps * 32, 2) swin_window_attention_mma_kernel(const WinAttnMmaParams p) {
    extern __shared__ __align__(16) unsigned char mm_smem[];
    WinAttnMmaStage *stage = reinterpret_cast<WinAttnMmaStage *>(mm_smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int g = lane >> 2, t4 = lane & 3;
    const int ws = p.ws, N = ws * ws;
    const int NK = (N + 15) & ~15;             // keys padded to whole k-steps of the second product
    const int ntiles = NK / 8;
    const int nWx = p.Wp / ws, nWy = p.Hp / ws;
    const int nb = (2 * ws - 1) * (2 * ws - 1);

    // token map + codes of `item` into stage st, then its q / k / v rows by cp.async (the caller commits / waits)
    auto prepare = [&](int item, WinAttnMmaStage &st) {
        int bid = item;
        const int h = bid % p.heads; bid /= p.heads;
        const int wx = bid % nWx; bid /= nWx;
        const int wy = bid % nWy;
        const int b = bid / nWy;
        for (int i = tid; i < NK * 4; i += nthreads) {          // 4 threads per token: each recomputes the token's map
            const int t = i >> 2, c8 = (i & 3) * 8;
            int tok = -2;                                         // key padding (beyond the window)
            if (t < N) {
                const int ty = t / ws, tx = t - ty * ws;
                const int ys = wy * ws + ty, xs = wx * ws + tx;
                int y = ys + p.shift, x = xs + p.shift;
                if (y >= p.Hp) y -= p.Hp;
                if (x >= p.Wp) x -= p.Wp;
                tok = (y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) : -1;
                if (c8 == 0) {
                    const int hr = ys < p.Hp - ws ? 0 : (ys < p.Hp - p.shift ? 1 : 2);
                    const int wr = xs < p.Wp - ws ? 0 : (xs < p.Wp - p.shift ? 1 : 2);
                    st.tok[t] = tok;
                    st.code[t] = ((ty * (2 * ws - 1) + tx) << 4) | (p.shift > 0 ? hr * 3 + wr : 0);
                }
            } else if (c8 == 0) {
                st.tok[t] = -2;
                st.code[t] = 0;
            }
            __nv_bfloat16 *qd = st.q + t * kMmPitch + c8, *kd = st.k + t * kMmPitch + c8, *vd = st.v + t * kMmPitch + c8;
            if (tok >= 0) {
                const __nv_bfloat16 *r = p.qkv + (size_t)tok * 3 * p.C + h * 32 + c8;
                cp_async16(qd, r);
                cp_async16(kd, r + p.C);
                cp_async16(vd, r + 2 * p.C);
            } else if (tok == -1) {            // zero-padded token: q / k / v are the projection bias
                const float *r = p.qkv_bias + h * 32 + c8;
                const float4 q0 = load4(r), q1 = load4(r + 4), k0 = load4(r + p.C), k1 = load4(r + p.C + 4);
                const float4 v0 = load4(r + 2 * p.C), v1 = load4(r + 2 * p.C + 4);
                *reinterpret_cast<uint4 *>(qd) = make_uint4(pack2_bf16(q0.x, q0.y), pack2_bf16(q0.z, q0.w), pack2_bf16(q1.x, q1.y), pack2_bf16(q1.z, q1.w));
                *reinterpret_cast<uint4 *>(kd) = make_uint4(pack2_bf16(k0.x, k0.y), pack2_bf16(k0.z, k0.w), pack2_bf16(k1.x, k1.y), pack2_bf16(k1.z, k1.w));
                *reinterpret_cast<uint4 *>(vd) = make_uint4(pack2_bf16(v0.x, v0.y), pack2_bf16(v0.z, v0.w), pack2_bf16(v1.x, v1.y), pack2_bf16(v1.z, v1.w));
            } else {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4 *>(qd) = z; *reinterpret_cast<uint4 *>(kd) = z; *reinterpret_cast<uint4 *>(vd) = z;
            }
        }
        const float *tb = p.table_t + (size_t)h * nb;
        for (int i = tid; i < nb; i += nthreads) st.bias[i] = __ldg(tb + i);
    };

    int item = blockIdx.x;
    if (item < p.items) prepare(item, stage[0]);
    cp_async_commit();
    for (int it = 0; item < p.items; item += gridDim.x, ++it) {
        WinAttnMmaStage &st = stage[it & 1];
        cp_async_wait_all();
        __syncthreads();                       // this item's tiles are complete; the other stage is no longer read
        const int next = item + gridDim.x;
        if (next < p.items) prepare(next, stage[(it + 1) & 1]);
        cp_async_commit();
        if (warp * 16 >= N) continue;          // (never for 12 x 12; smaller windows use fewer warps' rows)
        const int h = item % p.heads;
        const uint32_t *sK32 = reinterpret_cast<const uint32_t *>(st.k);
        const uint32_t *sQ32 = reinterpret_cast<const uint32_t *>(st.q);

        // ---- S = q k^T for this warp's 16 rows
        const int r0 = warp * 16 + g, r1 = r0 + 8;               // this thread's two query rows
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            qa[ks][0] = sQ32[(r0 * kMmPitch + ks * 16 + 2 * t4) >> 1];
            qa[ks][1] = sQ32[(r1 * kMmPitch + ks * 16 + 2 * t4) >> 1];
            qa[ks][2] = sQ32[(r0 * kMmPitch + ks * 16 + 8 + 2 * t4) >> 1];
            qa[ks][3] = sQ32[(r1 * kMmPitch + ks * 16 + 8 + 2 * t4) >> 1];
        }
        float s[kMmMaxTiles][4];
#pragma unroll
        for (int j = 0; j < kMmMaxTiles; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            if (j < ntiles) {
                const int key = j * 8 + g;                        // B fragment: (k = 2 t4 .. +1 (+8), n = g)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    mma_bf16_16816(s[j], qa[ks], sK32[(key * kMmPitch + ks * 16 + 2 * t4) >> 1],
                                   sK32[(key * kMmPitch + ks * 16 + 8 + 2 * t4) >> 1]);
            }
        }
        // ---- scale + relative-position bias + region mask, row max
        const int c0 = st.code[min(r0, N - 1)], c1 = st.code[min(r1, N - 1)];
        const int off = (ws - 1) * (2 * ws - 1) + (ws - 1);
        const int ib0 = (c0 >> 4) + off, ib1 = (c1 >> 4) + off, ireg0 = c0 & 15, ireg1 = c1 & 15;
        bool masked = false;                   // does this window straddle shift regions at all?
        if (p.shift > 0) {
            int bid2 = item / p.heads;
            const int wx2 = bid2 % nWx, wy2 = (bid2 / nWx) % nWy;
            masked = wx2 == nWx - 1 || wy2 == nWy - 1;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < kMmMaxTiles; ++j) {
            if (j < ntiles) {
                const int2 jc2 = *reinterpret_cast<const int2 *>(&st.code[j * 8 + 2 * t4]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = j * 8 + 2 * t4 + e;
                    if (col < N) {
                        const int jc = e ? jc2.y : jc2.x;
                        const int kb = jc >> 4, kr = jc & 15;
                        float a = fmaf(s[j][e], p.scale, st.bias[ib0 - kb]);
                        float c = fmaf(s[j][2 + e], p.scale, st.bias[ib1 - kb]);
                        if (masked) {              // (uniform per item: only the last window row / column of a shifted block)
                            if (kr != ireg0) a += -100.0f;
                            if (kr != ireg1) c += -100.0f;
                        }
                        s[j][e] = a; s[j][2 + e] = c;
                        mx0 = fmaxf(mx0, a); mx1 = fmaxf(mx1, c);
                    } else {
                        s[j][e] = -INFINITY; s[j][2 + e] = -INFINITY;
                    }
                }
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
        // ---- O = P v: exponentials packed straight into the A fragments of the second product; V fragments by ldmatrix.trans
        float o[4][4];
#pragma unroll
        for (int d = 0; d < 4; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
        // ldmatrix row addresses: lanes 0-15 -> keys 0..15 of the k-step at dims [0,8), lanes 16-31 -> same keys at dims [8,16)
        const __nv_bfloat16 *vrow = st.v + (lane & 15) * kMmPitch + (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < kMmMaxTiles / 2; ++kk) {
            if (2 * kk < ntiles) {
                float e[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    e[4 * q + 0] = __expf(s[2 * kk + q][0] - mx0); e[4 * q + 1] = __expf(s[2 * kk + q][1] - mx0);
                    e[4 * q + 2] = __expf(s[2 * kk + q][2] - mx1); e[4 * q + 3] = __expf(s[2 * kk + q][3] - mx1);
                }
                sum0 += (e[0] + e[1]) + (e[4] + e[5]);
                sum1 += (e[2] + e[3]) + (e[6] + e[7]);
                uint32_t pa[4] = {pack2_bf16(e[0], e[1]), pack2_bf16(e[2], e[3]), pack2_bf16(e[4], e[5]), pack2_bf16(e[6], e[7])};
#pragma unroll
                for (int dp = 0; dp < 2; ++dp) {                  // two dim tiles (16 dims) per ldmatrix.x4
                    uint32_t vb[4];                               // {b0, b1} of dim tile 2dp, {b0, b1} of dim tile 2dp+1
                    ldmatrix_x4_trans(vb, vrow + kk * 16 * kMmPitch + dp * 16);
                    mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
                    mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
                }
            }
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
        const int tok0 = r0 < N ? st.tok[r0] : -1, tok1 = r1 < N ? st.tok[r1] : -1;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int col = h * 32 + d * 8 + 2 * t4;
            if (tok0 >= 0) *reinterpret_cast<uint32_t *>(p.out + (size_t)tok0 * p.C + col) = pack2_bf16(o[d][0] * inv0, o[d][1] * inv0);
            if (tok1 >= 0) *reinterpret_cast<uint32_t *>(p.out + (size_t)tok1 * p.C + col) = pack2_bf16(o[d][2] * inv1, o[d][3] * inv1);
        }
    }
    cp_async_wait_all();
}

}  // namespace vnext

using namespace vnext;

extern "C" int vnext_swin_window_attention_bf16(const void *qkv, const float *qkv_bias, const float *rel_pos_bias_table_t, void *out,
                                                int B, int H, int W, int C, int heads, int window, int shift, void *stream) {
    VNEXT_CHECK_ARG(B >= 0 && H > 0 && W > 0 && heads > 0 && C == heads * 32,
                    "vnext_swin_window_attention_bf16: needs C == heads * 32 (got C=%d heads=%d)", C, heads);
    VNEXT_CHECK_ARG(window >= 2 && window <= 12 && shift >= 0 && shift < window,
                    "vnext_swin_window_attention_bf16: window must be in [2, 12] and 0 <= shift < window (got %d, %d)", window, shift);
    if (B == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(qkv && qkv_bias && rel_pos_bias_table_t && out && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (reinterpret_cast<uintptr_t>(qkv_bias) & 15) == 0,
                    "vnext_swin_window_attention_bf16: null or unaligned pointer");
    WinAttnMmaParams p;
    p.qkv = static_cast<const __nv_bfloat16 *>(qkv); p.qkv_bias = qkv_bias; p.table_t = rel_pos_bias_table_t;
    p.out = static_cast<__nv_bfloat16 *>(out);
    p.B = B; p.H = H; p.W = W; p.C = C; p.heads = heads; p.ws = window; p.shift = shift;
    p.Hp = (H + window - 1) / window * window;
    p.Wp = (W + window - 1) / window * window;
    p.scale = 1.0f / sqrtf(32.0f);
    const long long items = (long long)B * (p.Hp / window) * (p.Wp / window) * heads;
    VNEXT_CHECK_ARG(items < (1ll << 31), "vnext_swin_window_attention_bf16: too many windows");
    p.items = (int)items;
    const int N = window * window;
    const int warps = (N + 15) / 16;
    const int grid = (int)(items < 2ll * num_sms() ? items : 2ll * num_sms());
    constexpr size_t smem = 2 * sizeof(WinAttnMmaStage);
    static bool configured = false;
    if (!configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(swin_window_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    swin_window_attention_mma_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
