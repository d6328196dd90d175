// Swin window attention on the 5th-generation tensor cores (sm_100a): tcgen05 TF32 MMAs, accumulators and the
// attention matrix in TMEM.  Same contract as swin_window_attention_kernel (swin_attention.cu) — one (frame, window,
// head) per work item, cyclic shift / window partition / zero-padded tokens / relative-position bias / shift mask by
// index arithmetic — but the two matrix products of WindowAttention.forward
// (projects/IDOL/idol/backbone/swin.py:141-165) run on the tensor cores:
//
//   S = (q * scale) k^T   : tcgen05.mma kind::tf32, A = q tile [128 x 32] and B = k tile [NK x 32] in shared memory
//                           (K-major, SWIZZLE_128B — written by the gathering threads in exactly the layout TMA would
//                           produce), D = S in TMEM [128 lanes x NK columns]
//   softmax               : TMEM lane == query row, so each of the 128 threads owns one full row: + bias + mask, max,
//                           exp, sum are thread-local (no shuffles); P = exp(S - max) is written back to TMEM in place
//   O = P v               : tcgen05.mma with A = P read straight from TMEM, B = v^T [32 x NK] in shared memory;
//                           the 1/sum normalisation is folded into the epilogue that scatters O to [B, H, W, C]
// N = window^2 tokens (144 for Swin-L) are handled as M-blocks of 128 rows; NK = N rounded up to 32 (padding rows /
// columns are zero and masked).  128 threads, 224 TMEM columns and ~75 KB of shared memory per CTA -> 2 CTAs per SM,
// persistent over the work items.
#include "common.cuh"

namespace vnext {

constexpr int kTcThreads = 128;
constexpr int kTcMaxN = 144;
constexpr int kTcNK = 160;                 // kTcMaxN rounded up to a multiple of 32
constexpr int kTcTmemCols = 256;           // S/P: 160, O block 0: 32, O block 1: 32

struct WinAttnTcParams {
    const void *qkv;               // IOT [B, H, W, 3C]
    const float *qkv_bias, *table;
    void *out;                     // IOT [B, H, W, C]
    int B, H, W, Hp, Wp, C, heads, ws, shift, items;
    float scale;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t tc_desc_sw128(uint32_t smem_addr) {   // K-major, 128-B rows, 8-row swizzle atoms
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {          // tf32 x tf32 -> f32, K-major A and B
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// shared-memory carve-up (dynamic, 1024-B aligned for the swizzle atoms)
constexpr int kOffQ = 0;                                   // [256 rows][128 B]
constexpr int kOffK = 256 * 128;                           // [kTcNK rows][128 B]
constexpr int kOffVt = kOffK + kTcNK * 128;                // [kTcNK / 32 chunks][32 rows][128 B]
constexpr int kOffBias = kOffVt + kTcNK * 128;             // 23*23 floats
constexpr int kOffTok = kOffBias + 532 * 4;                // kTcMaxN ints
constexpr int kOffReg = kOffTok + kTcMaxN * 4;
constexpr int kTcSmemBytes = kOffReg + kTcMaxN * 4 + 1024;

// IOT: storage type of qkv and out (fp32 or bf16); the MMAs are TF32 on fp32 shared-memory operands either way (the gather
// widens bf16 on the fly: half the global bytes in, half out).
template <typename IOT>
__global__ void __launch_bounds__(kTcThreads, 2) swin_window_attention_tc_kernel(const WinAttnTcParams p) {
    extern __shared__ unsigned char tc_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ws = p.ws, N = ws * ws;
    const int NK = (N + 31) & ~31;
    const int nblk = (N + 127) / 128;
    float *sbias = reinterpret_cast<float *>(sm + kOffBias);
    int *stok = reinterpret_cast<int *>(sm + kOffTok);
    int *sreg = reinterpret_cast<int *>(sm + kOffReg);

    // zero the operand tiles once: padding rows / columns are never written again
    for (int i = tid; i < kOffBias / 16; i += kTcThreads) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)),
                     "n"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's TMEM lane quarter
    const uint32_t idesc_s = tc_idesc(128, NK), idesc_o = tc_idesc(128, 32);
    const int nWx = p.Wp / ws, nWy = p.Hp / ws;
    uint32_t parity = 0;

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int bid = item;
        const int h = bid % p.heads; bid /= p.heads;
        const int wx = bid % nWx; bid /= nWx;
        const int wy = bid % nWy;
        const int b = bid / nWy;

        // ---- token map, shift regions, this head's bias column
        for (int t = tid; t < N; t += kTcThreads) {
            const int ys = wy * ws + t / ws, xs = wx * ws + t % ws;
            int y = ys + p.shift, x = xs + p.shift;
            if (y >= p.Hp) y -= p.Hp;
            if (x >= p.Wp) x -= p.Wp;
            stok[t] = (y < p.H && x < p.W) ? ((b * p.H + y) * p.W + x) : -1;
            const int hr = ys < p.Hp - ws ? 0 : (ys < p.Hp - p.shift ? 1 : 2);
            const int wr = xs < p.Wp - ws ? 0 : (xs < p.Wp - p.shift ? 1 : 2);
            // packed per-token code: bias-table offset of the token as a KEY ((t / ws) * (2ws-1) + t % ws) and its shift
            // region, so the softmax loop needs no integer division
            sreg[t] = (((t / ws) * (2 * ws - 1) + t % ws) << 4) | (p.shift > 0 ? hr * 3 + wr : 0);
        }
        for (int i = tid; i < (2 * ws - 1) * (2 * ws - 1); i += kTcThreads) sbias[i] = p.table[(size_t)i * p.heads + h];
        __syncthreads();
        // ---- gather q (pre-scaled), k, v^T into the UMMA operand layouts
        for (int i = tid; i < N * 8; i += kTcThreads) {
            const int t = i >> 3, cq = i & 7, c4 = cq * 4;
            const int tok = stok[t];
            float4 q, k, v;
            if (tok >= 0) {
                const IOT *r = reinterpret_cast<const IOT *>(p.qkv) + (size_t)tok * 3 * p.C + h * 32 + c4;
                q = load4(r); k = load4(r + p.C); v = load4(r + 2 * p.C);
            } else {        // zero-padded token: q / k / v are the projection bias
                const float *r = p.qkv_bias + h * 32 + c4;
                q = load4(r); k = load4(r + p.C); v = load4(r + 2 * p.C);
            }
            const int sw = (cq ^ (t & 7)) << 4;
            *reinterpret_cast<float4 *>(sm + kOffQ + t * 128 + sw) = make_float4(q.x * p.scale, q.y * p.scale, q.z * p.scale, q.w * p.scale);
            *reinterpret_cast<float4 *>(sm + kOffK + t * 128 + sw) = k;
            const int chunk = t >> 5, col = t & 31;
            unsigned char *vt = sm + kOffVt + chunk * 4096 + (col & 3) * 4;
            const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c4 + e;
                *reinterpret_cast<float *>(vt + c * 128 + (((col >> 2) ^ (c & 7)) << 4)) = ve[e];
            }
        }
        fence_proxy_async();
        __syncthreads();

        for (int mb = 0; mb < nblk; ++mb) {
            // ---- S = q k^T
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = tc_desc_sw128(tc_smem_u32(sm + kOffQ + mb * 128 * 128));
                const uint64_t db = tc_desc_sw128(tc_smem_u32(sm + kOffK));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_ss(tmem, da + 2 * k, db + 2 * k, idesc_s, k != 0);
                tc_commit(&mma_bar);
            }
            mbar_wait(&mma_bar, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            // ---- softmax of this thread's row (TMEM lane == row).  Warps whose 32 rows are all padding skip it.
            const int i = mb * 128 + tid;
            const bool warp_live = mb * 128 + warp * 32 < N;
            float sum = 1.f;
            if (warp_live) {
                const int ic = min(i, N - 1);
                const int icode = sreg[ic];
                const int ireg = icode & 15;
                // bias index = (iy - jy + ws-1)*(2ws-1) + (ix - jx + ws-1) = key_i - key_j + (ws-1)*(2ws-1) + (ws-1)
                const int ibase = (icode >> 4) + (ws - 1) * (2 * ws - 1) + (ws - 1);
                float mx = -INFINITY;
                for (int c0 = 0; c0 < NK; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(lane_addr + c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = c0 + j;
                        float s = -INFINITY;
                        if (col < N) {
                            const int jc = sreg[col];
                            s = __uint_as_float(v[j]) + sbias[ibase - (jc >> 4)];
                            if ((jc & 15) != ireg) s += -100.0f;
                        }
                        mx = fmaxf(mx, s);
                        v[j] = __float_as_uint(s);
                    }
                    tc_st16(lane_addr + c0, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                sum = 0.f;
                for (int c0 = 0; c0 < NK; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(lane_addr + c0, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float e = __expf(__uint_as_float(v[j]) - mx);    // exp(-inf) = 0 for the padding columns
                        sum += e;
                        v[j] = __float_as_uint(e);
                    }
                    tc_st16(lane_addr + c0, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();

            // ---- O = P v   (A = P from TMEM)
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + kTcNK + mb * 32;
                for (int kk = 0; kk < NK / 8; ++kk) {
                    const uint64_t db = tc_desc_sw128(tc_smem_u32(sm + kOffVt + (kk >> 2) * 4096)) + 2 * (kk & 3);
                    tc_mma_ts(d, tmem + kk * 8, db, idesc_o, kk != 0);
                }
                tc_commit(&mma_bar);
            }
            mbar_wait(&mma_bar, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            // ---- epilogue: O / sum -> out[token]
            if (warp_live) {
                uint32_t v[16];
                const float inv = 1.f / sum;
                const int tok = i < N ? stok[i] : -1;
                IOT *orow = reinterpret_cast<IOT *>(p.out) + (size_t)max(tok, 0) * p.C + h * 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tc_ld16(lane_addr + kTcNK + mb * 32 + half * 16, v);
                    if (tok >= 0) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            store4(orow + half * 16 + j,
                                   make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv,
                                               __uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();       // S/P columns and the operand tiles are free again
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols));
    }
}

}  // namespace vnext

using namespace vnext;

static int swin_tc_launch(int io_dtype, const void *qkv, const float *qkv_bias, const float *rel_pos_bias_table, void *out, int B,
                          int H, int W, int C, int heads, int window, int shift, void *stream) {
    VNEXT_CHECK_ARG(io_dtype == VNEXT_F32 || io_dtype == VNEXT_BF16, "vnext_swin_window_attention_tc: dtype must be fp32 (0) or bf16 (2)");
    VNEXT_CHECK_ARG(B >= 0 && H > 0 && W > 0 && heads > 0 && C == heads * 32,
                    "vnext_swin_window_attention_tc: needs C == heads * 32 (got C=%d heads=%d)", C, heads);
    VNEXT_CHECK_ARG(window >= 4 && window <= 12 && shift >= 0 && shift < window,
                    "vnext_swin_window_attention_tc: window must be in [4, 12] and 0 <= shift < window (got %d, %d)", window, shift);
    if (B == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(qkv && qkv_bias && rel_pos_bias_table && out && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_bias) & 15) == 0,
                    "vnext_swin_window_attention_tc: null or unaligned pointer");
    WinAttnTcParams p;
    p.qkv = qkv; p.qkv_bias = qkv_bias; p.table = rel_pos_bias_table; p.out = out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.heads = heads; p.ws = window; p.shift = shift;
    p.Hp = (H + window - 1) / window * window;
    p.Wp = (W + window - 1) / window * window;
    p.scale = 1.0f / sqrtf(32.0f);
    const long long items = (long long)B * (p.Hp / window) * (p.Wp / window) * heads;
    VNEXT_CHECK_ARG(items < (1ll << 31), "vnext_swin_window_attention_tc: too many windows");
    p.items = (int)items;
    static bool configured = false;
    if (!configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(swin_window_attention_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        VNEXT_CUDA(cudaFuncSetAttribute(swin_window_attention_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        configured = true;
    }
    const int grid = (int)(items < 2ll * num_sms() ? items : 2ll * num_sms());
    if (io_dtype == VNEXT_F32)
        swin_window_attention_tc_kernel<float><<<grid, kTcThreads, kTcSmemBytes, (cudaStream_t)stream>>>(p);
    else
        swin_window_attention_tc_kernel<__nv_bfloat16><<<grid, kTcThreads, kTcSmemBytes, (cudaStream_t)stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

extern "C" int vnext_swin_window_attention_tc(const float *qkv, const float *qkv_bias, const float *rel_pos_bias_table,
                                              float *out, int B, int H, int W, int C, int heads, int window, int shift,
                                              void *stream) {
    return swin_tc_launch(VNEXT_F32, qkv, qkv_bias, rel_pos_bias_table, out, B, H, W, C, heads, window, shift, stream);
}

extern "C" int vnext_swin_window_attention_tc_dt(int io_dtype, const void *qkv, const float *qkv_bias,
                                                 const float *rel_pos_bias_table, void *out, int B, int H, int W, int C,
                                                 int heads, int window, int shift, void *stream) {
    return swin_tc_launch(io_dtype, qkv, qkv_bias, rel_pos_bias_table, out, B, H, W, C, heads, window, shift, stream);
}
