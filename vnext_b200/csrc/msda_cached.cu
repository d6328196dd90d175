// Multi-scale deformable attention forward for SELF-attention over the pixel grid (encoder shape, Lq == S) —
// cache-blocked kernel for sm_100a: the same gather as the tile kernel of msda_forward.cu, re-ordered so that the L1
// holds what neighbouring queries share.
//
// Why: the tile kernel walks (query, head) pairs in memory order, so the 64 corner rows of a pair (128 B each) are
// re-fetched from L2 by every neighbouring query: 5.0 GB of L2 -> L1 sectors per 720p launch (14x the algorithmic
// bytes) at the L2 throughput cap, L1 hit rate 16 % (profiles/r01_msda_enc_v3_ncu_full.txt).  The shared-memory window
// kernel (msda_window.cu) removes that traffic but pays for the window management with occupancy (16 warps/SM, one
// CTA-wide barrier per unit) and ends up slower (profiles/r02_msda_window_v3_*).  This kernel keeps the tile kernel's
// occupancy (32 warps/SM) and lets the hardware cache do the staging:
//   * ONE persistent CTA of 32 warps per SM owns that SM's L1 (static shared memory is only the 32-KB sample table, the
//     carve-out is set to maximise L1);
//   * work unit = (batch n, th x tw query tile of ONE query level, head m), tw a power of two, sized on the host so that
//     the unit's footprint in all sampled levels (tile scaled to the level + a reach of a few pixels, one head = one
//     cache line per pixel) fits the L1.  All 32 warps work on the same one or two units at any time (a unit's passes are
//     dealt round-robin to the warps, no barrier between units), so a corner line is fetched from L2 about once per unit
//     instead of once per sample;
//   * per 32-sample pass: phase 1 (one lane per sample) as in the tile kernel, but the sample is parked as ONE 32-byte
//     entry: two 32-bit row offsets, the horizontal step to the right-hand pixel and the four corner weights.  Corners
//     the reference skips (outside [0,H-1]x[0,W-1], ms_deform_im2col_cuda.cuh:55-78) get weight 0 and the address of the
//     sample's OTHER, valid row / column (dx = 0), samples the reference skips entirely get weight 0 and a clamped
//     in-level address: every load is in bounds, there are no predicates, and nothing is read that is not level data;
//   * phase 2: LPC lanes x 16 B = one pixel row of 32 channels; two LDS.128 + four LDG.128 (L1 hits) + packed FFMA2 per
//     sample and lane; transposing shuffle reduction, one coalesced row store per pair;
//   * the sampling inputs of a warp's NEXT pass are loaded (L1::no_allocate: they are streamed once) before the gather of
//     the current one.
// Exact for any sampling locations (the L1 is only a cache).  Needs the level shapes on the host (tile plan).
#include <cuda.h>

#include <cstdlib>

#include "msda_device.cuh"

namespace vnext {

constexpr int kCachedMaxL = 4;
constexpr int kCachedWarps = 32;
constexpr int kCachedThreads = kCachedWarps * 32;

struct CDiv {          // exact n / d for 0 <= n < 2^31 (Granlund-Montgomery, 32-bit magic)
    uint32_t mul, shr;
};
static CDiv make_cdiv(uint32_t d) {
    CDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    f.shr = 31 + l;
    f.mul = (uint32_t)(((1ull << f.shr) + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t cdiv(uint32_t n, const CDiv &f) { return (uint32_t)(((uint64_t)n * f.mul) >> f.shr); }

struct CachedClass {   // all units of one query level
    int H, W, start;
    int th, tw_shift;
    int np_shift;      // log2(passes per unit) = log2(th * tw * LK / 32)
    int tiles_x, tiles_y;
    int units;         // N * tiles_y * tiles_x * M
    CDiv d_tx, d_ty;
};
struct CachedParams {
    CachedClass cls[kCachedMaxL];
    LevelMeta meta[kCachedMaxL];
    const void *value;
    const float *loc, *attn;                 // classic front-end
    const float *oa;                         // fused front-end (see msda_forward.cu)
    int oa_pitch, n_off;
    const float *ref;
    long long ref_batch_stride;
    void *out;
    int N, S, M, L, K, k_shift;
    CDiv d_M;
};

__device__ __forceinline__ float2 ldg_stream_f2(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_cached_128(const unsigned char *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ void cached_ffma2(float2 &acc, float w, uint32_t lo, uint32_t hi) {   // acc += w * (lo, hi)
    unsigned long long d;
    const unsigned long long a = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
    const unsigned long long b = ((unsigned long long)hi << 32) | lo;
    const unsigned long long c = ((unsigned long long)__float_as_uint(acc.y) << 32) | __float_as_uint(acc.x);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    acc.x = __uint_as_float((uint32_t)d);
    acc.y = __uint_as_float((uint32_t)(d >> 32));
}

template <typename VT>
struct CachedRow;
template <>
struct CachedRow<float> {
    static constexpr int VEC = 4, LPC = 8, ROWB = 128;
    __device__ static __forceinline__ void store(void *out, long long idx, float v) { reinterpret_cast<float *>(out)[idx] = v; }
};
template <>
struct CachedRow<__nv_bfloat16> {
    static constexpr int VEC = 8, LPC = 4, ROWB = 64;
    __device__ static __forceinline__ void store(void *out, long long idx, float v) {
        reinterpret_cast<__nv_bfloat16 *>(out)[idx] = __float2bfloat16_rn(v);
    }
};

// Fold VEC partial sums held by each of NG = VEC lane groups into one value per lane (see msda_forward.cu):
// on return lane (grp = lane / LPC, sub = lane % LPC) holds channel VEC*sub + grp.
template <int VEC>
__device__ __forceinline__ float cached_transpose_reduce(float (&a)[VEC], int lane) {
    int mask = 16;
#pragma unroll
    for (int n = VEC; n > 1; n >>= 1, mask >>= 1) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? a[n / 2 + i] : a[i];
            const float send = up ? a[i] : a[n / 2 + i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    return a[0];
}

template <bool FUSED>
struct CachedIn;           // one sample's raw front-end inputs
template <>
struct CachedIn<false> {
    float a0, a1, a2;      // x, y, attention weight
};
template <>
struct CachedIn<true> {
    float a0, a1, a2, r0, r1;   // offset x, offset y, logit, reference point
};
struct CachedPass {        // where this lane's sample of a pass lives
    int row;               // n * S + q of the lane's query (undefined when !live)
    int n, m;
    bool live;
};

// FRONT: 0 = classic (sampling_loc / attn_weight tensors); 1 = fused (raw offsets | logits rows + reference points: location
// arithmetic and softmax here); 2 = the same row layout, but already holding locations | normalised weights (written by the
// projection GEMM's epilogue, vnext_linear_msda_front)
template <typename VT, int LK, int FRONT>
__global__ void __launch_bounds__(kCachedThreads, 1) msda_fwd_cached_kernel(const __grid_constant__ CachedParams p) {
    constexpr bool FUSED = FRONT == 1;
    using RW = CachedRow<VT>;
    constexpr int VEC = RW::VEC, LPC = RW::LPC, NG = 32 / LPC, ROWB = RW::ROWB;
    constexpr int PPP = 32 / LK;          // (query, head) pairs per 32-sample warp pass
    constexpr int LKS = LK == 8 ? 3 : (LK == 16 ? 4 : 5);
    static_assert(LK % NG == 0 && (LK == 8 || LK == 16 || LK == 32), "unsupported L*K for the cached kernel");

    __shared__ __align__(16) uint4 table[kCachedWarps][32][2];   // per sample: {off0, off1, dx, w00} {w01, w10, w11, -}
    __shared__ __align__(16) float4 lvl_f[kCachedMaxL];          // (W, H) as float, -, -
    __shared__ __align__(16) int4 lvl_i[kCachedMaxL];            // W, H, start, -

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane / LPC, sub = lane % LPC;
    if (tid < p.L) {
        lvl_f[tid] = make_float4((float)p.meta[tid].W, (float)p.meta[tid].H, 0.f, 0.f);
        lvl_i[tid] = make_int4(p.meta[tid].W, p.meta[tid].H, p.meta[tid].start, 0);
    }
    __syncthreads();

    const unsigned char *vlane = reinterpret_cast<const unsigned char *>(p.value) + sub * 16;
    const uint32_t tab = smem_u32(&table[warp][0][0]);
    const int smp = lane & (LK - 1), pl = lane >> LKS;
    const int lvl = smp >> p.k_shift;
    const uint32_t pix_b = (uint32_t)(p.M * ROWB);               // bytes between horizontally adjacent pixels

    for (int c = 0; c < p.L; ++c) {
        const CachedClass &C = p.cls[c];
        const int units_here = C.units > (int)blockIdx.x ? (C.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        const int total = units_here << C.np_shift;              // this CTA's passes in this class
        const int qH = C.H, qW = C.W, qstart = C.start, tws = C.tw_shift, nps = C.np_shift;

        auto locate = [&](int i) {                               // pass i of this CTA's stream -> this lane's query
            CachedPass s;
            const int k = i >> nps, pass = i & ((1 << nps) - 1);
            const uint32_t u = blockIdx.x + (uint32_t)k * gridDim.x;
            uint32_t t = cdiv(u, p.d_M);
            s.m = (int)(u - t * p.M);
            const uint32_t t2 = cdiv(t, C.d_tx);
            const int tx = (int)(t - t2 * C.tiles_x);
            const uint32_t n = cdiv(t2, C.d_ty);
            const int ty = (int)(t2 - n * C.tiles_y);
            const int qi = pass * PPP + pl;
            const int qy = ty * C.th + (qi >> tws), qx = (tx << tws) + (qi & ((1 << tws) - 1));
            s.live = qy < qH && qx < qW;
            s.n = (int)n;
            s.row = s.n * p.S + qstart + qy * qW + qx;
            return s;
        };
        auto fetch = [&](const CachedPass &s) {
            CachedIn<FUSED> in;
            in.a0 = in.a1 = in.a2 = 0.f;
            if constexpr (FUSED) in.r0 = in.r1 = 0.f;
            if (s.live) {
                if constexpr (FRONT != 0) {
                    const float *rowp = p.oa + (size_t)s.row * p.oa_pitch;
                    const float2 off = ldg_stream_f2(rowp + (s.m * LK + smp) * 2);
                    in.a0 = off.x; in.a1 = off.y;
                    in.a2 = ldg_stream_f1(rowp + p.n_off + s.m * LK + smp);
                }
                if constexpr (FUSED) {
                    // (no L1 allocation either: the reference grid is 0.6 MB per frame set and would evict corner rows)
                    const float2 rf = ldg_stream_f2(p.ref + s.n * p.ref_batch_stride + ((size_t)(s.row - s.n * p.S) * p.L + lvl) * 2);
                    in.r0 = rf.x; in.r1 = rf.y;
                }
                if constexpr (FRONT == 0) {
                    const size_t si = ((size_t)s.row * p.M + s.m) * LK + smp;
                    const float2 xy = ldg_stream_f2(p.loc + 2 * si);
                    in.a0 = xy.x; in.a1 = xy.y;
                    in.a2 = ldg_stream_f1(p.attn + si);
                }
            }
            return in;
        };

        int i = warp;
        if (i >= total) continue;
        CachedPass cur = locate(i);
        CachedIn<FUSED> in = fetch(cur);
#pragma unroll 1
        while (true) {
            const int inext = i + kCachedWarps;
            const bool more = inext < total;
            CachedPass nxt = cur;
            CachedIn<FUSED> in_n = in;
            if (more) {
                nxt = locate(inext);
                in_n = fetch(nxt);
            }
            // ---------------------------------------------------------------- phase 1: one sample per lane
            {
                const float4 lf = lvl_f[lvl];
                const int4 li = lvl_i[lvl];
                const float fW = lf.x, fH = lf.y;
                const int W = li.x, H = li.y;
                float x, y, a = in.a2;
                if constexpr (FUSED) {
                    x = in.r0 + in.a0 / fW;                       // ref + offset / (W_l, H_l)   (module :102-105)
                    y = in.r1 + in.a1 / fH;
                    float mx = cur.live ? a : -INFINITY;
#pragma unroll
                    for (int o = LK / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    // fast exp / divide (2 ulp): the softmax weights enter a 1e-4-tolerance sum, not an index
                    const float e = cur.live ? __expf(a - mx) : 0.f;
                    float sum = e;
#pragma unroll
                    for (int o = LK / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    a = cur.live ? __fdividef(e, sum) : 0.f;
                } else {
                    x = in.a0;
                    y = in.a1;
                }
                const float h_im = __fsub_rn(__fmul_rn(y, fH), 0.5f);
                const float w_im = __fsub_rn(__fmul_rn(x, fW), 0.5f);
                const bool inside = cur.live && (h_im > -1.f) && (w_im > -1.f) && (h_im < fH) && (w_im < fW);
                const float hf = floorf(h_im), wf = floorf(w_im);
                // skipped samples (out of range, non-finite, dead lanes): weight 0 at pixel (0, 0) of the level
                const int h_low = inside ? (int)hf : 0, w_low = inside ? (int)wf : 0;
                const float lh = inside ? h_im - hf : 0.f, lw = inside ? w_im - wf : 0.f;
                const float av = inside ? a : 0.f;
                const bool h0 = h_low >= 0, h1 = h_low + 1 <= H - 1, w0 = w_low >= 0, w1 = w_low + 1 <= W - 1;
                const float rw0 = h0 ? av * (1.f - lh) : 0.f, rw1 = h1 ? av * lh : 0.f;
                const float cw0 = w0 ? 1.f - lw : 0.f, cw1 = w1 ? lw : 0.f;
                const int y0 = h0 ? h_low : h_low + 1, y1 = h1 ? h_low + 1 : h_low;      // both rows are level rows
                const int x0 = w0 ? w_low : w_low + 1;                                     // left column is a level column
                const uint32_t dx = (w0 && w1) ? pix_b : 0u;                               // right column = left + dx
                const int pbase = cur.n * p.S + li.z + x0;
                const uint32_t off0 = (uint32_t)((pbase + y0 * W) * p.M + cur.m) * (uint32_t)ROWB;
                const uint32_t off1 = (uint32_t)((pbase + y1 * W) * p.M + cur.m) * (uint32_t)ROWB;
                const uint32_t ea = tab + lane * 32;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ea), "r"(off0), "r"(off1), "r"(dx),
                             "r"(__float_as_uint(rw0 * cw0))
                             : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ea + 16), "r"(__float_as_uint(rw0 * cw1)),
                             "r"(__float_as_uint(rw1 * cw0)), "r"(__float_as_uint(rw1 * cw1)), "r"(0u)
                             : "memory");
            }
            __syncwarp();
            // ---------------------------------------------------------------- phase 2: LPC lanes per sample
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                float2 acc[VEC / 2];
#pragma unroll
                for (int v = 0; v < VEC / 2; ++v) acc[v] = make_float2(0.f, 0.f);
#pragma unroll
                for (int it = 0; it < LK / NG; ++it) {
                    const uint32_t ea = tab + (uint32_t)((j * LK + it * NG + grp) * 32);
                    uint4 e0, e1;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(ea));
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(ea + 16));
                    const uint4 t00 = ldg_cached_128(vlane + e0.x);
                    const uint4 t01 = ldg_cached_128(vlane + e0.x + e0.z);
                    const uint4 t10 = ldg_cached_128(vlane + e0.y);
                    const uint4 t11 = ldg_cached_128(vlane + e0.y + e0.z);
                    const float w00 = __uint_as_float(e0.w), w01 = __uint_as_float(e1.x);
                    const float w10 = __uint_as_float(e1.y), w11 = __uint_as_float(e1.z);
                    if constexpr (VEC == 4) {
                        cached_ffma2(acc[0], w00, t00.x, t00.y); cached_ffma2(acc[1], w00, t00.z, t00.w);
                        cached_ffma2(acc[0], w01, t01.x, t01.y); cached_ffma2(acc[1], w01, t01.z, t01.w);
                        cached_ffma2(acc[0], w10, t10.x, t10.y); cached_ffma2(acc[1], w10, t10.z, t10.w);
                        cached_ffma2(acc[0], w11, t11.x, t11.y); cached_ffma2(acc[1], w11, t11.z, t11.w);
                    } else {
                        const uint4 tt[4] = {t00, t01, t10, t11};
                        const float ww[4] = {w00, w01, w10, w11};
#pragma unroll
                        for (int r = 0; r < 4; ++r) {              // bf16 -> fp32 is a 16-bit shift
                            cached_ffma2(acc[0], ww[r], tt[r].x << 16, tt[r].x & 0xffff0000u);
                            cached_ffma2(acc[1], ww[r], tt[r].y << 16, tt[r].y & 0xffff0000u);
                            cached_ffma2(acc[2], ww[r], tt[r].z << 16, tt[r].z & 0xffff0000u);
                            cached_ffma2(acc[3], ww[r], tt[r].w << 16, tt[r].w & 0xffff0000u);
                        }
                    }
                }
                float av[VEC];
#pragma unroll
                for (int v = 0; v < VEC / 2; ++v) {
                    av[2 * v] = acc[v].x;
                    av[2 * v + 1] = acc[v].y;
                }
                const float r0 = cached_transpose_reduce<VEC>(av, lane);
                // pair j of the pass = the query of lane j*LK
                const int orow = __shfl_sync(0xffffffffu, cur.row, j * LK);
                const bool olive = __shfl_sync(0xffffffffu, (int)cur.live, j * LK) != 0;
                if (olive) RW::store(p.out, ((long long)orow * p.M + cur.m) * 32 + VEC * sub + grp, r0);
            }
            __syncwarp();
            if (!more) break;
            cur = nxt;
            in = in_n;
            i = inext;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tile plan
static int cached_span(int nq_axis, int n_axis, int t) {   // pixels of a level covered by the centres of t queries
    return (int)(((long long)t * n_axis + nq_axis - 1) / nq_axis) + 1;
}

template <typename VT, int LK, int FRONT>
static int launch_cached(const CachedParams &p, int grid, cudaStream_t stream) {
    auto kern = msda_fwd_cached_kernel<VT, LK, FRONT>;
    static bool configured = false;
    if (!configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
        configured = true;
    }
    kern<<<grid, kCachedThreads, 0, stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

template <typename VT, int FRONT>
static int dispatch_cached(const CachedParams &p, int LK, int grid, cudaStream_t stream) {
    switch (LK) {
        case 8: return launch_cached<VT, 8, FRONT>(p, grid, stream);
        case 16: return launch_cached<VT, 16, FRONT>(p, grid, stream);
        case 32: return launch_cached<VT, 32, FRONT>(p, grid, stream);
        default: break;
    }
    set_error("msda cached kernel: unsupported L*K = %d", LK);
    return VNEXT_EINVAL;
}

// value [N,S,M,32]; queries are the S pixels of the pyramid in level order.  Exactly one front-end is non-null:
// (loc, attn) classic or (oa, ref) fused.  hshapes / hlsi: HOST copies of spatial_shapes / level_start_index.
int msda_forward_cached_device(int dtype, const void *value, const int64_t *hshapes, const int64_t *hlsi, const float *loc,
                               const float *attn, const float *oa, int oa_pitch, const float *ref, int ref_dim,
                               int ref_shared, int N, int S, int M, int L, int K, void *out, cudaStream_t stream) {
    if ((long long)N * S == 0) {
        set_variant("fwd_empty");
        return VNEXT_OK;
    }
    const int LK = L * K;
    const int esz = dtype == VNEXT_F32 ? 4 : 2, rowb = 32 * esz;
    VNEXT_CHECK_ARG((dtype == VNEXT_F32 || dtype == VNEXT_BF16) && L >= 1 && L <= kCachedMaxL && (LK == 8 || LK == 16 || LK == 32),
                    "msda cached kernel: unsupported dtype/L/K (%d, %d, %d)", dtype, L, K);
    VNEXT_CHECK_ARG((long long)N * S * M * rowb < (1ll << 32) && (long long)N * S * M * LK < (1ll << 30),
                    "msda cached kernel: tensor too large for 32-bit offsets");
    VNEXT_CHECK_ARG((reinterpret_cast<uintptr_t>(value) & 15u) == 0, "msda cached kernel: value must be 16-byte aligned");
    const bool fused = oa != nullptr && ref != nullptr;
    const bool pre = oa != nullptr && ref == nullptr;        // rows already hold locations | normalised attention weights
    if (fused) {
        VNEXT_CHECK_ARG(ref_dim == 2 && oa_pitch >= M * LK * 3 && oa_pitch % 2 == 0 &&
                            (reinterpret_cast<uintptr_t>(oa) & 7u) == 0 && (reinterpret_cast<uintptr_t>(ref) & 7u) == 0,
                        "msda cached kernel: bad fused front-end arguments (2-d reference points only: encoder self-attention)");
    } else if (pre) {
        VNEXT_CHECK_ARG(oa_pitch >= M * LK * 3 && oa_pitch % 2 == 0 && (reinterpret_cast<uintptr_t>(oa) & 7u) == 0,
                        "msda cached kernel: bad pre-located front-end arguments");
    } else {
        VNEXT_CHECK_ARG(loc && attn && (reinterpret_cast<uintptr_t>(loc) & 7u) == 0, "msda cached kernel: bad loc/attn pointers");
    }
    int H[kCachedMaxL], W[kCachedMaxL], start[kCachedMaxL];
    long long total = 0;
    for (int l = 0; l < L; ++l) {
        H[l] = (int)hshapes[2 * l]; W[l] = (int)hshapes[2 * l + 1]; start[l] = (int)hlsi[l];
        VNEXT_CHECK_ARG(H[l] > 0 && W[l] > 0 && start[l] == total, "msda cached kernel: levels must tile [0, S) in order");
        total += (long long)H[l] * W[l];
    }
    VNEXT_CHECK_ARG(total == S, "msda cached kernel: sum of level sizes (%lld) != spatial_size (%d)", total, S);

    CachedParams p;
    memset(&p, 0, sizeof(p));
    static const char *env_budget = getenv("VNEXT_MSDA_CACHED_KB");
    static const bool debug = getenv("VNEXT_WIN_DEBUG") != nullptr;
    // footprint budget of one unit: the L1 is 228 KB minus the 33 KB of static shared memory; two units overlap in time
    const double budget = (env_budget ? atof(env_budget) : 150.0) * 1024.0;
    const int reach = 4;                                           // pixels a sample typically strays from its query
    static const int cand[][2] = {{32, 32}, {16, 32}, {16, 16}, {8, 16}, {8, 8}, {4, 8}, {4, 4}, {2, 4}, {2, 2}, {1, 2}, {1, 1}};
    const int PPP = 32 / LK;
    for (int lq = 0; lq < L; ++lq) {
        int th = 1, tw = 1;
        for (auto &tt : cand) {
            if (tt[0] * tt[1] < PPP) continue;                     // a unit is at least one pass
            if (tt[0] > 1 && tt[0] / 2 >= H[lq]) continue;
            if (tt[1] > 1 && tt[1] / 2 >= W[lq]) continue;
            double bytes = 0.0;
            for (int l = 0; l < L; ++l)
                bytes += (double)(cached_span(H[lq], H[l], tt[0]) + 2 * reach) * (cached_span(W[lq], W[l], tt[1]) + 2 * reach) * 128.0;
            // (a pixel row of one head occupies a 128-B line in L1 for both fp32 and bf16 value: bf16 fills half of it)
            th = tt[0]; tw = tt[1];
            if (bytes <= budget) break;
        }
        if (th * tw < PPP) { th = 1; tw = PPP; }
        CachedClass &c = p.cls[lq];
        c.H = H[lq]; c.W = W[lq]; c.start = start[lq];
        c.th = th;
        c.tw_shift = 0;
        while ((1 << c.tw_shift) < tw) ++c.tw_shift;
        int np = th * tw * LK / 32;
        c.np_shift = 0;
        while ((1 << c.np_shift) < np) ++c.np_shift;
        c.tiles_x = (W[lq] + tw - 1) / tw;
        c.tiles_y = (H[lq] + th - 1) / th;
        c.units = N * c.tiles_x * c.tiles_y * M;
        c.d_tx = make_cdiv((uint32_t)c.tiles_x);
        c.d_ty = make_cdiv((uint32_t)c.tiles_y);
        if (debug) fprintf(stderr, "[msda cached] query level %d (%dx%d): tile %dx%d, %d units x %d passes\n", lq, H[lq], W[lq], th, tw, c.units, np);
    }
    for (int l = 0; l < L; ++l) {
        p.meta[l].H = H[l]; p.meta[l].W = W[l]; p.meta[l].start = start[l]; p.meta[l].pad = 0;
    }
    p.value = value; p.loc = loc; p.attn = attn; p.oa = oa; p.oa_pitch = oa_pitch; p.n_off = M * LK * 2;
    p.ref = ref; p.ref_batch_stride = ref_shared ? 0 : (long long)S * L * ref_dim;
    p.out = out; p.N = N; p.S = S; p.M = M; p.L = L; p.K = K;
    p.k_shift = 0;
    while ((1 << p.k_shift) < K) ++p.k_shift;
    p.d_M = make_cdiv((uint32_t)M);
    const int grid = num_sms();
    int rc;
    if (dtype == VNEXT_F32)
        rc = fused ? dispatch_cached<float, 1>(p, LK, grid, stream)
                   : (pre ? dispatch_cached<float, 2>(p, LK, grid, stream) : dispatch_cached<float, 0>(p, LK, grid, stream));
    else
        rc = fused ? dispatch_cached<__nv_bfloat16, 1>(p, LK, grid, stream)
                   : (pre ? dispatch_cached<__nv_bfloat16, 2>(p, LK, grid, stream) : dispatch_cached<__nv_bfloat16, 0>(p, LK, grid, stream));
    set_variant(dtype == VNEXT_F32 ? (fused ? "fwd_cached_fused_f32" : (pre ? "fwd_cached_pre_f32" : "fwd_cached_f32"))
                                   : (fused ? "fwd_cached_fused_bf16" : (pre ? "fwd_cached_pre_bf16" : "fwd_cached_bf16")));
    return rc;
}

}  // namespace vnext
