// Multi-scale deformable attention forward for SELF-attention over the pixel grid (encoder shape: the queries ARE the
// pixels of the flattened pyramid, Lq == S) — shared-memory window kernel for sm_100a.
//
// Why: for the encoder shape the tile kernel of msda_forward.cu is bound by the L2 -> L1 gather path, not by HBM
// (profiles/r01_msda_enc_v3_ncu_full.txt: 5.0 GB of L2 sectors per launch = 14x the algorithmic bytes, at the chip's
// L2 throughput cap): every (query, head) gathers L*K*4 corner rows from L2 and neighbouring queries re-fetch the
// same rows.  In the encoder a query's samples land around its OWN pixel in every level (reference point = pixel
// centre, learned offsets of a few pixels: projects/IDOL/idol/models/deformable_transformer.py:248-261,
// ops/modules/ms_deform_attn.py:98-105), so a spatial block of queries shares a small window of every level.
//
// Design:
//   * work unit = (batch n, query tile th x tw of ONE query level, head m).  One elected thread stages, per sampled
//     level, the window [tile footprint scaled to that level +- margin] of head m's 32 channels into shared memory
//     with ONE 4-D TMA box load (cp.async.bulk.tensor, dims {M*D, W_l, H_l, N}): the hardware zero-fills pixels
//     outside the level, which IS the reference's border rule (corners outside [0,H-1]x[0,W-1] contribute 0,
//     ms_deform_im2col_cuda.cuh:55-78), so the gather needs no corner predicates;
//   * phase 1 (one lane per sample, as in the tile kernel): sampling location (optionally fused softmax + location
//     arithmetic), floor, bilinear fractions; the lane parks ONE 16-byte entry per sample:
//     {shared-memory offsets of the two corner rows' left pixels, attn*(1-lh), attn*lh, lw};
//   * phase 2: a lane group (8 lanes x 16 B = one 128-B fp32 row; 4 lanes for bf16) reads the entry (one LDS.128,
//     broadcast) and the sample's four corner rows from the window (four LDS.128), 16 (32) FFMA; L2 sees each
//     window byte once per unit instead of once per sample.  Shared-memory bandwidth (128 B/clk/SM) is the new bound;
//   * samples whose 2x2 footprint leaves the staged window (offset larger than the margin), and levels whose window
//     does not fit the shared-memory budget for this tile class (fine levels under coarse queries), take the global
//     path inside the same loop: same arithmetic, corner rows read with LDG and explicit border predicates — exact
//     for ANY sampling locations, the window is only a cache;
//   * one persistent CTA of 16 warps per SM with TWO window buffers: the TMA fill of unit i+1 and the global loads of
//     its sampling inputs (kept in registers) fly during the gather of unit i (a first version with two independent
//     8-warp CTAs and no prefetch was latency-bound: 558 us vs 385 us for the tile kernel at 720p, see profiles/).
// Host side needs the level shapes on the HOST (the TMA descriptors are built per call from them): the entry point
// takes host copies of spatial_shapes / level_start_index (the Python layer caches them per shapes tensor).
#include <cuda.h>

#include "msda_device.cuh"

namespace vnext {

constexpr int kWinMaxL = 4;
constexpr int kWinWarps = 16;
constexpr int kWinThreads = kWinWarps * 32;
constexpr int kWinMaxPass = 4;               // 32-sample passes per warp per unit: a unit has <= 16 * 4 * 32 samples
constexpr int kWinZeroBytes = 256;           // zero rows at the start of dynamic smem (target of skipped samples)
constexpr uint32_t kWinGlobalFlag = 1u << 31;

struct WinLevelCfg {   // window of sampled level l under tiles of query level lq
    int staged;        // 0: samples of this level take the global path
    int bw, bh;        // box in pixels
    int smem_off;      // byte offset inside dynamic smem (128-B aligned)
    int pitch;         // bytes per window row
    int bytes;
};
struct WinClass {      // all tiles of one query level
    int th, tw, tiles_x, tiles_y;
    int unit_begin;
    int H, W, start;
    WinLevelCfg lv[kWinMaxL];
};
struct WinParams {
    CUtensorMap tmap[kWinMaxL * kWinMaxL];   // [lq][l]
    WinClass cls[kWinMaxL];
    LevelMeta meta[kWinMaxL];
    const void *value;
    const float *loc, *attn;                 // classic front-end
    const float *oa;                         // fused front-end (see msda_forward.cu)
    int oa_pitch, n_off;
    const float *ref;
    int ref_dim;
    long long ref_batch_stride;
    void *out;
    int N, S, M, L, K, margin, total_units;
    int buf_bytes;                           // size of one window buffer (two of them follow the zero rows)
};

__device__ __forceinline__ void tma_load_box4(uint32_t smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_dst),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int floordiv(int a, int b) {   // b > 0
    const int q = a / b;
    return (a % b < 0) ? q - 1 : q;
}

template <typename VT>
struct WinRow;
template <>
struct WinRow<float> {
    static constexpr int VEC = 4, LPC = 8, ROWB = 128;
    __device__ static __forceinline__ void widen(const uint4 &t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
    __device__ static __forceinline__ void store(void *out, long long idx, float v) { reinterpret_cast<float *>(out)[idx] = v; }
};
template <>
struct WinRow<__nv_bfloat16> {
    static constexpr int VEC = 8, LPC = 4, ROWB = 64;
    __device__ static __forceinline__ void widen(const uint4 &t, float (&v)[8]) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ void store(void *out, long long idx, float v) {
        reinterpret_cast<__nv_bfloat16 *>(out)[idx] = __float2bfloat16_rn(v);
    }
};

template <int VEC>
__device__ __forceinline__ float win_transpose_reduce(float (&a)[VEC], int lane) {   // see msda_forward.cu
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

struct WinUnit {
    int lq, m, n, qy0, qx0;
};
__device__ __forceinline__ WinUnit win_decode(const WinParams &p, int u) {
    WinUnit w;
    w.lq = 0;
    while (w.lq + 1 < p.L && u >= p.cls[w.lq + 1].unit_begin) ++w.lq;
    const WinClass &c = p.cls[w.lq];
    int r = u - c.unit_begin;
    w.m = r % p.M;
    r /= p.M;
    const int tx = r % c.tiles_x;
    r /= c.tiles_x;
    const int ty = r % c.tiles_y;
    w.n = r / c.tiles_y;
    w.qy0 = ty * c.th;
    w.qx0 = tx * c.tw;
    return w;
}

// One sample's raw front-end inputs, fetched one unit ahead of their use (the loads of unit i+1 fly during the gather of
// unit i): classic = (x, y, attention weight); fused = (offset x, offset y, logit) + reference point (x, y).
struct WinSampleIn {
    float a0, a1, a2, r0, r1;
};

template <typename VT, int LK, bool FUSED>
__global__ void __launch_bounds__(kWinThreads, 1) msda_fwd_window_kernel(const __grid_constant__ WinParams p) {
    using RW = WinRow<VT>;
    constexpr int VEC = RW::VEC, LPC = RW::LPC, NG = 32 / LPC, ROWB = RW::ROWB;
    constexpr int PPP = 32 / LK;          // (query, head) pairs per 32-sample warp pass
    constexpr int UPR = 32 / VEC;         // 16-byte units per (pixel, head) row
    static_assert(LK % NG == 0 && 32 % LK == 0, "unsupported L*K for the window kernel");

    extern __shared__ __align__(128) unsigned char win_smem[];   // [zero rows][window buffer 0][window buffer 1]
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ __align__(16) uint4 table[kWinWarps][32];   // {row offsets | flags, attn*hh, attn*lh, lw}
    __shared__ int2 gtab[kWinWarps][32];                   // global path: {unit offset of the top-left corner, row stride}

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane / LPC, sub = lane % LPC;
    const uint32_t smem_base = smem_u32(win_smem);

    if (tid < kWinZeroBytes / 4) reinterpret_cast<uint32_t *>(win_smem)[tid] = 0u;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // window fill of unit u into buffer b (one thread); returns nothing — whether a fill exists is recomputed by everyone
    auto issue = [&](int u, int b) {
        const WinUnit w = win_decode(p, u);
        const WinClass &c = p.cls[w.lq];
        uint32_t tx_bytes = 0;
#pragma unroll
        for (int l = 0; l < kWinMaxL; ++l)
            if (l < p.L && c.lv[l].staged) tx_bytes += (uint32_t)c.lv[l].bytes;
        if (!tx_bytes) return;
        fence_proxy_async();
        mbar_expect_tx(&bar[b], tx_bytes);
        const uint32_t base = smem_base + kWinZeroBytes + (uint32_t)b * (uint32_t)p.buf_bytes;
#pragma unroll
        for (int l = 0; l < kWinMaxL; ++l)
            if (l < p.L && c.lv[l].staged)
                tma_load_box4(base + c.lv[l].smem_off, &p.tmap[w.lq * kWinMaxL + l], w.m * 32,
                              floordiv((2 * w.qx0 + 1) * p.meta[l].W - c.W, 2 * c.W) - p.margin,
                              floordiv((2 * w.qy0 + 1) * p.meta[l].H - c.H, 2 * c.H) - p.margin, w.n, &bar[b]);
    };
    // raw inputs of this lane's samples in unit u (pass = warp + pi * kWinWarps)
    auto prefetch = [&](int u, WinSampleIn (&in)[kWinMaxPass]) {
        const WinUnit w = win_decode(p, u);
        const WinClass &c = p.cls[w.lq];
        const int nq = c.th * c.tw;
#pragma unroll
        for (int pi = 0; pi < kWinMaxPass; ++pi) {
            const int idx = (warp + pi * kWinWarps) * 32 + lane;
            const int qi = idx / LK, smp = idx % LK;
            const int qr = qi / c.tw, qc = qi - qr * c.tw;
            const int qy = w.qy0 + qr, qx = w.qx0 + qc;
            in[pi].a0 = in[pi].a1 = in[pi].a2 = in[pi].r0 = in[pi].r1 = 0.f;
            if (qi < nq && qy < c.H && qx < c.W) {
                const int q = c.start + qy * c.W + qx;
                const long long row = (long long)w.n * p.S + q;
                if (FUSED) {
                    const float *rowp = p.oa + row * p.oa_pitch;
                    const float2 off = __ldg(reinterpret_cast<const float2 *>(rowp + (w.m * LK + smp) * 2));
                    in[pi].a0 = off.x; in[pi].a1 = off.y;
                    in[pi].a2 = __ldg(rowp + p.n_off + w.m * LK + smp);
                    const float2 rf = __ldg(reinterpret_cast<const float2 *>(
                        p.ref + w.n * p.ref_batch_stride + ((size_t)q * p.L + smp / p.K) * 2));
                    in[pi].r0 = rf.x; in[pi].r1 = rf.y;
                } else {
                    const long long si = (row * p.M + w.m) * LK + smp;
                    const float2 xy = __ldg(reinterpret_cast<const float2 *>(p.loc) + si);
                    in[pi].a0 = xy.x; in[pi].a1 = xy.y;
                    in[pi].a2 = __ldg(p.attn + si);
                }
            }
        }
    };

    const unsigned char *vbytes = reinterpret_cast<const unsigned char *>(p.value);
    WinSampleIn cur[kWinMaxPass], nxt[kWinMaxPass];
    int u = blockIdx.x;
    if (u < p.total_units) {
        if (tid == 0) issue(u, 0);
        prefetch(u, cur);
    }
    uint32_t phase0 = 0, phase1 = 0;             // parity of the next fill of each buffer
    for (int it = 0; u < p.total_units; u += gridDim.x, ++it) {
        const int b = it & 1;
        const int un = u + gridDim.x;
        if (un < p.total_units) {                // buffer b^1 was released by the barrier that ended unit it-1
            if (tid == 0) issue(un, b ^ 1);
            prefetch(un, nxt);
        }
        const WinUnit w = win_decode(p, u);
        const WinClass &c = p.cls[w.lq];
        const int m = w.m, n = w.n, qy0 = w.qy0, qx0 = w.qx0;
        bool filled = false;
#pragma unroll
        for (int l = 0; l < kWinMaxL; ++l)
            if (l < p.L && c.lv[l].staged) filled = true;
        const uint32_t buf_off = kWinZeroBytes + (uint32_t)b * (uint32_t)p.buf_bytes;
        const int nq = c.th * c.tw;
        const int npass = (nq * LK + 31) / 32;
        bool waited = !filled;
#pragma unroll
        for (int pi = 0; pi < kWinMaxPass; ++pi) {
            const int pass = warp + pi * kWinWarps;
            if (pass >= npass) break;
            // ---- phase 1: one sample per lane
            const int idx = pass * 32 + lane;
            const int qi = idx / LK, smp = idx % LK;
            const int qr = qi / c.tw, qc = qi - qr * c.tw;
            const int qy = qy0 + qr, qx = qx0 + qc;
            const bool live = qi < nq && qy < c.H && qx < c.W;
            const int lvl = smp / p.K;
            const LevelMeta lm = p.meta[lvl];
            float x, y, a = cur[pi].a2;
            if (FUSED) {
                x = cur[pi].r0 + cur[pi].a0 / (float)lm.W;       // ref + offset / (W_l, H_l)   (module :102-105)
                y = cur[pi].r1 + cur[pi].a1 / (float)lm.H;
                float mx = live ? a : -INFINITY;
#pragma unroll
                for (int o = LK / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float e = live ? expf(a - mx) : 0.f;
                float sum = e;
#pragma unroll
                for (int o = LK / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                a = live ? e / sum : 0.f;
            } else {
                x = cur[pi].a0;
                y = cur[pi].a1;
            }
            {
                const float h_im = __fsub_rn(__fmul_rn(y, (float)lm.H), 0.5f);
                const float w_im = __fsub_rn(__fmul_rn(x, (float)lm.W), 0.5f);
                const bool inside = live && (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)lm.H) && (w_im < (float)lm.W);
                const float hf = floorf(h_im), wf = floorf(w_im);
                uint32_t packed = 0u;                                 // both rows -> the zero region
                float ahh = 0.f, alh = 0.f, lw = 0.f;                 // skipped samples: all-zero weights (no NaN leaks)
                if (inside) {
                    const int h_low = (int)hf, w_low = (int)wf;
                    const float lh = h_im - hf;
                    lw = w_im - wf;
                    ahh = a * (1.f - lh);
                    alh = a * lh;
                    const WinLevelCfg cfg = c.lv[lvl];
                    const int rx = w_low - (floordiv((2 * qx0 + 1) * lm.W - c.W, 2 * c.W) - p.margin);
                    const int ry = h_low - (floordiv((2 * qy0 + 1) * lm.H - c.H, 2 * c.H) - p.margin);
                    if (cfg.staged && rx >= 0 && rx + 1 < cfg.bw && ry >= 0 && ry + 1 < cfg.bh) {
                        const uint32_t off0 = buf_off + (uint32_t)(cfg.smem_off + ry * cfg.pitch + rx * ROWB);
                        packed = (off0 >> 4) | (((off0 + (uint32_t)cfg.pitch) >> 4) << 14);
                    } else {
                        const uint32_t v = (h_low >= 0 ? 1u : 0u) | (h_low + 1 <= lm.H - 1 ? 2u : 0u) |
                                           (w_low >= 0 ? 4u : 0u) | (w_low + 1 <= lm.W - 1 ? 8u : 0u);
                        packed = kWinGlobalFlag | v;
                        gtab[warp][lane] = make_int2((((int)(n * p.S) + lm.start + h_low * lm.W + w_low) * p.M + m) * UPR,
                                                     lm.W * p.M * UPR);
                    }
                }
                table[warp][lane] = make_uint4(packed, __float_as_uint(ahh), __float_as_uint(alh), __float_as_uint(lw));
            }
            __syncwarp();
            if (!waited) {
                mbar_wait(&bar[b], b ? phase1 : phase0);
                waited = true;
            }
            // ---- phase 2: one lane group per sample, four corner rows per sample
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                float acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
                for (int i = 0; i < LK / NG; ++i) {
                    const int s = j * LK + i * NG + grp;
                    const uint4 e = table[warp][s];
                    const float ahh = __uint_as_float(e.y), alh = __uint_as_float(e.z), lw = __uint_as_float(e.w);
                    const float hw = 1.f - lw;
                    const float w00 = ahh * hw, w01 = ahh * lw, w10 = alh * hw, w11 = alh * lw;
                    uint4 t00, t01, t10, t11;
                    if (!(e.x & kWinGlobalFlag)) {
                        const uint32_t a0 = smem_base + ((e.x & 0x3fffu) << 4) + sub * 16;
                        const uint32_t a1 = smem_base + (((e.x >> 14) & 0x3fffu) << 4) + sub * 16;
                        t00 = lds128(a0);
                        t01 = lds128(a0 + ROWB);
                        t10 = lds128(a1);
                        t11 = lds128(a1 + ROWB);
                    } else {
                        const int2 g = gtab[warp][s];
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        const uint4 *bp = reinterpret_cast<const uint4 *>(vbytes + (long long)g.x * 16) + sub;
                        const bool h0 = e.x & 1u, h1 = e.x & 2u, w0 = e.x & 4u, w1 = e.x & 8u;
                        t00 = (h0 && w0) ? __ldg(bp) : z;
                        t01 = (h0 && w1) ? __ldg(bp + p.M * UPR) : z;
                        t10 = (h1 && w0) ? __ldg(bp + g.y) : z;
                        t11 = (h1 && w1) ? __ldg(bp + g.y + p.M * UPR) : z;
                    }
                    float v[VEC];
                    RW::widen(t00, v);
#pragma unroll
                    for (int cc = 0; cc < VEC; ++cc) acc[cc] = fmaf(w00, v[cc], acc[cc]);
                    RW::widen(t01, v);
#pragma unroll
                    for (int cc = 0; cc < VEC; ++cc) acc[cc] = fmaf(w01, v[cc], acc[cc]);
                    RW::widen(t10, v);
#pragma unroll
                    for (int cc = 0; cc < VEC; ++cc) acc[cc] = fmaf(w10, v[cc], acc[cc]);
                    RW::widen(t11, v);
#pragma unroll
                    for (int cc = 0; cc < VEC; ++cc) acc[cc] = fmaf(w11, v[cc], acc[cc]);
                }
                const float r0 = win_transpose_reduce<VEC>(acc, lane);
                // pair j of this pass = query pass*PPP + j of the tile
                const int qj = pass * PPP + j;
                const int jr = qj / c.tw, jc = qj - jr * c.tw;
                if (qj < nq && qy0 + jr < c.H && qx0 + jc < c.W) {
                    const long long orow = (long long)n * p.S + c.start + (qy0 + jr) * c.W + (qx0 + jc);
                    RW::store(p.out, (orow * p.M + m) * 32 + VEC * sub + grp, r0);
                }
            }
            __syncwarp();
        }
        if (!waited) mbar_wait(&bar[b], b ? phase1 : phase0);   // every thread observes every fill
        if (filled) {
            if (b) phase1 ^= 1u; else phase0 ^= 1u;
        }
        __syncthreads();                        // all gathers of this unit are done: its buffer may be refilled
#pragma unroll
        for (int pi = 0; pi < kWinMaxPass; ++pi) cur[pi] = nxt[pi];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tile classes, windows, tensor maps
typedef CUresult (*PFN_encodeTiledW)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledW win_get_encode() {
    static PFN_encodeTiledW fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiledW>(ptr);
    }
    return fn;
}

static inline int host_floordiv(long long a, long long b) {
    long long q = a / b;
    if (a % b < 0) --q;
    return (int)q;
}

// widest span floor(centre(last query of a tile)) - floor(centre(first query)) over all tiles of `t` queries along an
// axis with nq_axis queries, mapped to a level with n_axis pixels
static int max_span(int nq_axis, int n_axis, int t) {
    int best = 0;
    for (int q0 = 0; q0 < nq_axis; q0 += t) {
        const int q1 = (q0 + t - 1 < nq_axis - 1) ? q0 + t - 1 : nq_axis - 1;
        const int f0 = host_floordiv((2ll * q0 + 1) * n_axis - nq_axis, 2ll * nq_axis);
        const int f1 = host_floordiv((2ll * q1 + 1) * n_axis - nq_axis, 2ll * nq_axis);
        if (f1 - f0 > best) best = f1 - f0;
    }
    return best;
}

struct WinPlanChoice {
    int th, tw;
    WinLevelCfg lv[kWinMaxL];
    double cost;
    size_t smem;
};

// Pick, for query level lq, the tile and the set of staged levels that minimise an L2-traffic estimate per query.
static bool plan_class(int lq, int L, int K, const int *H, const int *W, int rowb, int margin, size_t budget,
                       WinPlanChoice *best) {
    static const int cand[][2] = {{16, 16}, {8, 16}, {8, 8}, {4, 8}, {4, 4}, {2, 4}, {2, 2}, {1, 2}, {1, 1}};
    bool found = false;
    best->cost = 1e300;
    for (auto &tt : cand) {
        const int th = tt[0] < H[lq] ? tt[0] : H[lq], tw = tt[1] < W[lq] ? tt[1] : W[lq];
        const double nq = (double)th * tw;
        if (th * tw * L * K > kWinWarps * kWinMaxPass * 32) continue;     // a unit's samples must fit the warps' passes
        // staging options: levels >= first (coarser or equal to `first`), first = 0 .. L (L = nothing staged)
        for (int first = 0; first <= L; ++first) {
            WinPlanChoice ch;
            ch.th = th; ch.tw = tw;
            size_t off = 0;                        // offsets are relative to the window buffer
            double staged_bytes = 0.0, unstaged = 0.0;
            bool ok = true;
            for (int l = 0; l < L; ++l) {
                WinLevelCfg &cfg = ch.lv[l];
                cfg.staged = 0; cfg.bw = cfg.bh = cfg.smem_off = cfg.pitch = cfg.bytes = 0;
                if (l < first) {
                    unstaged += 1.0;
                    continue;
                }
                const int bw = max_span(W[lq], W[l], tw) + 2 * margin + 2;
                const int bh = max_span(H[lq], H[l], th) + 2 * margin + 2;
                if (bw > 256 || bh > 256) { ok = false; break; }
                cfg.staged = 1; cfg.bw = bw; cfg.bh = bh; cfg.pitch = bw * rowb; cfg.bytes = bw * bh * rowb;
                cfg.smem_off = (int)off;
                off += ((size_t)cfg.bytes + 127) / 128 * 128;
                staged_bytes += cfg.bytes;
            }
            if (!ok || off > budget || off >= (1u << 18)) continue;
            // per query: window bytes pulled from L2 + un-staged levels' corner rows (L2 gather, penalised 2x: latency-
            // bound path) + a fixed per-unit cost (barriers, pipeline bubbles of small units) worth ~64 KB of traffic
            const double cost = (staged_bytes + 65536.0) / nq + unstaged * K * 4 * rowb * 2.0;
            if (cost < best->cost) {
                *best = ch;
                best->cost = cost;
                best->smem = off;
                found = true;
            }
        }
    }
    return found;
}

template <typename VT, int LK, bool FUSED>
static int launch_window(const WinParams &p, size_t smem, cudaStream_t stream) {
    auto kern = msda_fwd_window_kernel<VT, LK, FUSED>;
    VNEXT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    VNEXT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWinThreads, smem));
    if (occ < 1) {
        set_error("msda window kernel: %zu bytes of shared memory do not fit", smem);
        return VNEXT_EINVAL;
    }
    const int grid = p.total_units < num_sms() * occ ? p.total_units : num_sms() * occ;
    kern<<<grid, kWinThreads, smem, stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

template <typename VT, bool FUSED>
static int dispatch_window(const WinParams &p, int LK, size_t smem, cudaStream_t stream) {
    switch (LK) {
        case 8: return launch_window<VT, 8, FUSED>(p, smem, stream);
        case 16: return launch_window<VT, 16, FUSED>(p, smem, stream);
        case 32: return launch_window<VT, 32, FUSED>(p, smem, stream);
        default: break;
    }
    set_error("msda window kernel: unsupported L*K = %d", LK);
    return VNEXT_EINVAL;
}

bool msda_window_supported(int dtype, int D, int L, int K) {
    const int LK = L * K;
    return (dtype == VNEXT_F32 || dtype == VNEXT_BF16) && D == 32 && L >= 1 && L <= kWinMaxL &&
           (LK == 8 || LK == 16 || LK == 32);
}

// value [N,S,M,32]; queries are the S pixels of the pyramid in level order.  Exactly one front-end is non-null:
// (loc, attn) classic or (oa, ref) fused.  hshapes / hlsi: HOST copies of spatial_shapes / level_start_index.
int msda_forward_window_device(int dtype, const void *value, const int64_t *hshapes, const int64_t *hlsi, const float *loc,
                               const float *attn, const float *oa, int oa_pitch, const float *ref, int ref_dim,
                               int ref_shared, int N, int S, int M, int L, int K, int margin, void *out,
                               cudaStream_t stream) {
    if ((long long)N * S == 0) {
        set_variant("fwd_empty");
        return VNEXT_OK;
    }
    const int LK = L * K;
    VNEXT_CHECK_ARG(msda_window_supported(dtype, 32, L, K), "msda window kernel: unsupported dtype/L/K (%d, %d, %d)", dtype, L, K);
    VNEXT_CHECK_ARG((long long)N * S * M * 32 < (1ll << 31) && (long long)N * S * M < (1ll << 27),
                    "msda window kernel: tensor too large for 32-bit offsets");
    VNEXT_CHECK_ARG((reinterpret_cast<uintptr_t>(value) & 15u) == 0, "msda window kernel: value must be 16-byte aligned");
    VNEXT_CHECK_ARG(margin >= 0 && margin <= 64, "msda window kernel: margin out of range");
    const bool fused = oa != nullptr;
    if (fused) {
        VNEXT_CHECK_ARG(ref && ref_dim == 2 && oa_pitch >= M * LK * 3 && oa_pitch % 2 == 0 &&
                            (reinterpret_cast<uintptr_t>(oa) & 7u) == 0 && (reinterpret_cast<uintptr_t>(ref) & 7u) == 0,
                        "msda window kernel: bad fused front-end arguments (2-d reference points only: encoder self-attention)");
    } else {
        VNEXT_CHECK_ARG(loc && attn && (reinterpret_cast<uintptr_t>(loc) & 7u) == 0, "msda window kernel: bad loc/attn pointers");
    }
    PFN_encodeTiledW enc = win_get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VNEXT_ECUDA;
    }
    int H[kWinMaxL], W[kWinMaxL], start[kWinMaxL];
    long long total = 0;
    for (int l = 0; l < L; ++l) {
        H[l] = (int)hshapes[2 * l]; W[l] = (int)hshapes[2 * l + 1]; start[l] = (int)hlsi[l];
        VNEXT_CHECK_ARG(H[l] > 0 && W[l] > 0 && start[l] == total, "msda window kernel: levels must tile [0, S) in order");
        total += (long long)H[l] * W[l];
    }
    VNEXT_CHECK_ARG(total == S, "msda window kernel: sum of level sizes (%lld) != spatial_size (%d)", total, S);

    WinParams p;
    memset(&p, 0, sizeof(p));
    const int esz = dtype == VNEXT_F32 ? 4 : 2, rowb = 32 * esz;
    const size_t budget = 106 * 1024;             // per window buffer; two buffers + zero rows + 12 KB static < 227 KB
    size_t smem = 0;
    int units = 0;
    for (int lq = 0; lq < L; ++lq) {
        WinPlanChoice ch;
        if (!plan_class(lq, L, K, H, W, rowb, margin, budget, &ch)) {
            set_error("msda window kernel: no tile plan for level %d", lq);
            return VNEXT_EINVAL;
        }
        WinClass &c = p.cls[lq];
        c.th = ch.th; c.tw = ch.tw;
        c.tiles_x = (W[lq] + ch.tw - 1) / ch.tw;
        c.tiles_y = (H[lq] + ch.th - 1) / ch.th;
        c.unit_begin = units;
        c.H = H[lq]; c.W = W[lq]; c.start = start[lq];
        units += N * c.tiles_x * c.tiles_y * M;
        if (ch.smem > smem) smem = ch.smem;
        for (int l = 0; l < L; ++l) {
            c.lv[l] = ch.lv[l];
            if (!ch.lv[l].staged) continue;
            cuuint64_t dims[4] = {(cuuint64_t)M * 32, (cuuint64_t)W[l], (cuuint64_t)H[l], (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)M * 32 * esz, (cuuint64_t)W[l] * M * 32 * esz, (cuuint64_t)S * M * 32 * esz};
            cuuint32_t box[4] = {32u, (cuuint32_t)ch.lv[l].bw, (cuuint32_t)ch.lv[l].bh, 1u};
            cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
            void *base = const_cast<unsigned char *>(reinterpret_cast<const unsigned char *>(value)) + (size_t)start[l] * M * 32 * esz;
            CUresult r = enc(&p.tmap[lq * kWinMaxL + l],
                             dtype == VNEXT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base,
                             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled (msda window lq=%d l=%d box %dx%d) failed (%d)", lq, l, ch.lv[l].bw, ch.lv[l].bh, (int)r);
                return VNEXT_ECUDA;
            }
        }
    }
    for (int l = 0; l < L; ++l) {
        p.meta[l].H = H[l]; p.meta[l].W = W[l]; p.meta[l].start = start[l]; p.meta[l].pad = 0;
    }
    p.value = value; p.loc = loc; p.attn = attn; p.oa = oa; p.oa_pitch = oa_pitch; p.n_off = M * LK * 2;
    p.ref = ref; p.ref_dim = ref_dim; p.ref_batch_stride = ref_shared ? 0 : (long long)S * L * ref_dim;
    p.out = out; p.N = N; p.S = S; p.M = M; p.L = L; p.K = K; p.margin = margin; p.total_units = units;
    p.buf_bytes = (int)((smem + 127) / 128 * 128);
    smem = kWinZeroBytes + 2 * (size_t)p.buf_bytes;
    int rc;
    if (dtype == VNEXT_F32)
        rc = fused ? dispatch_window<float, true>(p, LK, smem, stream) : dispatch_window<float, false>(p, LK, smem, stream);
    else
        rc = fused ? dispatch_window<__nv_bfloat16, true>(p, LK, smem, stream)
                   : dispatch_window<__nv_bfloat16, false>(p, LK, smem, stream);
    set_variant(dtype == VNEXT_F32 ? (fused ? "fwd_window_fused_f32" : "fwd_window_f32")
                                   : (fused ? "fwd_window_fused_bf16" : "fwd_window_bf16"));
    return rc;
}

}  // namespace vnext
