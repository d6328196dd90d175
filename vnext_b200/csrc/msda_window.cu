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
// Design (third version; the second one was issue-bound: 460 warp instructions per (query, head), ncu
// profiles/r02_msda_window_v2_ncu.txt — integer divisions, per-lane parameter-bank indexing, bf16 widening):
//   * work unit = (batch n, query tile th x tw of ONE query level, head m), tw a power of two.  One elected thread
//     decodes units ONE AHEAD into a shared-memory descriptor (window origins, shared addresses, pitches per level) and
//     stages, per sampled level, the window [tile footprint scaled to that level +- margin] of head m's 32 channels with
//     ONE 4-D TMA box load (cp.async.bulk.tensor, dims {M*D, W_l, H_l, N}): the hardware zero-fills pixels outside the
//     level, which IS the reference's border rule (corners outside [0,H-1]x[0,W-1] contribute 0,
//     ms_deform_im2col_cuda.cuh:55-78), so the gather needs no corner predicates;
//   * phase 1 (one lane per sample): sampling location (optionally fused softmax + location arithmetic), floor,
//     bilinear weights; the lane parks a 32-byte entry per sample holding the two shared-memory row addresses and the
//     four corner weights — everything phase 2 needs, so phase 2 has no arithmetic besides the gather itself;
//   * phase 2: fp32 value: 8 lanes x 16 B = one 128-B pixel row, four LDS.128 + eight packed FFMA2 per sample and lane
//     (fma.rn.f32x2); bf16 value: the two horizontally adjacent pixels of a corner pair are ONE 128-B contiguous chunk
//     (2 x 64 B), read by 8 lanes with one LDS.128 each (lanes 0-3 left pixel, lanes 4-7 right pixel) — always
//     bank-conflict-free, which 64-B rows addressed independently are not;
//   * samples whose 2x2 footprint leaves the staged window (offset larger than the margin), and levels whose window
//     does not fit the shared-memory budget for this tile class (fine levels under coarse queries), take the global
//     path (LDG with explicit border predicates) — exact for ANY sampling locations, the window is only a cache.  A
//     warp vote per pass selects the loop version without per-sample checks when the whole pass is in-window;
//   * one persistent CTA of 16 warps per SM with TWO window buffers: the TMA fill of unit i+1 and the global loads of
//     its sampling inputs (kept in registers) fly during the gather of unit i.
// Host side needs the level shapes on the HOST (the TMA descriptors are built per call from them): the entry point
// takes host copies of spatial_shapes / level_start_index (the Python layer caches them per shapes tensor).
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

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
    int smem_off;      // byte offset inside a window buffer (128-B aligned)
    int pitch;         // bytes per window row
    int bytes;
};
struct WinClass {      // all tiles of one query level
    int th, tw, tw_shift, tiles_x, tiles_y;
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
    int N, S, M, L, K, k_shift, margin, total_units;
    int buf_bytes;                           // size of one window buffer (two of them follow the zero rows)
};

// What phase 1 needs to know about sampled level l in the current unit: two 16-byte shared-memory reads per lane.
struct __align__(16) WinLevelInfo {
    float fW, fH;                            // level size
    int ox, oy;                              // window origin in level pixels (may be negative: zero-filled by TMA)
    uint32_t bwm1, bhm1;                     // box size - 1 (0: level not staged -> every sample takes the global path)
    uint32_t saddr, pitch;                   // shared byte address of the window in this unit's buffer, row pitch
    int W, H, start, pad;                    // global path
};
struct __align__(16) WinUnitInfo {
    WinLevelInfo lv[kWinMaxL];
    int qy0, qx0, qH, qW;                    // tile origin, query-level size
    int row0, tw_shift, nq, npass;           // row0 = n*S + start of the query level
    int m, n, valid, tx_bytes;
    int lq, pad0, pad1, pad2;
};
struct __align__(16) WinEntry {              // one sample, written by its phase-1 lane
    uint32_t a0, a1;                         // shared byte addresses of the (y, x) and (y+1, x) rows; a0 bit 31: global path
    float w0, w1;                            // fp32: w00, w01            bf16: left pixel  (w00, w10)
    float w2, w3;                            // fp32: w10, w11            bf16: a0/a1 again are in e1 (see below)
    uint32_t g0, g1;                         // global path: unit offset of the top-left corner, row stride (16-B units)
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
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int floordiv(int a, int b) {   // b > 0
    const int q = a / b;
    return (a % b < 0) ? q - 1 : q;
}
// acc(2 x fp32) += w * v(2 x fp32): one packed FFMA2 (sm_100)
__device__ __forceinline__ void ffma2(float2 &acc, float w, uint32_t lo, uint32_t hi) {
    unsigned long long d;
    const unsigned long long a = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
    const unsigned long long b = ((unsigned long long)hi << 32) | lo;
    const unsigned long long c = ((unsigned long long)__float_as_uint(acc.y) << 32) | __float_as_uint(acc.x);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    acc.x = __uint_as_float((uint32_t)d);
    acc.y = __uint_as_float((uint32_t)(d >> 32));
}

template <typename VT>
struct WinRow;
template <>
struct WinRow<float> {
    static constexpr int VEC = 4, LPC = 8, ROWB = 128;
    __device__ static __forceinline__ void store(void *out, int idx, float v) { reinterpret_cast<float *>(out)[idx] = v; }
};
template <>
struct WinRow<__nv_bfloat16> {
    static constexpr int VEC = 8, LPC = 8, ROWB = 64;       // 8 lanes per sample: 4 per pixel of the horizontal pair
    __device__ static __forceinline__ void store(void *out, int idx, float v) {
        reinterpret_cast<__nv_bfloat16 *>(out)[idx] = __float2bfloat16_rn(v);
    }
};

// Sum a[0..N) over the lanes that differ in the bits of `mask` .. 4 (N = 2, 4 or 8 -> 1, 2 or 3 lane bits), leaving ONE
// value per lane: at each step a lane keeps the half of the vector selected by its own bit and adds the partner's.
template <int N>
__device__ __forceinline__ float win_transpose_reduce(float (&a)[N], int lane, int mask) {
#pragma unroll
    for (int n = N; n > 1; n >>= 1, mask >>= 1) {
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

// One sample's raw front-end inputs, fetched one unit ahead of their use (the loads of unit i+1 fly during the gather of
// unit i): classic = (x, y, attention weight); fused = (offset x, offset y, logit) + reference point (x, y).
template <bool FUSED>
struct WinSampleIn;
template <>
struct WinSampleIn<false> {
    float a0, a1, a2;
};
template <>
struct WinSampleIn<true> {
    float a0, a1, a2, r0, r1;
};

template <typename VT, int LK, bool FUSED>
__global__ void __launch_bounds__(kWinThreads, 1) msda_fwd_window_kernel(const __grid_constant__ WinParams p) {
    using RW = WinRow<VT>;
    constexpr bool BF = sizeof(VT) == 2;
    constexpr int NG = 4;                 // samples per warp instruction of phase 2 (8 lanes each)
    constexpr int PPP = 32 / LK;          // (query, head) pairs per 32-sample warp pass
    constexpr int LKS = LK == 8 ? 3 : (LK == 16 ? 4 : 5);
    constexpr int UPR = BF ? 4 : 8;       // 16-byte units per (pixel, head) row
    constexpr int ROWB = RW::ROWB;
    static_assert(LK == 8 || LK == 16 || LK == 32, "unsupported L*K for the window kernel");

    extern __shared__ __align__(128) unsigned char win_smem[];   // [zero rows][window buffer 0][window buffer 1]
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ WinUnitInfo info[3];
    __shared__ WinEntry table[kWinWarps][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, lane7 = lane & 7;
    const uint32_t smem_base = smem_u32(win_smem);
    const uint32_t tab_base = smem_u32(&table[warp][0]);

    if (tid < kWinZeroBytes / 4) reinterpret_cast<uint32_t *>(win_smem)[tid] = 0u;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }

    // ---- thread 0: decode unit u into info[slot] (its window goes to buffer b)
    auto decode = [&](int u, int slot, int b) {
        WinUnitInfo &I = info[slot];
        if (u >= p.total_units) {
            I.valid = 0;
            return;
        }
        int lq = 0;
        while (lq + 1 < p.L && u >= p.cls[lq + 1].unit_begin) ++lq;
        const WinClass &c = p.cls[lq];
        int r = u - c.unit_begin;
        const int m = r % p.M;
        r /= p.M;
        const int tx = r % c.tiles_x;
        r /= c.tiles_x;
        const int ty = r % c.tiles_y;
        const int n = r / c.tiles_y;
        const int qy0 = ty * c.th, qx0 = tx * c.tw;
        const uint32_t base = smem_base + kWinZeroBytes + (uint32_t)b * (uint32_t)p.buf_bytes;
        int tx_bytes = 0;
        for (int l = 0; l < p.L; ++l) {
            const WinLevelCfg &cfg = c.lv[l];
            WinLevelInfo &V = I.lv[l];
            const int W = p.meta[l].W, H = p.meta[l].H;
            V.fW = (float)W; V.fH = (float)H;
            V.ox = floordiv((2 * qx0 + 1) * W - c.W, 2 * c.W) - p.margin;
            V.oy = floordiv((2 * qy0 + 1) * H - c.H, 2 * c.H) - p.margin;
            V.bwm1 = cfg.staged ? (uint32_t)(cfg.bw - 1) : 0u;
            V.bhm1 = cfg.staged ? (uint32_t)(cfg.bh - 1) : 0u;
            V.saddr = base + (uint32_t)cfg.smem_off;
            V.pitch = (uint32_t)cfg.pitch;
            V.W = W; V.H = H; V.start = p.meta[l].start; V.pad = 0;
            if (cfg.staged) tx_bytes += cfg.bytes;
        }
        const int nq = c.th * c.tw;
        I.qy0 = qy0; I.qx0 = qx0; I.qH = c.H; I.qW = c.W;
        I.row0 = n * p.S + c.start; I.tw_shift = c.tw_shift; I.nq = nq; I.npass = (nq * LK + 31) >> 5;
        I.m = m; I.n = n; I.valid = 1; I.tx_bytes = tx_bytes; I.lq = lq;
    };
    // ---- thread 0: start the window fill described by info[slot] (buffer b)
    auto issue = [&](int slot, int b) {
        const WinUnitInfo &I = info[slot];
        if (!I.valid || !I.tx_bytes) return;
        fence_proxy_async();
        mbar_expect_tx(&bar[b], (uint32_t)I.tx_bytes);
        for (int l = 0; l < p.L; ++l)
            if (I.lv[l].bwm1)
                tma_load_box4(I.lv[l].saddr, &p.tmap[I.lq * kWinMaxL + l], I.m * 32, I.lv[l].ox, I.lv[l].oy, I.n, &bar[b]);
    };
    // ---- raw inputs of this lane's samples in the unit described by I (pass = warp + pi * kWinWarps)
    auto prefetch = [&](const WinUnitInfo &I, WinSampleIn<FUSED> (&in)[kWinMaxPass]) {
        const int tws = I.tw_shift, qy0 = I.qy0, qx0 = I.qx0, qH = I.qH, qW = I.qW, nq = I.nq, row0 = I.row0, m = I.m;
        const int smp = lane & (LK - 1);
#pragma unroll
        for (int pi = 0; pi < kWinMaxPass; ++pi) {
            const int qi = (warp + pi * kWinWarps) * PPP + (lane >> LKS);
            const int qy = qy0 + (qi >> tws), qx = qx0 + (qi & ((1 << tws) - 1));
            in[pi].a0 = in[pi].a1 = in[pi].a2 = 0.f;
            if constexpr (FUSED) in[pi].r0 = in[pi].r1 = 0.f;
            if (qi < nq && qy < qH && qx < qW) {
                const int row = row0 + qy * qW + qx;               // n * S + q
                if constexpr (FUSED) {
                    const float *rowp = p.oa + (size_t)row * p.oa_pitch;
                    const float2 off = __ldg(reinterpret_cast<const float2 *>(rowp + (m * LK + smp) * 2));
                    in[pi].a0 = off.x; in[pi].a1 = off.y;
                    in[pi].a2 = __ldg(rowp + p.n_off + m * LK + smp);
                    const float2 rf = __ldg(reinterpret_cast<const float2 *>(
                        p.ref + I.n * p.ref_batch_stride + ((size_t)(row - I.n * p.S) * p.L + (smp >> p.k_shift)) * 2));
                    in[pi].r0 = rf.x; in[pi].r1 = rf.y;
                } else {
                    const size_t si = ((size_t)row * p.M + m) * LK + smp;
                    const float2 xy = __ldg(reinterpret_cast<const float2 *>(p.loc) + si);
                    in[pi].a0 = xy.x; in[pi].a1 = xy.y;
                    in[pi].a2 = __ldg(p.attn + si);
                }
            }
        }
    };

    const unsigned char *vbytes = reinterpret_cast<const unsigned char *>(p.value);
    WinSampleIn<FUSED> cur[kWinMaxPass], nxt[kWinMaxPass];
    const int u0 = blockIdx.x, ustep = gridDim.x;
    if (tid == 0) {
        decode(u0, 0, 0);
        decode(u0 + ustep, 1, 1);
    }
    __syncthreads();                             // zero rows, barriers and the first two descriptors are visible
    if (tid == 0) issue(0, 0);
    if (info[0].valid) prefetch(info[0], cur);
    uint32_t phase0 = 0, phase1 = 0;             // parity of the next fill of each buffer
    int slot = 0;
    for (int it = 0;; ++it) {
        const WinUnitInfo &I = info[slot];
        if (!I.valid) break;
        const int b = it & 1;
        const int slot_n = slot == 2 ? 0 : slot + 1, slot_nn = slot_n == 2 ? 0 : slot_n + 1;
        const WinUnitInfo &In = info[slot_n];
        if (tid == 0) {
            issue(slot_n, b ^ 1);                // buffer b^1 was released by the barrier that ended unit it-1
            decode(u0 + (it + 2) * ustep, slot_nn, b);
        }
        const bool more = In.valid != 0;
        if (more) prefetch(In, nxt);

        const int m = I.m, n = I.n, qy0 = I.qy0, qx0 = I.qx0, qH = I.qH, qW = I.qW, tws = I.tw_shift, nq = I.nq;
        const int npass = I.npass, row0 = I.row0;
        const bool filled = I.tx_bytes != 0;
        bool waited = !filled;
#pragma unroll 1
        for (int pi = 0; pi < kWinMaxPass; ++pi) {
            const int pass = warp + pi * kWinWarps;
            if (pass >= npass) break;
            // ---------------------------------------------------------------- phase 1: one sample per lane
            bool any_global;
            {
                const int smp = lane & (LK - 1);
                const int qi = pass * PPP + (lane >> LKS);
                const int qy = qy0 + (qi >> tws), qx = qx0 + (qi & ((1 << tws) - 1));
                const bool live = qi < nq && qy < qH && qx < qW;
                const int lvl = smp >> p.k_shift;
                const uint32_t lvaddr = smem_u32(&I.lv[lvl]);
                const uint4 la = lds128(lvaddr), lb = lds128(lvaddr + 16);
                const float fW = __uint_as_float(la.x), fH = __uint_as_float(la.y);
                WinSampleIn<FUSED> in = cur[0];
#pragma unroll
                for (int k = 1; k < kWinMaxPass; ++k)
                    if (pi == k) in = cur[k];
                float x, y, a = in.a2;
                if constexpr (FUSED) {
                    x = in.r0 + in.a0 / fW;                       // ref + offset / (W_l, H_l)   (module :102-105)
                    y = in.r1 + in.a1 / fH;
                    float mx = live ? a : -INFINITY;
#pragma unroll
                    for (int o = LK / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    const float e = live ? expf(a - mx) : 0.f;
                    float sum = e;
#pragma unroll
                    for (int o = LK / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    a = live ? e / sum : 0.f;
                } else {
                    x = in.a0;
                    y = in.a1;
                }
                const float h_im = __fsub_rn(__fmul_rn(y, fH), 0.5f);
                const float w_im = __fsub_rn(__fmul_rn(x, fW), 0.5f);
                const bool inside = live && (h_im > -1.f) && (w_im > -1.f) && (h_im < fH) && (w_im < fW);
                const float hf = floorf(h_im), wf = floorf(w_im);
                const int h_low = (int)hf, w_low = (int)wf;
                const float lh = h_im - hf;
                // skipped samples (incl. non-finite locations): all-zero weights (no NaN leaks), both rows -> the zero region
                const float lw = inside ? w_im - wf : 0.f;
                const float ahh = inside ? a * (1.f - lh) : 0.f, alh = inside ? a * lh : 0.f;
                const float hw = 1.f - lw;
                const float w00 = ahh * hw, w01 = ahh * lw, w10 = alh * hw, w11 = alh * lw;
                const uint32_t rx = (uint32_t)(w_low - (int)la.z), ry = (uint32_t)(h_low - (int)la.w);
                const bool inwin = rx < lb.x && ry < lb.y;
                const bool glob = inside && !inwin;
                uint32_t a0 = smem_base, a1 = smem_base, g0 = 0u, g1 = 0u;
                if (inside && inwin) {
                    a0 = lb.z + ry * lb.w + rx * (uint32_t)ROWB;
                    a1 = a0 + lb.w;
                }
                if (glob) {
                    const uint4 lc = lds128(lvaddr + 32);        // W, H, start
                    const int W = (int)lc.x, H = (int)lc.y;
                    a0 = kWinGlobalFlag | (h_low >= 0 ? 1u : 0u) | (h_low + 1 <= H - 1 ? 2u : 0u) | (w_low >= 0 ? 4u : 0u) |
                         (w_low + 1 <= W - 1 ? 8u : 0u);
                    g0 = (uint32_t)((((n * p.S) + (int)lc.z + h_low * W + w_low) * p.M + m) * UPR);
                    g1 = (uint32_t)(W * p.M * UPR);
                }
                any_global = __any_sync(0xffffffffu, glob);
                const uint32_t ea = tab_base + lane * 32;
                if (BF) {      // e0 = left pixel (a0, a1, w00, w10), e1 = right pixel (a0, a1, w01, w11); g0/g1 only if global
                    sts128(ea, a0, a1, __float_as_uint(w00), __float_as_uint(w10));
                    sts128(ea + 16, glob ? g0 : a0, glob ? g1 : a1, __float_as_uint(w01), __float_as_uint(w11));
                } else {
                    sts128(ea, a0, a1, __float_as_uint(w00), __float_as_uint(w01));
                    sts128(ea + 16, __float_as_uint(w10), __float_as_uint(w11), g0, g1);
                }
            }
            __syncwarp();
            if (!waited) {
                mbar_wait(&bar[b], b ? phase1 : phase0);
                waited = true;
            }
            // ---------------------------------------------------------------- phase 2: 8 lanes per sample
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                float2 acc[BF ? 4 : 2];
#pragma unroll
                for (int v = 0; v < (BF ? 4 : 2); ++v) acc[v] = make_float2(0.f, 0.f);
                auto gather = [&](auto checked) {
                    constexpr bool CHECK = decltype(checked)::value;
#pragma unroll
                    for (int i = 0; i < LK / NG; ++i) {
                        const uint32_t ea = tab_base + (uint32_t)((j * LK + i * NG + grp) * 32);
                        if constexpr (BF) {
                            const int side = lane7 >> 2;
                            const uint4 e = lds128(ea + side * 16);
                            uint4 r0, r1;
                            bool is_g = false;
                            if constexpr (CHECK) {
                                const uint4 e0 = side ? lds128(ea) : e;       // flag + validity live in e0.x
                                is_g = (e0.x & kWinGlobalFlag) != 0;
                                if (is_g) {
                                    const uint2 g = side ? make_uint2(e.x, e.y) : lds64(ea + 16);
                                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                                    const uint4 *bp = reinterpret_cast<const uint4 *>(vbytes) + ((int)g.x + (lane7 & 3) + side * (p.M * UPR));
                                    const bool h0 = e0.x & 1u, h1 = e0.x & 2u, ws = side ? (e0.x & 8u) : (e0.x & 4u);
                                    r0 = (h0 && ws) ? __ldg(bp) : z;
                                    r1 = (h1 && ws) ? __ldg(bp + (int)g.y) : z;
                                }
                            }
                            if (!is_g) {
                                r0 = lds128(e.x + lane7 * 16);
                                r1 = lds128(e.y + lane7 * 16);
                            }
                            const float wa = __uint_as_float(e.z), wb = __uint_as_float(e.w);
                            ffma2(acc[0], wa, r0.x << 16, r0.x & 0xffff0000u);
                            ffma2(acc[1], wa, r0.y << 16, r0.y & 0xffff0000u);
                            ffma2(acc[2], wa, r0.z << 16, r0.z & 0xffff0000u);
                            ffma2(acc[3], wa, r0.w << 16, r0.w & 0xffff0000u);
                            ffma2(acc[0], wb, r1.x << 16, r1.x & 0xffff0000u);
                            ffma2(acc[1], wb, r1.y << 16, r1.y & 0xffff0000u);
                            ffma2(acc[2], wb, r1.z << 16, r1.z & 0xffff0000u);
                            ffma2(acc[3], wb, r1.w << 16, r1.w & 0xffff0000u);
                        } else {
                            const uint4 e0 = lds128(ea);
                            const uint2 e1 = lds64(ea + 16);
                            uint4 t00, t01, t10, t11;
                            bool is_g = false;
                            if constexpr (CHECK) {
                                is_g = (e0.x & kWinGlobalFlag) != 0;
                                if (is_g) {
                                    const uint2 g = lds64(ea + 24);
                                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                                    const uint4 *bp = reinterpret_cast<const uint4 *>(vbytes) + ((int)g.x + lane7);
                                    const int gy = (int)g.y;
                                    const bool h0 = e0.x & 1u, h1 = e0.x & 2u, w0 = e0.x & 4u, w1 = e0.x & 8u;
                                    t00 = (h0 && w0) ? __ldg(bp) : z;
                                    t01 = (h0 && w1) ? __ldg(bp + p.M * UPR) : z;
                                    t10 = (h1 && w0) ? __ldg(bp + gy) : z;
                                    t11 = (h1 && w1) ? __ldg(bp + gy + p.M * UPR) : z;
                                }
                            }
                            if (!is_g) {
                                const uint32_t a0 = e0.x + lane7 * 16, a1 = e0.y + lane7 * 16;
                                t00 = lds128(a0);
                                t01 = lds128(a0 + ROWB);
                                t10 = lds128(a1);
                                t11 = lds128(a1 + ROWB);
                            }
                            const float w00 = __uint_as_float(e0.z), w01 = __uint_as_float(e0.w);
                            const float w10 = __uint_as_float(e1.x), w11 = __uint_as_float(e1.y);
                            ffma2(acc[0], w00, t00.x, t00.y); ffma2(acc[1], w00, t00.z, t00.w);
                            ffma2(acc[0], w01, t01.x, t01.y); ffma2(acc[1], w01, t01.z, t01.w);
                            ffma2(acc[0], w10, t10.x, t10.y); ffma2(acc[1], w10, t10.z, t10.w);
                            ffma2(acc[0], w11, t11.x, t11.y); ffma2(acc[1], w11, t11.z, t11.w);
                        }
                    }
                };
                if (any_global) gather(std::true_type{}); else gather(std::false_type{});
                // pair j of this pass = query pass*PPP + j of the tile
                const int qj = pass * PPP + j;
                const int jy = qy0 + (qj >> tws), jx = qx0 + (qj & ((1 << tws) - 1));
                const bool ok = qj < nq && jy < qH && jx < qW;
                const int obase = ((row0 + jy * qW + jx) * p.M + m) * 32;
                if constexpr (BF) {
                    // lane = grp*8 + side*4 + sub holds channels 8*sub .. 8*sub+7 of its side: sum over grp (bits 4, 3) and
                    // side (bit 2), one channel left per lane: 8*sub + 4*bit4 + 2*bit3 + bit2
                    float a8[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y};
                    const float r = win_transpose_reduce<8>(a8, lane, 16);
                    if (ok) RW::store(p.out, obase + 8 * (lane & 3) + ((lane >> 2) & 1) + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1) * 4, r);
                } else {
                    float a4[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
                    const float r = win_transpose_reduce<4>(a4, lane, 16);
                    if (ok) RW::store(p.out, obase + 4 * lane7 + ((lane >> 3) & 1) + ((lane >> 4) & 1) * 2, r);
                }
            }
            __syncwarp();
        }
        if (!waited) mbar_wait(&bar[b], b ? phase1 : phase0);   // every thread observes every fill
        if (filled) {
            if (b) phase1 ^= 1u; else phase0 ^= 1u;
        }
        __syncthreads();                        // all gathers of this unit are done: its buffer may be refilled
        if (!more) break;
#pragma unroll
        for (int pi = 0; pi < kWinMaxPass; ++pi) cur[pi] = nxt[pi];
        slot = slot_n;
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
        // tw stays a power of two (the kernel splits a tile-local query index with a shift); a tile may overhang the level
        const int th = tt[0] < H[lq] ? tt[0] : H[lq], tw = tt[1];
        if (tw > 1 && tw / 2 >= W[lq]) continue;                          // more than half of every tile would be empty
        const double nq = (double)th * (tw < W[lq] ? tw : W[lq]);
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
            const double cost = (staged_bytes + 65536.0) / nq + unstaged * K * 4 * rowb * 4.0;
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
    const size_t budget = 103 * 1024;             // per window buffer; two buffers + zero rows + 17.5 KB static < 227 KB
    size_t smem = 0;
    int units = 0;
    static const bool debug = getenv("VNEXT_WIN_DEBUG") != nullptr;
    for (int lq = 0; lq < L; ++lq) {
        WinPlanChoice ch;
        if (!plan_class(lq, L, K, H, W, rowb, margin, budget, &ch)) {
            set_error("msda window kernel: no tile plan for level %d", lq);
            return VNEXT_EINVAL;
        }
        WinClass &c = p.cls[lq];
        c.th = ch.th; c.tw = ch.tw;
        c.tw_shift = 0;
        while ((1 << c.tw_shift) < ch.tw) ++c.tw_shift;
        c.tiles_x = (W[lq] + ch.tw - 1) / ch.tw;
        c.tiles_y = (H[lq] + ch.th - 1) / ch.th;
        c.unit_begin = units;
        c.H = H[lq]; c.W = W[lq]; c.start = start[lq];
        units += N * c.tiles_x * c.tiles_y * M;
        if (ch.smem > smem) smem = ch.smem;
        if (debug)
            fprintf(stderr, "[msda window] query level %d (%dx%d): tile %dx%d, %d units, staged levels {%d,%d,%d,%d}, %zu B/buffer\n", lq,
                    H[lq], W[lq], ch.th, ch.tw, N * c.tiles_x * c.tiles_y * M, ch.lv[0].staged, L > 1 ? ch.lv[1].staged : -1,
                    L > 2 ? ch.lv[2].staged : -1, L > 3 ? ch.lv[3].staged : -1, ch.smem);
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
    p.k_shift = 0;
    while ((1 << p.k_shift) < K) ++p.k_shift;
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
