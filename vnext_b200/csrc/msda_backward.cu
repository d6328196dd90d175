// Multi-scale deformable attention — backward kernels for sm_100a.
//
// Replaces the reference's six ms_deformable_col2im_gpu_kernel_* variants (projects/IDOL/idol/models/
// ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920; for D=32 a 32-thread block per (query, head), two
// __syncthreads and a serial 32-way sum per sample point, four scalar atomicAdd per thread per point).
//
// B200 design (fp32, D = 32): same tiling as the forward — (query, head) pairs, TMA-staged
// sampling_loc / attn_weight tiles, one LANE per sample in phase 1, 8 lanes x 16 B per corner row in
// phase 2 — plus:
//   * grad_value: one vectorised `red.global.add.v4.f32` per lane per corner (4 channels / RED instead
//     of 4 scalar atomics);
//   * grad_sampling_loc / grad_attn_weight need, per sample, the four dot products <corner row, grad_out
//     row>.  Each lane forms a 4-channel partial; the 8 partials of a corner are folded with a
//     transposing shuffle reduction over the 4 corners (4 shuffles per 4 dots), parked in a per-warp
//     shared-memory table and read back as one float4 by the lane that OWNS the sample, which then
//     writes its (gx, gy, ga) directly — no atomics, no zero-fill, coalesced stores.
// The summation order of grad_value is atomics-defined (as in the reference); tests use tolerances.
#include "msda_device.cuh"

namespace vnext {

struct MsdaBwdParams {
    const float *value;
    const int64_t *shapes;
    const int64_t *lsi;
    const float *loc;
    const float *attn;
    const float *gout;
    float *gvalue;
    float *gloc;
    float *gattn;
    int S, M, L, K, Lq;
    int pairs, TP, num_tiles, use_tma;
};

constexpr int kBwdWarps = 4;
constexpr int kBwdThreads = kBwdWarps * 32;
constexpr int kBwdStages = 2;

template <int LK>
__global__ void __launch_bounds__(kBwdThreads, 4) msda_bwd_tile_kernel(const MsdaBwdParams p) {
    constexpr int PPP = 32 / LK;
    static_assert(LK % 4 == 0 && 32 % LK == 0, "unsupported L*K for the tile kernel");

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[kBwdStages];
    __shared__ LevelMeta meta[kMaxLevels];
    __shared__ __align__(16) float dots[kBwdWarps][32][4];   // [warp][sample][corner]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, sub = lane & 7;
    const int stage_bytes = p.TP * LK * 12;

    load_level_meta(meta, p.shapes, p.lsi, p.L);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kBwdStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto stage_loc = [&](int s) { return reinterpret_cast<float2 *>(dyn_smem + s * stage_bytes); };
    auto stage_attn = [&](int s) { return reinterpret_cast<float *>(dyn_smem + s * stage_bytes + p.TP * LK * 8); };
    auto issue = [&](int tile, int s) {
        const int here = min(p.TP, p.pairs - tile * p.TP);
        const uint32_t nloc = (uint32_t)here * LK * 8, nattn = (uint32_t)here * LK * 4;
        mbar_expect_tx(&full_bar[s], nloc + nattn);
        tma_load_1d(stage_loc(s), p.loc + (size_t)tile * p.TP * LK * 2, nloc, &full_bar[s]);
        tma_load_1d(stage_attn(s), p.attn + (size_t)tile * p.TP * LK, nattn, &full_bar[s]);
    };
    if (p.use_tma && tid == 0) {
#pragma unroll
        for (int s = 0; s < kBwdStages; ++s) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < p.num_tiles) issue(t, s);
        }
    }

    const float4 *value4 = reinterpret_cast<const float4 *>(p.value);
    const float4 *gout4 = reinterpret_cast<const float4 *>(p.gout);

    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kBwdStages;
        const int here = min(p.TP, p.pairs - tile * p.TP);
        const float2 *sl = stage_loc(s);
        const float *sa = stage_attn(s);
        if (p.use_tma) {
            mbar_wait(&full_bar[s], (it / kBwdStages) & 1);
        } else {
            float *dl = reinterpret_cast<float *>(stage_loc(s));
            float *da = stage_attn(s);
            const float *gl = p.loc + (size_t)tile * p.TP * LK * 2;
            const float *ga = p.attn + (size_t)tile * p.TP * LK;
            for (int i = tid; i < here * LK * 2; i += kBwdThreads) dl[i] = gl[i];
            for (int i = tid; i < here * LK; i += kBwdThreads) da[i] = ga[i];
            __syncthreads();
        }

        const int nsamples = here * LK;
        for (int base_idx = warp * 32; base_idx < nsamples; base_idx += kBwdWarps * 32) {
            // ---- phase 1: one sample per lane
            const int idx = base_idx + lane;
            const int pl = idx / LK, smp = idx % LK;
            int off[4] = {-1, -1, -1, -1};
            float wa[4] = {0.f, 0.f, 0.f, 0.f};      // bilinear weight * attention weight
            float a = 0.f, hh = 0.f, hw = 0.f, lh = 0.f, lw = 0.f, fH = 0.f, fW = 0.f;
            float bw[4] = {0.f, 0.f, 0.f, 0.f};      // bare bilinear weights
            if (idx < nsamples) {
                const float2 xy = sl[idx];
                a = sa[idx];
                const int gp = tile * p.TP + pl;
                const int m = gp % p.M, n = (gp / p.M) / p.Lq;
                const LevelMeta lm = meta[smp / p.K];
                const int base = ((n * p.S + lm.start) * p.M + m) * 8;
                const SamplePrep sp = prep_sample(xy.x, xy.y, lm.H, lm.W, base, p.M * 8);
                fH = (float)lm.H;
                fW = (float)lm.W;
                if (sp.inside) { hh = sp.hh; hw = sp.hw; lh = sp.lh; lw = sp.lw; }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    off[r] = sp.off[r];
                    bw[r] = sp.w[r];
                    wa[r] = sp.w[r] * a;
                }
            }
            // ---- phase 2
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                const int pj = base_idx / LK + j;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pj < here) g = __ldg(gout4 + (size_t)(tile * p.TP + pj) * 8 + sub);
#pragma unroll
                for (int i = 0; i < LK / 4; ++i) {
                    const int src = j * LK + i * 4 + grp;   // the sample this lane group gathers for
                    float part[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int o = __shfl_sync(0xffffffffu, off[r], src);
                        const float w = __shfl_sync(0xffffffffu, wa[r], src);
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (o >= 0) {
                            v = __ldg(value4 + o + sub);
                            red_add_v4(p.gvalue + 4 * (size_t)(o + sub), w * g.x, w * g.y, w * g.z, w * g.w);
                        }
                        part[r] = fmaf(v.x, g.x, fmaf(v.y, g.y, fmaf(v.z, g.z, v.w * g.w)));
                    }
                    // fold the 8 lane partials of each corner: 4 corners x 8 lanes -> transposing reduce
                    {
                        const bool up4 = (lane & 4) != 0;
                        const float k0 = up4 ? part[2] : part[0], s0 = up4 ? part[0] : part[2];
                        const float k1 = up4 ? part[3] : part[1], s1 = up4 ? part[1] : part[3];
                        const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
                        const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
                        const bool up2 = (lane & 2) != 0;
                        const float k = up2 ? a1 : a0, sn = up2 ? a0 : a1;
                        float d = k + __shfl_xor_sync(0xffffffffu, sn, 2);
                        d += __shfl_xor_sync(0xffffffffu, d, 1);
                        // lane now holds the dot of corner r = 2*bit2 + bit1 for sample `src`
                        if ((lane & 1) == 0) dots[warp][src][((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = d;
                    }
                }
            }
            __syncwarp();
            // ---- phase 3: the owning lane finishes its sample
            if (idx < nsamples) {
                const float4 d = *reinterpret_cast<const float4 *>(&dots[warp][lane][0]);
                const float ga = bw[0] * d.x + bw[1] * d.y + bw[2] * d.z + bw[3] * d.w;
                // d(val)/dw_im = hh*(v2 - v1) + lh*(v4 - v3);  d(val)/dh_im = hw*(v3 - v1) + lw*(v4 - v2)
                const float gw = hh * (d.y - d.x) + lh * (d.w - d.z);
                const float gh = hw * (d.z - d.x) + lw * (d.w - d.y);
                const size_t si = (size_t)tile * p.TP * LK + idx;
                reinterpret_cast<float2 *>(p.gloc)[si] = make_float2(fW * gw * a, fH * gh * a);
                p.gattn[si] = ga;
            }
            __syncwarp();
        }
        __syncthreads();
        if (p.use_tma && tid == 0) {
            const int next = tile + kBwdStages * gridDim.x;
            if (next < p.num_tiles) {
                fence_proxy_async();
                issue(next, s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic backward: one thread per (pair, channel); grad_loc / grad_attn via atomics on zero-filled
// outputs (warp-aggregated when a whole warp shares the pair).  VT = storage type of value/grad_out,
// GT = type of the grad_value accumulator buffer, AT = arithmetic type.
template <typename AT>
__device__ __forceinline__ void atomic_add_t(AT *addr, AT v) { atomicAdd(addr, v); }

template <typename VT, typename LT, typename GT, typename AT>
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const VT *__restrict__ value,
                                                               const int64_t *__restrict__ shapes,
                                                               const int64_t *__restrict__ lsi,
                                                               const LT *__restrict__ loc,
                                                               const LT *__restrict__ attn,
                                                               const VT *__restrict__ gout, GT *gvalue, LT *gloc,
                                                               LT *gattn, int S, int M, int D, int L, int K,
                                                               int Lq, long long total) {
    const bool warp_uniform_pair = (D % 32) == 0;   // all 32 lanes of a warp share (pair, sample)
    for (long long e0 = (long long)blockIdx.x * blockDim.x; e0 < total; e0 += (long long)gridDim.x * blockDim.x) {
        const long long e = e0 + threadIdx.x;
        const bool live = e < total;
        const long long ee = live ? e : total - 1;
        const int c = (int)(ee % D);
        const long long pair = ee / D;
        const int m = (int)(pair % M);
        const long long n = pair / M / Lq;
        const AT top = live ? (AT)gout[ee] : (AT)0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const long long lbase = ((n * S + (int)lsi[l]) * M + m) * (long long)D + c;
            const long long ws = (long long)M * D, hs = (long long)W * ws;
            for (int k = 0; k < K; ++k) {
                const long long si = pair * L * K + l * K + k;
                const AT x = (AT)loc[si * 2], y = (AT)loc[si * 2 + 1], a = (AT)attn[si];
                AT h_im, w_im;
                if (sizeof(AT) == 4) {
                    h_im = (AT)__fsub_rn(__fmul_rn((float)y, (float)H), 0.5f);
                    w_im = (AT)__fsub_rn(__fmul_rn((float)x, (float)W), 0.5f);
                } else {
                    h_im = (AT)__dsub_rn(__dmul_rn((double)y, (double)H), 0.5);
                    w_im = (AT)__dsub_rn(__dmul_rn((double)x, (double)W), 0.5);
                }
                AT g_a = 0, g_x = 0, g_y = 0;
                const bool inside = h_im > -1 && w_im > -1 && h_im < H && w_im < W;
                if (inside && live) {
                    const AT hf = floor(h_im), wf = floor(w_im);
                    const int h_low = (int)hf, w_low = (int)wf;
                    const AT lh = h_im - hf, lw = w_im - wf, hh = 1 - lh, hw = 1 - lw;
                    const AT tgv = top * a;
                    AT v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                    if (h_low >= 0 && w_low >= 0) {
                        const long long o = lbase + h_low * hs + w_low * ws;
                        v1 = (AT)value[o];
                        atomic_add_t<GT>(gvalue + o, (GT)(hh * hw * tgv));
                    }
                    if (h_low >= 0 && w_low + 1 <= W - 1) {
                        const long long o = lbase + h_low * hs + (w_low + 1) * ws;
                        v2 = (AT)value[o];
                        atomic_add_t<GT>(gvalue + o, (GT)(hh * lw * tgv));
                    }
                    if (h_low + 1 <= H - 1 && w_low >= 0) {
                        const long long o = lbase + (h_low + 1) * hs + w_low * ws;
                        v3 = (AT)value[o];
                        atomic_add_t<GT>(gvalue + o, (GT)(lh * hw * tgv));
                    }
                    if (h_low + 1 <= H - 1 && w_low + 1 <= W - 1) {
                        const long long o = lbase + (h_low + 1) * hs + (w_low + 1) * ws;
                        v4 = (AT)value[o];
                        atomic_add_t<GT>(gvalue + o, (GT)(lh * lw * tgv));
                    }
                    g_a = top * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
                    g_x = (AT)W * (hh * (v2 - v1) + lh * (v4 - v3)) * tgv;
                    g_y = (AT)H * (hw * (v3 - v1) + lw * (v4 - v2)) * tgv;
                }
                if (warp_uniform_pair) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        g_a += __shfl_xor_sync(0xffffffffu, g_a, o);
                        g_x += __shfl_xor_sync(0xffffffffu, g_x, o);
                        g_y += __shfl_xor_sync(0xffffffffu, g_y, o);
                    }
                    if ((threadIdx.x & 31) == 0 && inside && live) {
                        atomic_add_t<LT>(gloc + si * 2, (LT)g_x);
                        atomic_add_t<LT>(gloc + si * 2 + 1, (LT)g_y);
                        atomic_add_t<LT>(gattn + si, (LT)g_a);
                    }
                } else if (inside && live) {
                    atomic_add_t<LT>(gloc + si * 2, (LT)g_x);
                    atomic_add_t<LT>(gloc + si * 2 + 1, (LT)g_y);
                    atomic_add_t<LT>(gattn + si, (LT)g_a);
                }
            }
        }
    }
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int LK>
static int launch_bwd_tile(const MsdaBwdParams &p0, cudaStream_t stream) {
    MsdaBwdParams p = p0;
    constexpr int PPP = 32 / LK;
    const int min_tp = kBwdWarps * PPP;
    int tp = 64;
    while (tp > min_tp && (p.pairs + tp - 1) / tp < 8 * num_sms()) tp >>= 1;
    while (tp > min_tp && (size_t)kBwdStages * tp * LK * 12 > 40 * 1024) tp >>= 1;   // stay under the 48 KB default with the static part
    if (tp < min_tp) tp = min_tp;
    p.TP = tp;
    p.num_tiles = (p.pairs + tp - 1) / tp;
    const size_t smem = (size_t)kBwdStages * tp * LK * 12;
    auto kern = msda_bwd_tile_kernel<LK>;
    int occ = 0;
    VNEXT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBwdThreads, smem));
    if (occ < 1) occ = 1;
    const int grid = min(p.num_tiles, num_sms() * occ);
    kern<<<grid, kBwdThreads, smem, stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

int msda_backward_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                         const void *attn, const void *gout, int N, int S, int M, int D, int L, int Lq, int K,
                         void *gvalue, void *gloc, void *gattn, void *workspace, cudaStream_t stream) {
    const long long pairs = (long long)N * Lq * M;
    const long long total = pairs * D;
    const long long nvalue = (long long)N * S * M * D;
    const size_t esz = dtype == VNEXT_F64 ? 8 : 4;          // element size of loc / attn (and f32/f64 value)
    const size_t lsz = dtype == VNEXT_F64 ? 8 : 4;
    (void)esz;
    if (dtype == VNEXT_BF16 && workspace == nullptr && nvalue > 0) {
        set_error("vnext_msda_backward: bf16 needs an fp32 workspace of %zu bytes", (size_t)nvalue * 4);
        return VNEXT_EINVAL;
    }
    void *gv_acc = dtype == VNEXT_BF16 ? workspace : gvalue;
    const size_t gv_bytes = (size_t)nvalue * (dtype == VNEXT_F64 ? 8 : 4);
    if (nvalue > 0) {
        VNEXT_CUDA(cudaMemsetAsync(gv_acc, 0, gv_bytes, stream));
        count_launch();
    }
    if (total == 0) {
        set_variant("bwd_empty");
        return VNEXT_OK;
    }
    const bool small_index = pairs < (1ll << 30) && nvalue < (1ll << 31) && L <= kMaxLevels;
    const int LK = L * K;
    if (dtype == VNEXT_F32 && D == 32 && small_index && aligned16(value) && aligned16(gout) && aligned16(gvalue) &&
        (reinterpret_cast<uintptr_t>(loc) & 7u) == 0 && (reinterpret_cast<uintptr_t>(gloc) & 7u) == 0 &&
        (LK == 4 || LK == 8 || LK == 16 || LK == 32)) {
        MsdaBwdParams p;
        p.value = (const float *)value; p.shapes = shapes; p.lsi = lsi;
        p.loc = (const float *)loc; p.attn = (const float *)attn; p.gout = (const float *)gout;
        p.gvalue = (float *)gvalue; p.gloc = (float *)gloc; p.gattn = (float *)gattn;
        p.S = S; p.M = M; p.L = L; p.K = K; p.Lq = Lq; p.pairs = (int)pairs;
        p.TP = 0; p.num_tiles = 0;
        p.use_tma = aligned16(loc) && aligned16(attn);
        set_variant(p.use_tma ? "bwd_f32_d32_tma" : "bwd_f32_d32_ldg");
        switch (LK) {
            case 4: return launch_bwd_tile<4>(p, stream);
            case 8: return launch_bwd_tile<8>(p, stream);
            case 16: return launch_bwd_tile<16>(p, stream);
            default: return launch_bwd_tile<32>(p, stream);
        }
    }
    // generic path: grad_loc / grad_attn are accumulated with atomics -> zero-fill first
    VNEXT_CUDA(cudaMemsetAsync(gloc, 0, (size_t)pairs * LK * 2 * lsz, stream));
    VNEXT_CUDA(cudaMemsetAsync(gattn, 0, (size_t)pairs * LK * lsz, stream));
    count_launch(2);
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case VNEXT_F32:
            msda_bwd_generic_kernel<float, float, float, float><<<(int)blocks, threads, 0, stream>>>(
                (const float *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const float *)gout,
                (float *)gvalue, (float *)gloc, (float *)gattn, S, M, D, L, K, Lq, total);
            break;
        case VNEXT_F64:
            msda_bwd_generic_kernel<double, double, double, double><<<(int)blocks, threads, 0, stream>>>(
                (const double *)value, shapes, lsi, (const double *)loc, (const double *)attn,
                (const double *)gout, (double *)gvalue, (double *)gloc, (double *)gattn, S, M, D, L, K, Lq, total);
            break;
        case VNEXT_BF16: {
            msda_bwd_generic_kernel<__nv_bfloat16, float, float, float><<<(int)blocks, threads, 0, stream>>>(
                (const __nv_bfloat16 *)value, shapes, lsi, (const float *)loc, (const float *)attn,
                (const __nv_bfloat16 *)gout, (float *)workspace, (float *)gloc, (float *)gattn, S, M, D, L, K, Lq,
                total);
            long long cb = (nvalue + 255) / 256;
            if (cb > cap) cb = cap;
            f32_to_bf16_kernel<<<(int)cb, 256, 0, stream>>>((const float *)workspace, (__nv_bfloat16 *)gvalue, nvalue);
            count_launch();
            break;
        }
        default:
            set_error("vnext_msda_backward: unsupported dtype %d", dtype);
            return VNEXT_EINVAL;
    }
    count_launch();
    set_variant("bwd_generic");
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

}  // namespace vnext
