// Multi-scale deformable attention — forward kernels for sm_100a.
//
// Replaces the reference's ms_deformable_im2col_gpu_kernel (projects/IDOL/idol/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:237-299, one thread per output element, every thread re-reading the
// (x, y, w) triples and the int64 shape tensors from global memory).
//
// B200 design ("tile" kernels, D = 32 channels per head):
//   * the work unit is a (query, head) PAIR; pairs are contiguous in sampling_loc / attn_weight, so a
//     CTA's tile of TP pairs is ONE contiguous byte range of each tensor.  A single elected thread
//     stages it into shared memory with 1-D TMA (cp.async.bulk, SASS UBLKCP) on an mbarrier, two
//     stages deep, while the warps chew the previous tile;
//   * phase 1: each LANE owns one sample (32 samples = 32/LK pairs per warp pass) and computes its
//     level, the reference's in-range test, the four corner offsets and the four attention-scaled
//     bilinear weights ONCE (the reference recomputes them in each of the D channel threads);
//   * phase 2: a corner row of D=32 channels is 128 B (fp32) / 64 B (bf16).  LPC = 8 (4) lanes each
//     take a 16-byte piece, so one warp-wide 16-byte load fetches NG = 4 (8) corner rows — 4 (8)
//     full cache lines per instruction and no redundant bytes.  Offsets/weights reach the gathering
//     lanes with two warp shuffles.  Every lane accumulates into its own registers across all the
//     samples of the pair; the NG partial sums are folded once per pair by a transposing shuffle
//     reduction that leaves exactly one output channel per lane (one coalesced row store).
//   * persistent grid (resident CTAs only), static round-robin over tiles.
// Anything else (fp64, D != 32, L*K not in {4,8,16,32}, misaligned pointers) takes the generic kernel.
#include "msda_device.cuh"

namespace vnext {

struct MsdaFwdParams {
    const void *value;
    const int64_t *shapes;
    const int64_t *lsi;
    const float *loc;
    const float *attn;
    void *out;
    int S, M, L, K, Lq;
    int pairs;      // N * Lq * M
    int TP;         // pairs per tile
    int num_tiles;
    int use_tma;
};

constexpr int kFwdWarps = 4;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kFwdStages = 2;

template <typename VT>
struct RowVec;  // how one 16-byte piece of a corner row is loaded and widened to fp32
template <>
struct RowVec<float> {
    static constexpr int VEC = 4;   // channels per lane
    static constexpr int LPC = 8;   // lanes per corner row (32 channels)
    __device__ static __forceinline__ void load(const void *base, int off, int sub, bool ok, float (&v)[4]) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) t = __ldg(reinterpret_cast<const float4 *>(base) + off + sub);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static __forceinline__ void store(void *out, long long idx, float v) {
        reinterpret_cast<float *>(out)[idx] = v;
    }
};
template <>
struct RowVec<__nv_bfloat16> {
    static constexpr int VEC = 8;
    static constexpr int LPC = 4;
    __device__ static __forceinline__ void load(const void *base, int off, int sub, bool ok, float (&v)[8]) {
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (ok) t = __ldg(reinterpret_cast<const uint4 *>(base) + off + sub);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ void store(void *out, long long idx, float v) {
        reinterpret_cast<__nv_bfloat16 *>(out)[idx] = __float2bfloat16_rn(v);
    }
};

// Fold VEC partial sums held by each of NG = VEC lane groups into one value per lane.
// On return lane (grp = lane / LPC, sub = lane % LPC) holds channel VEC*sub + grp in a[0].
template <int VEC>
__device__ __forceinline__ float transpose_reduce(float (&a)[VEC], int lane) {
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

template <typename VT, int LK>
__global__ void __launch_bounds__(kFwdThreads, 4) msda_fwd_tile_kernel(const MsdaFwdParams p) {
    using RV = RowVec<VT>;
    constexpr int VEC = RV::VEC, LPC = RV::LPC, NG = 32 / LPC;
    constexpr int PPP = 32 / LK;         // pairs per 32-sample warp pass
    constexpr int UPR = 32 / VEC;        // 16-byte units per (pixel, head) row
    static_assert(LK % NG == 0 && 32 % LK == 0, "unsupported L*K for the tile kernel");

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[kFwdStages];
    __shared__ LevelMeta meta[kMaxLevels];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane / LPC, sub = lane % LPC;
    const int stage_bytes = p.TP * LK * 12;

    load_level_meta(meta, p.shapes, p.lsi, p.L);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto stage_loc = [&](int s) { return reinterpret_cast<float2 *>(dyn_smem + s * stage_bytes); };
    auto stage_attn = [&](int s) { return reinterpret_cast<float *>(dyn_smem + s * stage_bytes + p.TP * LK * 8); };
    auto issue = [&](int tile, int s) {   // one thread
        const int here = min(p.TP, p.pairs - tile * p.TP);
        const uint32_t nloc = (uint32_t)here * LK * 8, nattn = (uint32_t)here * LK * 4;
        mbar_expect_tx(&full_bar[s], nloc + nattn);
        tma_load_1d(stage_loc(s), p.loc + (size_t)tile * p.TP * LK * 2, nloc, &full_bar[s]);
        tma_load_1d(stage_attn(s), p.attn + (size_t)tile * p.TP * LK, nattn, &full_bar[s]);
    };

    if (p.use_tma && tid == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < p.num_tiles) issue(t, s);
        }
    }

    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kFwdStages;
        const int here = min(p.TP, p.pairs - tile * p.TP);
        const float2 *sl = stage_loc(s);
        const float *sa = stage_attn(s);
        if (p.use_tma) {
            mbar_wait(&full_bar[s], (it / kFwdStages) & 1);
        } else {   // pointers not 16-B aligned: plain cooperative staging
            float *dl = reinterpret_cast<float *>(stage_loc(s));
            float *da = stage_attn(s);
            const float *gl = p.loc + (size_t)tile * p.TP * LK * 2;
            const float *ga = p.attn + (size_t)tile * p.TP * LK;
            for (int i = tid; i < here * LK * 2; i += kFwdThreads) dl[i] = gl[i];
            for (int i = tid; i < here * LK; i += kFwdThreads) da[i] = ga[i];
            __syncthreads();
        }

        const int nsamples = here * LK;
        for (int base_idx = warp * 32; base_idx < nsamples; base_idx += kFwdWarps * 32) {
            // ---- phase 1: one sample per lane
            const int idx = base_idx + lane;
            const int pl = idx / LK, smp = idx % LK;
            int off[4] = {-1, -1, -1, -1};
            float wt[4] = {0.f, 0.f, 0.f, 0.f};
            if (idx < nsamples) {
                const float2 xy = sl[idx];
                const float a = sa[idx];
                const int gp = tile * p.TP + pl;
                const int m = gp % p.M, n = (gp / p.M) / p.Lq;
                const LevelMeta lm = meta[smp / p.K];
                const int base = ((n * p.S + lm.start) * p.M + m) * UPR;
                const SamplePrep sp = prep_sample(xy.x, xy.y, lm.H, lm.W, base, p.M * UPR);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    off[r] = sp.off[r];
                    wt[r] = sp.w[r] * a;
                }
            }
            // ---- phase 2: gather NG corner rows per load, accumulate per pair
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                float acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
                for (int i = 0; i < LK / NG; ++i) {
                    const int src = j * LK + i * NG + grp;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int o = __shfl_sync(0xffffffffu, off[r], src);
                        const float w = __shfl_sync(0xffffffffu, wt[r], src);
                        float v[VEC];
                        RV::load(p.value, o, sub, o >= 0, v);
#pragma unroll
                        for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w, v[c], acc[c]);
                    }
                }
                const float r0 = transpose_reduce<VEC>(acc, lane);
                const int pj = base_idx / LK + j;   // pair index inside the tile
                if (pj < here) RV::store(p.out, (long long)(tile * p.TP + pj) * 32 + VEC * sub + grp, r0);
            }
        }
        __syncthreads();   // everyone is done reading stage s
        if (p.use_tma && tid == 0) {
            const int next = tile + kFwdStages * gridDim.x;
            if (next < p.num_tiles) {
                fence_proxy_async();
                issue(next, s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic forward: any dtype / channel count / L*K.  One thread per (pair, channel); the 32 lanes of
// a warp walk consecutive channels, so each corner read is still a coalesced row segment.
template <typename T, typename LT, typename AT>
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const T *__restrict__ value,
                                                               const int64_t *__restrict__ shapes,
                                                               const int64_t *__restrict__ lsi,
                                                               const LT *__restrict__ loc,
                                                               const LT *__restrict__ attn, T *__restrict__ out,
                                                               int S, int M, int D, int L, int K, int Lq,
                                                               long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % D);
        const long long pair = e / D;
        const int m = (int)(pair % M);
        const long long n = pair / M / Lq;
        const LT *lc = loc + pair * L * K * 2;
        const LT *aw = attn + pair * L * K;
        AT col = 0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const T *lvl = value + ((n * S + (int)lsi[l]) * M + m) * (long long)D + c;
            const long long ws = (long long)M * D, hs = (long long)W * ws;
            for (int k = 0; k < K; ++k) {
                const AT x = (AT)lc[(l * K + k) * 2], y = (AT)lc[(l * K + k) * 2 + 1];
                const AT a = (AT)aw[l * K + k];
                AT h_im, w_im;
                if (sizeof(AT) == 4) {
                    h_im = (AT)__fsub_rn(__fmul_rn((float)y, (float)H), 0.5f);
                    w_im = (AT)__fsub_rn(__fmul_rn((float)x, (float)W), 0.5f);
                } else {
                    h_im = (AT)__dsub_rn(__dmul_rn((double)y, (double)H), 0.5);
                    w_im = (AT)__dsub_rn(__dmul_rn((double)x, (double)W), 0.5);
                }
                if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;
                const AT hf = floor(h_im), wf = floor(w_im);
                const int h_low = (int)hf, w_low = (int)wf;
                const AT lh = h_im - hf, lw = w_im - wf, hh = 1 - lh, hw = 1 - lw;
                AT v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                if (h_low >= 0 && w_low >= 0) v1 = (AT)lvl[h_low * hs + w_low * ws];
                if (h_low >= 0 && w_low + 1 <= W - 1) v2 = (AT)lvl[h_low * hs + (w_low + 1) * ws];
                if (h_low + 1 <= H - 1 && w_low >= 0) v3 = (AT)lvl[(h_low + 1) * hs + w_low * ws];
                if (h_low + 1 <= H - 1 && w_low + 1 <= W - 1) v4 = (AT)lvl[(h_low + 1) * hs + (w_low + 1) * ws];
                col += (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * a;
            }
        }
        out[e] = (T)col;
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename VT, int LK>
static int launch_tile(const MsdaFwdParams &p0, cudaStream_t stream) {
    MsdaFwdParams p = p0;
    constexpr int PPP = 32 / LK;
    const int min_tp = kFwdWarps * PPP;
    int tp = 64;
    while (tp > min_tp && (p.pairs + tp - 1) / tp < 8 * num_sms()) tp >>= 1;
    if (tp < min_tp) tp = min_tp;
    p.TP = tp;
    p.num_tiles = (p.pairs + tp - 1) / tp;
    const size_t smem = (size_t)kFwdStages * tp * LK * 12;
    auto kern = msda_fwd_tile_kernel<VT, LK>;
    static int max_blocks = 0;   // per instantiation
    if (max_blocks == 0) {
        VNEXT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        int occ = 0;
        VNEXT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFwdThreads, 2 * 64 * 32 * 12));
        max_blocks = (occ > 0 ? occ : 1);
    }
    int occ_now = 0;
    VNEXT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_now, kern, kFwdThreads, smem));
    if (occ_now < 1) occ_now = 1;
    const int grid = min(p.num_tiles, num_sms() * occ_now);
    kern<<<grid, kFwdThreads, smem, stream>>>(p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

template <typename VT>
static int dispatch_tile(const MsdaFwdParams &p, int LK, cudaStream_t stream, bool *handled) {
    *handled = true;
    switch (LK) {
        case 8: return launch_tile<VT, 8>(p, stream);
        case 16: return launch_tile<VT, 16>(p, stream);
        case 32: return launch_tile<VT, 32>(p, stream);
        case 4:
            if (RowVec<VT>::LPC == 8) return launch_tile<float, 4>(p, stream);
            break;
        default: break;
    }
    *handled = false;
    return VNEXT_OK;
}

int msda_forward_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                        const void *attn, int N, int S, int M, int D, int L, int Lq, int K, void *out,
                        cudaStream_t stream) {
    const long long pairs = (long long)N * Lq * M;
    const long long total = pairs * D;
    if (total == 0) {
        set_variant("fwd_empty");
        return VNEXT_OK;
    }
    const bool small_index = pairs < (1ll << 30) && (long long)N * S * M * D < (1ll << 31) && L <= kMaxLevels;
    if ((dtype == VNEXT_F32 || dtype == VNEXT_BF16) && D == 32 && small_index && aligned16(value) &&
        (reinterpret_cast<uintptr_t>(loc) & 7u) == 0) {
        MsdaFwdParams p;
        p.value = value; p.shapes = shapes; p.lsi = lsi;
        p.loc = static_cast<const float *>(loc); p.attn = static_cast<const float *>(attn);
        p.out = out; p.S = S; p.M = M; p.L = L; p.K = K; p.Lq = Lq; p.pairs = (int)pairs;
        p.TP = 0; p.num_tiles = 0;
        p.use_tma = aligned16(loc) && aligned16(attn);
        bool handled = false;
        int rc = (dtype == VNEXT_F32) ? dispatch_tile<float>(p, L * K, stream, &handled)
                                      : dispatch_tile<__nv_bfloat16>(p, L * K, stream, &handled);
        if (handled) {
            set_variant(dtype == VNEXT_F32 ? (p.use_tma ? "fwd_f32_d32_tma" : "fwd_f32_d32_ldg")
                                           : (p.use_tma ? "fwd_bf16_d32_tma" : "fwd_bf16_d32_ldg"));
            return rc;
        }
    }
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case VNEXT_F32:
            msda_fwd_generic_kernel<float, float, float><<<(int)blocks, threads, 0, stream>>>(
                (const float *)value, shapes, lsi, (const float *)loc, (const float *)attn, (float *)out, S, M, D,
                L, K, Lq, total);
            break;
        case VNEXT_F64:
            msda_fwd_generic_kernel<double, double, double><<<(int)blocks, threads, 0, stream>>>(
                (const double *)value, shapes, lsi, (const double *)loc, (const double *)attn, (double *)out, S, M,
                D, L, K, Lq, total);
            break;
        case VNEXT_BF16:
            msda_fwd_generic_kernel<__nv_bfloat16, float, float><<<(int)blocks, threads, 0, stream>>>(
                (const __nv_bfloat16 *)value, shapes, lsi, (const float *)loc, (const float *)attn,
                (__nv_bfloat16 *)out, S, M, D, L, K, Lq, total);
            break;
        default:
            set_error("vnext_msda_forward: unsupported dtype %d", dtype);
            return VNEXT_EINVAL;
    }
    count_launch();
    set_variant("fwd_generic");
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

}  // namespace vnext
