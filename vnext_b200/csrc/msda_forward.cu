// Multi-scale deformable attention — forward kernels for sm_100a.
//
// Replaces the reference's ms_deformable_im2col_gpu_kernel (projects/IDOL/idol/models/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:237-299, one thread per output element, every thread re-reading the
// (x, y, w) triples and the int64 shape tensors from global memory) and, in its fused form, the ATen
// passes the module wraps around it (softmax over L*K, sampling-location arithmetic;
// ops/modules/ms_deform_attn.py:98-108).
//
// B200 design ("tile" kernels, D = 32 channels per head):
//   * the work unit is a (query, head) PAIR; pairs are contiguous in sampling_loc / attn_weight (and in the
//     raw offset|logit projection), so a CTA's tile of TP pairs is ONE contiguous byte range.  A single
//     elected thread stages it into shared memory with 1-D TMA (cp.async.bulk, SASS UBLKCP) on an
//     mbarrier, two stages deep, while the warps chew the previous tile;
//   * phase 1: each LANE owns one sample (32 samples = 32/LK pairs per warp pass).  It computes — once,
//     not once per channel thread as the reference does — the level, the reference's in-range test, the
//     four corner addresses and the four attention-scaled bilinear weights, and parks them in a per-warp
//     shared-memory table (16 B per corner).  Fused front-end: the lane also forms its sampling location
//     from the reference point and the raw offset, and the softmax over the pair's L*K logits is a
//     shuffle reduction across the pair's lanes;
//   * phase 2: a corner row of D=32 channels is 128 B (fp32) / 64 B (bf16).  LPC = 8 (4) lanes each take a
//     16-byte piece, so one warp-wide 16-byte load fetches NG = 4 (8) corner rows — full cache lines, no
//     redundant bytes.  The inner loop is: one LDS.128 (address + weight, broadcast per lane group), one
//     LDG.128, four (eight) FFMA.  Out-of-range corners point at a zero row (weight untouched), skipped
//     samples have zero weights: no predicates, no divergence, and no value is read that the reference
//     would not read, other than zeros;
//   * every lane accumulates into its own registers across all samples of the pair; the NG partial sums
//     are folded once per pair by a transposing shuffle reduction that leaves exactly one output channel
//     per lane (one coalesced row store);
//   * persistent grid (resident CTAs only), static round-robin over tiles.
// Anything else (fp64, D != 32, L*K not in {4,8,16,32}, misaligned pointers) takes the generic kernel.
#include "msda_device.cuh"

namespace vnext {

__device__ __align__(256) float g_zero_row[64];   // target of out-of-range corners (never written)

struct FastDiv {       // exact n / d for 0 <= n < 2^31 (Granlund-Montgomery, 32-bit magic)
    uint32_t mul, shr, d;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    f.shr = 31 + l;
    f.mul = (uint32_t)(((1ull << f.shr) + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) {
    return (uint32_t)(((uint64_t)n * f.mul) >> f.shr);
}

struct MsdaFwdParams {
    const void *value;
    const int64_t *shapes;
    const int64_t *lsi;
    // classic front-end: sampling locations + normalised attention weights
    const float *loc;
    const float *attn;
    // fused front-end: raw projection rows [N*Lq, oa_pitch] = [offsets (M*L*K*2) | logits (M*L*K)], reference points
    const float *oa;
    int oa_pitch, n_off;
    const float *ref;          // [ref_batch, Lq, L, ref_dim]
    int ref_dim;               // 2 or 4
    long long ref_batch_stride;   // 0 when one set of reference points serves every batch element
    void *out;
    int S, M, L, K, Lq;
    int pairs;      // N * Lq * M
    int TP;         // pairs per tile
    int num_tiles;
    int use_tma;
    FastDiv divM, divLq;
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
    __device__ static __forceinline__ void load(const unsigned char *addr, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(addr));
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
    __device__ static __forceinline__ void load(const unsigned char *addr, float (&v)[8]) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(addr));
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

template <typename VT, int LK, bool FUSED>
__global__ void __launch_bounds__(kFwdThreads, 8) msda_fwd_tile_kernel(const MsdaFwdParams p) {
    using RV = RowVec<VT>;
    constexpr int VEC = RV::VEC, LPC = RV::LPC, NG = 32 / LPC;
    constexpr int PPP = 32 / LK;         // pairs per 32-sample warp pass
    constexpr int UPR = 32 / VEC;        // 16-byte units per (pixel, head) row
    static_assert(LK % NG == 0 && 32 % LK == 0, "unsupported L*K for the tile kernel");

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[kFwdStages];
    __shared__ LevelMeta meta[kMaxLevels];
    __shared__ __align__(16) uint4 table[kFwdWarps][4][32];   // [warp][corner][sample] = {addr lo, addr hi, weight, -}

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane / LPC, sub = lane % LPC;
    // bytes per stage: classic = TP*LK*(8 + 4); fused = (TP / M) rows of oa_pitch floats
    const int stage_bytes = FUSED ? (p.TP / p.M) * p.oa_pitch * 4 : p.TP * LK * 12;

    load_level_meta(meta, p.shapes, p.lsi, p.L);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto stage_ptr = [&](int s) { return dyn_smem + (size_t)s * stage_bytes; };
    auto issue = [&](int tile, int s) {   // one thread
        const int here = min(p.TP, p.pairs - tile * p.TP);
        if (FUSED) {
            const uint32_t nb = (uint32_t)(here / p.M) * p.oa_pitch * 4;
            mbar_expect_tx(&full_bar[s], nb);
            tma_load_1d(stage_ptr(s), p.oa + (size_t)tile * (p.TP / p.M) * p.oa_pitch, nb, &full_bar[s]);
        } else {
            const uint32_t nloc = (uint32_t)here * LK * 8, nattn = (uint32_t)here * LK * 4;
            mbar_expect_tx(&full_bar[s], nloc + nattn);
            tma_load_1d(stage_ptr(s), p.loc + (size_t)tile * p.TP * LK * 2, nloc, &full_bar[s]);
            tma_load_1d(stage_ptr(s) + p.TP * LK * 8, p.attn + (size_t)tile * p.TP * LK, nattn, &full_bar[s]);
        }
    };

    if (p.use_tma && tid == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < p.num_tiles) issue(t, s);
        }
    }

    const unsigned char *vbase = reinterpret_cast<const unsigned char *>(p.value) + sub * 16;
    const unsigned char *zero_row = reinterpret_cast<const unsigned char *>(g_zero_row) + sub * 16;
    (void)zero_row;

    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it % kFwdStages;
        const int here = min(p.TP, p.pairs - tile * p.TP);
        const float *stg = reinterpret_cast<const float *>(stage_ptr(s));
        if (p.use_tma) {
            mbar_wait(&full_bar[s], (it / kFwdStages) & 1);
        } else {   // pointers not 16-B aligned: plain cooperative staging
            float *dst = reinterpret_cast<float *>(stage_ptr(s));
            if (FUSED) {
                const float *g = p.oa + (size_t)tile * (p.TP / p.M) * p.oa_pitch;
                for (int i = tid; i < (here / p.M) * p.oa_pitch; i += kFwdThreads) dst[i] = g[i];
            } else {
                const float *gl = p.loc + (size_t)tile * p.TP * LK * 2;
                const float *ga = p.attn + (size_t)tile * p.TP * LK;
                for (int i = tid; i < here * LK * 2; i += kFwdThreads) dst[i] = gl[i];
                for (int i = tid; i < here * LK; i += kFwdThreads) dst[p.TP * LK * 2 + i] = ga[i];
            }
            __syncthreads();
        }

        const int nsamples = here * LK;
        for (int base_idx = warp * 32; base_idx < nsamples; base_idx += kFwdWarps * 32) {
            // ---- phase 1: one sample per lane -> table of (corner address, weight)
            const int idx = base_idx + lane;
            const int pl = idx / LK, smp = idx % LK;
            const bool live = idx < nsamples;
            float x = 0.f, y = 0.f, a = 0.f;
            int base = 0;
            LevelMeta lm = meta[0];
            if (live) {
                const uint32_t gp = (uint32_t)(tile * p.TP + pl);
                const uint32_t row = fdiv(gp, p.divM);            // global query row = n * Lq + q
                const uint32_t m = gp - row * p.M;
                const uint32_t n = fdiv(row, p.divLq);
                const int lvl = smp / p.K;
                lm = meta[lvl];
                base = ((n * p.S + lm.start) * p.M + m) * UPR;
                if (FUSED) {
                    const float *rowp = stg + (pl / p.M) * p.oa_pitch;
                    const float2 off = *reinterpret_cast<const float2 *>(rowp + (m * LK + smp) * 2);
                    a = rowp[p.n_off + m * LK + smp];             // raw logit; softmax below
                    const uint32_t q = row - n * p.Lq;
                    const float *rp = p.ref + n * p.ref_batch_stride + ((size_t)q * p.L + lvl) * p.ref_dim;
                    if (p.ref_dim == 2) {
                        const float2 r = *reinterpret_cast<const float2 *>(rp);
                        x = r.x + off.x / (float)lm.W;            // ref + offset / (W_l, H_l)   (module :102-105)
                        y = r.y + off.y / (float)lm.H;
                    } else {
                        const float4 r = *reinterpret_cast<const float4 *>(rp);
                        x = r.x + off.x / (float)p.K * r.z * 0.5f;   // ref_xy + offset / P * ref_wh * 0.5 (:106-108)
                        y = r.y + off.y / (float)p.K * r.w * 0.5f;
                    }
                } else {
                    const float2 xy = reinterpret_cast<const float2 *>(stg)[idx];
                    x = xy.x; y = xy.y;
                    a = stg[p.TP * LK * 2 + idx];
                }
            }
            if (FUSED) {   // softmax over the LK logits of the pair = the LK lanes of this lane's aligned group
                float mx = live ? a : -INFINITY;
#pragma unroll
                for (int o = LK / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float e = live ? expf(a - mx) : 0.f;
                float sum = e;
#pragma unroll
                for (int o = LK / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                a = live ? e / sum : 0.f;
            }
            {
                const SamplePrep sp = prep_sample(x, y, lm.H, lm.W, base, p.M * UPR);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool ok = live && sp.off[r] >= 0;
                    const unsigned char *addr = ok ? reinterpret_cast<const unsigned char *>(p.value) + (size_t)sp.off[r] * 16
                                                   : reinterpret_cast<const unsigned char *>(g_zero_row);
                    const uint64_t a64 = reinterpret_cast<uint64_t>(addr);
                    const float w = live ? sp.w[r] * a : 0.f;
                    table[warp][r][lane] = make_uint4((uint32_t)a64, (uint32_t)(a64 >> 32), __float_as_uint(w), 0u);
                }
            }
            __syncwarp();
            // ---- phase 2: gather NG corner rows per load, accumulate per pair
#pragma unroll
            for (int j = 0; j < PPP; ++j) {
                float acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
                for (int i = 0; i < LK / NG; ++i) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint4 e = table[warp][r][j * LK + i * NG + grp];
                        const unsigned char *addr =
                            reinterpret_cast<const unsigned char *>(((uint64_t)e.y << 32) | e.x) + sub * 16;
                        const float w = __uint_as_float(e.z);
                        float v[VEC];
                        RV::load(addr, v);
#pragma unroll
                        for (int c = 0; c < VEC; ++c) acc[c] = fmaf(w, v[c], acc[c]);
                    }
                }
                const float r0 = transpose_reduce<VEC>(acc, lane);
                const int pj = base_idx / LK + j;   // pair index inside the tile
                if (pj < here) RV::store(p.out, (long long)(tile * p.TP + pj) * 32 + VEC * sub + grp, r0);
            }
            __syncwarp();
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
    (void)vbase;
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

template <typename VT, int LK, bool FUSED>
static int launch_tile(const MsdaFwdParams &p0, cudaStream_t stream) {
    MsdaFwdParams p = p0;
    constexpr int PPP = 32 / LK;
    int min_tp = kFwdWarps * PPP;
    if (FUSED) {   // whole projection rows per tile: TP must be a multiple of M
        while (min_tp % p.M != 0) min_tp += kFwdWarps * PPP;
    }
    int tp = 32;          // 2 stages x 6 KB staging + 8.6 KB static: 8 CTAs (32 warps) per SM
    if (FUSED) tp = ((32 + p.M - 1) / p.M) * p.M;
    while (tp / 2 >= min_tp && (tp / 2) % min_tp == 0 && (p.pairs + tp - 1) / tp < 8 * num_sms()) tp >>= 1;
    if (tp < min_tp || tp % min_tp != 0) tp = min_tp;
    p.TP = tp;
    p.num_tiles = (p.pairs + tp - 1) / tp;
    const size_t smem = (size_t)kFwdStages * (FUSED ? (size_t)(tp / p.M) * p.oa_pitch * 4 : (size_t)tp * LK * 12);
    auto kern = msda_fwd_tile_kernel<VT, LK, FUSED>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) {
            set_error("msda forward: tile staging of %zu bytes does not fit shared memory", smem);
            return VNEXT_EINVAL;
        }
        VNEXT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

template <typename VT, bool FUSED>
static int dispatch_tile(const MsdaFwdParams &p, int LK, cudaStream_t stream, bool *handled) {
    *handled = true;
    switch (LK) {
        case 8: return launch_tile<VT, 8, FUSED>(p, stream);
        case 16: return launch_tile<VT, 16, FUSED>(p, stream);
        case 32: return launch_tile<VT, 32, FUSED>(p, stream);
        case 4:
            if (RowVec<VT>::LPC == 8) return launch_tile<float, 4, FUSED>(p, stream);
            break;
        default: break;
    }
    *handled = false;
    return VNEXT_OK;
}

static void fill_common(MsdaFwdParams &p, const void *value, const int64_t *shapes, const int64_t *lsi, void *out, int S,
                        int M, int L, int K, int Lq, long long pairs) {
    p.value = value; p.shapes = shapes; p.lsi = lsi; p.out = out;
    p.S = S; p.M = M; p.L = L; p.K = K; p.Lq = Lq; p.pairs = (int)pairs;
    p.TP = 0; p.num_tiles = 0;
    p.loc = p.attn = p.oa = p.ref = nullptr;
    p.oa_pitch = p.n_off = p.ref_dim = 0; p.ref_batch_stride = 0;
    p.divM = make_fastdiv((uint32_t)M);
    p.divLq = make_fastdiv((uint32_t)(Lq > 0 ? Lq : 1));
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
        fill_common(p, value, shapes, lsi, out, S, M, L, K, Lq, pairs);
        p.loc = static_cast<const float *>(loc); p.attn = static_cast<const float *>(attn);
        p.use_tma = aligned16(loc) && aligned16(attn);
        bool handled = false;
        int rc = (dtype == VNEXT_F32) ? dispatch_tile<float, false>(p, L * K, stream, &handled)
                                      : dispatch_tile<__nv_bfloat16, false>(p, L * K, stream, &handled);
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

// Fused front-end: softmax + sampling-location arithmetic + sampling in one kernel (tile kernels only).
int msda_forward_fused_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const float *oa,
                              int oa_pitch, int n_off, const float *ref, int ref_dim, int ref_shared, int N, int S,
                              int M, int D, int L, int Lq, int K, void *out, cudaStream_t stream) {
    const long long pairs = (long long)N * Lq * M;
    if (pairs == 0) {
        set_variant("fwd_empty");
        return VNEXT_OK;
    }
    const int LK = L * K;
    const bool small_index = pairs < (1ll << 30) && (long long)N * S * M * D < (1ll << 31) && L <= kMaxLevels;
    if (!((dtype == VNEXT_F32 || dtype == VNEXT_BF16) && D == 32 && small_index && aligned16(value) && aligned16(oa) &&
          oa_pitch % 4 == 0 && aligned16(ref) && (ref_dim == 2 || ref_dim == 4) && n_off == M * LK * 2 &&
          oa_pitch >= M * LK * 3 && (LK == 8 || LK == 16 || LK == 32 || (LK == 4 && dtype == VNEXT_F32)))) {
        set_error("vnext_msda_forward_fused: unsupported configuration (needs D == 32, L*K in {4,8,16,32}, fp32/bf16 value, "
                  "16-byte aligned value/oa/ref, oa_pitch %% 4 == 0, ref_dim 2|4); got D=%d L*K=%d oa_pitch=%d ref_dim=%d",
                  D, LK, oa_pitch, ref_dim);
        return VNEXT_EINVAL;
    }
    MsdaFwdParams p;
    fill_common(p, value, shapes, lsi, out, S, M, L, K, Lq, pairs);
    p.oa = oa; p.oa_pitch = oa_pitch; p.n_off = n_off; p.ref = ref; p.ref_dim = ref_dim;
    p.ref_batch_stride = ref_shared ? 0 : (long long)Lq * L * ref_dim;
    p.use_tma = 1;
    bool handled = false;
    int rc = (dtype == VNEXT_F32) ? dispatch_tile<float, true>(p, LK, stream, &handled)
                                  : dispatch_tile<__nv_bfloat16, true>(p, LK, stream, &handled);
    set_variant(dtype == VNEXT_F32 ? "fwd_fused_f32_d32_tma" : "fwd_fused_bf16_d32_tma");
    return rc;
}

}  // namespace vnext
