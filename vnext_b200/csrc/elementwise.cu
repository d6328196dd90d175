// Memory-bound epilogue / normalisation kernels of the backbone and input projection (sm_100a), NHWC fp32.
//
// The reference runs these as separate ATen passes after each cuDNN conv (FrozenBN affine, ReLU, residual
// add: detectron2/modeling/backbone/resnet.py:194-210,355-359, detectron2/layers/batch_norm.py:44-65;
// max-pool resnet.py:358; GroupNorm(32) of input_proj: projects/SeqFormer/seqformer/models/
// deformable_detr.py:64-78; pixel normalisation + /32 zero padding: seqformer.py:413-429, util/misc.py:288-316).
// Here each is ONE pass over HBM with 16-byte accesses; grids are sized to the resident-CTA count and
// grid-stride over the tensor.
#include "common.cuh"

namespace vnext {

// out = act(x + bias[c] (+ residual))   rows x C, C % 4 == 0; T = storage type of x / residual / out (fp32 or bf16)
template <typename T>
__global__ void __launch_bounds__(256) bias_act_kernel(const T *__restrict__ x, const float4 *__restrict__ bias,
                                                       const T *__restrict__ residual, T *__restrict__ out,
                                                       long long n4, int c4, int relu) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = load4(x + i * 4);
        const float4 b = __ldg(bias + (int)(i % c4));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (residual) {
            const float4 r = load4(residual + i * 4);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        store4(out + i * 4, v);
    }
}

// Stem tail: out[n, ho, wo, c] = max over the 3x3 / stride 2 / pad 1 window of relu(x + bias[c]).
// relu and +bias are monotone, so this is relu(max(x) + bias): one read of the conv output, one quarter-size write.
template <typename IT, typename OT>
__global__ void __launch_bounds__(256) stem_pool_kernel(const IT *__restrict__ x, const float4 *__restrict__ bias,
                                                        OT *__restrict__ out, int N, int H, int W, int Ho, int Wo,
                                                        int c4) {
    const long long total = (long long)N * Ho * Wo * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        long long t = i / c4;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int h = 2 * ho - 1 + dy;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int w = 2 * wo - 1 + dx;
                if (w < 0 || w >= W) continue;
                const float4 v = load4(x + (((long long)(n * H + h) * W + w) * c4 + c) * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        const float4 b = __ldg(bias + c);
        m.x = fmaxf(m.x + b.x, 0.f); m.y = fmaxf(m.y + b.y, 0.f); m.z = fmaxf(m.z + b.z, 0.f); m.w = fmaxf(m.w + b.w, 0.f);
        store4(out + i * 4, m);
    }
}

// GroupNorm over [N, HW, C] (channels-last), G groups of C/G = 8 channels: partial statistics -> fixed-order combine
// -> apply.  Deterministic by construction (no atomics): every CTA of the statistics pass writes its own (sum, sumsq)
// per group, and ONE thread per (n, group) folds them in block order with Chan's parallel-variance update in fp64, so the
// result does not depend on CTA scheduling (two concurrent replays of one CUDA graph give identical bits) and the
// variance does not suffer the E[x^2] - mean^2 cancellation over the whole level.
constexpr int kGnThreads = 256;
constexpr int kGnRowsPerCta = 64;
template <typename IT>
__global__ void __launch_bounds__(kGnThreads) groupnorm_partial_kernel(const IT *__restrict__ x, float *__restrict__ part,
                                                                       int HW, int C) {
    // thread -> (row offset, float4 column); C/4 float4 columns, two per 8-channel group
    const int c4 = C / 4;
    const int n = blockIdx.y;
    const int col = threadIdx.x % c4, rsub = threadIdx.x / c4, rstep = kGnThreads / c4;
    const int r0 = blockIdx.x * kGnRowsPerCta;
    const int r1 = min(HW, r0 + kGnRowsPerCta);
    float s = 0.f, ss = 0.f;
    for (int r = r0 + rsub; r < r1; r += rstep) {
        const float4 v = load4(x + (((long long)n * HW + r) * c4 + col) * 4);
        s += (v.x + v.y) + (v.z + v.w);
        ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    __shared__ float sh_s[kGnThreads], sh_ss[kGnThreads];
    sh_s[threadIdx.x] = s;
    sh_ss[threadIdx.x] = ss;
    __syncthreads();
    const int G = C / 8;
    if (threadIdx.x < G) {      // group g owns float4 columns 2g and 2g+1 of every row-subset
        float a = 0.f, b = 0.f;
        for (int k = 0; k < rstep; ++k)
            for (int h = 0; h < 2; ++h) {
                a += sh_s[k * c4 + 2 * threadIdx.x + h];
                b += sh_ss[k * c4 + 2 * threadIdx.x + h];
            }
        float *dst = part + (((long long)n * gridDim.x + blockIdx.x) * G + threadIdx.x) * 2;
        dst[0] = a;
        dst[1] = b;
    }
}

// stats[n][g] = (mean, rstd).  One WARP per (n, group): lane l folds blocks l, l+32, ... in index order, then the 32
// partial (count, mean, M2) triples are merged by a fixed xor-shuffle tree — Chan et al.'s pairwise update in double
// precision at every step, so the order of operations is a function of (nblk, HW) only: deterministic, and 30x
// shorter than the serial fold over all blocks it replaces (that loop cost ~40 us per level at 720p).
__device__ __forceinline__ void chan_merge(double &cnt, double &mean, double &m2, double cb, double mb, double m2b) {
    if (cb == 0.0) return;
    const double tot = cnt + cb, delta = mb - mean;
    m2 += m2b + delta * delta * cnt * cb / tot;
    mean += delta * cb / tot;
    cnt = tot;
}
__global__ void groupnorm_combine_kernel(const float *__restrict__ part, float *__restrict__ stats, int nblk, int HW, int G,
                                         float eps) {
    const int n = blockIdx.x, g = blockIdx.y * 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= G) return;
    double cnt = 0.0, mean = 0.0, m2 = 0.0;
    for (int b = lane; b < nblk; b += 32) {
        const float *p = part + (((long long)n * nblk + b) * G + g) * 2;
        const double cb = 8.0 * (double)min(kGnRowsPerCta, HW - b * kGnRowsPerCta);
        const double sb = (double)p[0], mb = sb / cb;
        chan_merge(cnt, mean, m2, cb, mb, fmax((double)p[1] - sb * mb, 0.0));
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {          // lower lane's triple first: the same association on every replay
        const double oc = __shfl_xor_sync(0xffffffffu, cnt, o), om = __shfl_xor_sync(0xffffffffu, mean, o),
                     o2 = __shfl_xor_sync(0xffffffffu, m2, o);
        if (lane & o) {
            double c2 = oc, mm = om, q2 = o2;
            chan_merge(c2, mm, q2, cnt, mean, m2);
            cnt = c2; mean = mm; m2 = q2;
        } else {
            chan_merge(cnt, mean, m2, oc, om, o2);
        }
    }
    if (lane == 0) {
        stats[((long long)n * G + g) * 2] = (float)mean;
        stats[((long long)n * G + g) * 2 + 1] = (float)(1.0 / sqrt(m2 / cnt + (double)eps));
    }
}

template <typename IT, typename OT>
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const IT *__restrict__ x, const float *__restrict__ stats,
                                                              const float4 *__restrict__ gamma, const float4 *__restrict__ beta,
                                                              OT *__restrict__ out, int N, int HW, int C,
                                                              long long out_batch_stride, float eps) {
    const int c4 = C / 4, G = C / 8;
    const long long per = (long long)HW * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per * N; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / per);
        const long long j = i % per;
        const int col = (int)(j % c4);
        const float *st = stats + ((long long)n * G + col / 2) * 2;
        const float mean = st[0], rstd = st[1];
        const float4 v = load4(x + i * 4), g = __ldg(gamma + col), b = __ldg(beta + col);
        float4 o;
        o.x = (v.x - mean) * rstd * g.x + b.x;
        o.y = (v.y - mean) * rstd * g.y + b.y;
        o.z = (v.z - mean) * rstd * g.z + b.z;
        o.w = (v.w - mean) * rstd * g.w + b.w;
        store4(out + (long long)n * out_batch_stride + j * 4, o);
    }
}

// uint8 / float frames [T, 3, H, W] (planar, 0..255) -> normalised, zero-padded channels-last [T, Hp, Wp, 3]
template <typename T, typename OT>
__global__ void __launch_bounds__(256) preprocess_kernel(const T *__restrict__ in, OT *__restrict__ out, int Tn, int H,
                                                         int W, int Hp, int Wp, float m0, float m1, float m2, float s0,
                                                         float s1, float s2) {
    const long long total = (long long)Tn * Hp * Wp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % Wp);
        const long long t2 = i / Wp;
        const int h = (int)(t2 % Hp), t = (int)(t2 / Hp);
        float r = 0.f, g = 0.f, b = 0.f;
        if (h < H && w < W) {
            const long long base = ((long long)t * 3 * H + h) * W + w;
            r = ((float)in[base] - m0) / s0;
            g = ((float)in[base + (long long)H * W] - m1) / s1;
            b = ((float)in[base + 2ll * H * W] - m2) / s2;
        }
        out[i * 3] = (OT)r;
        out[i * 3 + 1] = (OT)g;
        out[i * 3 + 2] = (OT)b;
    }
}

// Same normalisation, written for the tensor-core stem (gemm_tcgen05.cu: stem_conv_device): [T, Hp + 6, Wp + 6, 4] fp32 with a
// 3-pixel zero border (the convolution's padding; the /32 padding of the frame is zero too) and a zero 4th channel.
template <typename T>
__global__ void __launch_bounds__(256) preprocess_padded4_kernel(const T *__restrict__ in, float4 *__restrict__ out, int Tn, int H,
                                                                 int W, int Hp, int Wp, float m0, float m1, float m2, float s0,
                                                                 float s1, float s2) {
    const int Wq = Wp + 6, Hq = Hp + 6;
    const long long total = (long long)Tn * Hq * Wq;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % Wq);
        const long long t2 = i / Wq;
        const int yq = (int)(t2 % Hq), t = (int)(t2 / Hq);
        const int h = yq - 3, w = xq - 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h >= 0 && h < H && w >= 0 && w < W) {
            const long long base = ((long long)t * 3 * H + h) * W + w;
            v.x = ((float)in[base] - m0) / s0;
            v.y = ((float)in[base + (long long)H * W] - m1) / s1;
            v.z = ((float)in[base + 2ll * H * W] - m2) / s2;
        }
        out[i] = v;
    }
}

static inline int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace vnext

namespace vnext {   // defined in gemm_tcgen05.cu
int stem_conv_device(int out_dt, const float *xpad, const float *w, const float *bias, void *out, int T, int Hp, int Wp, int Cout,
                     int relu, cudaStream_t stream);
}

using namespace vnext;

template <typename IT, typename OT>
static void groupnorm_launch(const void *x, const float *gamma, const float *beta, void *out, long long out_batch_stride,
                             float *stats_ws, int N, int HW, int C, int groups, float eps, cudaStream_t st) {
    const int nblk = (HW + kGnRowsPerCta - 1) / kGnRowsPerCta;
    float *part = stats_ws + (size_t)N * groups * 2;          // [N][nblk][G][2] after the [N][G][2] (mean, rstd) table
    dim3 grid(nblk, N);
    groupnorm_partial_kernel<IT><<<grid, kGnThreads, 0, st>>>((const IT *)x, part, HW, C);
    groupnorm_combine_kernel<<<dim3(N, (groups + 31) / 32), 1024, 0, st>>>(part, stats_ws, nblk, HW, groups, eps);
    const long long total = (long long)N * HW * (C / 4);
    groupnorm_apply_kernel<IT, OT><<<grid_for(total, 256), 256, 0, st>>>((const IT *)x, stats_ws, (const float4 *)gamma,
                                                                       (const float4 *)beta, (OT *)out, N, HW, C,
                                                                       out_batch_stride, eps);
}

extern "C" {

static inline bool dt_ok(int dt) { return dt == VNEXT_F32 || dt == VNEXT_BF16; }

int vnext_bias_act_dt(int dtype, const void *x, const float *bias, const void *residual, void *out, long long rows, int C,
                      int relu, void *stream) {
    VNEXT_CHECK_ARG(dt_ok(dtype), "vnext_bias_act: dtype must be fp32 or bf16");
    VNEXT_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0, "vnext_bias_act: C must be a positive multiple of 4 (got %d)", C);
    if (rows == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(x && bias && out, "vnext_bias_act: null pointer");
    const long long n4 = rows * (C / 4);
    if (dtype == VNEXT_F32)
        bias_act_kernel<float><<<grid_for(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            (const float *)x, (const float4 *)bias, (const float *)residual, (float *)out, n4, C / 4, relu);
    else
        bias_act_kernel<__nv_bfloat16><<<grid_for(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16 *)x, (const float4 *)bias, (const __nv_bfloat16 *)residual, (__nv_bfloat16 *)out, n4, C / 4, relu);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
int vnext_bias_act(const float *x, const float *bias, const float *residual, float *out, long long rows, int C, int relu,
                   void *stream) {
    return vnext_bias_act_dt(VNEXT_F32, x, bias, residual, out, rows, C, relu, stream);
}

int vnext_stem_pool_dt(int in_dtype, int out_dtype, const void *x, const float *bias, void *out, int N, int H, int W, int C,
                       void *stream) {
    VNEXT_CHECK_ARG(dt_ok(in_dtype) && dt_ok(out_dtype), "vnext_stem_pool: dtypes must be fp32 or bf16");
    VNEXT_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "vnext_stem_pool: bad shape");
    if (N == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(x && bias && out, "vnext_stem_pool: null pointer");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo * (C / 4);
    const int g = grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const float4 *b4 = (const float4 *)bias;
    if (in_dtype == VNEXT_F32 && out_dtype == VNEXT_F32)
        stem_pool_kernel<float, float><<<g, 256, 0, st>>>((const float *)x, b4, (float *)out, N, H, W, Ho, Wo, C / 4);
    else if (in_dtype == VNEXT_F32)
        stem_pool_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float *)x, b4, (__nv_bfloat16 *)out, N, H, W, Ho, Wo, C / 4);
    else if (out_dtype == VNEXT_F32)
        stem_pool_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16 *)x, b4, (float *)out, N, H, W, Ho, Wo, C / 4);
    else
        stem_pool_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16 *)x, b4, (__nv_bfloat16 *)out, N, H, W, Ho, Wo, C / 4);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
int vnext_stem_pool(const float *x, const float *bias, float *out, int N, int H, int W, int C, void *stream) {
    return vnext_stem_pool_dt(VNEXT_F32, VNEXT_F32, x, bias, out, N, H, W, C, stream);
}

size_t vnext_groupnorm_workspace_floats(int N, int HW, int groups) {
    if (N <= 0 || HW <= 0 || groups <= 0) return 0;
    const size_t nblk = (size_t)(HW + kGnRowsPerCta - 1) / kGnRowsPerCta;
    return (nblk + 1) * (size_t)N * groups * 2;
}

int vnext_groupnorm_nhwc_dt(int in_dtype, int out_dtype, const void *x, const float *gamma, const float *beta, void *out,
                            long long out_batch_stride, float *stats_ws, int N, int HW, int C, int groups, float eps,
                            void *stream) {
    VNEXT_CHECK_ARG(dt_ok(in_dtype) && dt_ok(out_dtype), "vnext_groupnorm_nhwc: dtypes must be fp32 or bf16");
    VNEXT_CHECK_ARG(N >= 0 && HW > 0 && C > 0 && groups > 0 && C == groups * 8 && C <= 1024 && kGnThreads % (C / 4) == 0,
                    "vnext_groupnorm_nhwc: supports C == 8*groups with C/4 dividing 256 (got C=%d groups=%d)", C, groups);
    if (N == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(x && gamma && beta && out && stats_ws && out_batch_stride % 4 == 0, "vnext_groupnorm_nhwc: bad pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == VNEXT_F32 && out_dtype == VNEXT_F32)
        groupnorm_launch<float, float>(x, gamma, beta, out, out_batch_stride, stats_ws, N, HW, C, groups, eps, st);
    else if (in_dtype == VNEXT_F32)
        groupnorm_launch<float, __nv_bfloat16>(x, gamma, beta, out, out_batch_stride, stats_ws, N, HW, C, groups, eps, st);
    else if (out_dtype == VNEXT_F32)
        groupnorm_launch<__nv_bfloat16, float>(x, gamma, beta, out, out_batch_stride, stats_ws, N, HW, C, groups, eps, st);
    else
        groupnorm_launch<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, out, out_batch_stride, stats_ws, N, HW, C, groups, eps, st);
    count_launch(3);
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
int vnext_groupnorm_nhwc(const float *x, const float *gamma, const float *beta, float *out, long long out_batch_stride,
                         float *stats_ws, int N, int HW, int C, int groups, float eps, void *stream) {
    return vnext_groupnorm_nhwc_dt(VNEXT_F32, VNEXT_F32, x, gamma, beta, out, out_batch_stride, stats_ws, N, HW, C, groups, eps,
                                   stream);
}

int vnext_preprocess_dt(int in_is_u8, int out_dtype, const void *frames, void *out, int T, int H, int W, int Hp, int Wp,
                        const float *mean3, const float *std3, void *stream) {
    VNEXT_CHECK_ARG(dt_ok(out_dtype), "vnext_preprocess: out dtype must be fp32 or bf16");
    VNEXT_CHECK_ARG(T >= 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "vnext_preprocess: bad shape");
    if (T == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(frames && out && mean3 && std3, "vnext_preprocess: null pointer");
    const long long total = (long long)T * Hp * Wp;
    const int g = grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define VNEXT_PP(IT, OT)                                                                                              \
    preprocess_kernel<IT, OT><<<g, 256, 0, st>>>((const IT *)frames, (OT *)out, T, H, W, Hp, Wp, mean3[0], mean3[1], \
                                                 mean3[2], std3[0], std3[1], std3[2])
    if (in_is_u8 && out_dtype == VNEXT_F32) VNEXT_PP(unsigned char, float);
    else if (in_is_u8) VNEXT_PP(unsigned char, __nv_bfloat16);
    else if (out_dtype == VNEXT_F32) VNEXT_PP(float, float);
    else VNEXT_PP(float, __nv_bfloat16);
#undef VNEXT_PP
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
int vnext_stem_conv7x7(int in_is_u8, int out_dtype, const void *frames, const float *weight_packed, const float *bias, void *out,
                       float *scratch, int T, int H, int W, int Hp, int Wp, int Cout, const float *mean3, const float *std3, int relu,
                       void *stream) {
    VNEXT_CHECK_ARG(dt_ok(out_dtype), "vnext_stem_conv7x7: out dtype must be fp32 or bf16");
    VNEXT_CHECK_ARG(T >= 0 && H > 0 && W > 0 && Hp >= H && Wp >= W && Hp % 2 == 0 && Wp % 2 == 0 && Cout > 0 && Cout % 8 == 0,
                    "vnext_stem_conv7x7: bad shape (H %d W %d Hp %d Wp %d Cout %d)", H, W, Hp, Wp, Cout);
    if (T == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(frames && weight_packed && out && scratch && mean3 && std3 &&
                        (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(weight_packed) & 15) == 0,
                    "vnext_stem_conv7x7: null or unaligned pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)T * (Hp + 6) * (Wp + 6);
    const int g = grid_for(total, 256);
    if (in_is_u8)
        preprocess_padded4_kernel<unsigned char><<<g, 256, 0, st>>>((const unsigned char *)frames, (float4 *)scratch, T, H, W, Hp, Wp,
                                                                   mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    else
        preprocess_padded4_kernel<float><<<g, 256, 0, st>>>((const float *)frames, (float4 *)scratch, T, H, W, Hp, Wp, mean3[0],
                                                           mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return stem_conv_device(out_dtype, scratch, weight_packed, bias, out, T, Hp, Wp, Cout, relu, st);
}

int vnext_preprocess(int in_is_u8, const void *frames, float *out, int T, int H, int W, int Hp, int Wp, const float *mean3,
                     const float *std3, void *stream) {
    return vnext_preprocess_dt(in_is_u8, VNEXT_F32, frames, out, T, H, W, Hp, Wp, mean3, std3, stream);
}

}  // extern "C"
