// C-ABI entry points of libvnext_b200.so (declared in include/vnext_b200.h).
#include <mutex>
#include <string>

#include "common.cuh"

namespace vnext {

static thread_local std::string t_error;
static thread_local const char *t_variant = "none";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}
void set_variant(const char *name) { t_variant = name; }

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int msda_forward_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                        const void *attn, int N, int S, int M, int D, int L, int Lq, int K, void *out,
                        cudaStream_t stream);
int msda_backward_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                         const void *attn, const void *gout, int N, int S, int M, int D, int L, int Lq, int K,
                         void *gvalue, void *gloc, void *gattn, void *workspace, cudaStream_t stream);

bool msda_window_supported(int dtype, int D, int L, int K);
int msda_forward_cached_device(int dtype, const void *value, const int64_t *hshapes, const int64_t *hlsi, const float *loc,
                               const float *attn, const float *oa, int oa_pitch, const float *ref, int ref_dim,
                               int ref_shared, int N, int S, int M, int L, int K, void *out, cudaStream_t stream);
int msda_forward_window_device(int dtype, const void *value, const int64_t *hshapes, const int64_t *hlsi, const float *loc,
                               const float *attn, const float *oa, int oa_pitch, const float *ref, int ref_dim,
                               int ref_shared, int N, int S, int M, int L, int K, int margin, void *out,
                               cudaStream_t stream);

int msda_forward_fused_device(int dtype, const void *value, const int64_t *shapes, const int64_t *lsi, const float *oa,
                              int oa_pitch, int n_off, const float *ref, int ref_dim, int ref_shared, int N, int S,
                              int M, int D, int L, int Lq, int K, void *out, cudaStream_t stream);

static int check_dims(const char *who, int dtype, int N, int S, int M, int D, int L, int Lq, int K) {
    VNEXT_CHECK_ARG(dtype == VNEXT_F32 || dtype == VNEXT_F64 || dtype == VNEXT_BF16, "%s: unsupported dtype %d", who,
                    dtype);
    VNEXT_CHECK_ARG(N >= 0 && S >= 0 && Lq >= 0, "%s: negative size (batch=%d spatial_size=%d num_query=%d)", who, N,
                    S, Lq);
    VNEXT_CHECK_ARG(M > 0 && D > 0 && L > 0 && K > 0,
                    "%s: num_heads, channels, num_levels, num_point must be positive (got %d %d %d %d)", who, M, D, L,
                    K);
    return VNEXT_OK;
}

static size_t value_esize(int dtype) { return dtype == VNEXT_F64 ? 8 : (dtype == VNEXT_BF16 ? 2 : 4); }
static size_t loc_esize(int dtype) { return dtype == VNEXT_F64 ? 8 : 4; }

// Device staging for the *_host entry points: grows, never shrinks, one set per process.
struct HostStage {
    std::mutex mu;
    void *buf[12] = {nullptr};
    size_t cap[12] = {0};
    cudaStream_t stream = nullptr;
    int ensure(int i, size_t bytes) {
        if (bytes <= cap[i]) return VNEXT_OK;
        if (buf[i]) cudaFree(buf[i]);
        buf[i] = nullptr;
        cap[i] = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&buf[i], want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return VNEXT_ENOMEM;
        }
        cap[i] = want;
        return VNEXT_OK;
    }
};
static HostStage g_stage;

}  // namespace vnext

using namespace vnext;

extern "C" {

const char *vnext_last_error(void) { return t_error.c_str(); }
const char *vnext_version(void) { return "vnext_b200 0.1.0 sm_100a"; }
unsigned long long vnext_launch_count(void) { return g_launches.load(); }
const char *vnext_msda_last_variant(void) { return t_variant; }

int vnext_msda_forward(int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point, void *output, void *stream) {
    int rc = check_dims("vnext_msda_forward", dtype, batch, spatial_size, num_heads, channels, num_levels, num_query,
                        num_point);
    if (rc) return rc;
    const bool empty = (long long)batch * num_query == 0;
    VNEXT_CHECK_ARG(empty || (value && spatial_shapes && level_start_index && sampling_loc && attn_weight && output),
                    "vnext_msda_forward: null tensor pointer");
    return msda_forward_device(dtype, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, batch,
                               spatial_size, num_heads, channels, num_levels, num_query, num_point, output,
                               static_cast<cudaStream_t>(stream));
}

int vnext_msda_forward_selfattn(int dtype, const void *value, const int64_t *spatial_shapes_host,
                                const int64_t *level_start_index_host, const void *sampling_loc, const void *attn_weight,
                                const float *offsets_logits, int row_pitch, const float *reference_points, int ref_dim,
                                int ref_shared, int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                int num_point, int margin, void *output, void *stream) {
    int rc = check_dims("vnext_msda_forward_selfattn", dtype, batch, spatial_size, num_heads, channels, num_levels,
                        spatial_size, num_point);
    if (rc) return rc;
    const bool empty = (long long)batch * spatial_size == 0;
    VNEXT_CHECK_ARG(empty || (value && spatial_shapes_host && level_start_index_host && output &&
                              ((sampling_loc && attn_weight) || (offsets_logits && reference_points) ||
                               (offsets_logits && margin < 0))),
                    "vnext_msda_forward_selfattn: null tensor pointer");
    VNEXT_CHECK_ARG(channels == 32 && msda_window_supported(dtype, channels, num_levels, num_point),
                    "vnext_msda_forward_selfattn: needs channels == 32, num_levels <= 4, num_levels*num_point in {8,16,32}, "
                    "fp32/bf16 value (got D=%d L=%d K=%d dtype=%d)", channels, num_levels, num_point, dtype);
    if (margin < 0)      // cache-blocked variant (csrc/msda_cached.cu): no shared-memory windows, the L1 does the staging
        return msda_forward_cached_device(dtype, value, spatial_shapes_host, level_start_index_host,
                                          static_cast<const float *>(sampling_loc), static_cast<const float *>(attn_weight),
                                          offsets_logits, row_pitch, reference_points, ref_dim, ref_shared, batch,
                                          spatial_size, num_heads, num_levels, num_point, output,
                                          static_cast<cudaStream_t>(stream));
    return msda_forward_window_device(dtype, value, spatial_shapes_host, level_start_index_host,
                                      static_cast<const float *>(sampling_loc), static_cast<const float *>(attn_weight),
                                      offsets_logits, row_pitch, reference_points, ref_dim, ref_shared, batch, spatial_size,
                                      num_heads, num_levels, num_point, margin, output, static_cast<cudaStream_t>(stream));
}

int vnext_msda_backward(int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_loc, const void *attn_weight, const void *grad_output, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                        void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, void *workspace,
                        void *stream) {
    int rc = check_dims("vnext_msda_backward", dtype, batch, spatial_size, num_heads, channels, num_levels, num_query,
                        num_point);
    if (rc) return rc;
    const bool empty = (long long)batch * num_query == 0;
    VNEXT_CHECK_ARG(empty || (value && spatial_shapes && level_start_index && sampling_loc && attn_weight &&
                              grad_output && grad_sampling_loc && grad_attn_weight),
                    "vnext_msda_backward: null tensor pointer");
    VNEXT_CHECK_ARG((long long)batch * spatial_size == 0 || grad_value, "vnext_msda_backward: null grad_value");
    return msda_backward_device(dtype, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                grad_output, batch, spatial_size, num_heads, channels, num_levels, num_query,
                                num_point, grad_value, grad_sampling_loc, grad_attn_weight, workspace,
                                static_cast<cudaStream_t>(stream));
}

int vnext_msda_forward_fused(int dtype, const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *offsets_logits, int row_pitch,
                             const float *reference_points, int ref_dim, int ref_shared, int batch, int spatial_size,
                             int num_heads, int channels, int num_levels, int num_query, int num_point, void *output,
                             void *stream) {
    int rc = check_dims("vnext_msda_forward_fused", dtype, batch, spatial_size, num_heads, channels, num_levels,
                        num_query, num_point);
    if (rc) return rc;
    const bool empty = (long long)batch * num_query == 0;
    VNEXT_CHECK_ARG(empty || (value && spatial_shapes && level_start_index && offsets_logits && reference_points && output),
                    "vnext_msda_forward_fused: null tensor pointer");
    return msda_forward_fused_device(dtype, value, spatial_shapes, level_start_index, offsets_logits, row_pitch,
                                     num_heads * num_levels * num_point * 2, reference_points, ref_dim, ref_shared, batch,
                                     spatial_size, num_heads, channels, num_levels, num_query, num_point, output,
                                     static_cast<cudaStream_t>(stream));
}

size_t vnext_msda_backward_workspace_bytes(int dtype, int batch, int spatial_size, int num_heads, int channels) {
    if (dtype != VNEXT_BF16) return 0;
    return (size_t)batch * spatial_size * num_heads * channels * sizeof(float);
}

int vnext_msda_forward_host(int dtype, const void *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                            int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                            int num_point, void *output) {
    int rc = check_dims("vnext_msda_forward_host", dtype, batch, spatial_size, num_heads, channels, num_levels,
                        num_query, num_point);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (!g_stage.stream) VNEXT_CUDA(cudaStreamCreateWithFlags(&g_stage.stream, cudaStreamNonBlocking));
    cudaStream_t st = g_stage.stream;
    const size_t ve = value_esize(dtype), le = loc_esize(dtype);
    const size_t pairs = (size_t)batch * num_query * num_heads, lk = (size_t)num_levels * num_point;
    const size_t nb[6] = {(size_t)batch * spatial_size * num_heads * channels * ve, (size_t)num_levels * 2 * 8,
                          (size_t)num_levels * 8, pairs * lk * 2 * le, pairs * lk * le, pairs * channels * ve};
    const void *src[5] = {value, spatial_shapes, level_start_index, sampling_loc, attn_weight};
    for (int i = 0; i < 6; ++i)
        if ((rc = g_stage.ensure(i, nb[i]))) return rc;
    for (int i = 0; i < 5; ++i)
        if (nb[i]) VNEXT_CUDA(cudaMemcpyAsync(g_stage.buf[i], src[i], nb[i], cudaMemcpyHostToDevice, st));
    rc = msda_forward_device(dtype, g_stage.buf[0], (const int64_t *)g_stage.buf[1], (const int64_t *)g_stage.buf[2],
                             g_stage.buf[3], g_stage.buf[4], batch, spatial_size, num_heads, channels, num_levels,
                             num_query, num_point, g_stage.buf[5], st);
    if (rc) return rc;
    if (nb[5]) VNEXT_CUDA(cudaMemcpyAsync(output, g_stage.buf[5], nb[5], cudaMemcpyDeviceToHost, st));
    VNEXT_CUDA(cudaStreamSynchronize(st));
    return VNEXT_OK;
}

int vnext_msda_backward_host(int dtype, const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight,
                             const void *grad_output, int batch, int spatial_size, int num_heads, int channels,
                             int num_levels, int num_query, int num_point, void *grad_value, void *grad_sampling_loc,
                             void *grad_attn_weight) {
    int rc = check_dims("vnext_msda_backward_host", dtype, batch, spatial_size, num_heads, channels, num_levels,
                        num_query, num_point);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (!g_stage.stream) VNEXT_CUDA(cudaStreamCreateWithFlags(&g_stage.stream, cudaStreamNonBlocking));
    cudaStream_t st = g_stage.stream;
    const size_t ve = value_esize(dtype), le = loc_esize(dtype);
    const size_t pairs = (size_t)batch * num_query * num_heads, lk = (size_t)num_levels * num_point;
    const size_t nvalue = (size_t)batch * spatial_size * num_heads * channels;
    const size_t nb[10] = {nvalue * ve, (size_t)num_levels * 2 * 8, (size_t)num_levels * 8, pairs * lk * 2 * le,
                           pairs * lk * le, pairs * channels * ve,
                           nvalue * ve, pairs * lk * 2 * le, pairs * lk * le,
                           vnext_msda_backward_workspace_bytes(dtype, batch, spatial_size, num_heads, channels)};
    const void *src[6] = {value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output};
    void *dst[3] = {grad_value, grad_sampling_loc, grad_attn_weight};
    for (int i = 0; i < 10; ++i)
        if ((rc = g_stage.ensure(i, nb[i]))) return rc;
    for (int i = 0; i < 6; ++i)
        if (nb[i]) VNEXT_CUDA(cudaMemcpyAsync(g_stage.buf[i], src[i], nb[i], cudaMemcpyHostToDevice, st));
    rc = msda_backward_device(dtype, g_stage.buf[0], (const int64_t *)g_stage.buf[1], (const int64_t *)g_stage.buf[2],
                              g_stage.buf[3], g_stage.buf[4], g_stage.buf[5], batch, spatial_size, num_heads,
                              channels, num_levels, num_query, num_point, g_stage.buf[6], g_stage.buf[7],
                              g_stage.buf[8], g_stage.buf[9], st);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i)
        if (nb[6 + i]) VNEXT_CUDA(cudaMemcpyAsync(dst[i], g_stage.buf[6 + i], nb[6 + i], cudaMemcpyDeviceToHost, st));
    VNEXT_CUDA(cudaStreamSynchronize(st));
    return VNEXT_OK;
}

}  // extern "C"
