// Fused AdamW over a flat fp32 segment — the optimiser step of the IDOL training path (reference: torch.optim.AdamW
// built per parameter by projects/IDOL/train_net.py:94-140, with full-model gradient clipping :110-125).
//
// All trainable parameters live in one flat buffer (vnext_b200/idol/train.py: FlatParams), so one launch per
// learning-rate group replaces ~360 per-tensor updates.  HBM-bound: 4 reads + 3 writes of 4 B per element = 28 B/elt.
// The gradient scale (clip coefficient / world size) is read from device memory so the host never synchronises
// between backward and the update.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace vnext {

__global__ void __launch_bounds__(256) adamw_kernel(float4 *__restrict__ w, const float4 *__restrict__ g,
                                                    float4 *__restrict__ m, float4 *__restrict__ v, long long n4,
                                                    float lr, float b1, float b2, float eps, float wd, float bc1,
                                                    float rsqrt_bc2, const float *__restrict__ grad_scale) {
    const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
    const float step = lr / bc1;
    const float decay = 1.0f - lr * wd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 wi = w[i], gi = g[i], mi = m[i], vi = v[i];
        float *wp = &wi.x, *gp = &gi.x, *mp = &mi.x, *vp = &vi.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gp[k] * gs;
            mp[k] = b1 * mp[k] + (1.0f - b1) * gr;
            vp[k] = b2 * vp[k] + (1.0f - b2) * gr * gr;
            const float denom = sqrtf(vp[k]) * rsqrt_bc2 + eps;
            wp[k] = wp[k] * decay - step * (mp[k] / denom);
        }
        w[i] = wi;
        m[i] = mi;
        v[i] = vi;
    }
}

}  // namespace vnext

extern "C" int vnext_adamw(float *w, const float *g, float *m, float *v, long long n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, float bias_correction1, float bias_correction2,
                           const float *grad_scale, void *stream) {
    VNEXT_CHECK_ARG(w && g && m && v, "vnext_adamw: null pointer");
    VNEXT_CHECK_ARG(n >= 0 && n % 4 == 0, "vnext_adamw: n must be a non-negative multiple of 4 (got %lld)", n);
    VNEXT_CHECK_ARG(((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0,
                    "vnext_adamw: buffers must be 16-byte aligned");
    VNEXT_CHECK_ARG(bias_correction1 > 0.f && bias_correction2 > 0.f, "vnext_adamw: bias corrections must be > 0");
    if (n == 0) return VNEXT_OK;
    const long long n4 = n / 4;
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)vnext::num_sms() * 8);
    vnext::adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (float4 *)w, (const float4 *)g, (float4 *)m, (float4 *)v, n4, lr, beta1, beta2, eps, weight_decay,
        bias_correction1, 1.0f / sqrtf(bias_correction2), grad_scale);
    vnext::count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
