// optim.cu -- Adam on one flat parameter buffer, with the data-parallel gradient scaling fused in
// (train.py:31 torch.optim.Adam(model.parameters(), lr, weight_decay), train.py:73 optimizer.step();
// SURVEY.md section 8 row f-4).
//
// The data-parallel trainer keeps parameters, gradients and both moments in four flat fp32 buffers: autograd
// accumulates straight into views of the gradient buffer, NCCL all-reduces that buffer in place, and this one
// kernel applies   g <- g * grad_scale (+ weight_decay * p);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
// p <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)   (torch.optim.Adam, amsgrad=False), i.e. the 1/G of the
// gradient average never takes a pass of its own.  HBM-bound: 16 bytes read + 12 written per parameter.
#include "common.cuh"
#include "launchers.h"

namespace gdn {

// step_dev != NULL: the step count lives on the device (incremented by the caller on the same stream) and the bias
// corrections are derived from it here -- a captured CUDA graph then advances Adam's step on every replay
__global__ void __launch_bounds__(256)
k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps, float weight_decay,
            float grad_scale, const long long* __restrict__ step_dev, float lr) {
    if (step_dev != nullptr) {
        __shared__ float bc[2];
        if (threadIdx.x == 0) {
            const double step = (double)*step_dev;
            bc[0] = (float)((double)lr / (1.0 - pow((double)beta1, step)));
            bc[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, step)));
        }
        __syncthreads();
        lr_over_bc1 = bc[0];
        inv_sqrt_bc2 = bc[1];
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pi = p[i];
        float gi = g[i] * grad_scale;
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);          // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);     // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = fmaf(sqrtf(vi), inv_sqrt_bc2, eps);
        p[i] = pi - lr_over_bc1 * (mi / denom);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Data parallel, NVSwitch multicast (SURVEY.md section 8 row f-4 as written): gradient all-reduce, Adam and the
// broadcast of the new parameters in ONE kernel.  `g_mc` / `p_mc` are MULTICAST addresses of the flat gradient /
// parameter buffers (one symmetric allocation per rank, bound to one multicast object): a `multimem.ld_reduce`
// on g_mc returns the sum over all ranks' gradients, reduced inside the switch; a `multimem.st` on p_mc writes
// every rank's replica.  Rank r owns elements [lo, lo + cnt) -- a 1/G slice, so the optimizer state (m, v) and the
// update arithmetic are sharded G-fold too -- reads the reduced gradient of its slice, updates its slice of the
// moments, and stores the new parameters to all replicas.  Every replica receives bit-identical parameters by
// construction.  The caller brackets the launch with cross-rank barriers (gradients complete before, parameters
// visible after).  Per parameter: 4 bytes pulled through the switch per rank-slice, 4 bytes pushed.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256)
k_nvls_adam(const float* __restrict__ p_local, float* __restrict__ p_mc, const float* __restrict__ g_mc,
            float* __restrict__ m, float* __restrict__ v, long long lo, long long cnt,
            float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps, float weight_decay,
            float grad_scale) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; k < cnt; k += stride) {
        const float4 gs = multimem_ld_reduce_add(g_mc + lo + k);
        const float4 p4 = *reinterpret_cast<const float4*>(p_local + lo + k);
        float4 m4 = *reinterpret_cast<const float4*>(m + k), v4 = *reinterpret_cast<const float4*>(v + k);
        const float pi[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {gs.x, gs.y, gs.z, gs.w};
        float mi[4] = {m4.x, m4.y, m4.z, m4.w}, vi[4] = {v4.x, v4.y, v4.z, v4.w}, po[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float gi = gg[u] * grad_scale;
            if (weight_decay != 0.f) gi = fmaf(weight_decay, pi[u], gi);
            mi[u] = fmaf(beta1, mi[u], (1.f - beta1) * gi);
            vi[u] = fmaf(beta2, vi[u], (1.f - beta2) * gi * gi);
            const float denom = fmaf(sqrtf(vi[u]), inv_sqrt_bc2, eps);
            po[u] = pi[u] - lr_over_bc1 * (mi[u] / denom);
        }
        *reinterpret_cast<float4*>(m + k) = make_float4(mi[0], mi[1], mi[2], mi[3]);
        *reinterpret_cast<float4*>(v + k) = make_float4(vi[0], vi[1], vi[2], vi[3]);
        multimem_st(p_mc + lo + k, make_float4(po[0], po[1], po[2], po[3]));
    }
    __threadfence_system();
}

int launch_nvls_adam(const float* p_local, float* p_mc, const float* g_mc, float* m, float* v, long long lo, long long cnt,
                     float lr, float beta1, float beta2, float eps, float weight_decay, long long step, float grad_scale,
                     cudaStream_t st) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long gsz = (cnt / 4 + 255) / 256;
    if (gsz > 2LL * num_sms()) gsz = 2LL * num_sms();
    if (gsz < 1) gsz = 1;
    k_nvls_adam<<<(int)gsz, 256, 0, st>>>(p_local, p_mc, g_mc, m, v, lo, cnt, (float)((double)lr / bc1),
                                          (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, weight_decay, grad_scale);
    GDN_CHECK_LAUNCH("k_nvls_adam");
    return 0;
}

int launch_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, long long step, const long long* step_dev, float grad_scale,
                     cudaStream_t st) {
    const double sh = (double)(step >= 1 ? step : 1);
    const double bc1 = 1.0 - pow((double)beta1, sh), bc2 = 1.0 - pow((double)beta2, sh);
    long long gsz = (n + 255) / 256;
    if (gsz > 8LL * num_sms()) gsz = 8LL * num_sms();
    if (gsz < 1) gsz = 1;
    k_adam_flat<<<(int)gsz, 256, 0, st>>>(p, g, m, v, n, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps,
                                          weight_decay, grad_scale, step_dev, lr);
    GDN_CHECK_LAUNCH("k_adam_flat");
    return 0;
}

}  // namespace gdn
