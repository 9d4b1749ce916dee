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

__global__ void __launch_bounds__(256)
k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps, float weight_decay,
            float grad_scale) {
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

int launch_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, long long step, float grad_scale, cudaStream_t st) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    long long gsz = (n + 255) / 256;
    if (gsz > 8LL * num_sms()) gsz = 8LL * num_sms();
    if (gsz < 1) gsz = 1;
    k_adam_flat<<<(int)gsz, 256, 0, st>>>(p, g, m, v, n, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps,
                                          weight_decay, grad_scale);
    GDN_CHECK_LAUNCH("k_adam_flat");
    return 0;
}

}  // namespace gdn
