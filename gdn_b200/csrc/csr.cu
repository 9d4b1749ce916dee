// csr.cu -- GraphLayer attention on an ARBITRARY edge list (CSR by target), any heads.
//
// This is GraphLayer's own module boundary (models/graph_layer.py:53-117) for callers that
// hand it a general edge_index instead of the window-shared top-k graph; the GDN model never
// takes this path (it uses attention.cu).  Sparse parts only: the dense contractions around
// it (lin weight / attention-vector gradients) are plain matmuls done by the caller.
//
//   xl = x lin^T [n, H*D];  s_i[r,h] = <xl[r,h,:], att_i[h]> + <emb[r], att_em_i[h]>  (s_j alike)
//   pre[e,h] = s_i[dst,h] + s_j[src,h];  alpha = segment-softmax_dst(leaky_relu(pre))
//   out_h[r,h,:] = sum_{e: dst=r} alpha[e,h] xl[src_e, h, :]
#include "common.cuh"
#include "launchers.h"

namespace gdn {

__global__ void k_csr_lin(const float* __restrict__ x, const float* __restrict__ emb, const float* __restrict__ lin,
                          const float* __restrict__ a_i, const float* __restrict__ a_j,
                          const float* __restrict__ ae_i, const float* __restrict__ ae_j,
                          int n, int W, int D, int H, float* __restrict__ xl, float* __restrict__ s_i,
                          float* __restrict__ s_j) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < n; r += nwarps) {
        for (int h = 0; h < H; ++h) {
            float si = 0.f, sj = 0.f;
            for (int d = lane; d < D; d += 32) {
                const int c = h * D + d;
                float acc = 0.f;
                for (int w = 0; w < W; ++w) acc = fmaf(lin[(size_t)c * W + w], x[(size_t)r * W + w], acc);
                xl[(size_t)r * H * D + c] = acc;
                const float ev = emb ? emb[(size_t)r * D + d] : 0.f;
                si += acc * a_i[c] + ev * ae_i[c];
                sj += acc * a_j[c] + ev * ae_j[c];
            }
            si = warp_sum(si);
            sj = warp_sum(sj);
            if (lane == 0) { s_i[(size_t)r * H + h] = si; s_j[(size_t)r * H + h] = sj; }
        }
    }
}

__global__ void k_csr_attn_fwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const float* __restrict__ xl, const float* __restrict__ s_i, const float* __restrict__ s_j,
                               int n, int D, int H, float slope, float* __restrict__ out_h, float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < (long long)n * H; t += nwarps) {
        const int r = (int)(t / H), h = (int)(t % H);
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        const float si = s_i[(size_t)r * H + h];
        float m = -INFINITY;
        for (int e = e0 + lane; e < e1; e += 32) {
            const float pre = si + s_j[(size_t)col[e] * H + h];
            m = fmaxf(m, pre > 0.f ? pre : slope * pre);
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int e = e0 + lane; e < e1; e += 32) {
            const float pre = si + s_j[(size_t)col[e] * H + h];
            const float p = __expf((pre > 0.f ? pre : slope * pre) - m);
            alpha[(size_t)e * H + h] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float linv = 1.f / (sum + GDN_SOFTMAX_EPS);
        for (int e = e0 + lane; e < e1; e += 32) alpha[(size_t)e * H + h] *= linv;
        __syncwarp();
        for (int d = lane; d < D; d += 32) {
            float acc = 0.f;
            for (int e = e0; e < e1; ++e)
                acc = fmaf(alpha[(size_t)e * H + h], xl[(size_t)col[e] * H * D + h * D + d], acc);
            out_h[((size_t)r * H + h) * D + d] = acc;
        }
    }
}

__global__ void k_csr_attn_bwd(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const float* __restrict__ xl, const float* __restrict__ s_i, const float* __restrict__ s_j,
                               const float* __restrict__ alpha, const float* __restrict__ g_out_h,
                               int n, int D, int H, float slope, float* __restrict__ g_xl, float* __restrict__ g_si,
                               float* __restrict__ g_sj) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < (long long)n * H; t += nwarps) {
        const int r = (int)(t / H), h = (int)(t % H);
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        const float si = s_i[(size_t)r * H + h];
        const float* go = g_out_h + ((size_t)r * H + h) * D;
        // dot = sum_e alpha_e <go, xl[src_e]>
        float dot = 0.f;
        for (int e = e0; e < e1; ++e) {
            const float* xs = xl + (size_t)col[e] * H * D + h * D;
            float ga = 0.f;
            for (int d = lane; d < D; d += 32) ga = fmaf(go[d], xs[d], ga);
            ga = warp_sum(ga);
            dot = fmaf(alpha[(size_t)e * H + h], ga, dot);
        }
        float gsi = 0.f;
        for (int e = e0; e < e1; ++e) {
            const int src = col[e];
            const float* xs = xl + (size_t)src * H * D + h * D;
            const float a = alpha[(size_t)e * H + h];
            float ga = 0.f;
            for (int d = lane; d < D; d += 32) {
                ga = fmaf(go[d], xs[d], ga);
                atomicAdd(g_xl + (size_t)src * H * D + h * D + d, a * go[d]);
            }
            ga = warp_sum(ga);
            const float pre = si + s_j[(size_t)src * H + h];
            const float gl = a * (ga - dot);
            const float gp = pre > 0.f ? gl : slope * gl;
            gsi += gp;
            if (lane == 0) atomicAdd(g_sj + (size_t)src * H + h, gp);
        }
        if (lane == 0) g_si[(size_t)r * H + h] = gsi;
    }
}

}  // namespace gdn

using namespace gdn;

extern "C" int gdn_csr_fwd(int n, int W, int D, int H, int64_t E, const int32_t* rowptr, const int32_t* col,
                           const float* x, const float* emb, const float* lin_weight, const float* att_i,
                           const float* att_j, const float* att_em_i, const float* att_em_j, float* xl, float* s_i,
                           float* s_j, float* out_h, float* alpha, float negative_slope, void* stream) {
    GDN_CHECK_ARG(n >= 1 && W >= 1 && D >= 1 && H >= 1 && E >= 0, "csr_fwd: bad shape");
    GDN_CHECK_ARG(rowptr && col && x && lin_weight && att_i && att_j && att_em_i && att_em_j && xl && s_i && s_j &&
                      out_h && alpha, "csr_fwd: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    int g = ceil_div(n, 8);
    if (g > 16 * num_sms()) g = 16 * num_sms();
    k_csr_lin<<<g, 256, 0, st>>>(x, emb, lin_weight, att_i, att_j, att_em_i, att_em_j, n, W, D, H, xl, s_i, s_j);
    GDN_CHECK_LAUNCH("k_csr_lin");
    long long g2 = ((long long)n * H + 7) / 8;
    if (g2 > 16 * num_sms()) g2 = 16 * num_sms();
    k_csr_attn_fwd<<<(int)g2, 256, 0, st>>>(rowptr, col, xl, s_i, s_j, n, D, H, negative_slope, out_h, alpha);
    GDN_CHECK_LAUNCH("k_csr_attn_fwd");
    return 0;
}

extern "C" int gdn_csr_bwd(int n, int W, int D, int H, int64_t E, const int32_t* rowptr, const int32_t* col,
                           const float* x, const float* emb, const float* lin_weight, const float* att_i,
                           const float* att_j, const float* att_em_i, const float* att_em_j, const float* xl,
                           const float* s_i, const float* s_j, const float* alpha, const float* g_out_h, float* g_xl,
                           float* g_si, float* g_sj, float negative_slope, void* stream) {
    (void)W; (void)x; (void)emb; (void)lin_weight; (void)att_i; (void)att_j; (void)att_em_i; (void)att_em_j;
    GDN_CHECK_ARG(n >= 1 && D >= 1 && H >= 1 && E >= 0, "csr_bwd: bad shape");
    GDN_CHECK_ARG(rowptr && col && xl && s_i && s_j && alpha && g_out_h && g_xl && g_si && g_sj,
                  "csr_bwd: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(g_xl, 0, (size_t)n * H * D * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset g_xl");
    e = cudaMemsetAsync(g_sj, 0, (size_t)n * H * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset g_sj");
    long long g2 = ((long long)n * H + 7) / 8;
    if (g2 > 16 * num_sms()) g2 = 16 * num_sms();
    k_csr_attn_bwd<<<(int)g2, 256, 0, st>>>(rowptr, col, xl, s_i, s_j, alpha, g_out_h, n, D, H, negative_slope, g_xl,
                                            g_si, g_sj);
    GDN_CHECK_LAUNCH("k_csr_attn_bwd");
    return 0;
}
