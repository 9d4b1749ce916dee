// attention.cu -- GraphLayer attention message pass on the window-shared top-k graph.
//
// Replaces models/graph_layer.py:82-117 + PyG propagate/softmax/scatter-add
// (SURVEY.md section 8 rows a3/a4).  Algebra (verified against the
// reference by tests/): with u_i = Wl^T a_i, u_j = Wl^T a_j (R^W), e_i = V ae_i,
// e_j = V ae_j (R^N):
//     s_i[b,i] = x[b,i].u_i + e_i[i]        s_j[b,j] = x[b,j].u_j + e_j[j]
//     pre[b,i,k] = s_i[b,i] + s_j[b,S(i)_k]   alpha = softmax_k(leaky_relu(pre))
//     A[b,i,:] = sum_k alpha[b,i,k] x[b,S(i)_k,:]          (lin has no bias, so the
//     D-wide transform Wl.A commutes with the aggregation and runs afterwards).
//
// Layout: the graph is identical for every window, so a warp owns one target sensor i and
// 32 windows (lane <-> b).  x is first transposed to xT[N][Bs/32][WP][32] (window index fastest):
// every gather of a neighbour is then one fully used 128-byte line per (source, w), the
// neighbour index is warp-uniform, and the backward's scatter into g_sj is one coalesced
// RED per edge instead of 32 scattered atomics.
#include "common.cuh"
#include "launchers.h"

namespace gdn {

// ---------------------------------------------------------------------------------------
// u_i, u_j (R^W) and e_i, e_j (R^N)
// ---------------------------------------------------------------------------------------
__global__ void k_node_scalars(const float* __restrict__ V, const float* __restrict__ Wl,
                               const float* __restrict__ a_i, const float* __restrict__ a_j,
                               const float* __restrict__ ae_i, const float* __restrict__ ae_j,
                               int N, int D, int W, float* __restrict__ uv, float* __restrict__ ev) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x < 64) {
        const int which = threadIdx.x >> 5, w = threadIdx.x & 31;
        float acc = 0.f;
        if (w < W) {
            const float* a = which ? a_j : a_i;
            for (int d = 0; d < D; ++d) acc = fmaf(Wl[d * W + w], a[d], acc);
        }
        uv[which * 32 + w] = acc;
    }
    for (int i = warp; i < N; i += nwarps) {
        float si = 0.f, sj = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = V[(size_t)i * D + d];
            si = fmaf(v, ae_i[d], si);
            sj = fmaf(v, ae_j[d], sj);
        }
        si = warp_sum(si);
        sj = warp_sum(sj);
        if (lane == 0) { ev[i] = si; ev[N + i] = sj; }
    }
}

// ---------------------------------------------------------------------------------------
// x[B][N][W] -> xT[N][C][WP][32] (C = Bs/32 window chunks, w padded with zeros to WP),
// s_iT[N][Bs], s_jT[N][Bs]   (zero padded for b >= B)
// A lane's 16 gathers of one neighbour are then base + w*32: immediate offsets, no address math.
// ---------------------------------------------------------------------------------------
__global__ void k_transpose_scalars(const float* __restrict__ x, const float* __restrict__ uv,
                                    const float* __restrict__ ev, int B, int N, int W, int WP, int Bs, int vec4,
                                    float* __restrict__ xT, float* __restrict__ siT, float* __restrict__ sjT) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    for (long long t = warp; t < tasks; t += nwarps) {
        const int i = (int)(t / chunks), c = (int)(t % chunks);
        const int b = c * 32 + lane;
        float si = 0.f, sj = 0.f;
        const bool ok = b < B;
        const float* row = x + ((size_t)b * N + i) * W;
        float* dst = xT + ((size_t)i * chunks + c) * WP * 32 + lane;
        if (vec4) {
            // 16-byte loads (W % 4 == 0 and x 16-byte aligned): a lane's row is W contiguous floats, rows of different lanes are N*W apart, so the
            // number of L1 requests (each touching 32 lines) is what bounds this kernel
            for (int w = 0; w < WP; w += 4) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && w < W) q = __ldg(reinterpret_cast<const float4*>(row + w));
                const float v4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    dst[(w + u) * 32] = v4[u];
                    if (w + u < W) {
                        si = fmaf(v4[u], uv[w + u], si);
                        sj = fmaf(v4[u], uv[32 + w + u], sj);
                    }
                }
            }
        } else {
            for (int w = 0; w < WP; ++w) {
                const float v = (ok && w < W) ? __ldg(row + w) : 0.f;
                dst[w * 32] = v;
                if (w < W) {
                    si = fmaf(v, uv[w], si);
                    sj = fmaf(v, uv[32 + w], sj);
                }
            }
        }
        siT[(size_t)i * Bs + b] = ok ? si + ev[i] : 0.f;
        sjT[(size_t)i * Bs + b] = ok ? sj + ev[N + i] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// forward: A, segment max m, 1/(segment sum + 1e-16); optional alpha
// tasks are chunk-major (all sensors of window chunk 0, then chunk 1, ...) so that the
// gather working set at any time is one chunk of xT (N*WP*128 bytes)
// ---------------------------------------------------------------------------------------
template <int WP>
__global__ void __launch_bounds__(256)
k_attn_fwd(const float* __restrict__ xT, const float* __restrict__ siT, const float* __restrict__ sjT,
           const int32_t* __restrict__ nbr, int B, int N, int W, int Kp, int Bs,
           float* __restrict__ A, float* __restrict__ mT, float* __restrict__ linvT,
           float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    const unsigned xstride = (unsigned)chunks * WP * 32;
    for (long long t = warp; t < tasks; t += nwarps) {
        const int c = (int)(t / N), i = (int)(t % N);
        const int b = c * 32 + lane;
        const int32_t* nb = nbr + (size_t)i * Kp;
        const float* sjb = sjT + b;
        const float* xb = xT + (size_t)c * WP * 32 + lane;
        const float si = siT[(size_t)i * Bs + b];
        float m = -INFINITY;
        int deg = 0;
        for (int k = 0; k < Kp; ++k) {
            const int src = __ldg(nb + k);
            if (src < 0) break;
            ++deg;
            m = fmaxf(m, leaky(si + sjb[(unsigned)src * (unsigned)Bs]));
        }
        float acc[WP];
#pragma unroll
        for (int w = 0; w < WP; ++w) acc[w] = 0.f;
        float sum = 0.f;
#pragma unroll 4
        for (int k = 0; k < deg; ++k) {
            const unsigned src = (unsigned)__ldg(nb + k);
            const float p = expf(leaky(si + sjb[src * (unsigned)Bs]) - m);
            sum += p;
            const float* xs = xb + src * xstride;
#pragma unroll
            for (int w = 0; w < WP; ++w) acc[w] = fmaf(p, xs[w * 32], acc[w]);
        }
        const float linv = 1.f / (sum + GDN_SOFTMAX_EPS);
        mT[(size_t)i * Bs + b] = m;
        linvT[(size_t)i * Bs + b] = linv;
        if (b < B) {
            float* out = A + ((size_t)b * N + i) * W;
            if ((W & 3) == 0) {
#pragma unroll
                for (int w = 0; w < WP; w += 4)
                    if (w < W)
                        *reinterpret_cast<float4*>(out + w) =
                            make_float4(acc[w] * linv, acc[w + 1] * linv, acc[w + 2] * linv, acc[w + 3] * linv);
            } else {
#pragma unroll
                for (int w = 0; w < WP; ++w)
                    if (w < W) out[w] = acc[w] * linv;
            }
            if (alpha != nullptr) {
                float* al = alpha + ((size_t)b * N + i) * Kp;
                for (int k = 0; k < Kp; ++k) {
                    const int src = __ldg(nb + k);
                    al[k] = src < 0 ? 0.f : expf(leaky(si + sjb[(unsigned)src * (unsigned)Bs]) - m) * linv;
                }
            }
        }
    }
}

// alpha only, from a saved ctx (GNNLayer.att_weight_1 on demand)
__global__ void k_attn_alpha(const float* __restrict__ siT, const float* __restrict__ sjT,
                             const float* __restrict__ mT, const float* __restrict__ linvT,
                             const int32_t* __restrict__ nbr, int B, int N, int Kp, int Bs,
                             float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    for (long long t = warp; t < tasks; t += nwarps) {
        const int i = (int)(t / chunks);
        const int b = (int)(t % chunks) * 32 + lane;
        if (b >= B) continue;
        const int32_t* nb = nbr + (size_t)i * Kp;
        const float si = siT[(size_t)i * Bs + b];
        const float m = mT[(size_t)i * Bs + b], linv = linvT[(size_t)i * Bs + b];
        float* al = alpha + ((size_t)b * N + i) * Kp;
        for (int k = 0; k < Kp; ++k) {
            const int src = __ldg(nb + k);
            al[k] = src < 0 ? 0.f : expf(leaky(si + sjT[(size_t)src * Bs + b]) - m) * linv;
        }
    }
}

// ---------------------------------------------------------------------------------------
// backward: g_A -> g_s_i (store), g_s_j (coalesced RED)
//   g_alpha[k] = g_A . x[S_k];  g_l = alpha (g_alpha - dot),  dot = sum_k alpha g_alpha;
//   g_pre = g_l * (pre > 0 ? 1 : 0.2);  g_si = sum_k g_pre;  g_sj[S_k] += g_pre
// dot needs no pass of its own: sum_k alpha_k (g_A . x_k) = g_A . (sum_k alpha_k x_k) = g_A . A, and A
// is saved by the forward.  One sweep over the edges, no stash, no second exp.
// ---------------------------------------------------------------------------------------
template <int WP>
__global__ void __launch_bounds__(256, 2)
k_attn_bwd(const float* __restrict__ xT, const float* __restrict__ siT, const float* __restrict__ sjT,
           const float* __restrict__ mT, const float* __restrict__ linvT,
           const int32_t* __restrict__ nbr, const float* __restrict__ gA, const float* __restrict__ A,
           int B, int N, int W, int Kp, int Bs, float* __restrict__ gsiT, float* __restrict__ gsjT) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    const unsigned xstride = (unsigned)chunks * WP * 32;
    for (long long t = warp; t < tasks; t += nwarps) {
        const int c = (int)(t / N), i = (int)(t % N);
        const int b = c * 32 + lane;
        const bool ok = b < B;
        const int32_t* nb = nbr + (size_t)i * Kp;
        const float* sjb = sjT + b;
        float* gsjb = gsjT + b;
        const float* xb = xT + (size_t)c * WP * 32 + lane;
        const float si = siT[(size_t)i * Bs + b];
        const float m = mT[(size_t)i * Bs + b], linv = linvT[(size_t)i * Bs + b];
        float g[WP];
        const size_t roff = ((size_t)(ok ? b : 0) * N + i) * W;
        float dot = 0.f;
#pragma unroll
        for (int w = 0; w < WP; ++w) {
            g[w] = (ok && w < W) ? gA[roff + w] : 0.f;
            if (ok && w < W) dot = fmaf(g[w], A[roff + w], dot);
        }
        int deg = 0;
        for (int k = 0; k < Kp; ++k) {
            if (__ldg(nb + k) < 0) break;
            ++deg;
        }
        float gsi = 0.f;
#pragma unroll 4
        for (int k = 0; k < deg; ++k) {
            const unsigned src = (unsigned)__ldg(nb + k);
            const float pre = si + sjb[src * (unsigned)Bs];
            const float a = expf(leaky(pre) - m) * linv;
            const float* xs = xb + src * xstride;
            float ga0 = 0.f, ga1 = 0.f;
#pragma unroll
            for (int w = 0; w < WP; w += 2) {
                ga0 = fmaf(g[w], xs[w * 32], ga0);
                ga1 = fmaf(g[w + 1], xs[(w + 1) * 32], ga1);
            }
            const float gl = a * ((ga0 + ga1) - dot);
            const float gp = pre > 0.f ? gl : GDN_NEG_SLOPE * gl;
            gsi += gp;
            // the L2 atomic units are the limiter of this kernel (one op per element): pack four windows
            // into one red.global.add.v4.f32 (padding lanes carry exact zeros)
            const float v1 = __shfl_down_sync(0xffffffffu, gp, 1);
            const float v2 = __shfl_down_sync(0xffffffffu, gp, 2);
            const float v3 = __shfl_down_sync(0xffffffffu, gp, 3);
            if ((lane & 3) == 0)
                atomicAdd(reinterpret_cast<float4*>(gsjb + src * (unsigned)Bs), make_float4(gp, v1, v2, v3));
        }
        gsiT[(size_t)i * Bs + b] = ok ? gsi : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// g_e_i[i] = sum_b g_si, g_e_j[i] = sum_b g_sj,
// g_u_i[w] = sum_{b,i} g_si x[b,i,w], g_u_j likewise -> per-CTA partial [2*32] floats
// ---------------------------------------------------------------------------------------
template <int WP>
__global__ void __launch_bounds__(256)
k_scalar_grads(const float* __restrict__ xT, const float* __restrict__ gsiT, const float* __restrict__ gsjT,
               int N, int W, int Bs, float* __restrict__ gev, float* __restrict__ part) {
    __shared__ float red[8][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    float aui[WP], auj[WP];
#pragma unroll
    for (int w = 0; w < WP; ++w) aui[w] = auj[w] = 0.f;
    for (int i = warp; i < N; i += nwarps) {
        float ssi = 0.f, ssj = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const int b = c * 32 + lane;
            const float gi = gsiT[(size_t)i * Bs + b], gj = gsjT[(size_t)i * Bs + b];
            ssi += gi;
            ssj += gj;
            const float* xs = xT + ((size_t)i * chunks + c) * WP * 32 + lane;
#pragma unroll
            for (int w = 0; w < WP; ++w) {
                const float xv = xs[w * 32];
                aui[w] = fmaf(gi, xv, aui[w]);
                auj[w] = fmaf(gj, xv, auj[w]);
            }
        }
        ssi = warp_sum(ssi);
        ssj = warp_sum(ssj);
        if (lane == 0) { gev[i] = ssi; gev[N + i] = ssj; }
    }
#pragma unroll
    for (int w = 0; w < WP; ++w) {
        const float a = warp_sum(aui[w]), c = warp_sum(auj[w]);
        if (lane == 0) { red[wid][w] = a; red[wid][32 + w] = c; }
    }
    if (WP < 32 && lane == 0)
        for (int w = WP; w < 32; ++w) red[wid][w] = red[wid][32 + w] = 0.f;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[q][threadIdx.x];
        part[(size_t)blockIdx.x * 64 + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------
// embedding-side gradients of the attention scalars:
//   g_ae_i[d] = sum_i V[i,d] g_e_i[i]   (partials [2*D] per CTA)
//   g_V[i,d] (+)= g_e_i[i] ae_i[d] + g_e_j[i] ae_j[d]
// ---------------------------------------------------------------------------------------
__global__ void k_embed_grads(const float* __restrict__ V, const float* __restrict__ gev,
                              const float* __restrict__ ae_i, const float* __restrict__ ae_j,
                              int N, int D, int accumulate, float* __restrict__ gV, float* __restrict__ part) {
    // blockDim.x == D (<= 256): thread <-> channel d, CTA strides over sensors
    const int d = threadIdx.x;
    const float ai = ae_i[d], aj = ae_j[d];
    float si = 0.f, sj = 0.f;
    for (int i = blockIdx.x; i < N; i += gridDim.x) {
        const float gi = gev[i], gj = gev[N + i];
        const float v = V[(size_t)i * D + d];
        si = fmaf(v, gi, si);
        sj = fmaf(v, gj, sj);
        const float add = fmaf(gi, ai, gj * aj);
        float* o = gV + (size_t)i * D + d;
        *o = accumulate ? *o + add : add;
    }
    part[(size_t)blockIdx.x * 2 * D + d] = si;
    part[(size_t)blockIdx.x * 2 * D + D + d] = sj;
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static int grid_for_warps(long long warps_needed, int warps_per_cta, int max_ctas) {
    long long g = (warps_needed + warps_per_cta - 1) / warps_per_cta;
    if (g < 1) g = 1;
    if (g > max_ctas) g = max_ctas;
    return (int)g;
}

int launch_prep(const Shape& s, const float* x, const float* V, const gdn_layer_params* p,
                char* ctx, const CtxLayout& L, cudaStream_t st) {
    float* uv = (float*)(ctx + L.uv);
    float* ev = (float*)(ctx + L.ev);
    k_node_scalars<<<grid_for_warps(s.N, 8, 4 * num_sms()), 256, 0, st>>>(
        V, p->lin_weight, p->att_i, p->att_j, p->att_em_i, p->att_em_j, s.N, s.D, s.W, uv, ev);
    GDN_CHECK_LAUNCH("k_node_scalars");
    const long long tasks = (long long)s.N * (s.Bs / 32);
    k_transpose_scalars<<<grid_for_warps(tasks, 8, 16 * num_sms()), 256, 0, st>>>(
        x, uv, ev, s.B, s.N, s.W, s.WP, s.Bs, ((s.W & 3) == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0, (float*)(ctx + L.xT),
        (float*)(ctx + L.siT), (float*)(ctx + L.sjT));
    GDN_CHECK_LAUNCH("k_transpose_scalars");
    return 0;
}

int launch_attn_fwd(const Shape& s, const int32_t* nbr, char* ctx, const CtxLayout& L, float* alpha,
                    cudaStream_t st) {
    const long long tasks = (long long)s.N * (s.Bs / 32);
    const int grid = grid_for_warps(tasks, 8, 32 * num_sms());
    const float* xT = (const float*)(ctx + L.xT);
    const float* siT = (const float*)(ctx + L.siT);
    const float* sjT = (const float*)(ctx + L.sjT);
    float* A = (float*)(ctx + L.A);
    float* mT = (float*)(ctx + L.mT);
    float* linvT = (float*)(ctx + L.linvT);
#define GDN_LAUNCH_AF(WPV)                                                                          \
    k_attn_fwd<WPV><<<grid, 256, 0, st>>>(xT, siT, sjT, nbr, s.B, s.N, s.W, s.Kp, s.Bs, A, mT, linvT, alpha)
    if (s.WP == 8) GDN_LAUNCH_AF(8);
    else if (s.WP == 16) GDN_LAUNCH_AF(16);
    else GDN_LAUNCH_AF(32);
#undef GDN_LAUNCH_AF
    GDN_CHECK_LAUNCH("k_attn_fwd");
    return 0;
}

int launch_attn_alpha(const Shape& s, const int32_t* nbr, const char* ctx, const CtxLayout& L, float* alpha,
                      cudaStream_t st) {
    const long long tasks = (long long)s.N * (s.Bs / 32);
    k_attn_alpha<<<grid_for_warps(tasks, 8, 32 * num_sms()), 256, 0, st>>>(
        (const float*)(ctx + L.siT), (const float*)(ctx + L.sjT), (const float*)(ctx + L.mT),
        (const float*)(ctx + L.linvT), nbr, s.B, s.N, s.Kp, s.Bs, alpha);
    GDN_CHECK_LAUNCH("k_attn_alpha");
    return 0;
}

// g_A (ws) -> g_siT, g_sjT (ws), g_e (ws.small floats), scalar partials (part region B)
int launch_attn_bwd(const Shape& s, const int32_t* nbr, const char* ctx, const CtxLayout& L,
                    const float* gA, float* gsiT, float* gsjT, float* gev, float* part_u, int* n_part_u,
                    cudaStream_t st) {
    const long long tasks = (long long)s.N * (s.Bs / 32);
    const float* xT = (const float*)(ctx + L.xT);
    const float* siT = (const float*)(ctx + L.siT);
    const float* sjT = (const float*)(ctx + L.sjT);
    const float* mT = (const float*)(ctx + L.mT);
    const float* linvT = (const float*)(ctx + L.linvT);
    cudaError_t e = cudaMemsetAsync(gsjT, 0, (size_t)s.N * s.Bs * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset g_sj");
    const int grid = grid_for_warps(tasks, 8, 32 * num_sms());
    const float* Arows = (const float*)(ctx + L.A);
#define GDN_LAUNCH_AB(WPV)                                                                              \
    k_attn_bwd<WPV><<<grid, 256, 0, st>>>(xT, siT, sjT, mT, linvT, nbr, gA, Arows, s.B, s.N, s.W, s.Kp, s.Bs, gsiT, gsjT)
    if (s.WP == 8) GDN_LAUNCH_AB(8);
    else if (s.WP == 16) GDN_LAUNCH_AB(16);
    else GDN_LAUNCH_AB(32);
#undef GDN_LAUNCH_AB
    GDN_CHECK_LAUNCH("k_attn_bwd");
    const int g2 = grid_for_warps(s.N, 8, 2 * num_sms());
#define GDN_LAUNCH_SG(WPV) k_scalar_grads<WPV><<<g2, 256, 0, st>>>(xT, gsiT, gsjT, s.N, s.W, s.Bs, gev, part_u)
    if (s.WP == 8) GDN_LAUNCH_SG(8);
    else if (s.WP == 16) GDN_LAUNCH_SG(16);
    else GDN_LAUNCH_SG(32);
#undef GDN_LAUNCH_SG
    GDN_CHECK_LAUNCH("k_scalar_grads");
    *n_part_u = g2;
    return 0;
}

int launch_embed_grads(const Shape& s, const float* V, const float* gev, const gdn_layer_params* p,
                       int accumulate, float* gV, float* part, int* n_part, cudaStream_t st) {
    int grid = s.N < 2 * num_sms() ? s.N : 2 * num_sms();
    k_embed_grads<<<grid, s.D, 0, st>>>(V, gev, p->att_em_i, p->att_em_j, s.N, s.D, accumulate, gV, part);
    GDN_CHECK_LAUNCH("k_embed_grads");
    *n_part = grid;
    return 0;
}

}  // namespace gdn
