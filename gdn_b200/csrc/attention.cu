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
// 32 windows (lane <-> b).  x is first transposed to xT[N][Bs/32][WP/4][32][4] (window index next to
// fastest, four taps of a window in one 16-byte word): every gather of a neighbour is WP/4 LDG.128 whose
// 32 lanes cover 512 contiguous bytes, the neighbour index is warp-uniform, and the backward's scatter
// into g_sj is one coalesced RED per edge instead of 32 scattered atomics.
#include <stdlib.h>
#include "common.cuh"
#include "launchers.h"
#include "f32x2.cuh"

namespace gdn {

// ---------------------------------------------------------------------------------------
// prep (ONE launch): u_i = Wl^T a_i, u_j = Wl^T a_j (every CTA recomputes the 2W numbers from the
// L2-resident 4DW-byte weight: cheaper than a launch of its own), e_i[i] = V[i].ae_i, e_j[i] = V[i].ae_j
// (by the warp that owns sensor i), and
//   x[B][N][W] -> xT[N][C][WP][32] (C = Bs/32 window chunks, w padded with zeros to WP),
//   s_iT[N][Bs], s_jT[N][Bs]   (zero padded for b >= B)
// A lane's 16 gathers of one neighbour are then base + w*32: immediate offsets, no address math.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prep(const float* __restrict__ x, const float* __restrict__ V, const float* __restrict__ Wl,
       const float* __restrict__ a_i, const float* __restrict__ a_j,
       const float* __restrict__ ae_i, const float* __restrict__ ae_j,
       int B, int N, int W, int D, int WP, int Bs, int vec4,
       float* __restrict__ xT, float* __restrict__ siT, float* __restrict__ sjT,
       float* __restrict__ uv_out, float* __restrict__ ev_out) {
    __shared__ float su[4][64];
    __shared__ float uv[64];
    const int lane = threadIdx.x & 31;
    {   // thread <-> (d-slice, which, w): four interleaved d-slices, summed in fixed order
        const int slice = threadIdx.x >> 6, which = (threadIdx.x >> 5) & 1, w = lane;
        float acc = 0.f;
        if (w < W) {
            const float* a = which ? a_j : a_i;
            for (int d = slice; d < D; d += 4) acc = fmaf(__ldg(Wl + (size_t)d * W + w), __ldg(a + d), acc);
        }
        su[slice][threadIdx.x & 63] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const float u = (su[0][threadIdx.x] + su[1][threadIdx.x]) + (su[2][threadIdx.x] + su[3][threadIdx.x]);
        uv[threadIdx.x] = u;
        if (blockIdx.x == 0) uv_out[threadIdx.x] = u;
    }
    __syncthreads();
    float aei[8], aej[8];                                   // D <= 256: channels lane, lane + 32, ...
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int d = lane + 32 * q;
        aei[q] = d < D ? __ldg(ae_i + d) : 0.f;
        aej[q] = d < D ? __ldg(ae_j + d) : 0.f;
    }
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    for (long long t = warp; t < tasks; t += nwarps) {
        const int i = (int)(t / chunks), c = (int)(t % chunks);
        const int b = c * 32 + lane;
        float ei = 0.f, ej = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int d = lane + 32 * q;
            if (d < D) {
                const float v = __ldg(V + (size_t)i * D + d);
                ei = fmaf(v, aei[q], ei);
                ej = fmaf(v, aej[q], ej);
            }
        }
        ei = warp_sum(ei);
        ej = warp_sum(ej);
        if (c == 0 && lane == 0) { ev_out[i] = ei; ev_out[N + i] = ej; }
        float si = 0.f, sj = 0.f;
        const bool ok = b < B;
        const float* row = x + ((size_t)b * N + i) * W;
        // xT quad layout: [N][C][WP/4][32 lanes][4 taps] -- a lane's four consecutive taps are one 16-byte word
        float* dst = xT + ((size_t)i * chunks + c) * WP * 32 + lane * 4;
        if (vec4) {
            // 16-byte loads (W % 4 == 0 and x 16-byte aligned): a lane's row is W contiguous floats, rows of different
            // lanes are N*W apart; 16-byte stores land as full 512-byte runs per warp
            for (int w = 0; w < WP; w += 4) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && w < W) q = __ldg(reinterpret_cast<const float4*>(row + w));
                *reinterpret_cast<float4*>(dst + (w >> 2) * 128) = q;
                const float v4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (w + u < W) {
                        si = fmaf(v4[u], uv[w + u], si);
                        sj = fmaf(v4[u], uv[32 + w + u], sj);
                    }
            }
        } else {
            for (int w = 0; w < WP; w += 4) {
                float v4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v4[u] = (ok && w + u < W) ? __ldg(row + w + u) : 0.f;
                    if (w + u < W) {
                        si = fmaf(v4[u], uv[w + u], si);
                        sj = fmaf(v4[u], uv[32 + w + u], sj);
                    }
                }
                *reinterpret_cast<float4*>(dst + (w >> 2) * 128) = make_float4(v4[0], v4[1], v4[2], v4[3]);
            }
        }
        siT[(size_t)i * Bs + b] = ok ? si + ei : 0.f;
        sjT[(size_t)i * Bs + b] = ok ? sj + ej : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// forward: A, segment max m, 1/(segment sum + 1e-16); optional alpha
// tasks are chunk-major (all sensors of window chunk 0, then chunk 1, ...) so that the
// gather working set at any time is one chunk of xT (N*WP*128 bytes)
//
// DPL > 0 (module boundary, D = 32 DPL): the lin transform out = Wl.A + bias
// (models/graph_layer.py:56,71-74) is this kernel's epilogue -- the warp parks its 32 aggregated rows in
// shared memory, re-reads them as broadcast float4 and every lane produces its DPL channels of each row,
// so the n*D output streams to HBM underneath the L2-bound gather instead of in a pass of its own.
// dynamic smem (DPL > 0): WlT [DPL*WP][32] | per warp: A tile [32][WP + 4]
// ---------------------------------------------------------------------------------------
extern __shared__ __align__(16) float attn_smem[];

template <int WP, int DPL, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_attn_fwd(const float* __restrict__ xT, const float* __restrict__ siT, const float* __restrict__ sjT,
           const int32_t* __restrict__ nbr, int B, int N, int W, int Kp, int Bs,
           float* __restrict__ A, float* __restrict__ mT, float* __restrict__ linvT,
           float* __restrict__ alpha,
           const float* __restrict__ Wl, const float* __restrict__ bias, float* __restrict__ out) {
    constexpr int DP = DPL > 0 ? DPL : 1;
    constexpr int AST = WP + 4;                                 // tile row stride: conflict-free float4 stores
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    const long long tasks = (long long)N * chunks;
    const unsigned xstride = (unsigned)chunks * WP * 32;
    float* sWl = attn_smem;
    float* sa = attn_smem + DP * WP * 32 + (threadIdx.x >> 5) * (32 * AST);
    float bs[DP];
    if (DPL > 0) {
        for (int e = threadIdx.x; e < DP * WP * 32; e += blockDim.x) {
            const int l = e & 31, jw = e >> 5, j = jw / WP, w = jw % WP;
            sWl[e] = w < W ? __ldg(Wl + (size_t)(l * DP + j) * W + w) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < DP; ++j) bs[j] = bias != nullptr ? __ldg(bias + lane * DP + j) : 0.f;
        __syncthreads();
    }
    for (long long t = warp; t < tasks; t += nwarps) {
        const int c = (int)(t / N), i = (int)(t % N);
        const int b = c * 32 + lane;
        const int32_t* nb = nbr + (size_t)i * Kp;
        const float* sjb = sjT + b;
        const float4* xb = reinterpret_cast<const float4*>(xT + (size_t)c * WP * 32) + lane;
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
            const float4* xs = xb + src * (xstride >> 2);
#pragma unroll
            for (int q = 0; q < WP / 4; ++q) {
                const float4 v = xs[q * 32];
                acc[4 * q] = fmaf(p, v.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(p, v.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(p, v.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(p, v.w, acc[4 * q + 3]);
            }
        }
        const float linv = 1.f / (sum + GDN_SOFTMAX_EPS);
        mT[(size_t)i * Bs + b] = m;
        linvT[(size_t)i * Bs + b] = linv;
#pragma unroll
        for (int w = 0; w < WP; ++w) acc[w] *= linv;
        if (b < B) {
            float* o = A + ((size_t)b * N + i) * W;
            if ((W & 3) == 0) {
#pragma unroll
                for (int w = 0; w < WP; w += 4)
                    if (w < W) *reinterpret_cast<float4*>(o + w) = make_float4(acc[w], acc[w + 1], acc[w + 2], acc[w + 3]);
            } else {
#pragma unroll
                for (int w = 0; w < WP; ++w)
                    if (w < W) o[w] = acc[w];
            }
            if (alpha != nullptr) {
                float* al = alpha + ((size_t)b * N + i) * Kp;
                for (int k = 0; k < Kp; ++k) {
                    const int src = __ldg(nb + k);
                    al[k] = src < 0 ? 0.f : expf(leaky(si + sjb[(unsigned)src * (unsigned)Bs]) - m) * linv;
                }
            }
        }
        if (DPL > 0) {
            __syncwarp();                                        // the previous task's readers are done with the tile
#pragma unroll
            for (int w = 0; w < WP; w += 4)
                *reinterpret_cast<float4*>(sa + lane * AST + w) = make_float4(acc[w], acc[w + 1], acc[w + 2], acc[w + 3]);
            float wl[DP][WP];
#pragma unroll
            for (int j = 0; j < DP; ++j)
#pragma unroll
                for (int w = 0; w < WP; ++w) wl[j][w] = sWl[(j * WP + w) * 32 + lane];
            __syncwarp();
            const int nbv = (B - c * 32) < 32 ? (B - c * 32) : 32;
            float* orow = out + ((size_t)c * 32 * N + i) * (DP * 32) + lane * DP;
#pragma unroll 2
            for (int bb = 0; bb < nbv; ++bb) {
                float a[WP], z[DP];
                const float4* ap = reinterpret_cast<const float4*>(sa + bb * AST);
#pragma unroll
                for (int q = 0; q < WP / 4; ++q) {
                    const float4 v = ap[q];                      // broadcast
                    a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < DP; ++j) {                   // same chain as k_lin_fwd: bit-identical rows
                    unsigned long long z2 = 0ull;                // (even taps, odd taps): one FFMA2 per pair
#pragma unroll
                    for (int w = 0; w < WP; w += 2)
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(z2) : "l"(pk2(wl[j][w], wl[j][w + 1])), "l"(pk2(a[w], a[w + 1])));
                    float acc0, acc1;
                    upk2(z2, acc0, acc1);
                    z[j] = (acc0 + acc1) + bs[j];
                }
                float* o = orow + (size_t)bb * N * (DP * 32);
                if (DP == 4) *reinterpret_cast<float4*>(o) = make_float4(z[0], z[1 % DP], z[2 % DP], z[3 % DP]);
                else if (DP == 2) *reinterpret_cast<float2*>(o) = make_float2(z[0], z[1 % DP]);
                else {
#pragma unroll
                    for (int j = 0; j < DP; ++j) o[j] = z[j];
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
template <int WP, int MINB>
__global__ void __launch_bounds__(256, MINB)
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
        const float4* xb = reinterpret_cast<const float4*>(xT + (size_t)c * WP * 32) + lane;
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
            const float4* xs = xb + src * (xstride >> 2);
            float ga0 = 0.f, ga1 = 0.f;
#pragma unroll
            for (int q = 0; q < WP / 4; ++q) {                 // same summation order as before: even taps, odd taps
                const float4 v = xs[q * 32];
                ga0 = fmaf(g[4 * q], v.x, ga0);
                ga1 = fmaf(g[4 * q + 1], v.y, ga1);
                ga0 = fmaf(g[4 * q + 2], v.z, ga0);
                ga1 = fmaf(g[4 * q + 3], v.w, ga1);
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
// tail of the backward (ONE launch; it used to be seven): everything that follows the edge sweep.
//   per sensor i (a warp):  g_e_i[i] = sum_b g_si, g_e_j[i] = sum_b g_sj,
//                           g_u_i[w] += sum_b g_si x[b,i,w], g_u_j likewise,
//                           g_ae_i[d] += V[i,d] g_e_i[i], g_ae_j likewise,
//                           g_V[i,d] (+)= g_e_i[i] ae_i[d] + g_e_j[i] ae_j[d]
//   per CTA:                partial records part_u[64], part_e[2D]; a slice of the lin-backward records
//                           [nrec][D*W + D] (doubles) is summed into `sums`
//   last CTA to finish (ticket counter, zeroed together with g_sj):
//                           g_Wl[d,w] = sums[d,w] + a_i[d] g_ui[w] + a_j[d] g_uj[w];  g_bias = sums[D*W + d]
//                           g_a_i[d] = sum_w Wl[d,w] g_ui[w];  g_a_j likewise;  g_ae_i, g_ae_j
// Every sum runs in a fixed order (records ascending): deterministic given its inputs.
// ---------------------------------------------------------------------------------------
template <int WP>
__global__ void __launch_bounds__(256, 2)
k_attn_tail(const float* __restrict__ xT, const float* __restrict__ gsiT, const float* __restrict__ gsjT,
            const float* __restrict__ V, const float* __restrict__ Wl,
            const float* __restrict__ a_i, const float* __restrict__ a_j,
            const float* __restrict__ ae_i, const float* __restrict__ ae_j,
            int N, int W, int D, int Bs, int accumulate,
            const double* __restrict__ part, int nrec,
            float* __restrict__ part_u, float* __restrict__ part_e, double* __restrict__ sums,
            unsigned int* __restrict__ counter,
            float* __restrict__ gV, float* __restrict__ g_Wl, float* __restrict__ g_bias,
            float* __restrict__ g_ai, float* __restrict__ g_aj, float* __restrict__ g_aei, float* __restrict__ g_aej) {
    __shared__ float red[8][64];
    __shared__ __align__(16) float rede[8][512];   // [warp][which * 256 + d], D <= 256
    __shared__ double gu[64];
    __shared__ int last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int chunks = Bs >> 5;
    float aui[WP], auj[WP];
#pragma unroll
    for (int w = 0; w < WP; ++w) aui[w] = auj[w] = 0.f;
    float aei[8], aej[8], esi[8], esj[8];          // channels lane, lane + 32, ...
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int d = lane + 32 * q;
        aei[q] = d < D ? __ldg(ae_i + d) : 0.f;
        aej[q] = d < D ? __ldg(ae_j + d) : 0.f;
        esi[q] = esj[q] = 0.f;
    }
    for (int i = warp; i < N; i += nwarps) {
        // the embedding row (and the gradient row it is added to) is requested before the window sweep: the
        // kernel is a chain of DRAM latencies at one sensor per warp, so every load of a sensor goes out together
        float vq[8], oq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int d = lane + 32 * q;
            vq[q] = d < D ? __ldg(V + (size_t)i * D + d) : 0.f;
            oq[q] = (d < D && accumulate) ? gV[(size_t)i * D + d] : 0.f;
        }
        float ssi = 0.f, ssj = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const int b = c * 32 + lane;
            const float gi = gsiT[(size_t)i * Bs + b], gj = gsjT[(size_t)i * Bs + b];
            ssi += gi;
            ssj += gj;
            const float4* xs = reinterpret_cast<const float4*>(xT + ((size_t)i * chunks + c) * WP * 32) + lane;
#pragma unroll
            for (int q = 0; q < WP / 4; ++q) {
                const float4 v = xs[q * 32];
                const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    aui[4 * q + u] = fmaf(gi, xv[u], aui[4 * q + u]);
                    auj[4 * q + u] = fmaf(gj, xv[u], auj[4 * q + u]);
                }
            }
        }
        ssi = warp_sum(ssi);
        ssj = warp_sum(ssj);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int d = lane + 32 * q;
            esi[q] = fmaf(vq[q], ssi, esi[q]);
            esj[q] = fmaf(vq[q], ssj, esj[q]);
            const float add = fmaf(ssi, aei[q], ssj * aej[q]);
            if (d < D) gV[(size_t)i * D + d] = accumulate ? oq[q] + add : add;
        }
    }
#pragma unroll
    for (int w = 0; w < WP; ++w) {
        const float a = warp_sum(aui[w]), c = warp_sum(auj[w]);
        if (lane == 0) { red[wid][w] = a; red[wid][32 + w] = c; }
    }
    if (WP < 32 && lane == 0)
        for (int w = WP; w < 32; ++w) red[wid][w] = red[wid][32 + w] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        rede[wid][lane + 32 * q] = esi[q];
        rede[wid][256 + lane + 32 * q] = esj[q];
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    if (threadIdx.x < 64) {
        float s = 0.f;
        for (int q = 0; q < nw; ++q) s += red[q][threadIdx.x];
        part_u[(size_t)blockIdx.x * 64 + threadIdx.x] = s;
    }
    for (int e = threadIdx.x; e < 2 * D; e += blockDim.x) {
        const int which = e / D, d = e % D;
        float s = 0.f;
        for (int q = 0; q < nw; ++q) s += rede[q][which * 256 + d];
        part_e[(size_t)blockIdx.x * 2 * D + e] = s;
    }
    // this CTA's slice of the lin-backward records: 16 threads per entry stride the records (all loads in
    // flight at once), then a fixed-order butterfly inside the 16-lane group
    const int rec = D * W + D;
    const int per = (rec + gridDim.x - 1) / gridDim.x;
    for (int k0 = 0; k0 < per; k0 += 16) {
        const int k = k0 + (threadIdx.x >> 4), sub = threadIdx.x & 15;
        const int e = blockIdx.x * per + k;
        double s = 0.0;
        if (k < per && e < rec) {
            const double* col = part + e;
            int q = sub;
            for (; q + 112 < nrec; q += 128) {                 // eight independent loads in flight
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = col[(size_t)(q + 16 * u) * rec];
#pragma unroll
                for (int u = 0; u < 8; ++u) s += t[u];
            }
            for (; q < nrec; q += 16) s += col[(size_t)q * rec];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (sub == 0 && k < per && e < rec) sums[e] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last CTA: the G per-CTA records of part_u [64] and part_e [2D].  thread <-> (four adjacent entries, record
    // phase): 128-bit loads, eight in flight per thread, P = 256 / quads phases combined in fixed order through
    // shared memory (the phase count depends on D only, so the summation order is a function of the shapes)
    const int G = gridDim.x;
    double* sred = reinterpret_cast<double*>(&rede[0][0]);       // <= 1024 doubles = 8 KB: rede is dead by now
    __syncthreads();
    if ((D & 1) == 0) {
        const int nq = (2 * D + 64) >> 2;                        // quads: 16 of part_u, D/2 of part_e (<= 144)
        const int P = blockDim.x / nq;
        const int quad = threadIdx.x % nq, ph = threadIdx.x / nq;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (ph < P) {
            const float4* col = quad < 16 ? reinterpret_cast<const float4*>(part_u) + quad
                                          : reinterpret_cast<const float4*>(part_e) + (quad - 16);
            const size_t stride = quad < 16 ? 16 : (size_t)(D >> 1);      // record pitch in float4
            float4 t[8];
            int q = ph;
            for (; q + 7 * P < G; q += 8 * P) {
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = __ldcg(col + (size_t)(q + u * P) * stride);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    s0 += (double)t[u].x; s1 += (double)t[u].y; s2 += (double)t[u].z; s3 += (double)t[u].w;
                }
            }
            for (; q < G; q += P) {
                const float4 v = __ldcg(col + (size_t)q * stride);
                s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
            }
            double* o = sred + ((size_t)ph * nq + quad) * 4;
            o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 4 * nq; e += blockDim.x) {
            double tot = 0.0;
            for (int h = 0; h < P; ++h) tot += sred[(size_t)h * nq * 4 + e];
            if (e < 64) gu[e] = tot;
            else if (e - 64 < D) g_aei[e - 64] = (float)tot;
            else g_aej[e - 64 - D] = (float)tot;
        }
        __syncthreads();
    } else {                                                     // odd D: scalar loads, four phases per entry
        for (int base = 0; base < 2 * D + 64; base += 64) {
            const int e = base + (threadIdx.x & 63), ph = threadIdx.x >> 6;
            double s = 0.0;
            if (e < 2 * D + 64) {
                const float* col = e < 64 ? part_u + e : part_e + (e - 64);
                const size_t stride = e < 64 ? 64 : (size_t)2 * D;
                for (int q = ph; q < G; q += 4) s += (double)__ldcg(col + (size_t)q * stride);
            }
            sred[ph * 64 + (threadIdx.x & 63)] = s;
            __syncthreads();
            if (threadIdx.x < 64 && e < 2 * D + 64) {
                const double tot = (sred[threadIdx.x] + sred[64 + threadIdx.x]) + (sred[128 + threadIdx.x] + sred[192 + threadIdx.x]);
                if (e < 64) gu[e] = tot;
                else if (e - 64 < D) g_aei[e - 64] = (float)tot;
                else g_aej[e - 64 - D] = (float)tot;
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < rec; e += blockDim.x) {
        double s = __ldcg(sums + e);
        if (e < D * W) {
            const int d = e / W, w = e % W;
            s += (double)a_i[d] * gu[w] + (double)a_j[d] * gu[32 + w];
            g_Wl[e] = (float)s;
        } else if (g_bias != nullptr) {
            g_bias[e - D * W] = (float)s;
        }
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double si = 0.0, sj = 0.0;
        for (int w = 0; w < W; ++w) {
            const double wl = (double)Wl[(size_t)d * W + w];
            si += wl * gu[w];
            sj += wl * gu[32 + w];
        }
        g_ai[d] = (float)si;
        g_aj[d] = (float)sj;
    }
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
    const long long tasks = (long long)s.N * (s.Bs / 32);
    k_prep<<<grid_for_warps(tasks, 8, 16 * num_sms()), 256, 0, st>>>(
        x, V, p->lin_weight, p->att_i, p->att_j, p->att_em_i, p->att_em_j, s.B, s.N, s.W, s.D, s.WP, s.Bs,
        ((s.W & 3) == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0, (float*)(ctx + L.xT), (float*)(ctx + L.siT),
        (float*)(ctx + L.sjT), (float*)(ctx + L.uv), (float*)(ctx + L.ev));
    GDN_CHECK_LAUNCH("k_prep");
    return 0;
}

// p/out != NULL: the lin transform rides as the epilogue when its weight slice fits the registers
// (DPL * WP <= 64: D <= 128 at slide_win <= 16); *fused_out tells the caller whether `out` was written.
int launch_attn_fwd(const Shape& s, const int32_t* nbr, char* ctx, const CtxLayout& L, float* alpha,
                    const gdn_layer_params* p, float* out, int* fused_out, cudaStream_t st) {
    const long long tasks = (long long)s.N * (s.Bs / 32);
    const int grid = grid_for_warps(tasks, 8, 32 * num_sms());
    const float* xT = (const float*)(ctx + L.xT);
    const float* siT = (const float*)(ctx + L.siT);
    const float* sjT = (const float*)(ctx + L.sjT);
    float* A = (float*)(ctx + L.A);
    float* mT = (float*)(ctx + L.mT);
    float* linvT = (float*)(ctx + L.linvT);
    // ... and only while the task count leaves SMs short of warps anyway: the epilogue's weight registers halve the
    // occupancy, which the L2-bound gather of a full-size sweep does not forgive (measured: C4 0.123 -> 0.108 ms
    // for sweep + transform, C5 0.505 -> 0.627)
    static int fuse_env = -1;                     // GDN_FUSED_LIN = 0 never | 1 always | unset: by size
    if (fuse_env < 0) { const char* e = getenv("GDN_FUSED_LIN"); fuse_env = e ? (atoi(e) ? 1 : 0) : 2; }
    const bool fits = p != nullptr && out != nullptr && s.DPL * s.WP <= 64 && (s.DPL == 1 || s.DPL == 2 || s.DPL == 4);
    const bool fuse = fits && (fuse_env == 1 || (fuse_env == 2 && tasks <= 64LL * num_sms()));
    if (fused_out != nullptr) *fused_out = fuse ? 1 : 0;
    // MINB = resident CTAs the register allocation must allow (256 / MINB registers per thread)
#define GDN_LAUNCH_AF(WPV, DPLV, MINB)                                                                             \
    do {                                                                                                          \
        const size_t sm__ = (DPLV) > 0 ? ((size_t)(DPLV) * (WPV) * 32 + 8 * 32 * ((WPV) + 4)) * sizeof(float) : 0; \
        cudaError_t e__ = ensure_dyn_smem(k_attn_fwd<WPV, DPLV, MINB>, sm__);                                     \
        if (e__ != cudaSuccess) return cuda_fail(e__, "smem attribute k_attn_fwd");                               \
        k_attn_fwd<WPV, DPLV, MINB><<<grid, 256, sm__, st>>>(xT, siT, sjT, nbr, s.B, s.N, s.W, s.Kp, s.Bs, A, mT,  \
                                                            linvT, alpha, fuse ? p->lin_weight : nullptr,         \
                                                            fuse ? p->bias : nullptr, fuse ? out : nullptr);      \
    } while (0)
    static int minb16 = -1;                       // diagnostics: GDN_ATTN_MINB = 1 | 2 | 4 (slide_win 9..16, no epilogue)
    if (minb16 < 0) { const char* e = getenv("GDN_ATTN_MINB"); minb16 = e ? atoi(e) : 4; }
    if (!fuse) {
        if (s.WP == 8) GDN_LAUNCH_AF(8, 0, 4);
        else if (s.WP == 16) {
            if (minb16 == 1) GDN_LAUNCH_AF(16, 0, 1);
            else if (minb16 == 2) GDN_LAUNCH_AF(16, 0, 2);
            else GDN_LAUNCH_AF(16, 0, 4);
        } else GDN_LAUNCH_AF(32, 0, 3);
    } else if (s.WP == 8) {
        if (s.DPL == 1) GDN_LAUNCH_AF(8, 1, 2);
        else if (s.DPL == 2) GDN_LAUNCH_AF(8, 2, 2);
        else GDN_LAUNCH_AF(8, 4, 2);
    } else if (s.WP == 16) {
        if (s.DPL == 1) GDN_LAUNCH_AF(16, 1, 2);
        else if (s.DPL == 2) GDN_LAUNCH_AF(16, 2, 2);
        else GDN_LAUNCH_AF(16, 4, 2);
    } else {
        if (s.DPL == 1) GDN_LAUNCH_AF(32, 1, 2);
        else GDN_LAUNCH_AF(32, 2, 2);
    }
#undef GDN_LAUNCH_AF
    GDN_CHECK_LAUNCH(fuse ? "k_attn_fwd_out" : "k_attn_fwd");
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

// g_A (ws) -> g_siT, g_sjT (ws); `zero_bytes` from gsjT on are cleared first (g_sj accumulators + the tail's ticket)
int launch_attn_bwd(const Shape& s, const int32_t* nbr, const char* ctx, const CtxLayout& L,
                    const float* gA, float* gsiT, float* gsjT, size_t zero_bytes, cudaStream_t st) {
    const long long tasks = (long long)s.N * (s.Bs / 32);
    const float* xT = (const float*)(ctx + L.xT);
    const float* siT = (const float*)(ctx + L.siT);
    const float* sjT = (const float*)(ctx + L.sjT);
    const float* mT = (const float*)(ctx + L.mT);
    const float* linvT = (const float*)(ctx + L.linvT);
    cudaError_t e = cudaMemsetAsync(gsjT, 0, zero_bytes, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset g_sj");
    const int grid = grid_for_warps(tasks, 8, 32 * num_sms());
    const float* Arows = (const float*)(ctx + L.A);
#define GDN_LAUNCH_AB(WPV, MINB)                                                                        \
    k_attn_bwd<WPV, MINB><<<grid, 256, 0, st>>>(xT, siT, sjT, mT, linvT, nbr, gA, Arows, s.B, s.N, s.W, s.Kp, s.Bs, gsiT, gsjT)
    static int minb = -1;                         // diagnostics: GDN_ATTN_BWD_MINB = 2 | 3 | 4 (slide_win 9..16)
    if (minb < 0) { const char* e = getenv("GDN_ATTN_BWD_MINB"); minb = e ? atoi(e) : 3; }
    if (s.WP == 8) GDN_LAUNCH_AB(8, 4);
    else if (s.WP == 16) {
        if (minb == 2) GDN_LAUNCH_AB(16, 2);
        else if (minb == 4) GDN_LAUNCH_AB(16, 4);
        else GDN_LAUNCH_AB(16, 3);
    } else GDN_LAUNCH_AB(32, 2);
#undef GDN_LAUNCH_AB
    GDN_CHECK_LAUNCH("k_attn_bwd");
    return 0;
}

// everything after the edge sweep in one launch (see k_attn_tail); part = the lin-backward records
int launch_attn_tail(const Shape& s, const char* ctx, const CtxLayout& L, const float* V, const gdn_layer_params* p,
                     const float* gsiT, const float* gsjT, int accumulate, const double* part, int nrec,
                     float* part_u, float* part_e, double* sums, unsigned int* counter, gdn_layer_grads* g,
                     cudaStream_t st) {
    const float* xT = (const float*)(ctx + L.xT);
    // two CTAs per SM: the kernel is latency-bound.  Small sensor counts still get enough CTAs for the slice
    // reduction of the lin-backward records (eight entries per CTA); CTAs without a sensor write zero records.
    const int rec_ctas = (s.D * s.W + s.D + 7) / 8;
    int grid = grid_for_warps(s.N, 8, tail_max_ctas());
    if (grid < rec_ctas) grid = rec_ctas < tail_max_ctas() ? rec_ctas : tail_max_ctas();
#define GDN_LAUNCH_AT(WPV)                                                                                       \
    k_attn_tail<WPV><<<grid, 256, 0, st>>>(xT, gsiT, gsjT, V, p->lin_weight, p->att_i, p->att_j, p->att_em_i,    \
                                           p->att_em_j, s.N, s.W, s.D, s.Bs, accumulate, part, nrec, part_u,     \
                                           part_e, sums, counter, g->embedding, g->lin_weight, g->bias,          \
                                           g->att_i, g->att_j, g->att_em_i, g->att_em_j)
    if (s.WP == 8) GDN_LAUNCH_AT(8);
    else if (s.WP == 16) GDN_LAUNCH_AT(16);
    else GDN_LAUNCH_AT(32);
#undef GDN_LAUNCH_AT
    GDN_CHECK_LAUNCH("k_attn_tail");
    return 0;
}

}  // namespace gdn
