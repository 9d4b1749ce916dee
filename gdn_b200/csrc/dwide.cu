// dwide.cu -- the D-wide part of the path: lin transform (W -> D), BatchNorm + ReLU,
// (x) embedding, BatchNorm + ReLU, Dropout, Linear(D, 1), and its backward.
//
// Replaces models/graph_layer.py:56,71-74 (lin / head-mean / bias, applied after the
// aggregation because lin has no bias), models/GDN.py:77-79 (GNNLayer BN+ReLU) and
// models/GDN.py:171-187 (head), plus autograd of all of it (SURVEY.md section 8 rows
// a3, a5, a6, a7).  In training ONE D-wide activation reaches HBM: the BatchNorm-1 input
// xh1 = BN1-affine(Wl.A), written by k_fwd_stats2 and read back by k_fwd_out and the three
// backward passes (reading it beats recomputing Wl.A from the W-wide aggregate A four times
// on the FMA pipe); eval recomputes and writes nothing D-wide.  The BatchNorm batch
// statistics -- which force grid-wide reductions -- are obtained
//   BN1: from the first and second moments of A (Z is affine in A, so its per-channel
//        mean/variance follow from mean(A) and cov(A): W + W^2 numbers),
//   BN2: from the k_fwd_stats2 pass.
// The two contractions of the lin backward (g_A = g_z.Wl, g_Wl = g_z^T.A) run on the tensor
// cores (mma_tf32.cuh) for D in {64, 128}, W <= 16; the forward transform stays on the FMA
// pipe on purpose (the MMA's truncating accumulation biases Z, which the
// moment-based BN1 statistics do not forgive).
//
// Thread mapping: lane <-> DPL = D/32 consecutive channels; a warp walks rows.  The fused
// passes are sensor-major (a warp owns sensor i and a range of windows b), so V[i,:] stays
// in registers and the embedding gradient needs no atomics.
#include <stdlib.h>
#include "common.cuh"
#include "launchers.h"
#include "mma_tf32.cuh"
#include "f32x2.cuh"

namespace gdn {

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
template <int DPL, int WP>
__device__ __forceinline__ void load_wl(const float* __restrict__ Wl, int W, int lane, float (&wl)[DPL][WP]) {
#pragma unroll
    for (int j = 0; j < DPL; ++j)
#pragma unroll
        for (int w = 0; w < WP; ++w) wl[j][w] = (w < W) ? Wl[(size_t)(lane * DPL + j) * W + w] : 0.f;
}

template <int WP>
__device__ __forceinline__ void load_arow(const float* __restrict__ row, int W, float (&a)[WP]) {
    if ((W & 3) == 0) {
#pragma unroll
        for (int w = 0; w < WP; w += 4) {
            if (w < W) {
                const float4 v = *reinterpret_cast<const float4*>(row + w);
                a[w] = v.x; a[w + 1] = v.y; a[w + 2] = v.z; a[w + 3] = v.w;
            } else {
                a[w] = a[w + 1] = a[w + 2] = a[w + 3] = 0.f;
            }
        }
    } else {
#pragma unroll
        for (int w = 0; w < WP; ++w) a[w] = (w < W) ? row[w] : 0.f;
    }
}

template <int DPL>
__device__ __forceinline__ void load_chan(const float* __restrict__ p, int lane, float (&v)[DPL]) {
#pragma unroll
    for (int j = 0; j < DPL; ++j) v[j] = p[lane * DPL + j];
}

template <int DPL>
__device__ __forceinline__ void store_chan(float* __restrict__ p, int lane, const float (&v)[DPL]) {
    if (DPL == 4) {
        *reinterpret_cast<float4*>(p + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (DPL == 2) {
        *reinterpret_cast<float2*>(p + lane * 2) = make_float2(v[0], v[1]);
    } else if (DPL == 8) {
        *reinterpret_cast<float4*>(p + lane * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + lane * 8 + 4) = make_float4(v[4 % DPL], v[5 % DPL], v[6 % DPL], v[7 % DPL]);
    } else {
#pragma unroll
        for (int j = 0; j < DPL; ++j) p[lane * DPL + j] = v[j];
    }
}

template <int DPL>
__device__ __forceinline__ void load_chan_vec(const float* __restrict__ p, int lane, float (&v)[DPL]) {
    if (DPL == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p + lane * 4);
        v[0] = t.x; v[1 % DPL] = t.y; v[2 % DPL] = t.z; v[3 % DPL] = t.w;
    } else if (DPL == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p + lane * 2);
        v[0] = t.x; v[1 % DPL] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < DPL; ++j) v[j] = p[lane * DPL + j];
    }
}

// sum_{lanes} v[w] for every w; returns the total of index *w_out on this lane.
// Butterfly reduce-scatter: log2(WP) halving steps, then plain xor-sums over the rest.
template <int WP>
__device__ __forceinline__ float reduce_scatter(float (&v)[WP], int lane, int* w_out) {
    int o = 16, widx = 0;
#pragma unroll
    for (int cnt = WP / 2; cnt >= 1; cnt >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int t = 0; t < cnt; ++t) {
            const float send = up ? v[t] : v[t + cnt];
            const float keep = up ? v[t + cnt] : v[t];
            v[t] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
        if (up) widx += cnt;
        o >>= 1;
    }
    float r = v[0];
    for (; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    *w_out = widx;
    return r;
}

// per-CTA reduction of per-lane channel accumulators over the CTA's warps -> part[] (double)
// vals[q] (q < NV) is the accumulator of channel lane*DPL + (q % DPL), quantity q / DPL.
// Written as part[(q / DPL) * D + lane*DPL + q % DPL].
template <int NV, int DPL>
__device__ __forceinline__ void cta_reduce_channels(const double (&vals)[NV], int D, double* __restrict__ part,
                                                    double* smem /* [warps][NV][32] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) smem[((size_t)wid * NV + q) * 32 + lane] = vals[q];
    __syncthreads();
    for (int e = threadIdx.x; e < NV * 32; e += blockDim.x) {
        const int q = e >> 5, l = e & 31;
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += smem[((size_t)w * NV + q) * 32 + l];
        part[(size_t)(q / DPL) * D + l * DPL + (q % DPL)] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// row staging: a warp pulls up to 32 rows of A (one row per lane, all loads in flight at once)
// into its private shared-memory tile, then walks them with broadcast float4 reads.  This
// replaces one dependent global load per row (8 warps/SM cannot hide that latency) by one
// batch of 32 independent loads.
// ---------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char dyn_smem[];

template <int WP>
struct RowStage {
    static constexpr int STRIDE = WP + 4;                    // 80-byte rows: conflict-free float4 stores
    static constexpr int WARP_FLOATS = 32 * STRIDE;
    static size_t bytes(int warps) { return (size_t)warps * WARP_FLOATS * sizeof(float); }

    __device__ __forceinline__ static float* tile() {
        return reinterpret_cast<float*>(dyn_smem) + (size_t)(threadIdx.x >> 5) * WARP_FLOATS;
    }
    // lane l < nb loads the W floats at first + l * row_stride
    __device__ __forceinline__ static void fill(float* sa, const float* __restrict__ first, size_t row_stride,
                                                int nb, int W, int lane) {
        __syncwarp();
        if (lane < nb) {
            const float* row = first + (size_t)lane * row_stride;
            float* dst = sa + lane * STRIDE;
            if ((W & 3) == 0) {
#pragma unroll
                for (int w = 0; w < WP; w += 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (w < W) v = __ldg(reinterpret_cast<const float4*>(row + w));
                    *reinterpret_cast<float4*>(dst + w) = v;
                }
            } else {
#pragma unroll
                for (int w = 0; w < WP; ++w) dst[w] = (w < W) ? __ldg(row + w) : 0.f;
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ static void get(const float* sa, int rr, float (&a)[WP]) {
        const float4* p = reinterpret_cast<const float4*>(sa + rr * STRIDE);
#pragma unroll
        for (int q = 0; q < WP / 4; ++q) {
            const float4 v = p[q];
            a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
        }
    }
};

// Per-lane prefetch buffer for the saved BatchNorm-1 input xh1[r, :] (training path): lane l copies
// its own DPL channels of up to 32 rows into shared memory with cp.async (all copies in flight at once,
// no registers held), then reads them back itself -- no cross-lane traffic, no barrier beyond the wait.
constexpr int XH_ROWS = 8;       // rows per batch of the buffered passes; two buffers per warp (8 KB at D = 128)
template <int DPL>
struct XhStage {
    static constexpr int BUF_FLOATS = XH_ROWS * DPL * 32;            // one buffer
    static constexpr int SIDE_FLOATS = XH_ROWS * (1 + DPL);          // g_pred + keep words of the batch's rows
    static constexpr int WARP_FLOATS = 2 * (BUF_FLOATS + SIDE_FLOATS);
    __device__ __forceinline__ static float* buf(float* sx, int k) { return sx + (k & 1) * (BUF_FLOATS + SIDE_FLOATS); }
    // asynchronous copy of nb rows of xh1 (row stride row_stride floats) and, if gpred != nullptr, of the
    // rows' g_pred and dropout keep words (r0 = flat row index of the first row, rows N apart); one group
    __device__ __forceinline__ static void fill(float* sx, const float* __restrict__ first, size_t row_stride, int nb,
                                                int lane, const float* __restrict__ gpred, const uint32_t* __restrict__ bits,
                                                size_t r0, size_t rstep) {
        for (int rr = 0; rr < nb; ++rr) {
            const float* src = first + (size_t)rr * row_stride + lane * DPL;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sx + (rr * 32 + lane) * DPL);
            if (DPL == 4) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            } else if (DPL == 8) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 4) : "memory");
            } else if (DPL == 2) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
            } else {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        if (gpred != nullptr && lane < nb) {
            float* side = sx + BUF_FLOATS;
            const size_t r = r0 + (size_t)lane * rstep;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(side + lane);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0), "l"(gpred + r) : "memory");
            if (bits != nullptr) {
                const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(side + XH_ROWS + lane * DPL);
#pragma unroll
                for (int j = 0; j < DPL; ++j)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1 + 4 * j), "l"(bits + r * DPL + j) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ static void commit_empty() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ static void wait_current() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    __device__ __forceinline__ static void get(const float* sx, int rr, int lane, float (&v)[DPL]) {
        load_chan_vec<DPL>(sx + rr * 32 * DPL, lane, v);
    }
};

// ---------------------------------------------------------------------------------------
// module boundary: out[r,:] = Wl.A[r,:] + bias            (models/graph_layer.py:56,71-74)
// ---------------------------------------------------------------------------------------
template <int DPL, int WP>
__global__ void __launch_bounds__(256)
k_lin_fwd(const float* __restrict__ A, const float* __restrict__ Wl, const float* __restrict__ bias,
          long long n, int W, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float* sa = RowStage<WP>::tile();
    float wl[DPL][WP], bs[DPL];
    load_wl<DPL, WP>(Wl, W, lane, wl);
#pragma unroll
    for (int j = 0; j < DPL; ++j) bs[j] = bias ? bias[lane * DPL + j] : 0.f;
    for (long long r0 = warp * 32; r0 < n; r0 += nwarps * 32) {
        const int nb = (int)((n - r0) < 32 ? (n - r0) : 32);
        RowStage<WP>::fill(sa, A + (size_t)r0 * W, (size_t)W, nb, W, lane);
#pragma unroll 2
        for (int rr = 0; rr < nb; ++rr) {
            float a[WP], z[DPL];
            RowStage<WP>::get(sa, rr, a);
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int w = 0; w < WP; w += 2) {
                    acc0 = fmaf(wl[j][w], a[w], acc0);
                    acc1 = fmaf(wl[j][w + 1], a[w + 1], acc1);
                }
                z[j] = (acc0 + acc1) + bs[j];
            }
            store_chan<DPL>(out + (size_t)(r0 + rr) * D, lane, z);
        }
    }
}

// g_out -> g_A[r,w] = sum_d g_out[r,d] Wl[d,w];  partial g_Wl[d,w] += g_out[r,d] A[r,w];
// partial g_bias[d] += g_out[r,d].   part record: [D*W + D] doubles.
//
// A warp stages 32 rows of g_out (coalesced 512-byte rows, all loads in flight) and of A in shared
// memory, then uses the tile twice with two thread mappings and no shuffles:
//   (a) lane <-> channels: g_Wl / g_bias accumulate in registers, A rows broadcast;
//   (b) lane <-> row:      g_A[r, 0:W] = sum_d g[r,d] Wl[d, 0:W], Wl broadcast from shared memory.
// dynamic smem: Wl [D][WP] | per warp: g tile [32][D+4], A tile [32][WP+4]; reduction scratch aliases it.
template <int DPL, int WP>
__global__ void __launch_bounds__(256)
k_lin_bwd(const float* __restrict__ gout, const float* __restrict__ A, const float* __restrict__ Wl,
          long long n, int W, int D, float* __restrict__ gA, double* __restrict__ part) {
    constexpr int DT = DPL * 32;              // == D
    constexpr int GS = DT + 4;                // g tile row stride (floats): conflict-free float4 column reads
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long warp = (long long)blockIdx.x * nw + wid;
    const long long nwarps = (long long)gridDim.x * nw;
    float* sWl = reinterpret_cast<float*>(dyn_smem);                       // [DT][WP]
    float* sG = sWl + DT * WP + (size_t)wid * (32 * GS + RowStage<WP>::WARP_FLOATS);
    float* sa = sG + 32 * GS;
    for (int e = threadIdx.x; e < DT * WP; e += blockDim.x) {
        const int d = e / WP, w = e % WP;
        sWl[e] = w < W ? Wl[(size_t)d * W + w] : 0.f;
    }
    __syncthreads();
    float gwl[DPL][WP], gb[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        gb[j] = 0.f;
#pragma unroll
        for (int w = 0; w < WP; ++w) gwl[j][w] = 0.f;
    }
    for (long long r0 = warp * 32; r0 < n; r0 += nwarps * 32) {
        const int nb = (int)((n - r0) < 32 ? (n - r0) : 32);
        RowStage<WP>::fill(sa, A + (size_t)r0 * W, (size_t)W, nb, W, lane);
        // stage g_out rows: row rr <- 32 lanes x DPL floats, coalesced
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            float go[DPL];
            if (rr < nb) load_chan_vec<DPL>(gout + (size_t)(r0 + rr) * D, lane, go);
            else {
#pragma unroll
                for (int j = 0; j < DPL; ++j) go[j] = 0.f;
            }
            store_chan<DPL>(sG + rr * GS, lane, go);
        }
        __syncwarp();
        // (a) lane <-> channels
#pragma unroll 2
        for (int rr = 0; rr < nb; ++rr) {
            float a[WP], go[DPL];
            RowStage<WP>::get(sa, rr, a);
            load_chan_vec<DPL>(sG + rr * GS, lane, go);
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                gb[j] += go[j];
#pragma unroll
                for (int w = 0; w < WP; ++w) gwl[j][w] = fmaf(go[j], a[w], gwl[j][w]);
            }
        }
        // (b) lane <-> row
        {
            float acc[WP];
#pragma unroll
            for (int w = 0; w < WP; ++w) acc[w] = 0.f;
            const float* grow = sG + lane * GS;
#pragma unroll 2
            for (int d4 = 0; d4 < DT / 4; ++d4) {
                const float4 g4 = *reinterpret_cast<const float4*>(grow + 4 * d4);
                const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4* wrow = reinterpret_cast<const float4*>(sWl + (size_t)(4 * d4 + c) * WP);
#pragma unroll
                    for (int q = 0; q < WP / 4; ++q) {
                        const float4 w4 = wrow[q];                                 // broadcast
                        acc[4 * q] = fmaf(gq[c], w4.x, acc[4 * q]);
                        acc[4 * q + 1] = fmaf(gq[c], w4.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(gq[c], w4.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(gq[c], w4.w, acc[4 * q + 3]);
                    }
                }
            }
            if (lane < nb) {
                float* o = gA + (size_t)(r0 + lane) * W;
                if ((W & 3) == 0) {
#pragma unroll
                    for (int w = 0; w < WP; w += 4)
                        if (w < W) *reinterpret_cast<float4*>(o + w) = make_float4(acc[w], acc[w + 1], acc[w + 2], acc[w + 3]);
                } else {
#pragma unroll
                    for (int w = 0; w < WP; ++w)
                        if (w < W) o[w] = acc[w];
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    float (*red)[33][32] = reinterpret_cast<float (*)[33][32]>(dyn_smem);
    double* prec = part + (size_t)blockIdx.x * ((size_t)D * W + D);
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
#pragma unroll
        for (int w = 0; w < WP; ++w) red[wid][w][lane] = gwl[j][w];
        red[wid][32][lane] = gb[j];
        __syncthreads();
        for (int e = threadIdx.x; e < 33 * 32; e += blockDim.x) {
            const int w = e >> 5, l = e & 31;
            if (w < W || w == 32) {
                double s = 0.0;
                for (int q = 0; q < nw; ++q) s += (double)red[q][w][l];
                const int d = l * DPL + j;
                if (w == 32) prec[(size_t)D * W + d] = s;
                else prec[(size_t)d * W + w] = s;
            }
        }
        __syncthreads();
    }
}

// Tensor-core contractions of the lin backward for D in {64, 128}, W <= 16 (mma.sync m16n8k8, 3xTF32),
// shared by k_lin_bwd_mma (module boundary) and k_bwd3_mma (fused head).  On a tile of TR = 16 rows,
// G [16 x D] (fp32, row stride D + 4) and A [16 x WP] (row stride WP + 4) in shared memory:
//   (1) g_A tile [16 x WP]   = G . Wl            (A fragments by ldmatrix, Wl pre-split {hi, lo} in smem)
//   (2) g_Wl^T  [WP x D]    += A^T . G           (accumulators stay in registers for the whole kernel)
// The k index of (2) walks the rows in the order 0,2,4,6,1,3,5,7 so that the B-fragment reads of G are
// bank-conflict free with the same row stride that makes the ldmatrix reads of (1) conflict free.
template <int DPL, int WP>
struct LinBwdMma {
    static constexpr int DT = DPL * 32, GS = DT + 4, TR = 16;
    static constexpr int WS = WP + 4;             // Wl row stride in 8-byte {hi, lo} pairs: conflict-free LDS.64
    static constexpr int AS = RowStage<WP>::STRIDE;
    static constexpr int NT1 = WP / 8, NT2 = DT / 8;
    static_assert(WP == 8 || WP == 16, "LinBwdMma: W <= 16");
    static_assert(DPL == 2 || DPL == 4, "LinBwdMma: D in {64, 128}");
    static constexpr size_t W_BYTES = (size_t)DT * WS * sizeof(uint2);
    static constexpr size_t G_BYTES = (size_t)TR * GS * sizeof(float);

    // all threads of the CTA; caller synchronises
    __device__ __forceinline__ static void fill_w(uint2* sW, const float* __restrict__ Wl, int W) {
        for (int e = threadIdx.x; e < DT * WP; e += blockDim.x) {
            const int d = e / WP, w = e % WP;
            uint2 v;
            split_tf32_rn(w < W ? Wl[(size_t)d * W + w] : 0.f, v.x, v.y);
            sW[d * WS + w] = v;
        }
    }
    // (1): c1[nt] holds rows (g, g + 8) x columns (8 nt + 2t, + 1) of the tile's g_A
    __device__ __forceinline__ static void ga_tile(const float* sG, const uint2* sW, int lane, float (&c1)[NT1][4]) {
        const int g = lane >> 2, t = lane & 3;
        // ldmatrix row address: matrix j = lane >> 3 -> rows (j & 1) * 8 + (lane & 7), columns + (j >> 1) * 4
        const float* ldm = sG + (((lane >> 3) & 1) * 8 + (lane & 7)) * GS + (lane >> 4) * 4;
        const uint2* wb = sW + t * WS + g;
        float c0[NT1][4];                         // two accumulator sets: shorter dependent MMA chains
#pragma unroll
        for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) c0[nt][q] = c1[nt][q] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < DT / 8; ks += 2) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint32_t raw[4], ah[4], al[4];
                ldmatrix_x4(raw, ldm + 8 * (ks + u));
#pragma unroll
                for (int q = 0; q < 4; ++q) split_tf32(__uint_as_float(raw[q]), ah[q], al[q]);
#pragma unroll
                for (int nt = 0; nt < NT1; ++nt) {
                    const uint2 b0 = wb[(8 * (ks + u)) * WS + 8 * nt];
                    const uint2 b1 = wb[(8 * (ks + u) + 4) * WS + 8 * nt];
                    if (u == 0) mma_3xtf32(c0[nt], ah, al, b0.x, b1.x, b0.y, b1.y);
                    else mma_3xtf32(c1[nt], ah, al, b0.x, b1.x, b0.y, b1.y);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) c1[nt][q] += c0[nt][q];
    }
    // (2): acc2[nt] holds rows w = (g, g + 8) x columns d = (8 nt + 2t, + 1) of g_Wl^T
    __device__ __forceinline__ static void gwl_acc(const float* sG, const float* sa, int lane, float (&acc2)[NT2][4]) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int ks = 0; ks < TR / 8; ++ks) {
            const int ra = 8 * ks + 2 * t, rb = ra + 1;
            uint32_t ah[4], al[4];
            split_tf32(sa[ra * AS + g], ah[0], al[0]);
            split_tf32(sa[rb * AS + g], ah[2], al[2]);
            if (WP == 16) {
                split_tf32(sa[ra * AS + g + 8], ah[1], al[1]);
                split_tf32(sa[rb * AS + g + 8], ah[3], al[3]);
            } else {
                ah[1] = al[1] = ah[3] = al[3] = 0u;
            }
            const float* ga = sG + ra * GS + g;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(ga[8 * nt], bh0, bl0);
                split_tf32(ga[GS + 8 * nt], bh1, bl1);
                mma_3xtf32(acc2[nt], ah, al, bh0, bh1, bl0, bl1);
            }
        }
    }
    // The tensor core adds into its fp32 accumulator with truncation, so a long chain of MMAs on one
    // accumulator leaks towards zero (measured: 7e-5 on g_Wl after ~330 chained MMAs at C5).  The
    // register accumulators are therefore flushed every 4 tiles into a CTA-wide fp32 image in shared
    // memory (red.shared.add.f32, round-to-nearest) and restarted from zero.
    static constexpr int ACS = DT + 8;                                    // accumulator image row stride
    static constexpr size_t ACC_BYTES = (size_t)(WP + 1) * ACS * sizeof(float);
    __device__ __forceinline__ static void zero_acc(float* sAcc) {        // all threads; caller synchronises
        for (int e = threadIdx.x; e < (WP + 1) * ACS; e += blockDim.x) sAcc[e] = 0.f;
    }
    __device__ __forceinline__ static void flush(float (&acc2)[NT2][4], float* sAcc, int lane) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) {
            float* r0 = sAcc + g * ACS + 8 * nt + 2 * t;
            atomicAdd(r0, acc2[nt][0]);
            atomicAdd(r0 + 1, acc2[nt][1]);
            if (WP == 16) {
                atomicAdd(r0 + 8 * ACS, acc2[nt][2]);
                atomicAdd(r0 + 8 * ACS + 1, acc2[nt][3]);
            }
            acc2[nt][0] = acc2[nt][1] = acc2[nt][2] = acc2[nt][3] = 0.f;
        }
    }
    // end of kernel (every warp has flushed): bias-gradient partials join the image, which becomes the CTA's
    // part record [D*W + D] doubles
    __device__ __forceinline__ static void finish(const float (&gb)[DPL], float* sAcc, int W, double* __restrict__ prec) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < DPL; ++j) atomicAdd(sAcc + WP * ACS + lane * DPL + j, gb[j]);
        __syncthreads();
        for (int e = threadIdx.x; e < (WP + 1) * DT; e += blockDim.x) {
            const int w = e / DT, d = e % DT;
            if (w == WP) prec[(size_t)DT * W + d] = (double)sAcc[w * ACS + d];
            else if (w < W) prec[(size_t)d * W + w] = (double)sAcc[w * ACS + d];
        }
    }
};

// module boundary, tensor-core version: a warp stages 16 rows of g_out and of A per tile.
// dynamic smem: Wl {hi,lo} [D][WP+4] | per warp: G tile [16][D+4], A tile [16][WP+4]; reduction scratch aliases it.
template <int DPL, int WP>
__global__ void __launch_bounds__(256, 2)
k_lin_bwd_mma(const float* __restrict__ gout, const float* __restrict__ A, const float* __restrict__ Wl,
              long long n, int W, int D, float* __restrict__ gA, double* __restrict__ part, int flush_every) {
    using M = LinBwdMma<DPL, WP>;
    constexpr int TR = M::TR, GS = M::GS, AS = M::AS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long warp = (long long)blockIdx.x * nw + wid;
    const long long nwarps = (long long)gridDim.x * nw;
    uint2* sW = reinterpret_cast<uint2*>(dyn_smem);
    float* sAcc = reinterpret_cast<float*>(dyn_smem + M::W_BYTES);
    float* sG = reinterpret_cast<float*>(dyn_smem + M::W_BYTES + M::ACC_BYTES) + (size_t)wid * (TR * GS + TR * AS);
    float* sa = sG + TR * GS;
    M::fill_w(sW, Wl, W);
    M::zero_acc(sAcc);
    __syncthreads();
    int since_flush = 0;
    float acc2[M::NT2][4], gb[DPL];
#pragma unroll
    for (int nt = 0; nt < M::NT2; ++nt) acc2[nt][0] = acc2[nt][1] = acc2[nt][2] = acc2[nt][3] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) gb[j] = 0.f;
    for (long long r0 = warp * TR; r0 < n; r0 += nwarps * TR) {
        const int nb = (int)((n - r0) < TR ? (n - r0) : TR);
        __syncwarp();
        if (lane < TR) {                                                      // A rows (zero beyond the tail)
            float* dst = sa + lane * AS;
            const float* row = A + (size_t)(r0 + lane) * W;
            if (lane < nb && (W & 3) == 0) {
#pragma unroll
                for (int w = 0; w < WP; w += 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (w < W) v = __ldg(reinterpret_cast<const float4*>(row + w));
                    *reinterpret_cast<float4*>(dst + w) = v;
                }
            } else {
#pragma unroll
                for (int w = 0; w < WP; ++w) dst[w] = (lane < nb && w < W) ? __ldg(row + w) : 0.f;
            }
        }
#pragma unroll 8
        for (int rr = 0; rr < TR; ++rr) {                                     // G rows, coalesced
            float go[DPL];
#pragma unroll
            for (int j = 0; j < DPL; ++j) go[j] = 0.f;
            if (rr < nb) load_chan_vec<DPL>(gout + (size_t)(r0 + rr) * D, lane, go);
            store_chan<DPL>(sG + rr * GS, lane, go);
#pragma unroll
            for (int j = 0; j < DPL; ++j) gb[j] += go[j];
        }
        __syncwarp();
        float c1[M::NT1][4];
        M::ga_tile(sG, sW, lane, c1);
#pragma unroll
        for (int nt = 0; nt < M::NT1; ++nt) {
            const int w0 = 8 * nt + 2 * t;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int rr = g + 8 * hh;
                if (rr < nb && w0 < W) {
                    float* o = gA + (size_t)(r0 + rr) * W + w0;
                    if ((W & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(c1[nt][2 * hh], c1[nt][2 * hh + 1]);
                    else {
                        o[0] = c1[nt][2 * hh];
                        if (w0 + 1 < W) o[1] = c1[nt][2 * hh + 1];
                    }
                }
            }
        }
        M::gwl_acc(sG, sa, lane, acc2);
        if (++since_flush == flush_every) {                                   // 24 chained MMAs per accumulator at 4
            M::flush(acc2, sAcc, lane);
            since_flush = 0;
        }
    }
    if (since_flush) M::flush(acc2, sAcc, lane);
    __syncthreads();
    M::finish(gb, sAcc, W, part + (size_t)blockIdx.x * ((size_t)D * W + D));
}

// ---------------------------------------------------------------------------------------
// moments of A: sum_r A[r,w], sum_r A[r,w] A[r,w']   -> part record [W*W + W] doubles
// lane <-> w' (WP lanes), 32/WP rows in flight per warp
// ---------------------------------------------------------------------------------------
template <int WP>
__global__ void __launch_bounds__(256)
k_moments(const float* __restrict__ A, long long n, int W, double* __restrict__ part) {
    constexpr int NW = WP == 32 ? 2 : 8;        // warps per CTA (static smem stays < 48 KB)
    __shared__ double red[NW][WP + 1][WP];
    __shared__ __align__(16) float stage[NW][RowStage<WP>::WARP_FLOATS];
    constexpr int SLOTS = 32 / WP;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int slot = lane / WP, wq = lane % WP;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float* sa = stage[wid];
    float acc[WP];
#pragma unroll
    for (int w = 0; w < WP; ++w) acc[w] = 0.f;
    float s1 = 0.f;
    double dacc[WP], ds1 = 0.0;
#pragma unroll
    for (int w = 0; w < WP; ++w) dacc[w] = 0.0;
    // 32 rows per batch staged in shared memory (one coalesced burst), then lane (slot, w') walks the
    // rows of its slot: fp32 partials are flushed into fp64 after every batch
    for (long long r0 = warp * 32; r0 < n; r0 += nwarps * 32) {
        const int nb = (int)((n - r0) < 32 ? (n - r0) : 32);
        RowStage<WP>::fill(sa, A + (size_t)r0 * W, (size_t)W, nb, W, lane);
        for (int rr = slot; rr < nb; rr += SLOTS) {
            float a[WP];
            RowStage<WP>::get(sa, rr, a);
            const float own = sa[rr * RowStage<WP>::STRIDE + wq];
            s1 += own;
#pragma unroll
            for (int w = 0; w < WP; ++w) acc[w] = fmaf(a[w], own, acc[w]);
        }
#pragma unroll
        for (int w = 0; w < WP; ++w) { dacc[w] += (double)acc[w]; acc[w] = 0.f; }
        ds1 += (double)s1; s1 = 0.f;
    }
#pragma unroll
    for (int w = 0; w < WP; ++w) dacc[w] += (double)acc[w];
    ds1 += (double)s1;
    // fold the row slots of the warp together
#pragma unroll
    for (int o = WP; o < 32; o <<= 1) {
#pragma unroll
        for (int w = 0; w < WP; ++w) dacc[w] += __shfl_xor_sync(0xffffffffu, dacc[w], o);
        ds1 += __shfl_xor_sync(0xffffffffu, ds1, o);
    }
    if (slot == 0) {
#pragma unroll
        for (int w = 0; w < WP; ++w) red[wid][w][wq] = dacc[w];
        red[wid][WP][wq] = ds1;
    }
    __syncthreads();
    double* prec = part + (size_t)blockIdx.x * ((size_t)W * W + W);
    for (int e = threadIdx.x; e < (WP + 1) * WP; e += blockDim.x) {
        const int w = e / WP, q = e % WP;
        if (q < W && (w < W || w == WP)) {
            double s = 0.0;
            for (int k = 0; k < NW; ++k) s += red[k][w][q];
            if (w == WP) prec[(size_t)W * W + q] = s;
            else prec[(size_t)w * W + q] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------
// fused head passes (sensor-major).  bnc = ctx.bn: 8 vectors of D floats:
//   0 mean1  1 istd1  2 k1a (= istd1)  3 k1b (= (bias - mean1) istd1)
//   4 mean2  5 istd2  6 k2a (= istd2)  7 k2b (= -mean2 istd2)
// xh1 = z' k1a + k1b with z' = Wl.A;  y1 = g1 xh1 + be1;  r1 = relu(y1);  p = r1 V[i]
// xh2 = p k2a + k2b;  y2 = g2 xh2 + be2;  h2 = relu(y2);  hm = h2 * keep * scale
// ---------------------------------------------------------------------------------------

template <int DPL, int WP>
struct RowEval {
    float wl[DPL][WP];
    float k1a[DPL], k1b[DPL], g1[DPL], be1[DPL];
    __device__ __forceinline__ void init(const HeadArgs& h, int lane, bool need_wl = true) {
        if (need_wl) load_wl<DPL, WP>(h.Wl, h.W, lane, wl);
        load_chan<DPL>(h.bnc + 2 * h.D, lane, k1a);
        load_chan<DPL>(h.bnc + 3 * h.D, lane, k1b);
        load_chan<DPL>(h.g1, lane, g1);
        load_chan<DPL>(h.be1, lane, be1);
    }
    // xh1, y1 for the lane's channels
    __device__ __forceinline__ void eval(const float (&a)[WP], float (&xh1)[DPL], float (&y1)[DPL]) const {
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            // two interleaved partial sums over even / odd window taps: one FFMA2 per pair of taps
            unsigned long long z2 = 0ull;
#pragma unroll
            for (int w = 0; w < WP; w += 2)
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(z2) : "l"(pk2(wl[j][w], wl[j][w + 1])), "l"(pk2(a[w], a[w + 1])));
            float z0, z1;
            upk2(z2, z0, z1);
            const float z = z0 + z1;
            xh1[j] = fmaf(z, k1a[j], k1b[j]);
            y1[j] = fmaf(g1[j], xh1[j], be1[j]);
        }
    }
    // same, xh1 taken from the buffer the forward saved
    __device__ __forceinline__ void from_saved(const float (&xh1)[DPL], float (&y1)[DPL]) const {
#pragma unroll
        for (int j = 0; j < DPL; ++j) y1[j] = fmaf(g1[j], xh1[j], be1[j]);
    }
};

// inputs of one row for the D-wide passes: A row (if the pass needs it) and (xh1, y1), either
// recomputed from A or read back from the saved buffer
template <int DPL, int WP, bool NEED_A, bool BUF>
__device__ __forceinline__ void row_inputs(const RowEval<DPL, WP>& re, const float* sa, const float* sx, int rr, int lane,
                                           float (&a)[WP], float (&xh1)[DPL], float (&y1)[DPL]) {
    if (NEED_A) RowStage<WP>::get(sa, rr, a);
    if (BUF) {
        XhStage<DPL>::get(sx, rr, lane, xh1);
        re.from_saved(xh1, y1);
    } else {
        re.eval(a, xh1, y1);
    }
}

#define GDN_TASK_LOOP_BEGIN(h)                                                                   \
    const int lane = threadIdx.x & 31;                                                           \
    const long long warp_ = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;             \
    const long long nwarps_ = ((long long)gridDim.x * blockDim.x) >> 5;                          \
    const long long tasks_ = (long long)(h).N * (h).S;                                           \
    const size_t rstride_ = (size_t)(h).N * (h).W;                                               \
    constexpr int WARP_SMEM_ = (NEED_A_ ? RowStage<WP>::WARP_FLOATS : 0) + (BUF_ ? XhStage<DPL>::WARP_FLOATS : 0); \
    float* sa_ = reinterpret_cast<float*>(dyn_smem) + (size_t)(threadIdx.x >> 5) * WARP_SMEM_;   \
    float* sxbase_ = sa_ + (NEED_A_ ? RowStage<WP>::WARP_FLOATS : 0);                            \
    int pk_ = 0;             /* buffered passes: batch counter (buffer parity) */                \
    bool pf_ = false;        /* ... and whether the coming batch is already in flight */         \
    for (long long task_ = warp_; task_ < tasks_; task_ += nwarps_) {                            \
        const int i = (int)(task_ / (h).S), sp = (int)(task_ % (h).S);                           \
        const int b_lo = sp * (h).rps, b_hi = min((h).B, b_lo + (h).rps);
#define GDN_TASK_LOOP_END }
// inside a task: batches of up to 32 windows of sensor i (8 in the buffered passes, whose staging is
// double-buffered: while a batch is processed the next one -- of this task or the warp's next task -- is
// already in flight; gp_/bits_ = g_pred and keep-word arrays staged with it, or nullptr)
#define GDN_BATCH_LOOP_BEGIN2(h, gp_, bits_)                                                     \
    for (int b0 = b_lo; b0 < b_hi; b0 += (BUF_ ? XH_ROWS : 32)) {                                \
        const int nb = min(BUF_ ? XH_ROWS : 32, b_hi - b0);                                      \
        float* sx_ = sxbase_;                                                                    \
        if (BUF_) {                                                                              \
            sx_ = XhStage<DPL>::buf(sxbase_, pk_);                                               \
            const size_t xrow_ = (size_t)(h).N * (h).D;                                          \
            if (!pf_) {                                                                          \
                __syncwarp();                                                                    \
                XhStage<DPL>::fill(sx_, (h).xh1 + ((size_t)b0 * (h).N + i) * (h).D, xrow_, nb, lane, gp_, bits_, \
                                   (size_t)b0 * (h).N + i, (size_t)(h).N);                       \
            }                                                                                    \
            int ni_ = i, nb0_ = b0 + XH_ROWS, nhi_ = b_hi;                                       \
            bool nh_ = true;                                                                     \
            if (nb0_ >= b_hi) {                                                                  \
                const long long t2_ = task_ + nwarps_;                                           \
                nh_ = t2_ < tasks_;                                                              \
                if (nh_) {                                                                       \
                    ni_ = (int)(t2_ / (h).S);                                                    \
                    nb0_ = (int)(t2_ % (h).S) * (h).rps;                                         \
                    nhi_ = min((h).B, nb0_ + (h).rps);                                           \
                }                                                                                \
            }                                                                                    \
            __syncwarp();                                                                        \
            if (nh_) XhStage<DPL>::fill(XhStage<DPL>::buf(sxbase_, pk_ + 1), (h).xh1 + ((size_t)nb0_ * (h).N + ni_) * (h).D, \
                                        xrow_, min(XH_ROWS, nhi_ - nb0_), lane, gp_, bits_,        \
                                        (size_t)nb0_ * (h).N + ni_, (size_t)(h).N);              \
            else XhStage<DPL>::commit_empty();                                                   \
            pf_ = nh_;                                                                           \
            ++pk_;                                                                               \
            XhStage<DPL>::wait_current();                                                        \
            __syncwarp();                                                                        \
        }                                                                                        \
        if (NEED_A_) RowStage<WP>::fill(sa_, (h).A + ((size_t)b0 * (h).N + i) * (h).W, rstride_, nb, (h).W, lane);
#define GDN_BATCH_LOOP_BEGIN(h) GDN_BATCH_LOOP_BEGIN2(h, ((const float*)nullptr), ((const uint32_t*)nullptr))
#define GDN_BATCH_LOOP_END }

// BN2 batch statistics: sum_r p, sum_r p^2 -> part record [2*D] doubles
template <int DPL, int WP>
__global__ void __launch_bounds__(256, (DPL <= 4 && DPL * WP <= 64) ? 2 : 1)
k_fwd_stats2(HeadArgs h, double* __restrict__ part) {
    constexpr bool NEED_A_ = true, BUF_ = false;
    RowEval<DPL, WP> re;
    re.init(h, threadIdx.x & 31);
    double acc[2 * DPL];
#pragma unroll
    for (int q = 0; q < 2 * DPL; ++q) acc[q] = 0.0;
    GDN_TASK_LOOP_BEGIN(h)
        float v[DPL], s1[DPL], s2[DPL];
        load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
#pragma unroll
        for (int j = 0; j < DPL; ++j) s1[j] = s2[j] = 0.f;
        GDN_BATCH_LOOP_BEGIN(h)
#pragma unroll 2
            for (int rr = 0; rr < nb; ++rr) {
                float a[WP], xh1[DPL], y1[DPL];
                row_inputs<DPL, WP, NEED_A_, BUF_>(re, sa_, sx_, rr, lane, a, xh1, y1);
                if (h.xh1 != nullptr) store_chan<DPL>(h.xh1 + ((size_t)(b0 + rr) * h.N + i) * h.D, lane, xh1);
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const float p = fmaxf(y1[j], 0.f) * v[j];
                    s1[j] += p;
                    s2[j] = fmaf(p, p, s2[j]);
                }
            }
        GDN_BATCH_LOOP_END
#pragma unroll
        for (int j = 0; j < DPL; ++j) { acc[j] += (double)s1[j]; acc[DPL + j] += (double)s2[j]; }
    GDN_TASK_LOOP_END
    __syncthreads();
    cta_reduce_channels<2 * DPL, DPL>(acc, h.D, part + (size_t)blockIdx.x * 2 * h.D, reinterpret_cast<double*>(dyn_smem));
}

// Dropout keep bits of one row, produced by the lane that owns the row (32 rows of a batch are
// generated in parallel by the 32 lanes).  Word j holds channel slot j of every lane: bit l <->
// channel l*DPL + j.  Explicit mask (tests): bit = mask != 0.  Otherwise Philox4x32-10: one call
// yields eight 16-bit uniforms, keep iff u16 >= round(p * 65536); counter = (row*DPL + j)*4 + q.
template <int DPL>
__device__ __forceinline__ uint32_t gen_keep_word(const HeadArgs& h, size_t r, int j) {
    uint32_t w = 0u;
    if (h.mask != nullptr) {
        const float* m = h.mask + r * h.D;
        for (int l = 0; l < 32; ++l) w |= (m[l * DPL + j] != 0.f) ? (1u << l) : 0u;
    } else {
        const uint32_t thr16 = (uint32_t)(h.p_drop * 65536.f + 0.5f);
        const unsigned long long offset = h.offset + (h.offset_dev ? *h.offset_dev : 0ull);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 rnd = philox4x32_10((unsigned long long)(r * DPL + j) * 4ull + q, offset, h.seed);
            const uint32_t x[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                w |= ((x[c] & 0xffffu) >= thr16 ? 1u : 0u) << (q * 8 + c * 2);
                w |= ((x[c] >> 16) >= thr16 ? 1u : 0u) << (q * 8 + c * 2 + 1);
            }
        }
    }
    return w;
}

// pred[b,i] = sum_d hm[d] wo[d] + bo;  training: dropout + keep bits saved
template <int DPL, int WP, bool BUF_>
__global__ void __launch_bounds__(256)
k_fwd_out(HeadArgs h, float* __restrict__ pred) {
    constexpr bool NEED_A_ = !BUF_;
    RowEval<DPL, WP> re;
    re.init(h, threadIdx.x & 31, NEED_A_);
    float k2a[DPL], k2b[DPL], g2[DPL], be2[DPL], wo[DPL];
    {
        const int ln = threadIdx.x & 31;
        load_chan<DPL>(h.bnc + 6 * h.D, ln, k2a);
        load_chan<DPL>(h.bnc + 7 * h.D, ln, k2b);
        load_chan<DPL>(h.g2, ln, g2);
        load_chan<DPL>(h.be2, ln, be2);
        load_chan<DPL>(h.wo, ln, wo);
    }
    const float bo = h.bo[0];
    const bool drop = h.training && h.p_drop > 0.f;
    GDN_TASK_LOOP_BEGIN(h)
        float v[DPL];
        load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
        GDN_BATCH_LOOP_BEGIN(h)
            float my_pred = 0.f;
            // keep words of the batch: in the buffered passes a batch is XH_ROWS = 8 rows, so four lanes share a
            // row (lane l: row l & 7, words [(l >> 3) * H, +H)) and all 32 lanes run the generator
            constexpr bool SPLIT = BUF_;
            constexpr int LPR = 32 / XH_ROWS;                        // lanes per row
            constexpr int H = SPLIT ? (DPL + LPR - 1) / LPR : DPL;   // words per lane
            const int brow = SPLIT ? (lane % XH_ROWS) : lane, bpart = SPLIT ? (lane / XH_ROWS) : 0;
            uint32_t my_bits[H];
#pragma unroll
            for (int jj = 0; jj < H; ++jj) my_bits[jj] = 0xffffffffu;
            if (drop && brow < nb) {
#pragma unroll
                for (int jj = 0; jj < H; ++jj)
                    if (bpart * H + jj < DPL) my_bits[jj] = gen_keep_word<DPL>(h, (size_t)(b0 + brow) * h.N + i, bpart * H + jj);
            }
            // one row: the lane's partial of pred (sum over its DPL channels), packed chain
            auto row_partial = [&](int rr) -> float {
                float a[WP], xh1[DPL], y1[DPL], kf[DPL], r1[DPL], p[DPL], xh2[DPL], y2[DPL], r2[DPL], wk[DPL], dv[DPL];
                row_inputs<DPL, WP, NEED_A_, BUF_>(re, sa_, sx_, rr, lane, a, xh1, y1);
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const uint32_t word = __shfl_sync(0xffffffffu, my_bits[j % H], SPLIT ? rr + XH_ROWS * (j / H) : rr);
                    kf[j] = drop ? (((word >> lane) & 1u) ? h.scale : 0.f) : 1.f;
                }
                vrelu<DPL>(r1, y1);
                vmul<DPL>(p, r1, v);
                vfma<DPL>(xh2, p, k2a, k2b);
                vfma<DPL>(y2, g2, xh2, be2);
                vrelu<DPL>(r2, y2);
                vmul<DPL>(wk, wo, kf);
                vmul<DPL>(dv, r2, wk);
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < DPL; ++j) dot += dv[j];
                return dot;
            };
            if (BUF_) {
                // the batch's XH_ROWS row sums in one butterfly reduce-scatter (9 shuffles instead of 5 per row)
                float dots[XH_ROWS];
#pragma unroll
                for (int rr = 0; rr < XH_ROWS; ++rr) dots[rr] = rr < nb ? row_partial(rr) : 0.f;   // nb is warp-uniform
                int ridx;
                const float tot = reduce_scatter<XH_ROWS>(dots, lane, &ridx);
                // lanes 4r .. 4r+3 now hold the total of row r = ridx; lane r fetches it for the coalesced store below
                const float mine = __shfl_sync(0xffffffffu, tot, (lane & (XH_ROWS - 1)) * (32 / XH_ROWS));
                (void)ridx;
                if (lane < XH_ROWS) my_pred = mine + bo;
            } else {
#pragma unroll 2
                for (int rr = 0; rr < nb; ++rr) {
                    const float dot = warp_sum(row_partial(rr));
                    if (lane == rr) my_pred = dot + bo;
                }
            }
            if (lane < nb) pred[(size_t)(b0 + lane) * h.N + i] = my_pred;
            if (drop && brow < nb) {
                const size_t r = (size_t)(b0 + brow) * h.N + i;
#pragma unroll
                for (int jj = 0; jj < H; ++jj)
                    if (bpart * H + jj < DPL) h.bits[r * DPL + bpart * H + jj] = my_bits[jj];
            }
        GDN_BATCH_LOOP_END
    GDN_TASK_LOOP_END
}

// shared recompute of the chain for the backward passes
template <int DPL, int WP>
struct BwdRow {
    float k2a[DPL], k2b[DPL], g2[DPL], be2[DPL], wo[DPL];
    __device__ __forceinline__ void init(const HeadArgs& h, int lane) {
        load_chan<DPL>(h.bnc + 6 * h.D, lane, k2a);
        load_chan<DPL>(h.bnc + 7 * h.D, lane, k2b);
        load_chan<DPL>(h.g2, lane, g2);
        load_chan<DPL>(h.be2, lane, be2);
        load_chan<DPL>(h.wo, lane, wo);
    }
};

// per-batch side data of the backward passes: lane l holds g_pred and the keep words of row b0+l
template <int DPL>
struct BwdSide {
    float gp;
    uint32_t bits[DPL];
    __device__ __forceinline__ void load(const HeadArgs& h, const float* __restrict__ gpred, int i, int b0, int nb,
                                         int lane) {
        const bool drop = h.training && h.p_drop > 0.f;
        gp = 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) bits[j] = 0xffffffffu;
        if (lane < nb) {
            const size_t r = (size_t)(b0 + lane) * h.N + i;
            gp = __ldg(gpred + r);
            if (drop) {
#pragma unroll
                for (int j = 0; j < DPL; ++j) bits[j] = h.bits[r * DPL + j];
            }
        }
    }
    // same, from the side area the batch staging filled (buffered passes)
    __device__ __forceinline__ void from_stage(const HeadArgs& h, const float* sx, int nb, int lane) {
        const bool drop = h.training && h.p_drop > 0.f;
        const float* side = sx + XhStage<DPL>::BUF_FLOATS;
        gp = lane < nb ? side[lane] : 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j)
            bits[j] = (drop && lane < nb) ? __float_as_uint(side[XH_ROWS + lane * DPL + j]) : 0xffffffffu;
    }
    // keep factor (0 or scale) of this lane's channels for row rr of the batch, and that row's g_pred
    __device__ __forceinline__ float row(const HeadArgs& h, int rr, int lane, float (&kf)[DPL]) const {
        const bool drop = h.training && h.p_drop > 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            const uint32_t word = __shfl_sync(0xffffffffu, bits[j], rr);
            kf[j] = drop ? (((word >> lane) & 1u) ? h.scale : 0.f) : 1.f;
        }
        return __shfl_sync(0xffffffffu, gp, rr);
    }
};

// pass 1: g_wo, g_gamma2, g_beta2, g_bo   -> part record [3*D + 32] doubles
template <int DPL, int WP, bool BUF_>
__global__ void __launch_bounds__(256)
k_bwd1(HeadArgs h, BwdArgs g, double* __restrict__ part) {
    constexpr bool NEED_A_ = !BUF_;
    RowEval<DPL, WP> re;
    BwdRow<DPL, WP> br;
    re.init(h, threadIdx.x & 31, NEED_A_);
    br.init(h, threadIdx.x & 31);
    double acc[3 * DPL];
#pragma unroll
    for (int q = 0; q < 3 * DPL; ++q) acc[q] = 0.0;
    double gbo = 0.0;
    GDN_TASK_LOOP_BEGIN(h)
        float v[DPL], t0[DPL], t1[DPL], t2[DPL];
        load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
#pragma unroll
        for (int j = 0; j < DPL; ++j) t0[j] = t1[j] = t2[j] = 0.f;
        float tb = 0.f;
        GDN_BATCH_LOOP_BEGIN2(h, g.gpred, ((h.training && h.p_drop > 0.f) ? (const uint32_t*)h.bits : (const uint32_t*)nullptr))
            BwdSide<DPL> side;
            if (BUF_) side.from_stage(h, sx_, nb, lane);
            else side.load(h, g.gpred, i, b0, nb, lane);
#pragma unroll 2
            for (int rr = 0; rr < nb; ++rr) {
                float a[WP], xh1[DPL], y1[DPL], kf[DPL];
                row_inputs<DPL, WP, NEED_A_, BUF_>(re, sa_, sx_, rr, lane, a, xh1, y1);
                const float gp = side.row(h, rr, lane, kf);
                tb += gp;
                // packed chain (pairs of channels per instruction)
                float r1[DPL], p[DPL], xh2[DPL], y2[DPL], q[DPL], r2[DPL], wq[DPL], gy2[DPL];
                vrelu<DPL>(r1, y1);
                vmul<DPL>(p, r1, v);
                vfma<DPL>(xh2, p, br.k2a, br.k2b);
                vfma<DPL>(y2, br.g2, xh2, br.be2);
                vscale<DPL>(q, kf, gp);                                          // g_pred * keep factor
                vrelu<DPL>(r2, y2);
                vfma<DPL>(t0, r2, q, t0);                                        // g_wo
                vmul<DPL>(wq, br.wo, q);
                vgate<DPL>(gy2, y2, wq);
                vfma<DPL>(t1, gy2, xh2, t1);                                     // g_gamma2
                vadd<DPL>(t2, t2, gy2);                                          // g_beta2
            }
        GDN_BATCH_LOOP_END
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            acc[j] += (double)t0[j]; acc[DPL + j] += (double)t1[j]; acc[2 * DPL + j] += (double)t2[j];
        }
        gbo += (double)tb;
    GDN_TASK_LOOP_END
    __syncthreads();
    double* dsm = reinterpret_cast<double*>(dyn_smem);
    double* prec = part + (size_t)blockIdx.x * (3 * h.D + 32);
    cta_reduce_channels<3 * DPL, DPL>(acc, h.D, prec, dsm);
    // g_bo: every lane of a warp holds the same value; one slot per warp
    if ((threadIdx.x & 31) == 0) dsm[threadIdx.x >> 5] = gbo;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += dsm[w];
        prec[3 * h.D] = s;
    }
}

// pass 2: g_V (sum over windows), g_gamma1, g_beta1     -> part record [2*D] doubles
template <int DPL, int WP, bool BUF_>
__global__ void __launch_bounds__(256)
k_bwd2(HeadArgs h, BwdArgs g, double* __restrict__ part) {
    constexpr bool NEED_A_ = !BUF_;
    RowEval<DPL, WP> re;
    BwdRow<DPL, WP> br;
    re.init(h, threadIdx.x & 31, NEED_A_);
    br.init(h, threadIdx.x & 31);
    float cB2[DPL], cG2[DPL], s2c[DPL];
    {
        const int ln = threadIdx.x & 31;
        load_chan<DPL>(g.c2, ln, cB2);
        load_chan<DPL>(g.c2 + h.D, ln, cG2);
#pragma unroll
        for (int j = 0; j < DPL; ++j) s2c[j] = br.g2[j] * br.k2a[j];
    }
    float ncB2[DPL], ncG2[DPL], wok[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) { ncB2[j] = -cB2[j]; ncG2[j] = -cG2[j]; wok[j] = br.wo[j]; }
    double acc[2 * DPL];
#pragma unroll
    for (int q = 0; q < 2 * DPL; ++q) acc[q] = 0.0;
    GDN_TASK_LOOP_BEGIN(h)
        float v[DPL], gv[DPL], t1[DPL], t2[DPL];
        load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
#pragma unroll
        for (int j = 0; j < DPL; ++j) gv[j] = t1[j] = t2[j] = 0.f;
        GDN_BATCH_LOOP_BEGIN2(h, g.gpred, ((h.training && h.p_drop > 0.f) ? (const uint32_t*)h.bits : (const uint32_t*)nullptr))
            BwdSide<DPL> side;
            if (BUF_) side.from_stage(h, sx_, nb, lane);
            else side.load(h, g.gpred, i, b0, nb, lane);
#pragma unroll 2
            for (int rr = 0; rr < nb; ++rr) {
                float a[WP], xh1[DPL], y1[DPL], kf[DPL];
                row_inputs<DPL, WP, NEED_A_, BUF_>(re, sa_, sx_, rr, lane, a, xh1, y1);
                const float gp = side.row(h, rr, lane, kf);
                // packed chain (pairs of channels per instruction)
                float r1[DPL], p[DPL], xh2[DPL], y2[DPL], t[DPL], gy2[DPL], gpp[DPL], gy1[DPL];
                vrelu<DPL>(r1, y1);
                vmul<DPL>(p, r1, v);
                vfma<DPL>(xh2, p, br.k2a, br.k2b);
                vfma<DPL>(y2, br.g2, xh2, br.be2);
                vmul<DPL>(t, wok, kf);                                           // w_o * keep factor
                vscale<DPL>(t, t, gp);
                vgate<DPL>(gy2, y2, t);
                vadd<DPL>(t, gy2, ncB2);
                vfma<DPL>(t, xh2, ncG2, t);                                      // gy2 - cB2 - xh2 cG2
                vmul<DPL>(gpp, s2c, t);                                          // d loss / d p
                vfma<DPL>(gv, gpp, r1, gv);
                vmul<DPL>(t, gpp, v);
                vgate<DPL>(gy1, y1, t);
                vfma<DPL>(t1, gy1, xh1, t1);                                     // g_gamma1
                vadd<DPL>(t2, t2, gy1);                                          // g_beta1
            }
        GDN_BATCH_LOOP_END
        store_chan<DPL>(g.gV + ((size_t)sp * h.N + i) * h.D, lane, gv);
#pragma unroll
        for (int j = 0; j < DPL; ++j) { acc[j] += (double)t1[j]; acc[DPL + j] += (double)t2[j]; }
    GDN_TASK_LOOP_END
    __syncthreads();
    cta_reduce_channels<2 * DPL, DPL>(acc, h.D, part + (size_t)blockIdx.x * 2 * h.D, reinterpret_cast<double*>(dyn_smem));
}

// pass 3: g_A[r,:] = g_z.Wl, partial g_Wl += g_z (x) A, partial g_bias += g_z
// part record: [D*W + D] doubles
template <int DPL, int WP, bool BUF_>
__global__ void __launch_bounds__(256)
k_bwd3(HeadArgs h, BwdArgs g, double* __restrict__ part) {
    constexpr bool NEED_A_ = true;
    RowEval<DPL, WP> re;
    BwdRow<DPL, WP> br;
    re.init(h, threadIdx.x & 31);
    br.init(h, threadIdx.x & 31);
    const int wid = threadIdx.x >> 5;
    float cB2[DPL], cG2[DPL], s2c[DPL], cB1[DPL], cG1[DPL], s1c[DPL];
    {
        const int ln = threadIdx.x & 31;
        load_chan<DPL>(g.c2, ln, cB2);
        load_chan<DPL>(g.c2 + h.D, ln, cG2);
        load_chan<DPL>(g.c1, ln, cB1);
        load_chan<DPL>(g.c1 + h.D, ln, cG1);
#pragma unroll
        for (int j = 0; j < DPL; ++j) { s2c[j] = br.g2[j] * br.k2a[j]; s1c[j] = re.g1[j] * re.k1a[j]; }
    }
    float gwl[DPL][WP], gb[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
        gb[j] = 0.f;
#pragma unroll
        for (int w = 0; w < WP; ++w) gwl[j][w] = 0.f;
    }
    GDN_TASK_LOOP_BEGIN(h)
        float v[DPL];
        load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
        GDN_BATCH_LOOP_BEGIN2(h, g.gpred, ((h.training && h.p_drop > 0.f) ? (const uint32_t*)h.bits : (const uint32_t*)nullptr))
            BwdSide<DPL> side;
            if (BUF_) side.from_stage(h, sx_, nb, lane);
            else side.load(h, g.gpred, i, b0, nb, lane);
            for (int rr = 0; rr < nb; ++rr) {
                const size_t r = (size_t)(b0 + rr) * h.N + i;
                float a[WP], xh1[DPL], y1[DPL], kf[DPL], pw[WP];
                row_inputs<DPL, WP, NEED_A_, BUF_>(re, sa_, sx_, rr, lane, a, xh1, y1);
                const float gp = side.row(h, rr, lane, kf);
#pragma unroll
                for (int w = 0; w < WP; ++w) pw[w] = 0.f;
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const float p = fmaxf(y1[j], 0.f) * v[j];
                    const float xh2 = fmaf(p, br.k2a[j], br.k2b[j]);
                    const float y2 = fmaf(br.g2[j], xh2, br.be2[j]);
                    const float gy2 = y2 > 0.f ? gp * br.wo[j] * kf[j] : 0.f;
                    const float gpp = s2c[j] * (gy2 - cB2[j] - xh2 * cG2[j]);
                    const float gy1 = y1[j] > 0.f ? gpp * v[j] : 0.f;
                    const float gz = s1c[j] * (gy1 - cB1[j] - xh1[j] * cG1[j]);   // d loss / d z
                    gb[j] += gz;
#pragma unroll
                    for (int w = 0; w < WP; ++w) {
                        pw[w] = fmaf(gz, re.wl[j][w], pw[w]);
                        gwl[j][w] = fmaf(gz, a[w], gwl[j][w]);
                    }
                }
                int widx;
                const float tot = reduce_scatter<WP>(pw, lane, &widx);
                if ((lane & ((32 / WP) - 1)) == 0 && widx < h.W) g.gA[r * h.W + widx] = tot;
            }
        GDN_BATCH_LOOP_END
    GDN_TASK_LOOP_END
    __syncthreads();
    float (*red)[33][32] = reinterpret_cast<float (*)[33][32]>(dyn_smem);
    double* prec = part + (size_t)blockIdx.x * ((size_t)h.D * h.W + h.D);
    const int ln = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
#pragma unroll
        for (int w = 0; w < WP; ++w) red[wid][w][ln] = gwl[j][w];
        red[wid][32][ln] = gb[j];
        __syncthreads();
        for (int e = threadIdx.x; e < 33 * 32; e += blockDim.x) {
            const int w = e >> 5, l = e & 31;
            if (w < h.W || w == 32) {
                double s = 0.0;
                for (int q = 0; q < 8; ++q) s += (double)red[q][w][l];
                const int d = l * DPL + j;
                if (w == 32) prec[(size_t)h.D * h.W + d] = s;
                else prec[(size_t)d * h.W + w] = s;
            }
        }
        __syncthreads();
    }
}

// pass 3, tensor-core version (saved xh1, D in {64, 128}, W <= 16).  Per batch of 16 windows of one sensor
// the warp walks the rows lane <-> channels (the BatchNorm/ReLU chain down to g_z, g_bias partials),
// overwriting the staged xh1 tile with g_z IN PLACE, then runs the two contractions of LinBwdMma on that tile.
// The next batch's xh1 and A rows are already in flight (cp.async, two buffers per warp) while the current
// one is computed, and its g_pred / keep words wait in registers: no exposed global latency in the loop.
// dynamic smem: Wl {hi,lo} | accumulator image | per warp 2 x (xh1/g_z tile [16][D+4], A tile [16][WP+4])
template <int DPL, int WP>
__global__ void __launch_bounds__(256)
k_bwd3_mma(HeadArgs h, BwdArgs g, double* __restrict__ part) {
    using M = LinBwdMma<DPL, WP>;
    constexpr int GS = M::GS, AS = M::AS, TR = M::TR;
    constexpr int BUF_FLOATS = TR * GS + TR * AS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RowEval<DPL, WP> re;
    BwdRow<DPL, WP> br;
    re.init(h, lane, false);
    br.init(h, lane);
    float cB2[DPL], cG2[DPL], s2c[DPL], cB1[DPL], cG1[DPL], s1c[DPL];
    load_chan<DPL>(g.c2, lane, cB2);
    load_chan<DPL>(g.c2 + h.D, lane, cG2);
    load_chan<DPL>(g.c1, lane, cB1);
    load_chan<DPL>(g.c1 + h.D, lane, cG1);
#pragma unroll
    for (int j = 0; j < DPL; ++j) { s2c[j] = br.g2[j] * br.k2a[j]; s1c[j] = re.g1[j] * re.k1a[j]; }
    float ncB2[DPL], ncG2[DPL], ncB1[DPL], ncG1[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) { ncB2[j] = -cB2[j]; ncG2[j] = -cG2[j]; ncB1[j] = -cB1[j]; ncG1[j] = -cG1[j]; }
    uint2* sW = reinterpret_cast<uint2*>(dyn_smem);
    float* sAcc = reinterpret_cast<float*>(dyn_smem + M::W_BYTES);
    float* sbuf = reinterpret_cast<float*>(dyn_smem + M::W_BYTES + M::ACC_BYTES) + (size_t)wid * 2 * BUF_FLOATS;
    M::fill_w(sW, h.Wl, h.W);
    M::zero_acc(sAcc);
    __syncthreads();
    int since_flush = 0;
    float acc2[M::NT2][4], gb[DPL];
#pragma unroll
    for (int nt = 0; nt < M::NT2; ++nt) acc2[nt][0] = acc2[nt][1] = acc2[nt][2] = acc2[nt][3] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) gb[j] = 0.f;

    const long long warp_ = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps_ = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long tasks_ = (long long)h.N * h.S;
    // work items of this warp: (task, b0) in task-major order; task -> (sensor i, window range [b_lo, b_hi))
    auto task_range = [&](long long task, int& i, int& b_lo, int& b_hi) {
        i = (int)(task / h.S);
        const int sp = (int)(task % h.S);
        b_lo = sp * h.rps;
        b_hi = min(h.B, b_lo + h.rps);
    };
    // stage the xh1 rows and the A rows of batch (i, b0, nb) into buffer `buf` (asynchronous)
    auto issue = [&](int i, int b0, int nb, float* buf) {
        float* sx = buf;
        float* sa = buf + TR * GS;
        const float* xsrc = h.xh1 + ((size_t)b0 * h.N + i) * h.D + lane * DPL;
        const uint32_t xdst = (uint32_t)__cvta_generic_to_shared(sx + lane * DPL);
        for (int rr = 0; rr < nb; ++rr) {
            const float* src = xsrc + (size_t)rr * h.N * h.D;
            const uint32_t dst = xdst + rr * GS * 4;
            if (DPL == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        }
        if (lane < nb) {
            const float* arow = h.A + ((size_t)(b0 + lane) * h.N + i) * h.W;
            const uint32_t adst = (uint32_t)__cvta_generic_to_shared(sa + lane * AS);
            if ((h.W & 3) == 0) {
#pragma unroll
                for (int w = 0; w < WP; w += 4) {
                    const int nbytes = w < h.W ? 16 : 0;                           // zero fill beyond W
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(adst + w * 4), "l"(arow + (w < h.W ? w : 0)),
                                 "r"(nbytes) : "memory");
                }
            } else {
#pragma unroll
                for (int w = 0; w < WP; ++w) {
                    const int nbytes = w < h.W ? 4 : 0;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(adst + w * 4), "l"(arow + (w < h.W ? w : 0)),
                                 "r"(nbytes) : "memory");
                }
            }
        }
    };

    long long task = warp_;
    int i = 0, b_lo = 0, b_hi = 0, b0 = 0;
    bool have = task < tasks_;
    BwdSide<DPL> side;
    if (have) {
        task_range(task, i, b_lo, b_hi);
        b0 = b_lo;
        issue(i, b0, min(TR, b_hi - b0), sbuf);
        side.load(h, g.gpred, i, b0, min(TR, b_hi - b0), lane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    float v[DPL];
    bool new_task = true;
    for (int k = 0; have; ++k) {
        const int nb = min(TR, b_hi - b0);
        // the batch after this one
        long long ntask = task;
        int ni = i, nb_lo = b_lo, nb_hi = b_hi, nb0 = b0 + TR;
        bool nhave = true;
        if (nb0 >= b_hi) {
            ntask = task + nwarps_;
            nhave = ntask < tasks_;
            if (nhave) { task_range(ntask, ni, nb_lo, nb_hi); nb0 = nb_lo; }
        }
        BwdSide<DPL> nside = side;
        __syncwarp();                                         // everyone is done with the buffer about to be refilled
        if (nhave) {
            issue(ni, nb0, min(TR, nb_hi - nb0), sbuf + ((k + 1) & 1) * BUF_FLOATS);
            nside.load(h, g.gpred, ni, nb0, min(TR, nb_hi - nb0), lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory"); // this batch's copies have landed
        __syncwarp();
        float* sG = sbuf + (k & 1) * BUF_FLOATS;
        float* sa = sG + TR * GS;
        if (new_task) load_chan_vec<DPL>(h.V + (size_t)i * h.D, lane, v);
#pragma unroll 2
        for (int rr = 0; rr < nb; ++rr) {
            float xh1[DPL], y1[DPL], kf[DPL], gz[DPL];
            load_chan_vec<DPL>(sG + rr * GS, lane, xh1);
            re.from_saved(xh1, y1);
            const float gp = side.row(h, rr, lane, kf);
            // packed chain (pairs of channels per instruction)
            float r1[DPL], p[DPL], xh2[DPL], y2[DPL], t[DPL], gy2[DPL], gpp[DPL], gy1[DPL];
            vrelu<DPL>(r1, y1);
            vmul<DPL>(p, r1, v);
            vfma<DPL>(xh2, p, br.k2a, br.k2b);
            vfma<DPL>(y2, br.g2, xh2, br.be2);
            vmul<DPL>(t, br.wo, kf);
            vscale<DPL>(t, t, gp);
            vgate<DPL>(gy2, y2, t);
            vadd<DPL>(t, gy2, ncB2);
            vfma<DPL>(t, xh2, ncG2, t);                                          // gy2 - cB2 - xh2 cG2
            vmul<DPL>(gpp, s2c, t);
            vmul<DPL>(t, gpp, v);
            vgate<DPL>(gy1, y1, t);
            vadd<DPL>(t, gy1, ncB1);
            vfma<DPL>(t, xh1, ncG1, t);                                          // gy1 - cB1 - xh1 cG1
            vmul<DPL>(gz, s1c, t);                                               // d loss / d z
            vadd<DPL>(gb, gb, gz);
            store_chan<DPL>(sG + rr * GS, lane, gz);                            // in place
        }
        if (nb < TR) {                                                        // tail: zero rows for the MMAs
            float z[DPL];
#pragma unroll
            for (int j = 0; j < DPL; ++j) z[j] = 0.f;
            for (int rr = nb; rr < TR; ++rr) store_chan<DPL>(sG + rr * GS, lane, z);
            for (int e = lane; e < (TR - nb) * AS; e += 32) sa[nb * AS + e] = 0.f;
        }
        __syncwarp();
        float c1[M::NT1][4];
        M::ga_tile(sG, sW, lane, c1);
        {
            const int gg = lane >> 2, t = lane & 3;
#pragma unroll
            for (int nt = 0; nt < M::NT1; ++nt) {
                const int w0 = 8 * nt + 2 * t;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int rr = gg + 8 * hh;
                    if (rr < nb && w0 < h.W) {
                        float* o = g.gA + ((size_t)(b0 + rr) * h.N + i) * h.W + w0;
                        if ((h.W & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(c1[nt][2 * hh], c1[nt][2 * hh + 1]);
                        else {
                            o[0] = c1[nt][2 * hh];
                            if (w0 + 1 < h.W) o[1] = c1[nt][2 * hh + 1];
                        }
                    }
                }
            }
        }
        M::gwl_acc(sG, sa, lane, acc2);
        if (++since_flush == 4) {                                             // 24 chained MMAs per accumulator
            M::flush(acc2, sAcc, lane);
            since_flush = 0;
        }
        new_task = ntask != task;
        task = ntask; i = ni; b_lo = nb_lo; b_hi = nb_hi; b0 = nb0; have = nhave;
        side = nside;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (since_flush) M::flush(acc2, sAcc, lane);
    __syncthreads();
    M::finish(gb, sAcc, h.W, part + (size_t)blockIdx.x * ((size_t)h.D * h.W + h.D));
}

// ---------------------------------------------------------------------------------------
// per-CTA partial records -> one record of doubles (fixed summation order: deterministic)
// block (32, 32): x <-> entry, y strides the records; grid = ceil(rec / 32)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
k_reduce_part(const T* __restrict__ part, int nrec, int rec, double* __restrict__ sums) {
    __shared__ double sh[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (e < rec) {
        const T* col = part + e;
        int q = threadIdx.y;
        for (; q + 224 < nrec; q += 256) {                 // eight independent loads in flight, summed in order
            T t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = col[(size_t)(q + 32 * u) * rec];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)t[u];
        }
        for (; q < nrec; q += 32) s += (double)col[(size_t)q * rec];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < rec) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += sh[k][threadIdx.x];
        sums[e] = t;
    }
}

template <typename T>
static int reduce_part(const T* part, int nrec, int rec, double* sums, cudaStream_t st) {
    k_reduce_part<T><<<(rec + 31) / 32, dim3(32, 32), 0, st>>>(part, nrec, rec, sums);
    GDN_CHECK_LAUNCH("k_reduce_part");
    return 0;
}

// ---------------------------------------------------------------------------------------
// finalize kernels (one CTA, 256 threads; they read ONE pre-reduced record)
// ---------------------------------------------------------------------------------------
// BN1 statistics from the moments of A; writes ctx.bn[0..3], updates running stats.
__global__ void k_fin_bn1(const double* __restrict__ part, int nrec, long long n, int W, int D,
                          const float* __restrict__ Wl, const float* __restrict__ bias,
                          float* __restrict__ bnc, float* __restrict__ rmean, float* __restrict__ rvar,
                          long long* __restrict__ nbt) {
    extern __shared__ double sm[];          // [W*W + W]
    const int rec = W * W + W;
    for (int e = threadIdx.x; e < rec; e += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < nrec; ++q) s += part[(size_t)q * rec + e];
        sm[e] = s / (double)n;
    }
    __syncthreads();
    const double* m2 = sm;                  // E[a_w a_w']
    const double* m1 = sm + W * W;          // E[a_w]
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double mean = bias ? (double)bias[d] : 0.0, var = 0.0;
        for (int w = 0; w < W; ++w) {
            const double wl = (double)Wl[(size_t)d * W + w];
            mean += wl * m1[w];
            double row = 0.0;
            for (int w2 = 0; w2 < W; ++w2)
                row += (m2[w * W + w2] - m1[w] * m1[w2]) * (double)Wl[(size_t)d * W + w2];
            var += wl * row;
        }
        if (var < 0.0) var = 0.0;
        const double istd = 1.0 / sqrt(var + (double)GDN_BN_EPS);
        bnc[d] = (float)mean;
        bnc[D + d] = (float)istd;
        bnc[2 * D + d] = (float)istd;
        bnc[3 * D + d] = (float)(((bias ? (double)bias[d] : 0.0) - mean) * istd);
        if (rmean != nullptr) {
            const double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
            rmean[d] = (1.f - GDN_BN_MOMENTUM) * rmean[d] + GDN_BN_MOMENTUM * (float)mean;
            rvar[d] = (1.f - GDN_BN_MOMENTUM) * rvar[d] + GDN_BN_MOMENTUM * (float)unb;
        }
    }
    if (threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
}

// BN2 statistics from sum p, sum p^2; writes ctx.bn[4..7], updates running stats.
__global__ void k_fin_bn2(const double* __restrict__ part, int nrec, long long n, int D,
                          float* __restrict__ bnc, float* __restrict__ rmean, float* __restrict__ rvar,
                          long long* __restrict__ nbt) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int q = 0; q < nrec; ++q) { s1 += part[(size_t)q * 2 * D + d]; s2 += part[(size_t)q * 2 * D + D + d]; }
        const double mean = s1 / (double)n;
        double var = s2 / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double istd = 1.0 / sqrt(var + (double)GDN_BN_EPS);
        bnc[4 * D + d] = (float)mean;
        bnc[5 * D + d] = (float)istd;
        bnc[6 * D + d] = (float)istd;
        bnc[7 * D + d] = (float)(-mean * istd);
        if (rmean != nullptr) {
            const double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
            rmean[d] = (1.f - GDN_BN_MOMENTUM) * rmean[d] + GDN_BN_MOMENTUM * (float)mean;
            rvar[d] = (1.f - GDN_BN_MOMENTUM) * rvar[d] + GDN_BN_MOMENTUM * (float)unb;
        }
    }
    if (threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
}

// eval mode: constants from the running statistics (both BNs)
__global__ void k_fin_bn_eval(int D, const float* __restrict__ bias,
                              const float* __restrict__ rm1, const float* __restrict__ rv1,
                              const float* __restrict__ rm2, const float* __restrict__ rv2,
                              float* __restrict__ bnc) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float i1 = 1.f / sqrtf(rv1[d] + GDN_BN_EPS);
        const float i2 = 1.f / sqrtf(rv2[d] + GDN_BN_EPS);
        bnc[d] = rm1[d]; bnc[D + d] = i1; bnc[2 * D + d] = i1;
        bnc[3 * D + d] = ((bias ? bias[d] : 0.f) - rm1[d]) * i1;
        bnc[4 * D + d] = rm2[d]; bnc[5 * D + d] = i2; bnc[6 * D + d] = i2;
        bnc[7 * D + d] = -rm2[d] * i2;
    }
}

// pass-1 sums -> g_wo, g_gamma2, g_beta2, g_bo and the pass-2 coefficients c2 = (g_beta2/n, g_gamma2/n)
// `glob` = the same record summed over the data-parallel ranks (SyncBN), n = rows of ALL ranks; NULL: this rank only
__global__ void k_fin_bwd1(const double* __restrict__ part, int nrec, const double* __restrict__ glob, long long n, int D,
                           float* __restrict__ g_wo, float* __restrict__ g_g2, float* __restrict__ g_b2,
                           float* __restrict__ g_bo, float* __restrict__ c2) {
    const int rec = 3 * D + 32;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int q = 0; q < nrec; ++q) {
            a += part[(size_t)q * rec + d]; b += part[(size_t)q * rec + D + d]; c += part[(size_t)q * rec + 2 * D + d];
        }
        g_wo[d] = (float)a; g_g2[d] = (float)b; g_b2[d] = (float)c;
        if (glob != nullptr) { b = glob[D + d]; c = glob[2 * D + d]; }
        c2[d] = (float)(c / (double)n);
        c2[D + d] = (float)(b / (double)n);
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < nrec; ++q) s += part[(size_t)q * rec + 3 * D];
        g_bo[0] = (float)s;
    }
}

__global__ void k_fin_bwd2(const double* __restrict__ part, int nrec, const double* __restrict__ glob, long long n, int D,
                           float* __restrict__ g_g1, float* __restrict__ g_b1, float* __restrict__ c1) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double b = 0.0, c = 0.0;
        for (int q = 0; q < nrec; ++q) { b += part[(size_t)q * 2 * D + d]; c += part[(size_t)q * 2 * D + D + d]; }
        g_g1[d] = (float)b; g_b1[d] = (float)c;
        if (glob != nullptr) { b = glob[d]; c = glob[D + d]; }
        c1[d] = (float)(c / (double)n);
        c1[D + d] = (float)(b / (double)n);
    }
}

// sum the S partial embedding gradients
__global__ void k_reduce_gV(const float* __restrict__ gVp, int S, long long ND, float* __restrict__ gV) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ND; e += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < S; ++q) s += gVp[(size_t)q * ND + e];
        gV[e] = s;
    }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
#define GDN_DISPATCH_DW(DPLV, WPV, CALL)                                             \
    do {                                                                              \
        if (DPLV == 1) { GDN_DISPATCH_W(1, WPV, CALL); }                              \
        else if (DPLV == 2) { GDN_DISPATCH_W(2, WPV, CALL); }                         \
        else if (DPLV == 4) { GDN_DISPATCH_W(4, WPV, CALL); }                         \
        else { GDN_DISPATCH_W(8, WPV, CALL); }                                        \
    } while (0)
#define GDN_DISPATCH_W(DPLC, WPV, CALL)                                               \
    do {                                                                              \
        if (WPV == 8) { CALL(DPLC, 8); }                                              \
        else if (WPV == 16) { CALL(DPLC, 16); }                                       \
        else { CALL(DPLC, 32); }                                                      \
    } while (0)

// dynamic shared memory of a D-wide kernel: the row stage, aliased after the main loop by the
// cross-warp reduction scratch (doubles for channel sums, red[8][33][32] floats for g_Wl)
template <int WP>
static size_t dw_smem(int nv_double, bool red33) {
    size_t b = RowStage<WP>::bytes(8);
    const size_t r1 = (size_t)8 * nv_double * 32 * sizeof(double);
    const size_t r2 = red33 ? (size_t)8 * 33 * 32 * sizeof(float) : 0;
    if (r1 > b) b = r1;
    if (r2 > b) b = r2;
    return b;
}
// per-warp staging of the fused passes: A tile (if needed) + saved-xh1 tile (if buffered)
template <int DPL, int WP>
static size_t dw_smem2(bool need_a, bool buf, int nv_double, bool red33) {
    size_t b = (size_t)8 * ((need_a ? RowStage<WP>::WARP_FLOATS : 0) + (buf ? XhStage<DPL>::WARP_FLOATS : 0)) * sizeof(float);
    const size_t r1 = (size_t)8 * nv_double * 32 * sizeof(double);
    const size_t r2 = red33 ? (size_t)8 * 33 * 32 * sizeof(float) : 0;
    if (r1 > b) b = r1;
    if (r2 > b) b = r2;
    return b;
}
#define GDN_LAUNCH_DYN(KERNEL, GRID, SMEM, ST, ...)                                                     \
    do {                                                                                               \
        const size_t sm__ = (SMEM);                                                                    \
        {                                                                                              \
            cudaError_t e__ = ensure_dyn_smem(KERNEL, sm__);                                           \
            if (e__ != cudaSuccess) return cuda_fail(e__, "smem attribute");                           \
        }                                                                                              \
        KERNEL<<<GRID, 256, sm__, ST>>>(__VA_ARGS__);                                                  \
    } while (0)

// GDN_MMA_FLUSH (diagnostics): tiles between two flushes of the tensor-core accumulators (default 4)
static int mma_flush() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GDN_MMA_FLUSH"); v = e ? atoi(e) : 4; if (v < 1) v = 1; }
    return v;
}

// GDN_NO_MMA (bit mask, diagnostics): 1 = k_lin_bwd, 2 = k_bwd3 fall back to the FMA kernels
static int no_mma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GDN_NO_MMA"); v = e ? atoi(e) : 0; }
    return v;
}

static int dw_grid(long long tasks) {
    long long g = (tasks + 7) / 8;
    const int cap = 4 * num_sms();      // ws partial-record regions are sized for this many CTAs
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_lin_fwd(const Shape& s, const float* A, const gdn_layer_params* p, float* out, cudaStream_t st) {
    const int grid = dw_grid(s.n);
#define CALL(DPLC, WPC) \
    GDN_LAUNCH_DYN((k_lin_fwd<DPLC, WPC>), grid, dw_smem<WPC>(0, false), st, A, p->lin_weight, p->bias, s.n, s.W, s.D, out)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_lin_fwd");
    return 0;
}

template <int DPL, int WP>
static int launch_lin_bwd_mma(const Shape& s, const float* gout, const float* A, const gdn_layer_params* p, float* gA,
                              double* part, int* nrec, cudaStream_t st) {
    using M = LinBwdMma<DPL, WP>;
    constexpr int TR = M::TR, nw = 8;
    const size_t smem = M::W_BYTES + M::ACC_BYTES + (size_t)nw * (M::G_BYTES + (size_t)TR * M::AS * 4);
    long long g = (s.n + (long long)TR * nw - 1) / ((long long)TR * nw);
    if (g > 2 * num_sms()) g = 2 * num_sms();
    const int grid = (int)(g < 1 ? 1 : g);
    cudaError_t e = ensure_dyn_smem(k_lin_bwd_mma<DPL, WP>, smem);
    if (e != cudaSuccess) return cuda_fail(e, "smem attribute k_lin_bwd_mma");
    k_lin_bwd_mma<DPL, WP><<<grid, nw * 32, smem, st>>>(gout, A, p->lin_weight, s.n, s.W, s.D, gA, part, mma_flush());
    GDN_CHECK_LAUNCH("k_lin_bwd");
    *nrec = grid;
    return 0;
}

int launch_lin_bwd(const Shape& s, const float* gout, const float* A, const gdn_layer_params* p, float* gA,
                   double* part, int* nrec, cudaStream_t st) {
    // tensor-core path for the shapes it is built for
    if (no_mma() & 1) {}
    else if (s.DPL == 4 && s.WP == 16) return launch_lin_bwd_mma<4, 16>(s, gout, A, p, gA, part, nrec, st);
    if (s.DPL == 4 && s.WP == 8) return launch_lin_bwd_mma<4, 8>(s, gout, A, p, gA, part, nrec, st);
    if (s.DPL == 2 && s.WP == 16) return launch_lin_bwd_mma<2, 16>(s, gout, A, p, gA, part, nrec, st);
    if (s.DPL == 2 && s.WP == 8) return launch_lin_bwd_mma<2, 8>(s, gout, A, p, gA, part, nrec, st);
    // warps per CTA: as many as fit the shared-memory tiles
    int nw = 8;
    auto bytes = [&](int w) {
        size_t b = (size_t)s.D * s.WP * 4 + (size_t)w * (32 * (s.D + 4) + 32 * (s.WP + 4)) * 4;
        const size_t red = (size_t)w * 33 * 32 * 4;
        return b > red ? b : red;
    };
    while (nw > 1 && bytes(nw) > 200 * 1024) nw >>= 1;
    const size_t smem = bytes(nw);
    long long g = (s.n + 32LL * nw - 1) / (32LL * nw);
    if (g > 2 * num_sms()) g = 2 * num_sms();
    const int grid = (int)(g < 1 ? 1 : g);
#define CALL(DPLC, WPC)                                                                                          \
    do {                                                                                                         \
        cudaError_t e__ = ensure_dyn_smem(k_lin_bwd<DPLC, WPC>, smem);                                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, "smem attribute k_lin_bwd");                               \
        k_lin_bwd<DPLC, WPC><<<grid, nw * 32, smem, st>>>(gout, A, p->lin_weight, s.n, s.W, s.D, gA, part);      \
    } while (0)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_lin_bwd");
    *nrec = grid;
    return 0;
}

int launch_moments(const Shape& s, const float* A, double* part, int* nrec, cudaStream_t st) {
    const int nw = s.WP == 32 ? 2 : 8;
    long long g = (s.n + 32LL * nw - 1) / (32LL * nw);
    if (g > 4 * num_sms()) g = 4 * num_sms();
    const int grid = (int)(g < 1 ? 1 : g);
    if (s.WP == 8) k_moments<8><<<grid, 256, 0, st>>>(A, s.n, s.W, part);
    else if (s.WP == 16) k_moments<16><<<grid, 256, 0, st>>>(A, s.n, s.W, part);
    else k_moments<32><<<grid, 64, 0, st>>>(A, s.n, s.W, part);
    GDN_CHECK_LAUNCH("k_moments");
    *nrec = grid;
    return 0;
}

// SyncBN: `count` reduced doubles at `local` -> their sum over the ranks at local + stride (the second half of the
// `sums` region); returns that pointer, or NULL when statistics are per rank
static bool sync_on(const gdn_sync* sy) { return sy != nullptr && sy->world > 1 && sy->allreduce_sum_f64 != nullptr; }
static int sync_sums(const gdn_sync* sy, double* local, int count, double** global_out, cudaStream_t st) {
    *global_out = nullptr;
    if (!sync_on(sy)) return 0;
    double* glob = local + sums_stride();
    cudaError_t e = cudaMemcpyAsync(glob, local, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "sync_sums copy");
    const int rc = sy->allreduce_sum_f64(glob, (long long)count, sy->user, (void*)st);
    if (rc != 0) { set_error("SyncBN all-reduce callback failed (rc=%d)", rc); return rc > 0 ? -2 : rc; }
    *global_out = glob;
    return 0;
}
static long long rows_total(const Shape& s, const gdn_sync* sy) { return sync_on(sy) ? s.n * (long long)sy->world : s.n; }

int launch_fin_bn1(const Shape& s, const double* part, int nrec, double* sums, const gdn_layer_params* p, float* bnc,
                   const gdn_bn* bn, const gdn_sync* sync, cudaStream_t st) {
    const size_t smem = ((size_t)s.W * s.W + s.W) * sizeof(double);
    if (int rc = reduce_part<double>(part, nrec, s.W * s.W + s.W, sums, st)) return rc;
    double* glob = nullptr;
    if (int rc = sync_sums(sync, sums, s.W * s.W + s.W, &glob, st)) return rc;
    k_fin_bn1<<<1, 256, smem, st>>>(glob ? glob : sums, 1, rows_total(s, sync), s.W, s.D, p->lin_weight, p->bias, bnc,
                                    bn->running_mean, bn->running_var, (long long*)bn->num_batches_tracked);
    GDN_CHECK_LAUNCH("k_fin_bn1");
    return 0;
}

int launch_fin_bn_eval(const Shape& s, const gdn_layer_params* p, const gdn_head_params* h, float* bnc,
                       cudaStream_t st) {
    k_fin_bn_eval<<<1, 256, 0, st>>>(s.D, p->bias, h->bn1.running_mean, h->bn1.running_var,
                                     h->bn2.running_mean, h->bn2.running_var, bnc);
    GDN_CHECK_LAUNCH("k_fin_bn_eval");
    return 0;
}


int launch_fwd_stats2(const Shape& s, const HeadArgs& h, double* part, double* sums, const gdn_bn* bn, float* bnc,
                      const gdn_sync* sync, cudaStream_t st) {
    const int grid = dw_grid((long long)s.N * s.S);
#define CALL(DPLC, WPC) GDN_LAUNCH_DYN((k_fwd_stats2<DPLC, WPC>), grid, (dw_smem2<DPLC, WPC>(true, false, 2 * DPLC, false)), st, h, part)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_fwd_stats2");
    if (int rc = reduce_part<double>(part, grid, 2 * s.D, sums, st)) return rc;
    double* glob = nullptr;
    if (int rc = sync_sums(sync, sums, 2 * s.D, &glob, st)) return rc;
    k_fin_bn2<<<1, 256, 0, st>>>(glob ? glob : sums, 1, rows_total(s, sync), s.D, bnc, bn->running_mean, bn->running_var,
                                 (long long*)bn->num_batches_tracked);
    GDN_CHECK_LAUNCH("k_fin_bn2");
    return 0;
}

int launch_fwd_out(const Shape& s, const HeadArgs& h, float* pred, cudaStream_t st) {
    const int grid = dw_grid((long long)s.N * s.S);
    const bool buf = h.xh1 != nullptr;
#define CALL(DPLC, WPC)                                                                                             \
    do {                                                                                                            \
        if (buf) GDN_LAUNCH_DYN((k_fwd_out<DPLC, WPC, true>), grid, (dw_smem2<DPLC, WPC>(false, true, 0, false)), st, h, pred); \
        else GDN_LAUNCH_DYN((k_fwd_out<DPLC, WPC, false>), grid, (dw_smem2<DPLC, WPC>(true, false, 0, false)), st, h, pred);    \
    } while (0)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_fwd_out");
    return 0;
}

int launch_bwd1(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, double* sums, gdn_head_grads* gh,
                float* c2, const gdn_sync* sync, cudaStream_t st) {
    const int grid = dw_grid((long long)s.N * s.S);
    const bool buf = h.xh1 != nullptr;
#define CALL(DPLC, WPC)                                                                                             \
    do {                                                                                                            \
        if (buf) GDN_LAUNCH_DYN((k_bwd1<DPLC, WPC, true>), grid, (dw_smem2<DPLC, WPC>(false, true, 3 * DPLC, false)), st, h, g, part); \
        else GDN_LAUNCH_DYN((k_bwd1<DPLC, WPC, false>), grid, (dw_smem2<DPLC, WPC>(true, false, 3 * DPLC, false)), st, h, g, part);    \
    } while (0)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_bwd1");
    if (int rc = reduce_part<double>(part, grid, 3 * s.D + 32, sums, st)) return rc;
    double* glob = nullptr;
    if (int rc = sync_sums(sync, sums, 3 * s.D + 32, &glob, st)) return rc;
    k_fin_bwd1<<<1, 256, 0, st>>>(sums, 1, glob, rows_total(s, sync), s.D, gh->out_w, gh->bn2_weight, gh->bn2_bias, gh->out_b, c2);
    GDN_CHECK_LAUNCH("k_fin_bwd1");
    return 0;
}

int launch_bwd2(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, double* sums, gdn_head_grads* gh,
                float* c1, float* gV_final, const gdn_sync* sync, cudaStream_t st) {
    const int grid = dw_grid((long long)s.N * s.S);
    const bool buf = h.xh1 != nullptr;
#define CALL(DPLC, WPC)                                                                                             \
    do {                                                                                                            \
        if (buf) GDN_LAUNCH_DYN((k_bwd2<DPLC, WPC, true>), grid, (dw_smem2<DPLC, WPC>(false, true, 2 * DPLC, false)), st, h, g, part); \
        else GDN_LAUNCH_DYN((k_bwd2<DPLC, WPC, false>), grid, (dw_smem2<DPLC, WPC>(true, false, 2 * DPLC, false)), st, h, g, part);    \
    } while (0)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_bwd2");
    if (int rc = reduce_part<double>(part, grid, 2 * s.D, sums, st)) return rc;
    double* glob = nullptr;
    if (int rc = sync_sums(sync, sums, 2 * s.D, &glob, st)) return rc;
    k_fin_bwd2<<<1, 256, 0, st>>>(sums, 1, glob, rows_total(s, sync), s.D, gh->bn1_weight, gh->bn1_bias, c1);
    GDN_CHECK_LAUNCH("k_fin_bwd2");
    if (s.S > 1) {
        const long long ND = (long long)s.N * s.D;
        int gr = (int)((ND + 255) / 256);
        if (gr > 8 * num_sms()) gr = 8 * num_sms();
        k_reduce_gV<<<gr, 256, 0, st>>>(g.gV, s.S, ND, gV_final);
        GDN_CHECK_LAUNCH("k_reduce_gV");
    }
    return 0;
}

template <int DPL, int WP>
static int launch_bwd3_mma(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, int* nrec, cudaStream_t st) {
    using M = LinBwdMma<DPL, WP>;
    const size_t smem = M::W_BYTES + M::ACC_BYTES + (size_t)8 * 2 * (M::G_BYTES + (size_t)M::TR * M::AS * sizeof(float));
    // one resident CTA per SM (shared-memory bound): a persistent grid, each CTA walks its share of the tasks
    int grid = dw_grid((long long)s.N * s.S);
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    const int cap = num_sms() * (per_sm < 1 ? 1 : per_sm);
    if (grid > cap) grid = cap;
    GDN_LAUNCH_DYN((k_bwd3_mma<DPL, WP>), grid, smem, st, h, g, part);
    GDN_CHECK_LAUNCH("k_bwd3");
    *nrec = grid;
    return 0;
}

int launch_bwd3(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, int* nrec, cudaStream_t st) {
    const int grid = dw_grid((long long)s.N * s.S);
    const bool buf = h.xh1 != nullptr;
    if (no_mma() & 2) {}
    else if (buf && s.DPL == 4 && s.WP == 16) return launch_bwd3_mma<4, 16>(s, h, g, part, nrec, st);
    if (buf && s.DPL == 4 && s.WP == 8) return launch_bwd3_mma<4, 8>(s, h, g, part, nrec, st);
    if (buf && s.DPL == 2 && s.WP == 16) return launch_bwd3_mma<2, 16>(s, h, g, part, nrec, st);
    if (buf && s.DPL == 2 && s.WP == 8) return launch_bwd3_mma<2, 8>(s, h, g, part, nrec, st);
#define CALL(DPLC, WPC)                                                                                             \
    do {                                                                                                            \
        if (buf) GDN_LAUNCH_DYN((k_bwd3<DPLC, WPC, true>), grid, (dw_smem2<DPLC, WPC>(true, true, 0, true)), st, h, g, part); \
        else GDN_LAUNCH_DYN((k_bwd3<DPLC, WPC, false>), grid, (dw_smem2<DPLC, WPC>(true, false, 0, true)), st, h, g, part);   \
    } while (0)
    GDN_DISPATCH_DW(s.DPL, s.WP, CALL);
#undef CALL
    GDN_CHECK_LAUNCH("k_bwd3");
    *nrec = grid;
    return 0;
}

}  // namespace gdn
