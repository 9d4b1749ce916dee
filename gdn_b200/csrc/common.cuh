// common.cuh -- shared helpers for libgdn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gdn_b200.h"

#define GDN_NEG_SLOPE 0.2f      // models/graph_layer.py:13
#define GDN_BN_EPS 1e-5f        // nn.BatchNorm1d default
#define GDN_BN_MOMENTUM 0.1f
#define GDN_SOFTMAX_EPS 1e-16f  // PyG 1.5.0 utils.softmax

namespace gdn {

void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define GDN_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            gdn::set_error(__VA_ARGS__);         \
            return -1;                           \
        }                                        \
    } while (0)

#define GDN_CHECK_LAUNCH(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return gdn::cuda_fail(e__, what); \
        gdn::count_launch();                                     \
        gdn::prof_mark(what);                                    \
    } while (0)

// optional per-kernel timing (gdn_profile_enable): an event is recorded after every launch
void count_launch();     // every kernel launch the library enqueues (also those recorded into a stream capture)
void prof_enter(cudaStream_t st, const char* api);
void prof_mark(const char* what);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int    ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int    round_up32(int v) { return (v + 31) & ~31; }

int num_sms();
// CTAs of k_attn_tail (its per-CTA records in the workspace are sized by this)
inline int tail_max_ctas() { return 2 * num_sms(); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the request grows: keeps the call out of
// steady-state launches (and out of CUDA-graph captures after the warm-up iterations)
cudaError_t ensure_dyn_smem_ptr(const void* kernel, size_t bytes);
template <typename K>
static inline cudaError_t ensure_dyn_smem(K kernel, size_t bytes) {
    return ensure_dyn_smem_ptr(reinterpret_cast<const void*>(kernel), bytes);
}

// ---------------------------------------------------------------------------------------
// ctx / ws layouts (byte offsets; every region 256-byte aligned)
// ---------------------------------------------------------------------------------------
struct Shape {
    int B, N, W, D, K, Kp, Bs;   // Kp = K+1 neighbour slots, Bs = B rounded up to 32
    long long n;                 // B*N
    int WP;                      // register padding of W: 8, 16 or 32
    int DPL;                     // channels per lane = D/32
    int S;                       // b-splits of the sensor-major D-wide passes
    int rows_per_split;
};

int make_shape(const gdn_dims* d, Shape* s, bool need_dwide);

struct CtxLayout {
    size_t xT;      // [N][Bs/32][WP/4][32][4]  x transposed: (window, tap quad) fastest, w zero-padded to WP
    size_t siT;     // [N][Bs]      s_i = x.u_i + e_i
    size_t sjT;     // [N][Bs]      s_j = x.u_j + e_j
    size_t mT;      // [N][Bs]      segment max of the leaky-relu'd logits
    size_t linvT;   // [N][Bs]      1 / (segment sum + 1e-16)
    size_t A;       // [n][W]       A[b,i,:] = sum_k alpha * x[b, S(i)_k, :]
    size_t uv;      // u_i[32], u_j[32]
    size_t ev;      // e_i[N], e_j[N]
    size_t bn;      // mean1[D], istd1[D], mean2[D], istd2[D]     (fused path)
    size_t bits;    // [n][D/32] uint32 dropout keep bits          (fused path)
    size_t xh1;     // [n][D] saved BatchNorm-1 input xh1 (fused training path, D <= 128)
    size_t flags;   // int[4]: training, ...
    size_t total;
};
CtxLayout ctx_layout(const Shape& s, bool fused);

struct WsLayout {
    size_t gA;       // [n][W]
    size_t gsiT;     // [N][Bs]
    size_t gsjT;     // [N][Bs]
    size_t tail_ctr; // one unsigned int right behind gsjT: k_attn_tail's ticket counter, zeroed by the same memset
    size_t part;     // per-CTA partial sums (doubles), size part_bytes
    size_t part_bytes;
    size_t gV;       // [S][N][D] partial embedding gradients (S > 1)
    size_t small;    // floats: coefficient vectors, small partials
    size_t sums;     // doubles: one reduced record (+ the 64 attention-scalar sums)
    size_t total;
};
WsLayout ws_layout(const Shape& s, bool fused);

size_t sums_stride();             // doubles between the local and the rank-summed copy inside WsLayout::sums (SyncBN)
#define GDN_SUMS_MAX_DOUBLES 8768  // >= max(D*W + D, 3D + 32, W*W + W) + 64 for D <= 256, W <= 32
#define GDN_MAX_PART_CTAS 1184     // 148 SMs x 8: upper bound on CTAs that write partials

}  // namespace gdn

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace gdn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : GDN_NEG_SLOPE * v; }

// Philox4x32-10 (Salmon et al. 2011), counter = (ctr_lo, ctr_hi, offset_lo, offset_hi).
__device__ __forceinline__ uint4 philox4x32_10(uint64_t ctr, uint64_t offset, uint64_t seed) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32);
    uint32_t c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

}  // namespace gdn
#endif
