// scoring.cu -- test-time anomaly scoring (replaces evaluate.py:48-68 get_err_scores,
// util/data.py:75-82 get_err_median_and_iqr, the driver loop evaluate.py:6-36 and the
// max over sensors of evaluate.py:134-139).  All float64.
//
//   delta[t] = |pred[t,i] - gt[t,i]|                      (per sensor i)
//   med = np.median(delta),  iqr = percentile75 - percentile25  (numpy 'linear')
//   err[t] = (delta[t] - med) / (|iqr| + 1e-2)
//   s[t] = mean(err[t-3 .. t]) for t >= 3, else 0;   top1[t] = max_i s_i[t]
//
// Pipeline: (1) |pred-gt| transposed to sensor-major float64; (2) one CTA per sensor finds
// the six order statistics the median / quartiles need with an MSB-first radix select over
// the float64 bit patterns (non-negative doubles order like their bits) -- no sort, eight
// passes over the series; (3) one kernel normalises, smooths and takes the max over sensors.
#include "common.cuh"
#include "launchers.h"

namespace gdn {

__global__ void k_delta_transpose(const float* __restrict__ pred, const float* __restrict__ gt, int T, int N,
                                  double* __restrict__ dT) {
    __shared__ double tile[32][33];
    const int t0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, i = i0 + threadIdx.x;
        double v = 0.0;
        if (t < T && i < N) v = fabs((double)pred[(size_t)t * N + i] - (double)gt[(size_t)t * N + i]);
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, t = t0 + threadIdx.x;
        if (t < T && i < N) dT[(size_t)i * T + t] = tile[threadIdx.x][r];
    }
}

#define SC_NQ 6

// numpy 'linear' interpolation between order statistics a <= b at fraction t
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
    const double d = __dsub_rn(b, a);        // explicit roundings: no FMA contraction, as numpy
    return t < 0.5 ? __dadd_rn(a, __dmul_rn(d, t)) : __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t)));
}

__global__ void __launch_bounds__(256)
k_select_stats(const double* __restrict__ dT, int T, double* __restrict__ stats) {
    __shared__ unsigned hist[SC_NQ][256];
    __shared__ unsigned long long prefix[SC_NQ];
    __shared__ long long rank[SC_NQ];       // rank still to find inside the current prefix
    __shared__ double frac[2];
    const int i = blockIdx.x;
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(dT + (size_t)i * T);
    if (threadIdx.x == 0) {
        const double p25 = 0.25 * (double)(T - 1), p75 = 0.75 * (double)(T - 1);
        const long long l25 = (long long)floor(p25), l75 = (long long)floor(p75);
        rank[0] = (T - 1) / 2;                               // lower median
        rank[1] = T / 2;                                     // upper median
        rank[2] = l25;
        rank[3] = l25 + 1 < T ? l25 + 1 : T - 1;
        rank[4] = l75;
        rank[5] = l75 + 1 < T ? l75 + 1 : T - 1;
        frac[0] = p25 - (double)l25;
        frac[1] = p75 - (double)l75;
        for (int q = 0; q < SC_NQ; ++q) prefix[q] = 0ull;
    }
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int e = threadIdx.x; e < SC_NQ * 256; e += blockDim.x) (&hist[0][0])[e] = 0u;
        __syncthreads();
        unsigned long long pf[SC_NQ];
#pragma unroll
        for (int q = 0; q < SC_NQ; ++q) pf[q] = prefix[q];
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const unsigned long long k = keys[t];
            const unsigned long long hi = pass == 0 ? 0ull : (k >> (shift + 8));
            const unsigned dig = (unsigned)(k >> shift) & 255u;
#pragma unroll
            for (int q = 0; q < SC_NQ; ++q)
                if (hi == pf[q]) atomicAdd(&hist[q][dig], 1u);
        }
        __syncthreads();
        if (threadIdx.x < SC_NQ) {
            const int q = threadIdx.x;
            long long r = rank[q];
            int dsel = 255;
            for (int dgt = 0; dgt < 256; ++dgt) {
                const long long c = (long long)hist[q][dgt];
                if (r < c) { dsel = dgt; break; }
                r -= c;
            }
            rank[q] = r;
            prefix[q] = (prefix[q] << 8) | (unsigned long long)dsel;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v[SC_NQ];
        for (int q = 0; q < SC_NQ; ++q) v[q] = __longlong_as_double((long long)prefix[q]);
        const double med = (T & 1) ? v[0] : (v[0] + v[1]) / 2.0;
        const double q25 = np_lerp(v[2], v[3], frac[0]);
        const double q75 = np_lerp(v[4], v[5], frac[1]);
        stats[2 * i] = med;
        stats[2 * i + 1] = q75 - q25;
    }
}

__global__ void __launch_bounds__(256)
k_scores(const double* __restrict__ dT, const double* __restrict__ stats, int T, int N,
         double* __restrict__ scores, double* __restrict__ top1) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double best = -INFINITY;
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        if (t >= 3) {
            const double med = stats[2 * i], den = fabs(stats[2 * i + 1]) + 1e-2;
            const double* row = dT + (size_t)i * T;
            double acc = -0.0;
#pragma unroll
            for (int q = 3; q >= 0; --q) acc += (row[t - q] - med) / den;
            s = acc / 4.0;
        }
        if (scores != nullptr) scores[(size_t)i * T + t] = s;
        best = s > best ? s : best;
    }
    if (top1 != nullptr) top1[t] = best;
}

size_t score_ws_bytes(int T, int N) {
    return align_up((size_t)T * N * sizeof(double), 256) + align_up((size_t)2 * N * sizeof(double), 256);
}

int launch_score(const float* pred, const float* gt, int T, int N, double* scores, double* top1, double* stats,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
    (void)ws_bytes;
    double* dT = (double*)ws;
    double* st_int = (double*)((char*)ws + align_up((size_t)T * N * sizeof(double), 256));
    double* st_out = stats ? stats : st_int;
    dim3 grid(ceil_div(T, 32), ceil_div(N, 32));
    GDN_CHECK_ARG(grid.y <= 65535, "score: too many sensors (%d)", N);
    k_delta_transpose<<<grid, dim3(32, 8), 0, st>>>(pred, gt, T, N, dT);
    GDN_CHECK_LAUNCH("k_delta_transpose");
    k_select_stats<<<N, 256, 0, st>>>(dT, T, st_out);
    GDN_CHECK_LAUNCH("k_select_stats");
    if (scores != nullptr || top1 != nullptr) {
        k_scores<<<ceil_div(T, 256), 256, 0, st>>>(dT, st_out, T, N, scores, top1);
        GDN_CHECK_LAUNCH("k_scores");
    }
    return 0;
}

}  // namespace gdn
