// metrics.cu -- threshold sweep / F1, confusion counts and ROC-AUC rank sum on the device
// (util/data.py:28-51 eval_scores, evaluate.py:129-158 get_best_performance_data).
//
// The reference calls sklearn.f1_score 400 times over T ticks plus a list.index() scan per step (2.6-5.5 s per
// evaluation).  With the ticks sorted once by score (ordinal ranks = positions in a stable ascending sort) a
// threshold step i marks exactly the ticks at sorted positions >= k_i as anomalies, so
//   TP_i = sum of labels over the suffix [k_i, T),  F1_i = 2 TP_i / (P + (T - k_i))       (sklearn's formula)
// and the step's score threshold is the sorted score at one known position.  The k_i come from the host (400
// float64 products, evaluated exactly as the reference does).
#include "common.cuh"
#include "launchers.h"

namespace gdn {

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (wid == 0) {
        t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;                                            // valid in warp 0
}

// one CTA per threshold step; labels_sorted[r] in {0, 1} in ascending score order
__global__ void __launch_bounds__(256)
k_f1_sweep(const double* __restrict__ sorted_scores, const float* __restrict__ labels_sorted, int T,
           const int* __restrict__ k_pred, const int* __restrict__ k_thr, double* __restrict__ fmeas,
           double* __restrict__ thresholds) {
    __shared__ double sh[8];
    const int i = blockIdx.x;
    const int k = k_pred[i];
    double tp = 0.0, pos = 0.0;
    for (int r = threadIdx.x; r < T; r += blockDim.x) {
        const double l = labels_sorted[r] != 0.f ? 1.0 : 0.0;
        pos += l;
        if (r >= k) tp += l;
    }
    tp = block_sum(tp, sh);
    pos = block_sum(pos, sh);
    if (threadIdx.x == 0) {
        const double denom = pos + (double)(T - k);      // true positives + predicted positives = 2TP + FP + FN
        fmeas[i] = denom > 0.0 ? 2.0 * tp / denom : 0.0;
        const int q = k_thr[i];
        thresholds[i] = (q >= 0 && q < T) ? sorted_scores[q] : nan("");
    }
}

// counts[0..3] = TP, FP, FN, TN of (scores > threshold) against labels
__global__ void __launch_bounds__(256)
k_binary_counts(const double* __restrict__ scores, const float* __restrict__ labels, int T, double threshold,
                unsigned long long* __restrict__ counts) {
    __shared__ double sh[8];
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < T; r += gridDim.x * blockDim.x) {
        const bool p = scores[r] > threshold, l = labels[r] != 0.f;
        c[(p ? 0 : 2) + (l ? 0 : 1)] += 1.0;             // (p,l): TP=0  (p,!l): FP=1  (!p,l): FN=2  (!p,!l): TN=3
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double t = block_sum(c[q], sh);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(counts + q, (unsigned long long)t);
    }
}

// sum over the positive ticks of their tie-averaged 1-based rank, and the number of positives:
// AUC = (ranksum - P (P + 1) / 2) / (P (T - P)), the Mann-Whitney form of the trapezoidal ROC area
__global__ void __launch_bounds__(256)
k_auc_ranksum(const double* __restrict__ sorted_scores, const float* __restrict__ labels_sorted, int T,
              double* __restrict__ ranksum, unsigned long long* __restrict__ npos) {
    __shared__ double sh[8];
    double rs = 0.0, np_ = 0.0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < T; r += gridDim.x * blockDim.x) {
        if (labels_sorted[r] == 0.f) continue;
        const double v = sorted_scores[r];
        int lo = 0, hi = r;                              // first index with value == v
        while (lo < hi) { const int m = (lo + hi) >> 1; if (sorted_scores[m] < v) lo = m + 1; else hi = m; }
        const int first = lo;
        lo = r; hi = T - 1;                              // last index with value == v
        while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (sorted_scores[m] > v) hi = m - 1; else lo = m; }
        rs += 0.5 * ((double)first + (double)lo) + 1.0;
        np_ += 1.0;
    }
    rs = block_sum(rs, sh);
    np_ = block_sum(np_, sh);
    if (threadIdx.x == 0) {
        if (rs != 0.0) atomicAdd(ranksum, rs);
        if (np_ != 0.0) atomicAdd(npos, (unsigned long long)np_);
    }
}

int launch_f1_sweep(const double* sorted_scores, const float* labels_sorted, int T, const int* k_pred, const int* k_thr,
                    int S, double* fmeas, double* thresholds, cudaStream_t st) {
    k_f1_sweep<<<S, 256, 0, st>>>(sorted_scores, labels_sorted, T, k_pred, k_thr, fmeas, thresholds);
    GDN_CHECK_LAUNCH("k_f1_sweep");
    return 0;
}

int launch_binary_counts(const double* scores, const float* labels, int T, double threshold, unsigned long long* counts,
                         cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset counts");
    int g = (T + 255) / 256;
    if (g > 2 * num_sms()) g = 2 * num_sms();
    k_binary_counts<<<g, 256, 0, st>>>(scores, labels, T, threshold, counts);
    GDN_CHECK_LAUNCH("k_binary_counts");
    return 0;
}

int launch_auc_ranksum(const double* sorted_scores, const float* labels_sorted, int T, double* ranksum,
                       unsigned long long* npos, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ranksum, 0, sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(npos, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset ranksum");
    int g = (T + 255) / 256;
    if (g > 2 * num_sms()) g = 2 * num_sms();
    k_auc_ranksum<<<g, 256, 0, st>>>(sorted_scores, labels_sorted, T, ranksum, npos);
    GDN_CHECK_LAUNCH("k_auc_ranksum");
    return 0;
}

}  // namespace gdn
