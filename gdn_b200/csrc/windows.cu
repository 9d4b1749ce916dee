// windows.cu -- sliding-window batches formed on the device from a resident sensor series
// (datasets/TimeDataset.py:33-62: x_i = data[:, i-w:i], y_i = data[:, i]).
//
// The reference materialises every window on the host ([num_windows, N, W] doubles, W-fold redundant) and
// ships a batch -- plus a fully-connected edge_index nobody reads -- over PCIe every step.  Here the series
// [N, T] lives in HBM once; a batch is described by B window-end indices and gathered by one kernel:
//   x[b, i, w] = series[i, e_b - W + w]     y[b, i] = series[i, e_b]     label[b] = labels[e_b]
// HBM-bound: reads ~4 B N (W+1) bytes (overlapping windows hit L2), writes 4 B N (W+1).
#include "common.cuh"
#include "launchers.h"

namespace gdn {

// thread <-> (b, i), adjacent threads adjacent sensors: the [B, N, W] output is written as contiguous W-float
// rows (a warp writes 32 W floats back to back); the reads are W+1 consecutive floats per thread.
__global__ void __launch_bounds__(256)
k_window_batch(const float* __restrict__ series, const float* __restrict__ labels, int N, int T, int W,
               const int* __restrict__ win_end, int B, float* __restrict__ x, float* __restrict__ y,
               float* __restrict__ lab, int* __restrict__ err) {
    const long long total = (long long)B * N;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(r / N), i = (int)(r % N);
        const int e = win_end[b];
        if (e < W || e >= T) {                       // window [e-W, e] must lie inside the series: flag it and
            if (i == 0) atomicExch(err, b + 1);      // hand zeros (never uninitialised memory) to the model
            float* z = x + (size_t)r * W;
            for (int w = 0; w < W; ++w) z[w] = 0.f;
            y[r] = 0.f;
            if (i == 0 && lab != nullptr) lab[b] = 0.f;
            continue;
        }
        const float* src = series + (size_t)i * T + (e - W);
        float* dst = x + (size_t)r * W;
        for (int w = 0; w < W; ++w) dst[w] = __ldg(src + w);
        y[r] = __ldg(src + W);
        if (i == 0 && lab != nullptr) lab[b] = labels != nullptr ? __ldg(labels + e) : 0.f;
    }
}

int launch_window_batch(const float* series, const float* labels, int N, int T, int W, const int* win_end, int B,
                        float* x, float* y, float* lab, int* err, cudaStream_t st) {
    const long long total = (long long)B * N;
    long long g = (total + 255) / 256;
    if (g > 8LL * num_sms()) g = 8LL * num_sms();
    if (g < 1) g = 1;
    k_window_batch<<<(int)g, 256, 0, st>>>(series, labels, N, T, W, win_end, B, x, y, lab, err);
    GDN_CHECK_LAUNCH("k_window_batch");
    return 0;
}

}  // namespace gdn
