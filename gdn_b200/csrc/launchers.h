// launchers.h -- host-side launch functions shared between the translation units.
#pragma once
#include "common.cuh"

namespace gdn {

// mirrors of the kernel argument structs in dwide.cu
struct HeadArgs {
    const float* A;
    const float* V;
    const float* Wl;
    const float* bnc;
    float* xh1;          // [n][D] saved BatchNorm-1 input (training, D <= 128) or NULL: recompute from A
    const float* g1; const float* be1; const float* g2; const float* be2;
    const float* wo; const float* bo;
    int B, N, W, D, S, rps;
    const float* mask;
    uint32_t* bits;
    unsigned long long seed, offset;
    const unsigned long long* offset_dev;   // optional device-side addend to `offset` (CUDA-graph replays)
    float p_drop, scale;
    int training;
};
struct BwdArgs {
    const float* gpred;
    const float* c2;
    const float* c1;
    float* gV;
    float* gA;
};

// attention.cu
int launch_prep(const Shape& s, const float* x, const float* V, const gdn_layer_params* p,
                char* ctx, const CtxLayout& L, cudaStream_t st);
int launch_attn_fwd(const Shape& s, const int32_t* nbr, char* ctx, const CtxLayout& L, float* alpha,
                    const gdn_layer_params* p, float* out, int* fused_out, cudaStream_t st);
int launch_attn_alpha(const Shape& s, const int32_t* nbr, const char* ctx, const CtxLayout& L, float* alpha,
                      cudaStream_t st);
int launch_attn_bwd(const Shape& s, const int32_t* nbr, const char* ctx, const CtxLayout& L,
                    const float* gA, float* gsiT, float* gsjT, size_t zero_bytes, cudaStream_t st);
int launch_attn_tail(const Shape& s, const char* ctx, const CtxLayout& L, const float* V, const gdn_layer_params* p,
                     const float* gsiT, const float* gsjT, int accumulate, const double* part, int nrec,
                     float* part_u, float* part_e, double* sums, unsigned int* counter, gdn_layer_grads* g,
                     cudaStream_t st);

// dwide.cu
int launch_lin_fwd(const Shape& s, const float* A, const gdn_layer_params* p, float* out, cudaStream_t st);
int launch_lin_bwd(const Shape& s, const float* gout, const float* A, const gdn_layer_params* p, float* gA,
                   double* part, int* nrec, cudaStream_t st);
int launch_moments(const Shape& s, const float* A, double* part, int* nrec, cudaStream_t st);
int launch_fin_bn1(const Shape& s, const double* part, int nrec, double* sums, const gdn_layer_params* p, float* bnc,
                   const gdn_bn* bn, const gdn_sync* sync, cudaStream_t st);
int launch_fin_bn_eval(const Shape& s, const gdn_layer_params* p, const gdn_head_params* h, float* bnc,
                       cudaStream_t st);
int launch_fwd_stats2(const Shape& s, const HeadArgs& h, double* part, double* sums, const gdn_bn* bn, float* bnc,
                      const gdn_sync* sync, cudaStream_t st);
int launch_fwd_out(const Shape& s, const HeadArgs& h, float* pred, cudaStream_t st);
int launch_bwd1(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, double* sums, gdn_head_grads* gh,
                float* c2, const gdn_sync* sync, cudaStream_t st);
int launch_bwd2(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, double* sums, gdn_head_grads* gh,
                float* c1, float* gV_final, const gdn_sync* sync, cudaStream_t st);
int launch_bwd3(const Shape& s, const HeadArgs& h, const BwdArgs& g, double* part, int* nrec, cudaStream_t st);

// graph_build.cu
size_t graph_build_ws_bytes(int N, int D, int K);
int launch_graph_build(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr, void* ws,
                       size_t ws_bytes, int use_tc, float* kth, float margin, cudaStream_t st);

// scoring.cu
// optim.cu
int launch_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, long long step, const long long* step_dev, float grad_scale,
                     cudaStream_t st);

int launch_nvls_adam(const float* p_local, float* p_mc, const float* g_mc, float* m, float* v, long long lo, long long cnt,
                     float lr, float beta1, float beta2, float eps, float weight_decay, long long step, float grad_scale,
                     cudaStream_t st);

// metrics.cu
int launch_f1_sweep(const double* sorted_scores, const float* labels_sorted, int T, const int* k_pred, const int* k_thr,
                    int S, double* fmeas, double* thresholds, cudaStream_t st);
int launch_binary_counts(const double* scores, const float* labels, int T, double threshold, unsigned long long* counts,
                         cudaStream_t st);
int launch_auc_ranksum(const double* sorted_scores, const float* labels_sorted, int T, double* ranksum,
                       unsigned long long* npos, cudaStream_t st);

// windows.cu
int launch_window_batch(const float* series, const float* labels, int N, int T, int W, const int* win_end, int B,
                        float* x, float* y, float* lab, int* err, cudaStream_t st);

size_t score_ws_bytes(int T, int N);
int launch_score(const float* pred, const float* gt, int T, int N, double* scores, double* top1, double* stats,
                 void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace gdn
