/* gdn_b200.h -- C ABI of libgdn_b200.so: the B200 (sm_100a) implementation of the GDN
 * forward/backward hot path (reference: SchlomoFeng/GDN).
 *
 * The reference is pure Python on top of PyTorch + torch-geometric 1.5.0; it has no FFI
 * of its own.  The entry points below are what a binding for this path replaces, one
 * per reference call site (file:line under the reference tree):
 *
 *   gdn_graph_build(_warm) models/GDN.py:143-159   cosine Gram + row-wise top-k
 *   gdn_graphlayer_fwd     models/graph_layer.py:53-117 (+ PyG propagate/softmax) on the
 *                          window-shared top-k graph built by models/GDN.py:161-165
 *   gdn_graphlayer_bwd     autograd of the above (train.py:72)
 *   gdn_csr_fwd / _bwd     models/graph_layer.py:53-117 on an arbitrary edge list
 *   gdn_fused_fwd          models/GDN.py:122-187 (GraphLayer + BN + ReLU + (x)embedding +
 *                          BN + ReLU + Dropout + Linear(D,1)), train or eval
 *   gdn_fused_bwd          autograd of models/GDN.py:122-187 (train.py:72)
 *   gdn_score              evaluate.py:48-68 + util/data.py:75-82 (+ evaluate.py:134-139)
 *   gdn_f1_sweep, gdn_binary_counts, gdn_auc_ranksum
 *                          util/data.py:28-51 (eval_scores) and evaluate.py:129-158
 *                          (get_best_performance_data): threshold sweep / F1, precision, recall, AUC
 *   gdn_adam_flat          train.py:31,73 (torch.optim.Adam step) on flat buffers, 1/G gradient scaling fused
 *   gdn_nvls_adam          the same step on G ranks: gradient all-reduce + Adam + parameter broadcast, one kernel
 *                          over NVSwitch multicast memory
 *   gdn_window_batch       datasets/TimeDataset.py:33-62 (+ the per-step transfer train.py:66):
 *                          window batches gathered from a device-resident series
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_; the library
 *     never allocates, frees or retains device memory: the caller owns every buffer,
 *     including the scratch `ws` and the saved-for-backward `ctx` blobs whose sizes the
 *     *_bytes() queries return;
 *   - all tensors are dense, row-major, float32 unless stated; sizes are ints;
 *   - `stream` is a cudaStream_t passed as void*; every kernel is enqueued on it and the
 *     call returns without synchronising (CUDA-graph capturable);
 *   - return value: 0 = ok, <0 = invalid argument / unsupported shape, >0 = cudaError_t;
 *     gdn_last_error() returns a thread-local message for the last non-zero return.
 *   - there is no CPU fallback anywhere.
 */
#ifndef GDN_B200_H
#define GDN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDN_B200_VERSION 100

/* Problem shape shared by the window-batched entry points.
 * n = B*N rows; the neighbour table `nbr` has Kp = K+1 slots per sensor. */
typedef struct gdn_dims {
    int B;   /* windows in the batch                          */
    int N;   /* sensors (node_num)                            */
    int W;   /* slide_win = GraphLayer in_channels  (<= 32)   */
    int D;   /* dim = GraphLayer out_channels       (<= 256, multiple of 4) */
    int K;   /* topk                                          */
} gdn_dims;

/* GraphLayer parameters (models/graph_layer.py:25-36), heads = 1. */
typedef struct gdn_layer_params {
    const float* lin_weight; /* [D, W]  gnn.lin.weight                      */
    const float* att_i;      /* [D]     gnn.att_i  (1,1,D flattened)        */
    const float* att_j;      /* [D]                                         */
    const float* att_em_i;   /* [D]                                         */
    const float* att_em_j;   /* [D]                                         */
    const float* bias;       /* [D]     gnn.bias (may be NULL)              */
} gdn_layer_params;

typedef struct gdn_layer_grads {
    float* lin_weight; /* [D, W] */
    float* att_i;      /* [D]    */
    float* att_j;      /* [D]    */
    float* att_em_i;   /* [D]    */
    float* att_em_j;   /* [D]    */
    float* bias;       /* [D] (may be NULL) */
    float* embedding;  /* [N, D] gradient w.r.t. the sensor embedding V (overwritten) */
} gdn_layer_grads;

/* BatchNorm1d(D) state (models/GDN.py:67 and :96). */
typedef struct gdn_bn {
    const float* weight;        /* [D] gamma */
    const float* bias;          /* [D] beta  */
    float* running_mean;        /* [D] updated in training mode            */
    float* running_var;         /* [D] updated in training mode (unbiased) */
    int64_t* num_batches_tracked; /* scalar, += 1 in training mode (may be NULL) */
} gdn_bn;

/* Head of the model for out_layer_num == 1 (models/GDN.py:36,171-187). */
typedef struct gdn_head_params {
    gdn_bn bn1;            /* gnn_layers.0.bn                 */
    gdn_bn bn2;            /* bn_outlayer_in                  */
    const float* out_w;    /* [D] out_layer.mlp.0.weight      */
    const float* out_b;    /* [1] out_layer.mlp.0.bias        */
} gdn_head_params;

typedef struct gdn_head_grads {
    float* bn1_weight; float* bn1_bias;   /* [D] each */
    float* bn2_weight; float* bn2_bias;   /* [D] each */
    float* out_w;                         /* [D] */
    float* out_b;                         /* [1] */
} gdn_head_grads;

/* Dropout(0.2) (models/GDN.py:114,182).  Either an explicit keep-mask (test hook:
 * values in {0, 1/(1-p)}, layout [B, N, D]) or counter-based Philox4x32-10 keyed by
 * (seed, offset): one call yields eight 16-bit uniforms (keep iff u16 >= round(p * 65536));
 * with DPL = D/32 and d = l*DPL + j, element (b, i, d) uses counter (((b*N+i)*DPL + j)*4 + l/8),
 * 16-bit field l % 8.  The forward stores the keep bits (1 bit per element) for the backward. */
typedef struct gdn_dropout {
    const float* mask;   /* NULL -> Philox */
    uint64_t seed;
    uint64_t offset;
    float p;             /* 0 disables dropout */
    const uint64_t* offset_dev; /* optional DEVICE counter added to `offset` when the kernel runs: lets a
                                   captured CUDA graph draw a fresh mask on every replay (may be NULL) */
} gdn_dropout;

/* Synchronised BatchNorm for data-parallel training (models/GDN.py:77,179 evaluated on the GLOBAL batch instead of
 * this rank's shard).  The library owns no communicator: at the four points where a BatchNorm
 * needs batch-wide sums (two in the forward, two in the backward) it calls back with a small buffer of doubles to be
 * summed over the ranks IN PLACE, ordered on `stream`.  world = number of ranks (rows per rank must be equal). */
typedef struct gdn_sync {
    int world;
    int (*allreduce_sum_f64)(void* buf, long long count, void* user, void* stream);  /* 0 = ok */
    void* user;
} gdn_sync;

int         gdn_version(void);
const char* gdn_last_error(void);

/* ---- measurement hook (bench.py): when enabled, a CUDA event is recorded on the call's
 * stream after every kernel launch; gdn_profile_collect() synchronises and writes one text
 * line per kernel "<name> <launches> <total_ms>" and returns the number of launches. */
/* Number of kernel launches the library has enqueued so far in this process (launches recorded into a stream
 * capture count once, at capture time). */
long long gdn_launch_count(void);

/* ---- host side of the feed --------------------------------------------------------
 * Replaces train.py:63-66 `x.float()` (one thread, an intermediate tensor) + the blocking pageable `.to(device)`:
 * converts n doubles at `src` (pageable is fine) to floats at `dst` (normally a pinned staging buffer) in ONE pass, split
 * over `threads` plain host threads (a persistent pool inside the library: independent of OMP_NUM_THREADS and of the
 * caller's interpreter lock; AVX-512 convert + non-temporal stores where the CPU has them).  Host pointers, no CUDA call.
 * Returns 0, or -1 for a NULL pointer. */
int gdn_stage_f64_to_f32(const double* src, float* dst, size_t n, int threads);
int gdn_profile_enable(int on);
int gdn_profile_collect(char* buf, size_t buf_bytes);

/* ---- a1: learned graph (models/GDN.py:143-159) ---------------------------------------
 * V [N, D] -> idx [N, K] int64, descending cosine; EXACTLY equal cosines rank the lower column
 *          first.  That is a deterministic rule of ours, NOT torch.topk's: the reference's
 *          tie order is an artefact of libstdc++'s partial_sort/nth_element (SURVEY.md
 *          section 7.1).  Rows whose top-(K+1) reference cosines are > 1e-6 apart equal
 *          torch.topk bit for bit; tie-affected rows are equal after canonicalising inside
 *          tau-clusters (oracle/topk_protocol.py, counts asserted in tests/ and recorded
 *          by bench.py);
 *          and nbr [N, K+1] int32: the neighbour list GraphLayer actually uses after
 *          remove_self_loops/add_self_loops (models/graph_layer.py:61-63): the non-self
 *          top-k entries in order, then the sensor itself, then negative padding (any negative value ends
 *          the list; when the sensor was in its own top-k the LAST slot holds -2 - its position there, the
 *          other padding slots -1, so idx is recoverable from nbr alone).
 * Either output may be NULL.  use_tensor_cores: 0 = exact fp32 CUDA-core Gram (a warp-per-row kernel up to
 * 2048 sensors, a 64x64 tile kernel beyond), 1 = tcgen05 split-precision Gram with exact fp32 re-scoring
 * (N >= 1024, dim 64 or 128, topk <= 72), -1 = choose by N.  All three produce the same bits. */
size_t gdn_graph_build_ws_bytes(int N, int D, int K);
int    gdn_graph_build(const float* V, int N, int D, int K, int64_t* idx, int32_t* nbr,
                       void* ws, size_t ws_bytes, int use_tensor_cores, void* stream);

/* Same, warm-started: kth [N] float (in/out).  In: the K-th largest cosine of every row from the
 * PREVIOUS build (anything <= -2 or NaN = no hint); the tcgen05 engine admits only values above
 * kth[i] - margin, which removes almost all selection work when the embedding moved by one optimiser
 * step.  Rows whose hint turns out stale are recomputed exactly, so the result never depends on the
 * hint.  margin = +infinity says "kth holds no valid hints yet" (first build): a cold sweep, kth is only written.
 * Out: this build's K-th largest cosine per row. */
int    gdn_graph_build_warm(const float* V, int N, int D, int K, int64_t* idx, int32_t* nbr,
                            void* ws, size_t ws_bytes, int use_tensor_cores, float* kth, float margin,
                            void* stream);
/* Rows [row0, row1) of the same graph only -- the row-sharded build of the window-sharded data-parallel
 * trainer (optional exchange step): every rank builds its rows, the neighbour tables
 * are all-gathered.  row0 must be a multiple of 128, row1 a multiple of 128 or N; idx / nbr are the full
 * [N, K] / [N, K+1] arrays (only the range is written); kth may be NULL (no warm start). */
int    gdn_graph_build_rows(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr,
                            void* ws, size_t ws_bytes, int use_tensor_cores, float* kth, float margin, void* stream);

/* ---- a3/a4: GraphLayer on the window-shared graph ------------------------------------
 * x [B, N, W], V [N, D], nbr [N, K+1] -> out [B*N, D]
 * alpha (optional) [B*N, K+1]: attention weight of slot k of row (b, i) (0 for padding).
 * ctx: saved-for-backward blob of gdn_graphlayer_ctx_bytes(); ws: scratch. */
size_t gdn_graphlayer_ctx_bytes(const gdn_dims* d);
size_t gdn_graphlayer_ws_bytes(const gdn_dims* d);
int    gdn_graphlayer_fwd(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr,
                          const gdn_layer_params* p, float* out, float* alpha,
                          void* ctx, void* ws, size_t ws_bytes, void* stream);
int    gdn_graphlayer_bwd(const gdn_dims* d, const float* g_out, const float* V, const int32_t* nbr,
                          const gdn_layer_params* p, const void* ctx, gdn_layer_grads* g,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- a3/a4 on an arbitrary edge list (GraphLayer's own module boundary) ----------------
 * CSR by target: rowptr [n+1], col [E] (sources; self loops already fixed up by the
 * caller exactly as models/graph_layer.py:61-63 does), heads >= 1.
 * x [n, W], emb [n, D] (may be NULL... the reference requires it), lin_weight [H*D, W],
 * att_* [H*D], out_h [n, H, D] (per-head aggregate; the caller applies concat/mean+bias),
 * alpha [E, H]. */
int    gdn_csr_fwd(int n, int W, int D, int H, int64_t E, const int32_t* rowptr, const int32_t* col,
                   const float* x, const float* emb, const float* lin_weight,
                   const float* att_i, const float* att_j, const float* att_em_i, const float* att_em_j,
                   float* xl /*[n,H*D] scratch+saved*/, float* s_i /*[n,H]*/, float* s_j /*[n,H]*/,
                   float* out_h, float* alpha, float negative_slope, void* stream);
int    gdn_csr_bwd(int n, int W, int D, int H, int64_t E, const int32_t* rowptr, const int32_t* col,
                   const float* x, const float* emb, const float* lin_weight,
                   const float* att_i, const float* att_j, const float* att_em_i, const float* att_em_j,
                   const float* xl, const float* s_i, const float* s_j, const float* alpha,
                   const float* g_out_h /*[n,H,D]*/,
                   float* g_xl /*[n,H*D] scratch*/, float* g_si /*[n,H]*/, float* g_sj /*[n,H]*/,
                   float negative_slope, void* stream);

/* ---- a3-a7: whole GDN forward / backward, out_layer_num == 1 ---------------------------
 * training != 0: batch statistics + running-stat update + dropout; else running stats.
 * pred [B, N].  ctx keeps what the backward needs; ws is scratch. */
size_t gdn_fused_ctx_bytes(const gdn_dims* d);
size_t gdn_fused_ws_bytes(const gdn_dims* d);
int    gdn_fused_fwd(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr,
                     const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp,
                     int training, float* pred, void* ctx, void* ws, size_t ws_bytes, void* stream);
int    gdn_fused_bwd(const gdn_dims* d, const float* g_pred, const float* V, const int32_t* nbr,
                     const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp,
                     const void* ctx, gdn_layer_grads* g, gdn_head_grads* gh,
                     void* ws, size_t ws_bytes, void* stream);
/* Same with BatchNorm statistics (forward) and the BatchNorm backward's batch sums shared by `sync->world` ranks;
 * parameter gradients stay this rank's partial sums (the caller's gradient all-reduce completes them).  sync == NULL
 * or world <= 1: identical to the calls above. */
int    gdn_fused_fwd_sync(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr,
                          const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp,
                          int training, float* pred, void* ctx, void* ws, size_t ws_bytes, const gdn_sync* sync,
                          void* stream);
int    gdn_fused_bwd_sync(const gdn_dims* d, const float* g_pred, const float* V, const int32_t* nbr,
                          const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp,
                          const void* ctx, gdn_layer_grads* g, gdn_head_grads* gh,
                          void* ws, size_t ws_bytes, const gdn_sync* sync, void* stream);
/* attention weights of the last gdn_fused_fwd / gdn_graphlayer_fwd held in ctx:
 * alpha [B*N, K+1] (GNNLayer.att_weight_1, materialised only on demand). */
int    gdn_ctx_alpha(const gdn_dims* d, const int32_t* nbr, const void* ctx, float* alpha, void* stream);

/* ---- a8: test-time scoring (evaluate.py:48-68, util/data.py:75-82) ---------------------
 * pred, gt [T, N] float32 -> scores [N, T] float64 (normalised, smoothed error per sensor)
 * and top1 [T] float64 (max over sensors, evaluate.py:134-139 with topk=1); either may be
 * NULL.  stats (optional) [N, 2] float64 = (median, IQR) per sensor. */
size_t gdn_score_ws_bytes(int T, int N);
int    gdn_score(const float* pred, const float* gt, int T, int N, double* scores, double* top1,
                 double* stats, void* ws, size_t ws_bytes, void* stream);

/* ---- optimiser (train.py:31, 73) ----
 * torch.optim.Adam (amsgrad off) on flat float32 buffers of n elements: grads are multiplied by grad_scale first
 * (1/world_size after a sum all-reduce), weight_decay is the L2 form (grad += wd * param), step counts from 1. */
int    gdn_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, long long step, float grad_scale,
                     void* stream);
/* Same, with the step count read from DEVICE memory when the kernel runs (*step_dev >= 1; the caller increments it
 * on the same stream before the call): a captured CUDA graph advances Adam's bias correction on every replay. */
int    gdn_adam_flat_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, const long long* step_dev,
                         float grad_scale, void* stream);
/* Data-parallel step over NVSwitch multicast memory (train.py:73 on G ranks): all-reduce of the flat gradient,
 * Adam and the broadcast of the new parameters in one kernel.  params_mc / grads_mc are the MULTICAST addresses of
 * the symmetric flat parameter / gradient buffers (every rank's replica bound to one multicast object, e.g. through
 * torch.distributed._symmetric_memory); params_local is this rank's replica.  This rank owns elements
 * [lo, lo + count) (both multiples of 4) and holds Adam moments for that slice only (exp_avg, exp_avg_sq: `count`
 * floats).  The caller brackets the call with cross-rank barriers on the same stream. */
int    gdn_nvls_adam(const float* params_local, float* params_mc, const float* grads_mc, float* exp_avg,
                     float* exp_avg_sq, long long lo, long long count, float lr, float beta1, float beta2, float eps,
                     float weight_decay, long long step, float grad_scale, void* stream);

/* ---- evaluation metrics (util/data.py:28-51, evaluate.py:129-158) ----
 * sorted_scores [T] float64 ascending (stable), labels_sorted [T] float32 in {0,1} in the same order.
 * gdn_f1_sweep: for each of S steps, ticks at sorted positions >= k_pred[s] are predicted anomalous:
 *   fmeas[s] = 2 TP / (P + T - k_pred[s]) (0 if the denominator is 0), thresholds[s] = sorted_scores[k_thr[s]]
 *   (NaN if k_thr[s] is outside [0, T)).
 * gdn_binary_counts: counts[0..3] = TP, FP, FN, TN of (scores[t] > threshold) against labels[t] (unsorted is fine).
 * gdn_auc_ranksum: ranksum = sum over positive ticks of their tie-averaged 1-based rank, npos = number of positives. */
int    gdn_f1_sweep(const double* sorted_scores, const float* labels_sorted, int T, const int* k_pred, const int* k_thr,
                    int S, double* fmeas, double* thresholds, void* stream);
int    gdn_binary_counts(const double* scores, const float* labels, int T, double threshold, unsigned long long* counts,
                         void* stream);
int    gdn_auc_ranksum(const double* sorted_scores, const float* labels_sorted, int T, double* ranksum,
                       unsigned long long* npos, void* stream);

/* ---- data feed (datasets/TimeDataset.py:33-62) ----
 * series [N, T] float32 and labels [T] float32 (or NULL) stay resident on the device; a batch is B
 * window-end indices win_end[b] in [W, T):  x[b,i,w] = series[i, win_end[b]-W+w],  y[b,i] = series[i, win_end[b]],
 * lab[b] = labels[win_end[b]] (lab may be NULL).  err: one int, zero-initialised by the caller; receives
 * 1 + b of an out-of-range window (the batch rows of that window are set to zero; the flag is sticky until the
 * caller clears it, so it can be checked once per epoch instead of once per batch). */
int    gdn_window_batch(const float* series, const float* labels, int N, int T, int W, const int* win_end, int B,
                        float* x, float* y, float* lab, int* err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDN_B200_H */
