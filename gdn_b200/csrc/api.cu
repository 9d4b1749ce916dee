// api.cu -- extern "C" entry points of libgdn_b200.so (see include/gdn_b200.h).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include "common.cuh"
#include "launchers.h"

namespace gdn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e > 0 ? (int)e : 1;
}

// ---------------------------------------------------------------------------------------
// per-kernel profiler: events between launches, read back by gdn_profile_collect()
// ---------------------------------------------------------------------------------------
#define GDN_PROF_CAP 16384
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_prof_ev[GDN_PROF_CAP];
static const char* g_prof_name[GDN_PROF_CAP];
static int g_prof_created = 0;
static cudaStream_t g_prof_stream = nullptr;

static void prof_push(const char* name) {
    if (g_prof_n >= GDN_PROF_CAP) return;
    if (g_prof_n >= g_prof_created) {
        if (cudaEventCreate(&g_prof_ev[g_prof_created]) != cudaSuccess) return;
        ++g_prof_created;
    }
    cudaEventRecord(g_prof_ev[g_prof_n], g_prof_stream);
    g_prof_name[g_prof_n] = name;
    ++g_prof_n;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void prof_enter(cudaStream_t st, const char* api) {
    g_prof_stream = st;
    if (g_prof_on) prof_push(api);      // names starting with '@' are API-entry markers
}

void prof_mark(const char* what) {
    if (g_prof_on) prof_push(what);
}

cudaError_t ensure_dyn_smem_ptr(const void* kernel, size_t bytes) {
    // high-water mark per (device, kernel): cudaFuncSetAttribute is per device/context, and template
    // instantiations share a type, hence the function pointer as key.  Guarded: several host threads (one per
    // GPU in a multi-device process) may launch concurrently.
    static std::mutex mu;
    static const void* keys[1024];
    static int devs[1024];
    static size_t have[1024];
    static int n = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    int k = 0;
    for (; k < n; ++k)
        if (keys[k] == kernel && devs[k] == dev) break;
    if (k == n) {
        if (n == 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        keys[n] = kernel;
        devs[n] = dev;
        have[n] = 0;   // static + dynamic may already exceed the 48 KB default: always opt in once
        ++n;
    }
    if (bytes <= have[k]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have[k] = bytes;
    return e;
}

int num_sms() {
    static std::atomic<int> cached[64];            // per device ordinal; 0 = not queried yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        int n = 0;
        v = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

size_t sums_stride() { return GDN_SUMS_MAX_DOUBLES; }

int make_shape(const gdn_dims* d, Shape* s, bool need_dwide) {
    GDN_CHECK_ARG(d != nullptr, "dims is NULL");
    GDN_CHECK_ARG(d->B >= 1 && d->N >= 1, "B=%d N=%d must be positive", d->B, d->N);
    GDN_CHECK_ARG(d->W >= 1 && d->W <= 32, "slide_win W=%d unsupported (1..32)", d->W);
    GDN_CHECK_ARG(d->K >= 1 && d->K <= d->N, "topk K=%d must be in 1..N=%d", d->K, d->N);
    GDN_CHECK_ARG(d->D >= 1, "dim D=%d must be positive", d->D);
    if (need_dwide)
        GDN_CHECK_ARG(d->D % 32 == 0 && d->D <= 256 && (d->D == 32 || d->D == 64 || d->D == 128 || d->D == 256),
                      "dim D=%d unsupported by the fused kernels (32, 64, 128, 256)", d->D);
    s->B = d->B; s->N = d->N; s->W = d->W; s->D = d->D; s->K = d->K;
    s->Kp = d->K + 1;
    s->Bs = round_up32(d->B);
    s->n = (long long)d->B * d->N;
    s->WP = d->W <= 8 ? 8 : (d->W <= 16 ? 16 : 32);
    s->DPL = d->D / 32;
    GDN_CHECK_ARG((long long)d->N * s->Bs * s->WP < (1ll << 31), "N*B*W = %lld elements exceed the 32-bit gather offsets",
                  (long long)d->N * s->Bs * s->WP);
    // sensor-major passes: tasks = N * S; aim for >= 8 tasks per resident warp slot
    const long long want = (long long)num_sms() * 8 * 4;
    int S = (int)((want + d->N - 1) / d->N);
    if (S < 1) S = 1;
    if (S > d->B) S = d->B;
    int rps = (d->B + S - 1) / S;
    S = (d->B + rps - 1) / rps;
    s->S = S;
    s->rows_per_split = rps;
    return 0;
}

static size_t take(size_t* off, size_t bytes) {
    const size_t at = *off;
    *off = align_up(at + bytes, 256);
    return at;
}

CtxLayout ctx_layout(const Shape& s, bool fused) {
    CtxLayout L;
    size_t off = 0;
    const size_t nb = (size_t)s.N * s.Bs * sizeof(float);
    L.xT = take(&off, nb * s.WP);
    L.siT = take(&off, nb);
    L.sjT = take(&off, nb);
    L.mT = take(&off, nb);
    L.linvT = take(&off, nb);
    L.A = take(&off, (size_t)s.n * s.W * sizeof(float));
    L.uv = take(&off, 64 * sizeof(float));
    L.ev = take(&off, (size_t)2 * s.N * sizeof(float));
    L.bn = take(&off, fused ? (size_t)8 * s.D * sizeof(float) : 0);
    L.bits = take(&off, fused ? (size_t)s.n * s.DPL * sizeof(uint32_t) : 0);
    L.flags = take(&off, 4 * sizeof(int));
    // last, so that the attention part of the layout is the same with and without it
    L.xh1 = take(&off, (fused && s.D <= 128) ? (size_t)s.n * s.D * sizeof(float) : 0);
    L.total = off;
    return L;
}

WsLayout ws_layout(const Shape& s, bool fused) {
    WsLayout L;
    size_t off = 0;
    L.gA = take(&off, (size_t)s.n * s.W * sizeof(float));
    const size_t nb = (size_t)s.N * s.Bs * sizeof(float);
    L.gsiT = take(&off, nb);
    L.gsjT = take(&off, nb);
    L.tail_ctr = take(&off, 256);
    size_t rec = (size_t)s.D * s.W + s.D;
    if ((size_t)3 * s.D + 32 > rec) rec = (size_t)3 * s.D + 32;
    if ((size_t)s.W * s.W + s.W > rec) rec = (size_t)s.W * s.W + s.W;
    L.part_bytes = (size_t)4 * num_sms() * rec * sizeof(double);
    L.part = take(&off, L.part_bytes);
    L.gV = take(&off, (fused && s.S > 1) ? (size_t)s.S * s.N * s.D * sizeof(float) : 0);
    // small: c2[2D] c1[2D] part_u[ctas*64] part_e[ctas*2D]   (floats; ctas = tail_max_ctas())
    L.small = take(&off, ((size_t)4 * s.D + (size_t)tail_max_ctas() * (64 + 2 * s.D)) * sizeof(float));
    L.sums = take(&off, (size_t)2 * GDN_SUMS_MAX_DOUBLES * sizeof(double));   // local record | rank-summed copy (SyncBN)
    (void)rec;
    L.total = off;
    return L;
}

struct Small {
    float *c2, *c1, *part_u, *part_e;
};
static Small small_ptrs(const Shape& s, char* ws, const WsLayout& L) {
    Small m;
    float* p = (float*)(ws + L.small);
    m.c2 = p; p += 2 * s.D;
    m.c1 = p; p += 2 * s.D;
    m.part_u = p; p += (size_t)tail_max_ctas() * 64;
    m.part_e = p;
    return m;
}

static HeadArgs head_args(const Shape& s, const char* ctx, const CtxLayout& L, const float* V,
                          const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp, int training) {
    HeadArgs a;
    a.A = (const float*)(ctx + L.A);
    a.V = V;
    a.Wl = p->lin_weight;
    a.bnc = (const float*)(ctx + L.bn);
    // training passes read the BatchNorm-1 input back instead of recomputing Wl.A (measured: the recompute
    // passes are FMA-issue bound at 0.30-0.64 ms each at C5; re-reading n*D floats is HBM-bound at ~0.1 ms)
    a.xh1 = (training && s.D <= 128) ? (float*)(const_cast<char*>(ctx) + L.xh1) : nullptr;
    a.g1 = h->bn1.weight; a.be1 = h->bn1.bias; a.g2 = h->bn2.weight; a.be2 = h->bn2.bias;
    a.wo = h->out_w; a.bo = h->out_b;
    a.B = s.B; a.N = s.N; a.W = s.W; a.D = s.D; a.S = s.S; a.rps = s.rows_per_split;
    a.mask = dp ? dp->mask : nullptr;
    a.bits = (uint32_t*)(const_cast<char*>(ctx) + L.bits);
    a.seed = dp ? dp->seed : 0ull;
    a.offset = dp ? dp->offset : 0ull;
    a.offset_dev = dp ? (const unsigned long long*)dp->offset_dev : nullptr;
    a.p_drop = dp ? dp->p : 0.f;
    a.scale = (dp && dp->p > 0.f && dp->p < 1.f) ? 1.f / (1.f - dp->p) : 1.f;
    a.training = training;
    return a;
}

}  // namespace gdn

using namespace gdn;

extern "C" {

int gdn_version(void) { return GDN_B200_VERSION; }
const char* gdn_last_error(void) { return g_err; }

long long gdn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------- profiler
int gdn_profile_enable(int on) {
    g_prof_on = on != 0;
    g_prof_n = 0;
    return 0;
}

// Synchronises on the last recorded event and writes one line per kernel name:
// "<name> <launches> <total_ms>\n" (time from the previous event on the stream to the event
// recorded after the launch).  Returns the number of launches recorded, <0 on error.
int gdn_profile_collect(char* buf, size_t buf_bytes) {
    if (buf == nullptr || buf_bytes == 0) { set_error("profile_collect: NULL buffer"); return -1; }
    buf[0] = 0;
    if (g_prof_n == 0) return 0;
    cudaError_t e = cudaEventSynchronize(g_prof_ev[g_prof_n - 1]);
    if (e != cudaSuccess) { cuda_fail(e, "profile_collect"); return -1; }
    const int MAXN = 64;
    const char* names[MAXN];
    int counts[MAXN];
    double totals[MAXN];
    int nn = 0, launches = 0;
    for (int i = 0; i < g_prof_n; ++i) {
        if (g_prof_name[i][0] == '@') continue;
        float ms = 0.f;
        if (i == 0 || cudaEventElapsedTime(&ms, g_prof_ev[i - 1], g_prof_ev[i]) != cudaSuccess) continue;
        int k = 0;
        for (; k < nn; ++k) if (strcmp(names[k], g_prof_name[i]) == 0) break;
        if (k == nn) { if (nn == MAXN) continue; names[nn] = g_prof_name[i]; counts[nn] = 0; totals[nn] = 0.0; ++nn; }
        counts[k] += 1;
        totals[k] += (double)ms;
        ++launches;
    }
    size_t off = 0;
    for (int k = 0; k < nn; ++k) {
        int w = snprintf(buf + off, buf_bytes - off, "%s %d %.6f\n", names[k], counts[k], totals[k]);
        if (w < 0 || (size_t)w >= buf_bytes - off) break;
        off += (size_t)w;
    }
    g_prof_n = 0;
    return launches;
}

// ------------------------------------------------------------------------------- graph
size_t gdn_graph_build_ws_bytes(int N, int D, int K) { return graph_build_ws_bytes(N, D, K); }

int gdn_graph_build(const float* V, int N, int D, int K, int64_t* idx, int32_t* nbr, void* ws, size_t ws_bytes,
                    int use_tensor_cores, void* stream) {
    GDN_CHECK_ARG(V != nullptr, "V is NULL");
    GDN_CHECK_ARG(N >= 1 && D >= 1 && K >= 1 && K <= N, "graph_build: bad shape N=%d D=%d K=%d", N, D, K);
    prof_enter((cudaStream_t)stream, "@graph_build");
    return launch_graph_build(V, N, D, K, 0, N, idx, nbr, ws, ws_bytes, use_tensor_cores, nullptr, 0.f, (cudaStream_t)stream);
}

int gdn_graph_build_warm(const float* V, int N, int D, int K, int64_t* idx, int32_t* nbr, void* ws, size_t ws_bytes,
                         int use_tensor_cores, float* kth, float margin, void* stream) {
    GDN_CHECK_ARG(V != nullptr, "V is NULL");
    GDN_CHECK_ARG(N >= 1 && D >= 1 && K >= 1 && K <= N, "graph_build: bad shape N=%d D=%d K=%d", N, D, K);
    GDN_CHECK_ARG(kth != nullptr && margin >= 0.f, "graph_build_warm: kth is NULL or margin < 0");
    prof_enter((cudaStream_t)stream, "@graph_build");
    return launch_graph_build(V, N, D, K, 0, N, idx, nbr, ws, ws_bytes, use_tensor_cores, kth, margin, (cudaStream_t)stream);
}

int gdn_graph_build_rows(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr, void* ws,
                         size_t ws_bytes, int use_tensor_cores, float* kth, float margin, void* stream) {
    GDN_CHECK_ARG(V != nullptr, "V is NULL");
    GDN_CHECK_ARG(N >= 1 && D >= 1 && K >= 1 && K <= N, "graph_build: bad shape N=%d D=%d K=%d", N, D, K);
    GDN_CHECK_ARG(kth == nullptr || margin >= 0.f, "graph_build_rows: margin < 0");
    prof_enter((cudaStream_t)stream, "@graph_build");
    return launch_graph_build(V, N, D, K, row0, row1, idx, nbr, ws, ws_bytes, use_tensor_cores, kth, margin,
                              (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- GraphLayer (shared graph)
size_t gdn_graphlayer_ctx_bytes(const gdn_dims* d) {
    Shape s;
    if (make_shape(d, &s, true)) return 0;
    return ctx_layout(s, false).total;
}
size_t gdn_graphlayer_ws_bytes(const gdn_dims* d) {
    Shape s;
    if (make_shape(d, &s, true)) return 0;
    return ws_layout(s, false).total;
}

int gdn_graphlayer_fwd(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr,
                       const gdn_layer_params* p, float* out, float* alpha, void* ctx_, void* ws, size_t ws_bytes,
                       void* stream) {
    Shape s;
    if (int rc = make_shape(d, &s, true)) return rc;
    GDN_CHECK_ARG(x && V && nbr && p && out && ctx_, "graphlayer_fwd: NULL argument");
    const CtxLayout L = ctx_layout(s, false);
    cudaStream_t st = (cudaStream_t)stream;
    char* ctx = (char*)ctx_;
    (void)ws; (void)ws_bytes;
    prof_enter(st, "@graphlayer_fwd");
    if (int rc = launch_prep(s, x, V, p, ctx, L, st)) return rc;
    int fused_out = 0;
    if (int rc = launch_attn_fwd(s, nbr, ctx, L, alpha, p, out, &fused_out, st)) return rc;
    if (fused_out) return 0;
    return launch_lin_fwd(s, (const float*)(ctx + L.A), p, out, st);
}

int gdn_graphlayer_bwd(const gdn_dims* d, const float* g_out, const float* V, const int32_t* nbr,
                       const gdn_layer_params* p, const void* ctx_, gdn_layer_grads* g, void* ws_, size_t ws_bytes,
                       void* stream) {
    Shape s;
    if (int rc = make_shape(d, &s, true)) return rc;
    GDN_CHECK_ARG(g_out && V && nbr && p && ctx_ && g && ws_, "graphlayer_bwd: NULL argument");
    GDN_CHECK_ARG(g->lin_weight && g->att_i && g->att_j && g->att_em_i && g->att_em_j && g->embedding,
                  "graphlayer_bwd: NULL gradient buffer");
    const CtxLayout L = ctx_layout(s, false);
    const WsLayout WL = ws_layout(s, false);
    GDN_CHECK_ARG(ws_bytes >= WL.total, "graphlayer_bwd: workspace too small (%zu < %zu)", ws_bytes, WL.total);
    cudaStream_t st = (cudaStream_t)stream;
    const char* ctx = (const char*)ctx_;
    char* ws = (char*)ws_;
    const Small sm = small_ptrs(s, ws, WL);
    float* gA = (float*)(ws + WL.gA);
    double* part = (double*)(ws + WL.part);
    int nrec = 0;
    prof_enter(st, "@graphlayer_bwd");
    if (int rc = launch_lin_bwd(s, g_out, (const float*)(ctx + L.A), p, gA, part, &nrec, st)) return rc;
    float* gsiT = (float*)(ws + WL.gsiT);
    float* gsjT = (float*)(ws + WL.gsjT);
    if (int rc = launch_attn_bwd(s, nbr, ctx, L, gA, gsiT, gsjT, WL.tail_ctr + 256 - WL.gsjT, st)) return rc;
    return launch_attn_tail(s, ctx, L, V, p, gsiT, gsjT, 0, part, nrec, sm.part_u, sm.part_e, (double*)(ws + WL.sums),
                            (unsigned int*)(ws + WL.tail_ctr), g, st);
}

// ------------------------------------------------------------------------------- fused GDN
size_t gdn_fused_ctx_bytes(const gdn_dims* d) {
    Shape s;
    if (make_shape(d, &s, true)) return 0;
    return ctx_layout(s, true).total;
}
size_t gdn_fused_ws_bytes(const gdn_dims* d) {
    Shape s;
    if (make_shape(d, &s, true)) return 0;
    return ws_layout(s, true).total;
}

int gdn_fused_fwd(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr, const gdn_layer_params* p,
                  const gdn_head_params* h, const gdn_dropout* dp, int training, float* pred, void* ctx_, void* ws_,
                  size_t ws_bytes, void* stream) {
    return gdn_fused_fwd_sync(d, x, V, nbr, p, h, dp, training, pred, ctx_, ws_, ws_bytes, nullptr, stream);
}

int gdn_fused_fwd_sync(const gdn_dims* d, const float* x, const float* V, const int32_t* nbr, const gdn_layer_params* p,
                       const gdn_head_params* h, const gdn_dropout* dp, int training, float* pred, void* ctx_, void* ws_,
                       size_t ws_bytes, const gdn_sync* sync, void* stream) {
    Shape s;
    if (int rc = make_shape(d, &s, true)) return rc;
    GDN_CHECK_ARG(x && V && nbr && p && h && pred && ctx_ && ws_, "fused_fwd: NULL argument");
    GDN_CHECK_ARG(p->bias != nullptr, "fused_fwd: gnn.bias must be given (pass zeros for bias=False)");
    GDN_CHECK_ARG(!training || s.n > 1, "fused_fwd: BatchNorm needs more than one row in training mode");
    GDN_CHECK_ARG(!dp || (dp->p >= 0.f && dp->p < 1.f), "fused_fwd: dropout p must be in [0, 1)");
    const CtxLayout L = ctx_layout(s, true);
    const WsLayout WL = ws_layout(s, true);
    GDN_CHECK_ARG(ws_bytes >= WL.total, "fused_fwd: workspace too small (%zu < %zu)", ws_bytes, WL.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ctx = (char*)ctx_;
    char* ws = (char*)ws_;
    float* bnc = (float*)(ctx + L.bn);
    double* part = (double*)(ws + WL.part);
    prof_enter(st, "@fused_fwd");
    if (int rc = launch_prep(s, x, V, p, ctx, L, st)) return rc;
    if (int rc = launch_attn_fwd(s, nbr, ctx, L, nullptr, nullptr, nullptr, nullptr, st)) return rc;
    const HeadArgs ha = head_args(s, ctx, L, V, p, h, dp, training);
    if (training) {
        int nrec = 0;
        if (int rc = launch_moments(s, ha.A, part, &nrec, st)) return rc;
        double* sums = (double*)(ws + WL.sums);
        if (int rc = launch_fin_bn1(s, part, nrec, sums, p, bnc, &h->bn1, sync, st)) return rc;
        if (int rc = launch_fwd_stats2(s, ha, part, sums, &h->bn2, bnc, sync, st)) return rc;
    } else {
        if (int rc = launch_fin_bn_eval(s, p, h, bnc, st)) return rc;
    }
    return launch_fwd_out(s, ha, pred, st);
}

int gdn_fused_bwd(const gdn_dims* d, const float* g_pred, const float* V, const int32_t* nbr,
                  const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp, const void* ctx_,
                  gdn_layer_grads* g, gdn_head_grads* gh, void* ws_, size_t ws_bytes, void* stream) {
    return gdn_fused_bwd_sync(d, g_pred, V, nbr, p, h, dp, ctx_, g, gh, ws_, ws_bytes, nullptr, stream);
}

int gdn_fused_bwd_sync(const gdn_dims* d, const float* g_pred, const float* V, const int32_t* nbr,
                       const gdn_layer_params* p, const gdn_head_params* h, const gdn_dropout* dp, const void* ctx_,
                       gdn_layer_grads* g, gdn_head_grads* gh, void* ws_, size_t ws_bytes, const gdn_sync* sync,
                       void* stream) {
    Shape s;
    if (int rc = make_shape(d, &s, true)) return rc;
    GDN_CHECK_ARG(g_pred && V && nbr && p && h && ctx_ && g && gh && ws_, "fused_bwd: NULL argument");
    GDN_CHECK_ARG(g->lin_weight && g->att_i && g->att_j && g->att_em_i && g->att_em_j && g->bias && g->embedding,
                  "fused_bwd: NULL layer gradient buffer");
    GDN_CHECK_ARG(gh->bn1_weight && gh->bn1_bias && gh->bn2_weight && gh->bn2_bias && gh->out_w && gh->out_b,
                  "fused_bwd: NULL head gradient buffer");
    const CtxLayout L = ctx_layout(s, true);
    const WsLayout WL = ws_layout(s, true);
    GDN_CHECK_ARG(ws_bytes >= WL.total, "fused_bwd: workspace too small (%zu < %zu)", ws_bytes, WL.total);
    cudaStream_t st = (cudaStream_t)stream;
    const char* ctx = (const char*)ctx_;
    char* ws = (char*)ws_;
    const Small sm = small_ptrs(s, ws, WL);
    double* part = (double*)(ws + WL.part);
    const HeadArgs ha = head_args(s, ctx, L, V, p, h, dp, /*training=*/1);
    BwdArgs ba;
    ba.gpred = g_pred;
    ba.c2 = sm.c2;
    ba.c1 = sm.c1;
    ba.gV = s.S > 1 ? (float*)(ws + WL.gV) : g->embedding;
    ba.gA = (float*)(ws + WL.gA);
    int nrec = 0;
    prof_enter(st, "@fused_bwd");
    double* sums = (double*)(ws + WL.sums);
    if (int rc = launch_bwd1(s, ha, ba, part, sums, gh, sm.c2, sync, st)) return rc;
    if (int rc = launch_bwd2(s, ha, ba, part, sums, gh, sm.c1, g->embedding, sync, st)) return rc;
    if (int rc = launch_bwd3(s, ha, ba, part, &nrec, st)) return rc;
    float* gsiT = (float*)(ws + WL.gsiT);
    float* gsjT = (float*)(ws + WL.gsjT);
    if (int rc = launch_attn_bwd(s, nbr, ctx, L, ba.gA, gsiT, gsjT, WL.tail_ctr + 256 - WL.gsjT, st)) return rc;
    return launch_attn_tail(s, ctx, L, V, p, gsiT, gsjT, 1, part, nrec, sm.part_u, sm.part_e, sums,
                            (unsigned int*)(ws + WL.tail_ctr), g, st);
}

int gdn_ctx_alpha(const gdn_dims* d, const int32_t* nbr, const void* ctx_, float* alpha, void* stream) {
    Shape s;
    if (int rc = make_shape(d, &s, false)) return rc;
    GDN_CHECK_ARG(nbr && ctx_ && alpha, "ctx_alpha: NULL argument");
    // the attention part of the ctx layout does not depend on the fused flag
    const CtxLayout L = ctx_layout(s, false);
    prof_enter((cudaStream_t)stream, "@ctx_alpha");
    return launch_attn_alpha(s, nbr, (const char*)ctx_, L, alpha, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- scoring
size_t gdn_score_ws_bytes(int T, int N) { return score_ws_bytes(T, N); }

int gdn_score(const float* pred, const float* gt, int T, int N, double* scores, double* top1, double* stats,
              void* ws, size_t ws_bytes, void* stream) {
    GDN_CHECK_ARG(pred && gt, "score: NULL input");
    GDN_CHECK_ARG(T >= 1 && N >= 1, "score: bad shape T=%d N=%d", T, N);
    GDN_CHECK_ARG(ws != nullptr && ws_bytes >= score_ws_bytes(T, N), "score: workspace too small");
    prof_enter((cudaStream_t)stream, "@score");
    return launch_score(pred, gt, T, N, scores, top1, stats, ws, ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- optimiser
int gdn_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, long long step, float grad_scale, void* stream) {
    GDN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 1, "adam_flat: bad argument");
    GDN_CHECK_ARG(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f,
                  "adam_flat: bad hyper-parameters (step=%lld)", step);
    prof_enter((cudaStream_t)stream, "@adam_flat");
    return launch_adam_flat(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, nullptr,
                            grad_scale, (cudaStream_t)stream);
}

int gdn_adam_flat_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, const long long* step_dev, float grad_scale,
                      void* stream) {
    GDN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_dev && n >= 1, "adam_flat_dev: bad argument");
    GDN_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f,
                  "adam_flat_dev: bad hyper-parameters");
    prof_enter((cudaStream_t)stream, "@adam_flat");
    return launch_adam_flat(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, 0, step_dev,
                            grad_scale, (cudaStream_t)stream);
}

int gdn_nvls_adam(const float* params_local, float* params_mc, const float* grads_mc, float* exp_avg, float* exp_avg_sq,
                  long long lo, long long count, float lr, float beta1, float beta2, float eps, float weight_decay,
                  long long step, float grad_scale, void* stream) {
    GDN_CHECK_ARG(params_local && params_mc && grads_mc && exp_avg && exp_avg_sq, "nvls_adam: NULL argument");
    GDN_CHECK_ARG(lo >= 0 && count >= 4 && (lo % 4) == 0 && (count % 4) == 0, "nvls_adam: slice [%lld, +%lld) must be 16-byte aligned", lo, count);
    GDN_CHECK_ARG(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f,
                  "nvls_adam: bad hyper-parameters (step=%lld)", step);
    prof_enter((cudaStream_t)stream, "@nvls_adam");
    return launch_nvls_adam(params_local, params_mc, grads_mc, exp_avg, exp_avg_sq, lo, count, lr, beta1, beta2, eps,
                            weight_decay, step, grad_scale, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- metrics
int gdn_f1_sweep(const double* sorted_scores, const float* labels_sorted, int T, const int* k_pred, const int* k_thr,
                 int S, double* fmeas, double* thresholds, void* stream) {
    GDN_CHECK_ARG(sorted_scores && labels_sorted && k_pred && k_thr && fmeas && thresholds, "f1_sweep: NULL argument");
    GDN_CHECK_ARG(T >= 1 && S >= 1 && S <= 65535, "f1_sweep: bad shape T=%d steps=%d", T, S);
    prof_enter((cudaStream_t)stream, "@f1_sweep");
    return launch_f1_sweep(sorted_scores, labels_sorted, T, k_pred, k_thr, S, fmeas, thresholds, (cudaStream_t)stream);
}

int gdn_binary_counts(const double* scores, const float* labels, int T, double threshold, unsigned long long* counts,
                      void* stream) {
    GDN_CHECK_ARG(scores && labels && counts && T >= 1, "binary_counts: bad argument");
    prof_enter((cudaStream_t)stream, "@binary_counts");
    return launch_binary_counts(scores, labels, T, threshold, counts, (cudaStream_t)stream);
}

int gdn_auc_ranksum(const double* sorted_scores, const float* labels_sorted, int T, double* ranksum,
                    unsigned long long* npos, void* stream) {
    GDN_CHECK_ARG(sorted_scores && labels_sorted && ranksum && npos && T >= 1, "auc_ranksum: bad argument");
    prof_enter((cudaStream_t)stream, "@auc_ranksum");
    return launch_auc_ranksum(sorted_scores, labels_sorted, T, ranksum, npos, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- data feed
int gdn_window_batch(const float* series, const float* labels, int N, int T, int W, const int* win_end, int B,
                     float* x, float* y, float* lab, int* err, void* stream) {
    GDN_CHECK_ARG(series && win_end && x && y && err, "window_batch: NULL argument");
    GDN_CHECK_ARG(N >= 1 && W >= 1 && T > W && B >= 1, "window_batch: bad shape N=%d T=%d W=%d B=%d", N, T, W, B);
    GDN_CHECK_ARG((long long)B * N * W < (1LL << 31), "window_batch: B*N*W must be < 2^31");
    prof_enter((cudaStream_t)stream, "@window_batch");
    return launch_window_batch(series, labels, N, T, W, win_end, B, x, y, lab, err, (cudaStream_t)stream);
}

}  // extern "C"
