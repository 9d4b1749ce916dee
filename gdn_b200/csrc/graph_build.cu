// graph_build.cu -- learned-graph builder: cosine Gram of the sensor embeddings fused with
// the row-wise top-k (replaces models/GDN.py:143-159).
//
//   cos[i,j] = fl( fl(v_i.v_j) / fl(|v_i| |v_j|) ),  idx[i,:] = topk(cos[i,:], K) (descending)
//
// The [N,N] matrix is never written: a CTA owns 64 rows, sweeps the columns in tiles of 64,
// and streams every tile into per-row sorted candidate lists kept in shared memory.
// Ranking rule (deterministic): larger cosine first; equal cosines -> lower column first
// (torch.topk breaks exact ties arbitrarily, see SURVEY.md section 7 hard part 1).
//
// Two Gram engines feed the same selection code:
//   * exact fp32 FMA (this file, all N): each dot product is one fmaf chain over d in
//     ascending order -- bit-reproducible and independent of the tiling;
//   * tcgen05 split-precision tensor-core Gram for large N (gram_tc.cu), whose candidates
//     are re-scored with the same fp32 chain before ranking.
// Also emits nbr[N][K+1]: the neighbour list GraphLayer uses after
// remove_self_loops/add_self_loops (models/graph_layer.py:61-63).
#include <stdlib.h>
#include "common.cuh"
#include "launchers.h"

namespace gdn {

#define GB_TI 64
#define GB_TJ 64
#define GB_DC 32
#define GB_MAXK 256

__global__ void k_row_norms(const float* __restrict__ V, int N, int D, float* __restrict__ nrm) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < N; i += nwarps) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = V[(size_t)i * D + d];
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) nrm[i] = sqrtf(s);
    }
}

// Insert candidate (v, j) into the descending list (vals, idxs) of current length *cnt
// (capacity K).  Executed by a full warp; entries with equal value keep arrival order.
__device__ __forceinline__ void list_insert(float* vals, int* idxs, int* cnt_p, int K, float v, int j, int lane) {
    const int cnt = *cnt_p;
    int ge = 0;
    for (int e = lane; e < cnt; e += 32) ge += (vals[e] >= v) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ge += __shfl_xor_sync(0xffffffffu, ge, o);
    const int pos = ge;                       // first slot whose value is < v
    if (pos >= K) return;                     // (cannot happen when the caller pre-filters)
    const int last = (cnt < K ? cnt : K - 1); // slot that receives the shifted tail end
    // shift [pos, last-1] -> [pos+1, last]; read everything first, then write
    float tv[GB_MAXK / 32];
    int ti[GB_MAXK / 32];
#pragma unroll
    for (int q = 0; q < GB_MAXK / 32; ++q) {
        const int e = pos + lane + 32 * q;
        if (e < last) { tv[q] = vals[e]; ti[q] = idxs[e]; }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < GB_MAXK / 32; ++q) {
        const int e = pos + lane + 32 * q;
        if (e < last) { vals[e + 1] = tv[q]; idxs[e + 1] = ti[q]; }
    }
    if (lane == 0) {
        vals[pos] = v;
        idxs[pos] = j;
        *cnt_p = cnt < K ? cnt + 1 : K;
    }
    __syncwarp();
}

// dynamic smem: lists  vals[GB_TI][K] (float), idxs[GB_TI][K] (int)
__global__ void __launch_bounds__(256)
k_gram_topk(const float* __restrict__ V, const float* __restrict__ nrm, int N, int D, int K, int blk0,
            int64_t* __restrict__ idx_out, int32_t* __restrict__ nbr_out, const int* __restrict__ block_flags,
            float* __restrict__ kth_out) {
    const int blk = blockIdx.x + blk0;                                     // 64-row block (blk0: first block of the row range)
    if (block_flags != nullptr && block_flags[blk] == 0) return;           // fix-up mode: flagged blocks only
    __shared__ float As[GB_TI][GB_DC + 1];
    __shared__ float Bs[GB_TJ][GB_DC + 1];
    __shared__ float Cs[GB_TI][GB_TJ + 1];
    __shared__ int cnts[GB_TI];
    extern __shared__ float lists[];
    float* lvals = lists;
    int* lidxs = reinterpret_cast<int*>(lists + (size_t)GB_TI * K);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int i0 = blk * GB_TI;
    if (tid < GB_TI) cnts[tid] = 0;
    __syncthreads();

    for (int j0 = 0; j0 < N; j0 += GB_TJ) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int d0 = 0; d0 < D; d0 += GB_DC) {
            // stage 64 x 32 slices of the row block and the column block
            for (int e = tid; e < GB_TI * GB_DC; e += 256) {
                const int r = e / GB_DC, c = e % GB_DC;
                const int gi = i0 + r, gj = j0 + r, gd = d0 + c;
                As[r][c] = (gi < N && gd < D) ? V[(size_t)gi * D + gd] : 0.f;
                Bs[r][c] = (gj < N && gd < D) ? V[(size_t)gj * D + gd] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int c = 0; c < GB_DC; ++c) {
                float av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) av[a] = As[ty * 4 + a][c];
#pragma unroll
                for (int b = 0; b < 4; ++b) bv[b] = Bs[tx * 4 + b][c];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gi = i0 + ty * 4 + a;
            const float ni = gi < N ? nrm[gi] : 1.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int gj = j0 + tx * 4 + b;
                const float nj = gj < N ? nrm[gj] : 1.f;
                Cs[ty * 4 + a][tx * 4 + b] = acc[a][b] / (ni * nj);
            }
        }
        __syncthreads();
        // stream the tile into the per-row candidate lists: warp `wid` owns rows wid*8..+7
        for (int rr = 0; rr < 8; ++rr) {
            const int r = wid * 8 + rr;
            if (i0 + r >= N) break;
            float* vals = lvals + (size_t)r * K;
            int* idxs = lidxs + (size_t)r * K;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c = half * 32 + lane;
                const int j = j0 + c;
                const float v = Cs[r][c];
                const int cnt = cnts[r];
                const float thr = cnt >= K ? vals[K - 1] : -INFINITY;
                const bool pass = (j < N) && (cnt < K ? (v == v) : (v > thr));
                unsigned mask = __ballot_sync(0xffffffffu, pass);
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float cv = __shfl_sync(0xffffffffu, v, l);
                    const int cj = __shfl_sync(0xffffffffu, j, l);
                    const int cn = cnts[r];
                    if (cn < K || cv > vals[K - 1]) list_insert(vals, idxs, &cnts[r], K, cv, cj, lane);
                }
            }
        }
        __syncthreads();
    }
    // emit idx (int64) and the self-loop-fixed neighbour list
    for (int rr = 0; rr < 8; ++rr) {
        const int r = wid * 8 + rr;
        const int gi = i0 + r;
        if (gi >= N) break;
        const int* idxs = lidxs + (size_t)r * K;
        const int cnt = cnts[r];
        if (idx_out != nullptr)
            for (int k = lane; k < K; k += 32) idx_out[(size_t)gi * K + k] = k < cnt ? (int64_t)idxs[k] : (int64_t)gi;
        if (kth_out != nullptr && lane == 0) kth_out[gi] = cnt >= K ? lvals[(size_t)r * K + K - 1] : -INFINITY;
        if (nbr_out != nullptr && lane == 0) {
            int32_t* nb = nbr_out + (size_t)gi * (K + 1);
            int o = 0, pos = -1;
            for (int k = 0; k < K; ++k) {
                const int j = k < cnt ? idxs[k] : gi;
                if (j != gi) nb[o++] = j;
                else if (pos < 0) pos = k;
            }
            nb[o++] = gi;
            // padding: any negative value ends the list; the last slot also records where the row itself sat in
            // the top-k (-2 - position), so that idx can be rebuilt from this table alone (data-parallel exchange)
            for (; o < K + 1; ++o) nb[o] = (o == K && pos >= 0) ? -2 - pos : -1;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Small graphs (N <= GR_MAXN; the reference's own data sets have 27..127 sensors): one WARP per row.
// k_gram_topk gives such a graph two or three CTAs, each threading every candidate of 64 rows through a
// serial sorted-list insert (0.33 ms at 127 sensors, 60 % of that train step).  Here a CTA owns eight rows,
// stages 64-column slices of V in shared memory, and each warp
//   1. computes its row's N cosines -- the same fmaf chain over d = 0..D-1 and the same acc / (n_i n_j) as the
//      tile kernel, norms summed in k_row_norms' order, so every value is bit-identical to that engine's;
//   2. finds the K-th largest by a 32-step bitwise search over order-preserving integer keys;
//   3. compacts the entries above it (plus the earliest ties) and ranks them among themselves:
//      descending value, equal values by ascending column -- the order list_insert produces.
// dynamic smem: Bs[GR_TJ][DP] | As[GR_ROWS][D] | nb[GR_TJ] | keys[GR_ROWS][NP] | sel_key, sel_j [GR_ROWS][K]
// ---------------------------------------------------------------------------------------
#define GR_ROWS 8
#define GR_TJ 64
#define GR_MAXN 2048

__device__ __forceinline__ unsigned cos_key(float v) {           // NaN -> 0 (never selected); -0 == +0
    if (!(v == v)) return 0u;
    const unsigned u = __float_as_uint(v + 0.f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_cos(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(GR_ROWS * 32)
k_gram_rows(const float* __restrict__ V, int N, int D, int K, int row0, int row1,
            int64_t* __restrict__ idx_out, int32_t* __restrict__ nbr_out, float* __restrict__ kth_out) {
    extern __shared__ float gr_sm[];
    const int DP = D | 1, NP = (N + 31) & ~31;
    float* Bs = gr_sm;
    float* As = Bs + (size_t)GR_TJ * DP;
    float* nb = As + (size_t)GR_ROWS * D;
    unsigned* keys = reinterpret_cast<unsigned*>(nb + GR_TJ);
    unsigned* selk = keys + (size_t)GR_ROWS * NP;
    int* selj = reinterpret_cast<int*>(selk + (size_t)GR_ROWS * K);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gi = row0 + blockIdx.x * GR_ROWS + wid;
    const bool row_ok = gi < row1;

    for (int e = tid; e < GR_ROWS * D; e += GR_ROWS * 32) {
        const int r = row0 + blockIdx.x * GR_ROWS + e / D;
        As[e] = r < row1 ? V[(size_t)r * D + e % D] : 0.f;
    }
    __syncthreads();
    const float* arow = As + (size_t)wid * D;
    float ni;
    {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(arow[d], arow[d], s);
        ni = sqrtf(warp_sum(s));
    }
    unsigned* krow = keys + (size_t)wid * NP;
    for (int j0 = 0; j0 < N; j0 += GR_TJ) {
        for (int e = tid; e < GR_TJ * D; e += GR_ROWS * 32) {
            const int r = e / D, c = e % D;
            Bs[(size_t)r * DP + c] = j0 + r < N ? V[(size_t)(j0 + r) * D + c] : 0.f;
        }
        __syncthreads();
        for (int r = wid; r < GR_TJ; r += GR_ROWS) {               // the slice's norms, k_row_norms' order
            float s = 0.f;
            for (int d = lane; d < D; d += 32) { const float v = Bs[(size_t)r * DP + d]; s = fmaf(v, v, s); }
            s = warp_sum(s);
            if (lane == 0) nb[r] = sqrtf(s);
        }
        __syncthreads();
        if (row_ok) {
            const float* b0 = Bs + (size_t)lane * DP;
            const float* b1 = Bs + (size_t)(lane + 32) * DP;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
            for (int d = 0; d < D; ++d) {
                const float a = arow[d];
                acc0 = fmaf(a, b0[d], acc0);
                acc1 = fmaf(a, b1[d], acc1);
            }
            const int ja = j0 + lane, jb = j0 + lane + 32;
            if (ja < NP) krow[ja] = ja < N ? cos_key(acc0 / (ni * nb[lane])) : 0u;
            if (jb < NP) krow[jb] = jb < N ? cos_key(acc1 / (ni * nb[lane + 32])) : 0u;
        }
        __syncthreads();
    }
    if (!row_ok) return;
    // ---- selection (warp-local from here on) ----
    const int per = NP >> 5;
    int nvalid = 0;
    for (int q = 0; q < per; ++q) nvalid += krow[q * 32 + lane] != 0u ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    unsigned T = 1u;                                               // fewer than K valid entries: take them all
    int cnt = nvalid, need = 0;
    if (nvalid >= K) {
        T = 0u;
        for (int bit = 31; bit >= 0; --bit) {                      // largest T with #{key >= T} >= K
            const unsigned c = T | (1u << bit);
            int n = 0;
            for (int q = 0; q < per; ++q) n += krow[q * 32 + lane] >= c ? 1 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
            if (n >= K) T = c;
        }
        int gt = 0;
        for (int q = 0; q < per; ++q) gt += krow[q * 32 + lane] > T ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gt += __shfl_xor_sync(0xffffffffu, gt, o);
        need = K - gt;                                             // ties at the threshold: the earliest columns
        cnt = K;
    }
    unsigned* sk = selk + (size_t)wid * K;
    int* sj = selj + (size_t)wid * K;
    {
        const unsigned lt = (1u << lane) - 1u;
        int run = 0, eq_run = 0;
        for (int q = 0; q < per; ++q) {
            const unsigned k = krow[q * 32 + lane];
            const bool eq = nvalid >= K && k == T;
            const unsigned eqm = __ballot_sync(0xffffffffu, eq);
            const bool take = (k > T) || (eq && eq_run + __popc(eqm & lt) < need) || (nvalid < K && k != 0u);
            const unsigned tm = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int pos = run + __popc(tm & lt);
                sk[pos] = k;
                sj[pos] = q * 32 + lane;
            }
            run += __popc(tm);
            eq_run += __popc(eqm);
        }
    }
    __syncwarp();
    // rank inside the selection; a lane holds entries lane, lane + 32, ... (K <= GB_MAXK)
    int rk[GB_MAXK / 32], rj[GB_MAXK / 32];
    int self_pos = -1;
#pragma unroll
    for (int m = 0; m < GB_MAXK / 32; ++m) {
        const int e = lane + 32 * m;
        rk[m] = -1;
        rj[m] = 0;
        if (e < cnt) {
            const unsigned k = sk[e];
            const int j = sj[e];
            int r = 0;
            for (int f = 0; f < cnt; ++f) {
                const unsigned kf = sk[f];
                r += (kf > k || (kf == k && sj[f] < j)) ? 1 : 0;
            }
            rk[m] = r;
            rj[m] = j;
            if (j == gi) self_pos = r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) self_pos = max(self_pos, __shfl_xor_sync(0xffffffffu, self_pos, o));
    // k_gram_topk's conventions: idx slots past the list hold the row itself; nbr = the list without the row,
    // then the row, then padding whose last slot records where the row sat (-2 - position)
    const int pos = self_pos >= 0 ? self_pos : (cnt < K ? cnt : -1);
    const int nn = cnt - (self_pos >= 0 ? 1 : 0);
#pragma unroll
    for (int m = 0; m < GB_MAXK / 32; ++m) {
        if (rk[m] < 0) continue;
        if (idx_out != nullptr) idx_out[(size_t)gi * K + rk[m]] = (int64_t)rj[m];
        if (nbr_out != nullptr && rj[m] != gi)
            nbr_out[(size_t)gi * (K + 1) + rk[m] - ((self_pos >= 0 && rk[m] > self_pos) ? 1 : 0)] = rj[m];
    }
    if (idx_out != nullptr)
        for (int k = cnt + lane; k < K; k += 32) idx_out[(size_t)gi * K + k] = (int64_t)gi;
    if (nbr_out != nullptr)
        for (int o = nn + lane; o < K + 1; o += 32)
            nbr_out[(size_t)gi * (K + 1) + o] = o == nn ? gi : ((o == K && pos >= 0) ? -2 - pos : -1);
    if (kth_out != nullptr && lane == 0) kth_out[gi] = cnt >= K && nvalid >= K ? key_cos(T) : -INFINITY;
}

static size_t gram_rows_smem(int N, int D, int K) {
    const size_t DP = (size_t)(D | 1), NP = (size_t)((N + 31) & ~31);
    return (GR_TJ * DP + (size_t)GR_ROWS * D + GR_TJ + GR_ROWS * NP + 2 * (size_t)GR_ROWS * K) * sizeof(float);
}

// gram_tc.cu
bool gram_tc_supported(int N, int D, int K);
size_t gram_tc_ws_bytes(int N, int D, int K);
int launch_gram_tc(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr, void* ws,
                   cudaStream_t st, float* kth, float margin, float** nrm_out, int** flags_out);

size_t graph_build_ws_bytes(int N, int D, int K) {
    size_t b = align_up((size_t)N * sizeof(float), 256);
    if (gram_tc_supported(N, D, K)) {
        const size_t t = gram_tc_ws_bytes(N, D, K);
        if (t > b) b = t;
    }
    return b;
}

// rows [row0, row1) of the graph only (row0 a multiple of 128, row1 a multiple of 128 or N): the row-sharded
// build of the data-parallel trainer; rows outside the range are not touched
int launch_graph_build(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr, void* ws,
                       size_t ws_bytes, int use_tc, float* kth, float margin, cudaStream_t st) {
    GDN_CHECK_ARG(K <= GB_MAXK, "topk K=%d unsupported (max %d)", K, GB_MAXK);
    GDN_CHECK_ARG(0 <= row0 && row0 < row1 && row1 <= N && row0 % 128 == 0 && (row1 % 128 == 0 || row1 == N),
                  "graph_build: row range [%d, %d) must be non-empty, inside [0, %d) and aligned to 128", row0, row1, N);
    GDN_CHECK_ARG(ws != nullptr && ws_bytes >= graph_build_ws_bytes(N, D, K), "graph_build: workspace too small");
    const bool tc_ok = gram_tc_supported(N, D, K);
    GDN_CHECK_ARG(use_tc <= 0 || tc_ok,
                  "graph_build: the tcgen05 engine needs N >= 1024, dim 64 or 128, topk <= 72 (N=%d D=%d K=%d)", N, D, K);
    const size_t smem = (size_t)GB_TI * K * (sizeof(float) + sizeof(int));
    {
        cudaError_t e = ensure_dyn_smem(k_gram_topk, smem);
        if (e != cudaSuccess) return cuda_fail(e, "smem attr k_gram_topk");
    }
    if (use_tc > 0 || (use_tc < 0 && tc_ok)) {
        float* nrm = nullptr;
        int* flags = nullptr;
        if (int rc = launch_gram_tc(V, N, D, K, row0, row1, idx, nbr, ws, st, kth, margin, &nrm, &flags)) return rc;
        // exact fix-up of the (normally zero) 64-row blocks whose candidate window was ambiguous
        k_gram_topk<<<ceil_div(row1, GB_TI) - row0 / GB_TI, 256, smem, st>>>(V, nrm, N, D, K, row0 / GB_TI, idx, nbr, flags, kth);
        GDN_CHECK_LAUNCH("k_gram_topk_fixup");
        return 0;
    }
    const char* rows_env = getenv("GDN_GRAM_ROWS");                  // diagnostic: 0 = always the tile kernel
    if (N <= GR_MAXN && gram_rows_smem(N, D, K) <= 160 * 1024 && !(rows_env && rows_env[0] == '0')) {
        // small graph: a warp per row, one launch
        const size_t rs = gram_rows_smem(N, D, K);
        cudaError_t e = ensure_dyn_smem(k_gram_rows, rs);
        if (e != cudaSuccess) return cuda_fail(e, "smem attr k_gram_rows");
        k_gram_rows<<<ceil_div(row1 - row0, GR_ROWS), GR_ROWS * 32, rs, st>>>(V, N, D, K, row0, row1, idx, nbr, kth);
        GDN_CHECK_LAUNCH("k_gram_rows");
        return 0;
    }
    float* nrm = (float*)ws;
    int g = ceil_div(N, 8);
    if (g > 8 * num_sms()) g = 8 * num_sms();
    k_row_norms<<<g, 256, 0, st>>>(V, N, D, nrm);
    GDN_CHECK_LAUNCH("k_row_norms");
    k_gram_topk<<<ceil_div(row1, GB_TI) - row0 / GB_TI, 256, smem, st>>>(V, nrm, N, D, K, row0 / GB_TI, idx, nbr, nullptr, kth);
    GDN_CHECK_LAUNCH("k_gram_topk");
    return 0;
}

}  // namespace gdn
