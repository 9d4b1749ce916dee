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
