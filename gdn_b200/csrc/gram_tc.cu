// gram_tc.cu -- tensor-core engine of the learned-graph builder for large sensor counts
// (models/GDN.py:143-159).
//
// The N x N cosine Gram is a dense contraction, so at large N it belongs on the 5th-generation
// tensor cores.  fp32 ranking accuracy is kept by a split-precision product followed by an exact
// re-score:
//   1. k_normalise_split: vn = v / |v| (fp32), split into two bf16 matrices hi + lo = vn
//      (|vn - hi - lo| <= 2^-18 |vn|).
//   2. k_gram_tc (this file's tcgen05 kernel): one CTA owns 128 rows and sweeps its share of the columns
//      in tiles of 128 (grid = row blocks x column splits).  Operand tiles arrive by TMA
//      (cp.async.bulk.tensor, SWIZZLE_128B, K-major; one pipeline stage per (tile, k-block)); one elected
//      thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) for the three products
//      hi.hi + hi.lo + lo.hi into two TMEM accumulators; two epilogue warpgroups (one per accumulator) pull
//      their tiles out of TMEM with software-pipelined tcgen05.ld (thread <-> row), filter them against the
//      row's threshold and append what passes to the (row, segment) candidate buffer in global memory
//      (L2-resident), compacting it by bit-bisection when it could overflow.  The [N, N] matrix is never
//      written.  |approx - exact| <= eps = 3e-5 (3 * 2^-18 from the dropped lo.lo term and the split
//      residuals, plus fp32 accumulation).
//   3. k_rescore: merges the row's candidate segments (the K + 8 ... K + 16 best approximate values, by
//      bit-bisection on order-preserving keys with an early exit), re-scores them with the SAME fp32 fmaf
//      chain the exact engine uses (graph_build.cu; candidate rows staged through shared memory with
//      coalesced loads) and ranks them (value desc, index asc) -> bit-identical output.  A row whose spare
//      candidates all sit within 2*eps of the K-th approximate value could have lost a true neighbour, and
//      a row with fewer than K + 8 candidates had a stale warm-start hint: its 64-row block is flagged and
//      recomputed by the exact engine (never observed on non-degenerate embeddings; exercised by the tests
//      with duplicates and absurd hints).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "launchers.h"

namespace gdn {

constexpr int TC_BM = 128;        // rows per CTA = UMMA M
constexpr int TC_BN = 128;        // columns per tile = UMMA N (N < 128 leaves the MMA bound by its A-operand reads)
constexpr int TC_BK = 64;         // bf16 per k-block: 128 bytes = one SWIZZLE_128B atom row
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 384;
constexpr int TC_MAXSPLIT = 8;     // column splits per row block (x 2 warpgroups = candidate segments per row)
constexpr int TC_SLACK = 8;
constexpr int TC_MAXL = 80;
constexpr int TC_MAXC = 256;      // buffer capacity (8 entries per lane in a compaction)
constexpr float TC_EPS = 3e-5f;
constexpr uint32_t TC_SPIN = 1u << 22;   // bounded mbarrier spins: a protocol bug must not hang the GPU

// ---------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    for (uint32_t it = 0; it < TC_SPIN; ++it)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, code);
    return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for the outstanding tcgen05.ld; the registers are in/out operands so that no use of them can be
// scheduled above the wait when the load was issued earlier (software pipelining)
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4,
// LBO = 0 (a single swizzle atom along K), SBO = 1024 bytes (8 rows x 128 bytes), version 1.
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = TC_BN, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

// ---------------------------------------------------------------------------------------
// 1. normalise + split
// ---------------------------------------------------------------------------------------
__global__ void k_normalise_split(const float* __restrict__ V, int N, int D, float* __restrict__ nrm,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
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
        const float n = sqrtf(s);                   // same value k_row_norms produces
        if (lane == 0) nrm[i] = n;
        for (int d = lane; d < D; d += 32) {
            const float vn = V[(size_t)i * D + d] / n;
            const __nv_bfloat16 h = __float2bfloat16_rn(vn);
            const __nv_bfloat16 l = __float2bfloat16_rn(vn - __bfloat162float(h));
            hi[(size_t)i * D + d] = h;
            lo[(size_t)i * D + d] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------
// 2. tcgen05 Gram + per-row candidate selection
// ---------------------------------------------------------------------------------------
// grid (row blocks, column splits); 12 warps: 0 = TMA, 1 = MMA, 2 = TMEM alloc, 4-7 / 8-11 = the two
// epilogue warpgroups.  Warpgroup g consumes the tiles whose accumulator lives in TMEM buffer g (every
// other tile), so the two filter streams overlap; each (column split, warpgroup) pair owns its own
// candidate segment of every row: segment = 2 * blockIdx.y + g, merged by k_rescore.
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gram_tc(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
          int N, int KB, int L, int C, int tiles_per_split, int rb0, float* __restrict__ bufv, int* __restrict__ bufj,
          int* __restrict__ rowcnt, const float* __restrict__ hint, float margin, int* __restrict__ err, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SWIZZLE_128B atoms: 1024-byte aligned
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t A_BYTES = (uint32_t)KB * TC_BM * 128;                 // one of hi / lo
    constexpr uint32_t B_BYTES = TC_BN * 128;                            // one of hi / lo, one k-block = half a stage
    const uint32_t sA = base;                                            // [2][KB][128 x 128 B]
    const uint32_t sB = sA + 2 * A_BYTES;                                // [STAGES][2][TC_BN x 128 B]; stage <-> (tile, kb)
    const uint32_t off_stage = 2 * A_BYTES + TC_STAGES * 2 * B_BYTES;
    float* p_vals = reinterpret_cast<float*>(gbase + off_stage);         // [2][16][128] filter staging per warpgroup
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_vals + 2 * 16 * TC_BM);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_afull = bar0, b_full = bar0 + 8, b_empty = b_full + 8 * TC_STAGES,
                   b_tfull = b_empty + 8 * TC_STAGES, b_tempty = b_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = (blockIdx.x + rb0) * TC_BM;                          // global first row (rb0: first row block of the range)
    const int l0 = blockIdx.x * TC_BM;                                   // ... and its position in the candidate buffers
    const int S = 2 * (int)gridDim.y;                                    // candidate segments per row
    const int tile0 = blockIdx.y * tiles_per_split;                      // this CTA sweeps tiles [tile0, tile0 + ntiles)
    const int ntiles = min(tiles_per_split, (N + TC_BN - 1) / TC_BN - tile0);

    if (threadIdx.x == 0) {
        mbar_init(b_afull, 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(b_tfull + 8 * a, 1); mbar_init(b_tempty + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(2u * TC_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        mbar_expect_tx(b_afull, 2 * A_BYTES);
        for (int h = 0; h < 2; ++h)
            for (int kb = 0; kb < KB; ++kb)
                for (int half = 0; half < 2; ++half)
                    tma_load_2d(sA + h * A_BYTES + kb * (TC_BM * 128) + half * (64 * 128), h ? &tm_lo : &tm_hi, b_afull,
                                kb * TC_BK, m0 + half * 64);
        for (int it = 0; it < ntiles * KB; ++it) {
            const int t = it / KB, kb = it - t * KB;
            const int s = it % TC_STAGES;
            const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
            if (!mbar_wait(b_empty + 8 * s, ph ^ 1u, err, 1)) break;
            mbar_expect_tx(b_full + 8 * s, 2 * B_BYTES);
            for (int h = 0; h < 2; ++h)
                for (int half = 0; half < TC_BN / 64; ++half)
                    tma_load_2d(sB + (s * 2 + h) * B_BYTES + half * (64 * 128), h ? &tm_lo : &tm_hi, b_full + 8 * s,
                                kb * TC_BK, (tile0 + t) * TC_BN + half * 64);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer (one thread) =====
        bool ok = mbar_wait(b_afull, 0, err, 2);
        for (int t = 0; ok && t < ntiles; ++t) {
            const int acc = t & 1;
            const uint32_t aph = (uint32_t)(t >> 1) & 1u;
            if (!mbar_wait(b_tempty + 8 * acc, aph ^ 1u, err, 3)) break;
            uint32_t accumulate = 0;
            for (int kb = 0; kb < KB; ++kb) {
                const int it = t * KB + kb;
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                if (!mbar_wait(b_full + 8 * s, ph, err, 4)) { ok = false; break; }
                tc_fence_after();
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {                  // hi.hi, hi.lo, lo.hi
                    if ((dbg & 2) && prod > 0) break;
                    const int ah = prod == 2 ? 1 : 0, bh = prod == 1 ? 1 : 0;
                    const uint32_t a0 = sA + ah * A_BYTES + kb * (TC_BM * 128);
                    const uint32_t b0 = sB + (s * 2 + bh) * B_BYTES;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        tc_mma_f16(tmem_base + acc * TC_BN, sw128_desc(a0 + k * 32), sw128_desc(b0 + k * 32), TC_IDESC,
                                   accumulate);
                        accumulate = 1;
                    }
                }
                tc_commit(b_empty + 8 * s);      // smem stage reusable once these MMAs have read it
            }
            if (!ok) break;
            tc_commit(b_tfull + 8 * acc);        // accumulator ready for the epilogue
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread <-> row =====
        // Candidate selection, thread <-> row for the streaming part:
        //   * a thread appends every value above its row's threshold to the row's segment buffer (capacity
        //     C = 256 entries, global memory, L2-resident; the append is two fire-and-forget stores);
        //   * when a buffer could overflow within the next tile (128 columns) the warp compacts it
        //     cooperatively: the L-th largest key by bit-bisection, keep the L best at slots 0..L-1, and raise
        //     the threshold to the L-th value.  The threshold is only refreshed at compactions, so a cold row is
        //     compacted ~1 + ln(N / C) times per sweep instead of paying a minimum search per admitted element;
        //     a warm-started row normally never compacts before the final pass.
        //   * the row's segments (one per column split and warpgroup) are merged by k_rescore.
        const int wg = (warp - 4) >> 2;              // epilogue warpgroup = TMEM accumulator buffer it drains
        const int seg = 2 * (int)blockIdx.y + wg;    // candidate segment of this (column split, warpgroup)
        const int wrow0 = (warp & 3) * 32;           // first row of this warp inside the CTA (= its TMEM lanes)
        const int row = wrow0 + lane;
        float* bv = bufv + ((size_t)(l0 + row) * S + seg) * C;
        int* bj = bufj + ((size_t)(l0 + row) * S + seg) * C;
        float* pv = p_vals + wg * 16 * TC_BM + row;  // staging slot q of row r at [q * 128 + r]
        int cnt = 0;
        // Warm start: the caller may pass last step's K-th cosine per row.  The embedding moves by one
        // optimiser step between graph builds, so (hint - margin) is a valid admission threshold for
        // nearly every row and the sweep appends ~L + margin*density entries instead of ~L ln(N/L): no
        // compaction before the final one.  A row that ends with fewer than L entries had a stale hint:
        // its block is flagged and recomputed by the exact engine.
        float thr = -INFINITY;
        if (hint != nullptr && m0 + row < N) {
            const float hv = hint[m0 + row];
            if (hv == hv && hv > -2.f) thr = hv - margin;
        }
        constexpr int NE = TC_MAXC / 32;
        // order-preserving float <-> unsigned key
        auto to_key = [](float f) -> unsigned {
            const unsigned b = __float_as_uint(f);
            return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
        };
        auto from_key = [](unsigned k) -> float {
            return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
        };
        // Cooperative compaction of the rows in `todo`: find the L-th largest key by bisection on its
        // bits (32 warp-wide counts), keep the L best entries (ties by position) in slots 0..L-1.
        auto compact = [&](unsigned todo) {
            while (todo) {
                const int rr = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cnt_r = __shfl_sync(0xffffffffu, cnt, rr);
                float* gv = bufv + ((size_t)(l0 + wrow0 + rr) * S + seg) * C;
                int* gj = bufj + ((size_t)(l0 + wrow0 + rr) * S + seg) * C;
                unsigned key[NE];
                int ej[NE];
#pragma unroll
                for (int t = 0; t < NE; ++t) {
                    const int e = lane + 32 * t;
                    const bool valid = e < cnt_r;
                    key[t] = valid ? to_key(gv[e]) : 0u;            // key 0 (= -NaN pattern) sorts last
                    ej[t] = valid ? gj[e] : -1;
                }
                if (cnt_r > L) {
                    unsigned T = 0u;
#pragma unroll 1
                    for (int bit = 30; bit >= 0; bit -= 2) {          // two key bits per step
                        const unsigned c1 = T | (1u << bit), c2 = T | (2u << bit), c3 = T | (3u << bit);
                        unsigned c = 0;                                   // three counts (<= 256 each), 10 bits apiece
#pragma unroll
                        for (int t = 0; t < NE; ++t)
                            c += (key[t] >= c1 ? 1u : 0u) + (key[t] >= c2 ? 1u << 10 : 0u) + (key[t] >= c3 ? 1u << 20 : 0u);
                        c = __reduce_add_sync(0xffffffffu, c);
                        const int n1 = c & 1023, n2 = (c >> 10) & 1023, n3 = c >> 20;
                        T = n3 >= L ? c3 : (n2 >= L ? c2 : (n1 >= L ? c1 : T));
                    }
                    int n_gt = 0;
#pragma unroll
                    for (int t = 0; t < NE; ++t) n_gt += (key[t] > T) ? 1 : 0;
                    n_gt = __reduce_add_sync(0xffffffffu, n_gt);
                    const int need_eq = L - n_gt;                    // >= 1 entries equal to T are kept
                    const unsigned lt = (1u << lane) - 1u;
                    int eq_before = 0, kept_before = 0;
#pragma unroll
                    for (int t = 0; t < NE; ++t) {
                        const bool is_eq = key[t] == T;
                        const unsigned eqm = __ballot_sync(0xffffffffu, is_eq);
                        const int eq_rank = eq_before + __popc(eqm & lt);
                        const bool keep = key[t] > T || (is_eq && eq_rank < need_eq);
                        const unsigned km = __ballot_sync(0xffffffffu, keep);
                        if (keep) {
                            const int slot = kept_before + __popc(km & lt);
                            gv[slot] = from_key(key[t]);
                            gj[slot] = ej[t];
                        }
                        eq_before += __popc(eqm);
                        kept_before += __popc(km);
                    }
                    if (lane == rr) { thr = from_key(T); cnt = L; }
                }
                __syncwarp();
            }
        };
        // One 16-column group of the row: build the pass mask, and only if some row of the warp passes
        // anything park the 16 values in shared memory (dynamic indexing) and append the passing ones.
        // Passing entries are rare (~L per row per sweep once the threshold is warm), so the cost of a
        // group is the 16 compares.
        auto filter = [&](const uint32_t (&r)[16], int cbase) {
            unsigned mask = 0u;
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) mask |= (__uint_as_float(r[cc]) > thr) ? (1u << cc) : 0u;
            const int lim = N - cbase;                               // columns >= lim are padding
            if (lim < 16) mask &= lim <= 0 ? 0u : ((1u << lim) - 1u);
            if (__any_sync(0xffffffffu, mask != 0u)) {
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) pv[cc * TC_BM] = __uint_as_float(r[cc]);
                while (mask) {
                    const int cc = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    bv[cnt] = pv[cc * TC_BM];
                    bj[cnt] = cbase + cc;
                    ++cnt;
                }
                __syncwarp();
            }
        };
        // The TMEM loads are software-pipelined one group ahead (two register sets); the loop body is
        // kept small on purpose: with one warp per scheduler an instruction-cache miss is fully exposed.
        bool ok = true;
        constexpr int NG = TC_BN / 16;
#pragma unroll 1
        for (int t = wg;; t += 2) {
            // compaction point: once per tile (the tile can append at most TC_BN), and once at the very end
            const bool last = t >= ntiles;
            const unsigned need = __ballot_sync(0xffffffffu, last ? cnt > L : cnt + TC_BN > C);
            if (need) compact(need);
            if (last) break;
            const uint32_t aph = (uint32_t)(t >> 1) & 1u;
            if (!mbar_wait(b_tfull + 8 * wg, aph, err, 5)) { ok = false; break; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + wg * TC_BN;
            const int j0 = (tile0 + t) * TC_BN;
            uint32_t ra[16], rb[16];
            tmem_ld16(taddr, ra);
#pragma unroll 1
            for (int grp = 0; grp < NG; grp += 2) {
                tmem_wait_ld16(ra);
                tmem_ld16(taddr + (grp + 1) * 16, rb);
                if (!(dbg & 1)) filter(ra, j0 + grp * 16);
                tmem_wait_ld16(rb);
                if (grp + 2 < NG) {
                    tmem_ld16(taddr + (grp + 2) * 16, ra);
                } else {                                             // whole tile read: TMEM stage is free again
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_tempty + 8 * wg);
                }
                if (!(dbg & 1)) filter(rb, j0 + (grp + 1) * 16);
            }
        }
        if (ok && m0 + row < N) rowcnt[(size_t)(l0 + row) * S + seg] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TC_BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// 3. exact re-score + ranking of the candidates (one warp per row)
// ---------------------------------------------------------------------------------------
constexpr int RS_WARPS = 2;
constexpr int RS_MAXSEL = 96;      // candidates re-scored per row: L <= n <= L + 8 <= 96
// order-preserving float <-> unsigned key
__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }

template <int D>
__global__ void __launch_bounds__(RS_WARPS * 32)
k_rescore(const float* __restrict__ V, const float* __restrict__ nrm, int row0, int row1, int K, int L, int C, int S,
          const float* __restrict__ cand_val, const int* __restrict__ cand_idx, const int* __restrict__ rowcnt,
          float* __restrict__ kth_out,
          int64_t* __restrict__ idx_out, int32_t* __restrict__ nbr_out, int* __restrict__ block_flags,
          const int* __restrict__ tc_err) {
    __shared__ unsigned long long s_key[RS_WARPS][RS_MAXSEL];  // ranking keys (exact cosine, index)
    __shared__ float s_av[RS_WARPS][RS_MAXSEL];             // approximate (tensor-core) values of the selection
    __shared__ int s_j[RS_WARPS][RS_MAXSEL];
    __shared__ int s_out[RS_WARPS][TC_MAXL];
    extern __shared__ __align__(16) float rs_smem[];        // per warp: tile [32][D/2+4] (first the pool keys/idx [2][S*L]), v_i [D]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int li = blockIdx.x * RS_WARPS + wid;               // row inside the range = row of the candidate buffers
    const int i = row0 + li;
    if (i >= row1) return;
    // a pipeline protocol error in k_gram_tc (bounded mbarrier wait expired; never observed) leaves candidate
    // lists incomplete: hand every block to the exact engine instead of trusting them
    if (*tc_err != 0 && lane == 0) block_flags[i / 64] = 1;
    // the candidate rows are staged HALF a row at a time (the fmaf chain simply continues): half the shared memory
    // per warp, twice the resident warps for a kernel whose 128-step dependent chain is latency-bound
    constexpr int DH = D / 2;
    constexpr int TS = DH + 4;                                   // tile row stride: conflict-free float4 row reads
    float* tile = rs_smem + (size_t)wid * (32 * TS + D);
    float* svi = tile + 32 * TS;
    unsigned* pk = reinterpret_cast<unsigned*>(tile);        // the pool is dead before the tile is first written
    int* pj = reinterpret_cast<int*>(pk + S * L);
    for (int d = 4 * lane; d < D; d += 128)
        *reinterpret_cast<float4*>(svi + d) = __ldg(reinterpret_cast<const float4*>(V + (size_t)i * D + d));
    const float ni = nrm[i];
    for (int k = lane; k < K; k += 32) s_out[wid][k] = i;     // placeholder if the row is short (it is flagged)
    // ---- merge the row's candidate segments: the L best approximate values of their union ----
    // (segment counts first, then one flat pass over all segments so that every load is independent)
    int my_cnt = lane < S ? min(rowcnt[(size_t)li * S + lane], L) : 0;    // S <= 2 * TC_MAXSPLIT <= 32
    int my_off = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, my_off, o);
        if (lane >= o) my_off += t;
    }
    const int total = __shfl_sync(0xffffffffu, my_off, 31);
    my_off -= my_cnt;                                                     // exclusive prefix
    for (int f0 = 0; f0 < S * L; f0 += 32) {
        const int f = f0 + lane;
        const int sg = min(f / L, S - 1), e = f - sg * L;
        const int c = __shfl_sync(0xffffffffu, my_cnt, sg), off = __shfl_sync(0xffffffffu, my_off, sg);
        if (e < c) {
            const size_t src = ((size_t)li * S + sg) * C + e;
            pk[off + e] = f2key(__ldg(cand_val + src));
            pj[off + e] = __ldg(cand_idx + src);
        }
    }
    __syncwarp();
    // Selection by approximate value: any T with L <= #{key >= T} <= Lmax = L + 8 will do (every extra
    // candidate is another 4*D bytes through L2, which is what bounds this kernel).  Bisection on the
    // key bits below the common prefix of the pool, two bits per step, stops at the first such T --
    // a handful of steps; only a tie cluster straddling the window runs to the last bit.
    const int Lmax = min(RS_MAXSEL, L + 8);
    int have = total;
    if (total <= Lmax) {
        for (int e = lane; e < total; e += 32) { s_av[wid][e] = key2f(pk[e]); s_j[wid][e] = pj[e]; }
    } else {
        unsigned kmx = 0u, kmn = 0xffffffffu;
        for (int e = lane; e < total; e += 32) { const unsigned k = pk[e]; kmx = max(kmx, k); kmn = min(kmn, k); }
        kmx = __reduce_max_sync(0xffffffffu, kmx);
        kmn = __reduce_min_sync(0xffffffffu, kmn);
        int bit = (31 - __clz((kmx ^ kmn) | 1u)) & ~1;                    // highest differing bit pair
        unsigned T = bit >= 30 ? 0u : (kmx >> (bit + 2)) << (bit + 2);   // common prefix
        int nT = total;                                                   // #{key >= T}
#pragma unroll 1
        for (; bit >= 0 && nT > Lmax; bit -= 2) {
            const unsigned c1 = T | (1u << bit), c2 = T | (2u << bit), c3 = T | (3u << bit);
            unsigned c = 0, d = 0;                                        // counts up to S*L (> 1023 possible): 16-bit fields
            for (int e = lane; e < total; e += 32) {
                const unsigned k = pk[e];
                c += (k >= c1 ? 1u : 0u) + (k >= c2 ? 1u << 16 : 0u);
                d += (k >= c3 ? 1u : 0u);
            }
            c = __reduce_add_sync(0xffffffffu, c);
            d = __reduce_add_sync(0xffffffffu, d);
            const int n1 = c & 0xffff, n2 = c >> 16, n3 = (int)d;
            if (n3 >= L) { T = c3; nT = n3; }
            else if (n2 >= L) { T = c2; nT = n2; }
            else if (n1 >= L) { T = c1; nT = n1; }
        }
        // keep keys > T, and keys == T in position order while there is room (nT <= Lmax: all of them)
        int n_gt = 0;
        for (int e = lane; e < total; e += 32) n_gt += pk[e] > T;
        n_gt = __reduce_add_sync(0xffffffffu, n_gt);
        const int need_eq = min(nT, Lmax) - n_gt;
        const unsigned lt = (1u << lane) - 1u;
        int eq_before = 0, kept_before = 0;
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + lane;
            const unsigned k = e < total ? pk[e] : 0u;
            const bool is_eq = e < total && k == T;
            const unsigned eqm = __ballot_sync(0xffffffffu, is_eq);
            const bool keep = e < total && (k > T || (is_eq && eq_before + __popc(eqm & lt) < need_eq));
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int slot = kept_before + __popc(km & lt);
                if (slot < RS_MAXSEL) {                       // (always true: nT <= Lmax <= RS_MAXSEL; belt and braces)
                    s_av[wid][slot] = key2f(k);
                    s_j[wid][slot] = pj[e];
                }
            }
            eq_before += __popc(eqm);
            kept_before += __popc(km);
        }
        have = min(kept_before, RS_MAXSEL);
    }
    if (lane == 0 && have < L) block_flags[i / 64] = 1;       // stale warm-start hint: exact fix-up
    __syncwarp();
    // ---- exact re-score, 32 candidates at a time ----
    float amin = INFINITY;
    for (int l0 = 0; l0 < have; l0 += 32) {
        const int l = l0 + lane;
        const int j = l < have ? s_j[wid][l] : -1;
        // stage the 32 candidate half-rows with coalesced 16-byte loads, 8 loads in flight per lane
        constexpr int LPR = DH / 4, RPP = 32 / LPR;           // lanes per half-row, rows per warp-wide load
        const int sub = lane / LPR, dq = 4 * (lane % LPR);
        float dot = 0.f;
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8 * RPP) {
                float4 r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int jc = __shfl_sync(0xffffffffu, j, c0 + u * RPP + sub);
                    r[u] = jc >= 0 ? __ldg(reinterpret_cast<const float4*>(V + (size_t)jc * D + hf * DH + dq))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<float4*>(tile + (c0 + u * RPP + sub) * TS + dq) = r[u];
            }
            __syncwarp();
            if (j >= 0) {
                // one fmaf chain over d = 0..D-1, exactly as graph_build.cu accumulates it
                const float4* vj4 = reinterpret_cast<const float4*>(tile + lane * TS);
                const float4* vi4 = reinterpret_cast<const float4*>(svi + hf * DH);
#pragma unroll 8
                for (int q = 0; q < DH / 4; ++q) {
                    const float4 a = vi4[q], b = vj4[q];
                    dot = fmaf(a.x, b.x, dot);
                    dot = fmaf(a.y, b.y, dot);
                    dot = fmaf(a.z, b.z, dot);
                    dot = fmaf(a.w, b.w, dot);
                }
            }
            __syncwarp();
        }
        if (j >= 0) {
            const float c = dot / (ni * nrm[j]);
            // ranking key: exact cosine descending, then index ascending
            s_key[wid][l] = ((unsigned long long)f2key(c) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j);
            amin = fminf(amin, s_av[wid][l]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
    int above = 0;
    for (int l = lane; l < have; l += 32) above += s_av[wid][l] > amin + 2.f * TC_EPS ? 1 : 0;
    above = __reduce_add_sync(0xffffffffu, above);
    if (lane == 0 && above < K) block_flags[i / 64] = 1;      // the slack window is ambiguous: exact fix-up
    __syncwarp();
    // ---- rank = number of larger keys; a lane ranks its (up to three) candidates in one sweep ----
    {
        unsigned long long k0 = lane < have ? s_key[wid][lane] : ~0ull;
        unsigned long long k1 = lane + 32 < have ? s_key[wid][lane + 32] : ~0ull;
        unsigned long long k2 = lane + 64 < have ? s_key[wid][lane + 64] : ~0ull;
        int r0 = 0, r1 = 0, r2 = 0;
        for (int q = 0; q < have; ++q) {
            const unsigned long long kq = s_key[wid][q];
            r0 += kq > k0; r1 += kq > k1; r2 += kq > k2;
        }
        const unsigned long long ks[3] = {k0, k1, k2};
        const int rs[3] = {r0, r1, r2};
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int l = lane + 32 * u;
            if (l < have && rs[u] < K) {
                s_out[wid][rs[u]] = (int)(0xffffffffu - (unsigned)(ks[u] & 0xffffffffull));
                if (rs[u] == K - 1 && kth_out != nullptr) kth_out[i] = key2f((unsigned)(ks[u] >> 32));
            }
        }
    }
    __syncwarp();
    if (idx_out != nullptr)
        for (int k = lane; k < K; k += 32) idx_out[(size_t)i * K + k] = (int64_t)s_out[wid][k];
    if (nbr_out != nullptr) {
        // neighbour list: the top-k without the row itself, then the row itself, then negative padding; the last slot
        // also records where the row itself sat in the top-k (-2 - position): idx can be rebuilt from this table alone
        int32_t* nb = nbr_out + (size_t)i * (K + 1);
        const unsigned lt = (1u << lane) - 1u;
        int o = 0, pos = -1;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const int j = k < K ? s_out[wid][k] : i;
            const unsigned m = __ballot_sync(0xffffffffu, j != i);
            const unsigned ms = __ballot_sync(0xffffffffu, k < K && j == i);
            if (pos < 0 && ms != 0u) pos = k0 + __ffs(ms) - 1;
            if (j != i) nb[o + __popc(m & lt)] = j;
            o += __popc(m);
        }
        if (lane == 0) nb[o] = i;
        for (int k = o + 1 + lane; k < K + 1; k += 32) nb[k] = (k == K && pos >= 0) ? -2 - pos : -1;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

bool gram_tc_supported(int N, int D, int K) {
    return N >= 1024 && (D == 64 || D == 128) && K + TC_SLACK <= TC_MAXL && K + TC_SLACK <= N;
}

// column splits per row block: fill the SMs when there are few row blocks (N = 4096 -> 32 blocks x 4; a rank's
// eighth of N = 16384 -> 16 blocks x 8; N = 16384 on one GPU -> 128 blocks x 2); bounded by what the re-score's merge
// pool holds
static void tc_split(int N, int D, int K, int blocks, bool cold, int* nsplit, int* tiles_per_split) {
    const int ntiles = ceil_div(N, TC_BN), L = K + TC_SLACK;
    int want = num_sms() / blocks;
    // warm-started sweeps never run as a single split: with two, a (row, warpgroup) segment ends below L entries and
    // skips its final compaction (measured at N = 16384: k_gram_tc 0.367 -> 0.26 ms, k_rescore 0.187 -> 0.21 for the two
    // extra segments it merges).  A cold sweep compacts anyway and prefers the longer columns (1.53 vs 2.26 ms).
    if (want < 2 && !cold) want = 2;
    static int force = -1;                       // diagnostics: GDN_TC_SPLIT forces the number of column splits
    if (force < 0) { const char* e = getenv("GDN_TC_SPLIT"); force = e ? atoi(e) : 0; }
    if (force > 0) want = force;
    if (want > TC_MAXSPLIT) want = TC_MAXSPLIT;
    while (want > 1 && 2 * (2 * want) * L > 32 * (D / 2 + 4)) --want;
    const int tps = ceil_div(ntiles, want);
    *tiles_per_split = tps;
    *nsplit = ceil_div(ntiles, tps);
}
// candidate-buffer capacity in (128-row block x column split) units: covers the full build and any row range
static size_t tc_units(int N, int D, int K) {
    int nsplit, tps;
    const int blocks = ceil_div(N, TC_BM);
    tc_split(N, D, K, blocks, false, &nsplit, &tps);
    size_t u = (size_t)blocks * nsplit;
    const size_t floor_u = (size_t)num_sms() + TC_MAXSPLIT;           // blocks_sub * nsplit_sub <= max(num_sms, blocks_sub)
    return u > floor_u ? u : floor_u;
}

size_t gram_tc_ws_bytes(int N, int D, int K) {
    const int C = TC_MAXC;
    const size_t units = tc_units(N, D, K);
    size_t b = align_up((size_t)N * sizeof(float), 256);
    b += 2 * align_up((size_t)N * D * sizeof(__nv_bfloat16), 256);
    b += 2 * align_up(units * TC_BM * 2 * C * sizeof(float), 256);
    b += align_up(units * TC_BM * 2 * sizeof(int), 256);
    b += align_up(((size_t)(N + 63) / 64 + 68) * sizeof(int), 256);
    return b;
}

int launch_gram_tc(const float* V, int N, int D, int K, int row0, int row1, int64_t* idx, int32_t* nbr, void* ws,
                   cudaStream_t st, float* kth, float margin, float** nrm_out, int** flags_out) {
    const int L = K + TC_SLACK, C = TC_MAXC, KB = D / TC_BK;
    const int rb0 = row0 / TC_BM, blocks = ceil_div(row1, TC_BM) - rb0;
    int nsplit, tps;
    const bool cold = kth == nullptr || !(margin < 1e30f);          // no hints, or the caller says they are not valid yet
    tc_split(N, D, K, blocks, cold, &nsplit, &tps);
    const int S = 2 * nsplit;
    const size_t units = tc_units(N, D, K);
    GDN_CHECK_ARG((size_t)blocks * nsplit <= units, "gram_tc: %d row blocks x %d splits exceed the workspace", blocks, nsplit);
    char* p = (char*)ws;
    float* nrm = (float*)p;                 p += align_up((size_t)N * sizeof(float), 256);
    __nv_bfloat16* hi = (__nv_bfloat16*)p;  p += align_up((size_t)N * D * sizeof(__nv_bfloat16), 256);
    __nv_bfloat16* lo = (__nv_bfloat16*)p;  p += align_up((size_t)N * D * sizeof(__nv_bfloat16), 256);
    float* bufv = (float*)p;                p += align_up(units * TC_BM * 2 * C * sizeof(float), 256);
    int* bufj = (int*)p;                    p += align_up(units * TC_BM * 2 * C * sizeof(float), 256);
    int* rowcnt = (int*)p;                  p += align_up(units * TC_BM * 2 * sizeof(int), 256);
    int* flags = (int*)p;                   // [ceil(N/64)] block flags, then the error word
    const int nblk64 = (N + 63) / 64;
    int* err = flags + nblk64;
    cudaError_t e = cudaMemsetAsync(flags, 0, ((size_t)nblk64 + 68) * sizeof(int), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset flags");

    int g = ceil_div(N, 8);
    if (g > 8 * num_sms()) g = 8 * num_sms();
    k_normalise_split<<<g, 256, 0, st>>>(V, N, D, nrm, hi, lo);
    GDN_CHECK_LAUNCH("k_normalise_split");

    PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
    GDN_CHECK_ARG(encode != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tm_hi, tm_lo;
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)N};
    const cuuint64_t gstr[1] = {(cuuint64_t)D * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, 64u};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r1 = encode(&tm_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hi, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tm_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lo, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GDN_CHECK_ARG(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);

    const size_t smem = 1024 + (size_t)2 * KB * TC_BM * 128 + (size_t)TC_STAGES * 2 * TC_BN * 128 +
                        (size_t)2 * 16 * TC_BM * sizeof(float) + (4 + 2 * TC_STAGES + 4) * 8 + 16;
    GDN_CHECK_ARG(smem <= 227 * 1024, "gram_tc: %zu bytes of shared memory needed", smem);
    e = ensure_dyn_smem(k_gram_tc, smem);
    if (e != cudaSuccess) return cuda_fail(e, "smem attr k_gram_tc");
    static int dbg = -1;
    if (dbg < 0) { const char* e_ = getenv("GDN_TC_DBG"); dbg = e_ ? atoi(e_) : 0; }
    k_gram_tc<<<dim3(blocks, nsplit), TC_THREADS, smem, st>>>(tm_hi, tm_lo, N, KB, L, C, tps, rb0, bufv, bufj, rowcnt, kth,
                                                              margin, err, dbg);
    GDN_CHECK_LAUNCH("k_gram_tc");
    GDN_CHECK_ARG(2 * S * L <= 32 * (D / 2 + 4), "gram_tc: candidate pool (%d) exceeds the re-score tile", 2 * S * L);
    const size_t rs_smem = (size_t)RS_WARPS * (32 * (D / 2 + 4) + D) * sizeof(float);
    auto rescore = D == 128 ? k_rescore<128> : k_rescore<64>;
    e = ensure_dyn_smem_ptr(reinterpret_cast<const void*>(rescore), rs_smem);
    if (e != cudaSuccess) return cuda_fail(e, "smem attr k_rescore");
    rescore<<<ceil_div(row1 - row0, RS_WARPS), RS_WARPS * 32, rs_smem, st>>>(V, nrm, row0, row1, K, L, C, S, bufv, bufj, rowcnt,
                                                                            kth, idx, nbr, flags, err);
    GDN_CHECK_LAUNCH("k_rescore");
    *nrm_out = nrm;
    *flags_out = flags;
    return 0;
}

}  // namespace gdn
