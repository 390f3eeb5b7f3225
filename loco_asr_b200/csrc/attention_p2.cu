// Fused self-attention with SpeechT5's relative-position bias on tcgen05 / TMEM / TMA: two independent pipelines per SM.
// (SpeechT5Attention, HF modeling_speecht5.py:872-986; SpeechT5RelativePositionalEncoding, HF:425-441.)
//
//   S[i, j] = q_i . k_j + q_i . pe_k[clip(i - j, -160, 159) + 160]      (q pre-scaled by 64^-1/2 * log2 e at load)
//   ctx_i   = sum_j softmax_j(S[i, :]) v_j        over the keys of the SAME utterance only
//
// Why two pipelines.  The phases of an attention item use different units: the bias-table drain (TMEM reads, F2FP, STS), the
// bias gather (one LDS.U16 per score: shared-memory wavefronts), the exponentials (MUFU, 16 / clk / SM), the epilogue (TMEM
// reads, STS / LDS / STG).  In attention_tc.cu all eight softmax warps of the SM work on ONE item and walk these phases in
// lock-step, so each phase is bound by its own unit while the others idle (phase trace at 128 frames: bias + max 1250 cycles
// at the LDS wavefront floor, exponentials 1400 at the MUFU floor, drain 900, epilogue 1200; 7300 per item with no unit
// above 20 % over the item).  Here the SM runs two items at once, each in its own half of the tensor-memory lanes, with its
// own loader warp, MMA-issuer warp, four softmax warps, K/V ring, bias table and barriers; the pipelines drift out of phase
// (different items, different lengths), so one gathers while the other exponentiates.
//
// An item is a (64-query tile, head) of one utterance: pipeline p owns TMEM lanes [64 p, 64 p + 64).  Every MMA is issued
// with M = 128 on the shared 128-row Q tile (rows 64 p.. hold the pipeline's queries; the other half's results land in lanes
// nobody reads), which doubles the tensor work of a kernel whose tensor pipe was 6-13 % busy, and buys 64-row granularity:
// 83 % of the lanes hold a query on the SLURP-shaped length mix against 72 % with 128-row tiles.
// TMEM columns of pipeline p (256 each): S [192, 256); G = Q pe_k^T in [0, 256) (a second round over [0, 64) for the rest of
// the up to 320 table columns -- utterances above 193 frames); once the table has been drained, O [0, 64), the row sums
// l [64, 80) and P [96, 128) (bf16 pairs, the A operand of P.V) take over G's columns.  When the table needs at most 192 columns (utterances up to 129 frames) S
// does not alias it and the item's first S is issued before the table exists; otherwise it waits for the drain (tab_done).
// The row sums are one more MMA, l = P . 1 on a tile of ones, accumulated beside O: what normalises O is exactly the bf16 P that
// was multiplied into it, and the softmax warps neither add nor exchange sums.  Per score: LDS.U16 + FHADD (bias), FMNMX3/2,
// FADD2/2, MUFU, F2FP/2.
// Each query row is shared by two threads (key halves of every 64-key block, in two warps of the same scheduler); they agree
// on the block maximum through shared memory, keep identical running maxima and split the accumulator rescale.
// Nothing overlaps between consecutive items of a pipeline except the TMA loads: the other pipeline is the overlap.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int PQ = 64;                  // queries per item
constexpr int FK = 64;                  // keys per block
constexpr int NS = 2;                   // ring entries per pipeline: (K_j | V_j)
constexpr int QT_LD = kRelCols + 8;     // fp16 table row pitch (656 B: conflict-free 16-byte row-per-thread stores)
constexpr int Q_TILE_B = 128 * 128;     // the shared 128-row Q tile of one slot (each pipeline fills its 64 rows)
constexpr int KV_TILE_B = FK * 128;
constexpr int ENTRY_B = 2 * KV_TILE_B;  // K part, then V part
constexpr int BOX_ROWS = 32;
constexpr int BOX_B = BOX_ROWS * 128;
constexpr int SM_PE = 0;
constexpr int SM_Q = SM_PE + kRelCols * 128;             // 2 slots
constexpr int SM_KV = SM_Q + 2 * Q_TILE_B;               // [2 pipelines][NS entries]
constexpr int SM_QT = SM_KV + 2 * NS * ENTRY_B;          // [2 pipelines][64 rows][QT_LD] fp16
constexpr int SM_ONES = SM_QT + 2 * PQ * QT_LD * 2;      // [16][64] bf16 ones: B operand of the row-sum MMA (2 KB)
#ifndef LOCO_P2_NH
#define LOCO_P2_NH 2
#endif
#ifndef LOCO_P2_SKEW
#define LOCO_P2_SKEW 0
#endif
constexpr int NH = LOCO_P2_NH;          // threads per query row: each takes FK / NH keys of every block, in NH warps of the same
                                        // TMEM lane quadrant.  NH = 4 (sixteen softmax warps at 96 registers, 16 scores per thread and
                                        // block) was built and measured: the gather, exponential and drain phases take the SAME number
                                        // of cycles as with NH = 2 (phase trace, profiles/r03_attention_p2_experiments.md) -- an item's
                                        // rows live in two TMEM lane quadrants = two SM sub-partitions whatever the warp count, and
                                        // each phase is bound by those two sub-partitions' MUFU / the SM's LSU, not by per-warp latency
#ifndef LOCO_P2_POLY_MASK
#define LOCO_P2_POLY_MASK 0xA           // every other pair of exponentials on the FMA pipe (common.cuh ex2_poly2): -5 % at 128-192 frames
#endif
constexpr int KW = FK / NH;             // keys per thread and block
constexpr int OW = kHeadDim / NH;       // accumulator columns per thread (rescale, epilogue)
static_assert(NH == 2 || NH == 4, "attention_p2: threads per row");
constexpr int SM_XH = SM_ONES + 16 * 128;                // [2 parities][2 pipelines][NH parts][64 rows] block maxima
constexpr int SM_DESC = SM_XH + 2 * 2 * NH * PQ * 4;     // [2 pipelines][4] int4 item descriptors
constexpr int SM_BARS = SM_DESC + 2 * 4 * 16;
constexpr int P2_SMEM = SM_BARS + 512 + 1024;
static_assert(P2_SMEM <= 232448, "attention_p2: shared memory budget");
constexpr int P2_SOFTMAX_WARPS = 4 * NH;
constexpr int WARP_LOAD = P2_SOFTMAX_WARPS;         // two loader warps: pipeline 0 / 1
constexpr int WARP_MMA = P2_SOFTMAX_WARPS + 2;      // two MMA issuer warps: pipeline 0 / 1
constexpr int P2_THREADS = (P2_SOFTMAX_WARPS + 4) * 32;
constexpr int PIPE_THREADS = 2 * NH * 32;           // softmax threads of one pipeline
constexpr int ROW_THREADS = NH * 32;                // softmax threads of one lane quadrant (the NH warps that share 32 rows)
constexpr int TM_S = 192, TM_P = 96, TM_O = 0, TM_L = 64, TM_PIPE = 256, P2_TMEM_COLS = 512;
constexpr int G_ROUND1 = 256;           // table columns of the first G round; the rest (<= 64) reuses columns [0, 64)
constexpr int G_LO_CHUNKS = 2;          // 32-column chunks that must be drained before the second round may be issued
constexpr int G_EARLY_S = 192;          // tables up to this many columns leave S's columns alone
#ifndef LOCO_LAZY_RESCALE
#define LOCO_LAZY_RESCALE 8.0f          // tools/parity_toggles.py builds a variant with 0 (rescale at every new maximum)
#endif
constexpr float kLazyRescale = LOCO_LAZY_RESCALE;    // log2 units

struct __align__(8) PBars {             // one set per pipeline
    uint64_t q_full[2], q_empty[2], kv_full[NS], kv_empty[NS];
    uint64_t g_full, g_lo_free, g2_full, tab_done, s_full, s_empty, p_full, pv_done, o_full, o_empty;
};
struct __align__(8) P2Bars {
    PBars p[2];
    uint64_t pe_full;
    uint32_t tmem_base;
};
static_assert(sizeof(P2Bars) <= 512, "P2Bars");

struct Item {
    int row0, i0, T, head, nr, n_kv, cbase, nc16;
    __device__ __forceinline__ int klen(int j) const { return min(FK, T - j * FK); }
    __device__ __forceinline__ void set(int4 d) {
        row0 = d.x; i0 = d.y; T = d.z; head = d.w;
        nr = min(PQ, T - i0);
        n_kv = (T + FK - 1) / FK;
        const int c_lo = max(i0 - (T - 1), -kMaxRel) + kMaxRel;
        const int c_hi = min(i0 + nr - 1, kMaxRel - 1) + kMaxRel;
        cbase = c_lo & ~15;
        nc16 = (c_hi - cbase + 16) & ~15;
    }
};

// tcgen05.ld / st of ONE 32-bit column of this warp's 32 lanes (the row sum)
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
// kind::f16 instruction descriptors, bf16 A (P, from TMEM) and B, M = 128, fp32 accumulate.  (Two exponentials per MUFU op --
// ex2.approx.f16x2 leaving P as fp16 pairs -- was tried: an fp16 A operand against the bf16 V is an illegal instruction, and
// with both descriptors forced to one format the kernel was no faster, 0.195 vs 0.189 ms per layer at 128 frames: the
// exponentials are not what the softmax warps wait for.)
constexpr uint32_t kIdescPV = umma_idesc_bf16(128, kHeadDim, /*b MN-major*/ 1);   // B = V, N = 64
constexpr uint32_t kIdescL = umma_idesc_bf16(128, 16, 0);                         // B = ones, K-major, N = 16

__device__ __forceinline__ uint32_t idesc_rt(int n) {      // M = 128, runtime N, both operands K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Bounded spin: a protocol bug records (tag, block, thread, parity) and traps -- the launch fails, nothing returns garbage.
__device__ int g_p2_timeout[8];
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity, int tag) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 22)) {
            if (atomicCAS(&g_p2_timeout[0], 0, tag) == 0) {
                g_p2_timeout[1] = (int)blockIdx.x;
                g_p2_timeout[2] = (int)threadIdx.x;
                g_p2_timeout[3] = (int)parity;
            }
            __threadfence();
            __trap();
        }
}

// Phase timeline of CTA 0, pipeline 0 (first 64 items), compiled in with -DLOCO_ATTN_TRACE: SM clock at fixed points of softmax
// warp 0 (role 0), softmax warp 4 (role 1: the other key half) and the MMA warp (role 2); printed when LOCO_ATTN_TRACE is set.
#ifdef LOCO_ATTN_TRACE
__device__ unsigned g_p2_trace[3][64][24];
#define TR(role, item_n, k)                                                                          \
    do {                                                                                             \
        if (blockIdx.x == 0 && lane == 0 && (item_n) < 64) {                                         \
            unsigned c_;                                                                             \
            asm volatile("mov.u32 %0, %%clock;" : "=r"(c_));                                         \
            g_p2_trace[role][item_n][k] = c_;                                                        \
        }                                                                                            \
    } while (0)
#else
#define TR(role, item_n, k) do { } while (0)
#endif

__global__ void __launch_bounds__(P2_THREADS, 1)
attention_p2_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap pe_map,
                    const PcTile* __restrict__ tiles, int n_items, bf16* __restrict__ ctx) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (sbase - smem_u32(smem_raw));
    P2Bars* bars = reinterpret_cast<P2Bars*>(smem_al + SM_BARS);
    int4* descs = reinterpret_cast<int4*>(smem_al + SM_DESC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&qkv_map);
        tma_prefetch_desc(&pe_map);
        mbar_init(smem_u32(&bars->pe_full), 1);
        for (int p = 0; p < 2; ++p) {
            PBars& b = bars->p[p];
            for (int s = 0; s < 2; ++s) {
                mbar_init(smem_u32(&b.q_full[s]), 1);
                mbar_init(smem_u32(&b.q_empty[s]), 1 + 2 * NH);      // the MMA warp + the pipeline's softmax warps
            }
            for (int s = 0; s < NS; ++s) {
                mbar_init(smem_u32(&b.kv_full[s]), 1);
                mbar_init(smem_u32(&b.kv_empty[s]), 1);
            }
            mbar_init(smem_u32(&b.g_full), 1);
            mbar_init(smem_u32(&b.g_lo_free), PIPE_THREADS);
            mbar_init(smem_u32(&b.g2_full), 1);
            mbar_init(smem_u32(&b.tab_done), PIPE_THREADS);
            mbar_init(smem_u32(&b.s_full), 1);
            mbar_init(smem_u32(&b.s_empty), PIPE_THREADS);
            mbar_init(smem_u32(&b.p_full), PIPE_THREADS);
            mbar_init(smem_u32(&b.pv_done), 1);
            mbar_init(smem_u32(&b.o_full), 1);
            mbar_init(smem_u32(&b.o_empty), PIPE_THREADS);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    for (int i = threadIdx.x; i < 16 * 128 / 4; i += P2_THREADS) reinterpret_cast<uint32_t*>(smem_al + SM_ONES)[i] = 0x3F803F80u;   // bf16 1.0
    fence_proxy_async_smem();            // read by the tensor core (async proxy)
    if (warp == WARP_MMA) tmem_alloc(smem_u32(&bars->tmem_base), P2_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    pdl_launch_dependents();
    pdl_wait();          // nothing before this line touches global data

    const int stride = 2 * (int)gridDim.x;
    if (warp == WARP_LOAD || warp == WARP_LOAD + 1) {
        // ===================== loaders: one per pipeline (pipeline 0's also brings pe_k) =====================
        const int p = warp - WARP_LOAD;
        PBars& B = bars->p[p];
        if (lane == 0) {
            if (p == 0) {
                const uint32_t pf = smem_u32(&bars->pe_full);
                mbar_arrive_expect_tx(pf, kRelCols * 128);
                tma_load_2d(sbase + SM_PE, &pe_map, pf, 0, 0);
                tma_load_2d(sbase + SM_PE + 160 * 128, &pe_map, pf, 0, 160);
            }
            uint32_t e = 0;
            int n = 0;
            for (int item = 2 * (int)blockIdx.x + p; item < n_items; item += stride, ++n) {
                const PcTile t = tiles[item / kHeads];
                const int head = item - (item / kHeads) * kHeads;
                const int qs = n & 1;
                bar_wait(smem_u32(&B.q_empty[qs]), (uint32_t)(((n >> 1) & 1) ^ 1), 101);
                descs[p * 4 + (n & 3)] = make_int4(t.row0, t.f0, t.t6, head);
                const int nr = min(PQ, t.t6 - t.f0);
                const int qbox = (nr + BOX_ROWS - 1) / BOX_ROWS;
                const uint32_t qf = smem_u32(&B.q_full[qs]);
                mbar_arrive_expect_tx(qf, qbox * BOX_B);
                for (int b = 0; b < qbox; ++b)
                    tma_load_2d(sbase + SM_Q + qs * Q_TILE_B + p * (PQ * 128) + b * BOX_B, &qkv_map, qf, head * kHeadDim, t.row0 + b * BOX_ROWS);
                const int kv_row = t.row0 - t.f0;
                const int n_kv = (t.t6 + FK - 1) / FK;
                for (int j = 0; j < n_kv; ++j, ++e) {
                    const int s = e % NS;
                    bar_wait(smem_u32(&B.kv_empty[s]), (uint32_t)(((e / NS) & 1) ^ 1), 102);
                    const int box = (min(FK, t.t6 - j * FK) + BOX_ROWS - 1) / BOX_ROWS;
                    const uint32_t kf = smem_u32(&B.kv_full[s]);
                    mbar_arrive_expect_tx(kf, 2 * box * BOX_B);
                    const uint32_t dst = sbase + SM_KV + (p * NS + s) * ENTRY_B;
                    for (int b = 0; b < box; ++b) {
                        tma_load_2d(dst + b * BOX_B, &qkv_map, kf, kHidden + head * kHeadDim, kv_row + j * FK + b * BOX_ROWS);
                        tma_load_2d(dst + KV_TILE_B + b * BOX_B, &qkv_map, kf, 2 * kHidden + head * kHeadDim, kv_row + j * FK + b * BOX_ROWS);
                    }
                }
            }
        }
    } else if (warp == WARP_MMA || warp == WARP_MMA + 1) {
        // ===================== MMA issuers: one per pipeline; warp-uniform control flow, one elected lane issues ==========
        const int p = warp - WARP_MMA;
        PBars& B = bars->p[p];
        const uint32_t tm = tmem + p * TM_PIPE;
        const uint32_t d_s = tm + TM_S, d_p = tm + TM_P, d_o = tm + TM_O, d_l = tm + TM_L;
        const uint64_t d_ones = umma_desc_sw128_kmajor(sbase + SM_ONES);
        const uint32_t ring = sbase + SM_KV + p * NS * ENTRY_B;
        uint32_t e = 0, n_blk = 0, n_g2 = 0;      // ring entries consumed, key blocks issued, items that needed a second G round
        int n = 0;
        if (p == 0) bar_wait(smem_u32(&bars->pe_full), 0, 204);
        for (int item = 2 * (int)blockIdx.x + p; item < n_items; item += stride, ++n) {
            const int qs = n & 1;
            if (p == 0) TR(2, n, 0);
            bar_wait(smem_u32(&B.q_full[qs]), (uint32_t)((n >> 1) & 1), 205);
            if (p == 0) TR(2, n, 1);
            if (p == 1 && n == 0) bar_wait(smem_u32(&bars->pe_full), 0, 204);
            Item it;
            it.set(descs[p * 4 + (n & 3)]);
            const uint64_t dq = umma_desc_sw128_kmajor(sbase + SM_Q + qs * Q_TILE_B);
            const uint64_t dp1 = umma_desc_sw128_kmajor(sbase + SM_PE + it.cbase * 128);
            const uint32_t id_g1 = idesc_rt(min(it.nc16, G_ROUND1));
            auto issue_s = [&](int j, uint32_t ee) {
                const int slot = ee % NS;
                bar_wait(smem_u32(&B.kv_full[slot]), (ee / NS) & 1, 203);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128_kmajor(ring + slot * ENTRY_B);
                const uint32_t id = idesc_rt((it.klen(j) + 15) & ~15);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_s, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), id, k);
                    umma_commit(smem_u32(&B.s_full));
                }
            };
            // S_0 first when its columns alias nothing of the table.  (The previous item's last S has been read: every softmax
            // thread arrives on s_empty before it arrives on p_full, and this warp waited for that block's p_full.  Each wait on
            // an already-complete mbarrier still costs this warp ~200 cycles of the item's critical path.)
            const bool early_s = it.nc16 <= G_EARLY_S;
            if (early_s) issue_s(0, e);
            if (p == 0) TR(2, n, 5);
            if (n > 0) bar_wait(smem_u32(&B.o_empty), (uint32_t)((n - 1) & 1), 206);      // the previous item's O has been read
            tc_fence_after();
            if (p == 0) TR(2, n, 2);
            if (elect_one()) {          // G = Q pe_k^T, first round
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tm, dq + (uint64_t)(k * 2), dp1 + (uint64_t)(k * 2), id_g1, k);
                umma_commit(smem_u32(&B.g_full));
            }
            if (p == 0) TR(2, n, 3);
            if (it.nc16 > G_ROUND1) {   // second round over the first 64 columns, once those are drained
                const uint64_t dp = umma_desc_sw128_kmajor(sbase + SM_PE + (it.cbase + G_ROUND1) * 128);
                const uint32_t id = idesc_rt(it.nc16 - G_ROUND1);
                bar_wait(smem_u32(&B.g_lo_free), n_g2 & 1, 207);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tm, dq + (uint64_t)(k * 2), dp + (uint64_t)(k * 2), id, k);
                    umma_commit(smem_u32(&B.g2_full));
                }
                ++n_g2;
            }
            if (!early_s) {
                bar_wait(smem_u32(&B.tab_done), (uint32_t)(n & 1), 208);      // the table has left TMEM: S may use its columns
                tc_fence_after();
                issue_s(0, e);
            }
            if (p == 0) TR(2, n, 4);
            // (P.V below waits for p_full, which every softmax thread gives only after it has drained its share of the table:
            //  O and P may then take over G's columns)
            for (int j = 0; j < it.n_kv; ++j, ++e, ++n_blk) {
                const int slot = e % NS;
                if (j + 1 < it.n_kv) {          // the next S goes out as soon as the softmax warps hold this one in registers
                    bar_wait(smem_u32(&B.s_empty), n_blk & 1, 213);
                    issue_s(j + 1, e + 1);
                }
                bar_wait(smem_u32(&B.p_full), n_blk & 1, 209);
                tc_fence_after();
                if (p == 0 && j == 0) TR(2, n, 6);
                if (p == 0 && j == it.n_kv - 1) TR(2, n, 8);
                const uint64_t dv = umma_desc_sw128_mnmajor(ring + slot * ENTRY_B + KV_TILE_B);
                const int ks = (it.klen(j) + 15) >> 4;
                if (elect_one()) {
                    for (int k = 0; k < ks; ++k) umma_bf16_ts(d_o, d_p + k * 8, dv + (uint64_t)(k * 128), kIdescPV, (j | k) != 0 ? 1u : 0u);
                    for (int k = 0; k < ks; ++k) umma_bf16_ts(d_l, d_p + k * 8, d_ones + (uint64_t)(k * 2), kIdescL, (j | k) != 0 ? 1u : 0u);   // l += P . 1
                    if (p == 0 && j == it.n_kv - 1) TR(2, n, 9);
                    umma_commit(smem_u32(&B.pv_done));
                    umma_commit(smem_u32(&B.kv_empty[slot]));
                }
                if (p == 0 && j == it.n_kv - 1) TR(2, n, 10);
            }
            if (elect_one()) {
                umma_commit(smem_u32(&B.o_full));
                umma_commit(smem_u32(&B.q_empty[qs]));
            }
            if (p == 0) TR(2, n, 7);
        }
    } else {
        // ===================== softmax warps =====================
        const int q = warp & 3;                    // TMEM lane quadrant
        const int h = warp >> 2;                   // which KW keys of a 64-key block (and which share of the drain / epilogue)
        const int p = q >> 1;                      // pipeline
        PBars& B = bars->p[p];
        const int rloc = (q & 1) * 32 + lane;      // row inside the item
        const uint32_t t_lane = tmem + p * TM_PIPE + ((uint32_t)(q * 32) << 16);
        const uint32_t t_s = t_lane + TM_S, t_o = t_lane + TM_O, t_p = t_lane + TM_P, t_l = t_lane + TM_L;
        __half* my_qt = reinterpret_cast<__half*>(smem_al + SM_QT) + (p * PQ + rloc) * QT_LD;
        float* xh = reinterpret_cast<float*>(smem_al + SM_XH);
        const int pair_id = 1 + q;                 // the two warps (key halves) that share this quadrant's rows
        const int pipe_id = 5 + p;                 // the pipeline's four softmax warps
        uint32_t cnt = 0, n_g2 = 0;                // key blocks seen, items with a second G round

        int n = 0;
#if LOCO_P2_SKEW > 0
        // Start pipeline 1 a fraction of an item late.  Both pipelines walk items of the same length (the batch is sorted), so
        // started together they stay in lock-step and meet in every phase -- both gather (LDS), both exponentiate (MUFU) -- and a
        // collision slows both equally, so it never dissolves.  Offset once, they stay offset for the same reason.
        if (p == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < LOCO_P2_SKEW) { }
        }
#endif
        for (int item = 2 * (int)blockIdx.x + p; item < n_items; item += stride, ++n) {
            if (q == 0 && h < 2) TR(h, n, 0);
            bar_wait(smem_u32(&B.g_full), (uint32_t)(n & 1), 310);      // also: the item's descriptor is in place
            tc_fence_after();
            if (q == 0 && h < 2) TR(h, n, 1);
            Item it;
            it.set(descs[p * 4 + (n & 3)]);
            const bool active = (q & 1) * 32 < it.nr;
            const int i = it.i0 + rloc;

            // ---- drain G (Q pe_k^T) into this row of the fp16 bias table: the two warps of the row split the 32-column chunks,
            // and only the columns this warp's 32 rows can reach are drained (T + 31 of the tile's T + 63)
            const int n_chunks = (it.nc16 + 31) >> 5;
            const int iw0 = it.i0 + (q & 1) * 32;
            const int w_lo = max(iw0 - (it.T - 1), -kMaxRel) + kMaxRel - it.cbase;
            const int w_hi = min(iw0 + 31, kMaxRel - 1) + kMaxRel - it.cbase;
            const int c_first = w_lo >> 5, c_last = min(w_hi >> 5, n_chunks - 1);
            auto store_chunk = [&](int c, const uint32_t (&v)[32]) {
                __half* dst = my_qt + c * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint4 o4;
                    __half2 hh;
                    hh = __floats2half2_rn(__uint_as_float(v[e + 0]), __uint_as_float(v[e + 1])); o4.x = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])); o4.y = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); o4.z = *reinterpret_cast<uint32_t*>(&hh);
                    hh = __floats2half2_rn(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])); o4.w = *reinterpret_cast<uint32_t*>(&hh);
                    *reinterpret_cast<uint4*>(dst + e) = o4;
                }
            };
            const bool two_rounds = it.nc16 > G_ROUND1;
            {   // round 1, one chunk at a time (keeping the next chunk's TMEM load in flight measured 7 % slower: registers)
                const int hi1 = min(c_last, G_ROUND1 / 32 - 1);
                if (active)
                    for (int c = h; c < G_LO_CHUNKS; c += NH)
                        if (c >= c_first && c <= hi1) {
                            uint32_t v[32];
                            tmem_ld_32x32(t_lane + c * 32, v);
                            tmem_ld_wait(v);
                            store_chunk(c, v);
                        }
                if (two_rounds) {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&B.g_lo_free));
                }
                if (active)
                    for (int c = G_LO_CHUNKS + ((h - G_LO_CHUNKS) & (NH - 1)); c <= hi1; c += NH)      // chunk c belongs to part c % NH
                        if (c >= c_first) {
                            uint32_t v[32];
                            tmem_ld_32x32(t_lane + c * 32, v);
                            tmem_ld_wait(v);
                            store_chunk(c, v);
                        }
            }
            if (two_rounds) {
                bar_wait(smem_u32(&B.g2_full), n_g2 & 1, 311);
                ++n_g2;
                tc_fence_after();
                if (active)
                    for (int c = G_ROUND1 / 32 + h; c <= c_last; c += NH) {
                        if (c < c_first) continue;
                        uint32_t v[32];
                        tmem_ld_32x32(t_lane + (c - G_ROUND1 / 32) * 32, v);
                        tmem_ld_wait(v);
                        store_chunk(c, v);
                    }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&B.tab_done));                   // G has left TMEM
            if (q == 0 && h < 2) TR(h, n, 2);
            if (active) named_bar_sync(pair_id, ROW_THREADS);     // all column sets of my rows are in the table

            float row_max = -INFINITY;
            const int ic = min(i, it.T - 1);                      // clamped row for table indexing on the slow path
            const int col0 = i + kMaxRel - it.cbase;              // table column of key 0
            for (int j = 0; j < it.n_kv; ++j, ++cnt) {
                const int jlen = it.klen(j);
                const bool mine = active && h * KW < jlen;        // this part of the block holds keys
                const int jc = j * FK + h * KW, mylen = jlen - h * KW;
                if (q == 0 && h < 2 && j == 0) TR(h, n, 3);
                bar_wait(smem_u32(&B.s_full), cnt & 1, 312);
                tc_fence_after();
                if (q == 0 && h < 2 && j == 0) TR(h, n, 4);
                if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 16);
                uint32_t su[KW];                   // scores, fp32 bit patterns (one array from the TMEM load to the exponentials)
                if (mine) tmem_ld_cols<KW>(t_s + h * KW, su);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < KW; ++e) asm volatile("" : "+r"(su[e]));
                if (q == 0 && h < 2 && j == 0) TR(h, n, 13);
                if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 17);
                tc_fence_before();
                mbar_arrive(smem_u32(&B.s_empty));                // the next S may be written
#define SC(e) __uint_as_float(su[e])
#define SET_SC(e, v) su[e] = __float_as_uint(v)
                float cbias = 0.f;                 // scalar bias of a fully clamped half block, folded into the exp argument
                float mloc = -INFINITY;
                if (mine) {
                    const int rel_max = iw0 + 31 - jc, rel_min = iw0 - (jc + KW - 1);
                    if (rel_max < kMaxRel && rel_min >= -kMaxRel) {
                        const unsigned short* base = reinterpret_cast<const unsigned short*>(my_qt) + (col0 - jc);
#pragma unroll
                        for (int e = 0; e < KW; ++e) SET_SC(e, add_f32_f16(SC(e), base[-e]));
                    } else if (rel_min >= kMaxRel - 1 || rel_max <= -kMaxRel) {
                        cbias = __half2float(my_qt[(rel_min >= kMaxRel - 1 ? kRelCols - 1 : 0) - it.cbase]);
                    } else {
                        const unsigned short* base = reinterpret_cast<const unsigned short*>(my_qt) + (kMaxRel - it.cbase);
#pragma unroll
                        for (int e = 0; e < KW; ++e) {
                            const int rel = max(-kMaxRel, min(kMaxRel - 1, ic - (jc + e)));
                            SET_SC(e, add_f32_f16(SC(e), base[rel]));
                        }
                    }
                    if (mylen < KW) {
#pragma unroll
                        for (int e = 0; e < KW; ++e)
                            if (e >= mylen) SET_SC(e, -INFINITY);
                    }
                    float cm[4] = {SC(0), SC(1), SC(2), SC(3)};      // four chains for ILP
#pragma unroll
                    for (int e = 4; e < KW; e += 4) {
                        cm[0] = fmaxf(cm[0], SC(e + 0));
                        cm[1] = fmaxf(cm[1], SC(e + 1));
                        cm[2] = fmaxf(cm[2], SC(e + 2));
                        cm[3] = fmaxf(cm[3], SC(e + 3));
                    }
                    mloc = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) + cbias;
                }
                // the NH threads of a row agree on the block maximum (double-buffered by block parity: the partners read my
                // value right after the barrier and I next write this buffer two barriers later)
                if (active) {
                    float* xhb = xh + ((cnt & 1) * 2 + p) * NH * PQ;
                    xhb[h * PQ + rloc] = mloc;
                    if (q == 0 && h < 2 && j == 0) TR(h, n, 5);
                    if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 18);
                    named_bar_sync(pair_id, ROW_THREADS);
#pragma unroll
                    for (int o = 1; o < NH; ++o) mloc = fmaxf(mloc, xhb[((h + o) & (NH - 1)) * PQ + rloc]);
                }
                if (q == 0 && h < 2 && j == 0) TR(h, n, 6);
                if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 19);
                if (j > 0) {            // P and O are free once the previous P.V has completed (issued a block ago)
                    bar_wait(smem_u32(&B.pv_done), (cnt - 1) & 1, 314);
                    tc_fence_after();
                }
                if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 20);
                // (only rows that exist vote: the lanes past the utterance's last query hold a neighbour's rows, and letting them
                //  trigger a rescale would make the valid rows' rounding depend on the batch the utterance travelled in.  All
                //  warps of a row see the same maxima, so they take the same decision and keep identical running maxima.)
                if (active && __any_sync(0xffffffffu, rloc < it.nr && mloc > row_max + kLazyRescale)) {
                    const float mx = fmaxf(row_max, mloc);
                    const float corr = ex2_approx(row_max - mx);      // first block: exp2(-inf) = 0
                    row_max = mx;
                    if (j > 0) {        // each of the row's warps rescales its share of the accumulator's columns; the first also the row sum
                        uint32_t v[OW];
                        tmem_ld_cols<OW>(t_o + h * OW, v);
                        uint32_t lsum = 0;
                        if (h == 0) lsum = tmem_ld_32x1(t_l);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < OW; ++e) asm volatile("" : "+r"(v[e]));
                        asm volatile("" : "+r"(lsum));
#pragma unroll
                        for (int e = 0; e < OW; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
                        tmem_st_cols<OW>(t_o + h * OW, v);
                        if (h == 0) tmem_st_32x1(t_l, __float_as_uint(__uint_as_float(lsum) * corr));
                    }
                }
                // ---- P = exp2(S - max) as packed bf16 pairs: my part's KW / 2 columns of the P buffer -----------------------
                if (mine) {
                    const float sub = cbias - row_max;
                    const float2 sub2 = make_float2(sub, sub);
                    uint32_t pp[KW / 2];
#pragma unroll
                    for (int e = 0; e < KW; e += 2) {
                        const float2 d = add_f32x2(make_float2(SC(e), SC(e + 1)), sub2);
                        const float2 pr = ex2_pair<LOCO_P2_POLY_MASK>(d, e >> 1);
                        pp[e >> 1] = pack_bf16(pr.x, pr.y);
                    }
                    if (q == 0 && h < 2 && j == 0) TR(h, n, 14);
                    tmem_st_cols<KW / 2>(t_p + h * (KW / 2), pp);
                }
                tmem_st_wait();
                if (q == 0 && h < 2 && j == 0) TR(h, n, 15);
                if (q == 0 && h < 2 && j == it.n_kv - 1) TR(h, n, 21);
                tc_fence_before();
                mbar_arrive(smem_u32(&B.p_full));
                if (q == 0 && h < 2 && j == 0) TR(h, n, 7);
#undef SC
#undef SET_SC
            }
            // ---- epilogue: O / l as bf16, staged in the item's own Q rows (its MMAs are done once o_full has completed; the
            // loader refills the slot only after this epilogue has arrived on q_empty) and written with full 128-byte lines
            const uint32_t stage = sbase + SM_Q + (n & 1) * Q_TILE_B + p * (PQ * 128);
            if (q == 0 && h < 2) TR(h, n, 8);
            bar_wait(smem_u32(&B.o_full), (uint32_t)(n & 1), 309);
            tc_fence_after();
            if (q == 0 && h < 2) TR(h, n, 9);
            if (active) {
                uint32_t v[OW];
                tmem_ld_cols<OW>(t_o + h * OW, v);
                uint32_t lsum = tmem_ld_32x1(t_l);            // l = P . 1, accumulated by the tensor core beside O
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < OW; ++e) asm volatile("" : "+r"(v[e]));
                asm volatile("" : "+r"(lsum));
                const float inv = 1.0f / __uint_as_float(lsum);
#pragma unroll
                for (int e = 0; e < OW; e += 8) {
                    uint4 o4;
                    o4.x = pack_bf16(__uint_as_float(v[e + 0]) * inv, __uint_as_float(v[e + 1]) * inv);
                    o4.y = pack_bf16(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv);
                    o4.z = pack_bf16(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv);
                    o4.w = pack_bf16(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv);
                    const int chunk = h * (OW / 8) + (e >> 3);          // 16-byte chunk of the row's 128-byte line
                    sts128(stage + rloc * 128 + ((chunk ^ (rloc & 7)) << 4), o4);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&B.o_empty));                    // O has left TMEM: the next item's G may be issued
            if (q == 0 && h < 2) TR(h, n, 10);
            named_bar_sync(pipe_id, PIPE_THREADS);                // all rows of the item are staged
            if (q == 0 && h < 2) TR(h, n, 11);
            {
                bf16* out = ctx + (int64_t)it.row0 * kHidden + it.head * kHeadDim;
                constexpr int RW = PQ / (2 * NH);                 // rows per warp: 8 lanes per row, 4 full lines per instruction
                const int w4 = (q & 1) * NH + h;                  // which of the pipeline's softmax warps: rows RW w4 + [0, RW)
#pragma unroll
                for (int r4 = 0; r4 < RW / 4; ++r4) {
                    const int r = w4 * RW + r4 * 4 + (lane >> 3), chunk = lane & 7;
                    if (r < it.nr) {
                        const uint4 o4 = lds128(stage + r * 128 + ((chunk ^ (r & 7)) << 4));
                        *reinterpret_cast<uint4*>(out + (int64_t)r * kHidden + chunk * 8) = o4;
                    }
                }
            }
            fence_proxy_async_smem();             // generic accesses to the slot are ordered before the TMA refill
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&B.q_empty[n & 1]));
            if (q == 0 && h < 2) TR(h, n, 12);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem, P2_TMEM_COLS);
    }
}

}  // namespace

int attention_p2_init() {
    return (int)cudaFuncSetAttribute(attention_p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
}

// tiles: 64-frame tiles of every utterance (PcTile.f0 a multiple of 64).
int launch_attention_p2(const void* qkv_map, const void* pe_map, const PcTile* tiles, int n_tiles, bf16* ctx, int num_sms,
                        cudaStream_t s) {
    if (n_tiles <= 0) return 0;
    const int n_items = n_tiles * kHeads;
    const int want = (n_items + 1) / 2;
    const int grid = want < num_sms ? want : num_sms;
    int rc = launch_pdl(attention_p2_kernel, dim3(grid), dim3(P2_THREADS), (size_t)P2_SMEM, s, *reinterpret_cast<const CUtensorMap*>(qkv_map),
                        *reinterpret_cast<const CUtensorMap*>(pe_map), tiles, n_items, ctx);
#ifdef LOCO_ATTN_TRACE
    static int traced = 0;
    if (!rc && getenv("LOCO_ATTN_TRACE") != nullptr && traced++ == 2) {
        static unsigned t[3][64][24];
        rc = (int)cudaStreamSynchronize(s);
        if (!rc) rc = (int)cudaMemcpyFromSymbol(t, g_p2_trace, sizeof t);
        const unsigned origin = t[0][6][0];
        fprintf(stderr, "attention_p2 phase trace of CTA 0 pipeline 0, items 6..9 (items %d, grid %d); SM clocks since softmax warp 0 entered item 6\n"
                        "softmax: 0 top, 1 g_full, 2 drained (tab_done), 3 at s_full wait, 4 s_full, 5 bias+max done, 6 maxima exchanged, 7 P stored (block 0), "
                        "8 key loop done, 9 o_full, 10 O staged, 11 all staged, 12 stored, 13 S in registers (block 0), 14 exponentials done (block 0), 15 P in TMEM (block 0)\n"
                        "          last block: 16 s_full, 17 S in registers, 18 bias + max done, 19 maxima exchanged, 20 pv_done, 21 P in TMEM\n"
                        "mma: 0 top, 1 q_full, 2 o_empty, 3 G issued, 4 tab_done, 5 S0 issued, 6 p_full(0), 7 item done, last block: 8 p_full, 9 P.V issued, 10 committed\n", n_items, grid);
        const char* names[3] = {"softmax h0", "softmax h1", "mma"};
        for (int r = 0; r < 3; ++r)
            for (int n = 6; n < 10; ++n) {
                fprintf(stderr, "%-10s item %d:", names[r], n);
                for (int k = 0; k < (r == 2 ? 11 : 22); ++k) fprintf(stderr, " %6d", (int)(t[r][n][k] - origin));
                fprintf(stderr, "\n");
            }
    }
#endif
    static const bool debug = getenv("LOCO_ATTN_DEBUG") != nullptr;
    if (debug && !rc) {
        int t[8] = {0};
        rc = (int)cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(t, g_p2_timeout, sizeof t);
        if (t[0]) {
            fprintf(stderr, "loco: attention_p2 wait timed out: tag %d block %d thread %d parity %d (items %d grid %d)\n", t[0], t[1], t[2], t[3],
                    n_items, grid);
            int z[8] = {0};
            cudaMemcpyToSymbol(g_p2_timeout, z, sizeof z);
        }
    }
    return rc;
}

}  // namespace loco
