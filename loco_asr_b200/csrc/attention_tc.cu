// Fused self-attention with SpeechT5's relative-position bias on tcgen05 / TMEM / TMA
// (SpeechT5Attention, HF modeling_speecht5.py:872-986; SpeechT5RelativePositionalEncoding, HF:425-441).
//
//   S[i, j] = q_i . k_j + q_i . pe_k[clip(i - j, -160, 159) + 160]      (q pre-scaled by 64^-1/2 * log2 e at load)
//   ctx_i   = sum_j softmax_j(S[i, :]) v_j        over the keys of the SAME utterance only
//
// Persistent, warp-specialised: one CTA per SM walks a list of work items (128-query tile x head of one utterance).
// pe_k (40 KB) is TMA-loaded once per CTA and stays in shared memory; Q tiles (2 slots) and 64-key K / V blocks (a 2-entry
// ring per softmax group, entry = V_j | K_{j+2}) stream in through TMA in 32-row boxes, so a short utterance moves only
// the rows it has.  TMEM map (512 columns):
//   S_g   = Q K_j^T        128 x 64 (N trimmed to the keys that exist)      [0, 64) group 0, [64, 128) group 1
//   P_g   bf16 pairs written by tcgen05.st, read back as the A operand      [128, 160) / [160, 192)
//   O_g  += P_j V_j        128 x 64; V_j is the MN-major B operand ([key][dim] exactly as TMA laid it out)   [192, 320)
//   G     = Q pe_k^T       128 x (table columns this tile can reach, <= 320) [320, 512): 192 columns, the rest in a second
//                          round over the first 128 columns once those are drained
// G is drained once per item into an fp16 table in shared memory (row i of the tile = row of the table), because the
// bias of key j sits at the per-row offset i - j + 160 -- a skew no TMEM load shape can express.  The reference's
// [T, T, 64] position_bias (575 MB at 30 s) and the T x T score matrix never exist.
//
// Twelve warps.  Two softmax groups of four warps ping-pong over the key blocks: group g takes blocks j = g, g+2, ... with
// one query row per thread (TMEM lane == row), its own S / P / O buffers, running maximum and row sum, so the groups never
// talk inside an item and one runs its exponentials (MUFU) while the other adds bias / takes maxima (LSU, ALU).  Each
// group has its own TMA loader warp and MMA issuer warp (group 0's also issues G and loads Q / pe_k): with 32-cycle MMAs
// (N = 64) a single issuing thread was the bottleneck (ncu: softmax warps 50 % in the s_full wait, MMA warp never idle).
// The issuer runs warp-uniform control flow with one elected lane, so descriptors stay in uniform registers.  S_{j+2} is
// issued as soon as the group has read S_j out of TMEM (s_empty), long before P_j exists, so the tensor round trip is off
// the critical path.  The running maximum is only raised when some row of the warp would exceed it by more than 2^8 (P stays
// <= 256), so the O rescale in TMEM is rare.  At the end of an item the two partial results are merged:
//   O = (2^(m_a - m) O_a + 2^(m_b - m) O_b) / (2^(m_a - m) l_a + 2^(m_b - m) l_b).
// Per score: LDS.U16 + FHADD (bias), FMNMX3, FADD2, MUFU.EX2, FADD2, F2FP -- about 5 issue slots.
//
// Items overlap: G_{n+1} is issued as soon as item n's last S is, S_{n+1,g} as soon as the group has read its last S of item
// n; the softmax warps drain G_{n+1} into the table and run item n's epilogue in between, hiding its last P.V and the
// second G round.
//
// mbarriers: pe_full; per item q_full[2]/q_empty[2] (slot n & 1; 2 arrivals: both MMA warps), g_full, g_lo_free (first 128 G
// columns drained, 256 arrivals), g2_full, ga_empty (table complete and previous o_full seen, 256 arrivals), o_full (2
// arrivals); per group and ring entry kv_full/kv_empty; per group and key block s_full, s_empty (128), p_full (128),
// pv_done.  Every softmax thread runs the same barrier skeleton whether or not its rows exist, and no waiter can be lapped
// by two phases of its barrier (see ga_empty / o_full).  Set LOCO_ATTN_DEBUG=1 to have a stuck wait reported per role.
//
// Measured (tools/attn_sweep.py, 64k frames per batch, per layer; profiles/r03b_attn_sweep.txt): 0.21 ms at T = 128, 0.29 ms at 256,
// 0.41 ms at 499, 0.84 ms at 1499, 1.49 ms at 2999 (411 TFLOP/s of Q K^T + P V + table FLOPs).  An item costs ~3.6 us + 0.95 us per
// key block whether its tile holds 128 rows or 4, so a tile that is at most half full goes to the two-pipeline kernel
// (attention_p2.cu) instead -- chosen per query tile from (frame count, row index) in api.cu.
// One pair of exponentials in four is evaluated on the FMA pipe (common.cuh ex2_poly2); setmaxnreg moves 64 registers per
// thread from the loader / MMA warpgroup to the softmax warpgroups (200 each: the key loop no longer spills).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int FQ = 128;                 // queries per item
constexpr int FK = 64;                  // keys per block
constexpr int NS = 2;                   // ring entries PER GROUP: (V_j | K_{j+2}) pairs, or a lone K for the group's first block of an item
constexpr int QT_LD = kRelCols + 8;     // fp16 table row pitch (656 B: conflict-free 16-byte row-per-thread stores)
constexpr int Q_TILE_B = FQ * 128;      // 128 rows x 64 dims bf16
constexpr int KV_TILE_B = FK * 128;
constexpr int ENTRY_B = 2 * KV_TILE_B;  // V part, then K part
constexpr int BOX_ROWS = 32;            // TMA box: 32 rows x 128 B
constexpr int BOX_B = BOX_ROWS * 128;
constexpr int SM_PE = 0;
constexpr int SM_Q = SM_PE + kRelCols * 128;
constexpr int SM_KV = SM_Q + 2 * Q_TILE_B;
constexpr int SM_QT = SM_KV + 2 * NS * ENTRY_B;
constexpr int SM_XM = SM_QT + FQ * QT_LD * 2;        // [2 groups][128 rows] running maxima at the end of an item
constexpr int SM_XL = SM_XM + 2 * FQ * 4;            // [2 groups][128 rows] row sums
constexpr int SM_DESC = SM_XL + 2 * FQ * 4;          // 4 x int4 item descriptors: slot n & 3 -- deeper than the Q ring, because a Q slot
                                                     // is released (S pre-issued) before the softmax warps have read its item's descriptor
constexpr int SM_BARS = SM_DESC + 64;
constexpr int FA_SMEM = SM_BARS + 256 + 1024;
static_assert(FA_SMEM <= 232448, "attention_tc: shared memory budget");
constexpr int FA_SOFTMAX_WARPS = 8;
constexpr int WARP_LOAD = FA_SOFTMAX_WARPS;                 // warps 8, 9: TMA loaders of group 0 / 1
constexpr int WARP_MMA = FA_SOFTMAX_WARPS + 2;              // warps 10, 11: MMA issuers of group 0 / 1
constexpr int FA_THREADS = (FA_SOFTMAX_WARPS + 4) * 32;     // 12 warps: the register file is allocated in 4-warp steps anyway
constexpr int FA_AUX_REGS = 104;         // setmaxnreg: the loader / MMA warpgroup (warps 8-11) releases 168 - 104 registers per thread,
constexpr int FA_SOFTMAX_REGS = 200;    // the two softmax warpgroups take 32 each: no spills (168 for everybody left 92 B of them in the key loop)
constexpr int TM_S = 0, TM_P = 128, TM_O = 192, TM_G = 320, FA_TMEM_COLS = 512;   // S 2x64, P 2x32, O 2x64, G 192
constexpr int G_ROUND1 = 192;           // G columns of the first MMA round (6 chunks); the rest reuses the first 128 columns
constexpr int G_LO_CHUNKS = 4;          // chunks that must be drained before the second round may be issued
#ifndef LOCO_LAZY_RESCALE
#define LOCO_LAZY_RESCALE 8.0f          // tools/parity_toggles.py builds a variant with 0 (rescale at every new maximum)
#endif
constexpr float kLazyRescale = LOCO_LAZY_RESCALE;    // log2 units
#ifndef LOCO_TC_POLY_MASK
#define LOCO_TC_POLY_MASK 0x8           // one pair of exponentials in four on the FMA pipe (common.cuh ex2_poly2): -2 % at 256 and 2999 frames;
#endif                                  // every other pair (0xA) measured no faster, three in four slower: FFMA2 costs the FMA pipe as much as two FFMAs

struct __align__(8) FaBars {
    uint64_t pe_full, g_full, g_lo_free, g2_full, ga_empty, o_full;
    uint64_t s_full[2], s_empty[2], p_full[2], pv_done[2];
    uint64_t q_full[2], q_empty[2], kv_full[2][NS], kv_empty[2][NS];
    uint32_t tmem_base;
};
static_assert(sizeof(FaBars) <= 256, "FaBars");

// Key blocks of an utterance: FK keys each, the last one ragged.  Block j covers keys [key0(j), key0(j + 1)).
// (Equal-sized blocks that give both groups the same work were tried and lost: every block then has a ragged tail and
// leaves the 32-aligned fast paths of the bias add.)
__device__ __forceinline__ int num_key_blocks(int T) { return (T + FK - 1) / FK; }
__device__ __forceinline__ int key0(int j, int T, int n_kv) { return min(j * FK, T); }

struct Item {            // geometry of one (query tile, head) work item
    int row0;            // row of the tile's first query in the [R6, *] buffers
    int i0;              // that query's index inside its utterance
    int T;               // frames of the utterance
    int head;
    int nr;              // valid query rows in the tile
    int n_kv;            // key blocks
    __device__ __forceinline__ int k0(int j) const { return key0(j, T, n_kv); }
    __device__ __forceinline__ int klen(int j) const { return key0(j + 1, T, n_kv) - key0(j, T, n_kv); }
    int cbase;           // first pe_k row (= bias table column) the tile's G covers; multiple of 16
    int nc16;            // table columns computed (multiple of 16)
    __device__ __forceinline__ void set(int4 d) {
        row0 = d.x; i0 = d.y; T = d.z; head = d.w;
        nr = min(FQ, T - i0);
        n_kv = num_key_blocks(T);
        const int c_lo = max(i0 - (T - 1), -kMaxRel) + kMaxRel;
        const int c_hi = min(i0 + nr - 1, kMaxRel - 1) + kMaxRel;
        cbase = c_lo & ~15;
        nc16 = (c_hi - cbase + 16) & ~15;
    }
};

__device__ __forceinline__ uint32_t idesc_rt(int n) {      // M = 128, runtime N
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FQ >> 4) << 24);
}

// Bounded spin without the printf of common.cuh's mbar_wait (the call would cost this kernel registers and spills).
// A protocol bug does not hang the GPU and never returns garbage: the first wait that runs out records (tag, block, thread,
// parity) in g_fa_timeout and TRAPS, so the launch fails and the next CUDA call of loco_encode / the caller's synchronise
// reports it (LOCO_ERR_CUDA); launch_attention_tc prints the record when LOCO_ATTN_DEBUG is set.
__device__ int g_fa_timeout[4 * 4 + 1];      // [role][tag + 1, block, thread, parity], then a "someone timed out" flag
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
    uint32_t spins = 0;
    volatile int* flag = &g_fa_timeout[16];
    // (a suspend-time hint -- NANOSLEEP.SYNCS between probes -- was measured: it frees issue slots but wakes up late;
    //  0.25 -> 0.28 ms at 128 frames, 1.56 -> 1.66 ms at 2999)
    while (!mbar_try_wait(bar, parity))
        if ((++spins & 1023u) == 0 && (spins > (1u << 20) || (*flag != 0 && spins > (1u << 14)))) {
            const int role = tag / 100 == 3 ? 2 + (int)(threadIdx.x >> 7) : tag / 100 - 1;
            *flag = 1;
            if (atomicCAS(&g_fa_timeout[role * 4], 0, tag + 1) == 0) {
                g_fa_timeout[role * 4 + 1] = (int)blockIdx.x;
                g_fa_timeout[role * 4 + 2] = (int)threadIdx.x;
                g_fa_timeout[role * 4 + 3] = (int)parity;
            }
            __threadfence();
            __trap();
        }
}

// Phase timeline of CTA 0 (first 64 items), compiled in with -DLOCO_ATTN_TRACE: SM clock at fixed points of the softmax
// warps 0 / 4 (groups 0 / 1) and of the two MMA warps; printed by launch_attention_tc when LOCO_ATTN_TRACE is set.
#ifdef LOCO_ATTN_TRACE
__device__ unsigned g_fa_trace[10][64][16];
#define TR(role, item_n, k)                                                                          \
    do {                                                                                             \
        if (blockIdx.x == 0 && lane == 0 && (item_n) < 64) {                                         \
            unsigned c_;                                                                             \
            asm volatile("mov.u32 %0, %%clock;" : "=r"(c_));                                         \
            g_fa_trace[role][item_n][k] = c_;                                                        \
        }                                                                                            \
    } while (0)
#else
#define TR(role, item_n, k) do { } while (0)
#endif

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap pe_map,
                    const PcTile* __restrict__ tiles, int n_items, bf16* __restrict__ ctx) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (sbase - smem_u32(smem_raw));
    FaBars* bars = reinterpret_cast<FaBars*>(smem_al + SM_BARS);
    int4* descs = reinterpret_cast<int4*>(smem_al + SM_DESC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&qkv_map);
        tma_prefetch_desc(&pe_map);
        mbar_init(smem_u32(&bars->pe_full), 1);
        mbar_init(smem_u32(&bars->g_full), 1);
        mbar_init(smem_u32(&bars->g_lo_free), FA_SOFTMAX_WARPS * 32);
        mbar_init(smem_u32(&bars->g2_full), 1);
        mbar_init(smem_u32(&bars->ga_empty), FA_SOFTMAX_WARPS * 32);
        mbar_init(smem_u32(&bars->o_full), 2);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars->s_full[s]), 1);
            mbar_init(smem_u32(&bars->p_full[s]), FA_SOFTMAX_WARPS * 16);
            mbar_init(smem_u32(&bars->s_empty[s]), FA_SOFTMAX_WARPS * 16);
            mbar_init(smem_u32(&bars->pv_done[s]), 1);
            mbar_init(smem_u32(&bars->q_full[s]), 1);
            mbar_init(smem_u32(&bars->q_empty[s]), 2 + FA_SOFTMAX_WARPS);     // both MMA warps + every softmax warp's epilogue
        }
        for (int s = 0; s < 2 * NS; ++s) {
            mbar_init(smem_u32(&bars->kv_full[0][s]), 1);
            mbar_init(smem_u32(&bars->kv_empty[0][s]), 1);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == WARP_MMA) tmem_alloc(smem_u32(&bars->tmem_base), FA_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    pdl_launch_dependents();
    pdl_wait();          // the prologue above overlapped the previous kernel's tail; nothing before this line touches global data

    if (warp == WARP_LOAD || warp == WARP_LOAD + 1) {
        setmaxnreg_dec<FA_AUX_REGS>();
        // ===================== loaders: one per group (group 0's also brings pe_k and the Q tiles) =====================
        const int g = warp - WARP_LOAD;
        if (lane == 0) {
            if (g == 0) {
                const uint32_t pf = smem_u32(&bars->pe_full);
                mbar_arrive_expect_tx(pf, kRelCols * 128);
                tma_load_2d(sbase + SM_PE, &pe_map, pf, 0, 0);
                tma_load_2d(sbase + SM_PE + 160 * 128, &pe_map, pf, 0, 160);
            }
            uint32_t e = 0;     // ring entries of this group issued so far
            int n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
                const PcTile t = tiles[item / kHeads];
                const int head = item - (item / kHeads) * kHeads;
                if (g == 0) {
                    const int qs = n & 1;
                    bar_wait(smem_u32(&bars->q_empty[qs]), (uint32_t)(((n >> 1) & 1) ^ 1), 101);
                    descs[n & 3] = make_int4(t.row0, t.f0, t.t6, head);
                    const int nr = min(FQ, t.t6 - t.f0);
                    const int qbox = (nr + BOX_ROWS - 1) / BOX_ROWS;
                    const uint32_t qf = smem_u32(&bars->q_full[qs]);
                    mbar_arrive_expect_tx(qf, qbox * BOX_B);
                    for (int b = 0; b < qbox; ++b)
                        tma_load_2d(sbase + SM_Q + qs * Q_TILE_B + b * BOX_B, &qkv_map, qf, head * kHeadDim, t.row0 + b * BOX_ROWS);
                }
                const int kv_row = t.row0 - t.f0;
                const int n_kv = num_key_blocks(t.t6);
                // one ring entry = the operands of one MMA-warp step: V_jv (P.V of block jv) and K_jk (S of block jk = jv + 2);
                // either may be absent (-1).  Entries are filled in the order the group's MMA warp consumes them.
                auto load_entry = [&](int jv, int jk) {
                    const int s = e % NS;
                    bar_wait(smem_u32(&bars->kv_empty[g][s]), (uint32_t)(((e / NS) & 1) ^ 1), 102);
                    const uint32_t kf = smem_u32(&bars->kv_full[g][s]);
                    const int vbox = jv >= 0 ? (key0(jv + 1, t.t6, n_kv) - key0(jv, t.t6, n_kv) + BOX_ROWS - 1) / BOX_ROWS : 0;
                    const int kbox = jk >= 0 ? (key0(jk + 1, t.t6, n_kv) - key0(jk, t.t6, n_kv) + BOX_ROWS - 1) / BOX_ROWS : 0;
                    mbar_arrive_expect_tx(kf, (vbox + kbox) * BOX_B);
                    const uint32_t dst = sbase + SM_KV + (g * NS + s) * ENTRY_B;
                    for (int b = 0; b < vbox; ++b)
                        tma_load_2d(dst + b * BOX_B, &qkv_map, kf, 2 * kHidden + head * kHeadDim, kv_row + key0(jv, t.t6, n_kv) + b * BOX_ROWS);
                    for (int b = 0; b < kbox; ++b)
                        tma_load_2d(dst + KV_TILE_B + b * BOX_B, &qkv_map, kf, kHidden + head * kHeadDim, kv_row + key0(jk, t.t6, n_kv) + b * BOX_ROWS);
                    ++e;
                };
                if (g < n_kv) load_entry(-1, g);
                for (int j = g; j < n_kv; j += 2) load_entry(j, j + 2 < n_kv ? j + 2 : -1);
            }
        }
    } else if (warp == WARP_MMA || warp == WARP_MMA + 1) {
        setmaxnreg_dec<FA_AUX_REGS>();
        // ===================== MMA issuers: one per group =====================
        // Group g's warp issues S_j and P_j.V_j for its blocks j = g, g+2, ...; group 0's also issues G.  The whole warp runs
        // the (uniform) control flow so descriptors live in uniform registers; one elected lane issues the tcgen05
        // instructions.  With 32-cycle MMAs (N = 64) the issue path itself is what has to stay short.
        const int g = warp - WARP_MMA;
        constexpr uint32_t idesc_o = umma_idesc_bf16(FQ, kHeadDim, /*b MN-major*/ 1);
        const uint32_t d_s = tmem + TM_S + g * FK, d_p = tmem + TM_P + g * (FK / 2), d_o = tmem + TM_O + g * kHeadDim;
        const uint32_t ring = sbase + SM_KV + g * NS * ENTRY_B;
        const uint32_t kvf = smem_u32(&bars->kv_full[g][0]), kve = smem_u32(&bars->kv_empty[g][0]);
        const uint32_t pfull = smem_u32(&bars->p_full[g]), sfull = smem_u32(&bars->s_full[g]);
        const uint32_t sempty = smem_u32(&bars->s_empty[g]), pvdone = smem_u32(&bars->pv_done[g]);
        uint32_t e = 0;                     // ring entries consumed so far
        uint32_t n_s = 0;                   // S blocks issued by this warp (block k may be issued once block k-1 was read: s_empty)
        uint32_t n_pv = 0;                  // P.V products issued (parity of p_full)
        auto issue_g = [&](const Item& it, int qs, int round) {
            const uint64_t dq = umma_desc_sw128_kmajor(sbase + SM_Q + qs * Q_TILE_B);
            const uint64_t dp = umma_desc_sw128_kmajor(sbase + SM_PE + (it.cbase + round * G_ROUND1) * 128);
            const uint32_t id = idesc_rt(round ? it.nc16 - G_ROUND1 : min(it.nc16, G_ROUND1));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem + TM_G, dq + (uint64_t)(k * 2), dp + (uint64_t)(k * 2), id, k);
                umma_commit(smem_u32(round ? &bars->g2_full : &bars->g_full));
            }
        };
        // S_j = Q K_j^T from the K part of ring entry `slot`, as soon as the group has read its previous S out of TMEM
        auto issue_s = [&](const Item& it, int qs, int j, int slot) {
            if (n_s > 0) bar_wait(sempty, (n_s - 1) & 1, 213);
            ++n_s;
            tc_fence_after();
            const uint64_t dq = umma_desc_sw128_kmajor(sbase + SM_Q + qs * Q_TILE_B);
            const uint64_t dk = umma_desc_sw128_kmajor(ring + slot * ENTRY_B + KV_TILE_B);
            const uint32_t id = idesc_rt((it.klen(j) + 15) & ~15);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d_s, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), id, k);
                umma_commit(sfull);
            }
        };
        // the group's first block of an item, from the lone-K ring entry `ee`
        auto issue_s_first = [&](const Item& it, int qs, uint32_t ee) {
            const int slot = ee % NS;
            bar_wait(kvf + slot * 8, (ee / NS) & 1, 203);
            issue_s(it, qs, g, slot);
            if (elect_one()) umma_commit(kve + slot * 8);
        };
        // this warp's part of "Q slot free": after its last S (and, for group 0, after the last G round)
        auto release_q = [&](const Item& it, int qs) {
            if (elect_one()) {
                if (g < it.n_kv || g == 0) umma_commit(smem_u32(&bars->q_empty[qs]));
                else mbar_arrive(smem_u32(&bars->q_empty[qs]));
            }
        };
        Item cur, nxt;
        if (g == 0) bar_wait(smem_u32(&bars->pe_full), 0, 204);
        bar_wait(smem_u32(&bars->q_full[0]), 0, 205);
        cur.set(descs[0]);
        tc_fence_after();
        if (g == 0) issue_g(cur, 0, 0);
        if (g < cur.n_kv) {
            issue_s_first(cur, 0, e);
            ++e;
        }
        // warp-uniform non-blocking probe
        auto ready = [&](uint32_t bar, uint32_t parity) { return __shfl_sync(0xffffffffu, (int)mbar_test_wait(bar, parity), 0) != 0; };
        int n = 0;
        for (int item = blockIdx.x;; item += gridDim.x, ++n) {
            const int qs = n & 1;
            const bool has_next = item + (int)gridDim.x < n_items;
            TR(8 + g, n, 0);
            if (g == 0 && cur.nc16 > G_ROUND1) {       // second G round over the first 128 columns, once they are drained
                bar_wait(smem_u32(&bars->g_lo_free), (uint32_t)(n & 1), 206);
                tc_fence_after();
                issue_g(cur, qs, 1);
            }
            TR(8 + g, n, 1);
            const uint32_t e_next = e + (cur.n_kv > g ? (uint32_t)((cur.n_kv - g + 1) >> 1) : 0u);   // ring entry of the next item's lone K
            bool ga_seen = false;
            // The next item's prologue (G_{n+1}, this group's first S_{n+1}) is issued as soon as its inputs exist, but never
            // by blocking in front of this item's P.V products: the warp polls while it waits for P.  (Blocking here on the
            // next Q tile held back P_n.V_n, with it o_full and the epilogue, and through the Q slot the load after that:
            // a dependency cycle that cost ~1500 of the ~8000 cycles per item at <= 128 frames.)
            //   0: this warp still has S blocks of `cur` to issue   1: needs (group 0: the table out of TMEM,) Q_{n+1} -> G_{n+1}
            //   2: needs the K entry and a free S buffer -> S_{n+1, first}        3: done
            int nx = 0;
            auto start_next = [&]() {
                release_q(cur, qs);
                nx = has_next ? 1 : 3;
            };
            auto advance_next = [&](bool block) {
                if (nx == 1) {
                    if (g == 0 && !ga_seen) {
                        const uint32_t ga = smem_u32(&bars->ga_empty);
                        if (block) bar_wait(ga, (uint32_t)(n & 1), 207);
                        else if (!ready(ga, (uint32_t)(n & 1))) return;
                        ga_seen = true;
                    }
                    const uint32_t qf = smem_u32(&bars->q_full[qs ^ 1]), qp = (uint32_t)(((n + 1) >> 1) & 1);
                    if (block) bar_wait(qf, qp, 205);
                    else if (!ready(qf, qp)) return;
                    nxt.set(descs[(n + 1) & 3]);
                    tc_fence_after();
                    if (g == 0) issue_g(nxt, qs ^ 1, 0);
                    TR(8 + g, n, 4);
                    nx = g < nxt.n_kv ? 2 : 3;
                }
                if (nx == 2) {
                    if (!block) {
                        if (!ready(kvf + (e_next % NS) * 8, (e_next / NS) & 1)) return;
                        if (n_s > 0 && !ready(sempty, (n_s - 1) & 1)) return;
                    }
                    issue_s_first(nxt, qs ^ 1, e_next);
                    nx = 3;
                }
            };
            if (g + 2 >= cur.n_kv) start_next();        // the pre-issued S was this warp's only one
            for (int j = g; j < cur.n_kv; j += 2, ++e, ++n_pv) {
                const int slot = e % NS;
                bar_wait(kvf + slot * 8, (e / NS) & 1, 203);
                if (j == g) TR(8 + g, n, 5);
                // ---- the next S of this group goes out first: it only needs the S buffer (read early in the block) ----------
                if (j + 2 < cur.n_kv) {
                    issue_s(cur, qs, j + 2, slot);
                    if (j + 4 >= cur.n_kv) start_next();
                }
                // ---- O_g += P_j V_j ---------------------------------------------------------------------------------------------
                if (j == g) TR(8 + g, n, 6);
                // (backing off between probes -- __nanosleep(64), or try_wait once nothing is left to advance -- measured slower:
                //  0.227-0.243 -> 0.242-0.249 ms per layer at 128 frames, 0.319-0.327 -> 0.344 at 256)
                for (uint32_t spins = 0; !ready(pfull, n_pv & 1);) {
                    if (nx == 1 || nx == 2) advance_next(false);
                    if (++spins > (1u << 20)) {
                        bar_wait(pfull, n_pv & 1, 208);       // records the stuck wait
                        break;
                    }
                }
                if (j == g) TR(8 + g, n, 7);
                tc_fence_after();
                const uint64_t dv = umma_desc_sw128_mnmajor(ring + slot * ENTRY_B);
                const int valid = cur.klen(j);
                const uint32_t acc0 = j >= 2 ? 1u : 0u;
                if (valid >= FK) {           // full block: four K = 16 steps, 16 V rows (2048 B) each
                    if (elect_one()) {
                        umma_bf16_ts(d_o, d_p, dv, idesc_o, acc0);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_bf16_ts(d_o, d_p + k * 8, dv + (uint64_t)(k * 128), idesc_o, 1u);
                        umma_commit(pvdone);
                        umma_commit(kve + slot * 8);
                    }
                } else {
                    const int ks = (valid + 15) >> 4;
                    if (elect_one()) {
                        for (int k = 0; k < ks; ++k) umma_bf16_ts(d_o, d_p + k * 8, dv + (uint64_t)(k * 128), idesc_o, k > 0 ? 1u : acc0);
                        umma_commit(pvdone);
                        umma_commit(kve + slot * 8);
                    }
                }
            }
            // o_full needs both groups' last P.V; it may only complete once every softmax thread has seen the previous
            // item's o_full (ga_empty), so a waiter is never lapped by two phases
            TR(8 + g, n, 8);
            if (!ga_seen) {
                bar_wait(smem_u32(&bars->ga_empty), (uint32_t)(n & 1), 207);
                ga_seen = true;
            }
            if (elect_one()) {
                if (g < cur.n_kv) umma_commit(smem_u32(&bars->o_full));
                else mbar_arrive(smem_u32(&bars->o_full));
            }
            TR(8 + g, n, 9);
            if (!has_next) break;
            advance_next(true);
            if (g < nxt.n_kv) ++e;
            cur = nxt;
        }
    } else {
        // ===================== softmax warps =====================
        setmaxnreg_inc<FA_SOFTMAX_REGS>();
        const int q = warp & 3;                    // TMEM lane quadrant
        const int g = warp >> 2;                   // group: key blocks j = g, g + 2, ...
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t t_s = t_lane + TM_S + g * FK;
        const uint32_t t_o = t_lane + TM_O + g * kHeadDim;
        const uint32_t t_p = t_lane + TM_P + g * (FK / 2);
        __half* my_qt = reinterpret_cast<__half*>(smem_al + SM_QT) + row * QT_LD;
        float* xm = reinterpret_cast<float*>(smem_al + SM_XM);
        float* xl = reinterpret_cast<float*>(smem_al + SM_XL);
        const int bar_id = 1 + q;

        // deferred epilogue state (previous item)
        bool prev_active = false, prev_two = false;
        int prev_nr = 0;
        bf16* prev_out = nullptr;                  // first row of the item's tile, this head's 64 columns
        uint32_t cnt = 0;                          // key blocks this group has seen (parity of s_full[g])
        uint32_t n_g2 = 0;                         // items that needed a second G round

        int n = 0;
        // Item n - 1's result: merge the two groups' partial accumulators, stage the 128 x 64 bf16 tile in the item's own Q
        // slot (its MMAs are done once o_full has completed; the loader refills the slot only after this epilogue has
        // arrived on q_empty) and write it out with full 128-byte lines per row.  Storing straight from the one-row-per-thread
        // registers touched 32 different lines per instruction and cost ~2 us per item on the LSU (a third of the kernel at
        // 2-3 s utterances).
        auto epilogue = [&]() {
            bar_wait(smem_u32(&bars->o_full), (uint32_t)((n - 1) & 1), 309);
            TR(warp, n, 11);
            tc_fence_after();
            const uint32_t stage = sbase + SM_Q + ((n - 1) & 1) * Q_TILE_B;
            if (prev_active) {
                const float ma = xm[row], mb = xm[FQ + row];
                const float m = fmaxf(ma, mb);
                float wa = ex2_approx(ma - m), wb = ex2_approx(mb - m);
                const float inv = 1.0f / (wa * xl[row] + wb * xl[FQ + row]);
                wa *= inv;
                wb *= inv;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t va[16], vb[16];
                    tmem_ld_32x16(t_lane + TM_O + g * 32 + hh * 16, va);
                    if (prev_two) tmem_ld_32x16(t_lane + TM_O + kHeadDim + g * 32 + hh * 16, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        asm volatile("" : "+r"(va[e]));
                        asm volatile("" : "+r"(vb[e]));
                    }
                    float o[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(va[e]) * wa;
                    if (prev_two) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = fmaf(__uint_as_float(vb[e]), wb, o[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 16; e += 8) {
                        uint4 o4;
                        o4.x = pack_bf16(o[e + 0], o[e + 1]);
                        o4.y = pack_bf16(o[e + 2], o[e + 3]);
                        o4.z = pack_bf16(o[e + 4], o[e + 5]);
                        o4.w = pack_bf16(o[e + 6], o[e + 7]);
                        const int chunk = g * 4 + hh * 2 + (e >> 3);        // 16-byte chunk of the row's 128-byte line
                        sts128(stage + row * 128 + ((chunk ^ (row & 7)) << 4), o4);
                    }
                }
                tc_fence_before();
                TR(warp, n, 12);
                named_bar_sync(bar_id, 64);       // both halves of the quad's rows are staged, and both groups have read both
                                                  // accumulators before either P.V restarts them
                // warp (q, g) writes rows 32 q + 16 g + [0, 16): 8 lanes per row, 4 full lines per instruction
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const int r = q * 32 + g * 16 + r4 * 4 + (lane >> 3), chunk = lane & 7;
                    if (r < prev_nr) {
                        const uint4 o4 = lds128(stage + r * 128 + ((chunk ^ (r & 7)) << 4));
                        *reinterpret_cast<uint4*>(prev_out + (int64_t)r * kHidden + chunk * 8) = o4;
                    }
                }
            }
            TR(warp, n, 13);
            fence_proxy_async_smem();             // generic accesses to the slot are ordered before the TMA refill
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->q_empty[(n - 1) & 1]));
        };

        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
            // G_n complete also means the item's descriptor is in place (the MMA warp read it before issuing G_n).  The
            // Q barriers cannot be used for that: a Q slot may be released and refilled before a softmax thread looks.
            TR(warp, n, 0);
            bar_wait(smem_u32(&bars->g_full), (uint32_t)(n & 1), 310);
            TR(warp, n, 1);
            tc_fence_after();
            Item it;
            it.set(descs[n & 3]);
            const bool active = q * 32 < it.nr;    // same for both warps of a row quad
            const int i = it.i0 + row;             // query index inside the utterance (rows >= T are never stored)

            // ---- drain G (Q pe_k^T) into this row of the fp16 bias table; the two groups split the 32-column chunks -----
            if (active || prev_active) named_bar_sync(bar_id, 64);   // the other group is done with the previous item's table rows
                                                                      // and has published its maxima / sums
            const int n_chunks = (it.nc16 + 31) >> 5;
            // Only the table columns this warp's 32 rows can reach are drained: rel = i - j for its rows i and the keys j of the
            // utterance, clamped -- a window of T + 31 columns out of the tile's T + 127.
            const int w_lo = max(it.i0 + q * 32 - (it.T - 1), -kMaxRel) + kMaxRel - it.cbase;
            const int w_hi = min(it.i0 + q * 32 + 31, kMaxRel - 1) + kMaxRel - it.cbase;
            const int c_first = w_lo >> 5, c_last = min(w_hi >> 5, n_chunks - 1);
            auto drain = [&](int c, int tcol) {
                if (c < c_first || c > c_last) return;
                uint32_t v[32];
                tmem_ld_32x32(t_lane + TM_G + tcol, v);
                tmem_ld_wait(v);
                __half* dst = my_qt + c * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint4 o4;
                    __half2 h;
                    h = __floats2half2_rn(__uint_as_float(v[e + 0]), __uint_as_float(v[e + 1])); o4.x = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])); o4.y = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); o4.z = *reinterpret_cast<uint32_t*>(&h);
                    h = __floats2half2_rn(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])); o4.w = *reinterpret_cast<uint32_t*>(&h);
                    *reinterpret_cast<uint4*>(dst + e) = o4;
                }
            };
            if (active)
                for (int c = g; c < min(n_chunks, G_LO_CHUNKS); c += 2) drain(c, c * 32);     // round two reuses these columns
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->g_lo_free));
            if (active)
                for (int c = G_LO_CHUNKS + g; c < min(n_chunks, G_ROUND1 / 32); c += 2) drain(c, c * 32);

            // ---- previous item's epilogue (its last P.V ran while the table was drained; the second G round runs now) ----
            TR(warp, n, 2);
            if (n > 0) epilogue();
            TR(warp, n, 3);

            if (it.nc16 > G_ROUND1) {
                bar_wait(smem_u32(&bars->g2_full), n_g2 & 1, 311);
                TR(warp, n, 14);
                ++n_g2;
                tc_fence_after();
                if (active)
                    for (int c = G_ROUND1 / 32 + g; c < n_chunks; c += 2) drain(c, (c - G_ROUND1 / 32) * 32);
            }
            TR(warp, n, 15);
            if (active) named_bar_sync(bar_id, 64);       // both column sets of my rows are in place
            // G has left TMEM -- and this thread has seen o_full of the previous item, so o_full (which the MMA warps only
            // complete for item n after this barrier) can never run two phases ahead of a waiter
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->ga_empty));
            TR(warp, n, 4);

            float row_max = -INFINITY, row_sum = 0.f;
            const int iw0 = it.i0 + q * 32;            // first query row of this warp
            const int ic = min(i, it.T - 1);           // clamped row for table indexing on the slow path
            const int col0 = i + kMaxRel - it.cbase;   // table column of key 0
            for (int j = g; j < it.n_kv; j += 2, ++cnt) {
                const int j0 = it.k0(j), jlen = it.klen(j);
                const int nch = active ? (jlen + 31) >> 5 : 0;      // 32-key chunks that hold keys
                bar_wait(smem_u32(&bars->s_full[g]), cnt & 1, 312);
                if (j == g) TR(warp, n, 5);
                tc_fence_after();
                uint32_t su[2][32];                // scores, fp32 bit patterns (one array from the TMEM load to the exponentials)
                if (nch > 0) tmem_ld_32x32(t_s, su[0]);
                if (nch > 1) tmem_ld_32x32(t_s + 32, su[1]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) asm volatile("" : "+r"(su[c][e]));
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->s_empty[g]));     // the group's next S may be written
                if (j == g) TR(warp, n, 6);
#define SC(c, e) __uint_as_float(su[c][e])
#define SET_SC(c, e, v) su[c][e] = __float_as_uint(v)
                float cbias[2] = {0.f, 0.f};       // per-chunk scalar bias (clamped regions), folded into the exp argument
                float mloc = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c >= nch) continue;
                    const int jc = j0 + c * 32;
                    const int rel_max = iw0 + 31 - jc, rel_min = iw0 - (jc + 31);
                    if (rel_max < kMaxRel && rel_min >= -kMaxRel) {
                        const unsigned short* base = reinterpret_cast<const unsigned short*>(my_qt) + (col0 - jc);
#pragma unroll
                        for (int e = 0; e < 32; ++e) SET_SC(c, e, add_f32_f16(SC(c, e), base[-e]));
                    } else if (rel_min >= kMaxRel - 1 || rel_max <= -kMaxRel) {
                        cbias[c] = __half2float(my_qt[(rel_min >= kMaxRel - 1 ? kRelCols - 1 : 0) - it.cbase]);
                    } else {
                        const unsigned short* base = reinterpret_cast<const unsigned short*>(my_qt) + (kMaxRel - it.cbase);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int rel = max(-kMaxRel, min(kMaxRel - 1, ic - (jc + e)));
                            SET_SC(c, e, add_f32_f16(SC(c, e), base[rel]));
                        }
                    }
                    if (c * 32 + 32 > jlen) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c * 32 + e >= jlen) SET_SC(c, e, -INFINITY);
                    }
                    float cm[4] = {SC(c, 0), SC(c, 1), SC(c, 2), SC(c, 3)};      // four chains for ILP
#pragma unroll
                    for (int e = 4; e < 32; e += 4) {
                        cm[0] = fmaxf(cm[0], SC(c, e + 0));
                        cm[1] = fmaxf(cm[1], SC(c, e + 1));
                        cm[2] = fmaxf(cm[2], SC(c, e + 2));
                        cm[3] = fmaxf(cm[3], SC(c, e + 3));
                    }
                    mloc = fmaxf(mloc, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) + cbias[c]);
                }
                if (j == g) TR(warp, n, 7);
                if (cnt > 0) {          // P_g and O_g are free once this group's previous P.V has completed (issued a block ago)
                    bar_wait(smem_u32(&bars->pv_done[g]), (cnt - 1) & 1, 314);
                    tc_fence_after();
                }
                if (j == g) TR(warp, n, 8);
                // (only rows that exist vote: the lanes past the utterance's last query hold a neighbour's rows, and letting them
                //  trigger a rescale made the valid rows' rounding depend on the batch the utterance travelled in)
                if (active && __any_sync(0xffffffffu, row < it.nr && mloc > row_max + kLazyRescale)) {
                    const float mx = fmaxf(row_max, mloc);
                    const float corr = ex2_approx(row_max - mx);      // first block: exp2(-inf) = 0
                    row_max = mx;
                    row_sum *= corr;
                    if (j >= 2) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t v[32];
                            tmem_ld_32x32(t_o + hh * 32, v);
                            tmem_ld_wait(v);
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
                            tmem_st_32x32(t_o + hh * 32, v);
                        }
                    }
                }
                // ---- P = exp2(S - max) as packed bf16 pairs over the first 32 columns of this group's S buffer -------------
                float2 ps[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c >= nch) continue;
                    const float sub = cbias[c] - row_max;
                    const float2 sub2 = make_float2(sub, sub);
                    uint32_t pp[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float2 d = add_f32x2(make_float2(SC(c, e), SC(c, e + 1)), sub2);
                        const float2 p = ex2_pair<LOCO_TC_POLY_MASK>(d, e >> 1);
                        ps[(e >> 1) & 1] = add_f32x2(ps[(e >> 1) & 1], p);
                        pp[e >> 1] = pack_bf16(p.x, p.y);
                    }
                    tmem_st_32x16(t_p + c * 16, pp);
                }
                row_sum += (ps[0].x + ps[0].y) + (ps[1].x + ps[1].y);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->p_full[g]));
                if (j == g) TR(warp, n, 9);
#undef SC
#undef SET_SC
            }
            // ---- partial results of this group; the epilogue is deferred behind the next item's table drain ---------------
            TR(warp, n, 10);
            prev_active = active;
            if (active) {
                xm[g * FQ + row] = row_max;
                xl[g * FQ + row] = row_sum;
                prev_nr = it.nr;
                prev_two = it.n_kv > 1;
                prev_out = ctx + (int64_t)it.row0 * kHidden + it.head * kHeadDim;
            }
        }
        if (n > 0) {
            if (prev_active) named_bar_sync(bar_id, 64);      // the other group's maxima / sums are in place
            epilogue();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem, FA_TMEM_COLS);
    }
}

}  // namespace

int attention_tc_init() {
    return (int)cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
}

int launch_attention_tc(const void* qkv_map, const void* pe_map, const PcTile* tiles, int n_tiles, bf16* ctx, int num_sms,
                        cudaStream_t s) {
    if (n_tiles <= 0) return 0;
    const int n_items = n_tiles * kHeads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    int rc = launch_pdl(attention_tc_kernel, dim3(grid), dim3(FA_THREADS), (size_t)FA_SMEM, s, *reinterpret_cast<const CUtensorMap*>(qkv_map),
                        *reinterpret_cast<const CUtensorMap*>(pe_map), tiles, n_items, ctx);
#ifdef LOCO_ATTN_TRACE
    static int traced = 0;
    if (!rc && getenv("LOCO_ATTN_TRACE") != nullptr && traced++ == 2) {       // third launch: caches and clocks are warm
        static unsigned t[10][64][16];
        rc = (int)cudaStreamSynchronize(s);
        if (!rc) rc = (int)cudaMemcpyFromSymbol(t, g_fa_trace, sizeof t);
        const char* names[10] = {"softmax g0 q0", "softmax g0 q1", "softmax g0 q2", "softmax g0 q3", "softmax g1 q0", "softmax g1 q1",
                                 "softmax g1 q2", "softmax g1 q3", "mma g0", "mma g1"};
        const unsigned origin = t[0][6][0];
        fprintf(stderr, "attention_tc phase trace of CTA 0, items 6..8 (items %d, grid %d); SM clocks since softmax warp 0 entered item 6\n", n_items, grid);
        fprintf(stderr, "softmax points: 0 top, 1 g_full, 2 drained round 1, 11 o_full, 12 staged, 13 stored, 3 epilogue done, 14 g2_full, 15 drained round 2,\n"
                        "  4 ga_empty arrived, 5 s_full, 6 S in registers, 7 bias+max, 8 pv_done, 9 P stored, 10 item end\n"
                        "mma points: 0 top, 1 after G round 2, 4 next G issued, 5 V entry full, 6 before p_full wait, 7 p_full, 8 P.V issued, 9 o_full committed\n");
        const int order[16] = {0, 1, 2, 11, 12, 13, 3, 14, 15, 4, 5, 6, 7, 8, 9, 10};
        for (int r = 0; r < 10; ++r)
            for (int n = 6; n < 9; ++n) {
                fprintf(stderr, "%-14s item %d:", names[r], n);
                if (r < 8)
                    for (int k = 0; k < 16; ++k) fprintf(stderr, " %6d", (int)(t[r][n][order[k]] - origin));
                else
                    for (int k = 0; k < 10; ++k) fprintf(stderr, " %6d", k == 2 || k == 3 ? 0 : (int)(t[r][n][k] - origin));
                fprintf(stderr, "\n");
            }
    }
#endif
    static const bool debug = getenv("LOCO_ATTN_DEBUG") != nullptr;
    if (debug && !rc) {
        int t[17] = {0};
        rc = (int)cudaStreamSynchronize(s);
        if (!rc) rc = (int)cudaMemcpyFromSymbol(t, g_fa_timeout, sizeof t);
        if (t[16]) {
            for (int r = 0; r < 4; ++r)
                if (t[r * 4])
                    fprintf(stderr, "loco: attention_tc wait timed out: role %d tag %d block %d thread %d parity %d (items %d grid %d)\n", r,
                            t[r * 4] - 1, t[r * 4 + 1], t[r * 4 + 2], t[r * 4 + 3], n_items, grid);
            int z[17] = {0};
            cudaMemcpyToSymbol(g_fa_timeout, z, sizeof z);
        }
    }
    return rc;
}

}  // namespace loco
