// Fused self-attention with SpeechT5's relative-position bias on tcgen05 / TMEM / TMA
// (SpeechT5Attention, HF modeling_speecht5.py:872-986; SpeechT5RelativePositionalEncoding, HF:425-441).
//
//   S[i, j] = q_i . k_j + q_i . pe_k[clip(i - j, -160, 159) + 160]      (q pre-scaled by 64^-1/2 * log2 e at load)
//   ctx_i   = sum_j softmax_j(S[i, :]) v_j        over the keys of the SAME utterance only
//
// One CTA = 128 queries of one head of one utterance.  TMA brings the Q tile, the whole pe_k table and then
// 128-key K / V blocks (two-stage ring) into SWIZZLE_128B shared memory; one thread issues every MMA:
//   G  = Q pe_k^T          128 x 320 (two N = 160 MMAs x 4 K-steps)  -> TMEM, drained once to an fp16 table in smem
//   S_j = Q K_j^T          128 x 128, double-buffered in TMEM so S_{j+1} is computed while block j's softmax runs
//   O  += P_j V_j          128 x 64, accumulated IN TMEM; P_j (bf16) goes through a swizzled smem tile as the A operand,
//                          V_j is consumed as an MN-major B operand exactly as TMA laid it out ([key][dim])
// Four softmax warps own one query row per thread (TMEM lane == row): tcgen05.ld the 128 scores, add the bias from
// the thread's own table row (three warp-uniform paths: never clamps / clamps / mixed), online softmax in log2
// units, rescale O in TMEM only when some row's maximum moved, write P.  The T x T score matrix and the
// reference's [T, T, 64] position_bias (575 MB at 30 s) never exist.
//
// mbarrier protocol (all single-CTA): q_full (Q + pe_k landed) -> g_full (G done) -> qt_done (table drained, 128 arrivals);
// per block j: kv_full[s] / kv_empty[s] (s = j & 1), s_full[b] / s_empty[b] (b = j & 1), p_full (128 arrivals), pv_done.
#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int FQ = 128;                 // queries per CTA
constexpr int FK = 128;                 // keys per block
constexpr int QT_LD = kRelCols + 8;     // fp16 table row pitch (656 B: conflict-free 16-byte row-per-thread stores)
constexpr int SM_Q = 0;                                   // 128 x 64 bf16
constexpr int SM_KV = SM_Q + FQ * 128;                    // 2 stages x (K | V), each 128 x 64 bf16; pe_k first lives here
constexpr int SM_P = SM_KV + 2 * 2 * FK * 128;            // 2 atoms x 128 rows x 128 B
constexpr int SM_QT = SM_P + 2 * FQ * 128;                // 128 x QT_LD halves
constexpr int SM_BARS = SM_QT + FQ * QT_LD * 2;
constexpr int FA_SMEM = SM_BARS + 256 + 1024;
constexpr int FA_THREADS = 192;         // warp 0 loader, warp 1 MMA, warps 2-5 softmax
constexpr int TM_S0 = 0, TM_S1 = 128, TM_O = 256, FA_TMEM_COLS = 512;

struct __align__(8) FaBars {
    uint64_t q_full, g_full, qt_done, p_full, pv_done;
    uint64_t kv_full[2], kv_empty[2], s_full[2], s_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ CUtensorMap pe_map,
                    const UttMeta* __restrict__ meta, bf16* __restrict__ ctx) {
    const UttMeta m = meta[blockIdx.z];
    const int T = m.t6;
    const int i0 = blockIdx.x * FQ;
    if (i0 >= T) return;
    const int head = blockIdx.y;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (sbase - smem_u32(smem_raw));
    FaBars* bars = reinterpret_cast<FaBars*>(smem_al + SM_BARS);
    __half* sqt = reinterpret_cast<__half*>(smem_al + SM_QT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kv = (T + FK - 1) / FK;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&qkv_map);
        tma_prefetch_desc(&pe_map);
        mbar_init(smem_u32(&bars->q_full), 1);
        mbar_init(smem_u32(&bars->g_full), 1);
        mbar_init(smem_u32(&bars->qt_done), 128);
        mbar_init(smem_u32(&bars->p_full), 128);
        mbar_init(smem_u32(&bars->pv_done), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars->kv_full[s]), 1);
            mbar_init(smem_u32(&bars->kv_empty[s]), 1);
            mbar_init(smem_u32(&bars->s_full[s]), 1);
            mbar_init(smem_u32(&bars->s_empty[s]), 128);
        }
        mbar_fence_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), FA_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int q_row = m.row6 + i0;          // first row of the Q tile in the [R6, 2304] qkv matrix
    const int kv_row = m.row6;

    if (warp == 0) {
        // ===================== loader =====================
        if (lane == 0) {
            const uint32_t qf = smem_u32(&bars->q_full);
            mbar_arrive_expect_tx(qf, FQ * 128 + kRelCols * 128);
            tma_load_2d(sbase + SM_Q, &qkv_map, qf, head * kHeadDim, q_row);
            tma_load_2d(sbase + SM_KV, &pe_map, qf, 0, 0);
            tma_load_2d(sbase + SM_KV + 160 * 128, &pe_map, qf, 0, 160);
            mbar_wait(smem_u32(&bars->qt_done), 0);      // pe_k consumed: its smem becomes the K/V ring
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                mbar_wait(smem_u32(&bars->kv_empty[s]), (uint32_t)(((j >> 1) & 1) ^ 1));
                const uint32_t kf = smem_u32(&bars->kv_full[s]);
                mbar_arrive_expect_tx(kf, 2 * FK * 128);
                const uint32_t dst = sbase + SM_KV + s * (2 * FK * 128);
                tma_load_2d(dst, &qkv_map, kf, kHidden + head * kHeadDim, kv_row + j * FK);
                tma_load_2d(dst + FK * 128, &qkv_map, kf, 2 * kHidden + head * kHeadDim, kv_row + j * FK);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_g = umma_idesc_bf16(FQ, 160);
            constexpr uint32_t idesc_s = umma_idesc_bf16(FQ, FK);
            constexpr uint32_t idesc_o = umma_idesc_bf16(FQ, kHeadDim, /*b MN-major*/ 1);
            const uint64_t dq = umma_desc_sw128_kmajor(sbase + SM_Q);
            mbar_wait(smem_u32(&bars->q_full), 0);
            tc_fence_after();
            for (int hlf = 0; hlf < 2; ++hlf) {
                const uint64_t dpe = umma_desc_sw128_kmajor(sbase + SM_KV + hlf * 160 * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem + hlf * 160, dq + (uint64_t)(k * 2), dpe + (uint64_t)(k * 2), idesc_g, k);
            }
            umma_commit(smem_u32(&bars->g_full));
            mbar_wait(smem_u32(&bars->qt_done), 0);      // G drained: its TMEM columns become S0 / S1 / O
            tc_fence_after();
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(smem_u32(&bars->kv_full[s]), (uint32_t)((j >> 1) & 1));
                mbar_wait(smem_u32(&bars->s_empty[s]), (uint32_t)(((j >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128_kmajor(sbase + SM_KV + s * (2 * FK * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem + (s ? TM_S1 : TM_S0), dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k);
                umma_commit(smem_u32(&bars->s_full[s]));
            };
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_s(j + 1);
                mbar_wait(smem_u32(&bars->p_full), (uint32_t)(j & 1));
                tc_fence_after();
                const uint32_t v_base = sbase + SM_KV + (j & 1) * (2 * FK * 128) + FK * 128;
#pragma unroll
                for (int k = 0; k < FK / 16; ++k) {
                    const uint64_t dp = umma_desc_sw128_kmajor(sbase + SM_P + (k >> 2) * (FQ * 128) + (k & 3) * 32);
                    const uint64_t dv = umma_desc_sw128_mnmajor(v_base + k * (16 * 128));
                    umma_bf16(tmem + TM_O, dp, dv, idesc_o, (j | k) != 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&bars->pv_done));
                umma_commit(smem_u32(&bars->kv_empty[j & 1]));
            }
        }
    } else {
        // ===================== softmax warps: one query row per thread =====================
        const int q = warp & 3;                    // TMEM lane quadrant
        const int row = q * 32 + lane;
        const int i = i0 + row;                    // query index inside the utterance (rows >= T are never stored)
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        __half* my_qt = sqt + row * QT_LD;

        // ---- drain G (Q pe_k^T) into this row of the fp16 bias table ------------------------------------------
        mbar_wait(smem_u32(&bars->g_full), 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kRelCols / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + c * 32, v);
            tmem_ld_wait(v);
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
                uint4 o4;
                __half2 h;
                h = __floats2half2_rn(__uint_as_float(v[e + 0]), __uint_as_float(v[e + 1])); o4.x = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])); o4.y = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5])); o4.z = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7])); o4.w = *reinterpret_cast<uint32_t*>(&h);
                *reinterpret_cast<uint4*>(my_qt + c * 32 + e) = o4;
            }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->qt_done));

        float row_max = -INFINITY, row_sum = 0.f;
        const int iw0 = i0 + q * 32;               // first query row of this warp
        const int ic = min(i, T - 1);              // clamped row for table indexing on the slow path
        for (int j = 0; j < n_kv; ++j) {
            const int b = j & 1;
            const int j0 = j * FK;
            mbar_wait(smem_u32(&bars->s_full[b]), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            uint32_t su[FK / 32][32];
#pragma unroll
            for (int c = 0; c < FK / 32; ++c) tmem_ld_32x32(t_lane + (b ? TM_S1 : TM_S0) + c * 32, su[c]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < FK / 32; ++c)
#pragma unroll
                for (int e = 0; e < 32; ++e) asm volatile("" : "+r"(su[c][e]));
            float sc[FK];
#pragma unroll
            for (int c = 0; c < FK / 32; ++c)
#pragma unroll
                for (int e = 0; e < 32; ++e) sc[c * 32 + e] = __uint_as_float(su[c][e]);
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->s_empty[b]));      // S buffer may be overwritten by S_{j+2}

            // ---- + relative-position bias, key mask, row maximum ----------------------------------------------
            float mx = row_max;
            float cbias[FK / 32];                  // per-chunk scalar bias (clamped regions), folded into the exp argument
#pragma unroll
            for (int c = 0; c < FK / 32; ++c) {
                const int jc = j0 + c * 32;
                const int rel_max = iw0 + 31 - jc, rel_min = iw0 - (jc + 31);
                cbias[c] = 0.f;
                if (jc >= T) continue;             // whole chunk past the utterance: skipped again below
                if (rel_max < kMaxRel && rel_min >= -kMaxRel) {
                    const __half* base = my_qt + (i - jc + kMaxRel);      // column for key jc; key jc+e is e columns lower
#pragma unroll
                    for (int e = 0; e < 32; ++e) sc[c * 32 + e] += __half2float(base[-e]);
                } else if (rel_min >= kMaxRel - 1 || rel_max <= -kMaxRel) {
                    cbias[c] = __half2float(my_qt[rel_min >= kMaxRel - 1 ? kRelCols - 1 : 0]);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int rel = max(-kMaxRel, min(kMaxRel - 1, ic - (jc + e))) + kMaxRel;
                        sc[c * 32 + e] += __half2float(my_qt[rel]);
                    }
                }
                if (jc + 32 > T) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (jc + e >= T) sc[c * 32 + e] = -INFINITY;
                }
                float cm = sc[c * 32];
#pragma unroll
                for (int e = 1; e < 32; ++e) cm = fmaxf(cm, sc[c * 32 + e]);
                mx = fmaxf(mx, cm + cbias[c]);
            }
            const float corr = ex2_approx(row_max - mx);      // first block: exp2(-inf) = 0
            row_max = mx;

            // ---- O (in TMEM) and the P tile are free once PV_{j-1} has completed ----------------------------
            if (j > 0) {
                mbar_wait(smem_u32(&bars->pv_done), (uint32_t)((j - 1) & 1));
                tc_fence_after();
                if (!__all_sync(0xffffffffu, corr == 1.0f)) {     // some row's maximum moved: rescale this warp's O rows
#pragma unroll
                    for (int c = 0; c < kHeadDim / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_lane + TM_O + c * 32, v);
                        tmem_ld_wait(v);
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
                        tmem_st_32x32(t_lane + TM_O + c * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            // ---- P = exp2(S - max) as bf16 into the swizzled A-operand tile -----------------------------------
            float ps = 0.f;
            const uint32_t p_row = sbase + SM_P + row * 128;
#pragma unroll
            for (int kk = 0; kk < FK / 8; ++kk) {          // 16-byte chunks of 8 keys
                uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
                if (j0 + (kk >> 2) * 32 < T) {             // chunks wholly past the utterance: P = 0, no exps
                    const float sub = mx - cbias[kk >> 2];
                    float p[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        p[e] = ex2_approx(sc[kk * 8 + e] - sub);
                        ps += p[e];
                    }
                    o4.x = pack_bf16(p[0], p[1]);
                    o4.y = pack_bf16(p[2], p[3]);
                    o4.z = pack_bf16(p[4], p[5]);
                    o4.w = pack_bf16(p[6], p[7]);
                }
                sts128(p_row + (kk >> 3) * (FQ * 128) + (((kk & 7) ^ (row & 7)) << 4), o4);
            }
            row_sum = row_sum * corr + ps;
            fence_proxy_async_smem();     // P (generic proxy) -> visible to the tensor core's async proxy
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->p_full));
        }
        // ---- epilogue: O / l -> bf16 -> ctx -------------------------------------------------------------------
        mbar_wait(smem_u32(&bars->pv_done), (uint32_t)((n_kv - 1) & 1));
        tc_fence_after();
        const float inv = 1.0f / row_sum;
        bf16* orow = ctx + (int64_t)(m.row6 + i) * kHidden + head * kHeadDim;
#pragma unroll
        for (int c = 0; c < kHeadDim / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + TM_O + c * 32, v);
            tmem_ld_wait(v);
            if (i < T) {
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    uint4 o4;
                    o4.x = pack_bf16(__uint_as_float(v[e + 0]) * inv, __uint_as_float(v[e + 1]) * inv);
                    o4.y = pack_bf16(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv);
                    o4.z = pack_bf16(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv);
                    o4.w = pack_bf16(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + e) = o4;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, FA_TMEM_COLS);
    }
}

}  // namespace

int attention_tc_init() {
    return (int)cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
}

int launch_attention_tc(const void* qkv_map, const void* pe_map, const UttMeta* meta, int n_utts, int max_t6, bf16* ctx,
                        cudaStream_t s) {
    if (n_utts <= 0 || max_t6 <= 0) return 0;
    dim3 grid((max_t6 + FQ - 1) / FQ, kHeads, n_utts);
    attention_tc_kernel<<<grid, FA_THREADS, FA_SMEM, s>>>(*reinterpret_cast<const CUtensorMap*>(qkv_map),
                                                          *reinterpret_cast<const CUtensorMap*>(pe_map), meta, ctx);
    return (int)cudaGetLastError();
}

}  // namespace loco
