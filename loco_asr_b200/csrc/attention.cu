// Fused (flash-style) self-attention with SpeechT5's query-dependent relative-position bias
// (SpeechT5Attention, HF modeling_speecht5.py:872-986; SpeechT5RelativePositionalEncoding, HF:425-441).
//
//   S[i, j] = q_i . k_j + q_i . pe_k[clip(i - j, -160, 159) + 160]
//   ctx_i   = sum_j softmax_j(S[i, :]) v_j        over the keys of the SAME utterance only
// (q arrives scaled by head_dim^-0.5 * log2(e): the 1/8 of HF:891 and the exp -> exp2 conversion are folded into
//  the q projection at weight-load time, so scores are in log2 units and the softmax is a bare ex2.)
//
// The reference materialises pe_k[clip(i-j)+160] as a [T, T, 64] tensor (575 MB at 30 s) and contracts it
// with q; here the bias goes through the equivalent table  QT = Q . pe_k^T  computed once per CTA on the tensor
// cores -- only the table columns this query tile can reach, kept in shared memory as fp16 -- and the T x T
// score matrix never exists: an online softmax runs in registers while keys / values stream by.
//
// Variable-length utterances: one CTA per (64-query tile, head, utterance); no padding, no mask tensor.
// Everything the CTA consumes (needed pe_k tiles of 64 table rows, then 32-key K|V tiles) is ONE stream of 9 KB
// tiles pulled through a 4-deep cp.async ring with a single __syncthreads per tile; at 3 s utterances the kernel is latency-
// and issue-bound, not math-bound (ncu r1a: 15 us per CTA for ~3 us of MMA), so prefetch depth, CTAs/SM (4, was
// 2) and instructions per score element are what count: the bias add has three tile-level fast paths (table
// index never clamps / always clamps high / always clamps low) that need no per-element index arithmetic.  The MMAs are mma.sync (1.9 % + 1.9 % of FLOPs at 3 s); moving S and O into TMEM with tcgen05 is the
// planned next step for the long-context configs.
#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int AQ = 64;       // queries per CTA (16 per warp)
constexpr int AT = 32;       // keys per streamed K|V tile
constexpr int APE = 64;      // pe_k rows per streamed table tile
constexpr int ALD = 72;      // padded bf16 row length (144 B: conflict-free ldmatrix)
constexpr int ARING = 4;     // ring depth (prefetch distance 3 tiles)
constexpr int TILE_ELEMS = 64 * ALD;   // a ring slot holds 64 rows: one pe_k tile, or 32 keys followed by their 32 values
constexpr int QKV_LD = 3 * kHidden;

constexpr int QT_PAD = 24;   // slack columns: rows past T index up to 15 columns beyond the computed range
__host__ __device__ constexpr int attn_smem_bytes(int qt_cols) {
    return ARING * TILE_ELEMS * 2 + AQ * (qt_cols + QT_PAD) * 2;
}

// copy 32 rows x 64 bf16 from global (row stride ld) into a padded smem tile; rows >= valid are zero-filled
__device__ __forceinline__ void load_tile32(bf16* dst, const bf16* src, int64_t ld, int valid, int tid) {
#pragma unroll
    for (int i = tid; i < 32 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < valid;
        cp_async_16(smem_u32(dst + r * ALD + c * 8), src + (int64_t)(ok ? r : 0) * ld + c * 8, ok);
    }
}

__global__ void __launch_bounds__(128) attention_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ pe_k,
                                                        const UttMeta* __restrict__ meta, const int32_t* __restrict__ utt_index,
                                                        bf16* __restrict__ ctx, int qt_cols) {
    pdl_launch_dependents();
    pdl_wait();
    const UttMeta m = meta[utt_index != nullptr ? utt_index[blockIdx.z] : (int)blockIdx.z];
    const int T = m.t6;
    const int i0 = blockIdx.x * AQ;
    if (i0 >= T) return;
    const int head = blockIdx.y;
    extern __shared__ __align__(16) uint8_t smem[];
    bf16* ring = reinterpret_cast<bf16*>(smem);
    __half* sqt = reinterpret_cast<__half*>(ring + ARING * TILE_ELEMS);
    const int qt_ld = qt_cols + QT_PAD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;

    const bf16* q_g = qkv + (int64_t)(m.row6 + i0) * QKV_LD + head * kHeadDim;
    const bf16* k_g = qkv + (int64_t)m.row6 * QKV_LD + kHidden + head * kHeadDim;
    const bf16* v_g = k_g + kHidden;

    // table columns this tile can reach: rel = i - j, i in [i0, min(i0+63, T-1)], j in [0, T-1]
    const int i_hi = min(i0 + AQ - 1, T - 1);
    const bool warp_active = i0 + warp * 16 < T;     // warps whose 16 rows are all past T only help with the loads
    const int c_lo = max(i0 - (T - 1), -kMaxRel) + kMaxRel;
    const int c_hi = min(i_hi, kMaxRel - 1) + kMaxRel;
    const int chunk_lo = c_lo / APE;
    const int n_pe = c_hi / APE - chunk_lo + 1;
    const int cbase = (c_lo / 16) * 16;          // QT column 0; c_hi - cbase < qt_cols by construction on the host
    const int n_kv = (T + AT - 1) / AT;
    const int n_tiles = n_pe + n_kv;             // stream: pe_k tiles, then K0|V0, K1|V1, ...
    // table columns THIS WARP's 16 rows can reach (a subset of the CTA's): n-tile pairs outside it are skipped
    const int cw_lo = max(i0 + warp * 16 - (T - 1), -kMaxRel) + kMaxRel;
    const int cw_hi = min(min(i0 + warp * 16 + 15, T - 1), kMaxRel - 1) + kMaxRel;

    auto issue = [&](int s) {
        if (s < n_tiles) {
            bf16* dst = ring + (s % ARING) * TILE_ELEMS;
            if (s < n_pe) {
                const bf16* src = pe_k + (int64_t)(chunk_lo + s) * APE * kHeadDim;
                load_tile32(dst, src, kHeadDim, 32, tid);
                load_tile32(dst + 32 * ALD, src + 32 * kHeadDim, kHeadDim, 32, tid);
            } else {
                const int kv = s - n_pe;
                load_tile32(dst, k_g + (int64_t)kv * AT * QKV_LD, QKV_LD, T - kv * AT, tid);
                load_tile32(dst + 32 * ALD, v_g + (int64_t)kv * AT * QKV_LD, QKV_LD, T - kv * AT, tid);
            }
        }
        cp_async_commit();
    };

    // Q fragments straight from global (read once): a0 = (row g, dims 2t..2t+1), a1 = row g+8, a2/a3 = dims +8
    uint32_t qf[4][4];
    {
        const int r0 = i0 + warp * 16 + gq, r1 = r0 + 8;
        const bf16* p0 = q_g + (int64_t)(warp * 16 + gq) * QKV_LD + tq * 2;
        const bf16* p1 = p0 + 8 * QKV_LD;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qf[ks][0] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16)) : 0u;
            qf[ks][1] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16)) : 0u;
            qf[ks][2] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8)) : 0u;
            qf[ks][3] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8)) : 0u;
        }
    }
#pragma unroll
    for (int s = 0; s < ARING - 1; ++s) issue(s);

    const int b_row = (lane & 7) + ((lane >> 4) << 3);
    const int b_col = ((lane >> 3) & 1) * 8;

    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
    float row_max[2] = {-INFINITY, -INFINITY};
    float row_sum[2] = {0.f, 0.f};
    uint32_t pf[2][4];  // P as A-operand fragments for the two 16-key k-steps of the current key tile
    const int qi[2] = {i0 + warp * 16 + gq, i0 + warp * 16 + gq + 8};
    const __half* qt_row[2] = {sqt + (warp * 16 + gq) * qt_ld, sqt + (warp * 16 + gq + 8) * qt_ld};

    for (int s = 0; s < n_tiles; ++s) {
        cp_async_wait<ARING - 2>();
        __syncthreads();            // tile s has landed for everyone; everyone is done with tile s-1
        issue(s + ARING - 1);       // refill the slot tile s-1 occupied
        const bf16* tile = ring + (s % ARING) * TILE_ELEMS;
        if (!warp_active) continue;
        if (s < n_pe) {
            // ---- QT[:, tile] = Q . pe_k[tile]^T  (16 x 64 per warp), stored as fp16 ------------------------
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int c16 = (chunk_lo + s) * APE + np * 16;
                if (c16 + 15 < cw_lo || c16 > cw_hi) continue;   // warp-uniform
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t b[4];
                    ldmatrix_x4(b, smem_u32(tile + (np * 16 + b_row) * ALD + ks * 16 + b_col));
                    const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                    mma_16816(c0, qf[ks], b0);
                    mma_16816(c1, qf[ks], b1);
                }
                __half* r0 = sqt + (warp * 16 + gq) * qt_ld + (c16 - cbase) + tq * 2;
                __half* r1 = r0 + 8 * qt_ld;
                *reinterpret_cast<__half2*>(r0) = __floats2half2_rn(c0[0], c0[1]);
                *reinterpret_cast<__half2*>(r1) = __floats2half2_rn(c0[2], c0[3]);
                *reinterpret_cast<__half2*>(r0 + 8) = __floats2half2_rn(c1[0], c1[1]);
                *reinterpret_cast<__half2*>(r1 + 8) = __floats2half2_rn(c1[2], c1[3]);
            }
        } else {
            // ---- K|V tile: S = Q K^T (+ bias), online softmax, O += P V ----------------------------------
            const int j0 = (s - n_pe) * AT;
            float sc[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) sc[n][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    uint32_t b[4];
                    ldmatrix_x4(b, smem_u32(tile + (np * 16 + b_row) * ALD + ks * 16 + b_col));
                    const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                    mma_16816(sc[np * 2], qf[ks], b0);
                    mma_16816(sc[np * 2 + 1], qf[ks], b1);
                }
            }
            // ---- + relative-position bias: rel = i - j over this warp's 16 rows x the tile's 32 keys -----------
            const int rel_max = i0 + warp * 16 + 15 - j0;
            const int rel_min = i0 + warp * 16 - (j0 + AT - 1);
            if (rel_max < kMaxRel && rel_min >= -kMaxRel) {
                // no clamping anywhere in the tile: column = (i - j) + 160, constant offsets from a per-row base
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const __half* base = qt_row[r] + (qi[r] - j0 - tq * 2 + kMaxRel - cbase);
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        sc[n][r * 2 + 0] += __half2float(base[-n * 8]);
                        sc[n][r * 2 + 1] += __half2float(base[-n * 8 - 1]);
                    }
                }
            } else if (rel_min >= kMaxRel - 1 || rel_max <= -kMaxRel) {
                // every pair clamps to the same table edge: one scalar per query row
                const int col = (rel_min >= kMaxRel - 1 ? kRelCols - 1 : 0) - cbase;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float bias = __half2float(qt_row[r][col]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        sc[n][r * 2 + 0] += bias;
                        sc[n][r * 2 + 1] += bias;
                    }
                }
            } else {
#pragma unroll
                for (int n = 0; n < 4; ++n) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = e >> 1;
                        int rel = min(qi[r], T - 1) - (j0 + n * 8 + tq * 2 + (e & 1));
                        rel = max(-kMaxRel, min(kMaxRel - 1, rel)) + kMaxRel - cbase;
                        sc[n][e] += __half2float(qt_row[r][rel]);
                    }
                }
            }
            if (j0 + AT > T) {   // ragged last tile: keys >= T do not exist
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (j0 + n * 8 + tq * 2 + (e & 1) >= T) sc[n][e] = -INFINITY;
            }
            float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], sc[n][e]);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            }
            float corr[2], ps[2] = {0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                corr[r] = ex2_approx(row_max[r] - mx[r]);   // first tile: exp2(-inf) = 0
                row_max[r] = mx[r];
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                float p[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    p[e] = ex2_approx(sc[n][e] - mx[e >> 1]);
                    ps[e >> 1] += p[e];
                }
                pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
                pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) row_sum[r] = row_sum[r] * corr[r] + ps[r];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                o[n][0] *= corr[0]; o[n][1] *= corr[0];
                o[n][2] *= corr[1]; o[n][3] *= corr[1];
            }
            // O += P V ([key][dim] storage; ldmatrix.trans yields the (k = key, n = dim) operand)
            const bf16* vt = tile + 32 * ALD;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int dp = 0; dp < 4; ++dp) {
                    uint32_t b[4];
                    const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int dim = dp * 16 + (lane >> 4) * 8;
                    ldmatrix_x4_trans(b, smem_u32(vt + key * ALD + dim));
                    const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                    mma_16816(o[dp * 2], pf[kk], b0);
                    mma_16816(o[dp * 2 + 1], pf[kk], b1);
                }
            }
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------
    if (!warp_active) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = row_sum[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const int row = i0 + warp * 16 + gq + r * 8;
        if (row >= T) continue;
        const float inv = 1.0f / l;
        bf16* orow = ctx + (int64_t)(m.row6 + row) * kHidden + head * kHeadDim;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            *reinterpret_cast<uint32_t*>(orow + n * 8 + tq * 2) = pack_bf16(o[n][r * 2] * inv, o[n][r * 2 + 1] * inv);
    }
}

// columns of the bias table a 64-query tile can need when no utterance in the batch exceeds max_t6 frames:
// c_hi - c_lo <= T + 62, plus up to 15 for aligning column 0 down to a multiple of 16
int qt_cols_for(int max_t6) {
    int cols = (max_t6 + 62 + 16 + 15) / 16 * 16;
    return cols > kRelCols ? kRelCols : cols;
}

}  // namespace

int attention_init() {
    return (int)cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(kRelCols));
}

int launch_attention(const bf16* qkv, const bf16* pe_k, const UttMeta* meta, const int32_t* utt_index, int n_utts, int max_t6, bf16* ctx,
                     cudaStream_t s) {
    if (n_utts <= 0 || max_t6 <= 0) return 0;
    const int qt_cols = qt_cols_for(max_t6);
    dim3 grid((max_t6 + AQ - 1) / AQ, kHeads, n_utts);
    return launch_pdl(attention_kernel, grid, dim3(128), (size_t)attn_smem_bytes(qt_cols), s, qkv, pe_k, meta, utt_index, ctx, qt_cols);
}

}  // namespace loco
