// Fused (flash-style) self-attention with SpeechT5's query-dependent relative-position bias
// (SpeechT5Attention, HF modeling_speecht5.py:872-986; SpeechT5RelativePositionalEncoding, HF:425-441).
//
//   S[i, j] = q_i . k_j + q_i . pe_k[clip(i - j, -160, 159) + 160]       (q already scaled by 1/8, HF:891)
//   ctx_i   = sum_j softmax_j(S[i, :]) v_j        over the keys of the SAME utterance only
//
// The reference materialises pe_k[clip(i-j)+160] as a [T, T, 64] tensor (575 MB at 30 s) and contracts it
// with q; here the bias goes through the equivalent table  QT = Q . pe_k^T  ([64 queries, 320]) computed once
// per CTA on the tensor cores and kept in shared memory, and the T x T score matrix never exists: keys/values
// stream through a double-buffered cp.async pipeline with an online softmax in registers.
// Variable-length utterances: one CTA per (64-query tile, head, utterance); no padding, no mask tensor.
// Round 1 uses mma.sync for QK^T / QT / PV (1.9 % + 1.9 % of FLOPs at 3 s); moving S and O into TMEM with
// tcgen05 is the planned next step for the long-context configs.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int AQ = 64;    // queries per CTA (16 per warp)
constexpr int AK = 32;    // keys per pipeline stage
constexpr int ALD = 72;   // padded bf16 row length of Q/K/V/pe tiles (144 B: conflict-free ldmatrix)
constexpr int QT_LD = 324;  // fp32 row length of the bias table
constexpr int A_SMEM = AQ * ALD * 2            // Q
                       + 2 * 2 * AK * ALD * 2  // K, V double buffered (also stages pe_k chunks)
                       + AQ * QT_LD * 4;       // QT
constexpr int QKV_LD = 3 * kHidden;

// copy `rows` x 64 bf16 from global (row stride ld) into a padded smem tile; rows >= valid are zero-filled
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int64_t ld, int rows, int valid, int tid) {
    for (int i = tid; i < rows * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < valid;
        cp_async_16(smem_u32(dst + r * ALD + c * 8), src + (int64_t)(ok ? r : 0) * ld + c * 8, ok);
    }
}

__global__ void __launch_bounds__(128) attention_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ pe_k,
                                                        const UttMeta* __restrict__ meta, bf16* __restrict__ ctx) {
    const UttMeta m = meta[blockIdx.z];
    const int T = m.t6;
    const int i0 = blockIdx.x * AQ;
    if (i0 >= T) return;
    const int head = blockIdx.y;
    extern __shared__ __align__(16) uint8_t smem[];
    bf16* sq = reinterpret_cast<bf16*>(smem);
    bf16* skv = sq + AQ * ALD;                               // [2 stages][K | V][AK][ALD]
    float* sqt = reinterpret_cast<float*>(skv + 4 * AK * ALD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;

    const bf16* q_g = qkv + (int64_t)(m.row6 + i0) * QKV_LD + head * kHeadDim;
    const bf16* k_g = qkv + (int64_t)m.row6 * QKV_LD + kHidden + head * kHeadDim;
    const bf16* v_g = k_g + kHidden;

    // ---- phase 0: Q tile + first pe_k chunk ------------------------------------------------------
    load_tile(sq, q_g, QKV_LD, AQ, T - i0, tid);
    load_tile(skv, pe_k, kHeadDim, 64, 64, tid);  // 64 pe rows fill the K|V halves of stage 0
    cp_async_commit();

    // ---- phase 1: QT[64, 320] = Q . pe_k^T, 5 chunks of 64 table rows ----------------------------
    uint32_t qf[4][4];  // Q fragments for the 4 k-steps (dims 0..63), reused by both phases
    const int a_row = warp * 16 + (lane & 15);
    const int a_col = (lane >> 4) * 8;
    const int b_row = (lane & 7) + ((lane >> 4) << 3);
    const int b_col = ((lane >> 3) & 1) * 8;
    for (int chunk = 0; chunk < kRelCols / 64; ++chunk) {
        if (chunk + 1 < kRelCols / 64)
            load_tile(skv + ((chunk + 1) & 1) * 2 * AK * ALD, pe_k + (int64_t)(chunk + 1) * 64 * kHeadDim, kHeadDim, 64, 64, tid);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (chunk == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], smem_u32(sq + a_row * ALD + ks * 16 + a_col));
        }
        const bf16* spe = skv + (chunk & 1) * 2 * AK * ALD;
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-column n-tiles: 64 table rows per chunk
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t b[4];
                ldmatrix_x4(b, smem_u32(spe + (np * 16 + b_row) * ALD + ks * 16 + b_col));
                const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                mma_16816(c0, qf[ks], b0);
                mma_16816(c1, qf[ks], b1);
            }
            float* r0 = sqt + (warp * 16 + gq) * QT_LD + chunk * 64 + np * 16 + tq * 2;
            float* r1 = r0 + 8 * QT_LD;
            r0[0] = c0[0]; r0[1] = c0[1]; r1[0] = c0[2]; r1[1] = c0[3];
            r0[8] = c1[0]; r0[9] = c1[1]; r1[8] = c1[2]; r1[9] = c1[3];
        }
        __syncthreads();
    }

    // ---- phase 2: stream keys / values ------------------------------------------------------------
    const int n_kv = (T + AK - 1) / AK;
    load_tile(skv, k_g, QKV_LD, AK, T, tid);
    load_tile(skv + AK * ALD, v_g, QKV_LD, AK, T, tid);
    cp_async_commit();

    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
    float row_max[2] = {-INFINITY, -INFINITY};
    float row_sum[2] = {0.f, 0.f};
    const int qi[2] = {i0 + warp * 16 + gq, i0 + warp * 16 + gq + 8};  // global query index of c0/c1 vs c2/c3
    const float* qt_row[2] = {sqt + (warp * 16 + gq) * QT_LD, sqt + (warp * 16 + gq + 8) * QT_LD};
    constexpr float kLog2e = 1.4426950408889634f;

    for (int kv = 0; kv < n_kv; ++kv) {
        const int j0 = kv * AK;
        if (kv + 1 < n_kv) {
            bf16* nxt = skv + ((kv + 1) & 1) * 2 * AK * ALD;
            load_tile(nxt, k_g + (int64_t)(j0 + AK) * QKV_LD, QKV_LD, AK, T - j0 - AK, tid);
            load_tile(nxt + AK * ALD, v_g + (int64_t)(j0 + AK) * QKV_LD, QKV_LD, AK, T - j0 - AK, tid);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const bf16* sk = skv + (kv & 1) * 2 * AK * ALD;
        const bf16* sv = sk + AK * ALD;

        // S = Q K^T : 16 x 32 per warp
        float s[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t b[4];
                ldmatrix_x4(b, smem_u32(sk + (np * 16 + b_row) * ALD + ks * 16 + b_col));
                const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                mma_16816(s[np * 2], qf[ks], b0);
                mma_16816(s[np * 2 + 1], qf[ks], b1);
            }
        }
        // + relative-position bias, key mask, running max
        float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int j = j0 + n * 8 + tq * 2 + (e & 1);
                int rel = qi[r] - j;
                rel = max(-kMaxRel, min(kMaxRel - 1, rel)) + kMaxRel;
                const float v = j < T ? s[n][e] + qt_row[r][rel] : -INFINITY;
                s[n][e] = v;
                mx[r] = fmaxf(mx[r], v);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], ps[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            corr[r] = exp2f((row_max[r] - mx[r]) * kLog2e);  // first tile: exp2(-inf) = 0
            row_max[r] = mx[r];
        }
        uint32_t pf[2][4];  // P as A-operand fragments for the two 16-key k-steps
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            float p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                p[e] = exp2f((s[n][e] - mx[r]) * kLog2e);
                ps[r] += p[e];
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
        // O += P V : V tile is [key][dim]; ldmatrix.trans yields the (k = key, n = dim) operand
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t b[4];
                const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int dim = dp * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(b, smem_u32(sv + key * ALD + dim));
                const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                mma_16816(o[dp * 2], pf[kk], b0);
                mma_16816(o[dp * 2 + 1], pf[kk], b1);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ---------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float l = row_sum[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        if (qi[r] >= T) continue;
        const float inv = 1.0f / l;
        bf16* orow = ctx + (int64_t)(m.row6 + qi[r]) * kHidden + head * kHeadDim;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            *reinterpret_cast<uint32_t*>(orow + n * 8 + tq * 2) = pack_bf16(o[n][r * 2] * inv, o[n][r * 2 + 1] * inv);
    }
}

}  // namespace

int attention_init() {
    return (int)cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM);
}

int launch_attention(const bf16* qkv, const bf16* pe_k, const UttMeta* meta, int n_utts, int max_t6, bf16* ctx, cudaStream_t s) {
    if (n_utts <= 0 || max_t6 <= 0) return 0;
    dim3 grid((max_t6 + AQ - 1) / AQ, kHeads, n_utts);
    attention_kernel<<<grid, 128, A_SMEM, s>>>(qkv, pe_k, meta, ctx);
    return (int)cudaGetLastError();
}

}  // namespace loco
