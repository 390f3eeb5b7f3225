// Grouped positional convolution on tcgen05, polyphase form (product path; posconv_tc.cu keeps the one-phase kernel and
// posconv.cu the mma.sync one as debug cross-checks).  SpeechT5PositionalConvEmbedding + SamePad,
// HF modeling_speecht5.py:355-397, 445-453: Conv1d(768 -> 768, k = 128, padding 64, groups 16), weight-norm folded at
// load, last frame dropped, + bias, GELU.
//
// Why another form.  Per group the conv is out[T, 48] = X_toeplitz[T, 128*48] * W_g[128*48, 48]; with one output frame per
// accumulator row the MMAs are M = 128, N = 48, K = 16, and such an MMA reads 5.6 KB of shared-memory operands for 24
// tensor-pipe cycles of work: the one-phase kernel ran at the shared-memory rate, under half the tensor rate (ncu r01g/r03j:
// tensor pipe 46-48 % active).  Here an accumulator row is FOUR consecutive output frames:
//     out[4s + p] = sum_tap X[4s + p + tap - 64] w[tap]        p = 0..3
//                 = sum_{j = 0..130} X[4s + j - 64] w[j - p]    (w[.] = 0 outside 0..127)
// so step j multiplies the SAME A operand (row s = frame 4s + j - 64) by the four taps j, j-1, j-2, j-3 side by side:
// M = 128, N = 192, K = 16 -- 10 KB of operands for 96 tensor-pipe cycles.  131 steps x 3 K-steps cover 512 output
// frames (the one-phase kernel needed 4 x 128 x 3 MMAs of a quarter the size at ~44 cycles each).
//
//   * A operand: frames 4s + j = 4(s + j/4) + j%4, so the window is staged de-interleaved by frame phase -- four arrays
//     x_b[r] = X[4r + b], each in the UMMA no-swizzle K-major layout with all rows 16 B apart; "step j" is phase array j%4
//     with the descriptor start advanced by (j/4) * 16 bytes.
//   * B operand: the group's weights live in global memory as [in/8][tap + 3 (zero taps at both ends)][out][in%8]; the 192
//     B rows of step j are the 4 consecutive taps j-3 .. j (column block q holds output phase 3 - q), i.e. a sliding
//     4-tap window over one array -- again only the descriptor start moves (768 B per step).  Weights stream through a
//     2-stage ring of 8 steps (+3 overlap taps) filled by 1-D bulk TMA copies.
//   * Utterances of any length share tiles: the batch is laid on a virtual TIMELINE, utterance after utterance with 64
//     zero frames between neighbours (the conv's zero padding for both), and an item is 512 consecutive timeline
//     frames of one group.  `vmap` (built with the plan) gives, per timeline frame, the row of the [R6, 768] buffers or
//     -1.  A result depends only on the utterance's own frames and the zeros around them, never on where in a tile it
//     sits: every column block and accumulator row sees the same K order (zero weight taps add exact zeros).
//   * Persistent, warp-specialised, one CTA per SM: weight loader warp, MMA issuer warp, four stager warps (cp.async of
//     the next item's window into the other A buffer), four epilogue warps (TMEM -> bias + GELU -> bf16 rows) on the
//     other accumulator -- staging and epilogue run under the MMAs of the neighbouring items.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int PP_PHASES = 4;
constexpr int PP_ROWS = 128;                                    // UMMA M: accumulator rows
constexpr int PP_TILE = PP_PHASES * PP_ROWS;                    // 512 timeline frames per item
constexpr int PP_N = PP_PHASES * kPosGroupCh;                   // 192
constexpr int PP_STEPS = kPosK + PP_PHASES - 1;                 // 131
constexpr int PP_AROWS = PP_ROWS + kPosK / PP_PHASES;           // 160 rows per phase array (s + j/4, j/4 <= 32)
constexpr int PP_WIN = PP_AROWS * PP_PHASES;                    // 640 staged frames: timeline tile start - 64 .. + 575
constexpr int PP_KC = kPosGroupCh / 8;                          // 6 sixteen-byte channel chunks
constexpr int PP_A_KC_BYTES = PP_AROWS * 16;                    // 2560
constexpr int PP_A_PHASE_BYTES = PP_KC * PP_A_KC_BYTES;         // 15360
constexpr int PP_A_BYTES = PP_PHASES * PP_A_PHASE_BYTES;        // 61440
constexpr int PP_STAGE_STEPS = 8;
constexpr int PP_STAGE_TAPS = PP_STAGE_STEPS + PP_PHASES - 1;   // 11
constexpr int PP_TAP_BYTES = kPosGroupCh * 16;                  // 768: [48 out][8 in] of one tap and channel chunk
constexpr int PP_W_KC_BYTES = PP_STAGE_TAPS * PP_TAP_BYTES;     // 8448
constexpr int PP_W_STAGE_BYTES = PP_KC * PP_W_KC_BYTES;         // 50688
constexpr int PP_W_STAGES = 2;
constexpr int PP_N_WSTAGES = (PP_STEPS + PP_STAGE_STEPS - 1) / PP_STAGE_STEPS;   // 17
static_assert(kPosPPTaps == PP_N_WSTAGES * PP_STAGE_STEPS + PP_PHASES - 1, "global weight array: taps per (group, chunk)");
static_assert(kPosPPTile == PP_TILE && kPosPPHalo == kPosK / 2, "timeline geometry shared with api.cu");
constexpr int PP_TMEM_COLS = 512;                               // 2 accumulators x 256-column slots (192 used)
constexpr int PP_THREADS = 384;       // warp 0 weights, warp 1 MMA, warps 2-3 idle, warps 4-7 epilogue, warps 8-11 stagers
constexpr int PP_SMEM = 2 * PP_A_BYTES + PP_W_STAGES * PP_W_STAGE_BYTES + 128 + 128;

struct __align__(8) PpBars {
    uint64_t w_full[PP_W_STAGES], w_empty[PP_W_STAGES];
    uint64_t a_full[2], a_empty[2];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(PP_THREADS, 1)
posconv_pp_kernel(const bf16* __restrict__ h, const bf16* __restrict__ w, const float* __restrict__ bias,
                  const int32_t* __restrict__ vmap, int n_items, bf16* __restrict__ pc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t a_base = smem_base;
    const uint32_t ring_base = smem_base + 2 * PP_A_BYTES;
    PpBars* bars = reinterpret_cast<PpBars*>(smem_al + 2 * PP_A_BYTES + PP_W_STAGES * PP_W_STAGE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < PP_W_STAGES; ++s) {
            mbar_init(smem_u32(&bars->w_full[s]), 1);
            mbar_init(smem_u32(&bars->w_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->a_full[b]), 128);
            mbar_init(smem_u32(&bars->a_empty[b]), 1);
            mbar_init(smem_u32(&bars->acc_full[b]), 1);
            mbar_init(smem_u32(&bars->acc_empty[b]), 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), PP_TMEM_COLS);
    pdl_launch_dependents();
    pdl_wait();          // barrier init / TMEM allocation overlap the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== weight loader: the item's group, 17 stages of 11 taps x 6 chunks =====================
        if (lane == 0) {
            uint32_t c = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int g = item % kPosGroups;
                const uint8_t* wg = reinterpret_cast<const uint8_t*>(w) + (size_t)g * PP_KC * kPosPPTaps * PP_TAP_BYTES;
                for (int st = 0; st < PP_N_WSTAGES; ++st, ++c) {
                    const uint32_t s = c % PP_W_STAGES, ph = (c / PP_W_STAGES) & 1u;
                    mbar_wait(smem_u32(&bars->w_empty[s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bars->w_full[s]);
                    mbar_arrive_expect_tx(full, PP_W_STAGE_BYTES);
#pragma unroll
                    for (int kc = 0; kc < PP_KC; ++kc)
                        bulk_load_1d(ring_base + s * PP_W_STAGE_BYTES + kc * PP_W_KC_BYTES,
                                     wg + ((size_t)kc * kPosPPTaps + (size_t)st * PP_STAGE_STEPS) * PP_TAP_BYTES, PP_W_KC_BYTES, full);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: warp-uniform control flow, one elected lane issues =====================
        constexpr uint32_t idesc = umma_idesc_bf16(PP_ROWS, PP_N);
        constexpr uint64_t kAStepK = (uint64_t)(2 * PP_A_KC_BYTES / 16);        // two channel chunks = one K = 16 step
        constexpr uint64_t kBStepK = (uint64_t)(2 * PP_W_KC_BYTES / 16);
        constexpr uint64_t kAPhase = (uint64_t)(PP_A_PHASE_BYTES / 16);
        constexpr uint64_t kBStep = (uint64_t)(PP_TAP_BYTES / 16);
        const uint64_t da0 = umma_desc_noswizzle_kmajor(a_base, PP_A_KC_BYTES, 128);
        const uint64_t db0 = umma_desc_noswizzle_kmajor(ring_base, PP_W_KC_BYTES, 128);
        uint32_t c = 0;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
            const int ab = n & 1;
            const uint32_t use = (uint32_t)(n >> 1) & 1u;
            mbar_wait(smem_u32(&bars->a_full[ab]), use);
            mbar_wait(smem_u32(&bars->acc_empty[ab]), use ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + ab * 256;
            const uint64_t da_item = da0 + (uint64_t)(ab * (PP_A_BYTES / 16));
            for (int st = 0; st < PP_N_WSTAGES; ++st, ++c) {
                const uint32_t s = c % PP_W_STAGES, ph = (c / PP_W_STAGES) & 1u;
                mbar_wait(smem_u32(&bars->w_full[s]), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t db_s = db0 + (uint64_t)(s * (PP_W_STAGE_BYTES / 16));
                    const uint64_t da_s = da_item + (uint64_t)(st * (PP_STAGE_STEPS / PP_PHASES));
                    const int n_steps = min(PP_STAGE_STEPS, PP_STEPS - st * PP_STAGE_STEPS);
#pragma unroll
                    for (int jl = 0; jl < PP_STAGE_STEPS; ++jl) {
                        if (jl < n_steps) {
                            const uint64_t da = da_s + (uint64_t)(jl & 3) * kAPhase + (uint64_t)(jl >> 2);
                            const uint64_t db = db_s + (uint64_t)jl * kBStep;
                            umma_bf16(d_tmem, da, db, idesc, (st | jl) != 0 ? 1u : 0u);
                            umma_bf16(d_tmem, da + kAStepK, db + kBStepK, idesc, 1u);
                            umma_bf16(d_tmem, da + 2 * kAStepK, db + 2 * kBStepK, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&bars->w_empty[s]));
                    if (st == PP_N_WSTAGES - 1) {
                        umma_commit(smem_u32(&bars->acc_full[ab]));
                        umma_commit(smem_u32(&bars->a_empty[ab]));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        // ===================== stagers: the item's 640-frame window, de-interleaved by frame phase =====================
        const int st_tid = tid - 256;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
            const int ab = n & 1;
            const uint32_t use = (uint32_t)(n >> 1) & 1u;
            const int vt = item / kPosGroups, g = item % kPosGroups;
            mbar_wait(smem_u32(&bars->a_empty[ab]), use ^ 1u);
            const int32_t* vm = vmap + (int64_t)vt * PP_TILE;
            const bf16* hg = h + g * kPosGroupCh;
            const uint32_t dst0 = a_base + ab * PP_A_BYTES;
#pragma unroll
            for (int i = 0; i < PP_WIN / 128; ++i) {
                const int t = st_tid + i * 128;
                const int r = __ldg(vm + t);
                const bf16* src = hg + (int64_t)(r >= 0 ? r : 0) * kHidden;
                const uint32_t dst = dst0 + (t & 3) * PP_A_PHASE_BYTES + (t >> 2) * 16;
#pragma unroll
                for (int kc = 0; kc < PP_KC; ++kc) cp_async_16(dst + kc * PP_A_KC_BYTES, src + kc * 8, r >= 0);
            }
            cp_async_commit();
            cp_async_wait<0>();
            fence_proxy_async_smem();     // generic-proxy writes (cp.async) -> visible to the tensor core's async proxy
            mbar_arrive(smem_u32(&bars->a_full[ab]));
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> bias + GELU -> bf16 rows of pc =====================
        const int q4 = warp & 3;
        const int s_row = q4 * 32 + lane;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
            const int ab = n & 1;
            const uint32_t use = (uint32_t)(n >> 1) & 1u;
            const int vt = item / kPosGroups, g = item % kPosGroups;
            const int32_t* vm = vmap + (int64_t)vt * PP_TILE + kPosPPHalo + 4 * s_row;
            int rows[PP_PHASES];
#pragma unroll
            for (int p = 0; p < PP_PHASES; ++p) rows[p] = __ldg(vm + p);
            mbar_wait(smem_u32(&bars->acc_full[ab]), use);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ab * 256);
#pragma unroll
            for (int q = 0; q < PP_PHASES; ++q) {
                const int r = rows[PP_PHASES - 1 - q];            // column block q holds output phase 3 - q
                bf16* orow = pc + (int64_t)(r >= 0 ? r : 0) * kHidden + g * kPosGroupCh;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_row + (uint32_t)(q * kPosGroupCh + c * 16), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(v[i]));
                    if (r >= 0) {
#pragma unroll
                        for (int j8 = 0; j8 < 16; j8 += 8) {
                            float2 f[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + g * kPosGroupCh + c * 16 + j8 + 2 * e));
                                f[e] = gelu_erf2(add_f32x2(make_float2(__uint_as_float(v[j8 + 2 * e]), __uint_as_float(v[j8 + 2 * e + 1])), b2));
                            }
                            uint4 o4;
                            o4.x = pack_bf16(f[0].x, f[0].y);
                            o4.y = pack_bf16(f[1].x, f[1].y);
                            o4.z = pack_bf16(f[2].x, f[2].y);
                            o4.w = pack_bf16(f[3].x, f[3].y);
                            *reinterpret_cast<uint4*>(orow + c * 16 + j8) = o4;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, PP_TMEM_COLS);
    }
}

}  // namespace

int posconv_pp_init() {
    return (int)cudaFuncSetAttribute(posconv_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM);
}

int launch_posconv_pp(const bf16* h, const bf16* w_pp, const float* bias, const int32_t* vmap, int n_vtiles, bf16* pc, int num_sms,
                      cudaStream_t s) {
    if (n_vtiles <= 0) return 0;
    const int n_items = n_vtiles * kPosGroups;
    const int grid = n_items < num_sms ? n_items : num_sms;
    return launch_pdl(posconv_pp_kernel, dim3(grid), dim3(PP_THREADS), (size_t)PP_SMEM, s, h, w_pp, bias, vmap, n_items, pc);
}

}  // namespace loco
