// Grouped positional convolution on tcgen05, one output frame per accumulator row (the product kernel until the polyphase
// form of posconv_pp.cu replaced it; now LOCO_DEBUG builds only, "posconv_impl" = 2, a cross-check with the same summation
// order).  SpeechT5PositionalConvEmbedding + SamePad, HF modeling_speecht5.py:355-397, 445-453:
// Conv1d(768 -> 768, k = 128, padding 64, groups 16), weight-norm folded at load, last frame dropped, + bias, GELU.
//
// Per group: out[T, 48] = X_toeplitz[T, 128*48] * W_g[128*48, 48].  The Toeplitz operand is never built.  A CTA
// stages, for each of up to two 128-frame output tiles (any utterances), the 255-frame x 48-channel window of
// its inputs in shared memory in the UMMA *no-swizzle K-major* core-matrix layout with all rows 16 B apart:
//     window[kc][row] (16 B = 8 channels),  kc = 0..5,  row = 0..255
// so "the operand of tap j" is the same window with the descriptor start address advanced by j * 16 bytes -- 128
// taps x 3 K-steps of tcgen05.mma (M = 128, N = 48, K = 16) read it in place, accumulating in TMEM (2 tiles x 48
// columns).  Frames outside the utterance are zero rows of the window: that is the conv's zero padding, and it is
// why tiles never straddle utterances.  The group's 590 KB of weights stream ONCE per CTA (taps outer, tiles
// inner) through a 3-stage ring filled by 1-D bulk TMA copies.  Two tiles per CTA and 105 KB of shared memory put TWO CTAs on an
// SM, so one CTA's window staging and epilogue run under the other's MMAs (four tiles per CTA and one CTA per SM halved the
// weight traffic but left those phases exposed: 1.22 -> 1.05 ms per step).
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

#ifndef PT_TILES_CFG
#define PT_TILES_CFG 2
#endif
constexpr int PT_TILES = PT_TILES_CFG;            // output tiles per CTA (2: two CTAs per SM, one stages / drains while the other's MMAs run)
constexpr int PT_ROWS = 128;                      // frames per tile (UMMA M)
constexpr int PT_WROWS = 256;                     // window rows (255 used)
constexpr int PT_KC = kPosGroupCh / 8;            // 6 sixteen-byte channel chunks
constexpr int PT_WIN_BYTES = PT_KC * PT_WROWS * 16;          // 24576
constexpr int PT_TAP_BYTES = PT_KC * kPosGroupCh * 16;       // 4608: [kc][out 48][8 in]
constexpr int PT_STAGE_TAPS = PT_TILES == 2 ? 4 : 8;
constexpr int PT_STAGE_BYTES = PT_STAGE_TAPS * PT_TAP_BYTES; // 36864
constexpr int PT_STAGES = 3;
constexpr int PT_N_ITERS = kPosK / PT_STAGE_TAPS;            // 16
constexpr int PT_TMEM_COLS = PT_TILES * 64;        // 64-column slot per tile (48 used)
constexpr int PT_THREADS = 192;                   // warp 0 weights, warp 1 MMA, warps 2-5 epilogue
constexpr int PT_SMEM = PT_TILES * PT_WIN_BYTES + PT_STAGES * PT_STAGE_BYTES + 128 + 128;

struct __align__(8) PtBars {
    uint64_t full[PT_STAGES];
    uint64_t empty[PT_STAGES];
    uint64_t done;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(PT_THREADS, PT_TILES == 2 ? 2 : 1)
posconv_tc_kernel(const bf16* __restrict__ h, const bf16* __restrict__ w, const float* __restrict__ bias,
                  const PcTile* __restrict__ tiles, int n_tiles_total, bf16* __restrict__ pc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t win_base = smem_base;
    const uint32_t ring_base = smem_base + PT_TILES * PT_WIN_BYTES;
    PtBars* bars = reinterpret_cast<PtBars*>(smem_al + PT_TILES * PT_WIN_BYTES + PT_STAGES * PT_STAGE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const int tile0 = blockIdx.x * PT_TILES;
    const int my_tiles = min(PT_TILES, n_tiles_total - tile0);

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        mbar_init(smem_u32(&bars->done), 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), PT_TMEM_COLS);
    pdl_launch_dependents();
    pdl_wait();          // barrier init / TMEM allocation overlap the previous kernel's tail

    // ---- stage the input windows (all threads): frames f0-64 .. f0+190, zero outside [0, t6) ----------------
    for (int t = 0; t < my_tiles; ++t) {
        const PcTile tl = tiles[tile0 + t];
        const bf16* src0 = h + (int64_t)(tl.row0 - tl.f0) * kHidden + g * kPosGroupCh;   // frame 0 of the utterance
        for (int i = tid; i < PT_WROWS * PT_KC; i += PT_THREADS) {
            const int wr = i / PT_KC, kc = i - wr * PT_KC;
            const int frame = tl.f0 - kPosK / 2 + wr;
            const bool ok = frame >= 0 && frame < tl.t6;
            cp_async_16(win_base + t * PT_WIN_BYTES + kc * (PT_WROWS * 16) + wr * 16,
                        src0 + (int64_t)(ok ? frame : 0) * kHidden + kc * 8, ok);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();     // generic-proxy writes (cp.async) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const bf16* wg = w + (int64_t)g * kPosK * kPosGroupCh * kPosGroupCh;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < PT_N_ITERS; ++it) {
                const int s = it % PT_STAGES;
                const uint32_t ph = (uint32_t)(it / PT_STAGES) & 1u;
                mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
                const uint32_t full = smem_u32(&bars->full[s]);
                mbar_arrive_expect_tx(full, PT_STAGE_BYTES);
                bulk_load_1d(ring_base + s * PT_STAGE_BYTES, reinterpret_cast<const uint8_t*>(wg) + (size_t)it * PT_STAGE_BYTES,
                             PT_STAGE_BYTES, full);
            }
        }
    } else if (warp == 1) {
        // MMA issuer.  Each MMA is only 128 x 48 x 16 (24 tensor-pipe cycles), so the issue path must be shorter than that:
        // the whole warp runs the uniform control flow (descriptors stay in uniform registers; a single divergent lane paid
        // an R2UR per operand and ~86 cycles per MMA -- 3.6x the tensor time, ncu r01c), one elected lane issues, and the
        // descriptors are base values plus constants: tap j is +j, a K-step is +2 * 256 (A) or +2 * 48 (B) sixteen-byte units.
        constexpr uint32_t idesc = umma_idesc_bf16(PT_ROWS, kPosGroupCh);
        constexpr uint64_t kAStepK = (uint64_t)(2 * PT_WROWS);             // 2 core-matrix columns of the window, in 16 B units
        constexpr uint64_t kBStepK = (uint64_t)(2 * kPosGroupCh);
        constexpr uint64_t kBStepTap = (uint64_t)(PT_TAP_BYTES / 16);
        const uint64_t da0 = umma_desc_noswizzle_kmajor(win_base, PT_WROWS * 16, 128);
        const uint64_t db0 = umma_desc_noswizzle_kmajor(ring_base, kPosGroupCh * 16, 128);
        for (int it = 0; it < PT_N_ITERS; ++it) {
            const int s = it % PT_STAGES;
            const uint32_t ph = (uint32_t)(it / PT_STAGES) & 1u;
            mbar_wait(smem_u32(&bars->full[s]), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t db_s = db0 + (uint64_t)(s * (PT_STAGE_BYTES / 16));
                const uint64_t da_it = da0 + (uint64_t)(it * PT_STAGE_TAPS);
#pragma unroll
                for (int tap = 0; tap < PT_STAGE_TAPS; ++tap) {
                    const uint64_t db = db_s + tap * kBStepTap;
                    for (int t = 0; t < my_tiles; ++t) {
                        const uint64_t da = da_it + (uint64_t)(t * (PT_WIN_BYTES / 16) + tap);
                        const uint32_t acc = (it | tap) != 0 ? 1u : 0u;
                        umma_bf16(tmem_base + t * 64, da, db, idesc, acc);
                        umma_bf16(tmem_base + t * 64, da + kAStepK, db + kBStepK, idesc, 1u);
                        umma_bf16(tmem_base + t * 64, da + 2 * kAStepK, db + 2 * kBStepK, idesc, 1u);
                    }
                }
                umma_commit(smem_u32(&bars->empty[s]));
                if (it == PT_N_ITERS - 1) umma_commit(smem_u32(&bars->done));
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: TMEM -> bias + GELU -> bf16 rows of pc ---------------------------------------------------
        const int q = warp & 3;
        mbar_wait(smem_u32(&bars->done), 0);
        tc_fence_after();
        for (int t = 0; t < my_tiles; ++t) {
            const PcTile tl = tiles[tile0 + t];
            const int row = q * 32 + lane;
            const bool ok = tl.f0 + row < tl.t6;
            bf16* orow = pc + (int64_t)(tl.row0 + row) * kHidden + g * kPosGroupCh;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t v[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 64 + c * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(v[i]));
                if (ok) {
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, PT_TMEM_COLS);
    }
}

}  // namespace

int posconv_tc_init() {
    return (int)cudaFuncSetAttribute(posconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM);
}

int launch_posconv_tc(const bf16* h, const bf16* w_tc, const float* bias, const PcTile* tiles, int n_tiles, bf16* pc, cudaStream_t s) {
    if (n_tiles <= 0) return 0;
    dim3 grid((n_tiles + PT_TILES - 1) / PT_TILES, kPosGroups);
    return launch_pdl(posconv_tc_kernel, grid, dim3(PT_THREADS), (size_t)PT_SMEM, s, h, w_tc, bias, tiles, n_tiles, pc);
}

}  // namespace loco
