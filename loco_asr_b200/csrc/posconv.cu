// Grouped positional convolution of the speech prenet (SpeechT5PositionalConvEmbedding + SamePad,
// HF modeling_speecht5.py:355-397, 445-453): Conv1d(768 -> 768, k = 128, padding 64, groups 16) with the
// weight-norm already folded (W = g * v / ||v||, api.cu), last output frame dropped, + bias, exact GELU.
//
// Per group this is an implicit GEMM  out[T, 48] = X_toeplitz[T, 128*48] * W_g[128*48, 48].  Activations are
// time-major [rows, 768], so one CTA stages the 191-frame x 48-channel window of its 64 output frames in
// shared memory ONCE (zero-filled outside the utterance -- that is the conv's zero padding) and every tap j
// is just the same window shifted down by j rows; the 128 per-tap 48x48 weight blocks stream through a
// double-buffered cp.async pipeline.  Round 1 issues the MMAs through mma.sync (legacy tensor path); the
// stage is 3.3 % of the FLOPs and is scheduled to move to tcgen05 with the same window trick.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int PC_ROWS = 64;              // output frames per CTA
constexpr int PC_WIN = PC_ROWS + kPosK;  // 192 window rows (191 used)
constexpr int PC_LD = 56;                // padded row length (bf16) -> conflict-free ldmatrix
constexpr int PC_TAPS = 8;               // taps per pipeline stage
constexpr int PC_STAGES = 2;
constexpr int PC_W_STAGE = PC_TAPS * kPosGroupCh * PC_LD;  // elements
constexpr int PC_SMEM = (PC_WIN * PC_LD + PC_STAGES * PC_W_STAGE) * 2;

__device__ __forceinline__ void load_w_stage(bf16* sw, const bf16* wg, int chunk, int tid) {
    // wg: this group's weights [128 taps][48 out][48 in]; copy taps chunk*8 .. chunk*8+7
    const bf16* src = wg + (int64_t)chunk * PC_TAPS * kPosGroupCh * kPosGroupCh;
    for (int i = tid; i < PC_TAPS * kPosGroupCh * 6; i += 128) {
        const int row = i / 6, c = i % 6;  // row = tap*48 + out
        cp_async_16(smem_u32(sw + row * PC_LD + c * 8), src + row * kPosGroupCh + c * 8, true);
    }
}

__global__ void __launch_bounds__(128) posconv_kernel(const bf16* __restrict__ h, const bf16* __restrict__ w,
                                                      const float* __restrict__ bias, const UttMeta* __restrict__ meta,
                                                      bf16* __restrict__ pc) {
    pdl_launch_dependents();
    pdl_wait();
    const UttMeta m = meta[blockIdx.z];
    const int f0 = blockIdx.x * PC_ROWS;
    if (f0 >= m.t6) return;
    const int g = blockIdx.y;
    extern __shared__ __align__(16) uint8_t smem[];
    bf16* sx = reinterpret_cast<bf16*>(smem);
    bf16* sw = sx + PC_WIN * PC_LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16* wg = w + (int64_t)g * kPosK * kPosGroupCh * kPosGroupCh;

    // window: frames f0-64 .. f0+127 of this utterance, channels 48g .. 48g+47; zero outside [0, t6)
    for (int i = tid; i < PC_WIN * 6; i += 128) {
        const int wr = i / 6, c = i % 6;
        const int frame = f0 - kPosK / 2 + wr;
        const bool ok = frame >= 0 && frame < m.t6;
        const bf16* src = h + (int64_t)(m.row6 + (ok ? frame : 0)) * kHidden + g * kPosGroupCh + c * 8;
        cp_async_16(smem_u32(sx + wr * PC_LD + c * 8), src, ok);
    }
    load_w_stage(sw, wg, 0, tid);
    cp_async_commit();

    float acc[6][4];
#pragma unroll
    for (int n = 0; n < 6; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;

    constexpr int n_chunks = kPosK / PC_TAPS;  // 16
    const uint32_t sx_base = smem_u32(sx);
    // per-lane ldmatrix offsets
    const int a_row = warp * 16 + (lane & 15);
    const int a_col = (lane >> 4) * 8;
    const int b_row = (lane & 7) + ((lane >> 4) << 3);
    const int b_col = ((lane >> 3) & 1) * 8;

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        if (chunk + 1 < n_chunks) load_w_stage(sw + ((chunk + 1) & 1) * PC_W_STAGE, wg, chunk + 1, tid);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t sw_base = smem_u32(sw + (chunk & 1) * PC_W_STAGE);
#pragma unroll 2
        for (int tap = 0; tap < PC_TAPS; ++tap) {
            const int j = chunk * PC_TAPS + tap;
#pragma unroll
            for (int kc = 0; kc < 3; ++kc) {
                uint32_t a[4];
                ldmatrix_x4(a, sx_base + (uint32_t)(((a_row + j) * PC_LD + kc * 16 + a_col) * 2));
#pragma unroll
                for (int np = 0; np < 3; ++np) {
                    uint32_t b[4];
                    ldmatrix_x4(b, sw_base + (uint32_t)(((tap * kPosGroupCh + np * 16 + b_row) * PC_LD + kc * 16 + b_col) * 2));
                    const uint32_t b0[2] = {b[0], b[1]};
                    const uint32_t b1[2] = {b[2], b[3]};
                    mma_16816(acc[np * 2], a, b0);
                    mma_16816(acc[np * 2 + 1], a, b1);
                }
            }
        }
        __syncthreads();
    }

    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int frame = f0 + warp * 16 + gq + half * 8;
        if (frame >= m.t6) continue;
        bf16* orow = pc + (int64_t)(m.row6 + frame) * kHidden + g * kPosGroupCh;
#pragma unroll
        for (int n = 0; n < 6; ++n) {
            const int col = n * 8 + tq * 2;
            const float v0 = gelu_erf(acc[n][half * 2 + 0] + __ldg(bias + g * kPosGroupCh + col));
            const float v1 = gelu_erf(acc[n][half * 2 + 1] + __ldg(bias + g * kPosGroupCh + col + 1));
            *reinterpret_cast<uint32_t*>(orow + col) = pack_bf16(v0, v1);
        }
    }
}

}  // namespace

int posconv_init() {
    return (int)cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM);
}

int launch_posconv(const bf16* h, const bf16* w, const float* bias, const UttMeta* meta, int n_utts, int max_t6, bf16* pc,
                   cudaStream_t s) {
    if (n_utts <= 0 || max_t6 <= 0) return 0;
    dim3 grid((max_t6 + PC_ROWS - 1) / PC_ROWS, kPosGroups, n_utts);
    return launch_pdl(posconv_kernel, grid, dim3(128), (size_t)PC_SMEM, s, h, w, bias, meta, pc);
}

}  // namespace loco
