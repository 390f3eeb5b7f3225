// Conv layer 0 + GroupNorm + GELU with mma.sync (round-1/2 product kernel, now the LOCO_DEBUG cross-check of conv0_tc.cu;
// "conv0_impl" = 1).  SpeechT5GroupNormConvLayer, HF modeling_speecht5.py:260-281; the GroupNorm scale / shift come from
// the waveform moments (frontend.cu).
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// conv0 + GroupNorm + GELU on the tensor cores.  A 10-tap, 1-input-channel convolution is a [frames, 10] x
// [10, 512] GEMM; done on CUDA cores it cost 10 of the kernel's ~35 FP32 instructions per output and the kernel
// was issue-bound at 8x its HBM floor (ncu r1a: 5.4 ms for a 4.2 GB write).  Here:
//   * the GroupNorm scale is folded into the weights (W'_c = w_c * scale[u, c]; one utterance per CTA) and the shift
//     is the MMA's initial accumulator, so normalisation costs no instruction per output;
//   * fp32 accuracy on bf16 tensor cores through a 3-term split:  x w' ~= x_hi w'_hi + x_lo w'_hi + x_hi w'_lo
//     (K = 3 x 16 with the 10 taps zero-padded to 16), relative error ~2^-16;
//   * per output only the GELU (9 instr + 2 MUFU) and the bf16 pack remain; rows are staged through shared memory
//     so every global store is a full 128-byte line of the time-major [T0, 512] activation.
constexpr int C0_WARPS = 8;
constexpr int C0_FRAMES = 512;               // frames per CTA (4 passes of 8 warps x 16 frames)
constexpr int C0_BLD = 40;                   // padded K (32: w_hi | w_lo) of the smem weight matrix -> conflict-free ldmatrix, 3 CTAs/SM
constexpr int C0_SLD = 72;                   // staging row: 64 channels + 8 pad (bf16)
constexpr int C0_SMEM = kConvDim * C0_BLD * 2 + (C0_FRAMES * 5 + 8) * 4 + kConvDim * 4 + C0_WARPS * 16 * C0_SLD * 2;


__global__ void __launch_bounds__(C0_WARPS * 32) conv0_mma_kernel(const float* __restrict__ wave, const UttMeta* __restrict__ meta,
                                                                   const float* __restrict__ w0, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, bf16* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.y;
    const UttMeta m = meta[u];
    const int slot0 = m.slot6 << 6;
    const int f0 = blockIdx.x * C0_FRAMES;
    if (f0 >= slot0) return;
    extern __shared__ __align__(16) uint8_t smem[];
    bf16* sB = reinterpret_cast<bf16*>(smem);                                   // [512][40]: w_hi taps 0..15, w_lo taps 0..15
    float* xs = reinterpret_cast<float*>(sB + kConvDim * C0_BLD);               // [512*5 + 8]
    float* ssh = xs + C0_FRAMES * 5 + 8;                                         // [512] shift
    bf16* stage = reinterpret_cast<bf16*>(ssh + kConvDim);                      // [8 warps][16][72]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;

    const float* x = wave + m.sample_off;
    for (int i = tid; i < C0_FRAMES * 5 + 8; i += C0_WARPS * 32) {
        const int sidx = f0 * 5 + i;
        xs[i] = sidx < m.n_samples ? __ldg(x + sidx) : 0.f;
    }
    for (int i = tid; i < kConvDim * 16; i += C0_WARPS * 32) {
        const int c = i >> 4, k = i & 15;
        bf16 hi = __float2bfloat16_rn(0.f), lo = hi;
        if (k < 10) split_bf16(__ldg(w0 + c * 10 + k) * scale[(int64_t)u * kConvDim + c], hi, lo);
        sB[c * C0_BLD + k] = hi;         // pairs with x_hi, then with x_lo
        sB[c * C0_BLD + 16 + k] = lo;    // pairs with x_hi
    }
    for (int i = tid; i < kConvDim; i += C0_WARPS * 32) ssh[i] = shift[(int64_t)u * kConvDim + i];
    __syncthreads();

    const int b_row = (lane & 7) + ((lane >> 4) << 3);
    const int b_col = ((lane >> 3) & 1) * 8;
    bf16* my_stage = stage + warp * 16 * C0_SLD;
    bf16* obase = out + ((int64_t)m.row6 << 6) * kConvDim;

#pragma unroll 1
    for (int pass = 0; pass < C0_FRAMES / (C0_WARPS * 16); ++pass) {
        const int fl = (pass * C0_WARPS + warp) * 16;   // first local frame of this warp's 16-frame block
        if (f0 + fl >= slot0) break;
        // A fragments: rows = frames fl+g / fl+g+8, k = taps (zero beyond 10)
        uint32_t a_hi[4], a_lo[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float* xr = xs + (fl + gq + r * 8) * 5;
            bf16 h0, l0, h1, l1;
            split_bf16(xr[tq * 2], h0, l0);
            split_bf16(xr[tq * 2 + 1], h1, l1);
            a_hi[r] = pack2(h0, h1);
            a_lo[r] = pack2(l0, l1);
            if (tq == 0) {
                split_bf16(xr[8], h0, l0);
                split_bf16(xr[9], h1, l1);
                a_hi[2 + r] = pack2(h0, h1);
                a_lo[2 + r] = pack2(l0, l1);
            } else {
                a_hi[2 + r] = 0u;
                a_lo[2 + r] = 0u;
            }
        }
#pragma unroll 1
        for (int cg = 0; cg < kConvDim / 64; ++cg) {   // 64 channels (8 n-tiles) at a time
            float acc[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float2 sh = *reinterpret_cast<const float2*>(ssh + cg * 64 + n * 8 + tq * 2);
                acc[n][0] = sh.x; acc[n][1] = sh.y; acc[n][2] = sh.x; acc[n][3] = sh.y;
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                    uint32_t b[4];
                    ldmatrix_x4(b, smem_u32(sB + (cg * 64 + np * 16 + b_row) * C0_BLD + (ks >> 1) * 16 + b_col));   // ks 0,1: w_hi; 2: w_lo
                    const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
                    if (ks == 1) {
                        mma_16816(acc[np * 2], a_lo, b0);
                        mma_16816(acc[np * 2 + 1], a_lo, b1);
                    } else {
                        mma_16816(acc[np * 2], a_hi, b0);
                        mma_16816(acc[np * 2 + 1], a_hi, b1);
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float2 g0 = gelu_erf2(make_float2(acc[n][0], acc[n][1])), g1 = gelu_erf2(make_float2(acc[n][2], acc[n][3]));
                *reinterpret_cast<uint32_t*>(my_stage + gq * C0_SLD + n * 8 + tq * 2) = pack_bf16(g0.x, g0.y);
                *reinterpret_cast<uint32_t*>(my_stage + (gq + 8) * C0_SLD + n * 8 + tq * 2) = pack_bf16(g1.x, g1.y);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // 16 rows x 8 chunks of 16 B = 128 chunks, 4 per lane
                const int idx = j * 32 + lane;
                const int row = idx >> 3, ch = idx & 7;
                const int f = f0 + fl + row;
                if (f < slot0) {
                    uint4 v = *reinterpret_cast<const uint4*>(my_stage + row * C0_SLD + ch * 8);
                    if (f >= m.t0) v = make_uint4(0u, 0u, 0u, 0u);   // slot padding frames
                    *reinterpret_cast<uint4*>(obase + (int64_t)f * kConvDim + cg * 64 + ch * 8) = v;
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

int conv0_mma_init() {
    return (int)cudaFuncSetAttribute(conv0_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM);
}

int launch_conv0(const float* wave, const UttMeta* meta, int n_utts, int max_slot0, const float* w0, const float* scale,
                 const float* shift, bf16* out, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    dim3 grid((max_slot0 + C0_FRAMES - 1) / C0_FRAMES, n_utts);
    return launch_pdl(conv0_mma_kernel, grid, dim3(C0_WARPS * 32), (size_t)C0_SMEM, s, wave, meta, w0, scale, shift, out);
}

}  // namespace loco
