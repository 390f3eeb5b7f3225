// Conv layer 0 of the wav2vec2-style feature encoder with its per-channel GroupNorm and GELU
// (SpeechT5GroupNormConvLayer, HF modeling_speecht5.py:260-281): Conv1d(1 -> 512, k = 10, s = 5, no bias),
// GroupNorm(num_groups = 512) == per-(utterance, channel) mean / biased variance over all T0 frames, eps 1e-5,
// affine, exact GELU.
//
// GroupNorm is a global reduction over time sitting between the conv and the GELU.  Because the conv is
// linear with 10 taps, its per-channel moments follow from 10 + 55 moments of the *waveform*:
//     mean_c  = sum_k w_c[k] m_k,                 m_k      = mean_t x[5t + k]
//     E[y_c^2] = sum_{k,k'} w_c[k] w_c[k'] R_kk',  R_{kk'}  = mean_t x[5t + k] x[5t + k']
// so one cheap pass over the waveform (fp32 per-thread partials, fp64 reduction) replaces a full extra pass
// over the 512 x T0 conv output, and conv0 -> GN -> GELU is then a single streaming kernel (conv0_tc.cu; conv0_mma.cu is
// the mma.sync cross-check) that writes the bf16 time-major [T0, 512] activation exactly once.  gn_finalize_kernel also
// folds the GroupNorm scale / shift into the per-utterance weight operand conv0_tc_kernel consumes.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int kStatFrames = 2048;  // frames per stats block
constexpr int kNumMoments = 65;    // 10 first + 55 second moments

// The chunk's samples are staged in shared memory with coalesced loads first: read straight from global, every thread's ten
// taps were a 40-byte access at a 20-byte stride and the kernel ran at 0.6 TB/s, one CTA per SM (156 registers).  Frames keep
// their thread (f0 + tid + 256 i) and the reduction its order, so the moments are bit-identical to the previous version.
__global__ void __launch_bounds__(256, 2) wave_moments_kernel(const float* __restrict__ wave, const UttMeta* __restrict__ meta,
                                                               double* __restrict__ partial, int chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.y, chunk = blockIdx.x;
    const UttMeta m = meta[u];
    const int f0 = chunk * kStatFrames;
    double* out = partial + ((int64_t)u * chunks + chunk) * kNumMoments;
    if (f0 >= m.t0) {
        if (threadIdx.x < kNumMoments) out[threadIdx.x] = 0.0;
        return;
    }
    const int f1 = min(f0 + kStatFrames, m.t0);
    const float* x = wave + m.sample_off + (int64_t)f0 * 5;
    __shared__ float xs[kStatFrames * 5 + 5];
    const int n_s = (f1 - f0) * 5 + 5;              // the last frame's ten taps end at sample 5 (f1 - 1) + 9 < n_samples
    for (int i = threadIdx.x; i < n_s; i += 256) xs[i] = __ldg(x + i);
    __syncthreads();
    float acc[kNumMoments];
#pragma unroll
    for (int i = 0; i < kNumMoments; ++i) acc[i] = 0.f;
    for (int f = threadIdx.x; f < f1 - f0; f += 256) {
        float v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = xs[5 * f + k];     // stride of 5 words between lanes: conflict-free
        int idx = 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            acc[k] += v[k];
#pragma unroll
            for (int k2 = k; k2 < 10; ++k2) {
                acc[idx] = fmaf(v[k], v[k2], acc[idx]);
                ++idx;
            }
        }
    }
    __shared__ double red[8][kNumMoments];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kNumMoments; ++i) {
        double d = warp_sum((double)acc[i]);
        if (lane == 0) red[warp][i] = d;
    }
    __syncthreads();
    if (threadIdx.x < kNumMoments) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += red[w][threadIdx.x];
        out[threadIdx.x] = d;
    }
}

__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// scale[u][c] = gamma_c / sqrt(var + eps), shift[u][c] = beta_c - mean * scale
__global__ void __launch_bounds__(512) gn_finalize_kernel(const double* __restrict__ partial, int chunks,
                                                           const UttMeta* __restrict__ meta, const float* __restrict__ w0,
                                                           const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                                                           float* __restrict__ scale, float* __restrict__ shift, bf16* __restrict__ wfold) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.x;
    __shared__ double mom[kNumMoments];
    if (threadIdx.x < kNumMoments) {
        double d = 0.0;
        for (int c = 0; c < chunks; ++c) d += partial[((int64_t)u * chunks + c) * kNumMoments + threadIdx.x];
        const int t0 = meta[u].t0;
        mom[threadIdx.x] = t0 > 0 ? d / (double)t0 : 0.0;
    }
    __syncthreads();
    const int c = threadIdx.x;
    double w[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) w[k] = (double)w0[c * 10 + k];
    double mean = 0.0, ex2 = 0.0;
    int idx = 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        mean += w[k] * mom[k];
#pragma unroll
        for (int k2 = k; k2 < 10; ++k2) {
            const double r = mom[idx++];
            ex2 += (k2 == k ? 1.0 : 2.0) * w[k] * w[k2] * r;
        }
    }
    double var = ex2 - mean * mean;
    if (var < 0.0) var = 0.0;
    const double sc = (double)gn_w[c] / sqrt(var + (double)kLnEps);
    const float scf = (float)sc, shf = (float)((double)gn_b[c] - mean * sc);
    scale[(int64_t)u * kConvDim + c] = scf;
    shift[(int64_t)u * kConvDim + c] = shf;
    if (wfold != nullptr) {
        // The utterance's B operand for conv0_tc_kernel (conv0_tc.cu): channel c's row of the K = 48 split GEMM, six 16-byte chunks
        // [w'_hi 0-7 | w'_hi 8 9, shift_hi, 0.. | w'_hi 0-7 | w'_hi 8 9, 0.. | w'_lo 0-7 | w'_lo 8 9, shift_lo, 0..], w' = w * scale,
        // stored [chunk][channel][8] = the UMMA no-swizzle K-major layout.
        uint32_t hi[6], lo[6];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            bf16 h0, l0, h1, l1;
            split_bf16(w0[c * 10 + 2 * k] * scf, h0, l0);
            split_bf16(w0[c * 10 + 2 * k + 1] * scf, h1, l1);
            hi[k] = pack2(h0, h1);
            lo[k] = pack2(l0, l1);
        }
        bf16 sh_hi, sh_lo;
        split_bf16(shf, sh_hi, sh_lo);
        hi[5] = (uint32_t)__bfloat16_as_ushort(sh_hi);
        lo[5] = (uint32_t)__bfloat16_as_ushort(sh_lo);
        uint4* dst = reinterpret_cast<uint4*>(wfold + (int64_t)u * (kConv0FoldBytes / 2)) + c;
        const uint4 c0 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[0 * kConvDim] = c0;
        dst[1 * kConvDim] = make_uint4(hi[4], hi[5], 0u, 0u);
        dst[2 * kConvDim] = c0;
        dst[3 * kConvDim] = make_uint4(hi[4], 0u, 0u, 0u);
        dst[4 * kConvDim] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dst[5 * kConvDim] = make_uint4(lo[4], lo[5], 0u, 0u);
    }
}

}  // namespace

int wave_stats_chunks(int max_t0) { return max_t0 <= 0 ? 1 : (max_t0 + kStatFrames - 1) / kStatFrames; }

int launch_wave_stats(const float* wave, const UttMeta* meta, int n_utts, int chunks, const float* w0, const float* gn_w,
                      const float* gn_b, double* partial, float* scale, float* shift, bf16* wfold, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    int rc = launch_pdl(wave_moments_kernel, dim3(chunks, n_utts), dim3(256), 0, s, wave, meta, partial, chunks);
    if (rc) return rc;
    return launch_pdl(gn_finalize_kernel, dim3(n_utts), dim3(512), 0, s, partial, chunks, meta, w0, gn_w, gn_b, scale, shift, wfold);
}

}  // namespace loco
