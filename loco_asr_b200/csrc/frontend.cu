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
// over the 512 x T0 conv output, and conv0 -> GN -> GELU is then a single streaming kernel that writes the
// bf16 time-major [T0, 512] activation exactly once.
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

constexpr int kStatFrames = 2048;  // frames per stats block
constexpr int kNumMoments = 65;    // 10 first + 55 second moments

__global__ void __launch_bounds__(256) wave_moments_kernel(const float* __restrict__ wave, const UttMeta* __restrict__ meta,
                                                            double* __restrict__ partial, int chunks) {
    const int u = blockIdx.y, chunk = blockIdx.x;
    const UttMeta m = meta[u];
    const int f0 = chunk * kStatFrames;
    double* out = partial + ((int64_t)u * chunks + chunk) * kNumMoments;
    if (f0 >= m.t0) {
        if (threadIdx.x < kNumMoments) out[threadIdx.x] = 0.0;
        return;
    }
    const int f1 = min(f0 + kStatFrames, m.t0);
    const float* x = wave + m.sample_off;
    float acc[kNumMoments];
#pragma unroll
    for (int i = 0; i < kNumMoments; ++i) acc[i] = 0.f;
    for (int f = f0 + threadIdx.x; f < f1; f += 256) {
        float v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = __ldg(x + 5 * f + k);
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

// scale[u][c] = gamma_c / sqrt(var + eps), shift[u][c] = beta_c - mean * scale
__global__ void __launch_bounds__(512) gn_finalize_kernel(const double* __restrict__ partial, int chunks,
                                                           const UttMeta* __restrict__ meta, const float* __restrict__ w0,
                                                           const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                                                           float* __restrict__ scale, float* __restrict__ shift) {
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
    scale[(int64_t)u * kConvDim + c] = (float)sc;
    shift[(int64_t)u * kConvDim + c] = (float)((double)gn_b[c] - mean * sc);
}

// One block = 64 output frames x 512 channels.  64 threads cover one row (8 channels / 16 B store each), the
// block's 4 thread-rows walk the 64 frames; the 325 input samples of the tile sit in shared memory.
constexpr int kC0Frames = 64;

__global__ void __launch_bounds__(256) conv0_gn_gelu_kernel(const float* __restrict__ wave, const UttMeta* __restrict__ meta,
                                                             const float* __restrict__ w0, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, bf16* __restrict__ out) {
    const int u = blockIdx.y;
    const UttMeta m = meta[u];
    const int slot0 = m.slot6 << 6;
    const int f0 = blockIdx.x * kC0Frames;
    if (f0 >= slot0) return;
    __shared__ float xs[kC0Frames * 5 + 8];
    const float* x = wave + m.sample_off;
    for (int i = threadIdx.x; i < kC0Frames * 5 + 5; i += 256) {
        const int s = f0 * 5 + i;
        xs[i] = s < m.n_samples ? __ldg(x + s) : 0.f;
    }
    const int cg = threadIdx.x & 63;  // channel group: channels 8*cg .. 8*cg+7
    const int fr = threadIdx.x >> 6;  // 0..3
    float w[8][10], sc[8], sh[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 10; ++k) w[c][k] = __ldg(w0 + (cg * 8 + c) * 10 + k);
        sc[c] = scale[(int64_t)u * kConvDim + cg * 8 + c];
        sh[c] = shift[(int64_t)u * kConvDim + cg * 8 + c];
    }
    __syncthreads();
    bf16* orow = out + ((int64_t)m.row6 << 6) * kConvDim;
#pragma unroll 1
    for (int i = fr; i < kC0Frames; i += 4) {
        const int f = f0 + i;
        if (f >= slot0) break;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (f < m.t0) {
            float xv[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) xv[k] = xs[i * 5 + k];
            float y[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 10; ++k) a = fmaf(w[c][k], xv[k], a);
                y[c] = gelu_erf(fmaf(a, sc[c], sh[c]));
            }
            o.x = pack_bf16(y[0], y[1]);
            o.y = pack_bf16(y[2], y[3]);
            o.z = pack_bf16(y[4], y[5]);
            o.w = pack_bf16(y[6], y[7]);
        }
        *reinterpret_cast<uint4*>(orow + (int64_t)f * kConvDim + cg * 8) = o;
    }
}

}  // namespace

int wave_stats_chunks(int max_t0) { return max_t0 <= 0 ? 1 : (max_t0 + kStatFrames - 1) / kStatFrames; }

int launch_wave_stats(const float* wave, const UttMeta* meta, int n_utts, int chunks, const float* w0, const float* gn_w,
                      const float* gn_b, double* partial, float* scale, float* shift, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    wave_moments_kernel<<<dim3(chunks, n_utts), 256, 0, s>>>(wave, meta, partial, chunks);
    gn_finalize_kernel<<<n_utts, 512, 0, s>>>(partial, chunks, meta, w0, gn_w, gn_b, scale, shift);
    return (int)cudaGetLastError();
}

int launch_conv0(const float* wave, const UttMeta* meta, int n_utts, int max_slot0, const float* w0, const float* scale,
                 const float* shift, bf16* out, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    dim3 grid((max_slot0 + kC0Frames - 1) / kC0Frames, n_utts);
    conv0_gn_gelu_kernel<<<grid, 256, 0, s>>>(wave, meta, w0, scale, shift, out);
    return (int)cudaGetLastError();
}

}  // namespace loco
