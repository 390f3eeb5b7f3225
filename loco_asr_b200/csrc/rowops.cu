// Row-wise, HBM-bound kernels: LayerNorm variants (one warp per row, 128-bit loads, warp-shuffle statistics
// in fp32), the prenet tail (positional-conv residual + sinusoidal positions + encoder input LayerNorm) and
// the final LayerNorm fused with the masked mean-pool.
//   a4  feature_projection.layer_norm            HF modeling_speecht5.py:498-510
//   a8  sinusoidal positions (row = frame + 2)   HF:285-351, 558-564
//   a10 encoder input LayerNorm                  HF:1292
//   a14 post-LN blocks                           HF:1047-1060
//   a17 mean over the utterance's own frames     intent_classifier.py:24-26 (masked: no padded frames exist)
#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {

template <int COLS>
struct RowVec {
    static constexpr int kChunks = COLS / 256;  // 16-byte chunks (8 bf16) per lane
    float v[kChunks * 8];

    __device__ __forceinline__ void load(const bf16* row, int lane) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + (c * 32 + lane) * 8));
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
            float* p = v + c * 8;
            p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y; p[4] = d.x; p[5] = d.y; p[6] = e.x; p[7] = e.y;
        }
    }
    __device__ __forceinline__ void add(const bf16* row, int lane) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + (c * 32 + lane) * 8));
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
            float* p = v + c * 8;
            p[0] += a.x; p[1] += a.y; p[2] += b.x; p[3] += b.y; p[4] += d.x; p[5] += d.y; p[6] += e.x; p[7] += e.y;
        }
    }
    __device__ __forceinline__ void add_f32(const float* row, int lane) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(row + (c * 32 + lane) * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(row + (c * 32 + lane) * 8 + 4));
            float* p = v + c * 8;
            p[0] += a.x; p[1] += a.y; p[2] += a.z; p[3] += a.w; p[4] += b.x; p[5] += b.y; p[6] += b.z; p[7] += b.w;
        }
    }
    // in-place LayerNorm with affine; statistics in fp32, two-pass over registers
    __device__ __forceinline__ void normalize(const float* __restrict__ gamma, const float* __restrict__ beta, int lane) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kChunks * 8; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.0f / COLS);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kChunks * 8; ++i) {
            const float d = v[i] - mean;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / COLS) + kLnEps);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int col = (c * 32 + lane) * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
            float* p = v + c * 8;
            p[0] = fmaf((p[0] - mean) * rstd, g0.x, b0.x);
            p[1] = fmaf((p[1] - mean) * rstd, g0.y, b0.y);
            p[2] = fmaf((p[2] - mean) * rstd, g0.z, b0.z);
            p[3] = fmaf((p[3] - mean) * rstd, g0.w, b0.w);
            p[4] = fmaf((p[4] - mean) * rstd, g1.x, b1.x);
            p[5] = fmaf((p[5] - mean) * rstd, g1.y, b1.y);
            p[6] = fmaf((p[6] - mean) * rstd, g1.z, b1.z);
            p[7] = fmaf((p[7] - mean) * rstd, g1.w, b1.w);
        }
    }
    __device__ __forceinline__ void store(bf16* row, int lane) const {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const float* p = v + c * 8;
            uint4 o;
            o.x = pack_bf16(p[0], p[1]);
            o.y = pack_bf16(p[2], p[3]);
            o.z = pack_bf16(p[4], p[5]);
            o.w = pack_bf16(p[6], p[7]);
            *reinterpret_cast<uint4*>(row + (c * 32 + lane) * 8) = o;
        }
    }
    __device__ __forceinline__ void store_f32(float* row, int lane) const {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const float* p = v + c * 8;
            *reinterpret_cast<float4*>(row + (c * 32 + lane) * 8) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4*>(row + (c * 32 + lane) * 8 + 4) = make_float4(p[4], p[5], p[6], p[7]);
        }
    }
};

template <int COLS>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         int rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    RowVec<COLS> r;
    r.load(x + (int64_t)row * COLS, lane);
    r.normalize(gamma, beta, lane);
    r.store(y + (int64_t)row * COLS, lane);
}

__global__ void __launch_bounds__(256) prenet_ln_kernel(const bf16* __restrict__ h, const bf16* __restrict__ pc,
                                                         const float* __restrict__ sin_table, const int32_t* __restrict__ row_frame,
                                                         bf16* __restrict__ y, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int frame = row_frame[row];
    RowVec<kHidden> r;
    if (frame < 0) {  // slot padding row: keep it finite and deterministic
#pragma unroll
        for (int i = 0; i < RowVec<kHidden>::kChunks * 8; ++i) r.v[i] = 0.f;
        r.store(y + (int64_t)row * kHidden, lane);
        return;
    }
    r.load(h + (int64_t)row * kHidden, lane);
    r.add(pc + (int64_t)row * kHidden, lane);
    r.add_f32(sin_table + (int64_t)(frame + 2) * kHidden, lane);  // position = frame + padding_idx + 1
    r.normalize(gamma, beta, lane);
    r.store(y + (int64_t)row * kHidden, lane);
}

// Text prenet + encoder input LayerNorm (SpeechT5TextEncoderPrenet, HF modeling_speecht5.py:765-779: embed_tokens then
// SpeechT5ScaledPositionalEncoding, HF:400-422: emb + alpha * pe[position]; SpeechT5Encoder.layer_norm, HF:1292).
// In the text layout a row IS a token (no slot padding), so tokens[row] is the row's id and row_frame[row] its position.
__global__ void __launch_bounds__(256) text_prenet_ln_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ embed,
                                                              const float* __restrict__ pe, float alpha, int vocab,
                                                              const int32_t* __restrict__ row_frame, bf16* __restrict__ y,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, int rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int pos = row_frame[row];
    RowVec<kHidden> r;
#pragma unroll
    for (int i = 0; i < RowVec<kHidden>::kChunks * 8; ++i) r.v[i] = 0.f;
    if (pos < 0) {
        r.store(y + (int64_t)row * kHidden, lane);
        return;
    }
    const int tok = min(max(tokens[row], 0), vocab - 1);      // ids are validated on the host side of the ABI
    r.add_f32(embed + (int64_t)tok * kHidden, lane);
    RowVec<kHidden> p;
#pragma unroll
    for (int i = 0; i < RowVec<kHidden>::kChunks * 8; ++i) p.v[i] = 0.f;
    p.add_f32(pe + (int64_t)pos * kHidden, lane);
#pragma unroll
    for (int i = 0; i < RowVec<kHidden>::kChunks * 8; ++i) r.v[i] = fmaf(alpha, p.v[i], r.v[i]);
    r.normalize(gamma, beta, lane);
    r.store(y + (int64_t)row * kHidden, lane);
}

// One block per utterance: LayerNorm every valid frame, accumulate column sums per warp in registers,
// reduce across the 8 warps in a fixed order (deterministic), write mean over the utterance's T frames.
__global__ void __launch_bounds__(256) final_ln_pool_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const UttMeta* __restrict__ meta,
                                                             float* __restrict__ pooled, float* __restrict__ hidden_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.x;
    const UttMeta m = meta[u];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NV = RowVec<kHidden>::kChunks * 8;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (int t = warp; t < m.t6; t += 8) {
        RowVec<kHidden> r;
        r.load(x + (int64_t)(m.row6 + t) * kHidden, lane);
        r.normalize(gamma, beta, lane);
        if (hidden_out != nullptr) r.store_f32(hidden_out + (int64_t)(m.out_row + t) * kHidden, lane);
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += r.v[i];
    }
    __shared__ float red[8][kHidden];
#pragma unroll
    for (int c = 0; c < RowVec<kHidden>::kChunks; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) red[warp][(c * 32 + lane) * 8 + e] = acc[c * 8 + e];
    __syncthreads();
    const float inv = m.t6 > 0 ? 1.0f / (float)m.t6 : 0.f;
    for (int col = threadIdx.x; col < kHidden; col += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][col];
        pooled[(int64_t)u * kHidden + col] = s * inv;
    }
}

__global__ void row_frames_kernel(const UttMeta* __restrict__ meta, int32_t* __restrict__ row_frame) {
    pdl_launch_dependents();
    pdl_wait();
    const UttMeta m = meta[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m.slot6) row_frame[m.row6 + t] = t < m.t6 ? t : -1;
}

}  // namespace

int launch_layernorm(const bf16* x, bf16* y, const float* gamma, const float* beta, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return 0;
    const int grid = (rows + 7) / 8;
    if (cols == kHidden) return launch_pdl(layernorm_kernel<kHidden>, dim3(grid), dim3(256), 0, s, x, y, gamma, beta, rows);
    else if (cols == kConvDim) return launch_pdl(layernorm_kernel<kConvDim>, dim3(grid), dim3(256), 0, s, x, y, gamma, beta, rows);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}

int launch_prenet_ln(const bf16* h, const bf16* pc, const float* sin_table, const int32_t* row_frame, bf16* y,
                     const float* gamma, const float* beta, int rows, cudaStream_t s) {
    if (rows <= 0) return 0;
    return launch_pdl(prenet_ln_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, h, pc, sin_table, row_frame, y, gamma, beta, rows);
}

int launch_text_prenet_ln(const int32_t* tokens, const float* embed, const float* pe, float alpha, int vocab, const int32_t* row_frame,
                          bf16* y, const float* gamma, const float* beta, int rows, cudaStream_t s) {
    if (rows <= 0) return 0;
    return launch_pdl(text_prenet_ln_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, tokens, embed, pe, alpha, vocab, row_frame, y, gamma, beta, rows);
}

int launch_final_ln_pool(const bf16* x, const float* gamma, const float* beta, const UttMeta* meta, int n_utts,
                         float* pooled, float* hidden_out_or_null, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    return launch_pdl(final_ln_pool_kernel, dim3(n_utts), dim3(256), 0, s, x, gamma, beta, meta, pooled, hidden_out_or_null);
}

int launch_row_frames(const UttMeta* meta, int n_utts, int max_slot6, int32_t* row_frame, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    return launch_pdl(row_frames_kernel, dim3((max_slot6 + 127) / 128, n_utts), dim3(128), 0, s, meta, row_frame);
}

}  // namespace loco
