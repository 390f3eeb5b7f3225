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
// HEAD adds the classifier's other poolings and its Linear as the same epilogue (speech_text/intent_classifier.py):
//   max            :28-30  max over the utterance's frames
//   self_attention :32-36  z_t = x_t . q, alpha = softmax_t(z), sum_t alpha_t x_t   (per-warp online softmax,
//                          the 8 partial (m, l, acc) states merged in warp order)
//   classifier     :20-22, 48  logits = W pooled + b, pooled being the method the head was configured with
template <bool HEAD>
__global__ void __launch_bounds__(256) final_ln_pool_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const UttMeta* __restrict__ meta,
                                                             float* __restrict__ pooled, float* __restrict__ hidden_out,
                                                             HeadArgs head) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.x;
    const UttMeta m = meta[u];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NV = RowVec<kHidden>::kChunks * 8;
    float acc[NV];
    float mx[HEAD ? NV : 1], att[HEAD ? NV : 1], qv[HEAD ? NV : 1];
    float m_run = -INFINITY, l_run = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    if constexpr (HEAD) {
#pragma unroll
        for (int i = 0; i < NV; ++i) mx[i] = -INFINITY, att[i] = 0.f, qv[i] = 0.f;
        if (head.method == kPoolAttention) {
#pragma unroll
            for (int c = 0; c < RowVec<kHidden>::kChunks; ++c)
#pragma unroll
                for (int e = 0; e < 8; ++e) qv[c * 8 + e] = __ldg(head.q + (c * 32 + lane) * 8 + e);
        }
    }
    for (int t = warp; t < m.t6; t += 8) {
        RowVec<kHidden> r;
        r.load(x + (int64_t)(m.row6 + t) * kHidden, lane);
        r.normalize(gamma, beta, lane);
        if (hidden_out != nullptr) r.store_f32(hidden_out + (int64_t)(m.out_row + t) * kHidden, lane);
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += r.v[i];
        if constexpr (HEAD) {
            if (head.method == kPoolMax) {
#pragma unroll
                for (int i = 0; i < NV; ++i) mx[i] = fmaxf(mx[i], r.v[i]);
            } else if (head.method == kPoolAttention) {
                float z = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) z = fmaf(r.v[i], qv[i], z);
                z = warp_sum(z);
                const float m_new = fmaxf(m_run, z);
                const float corr = __expf(m_run - m_new), p = __expf(z - m_new);   // first frame: exp(-inf) = 0
                l_run = fmaf(l_run, corr, p);
#pragma unroll
                for (int i = 0; i < NV; ++i) att[i] = fmaf(att[i], corr, p * r.v[i]);
                m_run = m_new;
            }
        }
    }
    __shared__ float red[8][kHidden];
    __shared__ float head_vec[HEAD ? kHidden : 1];
    __shared__ float warp_m[8], warp_l[8];
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
        if constexpr (HEAD)
            if (head.method == kPoolAverage) head_vec[col] = s * inv;
    }
    if constexpr (HEAD) {
        if (head.method != kPoolAverage) {
            __syncthreads();
            const bool is_max = head.method == kPoolMax;
#pragma unroll
            for (int c = 0; c < RowVec<kHidden>::kChunks; ++c)
#pragma unroll
                for (int e = 0; e < 8; ++e) red[warp][(c * 32 + lane) * 8 + e] = is_max ? mx[c * 8 + e] : att[c * 8 + e];
            if (lane == 0) warp_m[warp] = m_run, warp_l[warp] = l_run;
            __syncthreads();
            float m_all = -INFINITY;
#pragma unroll
            for (int w = 0; w < 8; ++w) m_all = fmaxf(m_all, warp_m[w]);
            float scale[8], l_all = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                scale[w] = warp_m[w] == -INFINITY ? 0.f : __expf(warp_m[w] - m_all);    // warps that saw no frame
                l_all = fmaf(warp_l[w], scale[w], l_all);
            }
            const float inv_l = l_all > 0.f ? 1.0f / l_all : 0.f;
            for (int col = threadIdx.x; col < kHidden; col += 256) {
                float v;
                if (is_max) {
                    v = -INFINITY;
#pragma unroll
                    for (int w = 0; w < 8; ++w) v = fmaxf(v, red[w][col]);
                    if (m.t6 <= 0) v = 0.f;
                } else {
                    v = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) v = fmaf(red[w][col], scale[w], v);
                    v *= inv_l;
                }
                head_vec[col] = v;
            }
        }
        __syncthreads();
        if (head.pooled_out != nullptr)
            for (int col = threadIdx.x; col < kHidden; col += 256) head.pooled_out[(int64_t)u * kHidden + col] = head_vec[col];
        if (head.logits_out != nullptr) {
            for (int c = warp; c < head.n_classes; c += 8) {
                const float* wr = head.w + (int64_t)c * kHidden;
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < kHidden / 128; ++k) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(wr + (k * 32 + lane) * 4));
                    const float4 b = *reinterpret_cast<const float4*>(head_vec + (k * 32 + lane) * 4);
                    d = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, d))));
                }
                d = warp_sum(d);
                if (lane == 0) head.logits_out[(int64_t)u * head.n_classes + c] = d + __ldg(head.b + c);
            }
        }
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
                         float* pooled, float* hidden_out_or_null, const HeadArgs& head, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    if (head.pooled_out != nullptr || head.logits_out != nullptr)
        return launch_pdl(final_ln_pool_kernel<true>, dim3(n_utts), dim3(256), 0, s, x, gamma, beta, meta, pooled, hidden_out_or_null, head);
    return launch_pdl(final_ln_pool_kernel<false>, dim3(n_utts), dim3(256), 0, s, x, gamma, beta, meta, pooled, hidden_out_or_null, head);
}

int launch_row_frames(const UttMeta* meta, int n_utts, int max_slot6, int32_t* row_frame, cudaStream_t s) {
    if (n_utts <= 0) return 0;
    return launch_pdl(row_frames_kernel, dim3((max_slot6 + 127) / 128, n_utts), dim3(128), 0, s, meta, row_frame);
}

}  // namespace loco
