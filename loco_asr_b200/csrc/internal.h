// Host-side declarations shared by the kernel translation units and api.cu.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace loco {

typedef __nv_bfloat16 bf16;

enum GemmEpilogue : int {
    EPI_BIAS = 0,           // C = A W^T (+ bias)
    EPI_BIAS_GELU = 1,      // C = gelu(A W^T (+ bias))
    EPI_BIAS_RESIDUAL = 2,  // C = A W^T + bias + R
    // Deferred LayerNorm (CTA-pair kernel only; N = 768 producers, any N consumers).  A post-LN block x' = LN(u), u = x + f(x),
    // never materialises x': the producer writes u and per-row partial statistics, the consumers normalise on the fly.
    EPI_LN_BIAS = 3,                 // C = rstd_r (A W'^T - mean_r c1[n]) + bias[n]        A = u (un-normalised), W' = gamma (.) W,
    EPI_LN_BIAS_GELU = 4,            // C = gelu(same)                                      c1[n] = sum_k W'[n,k], bias = beta W^T + b
    EPI_BIAS_RESIDUAL_STATS = 5,     // C = A W^T + bias + R, and the row statistics of C
    EPI_BIAS_LNRESIDUAL_STATS = 6,   // C = A W^T + bias' + (R - mean_r) rstd_r gamma[n], bias' = bias + beta, and the row statistics of C
};
constexpr int kStatSlots = 6;        // row statistics of an [M, 768] tensor: (mean, M2) of each 128-column slice, kStatSlots * 2 floats per row

// One GEMM problem: C[M, N] (bf16, row stride ldc) = epi(A[M, K] * W[N, K]^T).
// A rows may overlap in memory (lda < K) -- that is how the strided convolutions are expressed
// (row t of the implicit-GEMM operand is the contiguous strip of k*512 inputs starting at frame 2t).
struct GemmArgs {
    const bf16* A;
    int64_t lda;          // elements between consecutive A rows
    int64_t a_rows_alloc; // rows of A that exist in memory (TMA bound; >= M)
    const bf16* W;        // [N, K] row-major (nn.Linear layout)
    bf16* C;
    int64_t ldc;
    const float* bias;    // [N] or nullptr
    const bf16* R;        // residual [M, N] with row stride ldr, or nullptr
    int64_t ldr;
    int M, N, K;
    int epilogue;
    // deferred-LayerNorm epilogues
    const float* stats_in = nullptr;   // [M, kStatSlots, 2] statistics of the rows of A (EPI_LN_*) or of R (EPI_BIAS_LNRESIDUAL_STATS)
    const float* c1 = nullptr;         // [N] column sums of W' (EPI_LN_*)
    const float* ln_gamma = nullptr;   // [N] scale of the LayerNorm applied to R (EPI_BIAS_LNRESIDUAL_STATS; its shift is folded into bias)
    float* stats_out = nullptr;        // [M, kStatSlots, 2] statistics of the rows of C (EPI_*_STATS; N must be 768)
};

// Launch helper.  With LOCO_PDL=1 kernels are launched with programmatic stream serialization (PDL): a kernel may be
// scheduled while its predecessor in the stream drains and blocks in pdl_wait() (griddepcontrol.wait) until the predecessor
// has completed.  OFF by default: on the SLURP-shaped bench it measured 1.7 % slower (the persistent kernels leave no SM
// resources for an early successor, so only the wait's own latency is added).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Host helpers (tensormap.cu).  Returns cudaError / CUresult as int (0 = ok).
int tensormap_init();  // resolves cuTensorMapEncodeTiled through the runtime
// 2-D bf16 tensor map, box = [64 elements (128 B), box_rows], SWIZZLE_128B.
int make_tensor_map_bf16_sw128(void* map /*CUtensorMap*/, const void* base, uint64_t inner, uint64_t rows,
                               uint64_t row_stride_elems, uint32_t box_rows);
// Single-CTA tcgen05/TMEM/TMA GEMM (gemm_tcgen05.cu): the CTA-pair kernel's predecessor, LOCO_DEBUG builds only (cross-check).
int gemm_tc_init();
int gemm_tc_launch(const GemmArgs& g, int num_sms, cudaStream_t stream);
// CTA-pair variant (gemm_tcgen05_2cta.cu): tcgen05.mma.cta_group::2, 256 x 256 tiles, each CTA loads half of the weight tile.
int gemm_tc2_init();
int gemm_tc2_launch(const GemmArgs& g, int num_sms, cudaStream_t stream);
// Debug-only SIMT reference GEMM (gemm_simt.cu): used by the tests to cross-check the tensor path.
int gemm_simt_launch(const GemmArgs& g, cudaStream_t stream);

// ---- front end (frontend.cu) -------------------------------------------------------------------
struct UttMeta {          // per-utterance geometry, device-resident
    int64_t sample_off;   // first sample in the packed waveform
    int32_t n_samples;
    int32_t t0;           // frames after conv layer 0
    int32_t t6;           // frames after conv layer 6 == encoder frames T
    int32_t row6;         // first row in every [R6, *] buffer; row in layer i buffers = row6 << (6 - i)
    int32_t slot6;        // rows reserved in [R6, *] buffers (>= t6); layer i reserves slot6 << (6 - i)
    int32_t out_row;      // first row in the compact hidden_out buffer (prefix sum of t6)
};

int wave_stats_chunks(int max_t0);
int launch_wave_stats(const float* wave, const UttMeta* meta, int n_utts, int chunks, const float* w0 /*[512,10]*/,
                      const float* gn_w, const float* gn_b, double* partial /*[n, chunks, 65]*/, float* scale /*[n,512]*/,
                      float* shift /*[n,512]*/, bf16* wfold /*[n, kConv0FoldBytes] or null*/, cudaStream_t s);
// conv0 + GroupNorm + GELU on tcgen05 (conv0_tc.cu; the product kernel).  wfold: per utterance the K = 48 split-GEMM B operand
// gn_finalize_kernel wrote; tile_start[u] = first 128-frame tile of utterance u in the batch's tile order (n_utts + 1 entries).
constexpr int kConv0FoldBytes = 6 * 512 * 16;
int conv0_tc_init();
int launch_conv0_tc(const float* wave, const UttMeta* meta, const int32_t* tile_start, int n_utts, int n_tiles, const bf16* wfold,
                    bf16* out /*[out_rows, 512]*/, int64_t out_rows, int num_sms, cudaStream_t s);
// mma.sync version (conv0_mma.cu; debug cross-check)
int conv0_mma_init();
int launch_conv0(const float* wave, const UttMeta* meta, int n_utts, int max_slot0, const float* w0, const float* scale,
                 const float* shift, bf16* out /*[R0, 512]*/, cudaStream_t s);

// ---- row-wise kernels (rowops.cu) --------------------------------------------------------------
// y = LayerNorm(x) over `cols` (512 or 768), bf16 in / bf16 out.
int launch_layernorm(const bf16* x, bf16* y, const float* gamma, const float* beta, int rows, int cols, cudaStream_t s);
// y = LayerNorm(h + pc + sinusoid(frame + 2)) : end of the prenet + encoder input LayerNorm; slot padding rows -> 0.
int launch_prenet_ln(const bf16* h, const bf16* pc, const float* sin_table, const int32_t* row_frame, bf16* y,
                     const float* gamma, const float* beta, int rows, cudaStream_t s);
// Classifier head fused into the final kernel (speech_text/intent_classifier.py:24-48); all pointers device, fp32.
enum { kPoolAverage = 0, kPoolMax = 1, kPoolAttention = 2 };
struct HeadArgs {
    int method = kPoolAverage;
    int n_classes = 0;
    const float* q = nullptr;       // [768]            (self_attention only)
    const float* w = nullptr;       // [n_classes, 768] classifier.0.weight
    const float* b = nullptr;       // [n_classes]      classifier.0.bias
    float* pooled_out = nullptr;    // [n_utts, 768] or null: the method's pooled vector
    float* logits_out = nullptr;    // [n_utts, n_classes] or null
};
// Final LayerNorm of the last layer fused with the masked mean-pool, the optional compact fp32 copy and, when
// `head` carries an output pointer, the classifier's pooling + Linear.
int launch_final_ln_pool(const bf16* x, const float* gamma, const float* beta, const UttMeta* meta, int n_utts,
                         float* pooled /*[n,768]*/, float* hidden_out_or_null, const HeadArgs& head, cudaStream_t s);
// row_frame[r] = frame index of row r inside its utterance, or -1 for slot padding rows.
int launch_row_frames(const UttMeta* meta, int n_utts, int max_slot6, int32_t* row_frame, cudaStream_t s);
// text modality: y[row] = LayerNorm(embed[tokens[row]] + alpha * pe[row_frame[row]])
int launch_text_prenet_ln(const int32_t* tokens, const float* embed /*[vocab, 768]*/, const float* pe /*[positions, 768]*/, float alpha,
                          int vocab, const int32_t* row_frame, bf16* y, const float* gamma, const float* beta, int rows, cudaStream_t s);

// ---- 128- / 64-frame tiles of one utterance: attention work lists and the one-phase positional conv (posconv_tc.cu) ----
struct PcTile {      // one 128-frame output tile of one utterance
    int32_t row0;    // row of the tile's first frame in the [R6, 768] buffers
    int32_t f0;      // that frame's index inside its utterance
    int32_t t6;      // the utterance's frame count
    int32_t pad_;
};
// ---- positional conv on tcgen05, polyphase (posconv_pp.cu; the product kernel) -------------------
// The batch on a virtual timeline: utterance after utterance, kPosPPHalo zero frames between neighbours; an item is
// kPosPPTile consecutive timeline frames of one group.  vmap[kPosPPHalo + v] = row of timeline frame v in the [R6, 768]
// buffers, or -1; it has n_vtiles * kPosPPTile + 2 * kPosPPHalo entries.
constexpr int kPosPPTile = 512;
constexpr int kPosPPHalo = 64;
constexpr int kPosPPTaps = 139;      // taps per (group, channel chunk) in w_pp: 3 zero taps, the 128 taps, 8 zero taps
int posconv_pp_init();
// w_pp: [16 groups][6 in-chunks][kPosPPTaps][48 out][8 in] bf16
int launch_posconv_pp(const bf16* h, const bf16* w_pp, const float* bias, const int32_t* vmap, int n_vtiles, bf16* pc, int num_sms,
                      cudaStream_t s);

// ---- positional conv on tcgen05, one output frame per accumulator row (posconv_tc.cu; debug cross-check) ----
int posconv_tc_init();
// w_tc: [16 groups][128 taps][6 in-chunks][48 out][8 in] bf16 -- one tap's block is a ready-made UMMA B operand.
int launch_posconv_tc(const bf16* h, const bf16* w_tc, const float* bias, const PcTile* tiles, int n_tiles, bf16* pc,
                      cudaStream_t s);

// ---- positional conv, mma.sync debug cross-check (posconv.cu) ----------------------------------
// pc[r, :] = gelu(bias + grouped_conv(h)[r, :]) with zero padding at each utterance's own boundaries.
int posconv_init();
int launch_posconv(const bf16* h, const bf16* w /*[16][128][48 out][48 in]*/, const float* bias, const UttMeta* meta,
                   int n_utts, int max_t6, bf16* pc, cudaStream_t s);

// ---- attention (attention.cu) ------------------------------------------------------------------
// ctx[r, h*64:(h+1)*64] = softmax_j(q_i.k_j + q_i.pe_k[clip(i-j)+160]) v_j within each utterance.
int attention_init();
// tcgen05/TMEM/TMA kernel (attention_tc.cu), persistent over (128-query tile, head) items.  maps: qkv [2304, R6] box 32 rows;
// pe_k [64, 320] box 160 rows.  `tiles` is the positional conv's tile list (same 128-frame tiling of every utterance).
int attention_tc_init();
int launch_attention_tc(const void* qkv_map /*CUtensorMap*/, const void* pe_map /*CUtensorMap*/, const PcTile* tiles, int n_tiles,
                        bf16* ctx /*[R6, 768]*/, int num_sms, cudaStream_t s);
// Two-pipeline variant (attention_p2.cu): two (64-query tile, head) items per SM at a time, each in its own half of the TMEM
// lanes.  `tiles`: 64-frame tiles of every utterance.
int attention_p2_init();
int launch_attention_p2(const void* qkv_map /*CUtensorMap*/, const void* pe_map /*CUtensorMap*/, const PcTile* tiles, int n_tiles,
                        bf16* ctx /*[R6, 768]*/, int num_sms, cudaStream_t s);
// utt_index: the utterances to process (indices into meta), or nullptr for all of 0..n_utts-1
int launch_attention(const bf16* qkv /*[R6, 2304]*/, const bf16* pe_k /*[320, 64]*/, const UttMeta* meta, const int32_t* utt_index,
                     int n_utts, int max_t6, bf16* ctx /*[R6, 768]*/, cudaStream_t s);

}  // namespace loco
