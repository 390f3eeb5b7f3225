/*
 * loco_asr.h -- C ABI of the B200-native SpeechT5 speech-encoder (LoCo-ASR hot path).
 *
 * The reference has no FFI/plugin interface for this path: the boundary is the Python attribute call
 *     out = model.speecht5.encoder(**audios); out.last_hidden_state
 * (speech_text/extract_speecht5_base_embeddings_slurp.py:98-109,
 *  speech_text/extract_speecht5_finetuned_embeddings_slurp.py:95-105), i.e. HuggingFace
 * SpeechT5EncoderWithSpeechPrenet.forward (transformers modeling_speecht5.py:1355-1374).  The entry points
 * below are what a binding for that call needs; loco_asr_b200/encoder.py binds them with ctypes and mirrors
 * the reference's Python call surface (see INTEGRATION.md for the stub a maintainer adds to the scripts).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary.
 *   - every function returns 0 on success or a negative loco_status; loco_last_error() gives the message.
 *     Nothing throws, nothing calls exit().
 *   - OWNERSHIP: the caller owns every device buffer it passes in (waveforms, outputs, workspace) -- the
 *     Python host allocates them through torch's caching allocator.  The handle owns only its weight copies.
 *   - STREAMS: all work is enqueued on the caller's stream (a cudaStream_t passed as void*).  Host synchronisation happens
 *     only in loco_encode_host (which must hand host memory back), in loco_sync_check and in loco_plan_create.  The FIRST time
 *     loco_encode / loco_encode_text see a batch geometry they build and cache its plan (host work + cudaMalloc) and upload it
 *     with one asynchronous copy on the caller's stream, ordered before the encode -- the device is not synchronised, so a
 *     queue of first-time encodes keeps the GPU busy.  Asynchronous CUDA errors surface at the next call or at loco_sync_check.
 *   - a handle is bound to one device, is not thread-safe; use one handle per rank.
 */
#ifndef LOCO_ASR_H_
#define LOCO_ASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOCO_ABI_VERSION 2

#if defined(__GNUC__)
#define LOCO_API __attribute__((visibility("default")))
#else
#define LOCO_API
#endif

typedef enum {
    LOCO_OK = 0,
    LOCO_ERR_INVALID = -1,     /* bad argument / unsupported config */
    LOCO_ERR_CUDA = -2,        /* a CUDA call failed (message has the CUDA error string) */
    LOCO_ERR_WEIGHTS = -3,     /* missing / misshapen / unknown tensor */
    LOCO_ERR_WORKSPACE = -4,   /* workspace too small */
    LOCO_ERR_STATE = -5        /* call order violated (e.g. encode before finalize) */
} loco_status;

typedef enum { LOCO_F32 = 0, LOCO_F16 = 1, LOCO_BF16 = 2, LOCO_F64 = 3 } loco_dtype;

/* Encoder-relevant fields of HF SpeechT5Config (configuration_speecht5.py:142-194); the reference gets
 * them from the checkpoint's config.json via from_pretrained (extract_speecht5_base_embeddings_slurp.py:98).
 * Only the SpeechT5-base shape family is implemented; loco_create rejects anything else. */
typedef struct {
    int32_t hidden_size;                    /* 768 */
    int32_t encoder_layers;                 /* 12 (any 1..48 accepted) */
    int32_t encoder_attention_heads;        /* 12 */
    int32_t encoder_ffn_dim;                /* 3072 */
    int32_t num_conv_layers;                /* 7 */
    int32_t conv_dim[8];                    /* 512 x7 */
    int32_t conv_kernel[8];                 /* 10,3,3,3,3,2,2 */
    int32_t conv_stride[8];                 /* 5,2,2,2,2,2,2 */
    int32_t num_conv_pos_embeddings;        /* 128 */
    int32_t num_conv_pos_embedding_groups;  /* 16 */
    int32_t max_speech_positions;           /* 4000 (sinusoid table grows on demand, as HF:331-333) */
    int32_t encoder_max_relative_position;  /* 160 */
    int32_t pad_token_id;                   /* 1 */
    int32_t feat_extract_norm_is_group;     /* 1 */
    int32_t activation_is_gelu;             /* 1 (hidden_act and feat_extract_activation) */
    int32_t conv_bias;                      /* 0 */
    float layer_norm_eps;                   /* 1e-5 */
} loco_config;

typedef struct loco_handle loco_handle;

LOCO_API int loco_abi_version(void);

/* Fill `cfg` with the SpeechT5-base defaults above. */
LOCO_API void loco_default_config(loco_config* cfg);

/* Replaces: SpeechT5ForSpeechToText.from_pretrained(...).to(device) -- object construction only
 * (extract_speecht5_base_embeddings_slurp.py:98). */
LOCO_API int loco_create(const loco_config* cfg, int device, loco_handle** out);
LOCO_API void loco_destroy(loco_handle* h);
LOCO_API const char* loco_last_error(const loco_handle* h); /* h may be NULL: returns the last create() error */

/* Replaces: load_state_dict on encoder.wrapped_encoder / encoder.prenet
 * (extract_speecht5_base_embeddings_slurp.py:99-100) and the key names map_speecht5_hf.py:34-168 produces.
 * `key` is an HF state-dict key; accepted spellings: with or without the "speecht5.encoder." / "encoder."
 * prefix; sub-module dicts must be prefixed by the caller with "prenet." / "wrapped_encoder."; weight-norm
 * as "...pos_conv_embed.conv.weight_g|weight_v" (transformers 4.30.2) or
 * "...conv.parametrizations.weight.original0|original1" (5.x).  "prenet.masked_spec_embed" and
 * "prenet.pos_sinusoidal_embed.weights" are accepted and ignored (unused in eval / regenerated).
 * `data` is HOST memory, contiguous, row-major, `shape[ndim]`. */
LOCO_API int loco_load_tensor(loco_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype);

/* Checks that every tensor arrived, folds weight-norm, scales q_proj by head_dim^-0.5 * log2(e), fuses q/k/v,
 * re-lays conv weights tap-major, casts GEMM operands to bf16 and uploads. */
LOCO_API int loco_finalize_weights(loco_handle* h);

/* Geometry of one batch.  n_samples[n_utts] (host).  Outputs (any may be NULL):
 *   frames[n_utts]   encoder frames T_u per utterance (HF _get_feat_extract_output_lengths, :585-598)
 *   rows[n_utts]     first row of utterance u in the slot-packed [R6, *] stage buffers (debug taps)
 *   total_frames     sum of T_u  (rows of the compact hidden_out)
 *   workspace_bytes  device scratch loco_encode needs for this batch.  The workspace pointer may have ANY alignment
 *                    (cudaMalloc: 256 B, torch's caching allocator: 512 B): the figure includes 1024 bytes of slack and
 *                    loco_encode / loco_encode_text / loco_encode_host round the base up to the 1024-byte boundary the
 *                    TMA boxes and swizzled tiles inside want.
 * At most 65535 utterances per call (LOCO_ERR_INVALID beyond); split larger sets into several calls. */
LOCO_API int loco_plan(loco_handle* h, const int32_t* n_samples, int n_utts, int32_t* frames, int32_t* rows,
              int64_t* total_frames, size_t* workspace_bytes);

/* Replaces: out = model.speecht5.encoder(**audios)  (extract_speecht5_base_embeddings_slurp.py:108).
 *   wave_dev      f32[sum n_samples]  packed, unpadded 16 kHz waveforms (device)
 *   n_samples     i32[n_utts]         (host)
 *   pooled_dev    f32[n_utts, 768]    mean of last_hidden_state over each utterance's own frames (device)
 *   hidden_dev    f32[total_frames, 768] or NULL: last_hidden_state, utterances concatenated (device)
 *   workspace_dev any alignment, at least loco_plan's workspace_bytes (LOCO_ERR_WORKSPACE otherwise); wave_dev 4-byte aligned
 * Asynchronous on `stream`. */
LOCO_API int loco_encode(loco_handle* h, const float* wave_dev, const int32_t* n_samples, int n_utts, float* pooled_dev,
                float* hidden_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- plans: a batch geometry made ready to launch --------------------------------------------------------------------
 * loco_plan_create computes the layout and the kernels' work lists for one batch (kind 0: speech, lengths = samples per
 * utterance; kind 1: text, lengths = tokens per text) and uploads them into a small device block owned by the plan.  It is
 * synchronous (cudaMalloc + cudaMemcpy): call it outside hot loops and outside stream captures.
 * loco_encode_planned is then a pure enqueue on `stream` -- kernels and memset nodes only, no host-to-device copy, no host
 * synchronisation -- so it may be captured into a CUDA graph: capture one call per bucket, then replay the graph after
 * refilling the same input buffer (same lengths).  The caller still owns input, outputs and workspace (loco_plan_info's
 * workspace_bytes; any alignment).  A plan may be used any number of times, from one handle, until loco_plan_destroy.
 * loco_encode / loco_encode_text are loco_plan_create (cached per geometry, up to 256 of them; uploaded asynchronously on the
 * encoding stream instead of synchronously) + loco_encode_planned. */
typedef struct loco_batch_plan loco_batch_plan;
LOCO_API int loco_plan_create(loco_handle* h, int kind, const int32_t* lengths, int n_utts, loco_batch_plan** out);
LOCO_API int loco_plan_info(const loco_batch_plan* plan, int32_t* frames, int32_t* rows, int64_t* total_frames, size_t* workspace_bytes);
LOCO_API int loco_encode_planned(loco_handle* h, const loco_batch_plan* plan, const void* input_dev /* f32 waveforms | i32 tokens */,
                                 float* pooled_dev, float* hidden_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
LOCO_API void loco_plan_destroy(loco_handle* h, loco_batch_plan* plan);

/* Synchronise `stream` and report any asynchronous CUDA error of the work enqueued so far (LOCO_ERR_CUDA + message). */
LOCO_API int loco_sync_check(loco_handle* h, void* stream);

/* Same call with HOST buffers: copies the waveforms H2D (staged in the tail of the workspace), encodes,
 * copies pooled (and hidden) D2H and synchronises `stream` before returning.  The workspace must be
 * loco_host_workspace_bytes() large.  This is the path bench.py's `e2e` number times. */
LOCO_API int loco_host_workspace_bytes(loco_handle* h, const int32_t* n_samples, int n_utts, int want_hidden, size_t* bytes);
LOCO_API int loco_encode_host(loco_handle* h, const float* wave_host, const int32_t* n_samples, int n_utts, float* pooled_host,
                     float* hidden_host, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- text modality (the `-m text` branch of the same scripts) ------------------------------------------------
 * Replaces: out = model.speecht5.encoder(texts.input_ids) on SpeechT5ForTextToSpeech's encoder
 * (extract_speecht5_base_embeddings_slurp.py:79-93) = SpeechT5EncoderWithTextPrenet (HF modeling_speecht5.py:1377-1415):
 * SpeechT5TextEncoderPrenet (embed_tokens + scaled positional encoding, HF:765-779, 400-422) then the same SpeechT5Encoder.
 * A handle serves this path when its state dict carried "prenet.embed_tokens.weight" [vocab, 768] and
 * "prenet.encode_positions.alpha" (map_speecht5_hf.py:170-181; "prenet.encode_positions.pe" is accepted and ignored);
 * speech and text prenet weights may be loaded into the same handle, the wrapped_encoder weights are shared.
 *   n_tokens       i32[n_utts] (host)  tokens per text; each text is encoded alone, unpadded
 *   tokens_dev     i32[sum n_tokens]   packed token ids (device); ids outside [0, vocab) are clamped
 *   pooled_dev     f32[n_utts, 768]    mean of last_hidden_state over each text's own tokens
 *   hidden_dev     f32[sum n_tokens, 768] or NULL
 * loco_plan_text: rows[n_utts] = first row of text u in the stage buffers (== its offset in tokens_dev). */
LOCO_API int loco_plan_text(loco_handle* h, const int32_t* n_tokens, int n_utts, int32_t* rows, int64_t* total_tokens,
                            size_t* workspace_bytes);
LOCO_API int loco_encode_text(loco_handle* h, const int32_t* tokens_dev, const int32_t* n_tokens, int n_utts, float* pooled_dev,
                              float* hidden_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- classifier head as the encoder's epilogue (the consumer of the embeddings) -------------------------------
 * Replaces IntentClassifier.forward (speech_text/intent_classifier.py:38-50) on one utterance's own frames: the pooling
 * the classifier was built with -- method 0 "average" (:24-26), 1 "max" (:28-30), 2 "self_attention" (:32-36:
 * alpha = softmax_t(x_t . q), sum_t alpha_t x_t) -- and classifier = Linear(768, n_classes) (:20-22), computed inside the
 * kernel that applies the last LayerNorm, so last_hidden_state never goes to HBM for it.
 *   q_host  f32[768]            IntentClassifier.q        (required for method 2, else may be NULL)
 *   w_host  f32[n_classes,768]  classifier.0.weight       (NULL: pooling only)
 *   b_host  f32[n_classes]      classifier.0.bias
 * loco_set_head copies the HOST arrays to the device (synchronous; may be called again to replace the head, which also
 * clears the output pointers).
 * loco_set_head_outputs names where the following loco_encode / loco_encode_text calls write, until changed:
 *   head_pooled_dev  f32[n_utts,768]        the method's pooled vector, or NULL
 *   logits_dev       f32[n_utts,n_classes]  classifier output, or NULL        (both NULL: head off, the default)
 * pooled_dev of the encode call stays the masked mean whatever the method. */
#define LOCO_POOL_AVERAGE 0
#define LOCO_POOL_MAX 1
#define LOCO_POOL_SELF_ATTENTION 2
LOCO_API int loco_set_head(loco_handle* h, int method, const float* q_host, const float* w_host, const float* b_host, int n_classes);
LOCO_API int loco_set_head_outputs(loco_handle* h, float* head_pooled_dev, float* logits_dev);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
LOCO_API int64_t loco_launch_count(const loco_handle* h);

/* Per-stage device timing for the roofline report: when enabled every launch of loco_encode is bracketed by a
 * CUDA event pair on the caller's stream.  loco_profile_collect synchronises the device, sums elapsed
 * milliseconds and launch counts per category -- 0 tcgen05 GEMMs, 1 attention, 2 positional conv, 3 conv0 +
 * GroupNorm statistics, 4 LayerNorm / pooling row kernels -- and resets the records. */
LOCO_API int loco_profile_enable(loco_handle* h, int on);
LOCO_API int loco_profile_collect(loco_handle* h, int n_cats, double* ms, int64_t* launches);

/* ---- debug / test hooks (not part of the product surface) -------------------------------------------
 * The product library (libloco_asr.so) contains only the product kernels; loco_debug_set fails on it.  The cross-check
 * kernels (SIMT GEMM, single-CTA tcgen05 GEMM, mma.sync conv0, mma.sync and one-phase tcgen05 positional conv, mma.sync attention, stand-alone LayerNorm path) and the
 * knobs that select them are compiled only with -DLOCO_DEBUG into libloco_asr_debug.so, which the unit tests load. */
LOCO_API int loco_is_debug_build(void);
/* name: "gemm_impl" (2 = tcgen05 CTA pair, cta_group::2 [default], 0 = tcgen05 single CTA, 1 = SIMT reference), "posconv_impl" (0 = polyphase tcgen05 [default],
 * 1 = mma.sync cross-check, 2 = one-phase tcgen05 cross-check), "conv0_impl" (0 = tcgen05 [default], 1 = mma.sync cross-check), "ln_impl" (0 = the transformer layers' LayerNorms deferred into the GEMM epilogues [default, needs gemm_impl 2],
 * 1 = LayerNorm kernels), "attn_impl" (-1 = per utterance by its own frame count [default], 0 = tcgen05, 1 = mma.sync), "attn_tc_min_frames" / "attn_tc_lo" / "attn_tc_hi" (the frame ranges that select the tcgen05 kernel), "stop_after_layer"
 * (-1 = run all).  */
LOCO_API int loco_debug_set(loco_handle* h, const char* name, int64_t value);
/* After an encode: device pointer / geometry of a named stage buffer inside the caller's workspace
 * ("conv0".."conv6", "proj_ln", "proj", "pos_conv", "enc_in", "qkv", "ctx", "attn_res", "ln1", "mid",
 * "ffn_res", "x").  Row r of utterance u's frame t is rows[u] * 2^(6-i) + t for conv stage i, rows[u] + t
 * for all others.  dtype is a loco_dtype. */
LOCO_API int loco_debug_buffer(loco_handle* h, const char* name, void** dev_ptr, int64_t* n_rows, int64_t* n_cols, int* dtype);
/* Stand-alone GEMM entry for unit tests: C[M,N] = epi(A[M,K] (row stride lda) * W[N,K]^T (+bias) (+R)). */
LOCO_API int loco_debug_gemm(loco_handle* h, int impl, const void* a_bf16, int64_t lda, int64_t a_rows_alloc, const void* w_bf16,
                    void* c_bf16, const float* bias, const void* r_bf16, int m, int n, int k, int epilogue, void* stream);

/* The deferred-LayerNorm epilogues of the CTA-pair GEMM (epilogue 3..6, see csrc/internal.h GemmEpilogue), N = 768 where
 * statistics are produced.  stats_in / stats_out: f32[M, 6, 2] = (mean, M2) of each 128-column slice of a 768-wide row;
 * c1: f32[N] column sums of W; gamma: f32[N].  Unused operands may be NULL. */
LOCO_API int loco_debug_gemm_ln(loco_handle* h, const void* a_bf16, const void* w_bf16, void* c_bf16, const float* bias,
                                const void* r_bf16, int m, int n, int k, int epilogue, const float* stats_in, const float* c1,
                                const float* gamma, float* stats_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOCO_ASR_H_ */
