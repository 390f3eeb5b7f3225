/*
 * loco_encode_file.c -- a compiled host of the C ABI (include/loco_asr.h): plain C99 + the CUDA runtime, no Python, no torch.
 *
 * What the reference does in Python at speech_text/extract_speecht5_base_embeddings_slurp.py:98-109 (load the checkpoint,
 * call model.speecht5.encoder(**audios), read last_hidden_state) done from C, twice:
 *   1. host buffers in / out      loco_host_workspace_bytes + loco_encode_host
 *   2. device buffers, async      loco_plan + loco_encode on a stream + loco_sync_check
 * and the two results are required to be bit-identical (exit code 3 otherwise).
 *
 *   loco_encode_file <weights.bin> <waves.bin> <pooled_out.bin>
 *
 * weights.bin   records until EOF: u32 key_len | key bytes | u32 ndim | i64 shape[ndim] | f32 data[prod(shape)]
 *               (HF state-dict keys, the names loco_load_tensor documents)
 * waves.bin     i32 n_utts | i32 n_samples[n_utts] | f32 samples[sum n_samples]   (16 kHz, packed, unpadded)
 * pooled_out    f32 [n_utts, 768]: the mean of last_hidden_state over each utterance's own frames
 *
 * Build (tests/test_c_host.py does exactly this):
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/loco_encode_file.c -o loco_encode_file \
 *       -Lloco_asr_b200 -l:libloco_asr.so -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/loco_asr_b200 -Wl,-rpath,/usr/local/cuda/lib64
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "loco_asr.h"

#define HIDDEN 768

static int die(const char* what, const char* detail, int code) {
    fprintf(stderr, "loco_encode_file: %s: %s\n", what, detail ? detail : "?");
    return code;
}

#define LOCO(call)                                                       \
    do {                                                                 \
        int rc_ = (call);                                                \
        if (rc_ != LOCO_OK) {                                            \
            fprintf(stderr, "loco_encode_file: %s -> %d: %s\n", #call, rc_, loco_last_error(h)); \
            return 2;                                                    \
        }                                                                \
    } while (0)

#define CUDA(call)                                                       \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) return die(#call, cudaGetErrorString(e_), 2); \
    } while (0)

static int load_weights(loco_handle* h, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return die("cannot open", path, 1);
    int n = 0;
    for (;;) {
        uint32_t klen, ndim;
        char key[512];
        int64_t shape[4], numel = 1;
        if (fread(&klen, 4, 1, f) != 1) break; /* EOF */
        if (klen == 0 || klen >= sizeof key || fread(key, 1, klen, f) != klen || fread(&ndim, 4, 1, f) != 1 || ndim > 4 ||
            fread(shape, 8, ndim, f) != ndim) {
            fclose(f);
            return die("malformed record in", path, 1);
        }
        key[klen] = 0;
        for (uint32_t i = 0; i < ndim; ++i) numel *= shape[i];
        float* data = (float*)malloc((size_t)(numel > 0 ? numel : 1) * sizeof(float));
        if (!data || fread(data, sizeof(float), (size_t)numel, f) != (size_t)numel) {
            free(data);
            fclose(f);
            return die("short tensor data for", key, 1);
        }
        int rc = loco_load_tensor(h, key, data, shape, (int)ndim, LOCO_F32); /* the library copies: the buffer is ours again */
        free(data);
        if (rc != LOCO_OK) {
            fprintf(stderr, "loco_encode_file: loco_load_tensor(%s) -> %d: %s\n", key, rc, loco_last_error(h));
            fclose(f);
            return 2;
        }
        ++n;
    }
    fclose(f);
    fprintf(stderr, "loco_encode_file: %d tensors loaded\n", n);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 4) return die("usage", "loco_encode_file <weights.bin> <waves.bin> <pooled_out.bin>", 1);
    if (loco_abi_version() != LOCO_ABI_VERSION) return die("ABI", "header and library disagree", 1);

    loco_config cfg;
    loco_default_config(&cfg);
    loco_handle* h = NULL;
    if (loco_create(&cfg, 0, &h) != LOCO_OK) return die("loco_create", loco_last_error(NULL), 2); /* no GPU: fails here, loudly */
    int rc = load_weights(h, argv[1]);
    if (rc) return rc;
    LOCO(loco_finalize_weights(h));

    FILE* f = fopen(argv[2], "rb");
    if (!f) return die("cannot open", argv[2], 1);
    int32_t n = 0;
    if (fread(&n, 4, 1, f) != 1 || n < 0 || n > 65535) return die("bad header in", argv[2], 1);
    int32_t* n_samples = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (fread(n_samples, 4, (size_t)n, f) != (size_t)n) return die("short length table in", argv[2], 1);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) total += n_samples[i];
    float* wave = (float*)malloc((size_t)(total > 0 ? total : 1) * sizeof(float));
    if (fread(wave, sizeof(float), (size_t)total, f) != (size_t)total) return die("short sample data in", argv[2], 1);
    fclose(f);

    const size_t pooled_bytes = (size_t)n * HIDDEN * sizeof(float);
    float* pooled_a = (float*)calloc((size_t)(n > 0 ? n : 1) * HIDDEN, sizeof(float));
    float* pooled_b = (float*)calloc((size_t)(n > 0 ? n : 1) * HIDDEN, sizeof(float));

    /* ---- 1. host buffers in, host buffers out (the call bench.py's e2e number times) */
    size_t ws_bytes = 0;
    void* ws = NULL;
    LOCO(loco_host_workspace_bytes(h, n_samples, n, /*want_hidden*/ 0, &ws_bytes));
    CUDA(cudaMalloc(&ws, ws_bytes)); /* cudaMalloc promises 256-byte alignment; the library accepts any */
    LOCO(loco_encode_host(h, wave, n_samples, n, pooled_a, NULL, ws, ws_bytes, /*stream*/ NULL));
    CUDA(cudaFree(ws));

    /* ---- 2. device buffers, asynchronous on a stream of ours */
    int32_t* frames = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    int64_t total_frames = 0;
    LOCO(loco_plan(h, n_samples, n, frames, NULL, &total_frames, &ws_bytes));
    cudaStream_t stream;
    float *wave_dev = NULL, *pooled_dev = NULL;
    CUDA(cudaStreamCreate(&stream));
    CUDA(cudaMalloc(&ws, ws_bytes));
    CUDA(cudaMalloc((void**)&wave_dev, (size_t)(total > 0 ? total : 1) * sizeof(float)));
    CUDA(cudaMalloc((void**)&pooled_dev, pooled_bytes ? pooled_bytes : 4));
    CUDA(cudaMemcpyAsync(wave_dev, wave, (size_t)total * sizeof(float), cudaMemcpyHostToDevice, stream));
    LOCO(loco_encode(h, wave_dev, n_samples, n, pooled_dev, NULL, ws, ws_bytes, stream));
    LOCO(loco_sync_check(h, stream)); /* asynchronous errors surface here */
    CUDA(cudaMemcpy(pooled_b, pooled_dev, pooled_bytes, cudaMemcpyDeviceToHost));

    if (memcmp(pooled_a, pooled_b, pooled_bytes) != 0) return die("mismatch", "loco_encode_host and loco_encode disagree", 3);

    f = fopen(argv[3], "wb");
    if (!f || fwrite(pooled_b, 1, pooled_bytes, f) != pooled_bytes) return die("cannot write", argv[3], 1);
    fclose(f);
    fprintf(stderr, "loco_encode_file: %d utterances, %lld frames, %lld kernel launches, both paths bit-identical\n", (int)n,
            (long long)total_frames, (long long)loco_launch_count(h));

    cudaFree(pooled_dev);
    cudaFree(wave_dev);
    cudaFree(ws);
    cudaStreamDestroy(stream);
    loco_destroy(h);
    free(frames);
    free(pooled_a);
    free(pooled_b);
    free(wave);
    free(n_samples);
    return 0;
}
